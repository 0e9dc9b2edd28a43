// tools/microbench.cu -- integer-pipe micro-benchmarks that size the 64-bit modular multiply on sm_100a.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/microbench tools/microbench.cu
// Prints per-op throughput in (thread-)ops per clock per SM, from CUDA-event time and the measured SM clock.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define ITER 512
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) k(u64 *out, u64 seed, u64 q, u64 qinv, long long *clk) {
    u64 x[CHAINS], w[CHAINS];
    for (int c = 0; c < CHAINS; ++c) { x[c] = seed + threadIdx.x * 977 + c * 131 + blockIdx.x; w[c] = (seed * (c + 3)) | 1; }
    const u64 twoq = 2 * q;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) x[c] = __umul64hi(x[c], w[c]) + 1;                       // mul.hi.u64
            else if (MODE == 1) x[c] = x[c] * w[c] + 1;                             // mul.lo.u64
            else if (MODE == 2) {                                                    // Shoup lazy modmul
                x[c] = x[c] * w[c] - __umul64hi(x[c], qinv ^ w[c]) * q;
            } else if (MODE == 3) {                                                  // Montgomery lazy modmul
                u64 hi = __umul64hi(x[c], w[c]), lo = x[c] * w[c];
                x[c] = hi - __umul64hi(lo * qinv, q) + q;
            } else if (MODE == 4) {                                                  // Harvey CT butterfly (Shoup)
                u64 U = x[c], V = w[c];
                U = U >= twoq ? U - twoq : U;
                u64 T = V * qinv - __umul64hi(V, seed) * q;
                x[c] = U + T; w[c] = U - T + twoq;
            } else if (MODE == 5) {                                                  // 32-bit IMAD
                unsigned a = (unsigned)x[c], b = (unsigned)w[c];
                a = a * b + 7u; x[c] = a;
            } else if (MODE == 6) {                                                  // IMAD.WIDE (mad.wide.u32)
                unsigned a = (unsigned)x[c], b = (unsigned)w[c];
                u64 r; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(x[c]));
                x[c] = r;
            } else if (MODE == 7) {                                                  // 64-bit add + conditional subtract
                u64 s = x[c] + w[c];
                x[c] = s >= twoq ? s - twoq : s;
            }
        }
    }
    long long t1 = clock64();
    u64 acc = 0;
    for (int c = 0; c < CHAINS; ++c) acc ^= x[c] ^ w[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE> void run(const char *name, int sms) {
    const int blocks = sms * 8, threads = 256;
    u64 *out; long long *clk, hclk = 0;
    cudaMalloc(&out, sizeof(u64) * blocks * threads); cudaMalloc(&clk, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const u64 q = 0x1000000000248001ULL, qinv = 0x9e3779b97f4a7c15ULL;
    k<MODE><<<blocks, threads>>>(out, 12345, q, qinv, clk);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) k<MODE><<<blocks, threads>>>(out, 12345 + r, q, qinv, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&hclk, clk, 8, cudaMemcpyDeviceToHost);
    double ops = (double)reps * blocks * threads * ITER * CHAINS;
    // per-SM cycles from the in-kernel clock of block 0 (8 resident blocks/SM share the SM for the whole run)
    double ops_per_clk_sm = (double)(8 * threads) * ITER * CHAINS / (double)hclk;
    printf("%-28s %8.3f ms  %9.2f Gop/s  %7.2f ops/clk/SM (clock64)  clk=%lld  err=%s\n", name, ms / reps, ops / (ms * 1e6),
           ops_per_clk_sm, hclk, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(clk);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s sms=%d smem_optin=%zu l2=%d MB\n", p.name, p.multiProcessorCount, p.sharedMemPerBlockOptin, p.l2CacheSize >> 20);
    int sms = p.multiProcessorCount;
    run<5>("imad32", sms);
    run<6>("imad.wide.u32", sms);
    run<7>("add64+csub", sms);
    run<0>("mul.hi.u64", sms);
    run<1>("mul.lo.u64", sms);
    run<2>("shoup_lazy modmul", sms);
    run<3>("montgomery_lazy modmul", sms);
    run<4>("harvey butterfly (shoup)", sms);
    return 0;
}
