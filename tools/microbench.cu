// tools/microbench.cu -- issue-rate micro-benchmarks of the sm_100a integer pipes (the resources that bound
// a 64-bit modular butterfly).  Every kernel runs ITER x CHAINS dependent-chain instructions per thread via
// asm volatile; rates are reported per SM per clock using the SM clock measured in-kernel (clock64 / event time).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

#define ITER 4096
#define CH 8

template <int MODE>
__global__ void __launch_bounds__(256) k(u32 *out, u32 seed, long long *clk) {
    u32 a[CH], b[CH];
    u64 w[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { a[c] = seed + threadIdx.x * 977 + c; b[c] = seed * 3 + c + blockIdx.x; w[c] = ((u64)a[c] << 32) | b[c]; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));                      // IADD3
            if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c]));               // IMAD
            if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(b[c]));  // IMAD.WIDE
            if (MODE == 3) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));                    // 1:1 IADD3 : IMAD
                             asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b[c]) : "r"(a[c])); }
            if (MODE == 4) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));                    // 1:1 IADD3 : IMAD.WIDE
                             asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(b[c])); }
            if (MODE == 5) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));                      // LOP3
            if (MODE == 6) asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %1, %0, p;}" : "+r"(a[c]) : "r"(b[c]));   // ISETP + SEL
            if (MODE == 7) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %2;" : "+r"(a[c]), "+r"(b[c]) : "r"(seed));         // IADD3 + IADD3.X
            if (MODE == 8) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));                    // 2:1 IMAD : IADD3
                             asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b[c]) : "r"(a[c]));
                             asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b[c]) : "r"(a[c])); }
            if (MODE == 9) asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c]));               // IMAD.HI
            if (MODE == 10) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[c]) : "r"(a[c]), "r"(b[c]));     // IMAD.WIDE, no accumulate
            if (MODE == 11) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));                   // 2:1 IADD3 : IMAD.WIDE
                              asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(a[c]));
                              asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(b[c])); }
            if (MODE == 12) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));                   // 2:1 IADD3 : IMAD
                              asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(a[c]));
                              asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b[c]) : "r"(a[c])); }
            if (MODE == 13) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));                   // 1:1 IADD3 : IMAD.HI
                              asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(b[c]) : "r"(a[c])); }
            if (MODE == 14) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));                   // 3:1 IADD3 : IMAD.WIDE
                              asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(a[c]));
                              asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
                              asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(b[c])); }
        }
    }
    long long t1 = clock64();
    u32 acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= a[c] ^ b[c] ^ (u32)w[c] ^ (u32)(w[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, int sms, int per_iter) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE>, 256, 0);
    const int blocks = sms * occ, threads = 256;
    u32 *out; long long *clk;
    cudaMalloc(&out, sizeof(u32) * blocks * threads); cudaMalloc(&clk, 8 * blocks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < 3; ++r) k<MODE><<<blocks, threads>>>(out, 12345, clk);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int r = 0; r < reps; ++r) k<MODE><<<blocks, threads>>>(out, 12345 + r, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long *h = (long long *)malloc(8 * blocks);
    cudaMemcpy(h, clk, 8 * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    const double inst = (double)occ * threads * ITER * CH * per_iter;       // thread-instructions per SM per launch
    const double per_clk = inst / avg;                                       // all resident blocks run concurrently for ~avg cycles
    const double ghz = avg / (ms / reps * 1e6);                              // cycles per launch / launch time
    printf("%-26s occ=%d blocks/SM  %7.1f thread-inst/clk/SM  = %.3f warp-inst/clk/SMSP   (%.3f ms/launch, SM clock ~%.2f GHz)\n",
           name, occ, per_clk, per_clk / 128.0, ms / reps, ghz);
    cudaFree(out); cudaFree(clk); free(h);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s sms=%d\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run<0>("IADD3", sms, 1);
    run<5>("LOP3", sms, 1);
    run<7>("IADD3 + IADD3.X", sms, 2);
    run<6>("ISETP + SEL", sms, 2);
    run<1>("IMAD (mad.lo.u32)", sms, 1);
    run<2>("IMAD.WIDE.U32", sms, 1);
    run<3>("IADD3 : IMAD  1:1", sms, 2);
    run<8>("IADD3 : IMAD  1:2", sms, 3);
    run<4>("IADD3 : IMAD.WIDE 1:1", sms, 2);
    run<9>("IMAD.HI.U32", sms, 1);
    run<10>("IMAD.WIDE.U32 (mul)", sms, 1);
    run<12>("IADD3 : IMAD  2:1", sms, 3);
    run<11>("IADD3 : IMAD.WIDE 2:1", sms, 3);
    run<14>("IADD3 : IMAD.WIDE 3:1", sms, 4);
    run<13>("IADD3 : IMAD.HI 1:1", sms, 2);
    return 0;
}
