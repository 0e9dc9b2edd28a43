// tools/pipebench2.cu -- issue cost of the instructions a 64-bit modular butterfly is made of, on real clocks.
// Every kernel reads %globaltimer and clock64 at both ends, so the SM clock during the loop is measured rather than
// assumed, and the cost is printed as SM cycles per warp-instruction group per SMSP.  Groups mix NW wide multiplies
// (IMAD.WIDE), NI narrow multiply-adds (IMAD), NA ALU instructions (IADD3 / LOP3 / SHF in turn) on independent chains:
// if the FMA and ALU pipes overlap, cost(NW, 0, NA) = max(cost(NW), cost(NA)); if they do not, the costs add.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/pipebench2 tools/pipebench2.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

#define CH 8

__device__ __forceinline__ u64 gtime() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int NW, int NI, int NA, int WKIND, int AK>
__global__ void __launch_bounds__(128) k(u32 *out, u32 seed, int iters, long long *clk, u64 *ns) {
    u32 a[CH], b[CH], e[CH], f[CH];
    u64 w[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        a[c] = seed + threadIdx.x * 977 + c; b[c] = seed * 3 + c + blockIdx.x + threadIdx.x * 131; w[c] = ((u64)a[c] << 32) | b[c];
        e[c] = a[c] ^ 0x5bd1e995u; f[c] = b[c] + 0x9e3779b9u;
    }
    const u32 s3 = seed ^ (threadIdx.x * 2654435761u);
    __syncthreads();
    const u64 g0 = gtime();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                if (WKIND == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((u32)w[c]), "r"(b[c]));
                if (WKIND == 1) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((u32)w[c]), "r"(b[c]));
                if (WKIND == 2) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[c]) : "r"((u32)w[c]), "r"(b[c]));
                if (WKIND == 3) asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c]));
            }
#pragma unroll
            for (int i = 0; i < NI; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(e[c]) : "r"(f[c]));
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                const int kind = AK == 0 ? i % 3 : AK;
                if (kind == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(f[c]));
                if (kind == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(f[c]) : "r"(a[c]), "r"(s3));
                if (kind == 2) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(f[c]), "r"(s3));
            }
        }
    }
    const long long t1 = clock64();
    const u64 g1 = gtime();
    u32 acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= a[c] ^ b[c] ^ e[c] ^ f[c] ^ (u32)w[c] ^ (u32)(w[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) { clk[blockIdx.x] = t1 - t0; ns[blockIdx.x] = g1 - g0; }
}

template <int NW, int NI, int NA, int WKIND = 0, int AK = 0> void run(const char *name, int sms, int warps_per_smsp) {
    const int threads = 128, blocks_per_sm = warps_per_smsp;   // 128 threads = 4 warps = one per SMSP
    const int blocks = sms * blocks_per_sm;
    const int iters = 4096;
    u32 *out; long long *clk; u64 *ns;
    cudaMalloc(&out, sizeof(u32) * blocks * threads); cudaMalloc(&clk, 8 * blocks); cudaMalloc(&ns, 8 * blocks);
    k<NW, NI, NA, WKIND, AK><<<blocks, threads>>>(out, 1, 64, clk, ns);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<NW, NI, NA, WKIND, AK><<<blocks, threads>>>(out, 12345, iters, clk, ns);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(blocks); std::vector<u64> hn(blocks);
    cudaMemcpy(h.data(), clk, 8 * blocks, cudaMemcpyDeviceToHost);
    cudaMemcpy(hn.data(), ns, 8 * blocks, cudaMemcpyDeviceToHost);
    double cyc = 0, nsec = 0; for (int i = 0; i < blocks; ++i) { cyc += h[i]; nsec += hn[i]; } cyc /= blocks; nsec /= blocks;
    const double groups = (double)warps_per_smsp * iters * CH;        // warp-groups issued per SMSP
    printf("%-30s w/SMSP=%2d  %6.2f clk64-cycles/group  %6.2f ns-cycles@%.3fGHz/group  (kernel %.3f ms, block %.3f ms, clk64 rate %.3f GHz)\n",
           name, warps_per_smsp, cyc / groups, nsec * 1e-9 * 1.965e9 / groups, 1.965, ms, nsec * 1e-6, cyc / nsec);
    cudaFree(out); cudaFree(clk); cudaFree(ns);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s sms=%d clockRate=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const int sms = p.multiProcessorCount;
    for (int wps : {4, 8, 12}) {
        run<1, 0, 0, 0>("WIDE.U32 (mad, chained)", sms, wps);
        run<1, 0, 0, 1>("WIDE.S32 (mad, chained)", sms, wps);
        run<1, 0, 0, 2>("WIDE.U32 (mul, no addend)", sms, wps);
        run<1, 0, 0, 3>("IMAD.HI", sms, wps);
        run<0, 1, 0>("IMAD", sms, wps);
        run<0, 0, 1>("IADD3", sms, wps);
        run<0, 0, 3>("IADD3+LOP3+SHF", sms, wps);
        run<0, 0, 2, 0, 1>("LOP3 x2", sms, wps);
        run<0, 0, 2, 0, 2>("SHF x2", sms, wps);
        run<1, 0, 2, 0, 1>("WIDE + 2 LOP3", sms, wps);
        run<1, 0, 4, 0, 1>("WIDE + 4 LOP3", sms, wps);
        run<1, 0, 6, 0, 1>("WIDE + 6 LOP3", sms, wps);
        run<0, 1, 2, 0, 1>("IMAD + 2 LOP3", sms, wps);
        run<1, 0, 1>("WIDE + 1 ALU", sms, wps);
        run<1, 0, 2>("WIDE + 2 ALU", sms, wps);
        run<1, 0, 3>("WIDE + 3 ALU", sms, wps);
        run<1, 0, 4>("WIDE + 4 ALU", sms, wps);
        run<1, 1, 0>("WIDE + 1 IMAD", sms, wps);
        run<1, 1, 2>("WIDE + 1 IMAD + 2 ALU", sms, wps);
        run<0, 1, 1>("IMAD + 1 ALU", sms, wps);
        run<0, 1, 2>("IMAD + 2 ALU", sms, wps);
        run<5, 0, 10>("5 WIDE + 10 ALU (fold bf)", sms, wps);
        run<5, 4, 12>("5 WIDE + 4 IMAD + 12 ALU", sms, wps);
    }
    return 0;
}
