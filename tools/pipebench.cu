// tools/pipebench.cu -- issue-rate probes of the sm_100a pipes a 64-bit modular butterfly touches, with loop-carried
// operands so that ptxas cannot hoist or fold anything (tools/microbench.cu's mad.wide row had loop-invariant factors
// and measured the add, not the multiply).  Rates are warp-instructions per SM-clock per SMSP from clock64().
// Every kernel's loop body is checked by tools/sass_census.py-style counting in profiles/ (see r02_pipebench.txt).
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/pipebench tools/pipebench.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

#define ITER 2048
#define CH 8

template <int MODE>
__global__ void __launch_bounds__(256) k(u32 *out, u32 seed, long long *clk) {
    u32 a[CH], b[CH];
    u64 w[CH];
    double d[CH], e[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        a[c] = seed + threadIdx.x * 977 + c; b[c] = seed * 3 + c + blockIdx.x + threadIdx.x * 131; w[c] = ((u64)a[c] << 32) | b[c];
        d[c] = 1.0 + 1e-9 * a[c]; e[c] = 1.0 - 1e-9 * b[c];
    }
    const double kd = 1.0000001 + 1e-12 * (seed + threadIdx.x);
    const u32 s3 = seed ^ (threadIdx.x * 2654435761u);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            // every multiply has a loop-carried factor (a[c] or lo(w[c])), so nothing is loop invariant
            if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((u32)w[c]), "r"(b[c]));                 // IMAD.WIDE.U32
            if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c]));                                   // IMAD
            if (MODE == 2) asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c]));                                   // IMAD.HI
            if (MODE == 3) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((u32)w[c]), "r"(b[c]));                 // IMAD.WIDE (signed)
            if (MODE == 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(kd), "d"(e[c]));                          // DFMA
            if (MODE == 5) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(e[c]));                                       // DADD
            if (MODE == 6) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(kd));                                         // DMUL
            if (MODE == 7) { asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d[c]) : "r"(a[c])); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c])); }   // I2F.F64.S32 + IADD
            if (MODE == 8) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[c]) : "r"(b[c]), "r"(s3));   // IADD3 (3 inputs)
            if (MODE == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(s3));                    // LOP3
            if (MODE == 10) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b[c]), "r"(s3));                  // SHF
            if (MODE == 11) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((u32)w[c]), "r"(b[c]));              // WIDE : DFMA 1:1
                              asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(kd), "d"(e[c])); }
            if (MODE == 12) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((u32)w[c]), "r"(b[c]));              // WIDE : IADD3 1:2
                              asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[c]) : "r"(b[c]), "r"(s3));
                              asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[c]) : "r"(a[c]), "r"(s3)); }
            if (MODE == 13) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((u32)w[c]), "r"(b[c]));              // WIDE : IADD3 : LOP3 : DFMA
                              asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[c]) : "r"(b[c]), "r"(s3));
                              asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[c]) : "r"(a[c]), "r"(s3));
                              asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(kd), "d"(e[c])); }
            if (MODE == 14) { asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c]));                                // IMAD : DFMA 1:1
                              asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(kd), "d"(e[c])); }
            if (MODE == 15) { asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c]));                                // IMAD : IADD3 : DFMA
                              asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[c]) : "r"(a[c]), "r"(s3));
                              asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(kd), "d"(e[c])); }
            if (MODE == 16) { asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d[c]) : "r"(a[c]));                                    // I2F.U32 + DFMA + bit extraction
                              asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(kd), "d"(e[c]));
                              asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %1; add.u32 %0, %0, lo;}" : "+r"(a[c]) : "d"(d[c])); }
            if (MODE == 17) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; shf.l.wrap.b32 hi, lo, hi, 1; mov.b64 %0, {lo, hi};}" : "+l"(w[c]));   // SHF (64-bit funnel)
            if (MODE == 18) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((u32)w[c]), "r"(b[c]));              // WIDE : IMAD 1:1
                              asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c])); }
            if (MODE == 19) { asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c]));                                // IMAD.HI : IADD3 1:1
                              asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[c]) : "r"(a[c]), "r"(s3)); }
        }
    }
    long long t1 = clock64();
    u32 acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= a[c] ^ b[c] ^ (u32)w[c] ^ (u32)(w[c] >> 32) ^ (u32)__double2loint(d[c]) ^ (u32)__double2hiint(d[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, int sms, int per_iter) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE>, 256, 0);
    if (occ > 4) occ = 4;   // 32 warps per SM (8 per SMSP) x 8 chains: enough to cover any latency, same for every row
    const int blocks = sms * occ, threads = 256;
    u32 *out; long long *clk;
    cudaMalloc(&out, sizeof(u32) * blocks * threads); cudaMalloc(&clk, 8 * blocks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < 3; ++r) k<MODE><<<blocks, threads>>>(out, 12345, clk);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int r = 0; r < reps; ++r) k<MODE><<<blocks, threads>>>(out, 12345 + r, clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(blocks);
    cudaMemcpy(h.data(), clk, 8 * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    const double inst = (double)occ * threads * ITER * CH * per_iter;
    const double per_clk = inst / avg;
    const double ghz = avg / (ms / reps * 1e6);
    printf("%-34s occ=%d  %6.1f thread-inst/clk/SM = %.3f warp-inst/clk/SMSP = %5.2f clk per warp-group of %d   (%.3f ms, ~%.2f GHz)\n",
           name, occ, per_clk, per_clk / 128.0, per_iter * 128.0 / per_clk, per_iter, ms / reps, ghz);
    cudaFree(out); cudaFree(clk);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s sms=%d\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run<0>("IMAD.WIDE.U32 (carried factor)", sms, 1);
    run<3>("IMAD.WIDE signed", sms, 1);
    run<1>("IMAD", sms, 1);
    run<2>("IMAD.HI", sms, 1);
    run<4>("DFMA", sms, 1);
    run<5>("DADD", sms, 1);
    run<6>("DMUL", sms, 1);
    run<7>("I2F.F64.S32 + IADD", sms, 2);
    run<8>("IADD3 (3 inputs)", sms, 1);
    run<9>("LOP3", sms, 1);
    run<10>("SHF", sms, 1);
    run<17>("SHF on a 64-bit pair", sms, 1);
    run<18>("WIDE : IMAD 1:1", sms, 2);
    run<11>("WIDE : DFMA 1:1", sms, 2);
    run<12>("WIDE : IADD3 : LOP3", sms, 3);
    run<13>("WIDE : IADD3 : LOP3 : DFMA", sms, 4);
    run<14>("IMAD : DFMA 1:1", sms, 2);
    run<15>("IMAD : LOP3 : DFMA", sms, 3);
    run<16>("I2F.U32 : DFMA : MOV+IADD", sms, 3);
    run<19>("IMAD.HI : LOP3 1:1", sms, 2);
    return 0;
}
