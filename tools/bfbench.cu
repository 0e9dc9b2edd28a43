// tools/bfbench.cu -- register-resident butterfly throughput: the integer Shoup butterfly of ntt_big.cu against a
// butterfly whose quotient comes from the FP64 pipe (B200 issues DFMA at 64 lanes/clk/SM, idle in the integer kernels).
//
//   x*w mod q with x = x1*2^32 + x0:  S = x1*W1 + x0*w  (W1 = w*2^32 mod q)  is congruent to x*w and S/q < 2^33, so
//   t = rint(x1*(W1/q) + x0*(w/q)) in double precision is within 1/2 + 2^-19 of S/q and  S - t*q  lies in
//   [-0.51q, 0.51q]: 3 wide + 4 narrow integer multiply-adds (instead of 5 + 4), no carry chain for the quotient and a
//   product that is centred, which removes the per-butterfly conditional subtraction of the forward butterfly.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bfbench tools/bfbench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../matrix-fhe-lattigo_b200/csrc/modarith.cuh"

typedef unsigned int u32;
typedef ulonglong2 tw_t;
struct __align__(32) twf_t { u64 w, w1; double a, b; };

struct QC {
    u64 q, twoq, fourq, nq, nfourq;
    unsigned csh, cmul;
    u64 c0;      // FP64 product: additive constant that cancels the 0x43300000 of the magic-add high word
};

__device__ __forceinline__ void bf_int(u64 &U, u64 &V, const tw_t w, const QC &c) {
    const u64 u = csub(U, c.fourq);
    const u64 t = shoup4(V, w.x, w.y, c.nq);
    U = u + t;
    V = u - t + c.fourq;
}

// shoup4 with the quotient summed on the ALU pipe: t = C + hi32(A) + hi32(B) as a 3-input add with carries instead of
// folding one term into the addend of the wide multiply (which costs two moves to build the register pair)
__device__ __forceinline__ u64 shoup4_alu(u64 x, u64 w, u64 ws, u64 nq) {
    const unsigned x0 = lo32(x), x1 = hi32(x);
    const u64 C = mulwide_u32(x1, hi32(ws));
    const unsigned a = hi32(mulwide_u32(x0, hi32(ws))), b = hi32(mulwide_u32(x1, lo32(ws)));
    unsigned t0, t1;
    asm("{\n\t.reg .u32 s, k;\n\tadd.cc.u32 s, %2, %3;\n\taddc.u32 k, 0, 0;\n\tadd.cc.u32 %0, %4, s;\n\taddc.u32 %1, %5, k;\n\t}"
        : "=r"(t0), "=r"(t1) : "r"(a), "r"(b), "r"(lo32(C)), "r"(hi32(C)));
    const u64 r = madwide_u32(t0, lo32(nq), mulwide_u32(x0, lo32(w)));
    unsigned r1 = hi32(r);
    r1 = madlo_u32(x0, hi32(w), r1);
    r1 = madlo_u32(x1, lo32(w), r1);
    r1 = madlo_u32(t0, hi32(nq), r1);
    r1 = madlo_u32(t1, lo32(nq), r1);
    return ((u64)r1 << 32) | lo32(r);
}
// quotient through carry-chained mad.hi: t = x1*wsh + hi32(x0*wsh) + hi32(x1*wsl) with the additions folded into the multiplies
__device__ __forceinline__ u64 shoup4_madhi(u64 x, u64 w, u64 ws, u64 nq) {
    const unsigned x0 = lo32(x), x1 = hi32(x);
    const u64 C = mulwide_u32(x1, hi32(ws));
    unsigned t0 = lo32(C), t1 = hi32(C);
    asm("mad.hi.cc.u32 %0, %2, %3, %0;\n\taddc.u32 %1, %1, 0;" : "+r"(t0), "+r"(t1) : "r"(x0), "r"(hi32(ws)));
    asm("mad.hi.cc.u32 %0, %2, %3, %0;\n\taddc.u32 %1, %1, 0;" : "+r"(t0), "+r"(t1) : "r"(x1), "r"(lo32(ws)));
    const u64 r = madwide_u32(t0, lo32(nq), mulwide_u32(x0, lo32(w)));
    unsigned r1 = hi32(r);
    r1 = madlo_u32(x0, hi32(w), r1);
    r1 = madlo_u32(x1, lo32(w), r1);
    r1 = madlo_u32(t0, hi32(nq), r1);
    r1 = madlo_u32(t1, lo32(nq), r1);
    return ((u64)r1 << 32) | lo32(r);
}
template <int V>
__device__ __forceinline__ void bf_int_v(u64 &U, u64 &V_, const tw_t w, const QC &c) {
    const u64 u = csub(U, c.fourq);
    const u64 t = V == 2 ? shoup4_madhi(V_, w.x, w.y, c.nq) : V == 1 ? shoup4_alu(V_, w.x, w.y, c.nq) : shoup4(V_, w.x, w.y, c.nq);
    U = u + t;
    V_ = u - t + c.fourq;
}

// x*w - t*q + off (mod 2^64) with |x*w - t*q| <= 0.51q; off is folded into c (c = off - (0x43300000*lo32(nq) << 32))
template <int CVT>
__device__ __forceinline__ u64 fmul(u64 x, const twf_t &tw, u64 nq, u64 c) {
    const u32 x0 = lo32(x), x1 = hi32(x);
    if (CVT == 2) {   // diagnostic: integer part only (the quotient is garbage)
        const u32 t0 = x0 ^ x1, t1 = 0x43300000u;
        u64 acc = madwide_u32(x0, lo32(tw.w), c);
        acc = madwide_u32(x1, lo32(tw.w1), acc);
        acc = madwide_u32(t0, lo32(nq), acc);
        u32 h = hi32(acc);
        h = madlo_u32(x0, hi32(tw.w), h);
        h = madlo_u32(x1, hi32(tw.w1), h);
        h = madlo_u32(t0, hi32(nq), h);
        h = madlo_u32(t1, lo32(nq), h);
        return ((u64)h << 32) | lo32(acc);
    }
    if (CVT == 3) {   // diagnostic: FP64 part only
        const double e = fma(__uint2double_rn(x0), tw.b, __uint2double_rn(x1) * tw.a);
        const double m = e + 4503599627370496.0;
        return (u64)__double_as_longlong(m) + c;
    }
    double d0, d1;
    if (CVT == 0) { d0 = __uint2double_rn(x0); d1 = __uint2double_rn(x1); }
    else if (CVT == 4) {   // diagnostic: no conversions (bit patterns reinterpreted)
        d0 = __hiloint2double((int)x1, (int)x0); d1 = __hiloint2double((int)x0, (int)x1);
    } else {
        d0 = __hiloint2double(0x43300000, (int)x0) - 4503599627370496.0;
        d1 = __hiloint2double(0x43300000, (int)x1) - 4503599627370496.0;
    }
    const double e = fma(d0, tw.b, d1 * tw.a);
    const double m = e + 4503599627370496.0;
    const u32 t0 = (u32)__double2loint(m), t1 = (u32)__double2hiint(m);
    u64 acc = madwide_u32(x0, lo32(tw.w), c);
    acc = madwide_u32(x1, lo32(tw.w1), acc);
    acc = madwide_u32(t0, lo32(nq), acc);
    u32 h = hi32(acc);
    h = madlo_u32(x0, hi32(tw.w), h);
    h = madlo_u32(x1, hi32(tw.w1), h);
    h = madlo_u32(t0, hi32(nq), h);
    h = madlo_u32(t1, lo32(nq), h);
    return ((u64)h << 32) | lo32(acc);
}
template <int CVT>
__device__ __forceinline__ void bf_f64(u64 &U, u64 &V, const twf_t &w, const QC &c) {
    const u64 d = fmul<CVT>(V, w, c.nq, c.c0);      // signed, |d| <= 0.51q
    V = U - d;
    U = U + d;
}
// x < 8q -> [0, 2q)
__device__ __forceinline__ u64 red2q(u64 x, const QC &c) {
    const unsigned k = __umulhi((unsigned)(x >> c.csh), c.cmul);
    const u64 r = madwide_u32(k, lo32(c.nq), x);
    return ((u64)madlo_u32(k, hi32(c.nq), hi32(r)) << 32) | lo32(r);
}

#define TWN 256
template <int MODE, int LAYERS>
__global__ void __launch_bounds__(128, 8) k_bf(u64 *data, const tw_t *twi, const twf_t *twf, QC c, int iters, long long *clk, int lanemask) {
    u64 x[8];
    u64 *p = data + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = p[k];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const unsigned P = (it * 7 + (threadIdx.x & lanemask)) & 31;
        if (MODE == 0 || MODE == 6 || MODE == 8) {
#pragma unroll
            for (int l = 0; l < LAYERS; ++l) {
                const int j = l % 3, half = 4 >> j;
#pragma unroll
                for (int g = 0; g < (1 << j); ++g) {
                    const tw_t w = twi[(((P + l / 3) << j) + g) & (TWN - 1)];
#pragma unroll
                    for (int o = 0; o < half; ++o) { if (MODE == 8) bf_int_v<2>(x[g * 2 * half + o], x[g * 2 * half + o + half], w, c); else if (MODE == 6) bf_int_v<1>(x[g * 2 * half + o], x[g * 2 * half + o + half], w, c); else bf_int(x[g * 2 * half + o], x[g * 2 * half + o + half], w, c); }
                }
            }
        } else if (MODE == 7) {
            // primes with 15q < 2^64: no conditional subtraction for 3 layers ([0,2q) -> [0,14q)), then one quotient-estimate reduction
#pragma unroll
            for (int l = 0; l < LAYERS; ++l) {
                const int j = l % 3, half = 4 >> j;
#pragma unroll
                for (int g = 0; g < (1 << j); ++g) {
                    const tw_t w = twi[(((P + l / 3) << j) + g) & (TWN - 1)];
#pragma unroll
                    for (int o = 0; o < half; ++o) {
                        u64 &U = x[g * 2 * half + o], &V = x[g * 2 * half + o + half];
                        const u64 t = shoup4(V, w.x, w.y, c.nq);
                        V = U - t + c.fourq;
                        U = U + t;
                    }
                }
                if (j == 2) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) x[k] = red2q(x[k], c);
                }
            }
        } else {
            // inputs in [0,2q) plus an offset C = k*q: the window [C - 0.51 L q, C + 2q + 0.51 L q] must stay inside
            // [0, 8q):  L <= 3 with C = 2q, L <= 5 with C = 3q
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] += (LAYERS > 3 ? c.twoq + c.q : c.twoq);
#pragma unroll
            for (int l = 0; l < LAYERS; ++l) {
                const int j = l % 3, half = 4 >> j;
#pragma unroll
                for (int g = 0; g < (1 << j); ++g) {
                    const twf_t w = twf[(((P + l / 3) << j) + g) & (TWN - 1)];
#pragma unroll
                    for (int o = 0; o < half; ++o) bf_f64<MODE - 1>(x[g * 2 * half + o], x[g * 2 * half + o + half], w, c);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = red2q(x[k], c);
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        u64 v = x[k];
        if (MODE == 0 || MODE == 6 || MODE == 8) v = csub(csub(csub(v, c.fourq), c.twoq), c.q); else v = csub(v, c.q);
        p[k] = v;
    }
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// raw issue-rate probes of the FP64 pipe next to the integer multiply pipe
template <int MODE>
__global__ void __launch_bounds__(256) k_rate(double *out, double seed, long long *clk) {
    double a[8]; u32 ia[8]; u64 wa[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { a[c] = seed + threadIdx.x + c; ia[c] = threadIdx.x * 977 + c; wa[c] = ia[c]; }
    const double b = seed * 0.5;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 4096; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a[c]) : "d"(b));
            if (MODE == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[c]) : "d"(b));
            if (MODE == 2) { asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(a[c]) : "r"(ia[c])); asm volatile("add.u32 %0, %0, 1;" : "+r"(ia[c])); }
            if (MODE == 3) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a[c]) : "d"(b));
                             asm volatile("mad.lo.u32 %0, %0, %0, %0;" : "+r"(ia[c])); }
            if (MODE == 4) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a[c]) : "d"(b));
                             asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(wa[c]) : "r"(ia[c])); }
            if (MODE == 5) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a[c]) : "d"(b));
                             asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(wa[c]) : "r"(ia[c]));
                             asm volatile("mad.lo.u32 %0, %0, %0, %0;" : "+r"(ia[c]));
                             asm volatile("add.u32 %0, %0, 1;" : "+r"(ia[c])); }
        }
    }
    const long long t1 = clock64();
    double acc = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) acc += a[c] + (double)ia[c] + (double)wa[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

static u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((unsigned __int128)a * b % q); }

template <int MODE> void rate(const char *name, int sms, int per_iter) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rate<MODE>, 256, 0);
    const int blocks = sms * occ;
    double *out; long long *clk;
    cudaMalloc(&out, 8 * blocks * 256); cudaMalloc(&clk, 8 * blocks);
    for (int r = 0; r < 3; ++r) k_rate<MODE><<<blocks, 256>>>(out, 1.0, clk);
    cudaDeviceSynchronize();
    std::vector<long long> h(blocks);
    cudaMemcpy(h.data(), clk, 8 * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += v; avg /= blocks;
    const double inst = (double)occ * 256 * 4096 * 8 * per_iter;
    printf("%-34s occ=%d  %.1f thread-inst/clk/SM = %.3f warp-inst/clk/SMSP\n", name, occ, inst / avg, inst / avg / 128);
    cudaFree(out); cudaFree(clk);
}

template <int MODE, int LAYERS> double bench(const char *name, int sms, u64 *d_data, const std::vector<u64> &host_in, const tw_t *twi, const twf_t *twf,
                                 QC c, std::vector<u64> *result) {
    const int iters = 64;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bf<MODE, LAYERS>, 128, 0);
    const int blocks = sms * occ * 4;
    long long *clk; cudaMalloc(&clk, 8 * blocks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    const int reps = getenv("BF_REPS") ? atoi(getenv("BF_REPS")) : 5;
    for (int r = 0; r < reps; ++r) {
        cudaMemcpy(d_data, host_in.data(), host_in.size() * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        k_bf<MODE, LAYERS><<<blocks, 128>>>(d_data, twi, twf, c, iters, clk, getenv("BF_LANEMASK") ? atoi(getenv("BF_LANEMASK")) : 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
    result->resize((size_t)blocks * 128 * 8);
    cudaMemcpy(result->data(), d_data, result->size() * 8, cudaMemcpyDeviceToHost);
    const double bfs = (double)blocks * 128 * iters * 4.0 * LAYERS;
    std::vector<long long> hc(blocks);
    cudaMemcpy(hc.data(), clk, 8 * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : hc) avg += v; avg /= blocks;
    // a block's warps share an SMSP with occ*4/4 = occ warps: cycles per warp-butterfly on one SMSP
    const double cyc_per_bf = avg / (iters * 4.0 * LAYERS) / occ;
    printf("%-34s occ=%d blocks/SM  %.3f ms  %.1f G butterflies/s  %.1f SMSP-cycles per warp-butterfly\n", name, occ, best, bfs / best / 1e6, cyc_per_bf);
    cudaFree(clk);
    return bfs / best / 1e6;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("device %s sms=%d\n", prop.name, sms);
    const u64 q = 0x1000000000248001ULL;
    QC c{};
    c.q = q; c.twoq = 2 * q; c.fourq = 4 * q; c.nq = 0 - q; c.nfourq = 0 - 4 * q;
    {   // quotient estimate for x < 16q (same construction as b200ring.cu)
        int bl = 0;
        while (bl < 64 && (q >> bl)) ++bl;
        c.csh = bl > 28 ? (unsigned)(bl - 28) : 0u;
        c.cmul = (unsigned)((1ULL << 32) / ((q >> c.csh) + 1));
    }
    c.c0 = 0 - ((u64)(0x43300000u * (u32)c.nq) << 32);
    std::vector<tw_t> ti(TWN); std::vector<twf_t> tf(TWN);
    u64 s = 0x9E3779B97F4A7C15ULL;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (int i = 0; i < TWN; ++i) {
        const u64 w = rnd() % q;
        ti[i].x = w; ti[i].y = (u64)(((unsigned __int128)w << 64) / q);
        tf[i].w = w; tf[i].w1 = (u64)(((unsigned __int128)w << 32) % q);
        tf[i].a = (double)((long double)tf[i].w1 / (long double)q);
        tf[i].b = (double)((long double)w / (long double)q);
    }
    tw_t *d_ti; twf_t *d_tf;
    cudaMalloc(&d_ti, sizeof(tw_t) * TWN); cudaMalloc(&d_tf, sizeof(twf_t) * TWN);
    cudaMemcpy(d_ti, ti.data(), sizeof(tw_t) * TWN, cudaMemcpyHostToDevice);
    cudaMemcpy(d_tf, tf.data(), sizeof(twf_t) * TWN, cudaMemcpyHostToDevice);

    rate<0>("DFMA", sms, 1);
    rate<1>("DADD", sms, 1);
    rate<2>("I2F.F64.U32 (+IADD)", sms, 2);
    rate<3>("DFMA : IMAD 1:1", sms, 2);
    rate<4>("DFMA : IMAD.WIDE 1:1", sms, 2);
    rate<5>("DFMA : WIDE : IMAD : IADD", sms, 4);

    const size_t words = (size_t)sms * 8 * 4 * 128 * 8;
    std::vector<u64> in(words);
    for (auto &v : in) v = rnd() % q;
    u64 *d; cudaMalloc(&d, words * 8);
    std::vector<u64> r0, r1, r2, r3, r4;
    bench<0, 3>("integer shoup4 + csub (current)", sms, d, in, d_ti, d_tf, c, &r0);
    bench<1, 3>("FP64 quotient, cvt.f64.u32", sms, d, in, d_ti, d_tf, c, &r1);
    bench<2, 3>("FP64 quotient, magic-number cvt", sms, d, in, d_ti, d_tf, c, &r2);
    bench<0, 5>("integer, 5 layers/iter", sms, d, in, d_ti, d_tf, c, &r3);
    bench<1, 5>("FP64 cvt, 5 layers per reduction", sms, d, in, d_ti, d_tf, c, &r4);
    { std::vector<u64> rx;
      bench<6, 5>("integer, quotient summed on ALU", sms, d, in, d_ti, d_tf, c, &rx);
      size_t badx = 0; for (size_t i = 0; i < rx.size() && i < r3.size(); ++i) badx += rx[i] != r3[i];
      printf("   parity vs integer: %zu mismatches\n", badx);
      bench<8, 5>("integer, quotient by mad.hi.cc chain", sms, d, in, d_ti, d_tf, c, &rx);
      badx = 0; for (size_t i = 0; i < rx.size() && i < r3.size(); ++i) badx += rx[i] != r3[i];
      printf("   parity vs integer: %zu mismatches\n", badx);
      bench<0, 6>("integer, 6 layers/iter", sms, d, in, d_ti, d_tf, c, &r3);
      bench<7, 6>("integer, csub-free 3-layer groups", sms, d, in, d_ti, d_tf, c, &rx);
      badx = 0; for (size_t i = 0; i < rx.size() && i < r3.size(); ++i) badx += rx[i] != r3[i];
      printf("   parity vs integer: %zu mismatches\n", badx);
      bench<3, 5>("diag: integer part only, 5 layers", sms, d, in, d_ti, d_tf, c, &rx);
      bench<4, 5>("diag: FP64 part only, 5 layers", sms, d, in, d_ti, d_tf, c, &rx);
      bench<5, 5>("diag: no conversions, 5 layers", sms, d, in, d_ti, d_tf, c, &rx); }
    size_t n = r0.size() < r1.size() ? r0.size() : r1.size();
    if (r2.size() < n) n = r2.size();
    size_t bad1 = 0, bad2 = 0;
    for (size_t i = 0; i < n; ++i) { bad1 += r0[i] != r1[i]; bad2 += r0[i] != r2[i]; }
    printf("parity vs integer butterflies over %zu words: cvt %zu mismatches, magic %zu mismatches\n", n, bad1, bad2);
    size_t m = r3.size() < r4.size() ? r3.size() : r4.size(), bad3 = 0;
    for (size_t i = 0; i < m; ++i) bad3 += r3[i] != r4[i];
    printf("parity 5-layer: %zu mismatches over %zu words\n", bad3, m);
    // host check of the first thread's words against exact arithmetic (3-layer variant, 64 iterations)
    {
        u64 x[8]; for (int k = 0; k < 8; ++k) x[k] = in[k];
        for (int it = 0; it < 64; ++it) {
            const unsigned P = (it * 7 + 0) & 31;   // thread 0: the same for every lane mask
            for (int j = 0; j < 3; ++j) {
                const int half = 4 >> j;
                for (int g = 0; g < (1 << j); ++g) {
                    const u64 w = ti[((P << j) + g) & (TWN - 1)].x;
                    for (int o = 0; o < half; ++o) {
                        u64 &U = x[g * 2 * half + o], &V = x[g * 2 * half + o + half];
                        const u64 t = mulmod(V, w, q);
                        const u64 nu = (U + t) % q, nv = (U + q - t) % q;
                        U = nu; V = nv;
                    }
                }
            }
        }
        int bad = 0; for (int k = 0; k < 8; ++k) bad += x[k] != r1[k];
        printf("host exact check (thread 0): %s\n", bad ? "MISMATCH" : "ok");
    }
    return 0;
}
