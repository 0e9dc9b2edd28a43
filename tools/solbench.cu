// tools/solbench.cu -- register-resident butterfly throughput of the fold ("Solinas") butterfly of solinas.cuh against
// the Shoup/Harvey butterfly of modarith.cuh (the round-1 kernels), same twiddles, bit-identical canonical results.
// 8 words per thread, groups of 3 radix-2 layers, twiddles loaded warp-uniformly from a 256-entry table.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/solbench tools/solbench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../matrix-fhe-lattigo_b200/csrc/modarith.cuh"
#include "../matrix-fhe-lattigo_b200/csrc/solinas.cuh"


// ---- design A (measured, not adopted): 31-bit split of y, twiddle halves of 32 bits, shift/mask fold.  The carry of the
// low product into the high one and the shift/mask cost more than design B's sixth multiply (profiles/r02_solbench.txt).
struct SolAConst { u64 q; u64 bias; int negc; unsigned sh; unsigned mask; unsigned pad_; };
__device__ __forceinline__ unsigned funnel_r(unsigned lo, unsigned hi, unsigned s) {
    unsigned r; asm("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(s)); return r;
}
__device__ __forceinline__ s64 sola_mul(s64 y, const stw_t w, const SolAConst &c) {
    const unsigned ylo = (unsigned)y, yhi = (unsigned)((u64)y >> 32);
    const int y0 = (int)(ylo & 0x7fffffffu);
    const int y1 = (int)funnel_r(ylo, yhi, 31);
    s64 A = madwide_s32(y0, w.w_lo, (s64)c.bias);
    A = madwide_s32(y1, w.v_lo, A);
    s64 B = madwide_s32(y0, w.w_hi, (s64)((u64)A >> 32));
    B = madwide_s32(y1, w.v_hi, B);
    const unsigned blo = (unsigned)B, bhi = (unsigned)((u64)B >> 32);
    const int shh = (int)funnel_r(blo, bhi, c.sh);
    return madwide_s32(shh, c.negc, pack64((unsigned)A, blo & c.mask));
}
__device__ __forceinline__ s64 sola_fold(s64 v, const SolAConst &c) {
    const unsigned lo = (unsigned)v, hi = (unsigned)((u64)v >> 32);
    const int h = (int)hi >> c.sh;
    return madwide_s32(h, c.negc, pack64(lo, hi & c.mask));
}
__device__ __forceinline__ u64 sola_canon(s64 v, const SolAConst &c) {
    s64 r = sola_fold(v, c);
    if (r < 0) r += (s64)c.q;
    if (r >= (s64)c.q) r -= (s64)c.q;
    return (u64)r;
}

typedef ulonglong2 tw_t;
struct QC { u64 q, twoq, fourq, nq, nfourq; };

__device__ __forceinline__ void bf_int(u64 &U, u64 &V, const tw_t w, const QC &c) {
    const u64 u = addback_neg(U + c.nfourq, c.fourq);
    const u64 t = shoup4(V, w.x, w.y, c.nq);
    U = u + t;
    V = u - t + c.fourq;
}
// forward butterfly on signed words: U' = U + T (- q when KU), V' = U - T (+ q when KV); T in about [0, q]
template <bool KU, bool KV>
__device__ __forceinline__ void bf_sola(s64 &U, s64 &V, const stw_t w, const SolAConst &c) {
    const s64 t = sola_mul(V, w, c);
    V = KV ? U - t + (s64)c.q : U - t;
    U = KU ? U + t - (s64)c.q : U + t;
}
template <bool KU, bool KV>
__device__ __forceinline__ void bf_sol(s64 &U, s64 &V, const stw_t w, const SolConst &c) {
    const s64 t = sol_mul(V, w, c);
    V = KV ? U - t + (s64)c.q : U - t;
    U = KU ? U + t - (s64)c.q : U + t;
}
template <bool KU, bool KV>
__device__ __forceinline__ void bf_sol_sh(s64 &U, s64 &V, const stw_t w, const SolConst &c) {
    const s64 t = sol_mul_sh(V, w, c);
    V = KV ? U - t + (s64)c.q : U - t;
    U = KU ? U + t - (s64)c.q : U + t;
}
// all adds with three inputs: kz is a zero the compiler cannot see (keeps add-with-carry off the multiply pipe)
template <int KU, int KV>
__device__ __forceinline__ void bf_sol_3(s64 &U, s64 &V, const stw_t w, const SolConst &c, s64 kz) {
    const s64 t = sol_mul_sh(V, w, c);
    V = U - t + (KV ? (s64)c.q : kz);
    U = U + t + (KU ? -(s64)c.q : kz);
}
__device__ __forceinline__ s64 sol_fold_3(s64 v, const SolConst &c, s64 kz) {
    const unsigned lo = (unsigned)v, hi = (unsigned)((u64)v >> 32);
    return mulwide_s32((int)hi >> c.sh, c.negc) + pack64(lo, hi & c.mask) + kz;
}
// products without fused 64-bit addends: five mul.wide and explicit add-with-carry chains (is IMAD.WIDE with a register addend dearer?)
__device__ __forceinline__ s64 add64_asm(s64 a, s64 b) {
    unsigned lo, hi;
    asm volatile("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(lo), "=r"(hi)
                 : "r"((unsigned)a), "r"((unsigned)((u64)a >> 32)), "r"((unsigned)b), "r"((unsigned)((u64)b >> 32)));
    return pack64(lo, hi);
}
__device__ __forceinline__ s64 sol_mul_unfused(s64 y, const stw_t w, const SolConst &c) {
    const unsigned ylo = (unsigned)y, yhi = (unsigned)((u64)y >> 32);
    const int y0 = (int)ylo, y1 = (int)(yhi + (ylo >> 31));
    const s64 A = add64_asm(mulwide_s32(y0, w.w_lo), mulwide_s32(y1, w.v_lo));
    const s64 B = add64_asm(mulwide_s32(y0, w.w_hi), mulwide_s32(y1, w.v_hi));
    const unsigned blo = (unsigned)B;
    const s64 sh = pack64(blo << c.sh, blo >> (32 - c.sh));
    return add64_asm(add64_asm(mulwide_s32((int)(B >> 32), c.negc), sh), A);
}
template <bool KU, bool KV>
__device__ __forceinline__ void bf_sol_unfused(s64 &U, s64 &V, const stw_t w, const SolConst &c) {
    const s64 t = sol_mul_unfused(V, w, c);
    V = KV ? U - t + (s64)c.q : U - t;
    U = KU ? U + t - (s64)c.q : U + t;
}
// the product chain starts from U, so X = U + y*w costs no addition; Y = 2U - X; no recentring (fold every 4 layers)
__device__ __forceinline__ void bf_sol_chain(s64 &U, s64 &V, const stw_t w, const SolConst &c) {
    const s64 y = V;
    const unsigned ylo = (unsigned)y, yhi = (unsigned)((u64)y >> 32);
    const int y0 = (int)ylo, y1 = (int)(yhi + (ylo >> 31));
    s64 B = mulwide_s32(y0, w.w_hi); B = madwide_s32(y1, w.v_hi, B);
    s64 T = madwide_s32(y0, w.w_lo, U); T = madwide_s32(y1, w.v_lo, T);
    T = madwide_s32((int)(B >> 32), c.negc, T);
    const unsigned blo = (unsigned)B;
    const s64 X = T + pack64(blo << c.sh, blo >> (32 - c.sh));
    V = U + U - X;
    U = X;
}
// the shifted term joins the two butterfly additions as a third input (X = U + T + sh, Y = U - T - sh) instead of being added
// to the product first: two 64-bit additions fewer per butterfly, no recentring (fold every 4 layers)
__device__ __forceinline__ void bf_sol_3in(s64 &U, s64 &V, const stw_t w, const SolConst &c) {
    const s64 y = V;
    const unsigned ylo = (unsigned)y, yhi = (unsigned)((u64)y >> 32);
    const int y0 = (int)ylo, y1 = (int)(yhi + (ylo >> 31));
    s64 A = mulwide_s32(y0, w.w_lo); A = madwide_s32(y1, w.v_lo, A);
    s64 B = mulwide_s32(y0, w.w_hi); B = madwide_s32(y1, w.v_hi, B);
    const s64 T = madwide_s32((int)(B >> 32), c.negc, A);
    const unsigned blo = (unsigned)B;
    const s64 sh = pack64(blo << c.sh, blo >> (32 - c.sh));
    V = U - T - sh;
    U = U + T + sh;
}
// inverse (Gentleman-Sande) butterfly: U' = U + V, V' = (U - V)*w
__device__ __forceinline__ void bf_sol_inv(s64 &U, s64 &V, const stw_t w, const SolConst &c) {
    const s64 d = U - V;
    U = U + V;
    V = sol_mul(d, w, c);
}

#define TWN 256
// MODE 0: Shoup/Harvey (round 1)
//      1: design A (31-bit split of y, shift/mask fold), fold of all words at the start of every 3-layer group
//      2: design A, recentred with +-q, one fold per LAYERS layers (timing only beyond 4 layers)
//      3: design B inverse (Gentleman-Sande) butterfly, sums folded every layer
//      4: design B (twiddles split at bit k-32), fold at the start of every 3-layer group, recentred with +-q
//      5: design B without any fold (timing only: the words overflow)
template <int MODE, int LAYERS>
__global__ void __launch_bounds__(128, 8) k_bf(u64 *data, const tw_t *twi, const stw_t *twa, const stw_t *tws, QC c, SolAConst sa, SolConst sc, int iters, long long *clk, s64 kz) {
    u64 x[8];
    s64 z[8];
    u64 *p = data + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = p[k]; z[k] = (s64)x[k]; }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const unsigned P = (it * 7) & 31;
        if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) z[k] = sola_fold(z[k], sa);
        }
#pragma unroll
        for (int l = 0; l < LAYERS; ++l) {
            const int j = l % 3, half = 4 >> j;
            if (MODE == 1 && j == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) z[k] = sola_fold(z[k], sa);
            }
            if (MODE == 9 && (l & 3) == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) z[k] = sol_fold_3(z[k], sc, kz);
            }
            if ((MODE == 6 || MODE == 7 || MODE == 10 || MODE == 11 || MODE == 12) && (l & 3) == 0) {   // fold every 4 layers
#pragma unroll
                for (int k = 0; k < 8; ++k) z[k] = sol_fold(z[k], sc);
            }
            if (MODE == 4 && j == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) z[k] = sol_fold(z[k], sc);
            }
#pragma unroll
            for (int g = 0; g < (1 << j); ++g) {
                const int ti = (((P + l / 3) << j) + g) & (TWN - 1);
#pragma unroll
                for (int o = 0; o < half; ++o) {
                    const int i0 = g * 2 * half + o, i1 = i0 + half;
                    if (MODE == 0) bf_int(x[i0], x[i1], twi[ti], c);
                    else if (MODE == 1) bf_sola<false, false>(z[i0], z[i1], twa[ti], sa);
                    else if (MODE == 2) {   // alternate the recentring so that every word's window stays symmetric
                        if (l & 1) bf_sola<false, true>(z[i0], z[i1], twa[ti], sa); else bf_sola<true, false>(z[i0], z[i1], twa[ti], sa);
                    } else if (MODE == 3) {
                        bf_sol_inv(z[i0], z[i1], tws[ti], sc);
                        z[i0] = sol_fold(z[i0], sc);
                    } else if (MODE == 4) {   // windows (in q): [0,1] -> +-1.48 -> +-2.45 -> +-3.43
                        if (j == 0) bf_sol<false, true>(z[i0], z[i1], tws[ti], sc); else bf_sol<true, false>(z[i0], z[i1], tws[ti], sc);
                    } else if (MODE == 6) {   // recentring pattern of a 4-layer block: (0,1) (1,0) (1,0) (0,0)
                        const int r = l & 3;
                        if (r == 0) bf_sol<false, true>(z[i0], z[i1], tws[ti], sc); else if (r == 3) bf_sol<false, false>(z[i0], z[i1], tws[ti], sc); else bf_sol<true, false>(z[i0], z[i1], tws[ti], sc);
                    } else if (MODE == 9) {
                        const int r = l & 3;
                        if (r == 0) bf_sol_3<0, 1>(z[i0], z[i1], tws[ti], sc, kz); else if (r == 3) bf_sol_3<0, 0>(z[i0], z[i1], tws[ti], sc, kz); else bf_sol_3<1, 0>(z[i0], z[i1], tws[ti], sc, kz);
                    } else if (MODE == 10) {
                        const int r = l & 3;
                        if (r == 0) bf_sol_unfused<false, true>(z[i0], z[i1], tws[ti], sc); else if (r == 3) bf_sol_unfused<false, false>(z[i0], z[i1], tws[ti], sc); else bf_sol_unfused<true, false>(z[i0], z[i1], tws[ti], sc);
                    } else if (MODE == 11) {
                        bf_sol_chain(z[i0], z[i1], tws[ti], sc);
                    } else if (MODE == 12) {
                        bf_sol_3in(z[i0], z[i1], tws[ti], sc);
                    } else if (MODE == 7) {
                        const int r = l & 3;
                        if (r == 0) bf_sol_sh<false, true>(z[i0], z[i1], tws[ti], sc); else if (r == 3) bf_sol_sh<false, false>(z[i0], z[i1], tws[ti], sc); else bf_sol_sh<true, false>(z[i0], z[i1], tws[ti], sc);
                    } else {
                        bf_sol<false, false>(z[i0], z[i1], tws[ti], sc);
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (MODE == 0) p[k] = csub(csub(csub(csub(x[k], c.fourq), c.twoq), c.q), c.q);
        else if (MODE == 1 || MODE == 2) p[k] = sola_canon(sola_fold(z[k], sa), sa);
        else p[k] = sol_canon(sol_fold(z[k], sc), sc);
    }
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

static u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((unsigned __int128)a * b % q); }

template <int MODE, int LAYERS>
double bench(const char *name, int sms, u64 *d_data, const std::vector<u64> &host_in, const tw_t *twi, const stw_t *twa, const stw_t *tws, QC c,
             SolAConst sa, SolConst sc, std::vector<u64> *result) {
    const int iters = 64;
    const char *only = getenv("SOL_ONLY");
    if (only && atoi(only) != MODE * 10 + LAYERS) { result->clear(); return 0; }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bf<MODE, LAYERS>, 128, 0);
    const int blocks = sms * occ * 4;
    long long *clk; cudaMalloc(&clk, 8 * blocks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < (only ? 1 : 5); ++r) {
        cudaMemcpy(d_data, host_in.data(), host_in.size() * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        k_bf<MODE, LAYERS><<<blocks, 128>>>(d_data, twi, twa, tws, c, sa, sc, iters, clk, getenv("SOL_KZ") ? atoll(getenv("SOL_KZ")) : 0);
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
    const double cyc_per_bf = avg / (iters * 4.0 * LAYERS) / occ;
    printf("%-52s occ=%d  %.3f ms  %7.1f G butterflies/s  %5.1f SMSP-cycles per warp-butterfly\n", name, occ, best, bfs / best / 1e6, cyc_per_bf);
    cudaFree(clk);
    return bfs / best / 1e6;
}

static void split_bal(s64 v, int s, int *lo, int *hi) {
    const s64 h = (v + (1LL << (s - 1))) >> s;
    if (h > 0x7fffffffLL || h < -0x80000000LL) { printf("twiddle not representable\n"); exit(1); }
    *hi = (int)h; *lo = (int)(v - (h << s));
}

int main(int argc, char **argv) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("device %s sms=%d\n", prop.name, sms);
    const u64 q = argc > 1 ? strtoull(argv[1], nullptr, 0) : 0x10000000042b0001ULL;   // largest c of the C3 chain
    int k = 63; while (!((q >> k) & 1)) --k;
    s64 cc = (s64)(q - (1ULL << k));
    if (cc > (s64)(1ULL << (k - 1))) { ++k; cc = (s64)q - (s64)(1ULL << k); }
    printf("q = 0x%llx = 2^%d %+lld\n", q, k, cc);
    QC c{q, 2 * q, 4 * q, 0 - q, 0 - 4 * q};
    SolAConst sa{};
    sa.q = q; sa.negc = (int)-cc; sa.sh = (unsigned)(k - 32); sa.mask = (1u << (k - 32)) - 1;
    sa.bias = (u64)(((unsigned __int128)1 << 63) / q) * q;
    if (((u64)1 << 63) - sa.bias > q / 2) sa.bias += q;
    SolConst sc{};
    sc.q = q; sc.negc = (int)-cc; sc.sh = (unsigned)(k - 32); sc.mask = (1u << (k - 32)) - 1; sc.two_s = 1u << (k - 32);
    sc.bias = (u64)(((unsigned __int128)1 << 63) / q) * q;
    std::vector<tw_t> ti(TWN); std::vector<stw_t> ta(TWN), ts(TWN);
    u64 s = 0x9E3779B97F4A7C15ULL;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (int i = 0; i < TWN; ++i) {
        const u64 w = rnd() % q;
        ti[i].x = w; ti[i].y = (u64)(((unsigned __int128)w << 64) / q);
        const u64 v = (u64)(((unsigned __int128)w << 31) % q);
        const s64 wb = w > q / 2 ? (s64)w - (s64)q : (s64)w, vb = v > q / 2 ? (s64)v - (s64)q : (s64)v;
        split_bal(wb, 32, &ta[i].w_lo, &ta[i].w_hi);
        split_bal(vb, 32, &ta[i].v_lo, &ta[i].v_hi);
        const u64 v2 = (u64)(((unsigned __int128)w << 32) % q);
        const s64 v2b = v2 > q / 2 ? (s64)v2 - (s64)q : (s64)v2;
        split_bal(wb, k - 32, &ts[i].w_lo, &ts[i].w_hi);
        split_bal(v2b, k - 32, &ts[i].v_lo, &ts[i].v_hi);
    }
    tw_t *d_ti; stw_t *d_ts, *d_ta;
    cudaMalloc(&d_ti, sizeof(tw_t) * TWN); cudaMalloc(&d_ts, sizeof(stw_t) * TWN); cudaMalloc(&d_ta, sizeof(stw_t) * TWN);
    cudaMemcpy(d_ta, ta.data(), sizeof(stw_t) * TWN, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ti, ti.data(), sizeof(tw_t) * TWN, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ts, ts.data(), sizeof(stw_t) * TWN, cudaMemcpyHostToDevice);
    const size_t words = (size_t)sms * 8 * 4 * 128 * 8;
    std::vector<u64> in(words);
    for (auto &v : in) v = rnd() % q;
    u64 *d; cudaMalloc(&d, words * 8);
    std::vector<u64> r0, r1, r2, r3, r4;
    bench<0, 3>("Shoup/Harvey, add-back csub (round 1)", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &r0);
    bench<1, 3>("fold butterfly, fold every 3 layers", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &r1);
    bench<2, 3>("fold butterfly, recentred, fold every 3 layers", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &r2);
    bench<0, 6>("Shoup/Harvey, 6 layers per iteration", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &r3);
    bench<2, 6>("fold butterfly, recentred, fold every 6 layers", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &r4);
    { std::vector<u64> rx, ry;
      bench<4, 3>("design B, fold every 3 layers, recentred", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &rx);
      size_t bad = 0; for (size_t i = 0; i < rx.size() && i < r0.size(); ++i) bad += rx[i] != r0[i];
      printf("   parity vs Shoup/Harvey: %zu mismatches\n", bad);
      bench<4, 6>("design B, fold every 3 layers, 6 layers/iter", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &ry);
      bad = 0; for (size_t i = 0; i < ry.size() && i < r3.size(); ++i) bad += ry[i] != r3[i];
      printf("   parity vs Shoup/Harvey: %zu mismatches\n", bad);
      { std::vector<u64> r8, s8, t8;
        bench<0, 8>("Shoup/Harvey, 8 layers per iteration", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &r8);
        bench<6, 8>("design B, fold every 4 layers, wide shifter", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &s8);
        bench<7, 8>("design B, fold every 4 layers, ALU shifter", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &t8);
        { std::vector<u64> u8; bench<9, 8>("design B, ALU shifter, every add with 3 inputs", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &u8);
          size_t b3 = 0; for (size_t i = 0; i < r8.size() && i < u8.size(); ++i) b3 += r8[i] != u8[i];
          printf("   parity vs Shoup/Harvey: %zu mismatches\n", b3); }
        { std::vector<u64> u8, v8; bench<10, 8>("design B, ALU shifter, unfused products (5 mul.wide + adds)", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &u8);
          bench<11, 8>("design B, product chain started from U, no recentring", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &v8);
          size_t b3 = 0, b4 = 0; for (size_t i = 0; i < r8.size() && i < u8.size() && i < v8.size(); ++i) { b3 += r8[i] != u8[i]; b4 += r8[i] != v8[i]; }
          printf("   parity vs Shoup/Harvey: unfused %zu, chain %zu mismatches\n", b3, b4);
          std::vector<u64> w8; bench<12, 8>("design B, shifted term as third input of both additions", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &w8);
          size_t b5 = 0; for (size_t i = 0; i < r8.size() && i < w8.size(); ++i) b5 += r8[i] != w8[i];
          printf("   parity vs Shoup/Harvey: three-input %zu mismatches\n", b5); }
        size_t b1 = 0, b2 = 0; for (size_t i = 0; i < r8.size() && i < s8.size() && i < t8.size(); ++i) { b1 += r8[i] != s8[i]; b2 += r8[i] != t8[i]; }
        printf("   parity vs Shoup/Harvey: %zu and %zu mismatches\n", b1, b2); }
      bench<5, 6>("design B, no fold (timing only)", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &rx);
      bench<3, 6>("design B inverse butterfly, sums folded every layer", sms, d, in, d_ti, d_ta, d_ts, c, sa, sc, &rx); }
    if (getenv("SOL_ONLY")) return 0;
    size_t n = std::min(r0.size(), std::min(r1.size(), r2.size())), bad1 = 0, bad2 = 0;
    for (size_t i = 0; i < n; ++i) { bad1 += r0[i] != r1[i]; bad2 += r0[i] != r2[i]; }
    printf("parity vs Shoup/Harvey over %zu words: fold %zu mismatches, recentred %zu mismatches\n", n, bad1, bad2);
    size_t m = std::min(r3.size(), r4.size()), bad3 = 0;
    for (size_t i = 0; i < m; ++i) bad3 += r3[i] != r4[i];
    printf("parity 6 layers: %zu mismatches over %zu words\n", bad3, m);
    {   // host check of the first thread's words against exact arithmetic (3-layer variant)
        u64 x[8]; for (int kk = 0; kk < 8; ++kk) x[kk] = in[kk];
        for (int it = 0; it < 64; ++it) {
            const unsigned P = (it * 7) & 31;
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
        int bad = 0; for (int kk = 0; kk < 8; ++kk) bad += x[kk] != r1[kk];
        printf("host exact check (thread 0): %s\n", bad ? "MISMATCH" : "ok");
    }
    return 0;
}
