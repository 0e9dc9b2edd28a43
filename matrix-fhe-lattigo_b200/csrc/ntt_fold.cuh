// ntt_fold.cuh -- the two-pass transforms of ntt_big.cu on the fold arithmetic of solinas.cuh (primes q = 2^k + c).
// Included by ntt_big.cu inside its anonymous namespace: same two-pass split, same twiddle indexing (heap order in the
// column pass, row-pass order in the row pass), same slot permutation; the word type (signed, lazy), the butterflies, the
// reductions and the tiling of the headline column pass (balanced kernels) differ.  Where the folds sit and which layers
// recentre by +-q is restated on intervals in tools/fold_ranges.py (tests/test_fold_ranges.py).
#pragma once
#include "solinas.cuh"

#ifndef B200_FOLD_WIDE_SHIFT
#define B200_FOLD_WIDE_SHIFT 0   // 1: 2^s*B.lo through a sixth wide multiply instead of two shifts (measured slower, tools/solbench)
#endif

__device__ __forceinline__ s64 fmul(s64 y, const stw_t w, const SolConst &c) {
#if B200_FOLD_WIDE_SHIFT
    return sol_mul(y, w, c);
#else
    return sol_mul_sh(y, w, c);
#endif
}

// forward (Cooley-Tukey) butterfly on signed words; KU / KV recentre the two outputs by -q / +q
// (ku, kv are compile-time constants after unrolling)
__device__ __forceinline__ void fbf_fwd(s64 &U, s64 &V, const stw_t w, const SolConst &c, bool ku, bool kv) {
    const s64 t = fmul(V, w, c);
    V = kv ? U - t + (s64)c.q : U - t;
    U = ku ? U + t - (s64)c.q : U + t;
}
// The radix-2 layers of a transform are numbered 0, 1, .. across both passes, starting after the last fold of the
// layers before them (the load, or the radix-3 layer of a 3N ring).  Every FWD_PERIOD layers all words are folded; in
// between, the two outputs are recentred alternately (tools/fold_ranges.py FWD_PATTERN): the window of a folded word,
// [0, 2^k], grows to +-5.6 * 2^k in five layers, inside the +-7.5 * 2^k a multiplied word may have.
constexpr int FWD_PERIOD = 5;
__host__ __device__ constexpr bool fwd_ku(int r) { return (r & 1) == 0; }
__host__ __device__ constexpr bool fwd_kv(int r) { return (r & 1) == 1; }

// where a register group finds the twiddle of level j, block g (TwHeap / TwRow of ntt_big.cu, on the fold tables)
struct TwHeapS {
    const stw_t *__restrict__ tw; unsigned P;
    __device__ __forceinline__ stw_t operator()(int j, int g) const { return tw[(P << j) + g]; }
};
#ifndef B200_TW_SPLIT64
#define B200_TW_SPLIT64 0   // 1: first and second register group of the row pass load a twiddle as two 64-bit loads: measured 2.5 % slower
#endif
// A 128-bit load of a warp costs at least 4 L1 data-pipe wavefronts whatever it touches (quarter-warp granularity,
// tools/l1bench), a 64-bit load one per 128-byte line.  The first two register groups of a row read warp-uniform or
// 8-distinct entries, so two 64-bit loads would be 2 wavefronts against 4 -- but the second load instruction costs more
// than the wavefronts save on B200 (rows_fwd 1.716 -> 1.758 ms), so the single 128-bit load stays.
template <int STRIDE>
struct TwRowS {
    const stw_t *__restrict__ base;
    __device__ __forceinline__ stw_t operator()(int j, int g) const {
        const stw_t *p = base + ((1 << j) - 1 + g) * STRIDE;
        if (B200_TW_SPLIT64 && STRIDE < 32) {
            u64 a, b;
            asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(a) : "l"(p));
            asm volatile("ld.global.nc.u64 %0, [%1+8];" : "=l"(b) : "l"(p));
            stw_t w;
            w.w_lo = (int)(unsigned)a; w.w_hi = (int)(unsigned)(a >> 32); w.v_lo = (int)(unsigned)b; w.v_hi = (int)(unsigned)(b >> 32);
            return w;
        }
        return *p;
    }
};

// LOGM forward layers of a register group whose first layer has number L0
template <int LOGM, int L0, class TW>
__device__ __forceinline__ void fwd_regs_ft(s64 (&x)[1 << LOGM], const TW tw, const SolConst &c) {
#pragma unroll
    for (int j = 0; j < LOGM; ++j) {
        const int half = (1 << LOGM) >> (j + 1);
        const int l = L0 + j, r = l % FWD_PERIOD;
        if (l > 0 && r == 0) {
#pragma unroll
            for (int k = 0; k < (1 << LOGM); ++k) x[k] = sol_fold(x[k], c);
        }
#pragma unroll
        for (int g = 0; g < (1 << j); ++g) {
            const stw_t w = tw(j, g);
#pragma unroll
            for (int o = 0; o < half; ++o) fbf_fwd(x[g * 2 * half + o], x[g * 2 * half + o + half], w, c, fwd_ku(r), fwd_kv(r));
        }
    }
}
template <int LOGM, int L0>
__device__ __forceinline__ void fwd_regs_f(s64 (&x)[1 << LOGM], unsigned P, const stw_t *__restrict__ tw, const SolConst &c) {
    fwd_regs_ft<LOGM, L0>(x, TwHeapS{tw, P}, c);
}
// the same on the row-pass order of the tables (b200ring.cu row_pass_order): base and STRIDE say which group of the row it is
template <int LOGM, int L0, int STRIDE>
__device__ __forceinline__ void fwd_regs_fr(s64 (&x)[1 << LOGM], const stw_t *__restrict__ base, const SolConst &c) {
    fwd_regs_ft<LOGM, L0>(x, TwRowS<STRIDE>{base}, c);
}

// inverse (Gentleman-Sande) butterfly: U' = U + V - KS*q (folded when FOLD), V' = (U - V)*w
template <int KS, bool FOLD>
__device__ __forceinline__ void fbf_inv(s64 &U, s64 &V, const stw_t w, const SolConst &c) {
    const s64 d = U - V;
    s64 s = KS ? U + V - (s64)c.q : U + V;
    if (FOLD) s = sol_fold(s, c);
    U = s;
    V = fmul(d, w, c);
}
// LOGM inverse layers (innermost first); L0 = number of the group's first layer (0 = first layer of the row pass).  The sums
// are recentred by -q in odd layers and folded after them (tools/fold_ranges.py inv_layers).
template <int LOGM, int L0, int JMIN, class TW>
__device__ __forceinline__ void inv_regs_ft(s64 (&x)[1 << LOGM], const TW tw, const SolConst &c) {
#pragma unroll
    for (int j = LOGM - 1; j >= JMIN; --j) {
        const int half = (1 << LOGM) >> (j + 1);
        const int l = L0 + (LOGM - 1 - j);
#pragma unroll
        for (int g = 0; g < (1 << j); ++g) {
            const stw_t w = tw(j, g);
#pragma unroll
            for (int o = 0; o < half; ++o) {
                if (l & 1) fbf_inv<1, true>(x[g * 2 * half + o], x[g * 2 * half + o + half], w, c);
                else fbf_inv<0, false>(x[g * 2 * half + o], x[g * 2 * half + o + half], w, c);
            }
        }
    }
}
template <int LOGM, int L0, int JMIN = 0>
__device__ __forceinline__ void inv_regs_f(s64 (&x)[1 << LOGM], unsigned P, const stw_t *__restrict__ tw, const SolConst &c) {
    inv_regs_ft<LOGM, L0, JMIN>(x, TwHeapS{tw, P}, c);
}
template <int LOGM, int L0, int STRIDE>
__device__ __forceinline__ void inv_regs_fr(s64 (&x)[1 << LOGM], const stw_t *__restrict__ base, const SolConst &c) {
    inv_regs_ft<LOGM, L0, 0>(x, TwRowS<STRIDE>{base}, c);
}

// Twiddles of a row in row-pass order: block twr = table + R + 255*row; [0,7) the first group (warp-uniform), then 7 runs of
// 8 entries (second group, indexed by hi3 = lane & 7), then 6 runs of 32 entries (third group, indexed by lane).
// The 8 layers of one 256-word row held by one warp; same exchanges as row_fwd_core (ntt_big.cu).
// in: c = k*32 + lane, words in (-2^(64-k)|c|, 2^k) (fold_u of anything)   out: c = lane*8 + k, |word| < 5.4 * 2^k
template <int L0>
__device__ __forceinline__ void row_fwd_core_f(s64 (&x)[8], u64 *buf, int lane, const stw_t *__restrict__ twr, const SolConst &c) {
    fwd_regs_fr<3, L0, 1>(x, twr, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) buf[pad16(k * 32 + lane)] = (u64)x[k];
    __syncwarp();
    const int hi3 = lane & 7, lo2 = lane >> 3;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)buf[pad16(hi3 * 32 + k * 4 + lo2)];
    fwd_regs_fr<3, L0 + 3, 8>(x, twr + 7 + hi3, c);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) buf[pad16(hi3 * 32 + k * 4 + lo2)] = (u64)x[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)buf[pad16(lane * 8 + k)];
    s64 (&lo)[4] = *reinterpret_cast<s64 (*)[4]>(&x[0]);
    s64 (&hi)[4] = *reinterpret_cast<s64 (*)[4]>(&x[4]);
    fwd_regs_fr<2, L0 + 6, 32>(lo, twr + 63 + lane, c);
    fwd_regs_fr<2, L0 + 6, 32>(hi, twr + 63 + 96 + lane, c);
    __syncwarp();
}
// in: c = lane*8 + k, words in (-eps, 2^k + eps)   out: c = k*32 + lane; sums folded, products in (-0.6, 1.6) * 2^k
__device__ __forceinline__ void row_inv_core_f(s64 (&x)[8], u64 *buf, int lane, const stw_t *__restrict__ twr, const SolConst &c) {
    {
        s64 (&lo)[4] = *reinterpret_cast<s64 (*)[4]>(&x[0]);
        s64 (&hi)[4] = *reinterpret_cast<s64 (*)[4]>(&x[4]);
        inv_regs_fr<2, 0, 32>(lo, twr + 63 + lane, c);
        inv_regs_fr<2, 0, 32>(hi, twr + 63 + 96 + lane, c);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) buf[pad16(lane * 8 + k)] = (u64)x[k];
    __syncwarp();
    const int hi3 = lane & 7, lo2 = lane >> 3;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)buf[pad16(hi3 * 32 + k * 4 + lo2)];
    inv_regs_fr<3, 2, 8>(x, twr + 7 + hi3, c);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) buf[pad16(hi3 * 32 + k * 4 + lo2)] = (u64)x[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)buf[pad16(k * 32 + lane)];
    inv_regs_fr<3, 5, 1>(x, twr, c);
}

// L0 < 0: the column pass was a Shoup kernel (unsigned words below 8q): fold at the load, layers 0..7.
// L0 >= 0: the column pass was a fold kernel that ran layers 0..L0-1 and left signed words.
template <bool MATRIX, int VT, int L0>
__global__ void __launch_bounds__((MATRIX ? 6 : 1) * VT * 32, MATRIX ? B200_MINB_ROWS_M3 * 2 / VT : B200_MINB_ROWS_STD_FWD * 4 / VT)
k_rows_fwd_f(const u64 *in, u64 *out, BigArgs a) {
    constexpr int RUN = 6 * VT;
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const SolConst sc = a.sol[limb];
    const stw_t *__restrict__ tw = a.fwd_s + (size_t)limb * a.N;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int rowcode;
    const int r = row_of<MATRIX, VT>(a, blockIdx.y, w, rowcode);
    u64 *buf = sm + w * WROW;
    const u64 *src = in + ur.in + (size_t)r * 256;
    const stw_t *__restrict__ twr = tw + a.R + 255 * r;
    s64 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = L0 >= 0 ? (s64)ldd(src + k * 32 + lane) : sol_fold_u(ldd(src + k * 32 + lane), sc);
    row_fwd_core_f<(L0 >= 0 ? L0 : 0)>(x, buf, lane, twr, sc);
    u64 y[8];
    sol_canon_n(y, x, sc);   // canonical also for NTTLazy (a subset of its range)
    if (!MATRIX || (a.flags & B200_TREE_ORDER)) {
        u64 *dst = out + ur.out + (size_t)r * 256 + lane * 8;
        stg256(dst, y[0], y[1], y[2], y[3]);
        stg256(dst + 4, y[4], y[5], y[6], y[7]);
    } else {
        __syncthreads();
        const int rl = (int)rev_bits((unsigned)lane, 5);
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[stage_row<RUN>((int)rev_bits((unsigned)k, 3) * 32 + rl) + rowcode] = y[k];
        __syncthreads();
        const size_t run = (size_t)6 << a.logU;
        u64 *dst = out + ur.out + (size_t)blockIdx.y * RUN;
        {
            const int cr0 = threadIdx.x / RUN, rem = threadIdx.x - cr0 * RUN;
            const u64 *sp = sm + stage_row<RUN>(cr0) + rem;
            u64 *gp = dst + (size_t)cr0 * run + rem;
            const size_t gstep = 32 * run;
#pragma unroll
            for (int i = 0; i < 8; ++i, gp += gstep) std_(gp, sp[i * 32 * (RUN + 1)]);
        }
    }
}

template <bool MATRIX, int VT, bool SIGNED_OUT>
__global__ void __launch_bounds__((MATRIX ? 6 : 1) * VT * 32, MATRIX ? B200_MINB_ROWS_M3 * 2 / VT : B200_MINB_ROWS_STD_INV * 4 / VT)
k_rows_inv_f(const u64 *in, u64 *out, BigArgs a) {
    constexpr int RUN = 6 * VT;
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const SolConst sc = a.sol[limb];
    const stw_t *__restrict__ tw = a.inv_s + (size_t)limb * a.N;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int rowcode;
    const int r = row_of<MATRIX, VT>(a, blockIdx.y, w, rowcode);
    u64 *buf = sm + w * WROW;
    const stw_t *__restrict__ twr = tw + a.R + 255 * r;
    u64 y[8];
    if (!MATRIX || (a.flags & B200_TREE_ORDER)) {
        const u64 *src = in + ur.in + (size_t)r * 256 + lane * 8;
        ldg256(src, y[0], y[1], y[2], y[3]);
        ldg256(src + 4, y[4], y[5], y[6], y[7]);
    } else {
        const size_t run = (size_t)6 << a.logU;
        const u64 *src = in + ur.in + (size_t)blockIdx.y * RUN;
        {
            const int cr0 = threadIdx.x / RUN, rem = threadIdx.x - cr0 * RUN;
            u64 v[8];
            const u64 *gp = src + (size_t)cr0 * run + rem;
            const size_t gstep = 32 * run;
#pragma unroll
            for (int i = 0; i < 8; ++i, gp += gstep) v[i] = ldd(gp);
            u64 *sp = sm + stage_row<RUN>(cr0) + rem;
#pragma unroll
            for (int i = 0; i < 8; ++i) sp[i * 32 * (RUN + 1)] = v[i];
        }
        __syncthreads();
        const int rl = (int)rev_bits((unsigned)lane, 5);
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = sm[stage_row<RUN>((int)rev_bits((unsigned)k, 3) * 32 + rl) + rowcode];
        __syncthreads();
    }
    s64 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = sol_fold_u(y[k], sc);      // any 64-bit input word
    row_inv_core_f(x, buf, lane, twr, sc);
    u64 *dst = out + ur.out + (size_t)r * 256;
    if (SIGNED_OUT) {
#pragma unroll
        for (int k = 0; k < 8; ++k) std_(dst + k * 32 + lane, (u64)x[k]);
    } else {   // the Shoup column pass takes unsigned words below 4q: everything is above -q here
#pragma unroll
        for (int k = 0; k < 8; ++k) std_(dst + k * 32 + lane, (u64)(x[k] + (s64)sc.q));
    }
}

// Fused row stage of the ring product on the fold arithmetic (k_rows_mul of ntt_big.cu): both operands arrive from a fold
// column pass (signed words, layers 0..L0-1 done), run the 8 forward row layers, are brought to [0, q) and multiplied
// (Montgomery product of the two canonical words: x*y*2^-64, half the wide multiplies of a Barrett product; the inverse
// column pass gives the 2^64 back with its scaling constants), and the product runs the 8 inverse row layers; signed words go on to the fold inverse column pass.
template <int L0>
__global__ void __launch_bounds__(128, B200_MINB_ROWS_MUL)
k_rows_mul_f(u64 *ta, const u64 *tb, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const LimbConst lc = a.consts[limb];
    const SolConst sc = a.sol[limb];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = blockIdx.y * 4 + w;
    const stw_t *__restrict__ twr = a.fwd_s + (size_t)limb * a.N + a.R + 255 * r;
    const stw_t *__restrict__ twri = a.inv_s + (size_t)limb * a.N + a.R + 255 * r;
    u64 *buf = sm + w * WROW;
    u64 *pa = ta + ur.in + (size_t)r * 256;
    const u64 *pb = tb + ur.in + (size_t)r * 256;
    s64 x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)pa[k * 32 + lane];
#pragma unroll
    for (int k = 0; k < 8; ++k) y[k] = (s64)pb[k * 32 + lane];
    row_fwd_core_f<L0>(x, buf, lane, twr, sc);
    row_fwd_core_f<L0>(y, buf, lane, twr, sc);
    u64 cx[8], cy[8];
    sol_canon_n(cx, x, sc);
    sol_canon_n(cy, y, sc);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)mred(cx[k], cy[k], lc.q, lc.qinv);   // x*y*2^-64, canonical; the inverse column pass scales by N^-1 * 2^64 (B200_FLAG_MONT_SCALE)
    row_inv_core_f(x, buf, lane, twri, sc);
#pragma unroll
    for (int k = 0; k < 8; ++k) pa[k * 32 + lane] = (u64)x[k];
}

// ====================================================================================================
// column pass on the fold arithmetic: the hand-tuned headline shapes (same tiling as k_cols_*_m3_15 / k_cols_*_std16).
// Inputs of the forward kernels are the caller's words, taken as they are.  Contract: below 2q, what the reference's
// transforms take (ring/ntt.go:155-171) and what the Shoup kernels of ntt_big.cu assume; the schedules hold up to
// 4 * 2^k (3N) and 2.5 * 2^k (2^16), tools/fold_ranges.py.
// ====================================================================================================

// MATRIX, a = 15.  G1: cyclotomic layer, radix-3 layer, fold, radix-2 layer 0; G2: layers 1..3; G3: layers 4, 5 (all
// words folded before layer 5); the row pass continues with layer 6.
template <int CW>
__global__ void __launch_bounds__(CW * 48, B200_MINB_COLS_M3_FWD * 8 / CW)
k_cols_fwd_m3_15_f(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * CW, obase = ur.out + (size_t)blockIdx.x * CW;
    const SolConst sc = a.sol[limb];
    const s64 q = (s64)sc.q;
    const stw_t *__restrict__ tw = a.fwd_s + (size_t)limb * a.N;
    const stw_t om = a.misc_s[limb * MISC_COUNT + MISC_OMEGA3];
    const int t = threadIdx.x, c = t % CW, g = t / CW;   // g in [0,48)
    if (g < 32) {   // G1: rows hd*64 + b*32 + g
        s64 x[6][2];
#pragma unroll
        for (int hd = 0; hd < 6; ++hd)
#pragma unroll
            for (int b = 0; b < 2; ++b) x[hd][b] = (s64)ldd(in + ibase + (size_t)(hd * 64 + b * 32 + g) * 256 + c);
        const stw_t zeta = tw[1];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {   // cyclotomic layer: [0, 4) -> lo' in (-1.6, 4.6), hi' in (-3.6, 6.6)  (units of 2^k)
                const s64 lo = x[d][b], hi = x[d + 3][b];
                const s64 tt = fmul(hi, zeta, sc);
                x[d][b] = lo + tt - q;
                x[d + 3][b] = lo + hi - tt - 2 * q;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {   // radix-3 layer, outputs folded
                const stw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                const s64 b0 = sol_fold(x[3 * h][b], sc);
                const s64 t1 = fmul(x[3 * h + 1][b], z1, sc);
                const s64 t2 = fmul(x[3 * h + 2][b], z2, sc);
                const s64 t3 = fmul(t1 - t2, om, sc);
                x[3 * h][b] = sol_fold(b0 + t1 + t2, sc);
                x[3 * h + 1][b] = sol_fold(b0 - t2 + t3, sc);
                x[3 * h + 2][b] = sol_fold(b0 - t1 - t3, sc);
            }
        }
#pragma unroll
        for (int hd = 0; hd < 6; ++hd) {
            fbf_fwd(x[hd][0], x[hd][1], tw[6 + hd], sc, fwd_ku(0), fwd_kv(0));   // radix-2 layer 0 of block hd
#pragma unroll
            for (int b = 0; b < 2; ++b) sm[pc((hd * 64 + b * 32 + g) * CW + c)] = (u64)x[hd][b];
        }
    }
    __syncthreads();
    s64 x[8];
    {   // G2: g = top*4 + lo2, top = hd*2 + u5 in [0,12); rows top*32 + k*4 + lo2
        const int lo2 = g & 3, top = g >> 2;
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (s64)sm[pc((top * 32 + k * 4 + lo2) * CW + c)];
        fwd_regs_f<3, 1>(x, 12u + top, tw, sc);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[pc((top * 32 + k * 4 + lo2) * CW + c)] = (u64)x[k];
    }
    __syncthreads();
    {   // G3: rows g*8 + k; sub-transform roots at depth 4: 96 + g*2 + {0,1}
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (s64)sm[pc((g * 8 + k) * CW + c)];
        s64 (&lo)[4] = *reinterpret_cast<s64 (*)[4]>(&x[0]);
        s64 (&hi)[4] = *reinterpret_cast<s64 (*)[4]>(&x[4]);
        fwd_regs_f<2, 4>(lo, 96u + g * 2, tw, sc);
        fwd_regs_f<2, 4>(hi, 96u + g * 2 + 1, tw, sc);
#pragma unroll
        for (int k = 0; k < 8; ++k) std_(out + obase + (size_t)(g * 8 + k) * 256 + c, (u64)x[k]);
    }
}
constexpr int M3_15_ROW_L0 = 6;   // the row pass of this shape starts with layer 6

// inverse: the row pass ran layers 0..7 and left signed words (sums folded, products in (-0.6, 1.6)); layers 8..13 here,
// then the inverse radix-3 layer and the inverse cyclotomic layer with the 2/N scaling, canonical output
template <int CW, bool GATHER = false>
__global__ void __launch_bounds__(CW * 48, B200_MINB_COLS_M3_INV * 8 / CW)
k_cols_inv_m3_15_f(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * CW, obase = ur.out + (size_t)blockIdx.x * CW;
    const SolConst sc = a.sol[limb];
    const stw_t *__restrict__ tw = a.inv_s + (size_t)limb * a.N;
    const stw_t *mc = a.misc_s + limb * MISC_COUNT;
    const int t = threadIdx.x, c = t % CW, g = t / CW;
    {
        s64 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (s64)ldd(in + ibase + (size_t)(g * 8 + k) * 256 + c);
        s64 (&lo)[4] = *reinterpret_cast<s64 (*)[4]>(&x[0]);
        s64 (&hi)[4] = *reinterpret_cast<s64 (*)[4]>(&x[4]);
        inv_regs_f<2, 8>(lo, 96u + g * 2, tw, sc);
        inv_regs_f<2, 8>(hi, 96u + g * 2 + 1, tw, sc);
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[pc((g * 8 + k) * CW + c)] = (u64)x[k];
        __syncthreads();
        const int lo2 = g & 3, top = g >> 2;
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (s64)sm[pc((top * 32 + k * 4 + lo2) * CW + c)];
        inv_regs_f<3, 10>(x, 12u + top, tw, sc);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[pc((top * 32 + k * 4 + lo2) * CW + c)] = (u64)x[k];
    }
    __syncthreads();
    if (g < 32) {
        const bool msc = a.flags & B200_FLAG_MONT_SCALE;   // fused ring product: the scaling constants carry the 2^64 of its Montgomery product
        const stw_t om = mc[MISC_OMEGA3], k1 = mc[msc ? MISC_K1M : MISC_K1], k4 = mc[msc ? MISC_K4M : MISC_K4];
        s64 x[6][2];
#pragma unroll
        for (int hd = 0; hd < 6; ++hd) {
#pragma unroll
            for (int b = 0; b < 2; ++b) x[hd][b] = (s64)sm[pc((hd * 64 + b * 32 + g) * CW + c)];
            fbf_inv<1, true>(x[hd][0], x[hd][1], tw[6 + hd], sc);   // layer 13
        }
        u64 y[6][2];
        s64 z[6][2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {   // inverse radix-3 (factor 3 deferred): words in (-0.6, 1.6) -> x0 in (-1.8, 4.8), products
                const stw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                const s64 y0 = x[3 * h][b], y1 = x[3 * h + 1][b], y2 = x[3 * h + 2][b];
                const s64 s = fmul(y1 - y2, om, sc);
                x[3 * h][b] = sol_fold(y0 + y1 + y2, sc);   // folded: lo + hi of the cyclotomic layer must stay a valid multiplicand
                x[3 * h + 1][b] = fmul(y0 - y1 - s, z1, sc);
                x[3 * h + 2][b] = fmul(y0 - y2 + s, z2, sc);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {   // inverse cyclotomic layer with the 2/N scaling: |lo +- hi| < 6.6
                // hi = (lo' - hi')/(2 zeta - 1), lo = (lo' + hi' - hi)/2: two products and an exact halving instead of three products
                const s64 lo = x[d][b], hi = x[d + 3][b];
                const s64 hh = fmul(lo - hi, k1, sc);
                const s64 ll = fmul(lo + hi, k4, sc) - sol_half(hh, sc);
                z[d][b] = ll;
                z[d + 3][b] = hh;
            }
        }
        sol_canon_n(reinterpret_cast<u64 (&)[12]>(y), reinterpret_cast<const s64 (&)[12]>(z), sc);
#pragma unroll
        for (int hd = 0; hd < 6; ++hd)
#pragma unroll
            for (int b = 0; b < 2; ++b) store_out<GATHER>(a.peers, out, obase + (size_t)(hd * 64 + b * 32 + g) * 256 + c, y[hd][b]);
    }
}

// STANDARD, log2(N) = 16: layers 0..7 on the row bits; the row pass continues with layer 8
template <int CW>
__global__ void __launch_bounds__(CW * 32, B200_MINB_COLS_STD * 8 / CW)
k_cols_fwd_std16_f(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * CW, obase = ur.out + (size_t)blockIdx.x * CW;
    const SolConst sc = a.sol[limb];
    const stw_t *__restrict__ tw = a.fwd_s + (size_t)limb * a.N;
    const int t = threadIdx.x, c = t % CW, g = t / CW;   // g in [0,32)
    s64 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)ldd(in + ibase + (size_t)(k * 32 + g) * 256 + c);      // rows k*32 + g
    fwd_regs_f<3, 0>(x, 1u, tw, sc);
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[pc((k * 32 + g) * CW + c)] = (u64)x[k];
    __syncthreads();
    const int lo2 = g & 3, hi3 = g >> 2;                                                 // rows hi3*32 + k*4 + lo2
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)sm[pc((hi3 * 32 + k * 4 + lo2) * CW + c)];
    fwd_regs_f<3, 3>(x, 8u + hi3, tw, sc);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[pc((hi3 * 32 + k * 4 + lo2) * CW + c)] = (u64)x[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)sm[pc((g * 8 + k) * CW + c)];                     // rows g*8 + k
    {
        s64 (&lo)[4] = *reinterpret_cast<s64 (*)[4]>(&x[0]);
        s64 (&hi)[4] = *reinterpret_cast<s64 (*)[4]>(&x[4]);
        fwd_regs_f<2, 6>(lo, 64u + g * 2, tw, sc);
        fwd_regs_f<2, 6>(hi, 64u + g * 2 + 1, tw, sc);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) std_(out + obase + (size_t)(g * 8 + k) * 256 + c, (u64)x[k]);
}
constexpr int STD16_ROW_L0 = 8;

template <int CW, bool GATHER = false>
__global__ void __launch_bounds__(CW * 32, B200_MINB_COLS_STD * 8 / CW)
k_cols_inv_std16_f(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * CW, obase = ur.out + (size_t)blockIdx.x * CW;
    const SolConst sc = a.sol[limb];
    const stw_t *__restrict__ tw = a.inv_s + (size_t)limb * a.N;
    const bool msc = a.flags & B200_FLAG_MONT_SCALE;   // fused ring product: N^-1 carries the 2^64 of its Montgomery product
    const stw_t ninv = a.misc_s[limb * MISC_COUNT + (msc ? MISC_NINVM : MISC_NINV)], ninv_w1 = a.misc_s[limb * MISC_COUNT + (msc ? MISC_NINV_W1M : MISC_NINV_W1)];
    const int t = threadIdx.x, c = t % CW, g = t / CW;
    s64 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)ldd(in + ibase + (size_t)(g * 8 + k) * 256 + c);
    {
        s64 (&lo)[4] = *reinterpret_cast<s64 (*)[4]>(&x[0]);
        s64 (&hi)[4] = *reinterpret_cast<s64 (*)[4]>(&x[4]);
        inv_regs_f<2, 8>(lo, 64u + g * 2, tw, sc);
        inv_regs_f<2, 8>(hi, 64u + g * 2 + 1, tw, sc);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[pc((g * 8 + k) * CW + c)] = (u64)x[k];
    __syncthreads();
    const int lo2 = g & 3, hi3 = g >> 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)sm[pc((hi3 * 32 + k * 4 + lo2) * CW + c)];
    inv_regs_f<3, 10>(x, 8u + hi3, tw, sc);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[pc((hi3 * 32 + k * 4 + lo2) * CW + c)] = (u64)x[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (s64)sm[pc((k * 32 + g) * CW + c)];
    inv_regs_f<3, 13, 1>(x, 1u, tw, sc);   // layers 13, 14
    u64 y[8];
#pragma unroll
    for (int o = 0; o < 4; ++o) {   // layer 15 carries N^-1: |U + V| < 6.4
        const s64 U = x[o], V = x[o + 4];
        x[o] = fmul(U + V, ninv, sc);
        x[o + 4] = fmul(U - V, ninv_w1, sc);
    }
    sol_canon_n(y, x, sc);
#pragma unroll
    for (int k = 0; k < 8; ++k) store_out<GATHER>(a.peers, out, obase + (size_t)(k * 32 + g) * 256 + c, y[k]);
}

// ====================================================================================================
// balanced column pass, MATRIX a = 15.  The tiled kernels above run their first (forward) / last (inverse) register
// group on two thirds of the CTA (32 of 48 row groups) while the rest waits at the barrier: a quarter of all stall
// samples of the forward kernel sit on that one barrier (profiles/r02_*).  Here a CTA is CW*16 threads and every
// thread walks through 4 : 3 : 3 work items of the three phases, so that all warps are busy in all of them:
//   phase A  cyclotomic + radix-3 layer on the 6 words (hd) of one (c, u):           CW*64 items, 4 per thread
//   phase B  radix-2 layers 0..2 (u bits 5..3) on 8 words of one (c, hd, u & 7):     CW*48 items, 3 per thread
//   phase C  radix-2 layers 3..5 (u bits 2..0) on 8 words of one (c, hd, u >> 3):    CW*48 items, 3 per thread
// (inverse: C, B, A with the inverse layers).  The tile lives in shared memory between the phases, indexed
// (row*CW + c) with bits 2-3 XOR-swizzled by the 32-word group so that all three access patterns are conflict-free
// without padding.  Small CTAs (two warps at CW = 4) also mean 16-17 independent CTAs per SM instead of 8.
// ====================================================================================================
#ifndef B200_MINB_COLS_BAL
#define B200_MINB_COLS_BAL 12    // CTAs per SM at CW = 4 (64 threads): 80 registers per thread.  Measured on B200 (cols_fwd / cols_inv, ms):
                                 // 17 CTAs (56 regs, spills) 1.89 / 2.08, 14: 1.61 / 1.94, 12: 1.61 / 1.93, 10: 1.67 / 1.95, 8: 1.67 / 1.95
#endif
#ifndef B200_BAL_UNROLL
#define B200_BAL_UNROLL 1
#endif
#define B200_PRAGMA_(x) _Pragma(#x)
#define B200_UNROLL_(n) B200_PRAGMA_(unroll n)
__device__ __forceinline__ int swz(int i) { return i ^ (((i >> 5) & 3) << 2); }

template <int CW>
__global__ void __launch_bounds__(CW * 16, (B200_MINB_COLS_BAL * 4 / CW) < 1 ? 1 : (B200_MINB_COLS_BAL * 4 / CW))
k_cols_fwd_m3_15_b(const u64 *in, u64 *out, BigArgs a) {
    constexpr int T = CW * 16;
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * CW, obase = ur.out + (size_t)blockIdx.x * CW;
    const SolConst sc = a.sol[limb];
    const s64 q = (s64)sc.q;
    const stw_t *__restrict__ tw = a.fwd_s + (size_t)limb * a.N;
    const int t = threadIdx.x;
    {   // phase A
        const stw_t om = a.misc_s[limb * MISC_COUNT + MISC_OMEGA3];
        const stw_t zeta = tw[1];
        const int c = t % CW, g = t / CW;   // g in [0,16)
B200_UNROLL_(B200_BAL_UNROLL)
        for (int r = 0; r < 4; ++r) {
            const int u = g + 16 * r;
            s64 x[6];
#pragma unroll
            for (int hd = 0; hd < 6; ++hd) x[hd] = (s64)ldd(in + ibase + (size_t)(hd * 64 + u) * 256 + c);
#pragma unroll
            for (int d = 0; d < 3; ++d) {   // cyclotomic layer
                const s64 lo = x[d], hi = x[d + 3];
                const s64 tt = fmul(hi, zeta, sc);
                x[d] = lo + tt - q;
                x[d + 3] = lo + hi - tt - 2 * q;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {   // radix-3 layer, outputs folded
                const stw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                const s64 b0 = sol_fold(x[3 * h], sc);
                const s64 t1 = fmul(x[3 * h + 1], z1, sc);
                const s64 t2 = fmul(x[3 * h + 2], z2, sc);
                const s64 t3 = fmul(t1 - t2, om, sc);
                x[3 * h] = sol_fold(b0 + t1 + t2, sc);
                x[3 * h + 1] = sol_fold(b0 - t2 + t3, sc);
                x[3 * h + 2] = sol_fold(b0 - t1 - t3, sc);
            }
#pragma unroll
            for (int hd = 0; hd < 6; ++hd) sm[swz((hd * 64 + u) * CW + c)] = (u64)x[hd];
        }
    }
    __syncthreads();
B200_UNROLL_(B200_BAL_UNROLL)
    for (int r = 0; r < 3; ++r) {   // phase B: rows hd*64 + k*8 + ulow
        const int it = t + T * r, c = it % CW, rest = it / CW, ulow = rest & 7, hd = rest >> 3;
        s64 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (s64)sm[swz((hd * 64 + k * 8 + ulow) * CW + c)];
        fwd_regs_f<3, 0>(x, 6u + hd, tw, sc);
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[swz((hd * 64 + k * 8 + ulow) * CW + c)] = (u64)x[k];
    }
    __syncthreads();
B200_UNROLL_(B200_BAL_UNROLL)
    for (int r = 0; r < 3; ++r) {   // phase C: rows hd*64 + uhigh*8 + k
        const int it = t + T * r, c = it % CW, rest = it / CW, uhigh = rest & 7, hd = rest >> 3;
        s64 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (s64)sm[swz((hd * 64 + uhigh * 8 + k) * CW + c)];
        fwd_regs_f<3, 3>(x, (6u + hd) * 8 + uhigh, tw, sc);
#pragma unroll
        for (int k = 0; k < 8; ++k) std_(out + obase + (size_t)(hd * 64 + uhigh * 8 + k) * 256 + c, (u64)x[k]);
    }
}

template <int CW, bool GATHER = false>
__global__ void __launch_bounds__(CW * 16, (B200_MINB_COLS_BAL * 4 / CW) < 1 ? 1 : (B200_MINB_COLS_BAL * 4 / CW))
k_cols_inv_m3_15_b(const u64 *in, u64 *out, BigArgs a) {
    constexpr int T = CW * 16;
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * CW, obase = ur.out + (size_t)blockIdx.x * CW;
    const SolConst sc = a.sol[limb];
    const stw_t *__restrict__ tw = a.inv_s + (size_t)limb * a.N;
    const int t = threadIdx.x;
B200_UNROLL_(B200_BAL_UNROLL)
    for (int r = 0; r < 3; ++r) {   // phase C: layers 8..10 on u bits 0..2
        const int it = t + T * r, c = it % CW, rest = it / CW, uhigh = rest & 7, hd = rest >> 3;
        s64 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (s64)ldd(in + ibase + (size_t)(hd * 64 + uhigh * 8 + k) * 256 + c);
        inv_regs_f<3, 8>(x, (6u + hd) * 8 + uhigh, tw, sc);
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[swz((hd * 64 + uhigh * 8 + k) * CW + c)] = (u64)x[k];
    }
    __syncthreads();
B200_UNROLL_(B200_BAL_UNROLL)
    for (int r = 0; r < 3; ++r) {   // phase B: layers 11..13 on u bits 3..5
        const int it = t + T * r, c = it % CW, rest = it / CW, ulow = rest & 7, hd = rest >> 3;
        s64 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (s64)sm[swz((hd * 64 + k * 8 + ulow) * CW + c)];
        inv_regs_f<3, 11>(x, 6u + hd, tw, sc);
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[swz((hd * 64 + k * 8 + ulow) * CW + c)] = (u64)x[k];
    }
    __syncthreads();
    {   // phase A: inverse radix-3 layer (factor 3 deferred) and inverse cyclotomic layer with the 2/N scaling
        const stw_t *mc = a.misc_s + limb * MISC_COUNT;
        const bool msc = a.flags & B200_FLAG_MONT_SCALE;   // fused ring product: the scaling constants carry the 2^64 of its Montgomery product
        const stw_t om = mc[MISC_OMEGA3], k1 = mc[msc ? MISC_K1M : MISC_K1], k4 = mc[msc ? MISC_K4M : MISC_K4];
        const int c = t % CW, g = t / CW;
B200_UNROLL_(B200_BAL_UNROLL)
        for (int r = 0; r < 4; ++r) {
            const int u = g + 16 * r;
            s64 x[6];
#pragma unroll
            for (int hd = 0; hd < 6; ++hd) x[hd] = (s64)sm[swz((hd * 64 + u) * CW + c)];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const stw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                const s64 y0 = x[3 * h], y1 = x[3 * h + 1], y2 = x[3 * h + 2];
                const s64 s = fmul(y1 - y2, om, sc);
                x[3 * h] = sol_fold(y0 + y1 + y2, sc);   // folded: lo + hi of the cyclotomic layer must stay a valid multiplicand
                x[3 * h + 1] = fmul(y0 - y1 - s, z1, sc);
                x[3 * h + 2] = fmul(y0 - y2 + s, z2, sc);
            }
            u64 y[6];
            s64 z[6];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const s64 lo = x[d], hi = x[d + 3];
                const s64 hh = fmul(lo - hi, k1, sc);
                const s64 ll = fmul(lo + hi, k4, sc) - sol_half(hh, sc);
                z[d] = ll;
                z[d + 3] = hh;
            }
            sol_canon_n(y, z, sc);
#pragma unroll
            for (int hd = 0; hd < 6; ++hd) store_out<GATHER>(a.peers, out, obase + (size_t)(hd * 64 + u) * 256 + c, y[hd]);
        }
    }
}

// ====================================================================================================
// generic column pass on the fold arithmetic: the shapes k_cols_std_g / k_cols_m3_g of ntt_big.cu cover (STANDARD
// N = 2^10..2^17, MATRIX N = 3*2^10..3*2^16), same tiling (GCW columns per CTA, register groups of <= 3 layers, padded
// shared-memory tile), signed lazy words.  The radix-2 layers are numbered across both passes as in the headline kernels
// (forward: from 0 after the load / the radix-3 layer, the row pass continues with LOG; inverse: the row pass ran 0..7, the
// column pass starts at 8); the fold schedule depends on the number only.
// ====================================================================================================
template <int S, int ROOTS, int LOG, bool INV, bool FROM_G, bool TO_G, bool LAST_STD, int L0, bool GATHER = false>
__device__ __forceinline__ void col_group_f(const u64 *gin, size_t ibase, u64 *gout, size_t obase, u64 *sm, int b0,
                                            const stw_t *__restrict__ tw, const SolConst &sc, const stw_t ninv, const stw_t ninv_w1,
                                            const PeerOut *po = nullptr) {
    constexpr int M = 1 << S;
    const int ntiles = (GCW * ROOTS) << (LOG - S);
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
        const int c = t & (GCW - 1), rest = t >> GCW_LOG;
        const int lo = rest & ((1 << b0) - 1), hi = rest >> b0;
        const int row0 = (hi << (b0 + S)) + lo;
        s64 x[M];
#pragma unroll
        for (int k = 0; k < M; ++k) {
            const int row = row0 + (k << b0);
            x[k] = (s64)(FROM_G ? ldd(gin + ibase + (size_t)row * 256 + c) : sm[pc(row * GCW + c)]);
        }
        const unsigned P = (unsigned)((ROOTS << (LOG - b0 - S)) + hi);
        u64 y[M];
        if (!INV) {
            fwd_regs_f<S, L0>(x, P, tw, sc);
        } else if (!LAST_STD) {
            inv_regs_f<S, L0>(x, P, tw, sc);
        } else {   // top group of the STANDARD inverse: the last layer carries N^-1, canonical output
            inv_regs_f<S, L0, 1>(x, P, tw, sc);
#pragma unroll
            for (int o = 0; o < M / 2; ++o) {
                const s64 U = x[o], V = x[o + M / 2];
                x[o] = fmul(U + V, ninv, sc);
                x[o + M / 2] = fmul(U - V, ninv_w1, sc);
            }
            sol_canon_n(y, x, sc);
        }
#pragma unroll
        for (int k = 0; k < M; ++k) {
            const int row = row0 + (k << b0);
            const u64 v = (INV && LAST_STD) ? y[k] : (u64)x[k];
            if (TO_G) { if (GATHER) store_out<true>(*po, gout, obase + (size_t)row * 256 + c, v); else std_(gout + obase + (size_t)row * 256 + c, v); }
            else sm[pc(row * GCW + c)] = v;
        }
    }
}

template <int LOGR, int S1, int S2, int S3, bool INV, bool GATHER = false>
__global__ void __launch_bounds__(std_col_threads(LOGR), minb_for(std_col_threads(LOGR), 1024))
k_cols_std_g_f(const u64 *in, u64 *out, BigArgs a) {
    static_assert(S1 + S2 + S3 == LOGR && S1 >= 1 && S2 >= 1, "group sizes");
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * GCW, obase = ur.out + (size_t)blockIdx.x * GCW;
    const SolConst sc = a.sol[limb];
    const stw_t *__restrict__ tw = (INV ? a.inv_s : a.fwd_s) + (size_t)limb * a.N;
    const bool msc = a.flags & B200_FLAG_MONT_SCALE;   // fused ring product: N^-1 carries the 2^64 of its Montgomery product
    const stw_t ninv = a.misc_s[limb * MISC_COUNT + (msc ? MISC_NINVM : MISC_NINV)], ninv_w1 = a.misc_s[limb * MISC_COUNT + (msc ? MISC_NINV_W1M : MISC_NINV_W1)];
    constexpr int S3e = S3 ? S3 : 1;
    if (!INV) {
        col_group_f<S1, 1, LOGR, false, true, false, false, 0>(in, ibase, out, obase, sm, S2 + S3, tw, sc, ninv, ninv_w1);
        __syncthreads();
        if (S3 == 0) {
            col_group_f<S2, 1, LOGR, false, false, true, false, S1>(in, ibase, out, obase, sm, 0, tw, sc, ninv, ninv_w1);
        } else {
            col_group_f<S2, 1, LOGR, false, false, false, false, S1>(in, ibase, out, obase, sm, S3, tw, sc, ninv, ninv_w1);
            __syncthreads();
            col_group_f<S3e, 1, LOGR, false, false, true, false, S1 + S2>(in, ibase, out, obase, sm, 0, tw, sc, ninv, ninv_w1);
        }
    } else {
        if (S3 == 0) {
            col_group_f<S2, 1, LOGR, true, true, false, false, 8>(in, ibase, out, obase, sm, 0, tw, sc, ninv, ninv_w1);
        } else {
            col_group_f<S3e, 1, LOGR, true, true, false, false, 8>(in, ibase, out, obase, sm, 0, tw, sc, ninv, ninv_w1);
            __syncthreads();
            col_group_f<S2, 1, LOGR, true, false, false, false, 8 + S3>(in, ibase, out, obase, sm, S3, tw, sc, ninv, ninv_w1);
        }
        __syncthreads();
        col_group_f<S1, 1, LOGR, true, false, true, true, 8 + S3 + S2, GATHER>(in, ibase, out, obase, sm, S2 + S3, tw, sc, ninv, ninv_w1, &a.peers);
    }
}

// MATRIX: first group = cyclotomic + radix-3 + S1 radix-2 layers on 6*2^S1 words; then groups on the u bits
template <int LOGU, int S1, int S2, int S3, bool INV, bool GATHER = false>
__global__ void __launch_bounds__(m3_col_threads(LOGU), minb_for(m3_col_threads(LOGU), 768))
k_cols_m3_g_f(const u64 *in, u64 *out, BigArgs a) {
    static_assert(S1 + S2 + S3 == LOGU && S2 >= 1, "group sizes");
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * GCW, obase = ur.out + (size_t)blockIdx.x * GCW;
    const SolConst sc = a.sol[limb];
    const s64 q = (s64)sc.q;
    const stw_t *__restrict__ tw = (INV ? a.inv_s : a.fwd_s) + (size_t)limb * a.N;
    const stw_t *mc = a.misc_s + limb * MISC_COUNT;
    const stw_t om = mc[MISC_OMEGA3], none = stw_t{0, 0, 0, 0};
    constexpr int MA = 1 << S1, U = 1 << LOGU, ULO = U >> S1, S3e = S3 ? S3 : 1;
    if (!INV) {
        for (int t = threadIdx.x; t < GCW * ULO; t += blockDim.x) {
            const int c = t & (GCW - 1), ulo = t >> GCW_LOG;
            s64 x[6][MA];
#pragma unroll
            for (int hd = 0; hd < 6; ++hd)
#pragma unroll
                for (int k = 0; k < MA; ++k) x[hd][k] = (s64)ldd(in + ibase + (size_t)(hd * U + k * ULO + ulo) * 256 + c);
            const stw_t zeta = tw[1];
#pragma unroll
            for (int k = 0; k < MA; ++k) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {   // cyclotomic layer
                    const s64 lo = x[d][k], hi = x[d + 3][k];
                    const s64 tt = fmul(hi, zeta, sc);
                    x[d][k] = lo + tt - q;
                    x[d + 3][k] = lo + hi - tt - 2 * q;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {   // radix-3 layer, outputs folded
                    const stw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                    const s64 b0 = sol_fold(x[3 * h][k], sc);
                    const s64 t1 = fmul(x[3 * h + 1][k], z1, sc);
                    const s64 t2 = fmul(x[3 * h + 2][k], z2, sc);
                    const s64 t3 = fmul(t1 - t2, om, sc);
                    x[3 * h][k] = sol_fold(b0 + t1 + t2, sc);
                    x[3 * h + 1][k] = sol_fold(b0 - t2 + t3, sc);
                    x[3 * h + 2][k] = sol_fold(b0 - t1 - t3, sc);
                }
            }
#pragma unroll
            for (int hd = 0; hd < 6; ++hd) {
                if (S1 > 0) fwd_regs_f<S1, 0>(x[hd], (unsigned)(6 + hd), tw, sc);
#pragma unroll
                for (int k = 0; k < MA; ++k) sm[pc((hd * U + k * ULO + ulo) * GCW + c)] = (u64)x[hd][k];
            }
        }
        __syncthreads();
        if (S3 == 0) {
            col_group_f<S2, 6, LOGU, false, false, true, false, S1>(in, ibase, out, obase, sm, 0, tw, sc, none, none);
        } else {
            col_group_f<S2, 6, LOGU, false, false, false, false, S1>(in, ibase, out, obase, sm, S3, tw, sc, none, none);
            __syncthreads();
            col_group_f<S3e, 6, LOGU, false, false, true, false, S1 + S2>(in, ibase, out, obase, sm, 0, tw, sc, none, none);
        }
    } else {
        if (S3 == 0) {
            col_group_f<S2, 6, LOGU, true, true, false, false, 8>(in, ibase, out, obase, sm, 0, tw, sc, none, none);
        } else {
            col_group_f<S3e, 6, LOGU, true, true, false, false, 8>(in, ibase, out, obase, sm, 0, tw, sc, none, none);
            __syncthreads();
            col_group_f<S2, 6, LOGU, true, false, false, false, 8 + S3>(in, ibase, out, obase, sm, S3, tw, sc, none, none);
        }
        __syncthreads();
        const bool msc = a.flags & B200_FLAG_MONT_SCALE;
        const stw_t k1 = mc[msc ? MISC_K1M : MISC_K1], k4 = mc[msc ? MISC_K4M : MISC_K4];
        for (int t = threadIdx.x; t < GCW * ULO; t += blockDim.x) {
            const int c = t & (GCW - 1), ulo = t >> GCW_LOG;
            s64 x[6][MA];
#pragma unroll
            for (int hd = 0; hd < 6; ++hd) {
#pragma unroll
                for (int k = 0; k < MA; ++k) x[hd][k] = (s64)sm[pc((hd * U + k * ULO + ulo) * GCW + c)];
                if (S1 > 0) inv_regs_f<S1, 8 + S3 + S2>(x[hd], (unsigned)(6 + hd), tw, sc);
            }
#pragma unroll
            for (int k = 0; k < MA; ++k) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {   // inverse radix-3 (factor 3 deferred), x0 folded
                    const stw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                    const s64 y0 = x[3 * h][k], y1 = x[3 * h + 1][k], y2 = x[3 * h + 2][k];
                    const s64 sg = fmul(y1 - y2, om, sc);
                    // the last radix-2 layer is number LOGU + 7: when that is even its sums arrive unfolded (up to 3.2 * 2^k)
                    // and three of them would pass 2^63: one of them is folded first (tools/fold_ranges.py inverse_3n)
                    x[3 * h][k] = sol_fold((((LOGU + 7) & 1) ? y0 : sol_fold(y0, sc)) + y1 + y2, sc);
                    x[3 * h + 1][k] = fmul(y0 - y1 - sg, z1, sc);
                    x[3 * h + 2][k] = fmul(y0 - y2 + sg, z2, sc);
                }
                s64 z[6];
                u64 y[6];
#pragma unroll
                for (int d = 0; d < 3; ++d) {   // inverse cyclotomic layer, two products and a halving
                    const s64 lo = x[d][k], hi = x[d + 3][k];
                    const s64 hh = fmul(lo - hi, k1, sc);
                    z[d] = fmul(lo + hi, k4, sc) - sol_half(hh, sc);
                    z[d + 3] = hh;
                }
                sol_canon_n(y, z, sc);
#pragma unroll
                for (int hd = 0; hd < 6; ++hd) store_out<GATHER>(a.peers, out, obase + (size_t)(hd * U + k * ULO + ulo) * 256 + c, y[hd]);
            }
        }
    }
}
