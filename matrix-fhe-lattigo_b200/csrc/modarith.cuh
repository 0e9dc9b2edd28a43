// modarith.cuh -- 64-bit modular arithmetic on the sm_100a integer pipes.
//
// Device restatement of the bit-level primitives of the reference's ring/modular_reduction.go
// (R = 2^64; hi/lo halves of the 128-bit product == math/bits.Mul64 == __umul64hi / *), plus the
// Shoup/Harvey constant-operand multiply the transform kernels use for twiddles (not in the
// reference: it keeps twiddles in Montgomery form; canonical outputs are identical).
#pragma once
#include <cstdint>

typedef unsigned long long u64;   // same width as uint64_t; the type __umul64hi takes

// Per-prime constants (ring.SubRing: Modulus, MRedConstant, BRedConstant; subring.go:35-55)
struct LimbConst {
    u64 q;        // Modulus
    u64 qinv;     // MRedConstant = q^-1 mod 2^64            (modular_reduction.go:68-75)
    u64 brc_hi;   // BRedConstant[0] = floor(2^128/q) >> 64   (modular_reduction.go:99-107)
    u64 brc_lo;   // BRedConstant[1]
    unsigned canon_shift;   // quotient estimate for x < 16q:  k = ((x >> canon_shift) * canon_mul) >> 32
    unsigned canon_mul;     //   with floor(x/q) - 1 <= k <= floor(x/q)   (not in the reference; see canon_est)
    u64 pad_;
};

__device__ __forceinline__ u64 csub(u64 x, u64 m) { return x >= m ? x - m : x; }   // CRed :200-205

// MRed :78-86   x*y*2^-64 mod q, canonical
__device__ __forceinline__ u64 mred(u64 x, u64 y, u64 q, u64 qinv) {
    u64 mhi = __umul64hi(x, y), mlo = x * y;
    u64 r = mhi - __umul64hi(mlo * qinv, q) + q;
    return r >= q ? r - q : r;
}
// MRedLazy :90-95   same, in [0, 2q-1]
__device__ __forceinline__ u64 mred_lazy(u64 x, u64 y, u64 q, u64 qinv) {
    u64 mhi = __umul64hi(x, y), mlo = x * y;
    return mhi - __umul64hi(mlo * qinv, q) + q;
}
// BRedLazy :166-197   x*y mod q in [0, 2q-1]; the carry chain of the four partial products
__device__ __forceinline__ u64 bred_lazy(u64 x, u64 y, u64 q, u64 uhi, u64 ulo) {
    u64 mhi = __umul64hi(x, y), mlo = x * y;
    u64 r = mhi * uhi;
    u64 hlo = mlo * uhi;
    r += __umul64hi(mlo, uhi);
    u64 s0 = hlo + __umul64hi(mlo, ulo);
    r += (s0 < hlo);
    u64 h2 = mhi * ulo;
    r += __umul64hi(mhi, ulo);
    r += ((h2 + s0) < h2);
    return mlo - r * q;
}
// BRed :127-162
__device__ __forceinline__ u64 bred(u64 x, u64 y, u64 q, u64 uhi, u64 ulo) {
    return csub(bred_lazy(x, y, q, uhi, ulo), q);
}
// BRedAddLazy :121-124 / BRedAdd :110-117   a mod q for any 64-bit a
__device__ __forceinline__ u64 bred_add_lazy(u64 a, u64 q, u64 uhi) { return a - __umul64hi(a, uhi) * q; }
__device__ __forceinline__ u64 bred_add(u64 a, u64 q, u64 uhi) { return csub(bred_add_lazy(a, q, uhi), q); }
// MFormLazy :40-45 / MForm :11-35   a*2^64 mod q
__device__ __forceinline__ u64 mform_lazy(u64 a, u64 q, u64 uhi, u64 ulo) {
    return (0ULL - (a * uhi + __umul64hi(a, ulo))) * q;
}
__device__ __forceinline__ u64 mform(u64 a, u64 q, u64 uhi, u64 ulo) { return csub(mform_lazy(a, q, uhi, ulo), q); }
// IMForm :49-56   a*2^-64 mod q
__device__ __forceinline__ u64 imform(u64 a, u64 q, u64 qinv) { return csub(q - __umul64hi(a * qinv, q), q); }

// Shoup multiply by a constant w with companion ws = floor(w*2^64/q): x*w mod q in [0, 2q) for ANY
// 64-bit x (Harvey, "Faster arithmetic for number-theoretic transforms", 2014).
__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {
    return x * w - __umul64hi(x, ws) * q;
}
__device__ __forceinline__ u64 shoup(u64 x, u64 w, u64 ws, u64 q) { return csub(shoup_lazy(x, w, ws, q), q); }

// ---- hand-scheduled variants for the tuned kernels (ntt_big.cu) -----------------------------------------
// On sm_100a a 64-bit multiply is emulated on the 32-bit integer pipes: IMAD.WIDE (32x32->64, two FMA-pipe
// slots), IMAD (one slot) and IADD3 (ALU pipe, half rate).  The ALU pipe is the binding resource of the
// butterflies, so these forms trade ALU work for FMA-pipe work and drop what laziness allows.
__device__ __forceinline__ u64 mulwide_u32(unsigned a, unsigned b) {
    u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ u64 madwide_u32(unsigned a, unsigned b, u64 c) {
    u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
__device__ __forceinline__ unsigned madlo_u32(unsigned a, unsigned b, unsigned c) {
    unsigned r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ unsigned lo32(u64 x) { return (unsigned)x; }
__device__ __forceinline__ unsigned hi32(u64 x) { return (unsigned)(x >> 32); }

// x*w mod q in [0, 4q) for ANY 64-bit x.  The quotient estimate t drops the low partial product and the
// low halves of the cross products of x*ws, so it is at most 2 below floor(x*ws/2^64); the remainder is
// accumulated as x*w + t*(2^64 - q) on the low 64 bits (2 wide + 4 narrow multiply-adds, no subtraction).
#ifndef B200_SHOUP_MULHI
#define B200_SHOUP_MULHI 0
#endif
// B200_SHOUP_ALU_SUM: sum the three quotient terms with explicit add-with-carry instructions.  Otherwise ptxas folds
// one high half into the 64-bit addend of a wide multiply and pays two moves to build that register pair.
#ifndef B200_SHOUP_ALU_SUM
#define B200_SHOUP_ALU_SUM 0
#endif
__device__ __forceinline__ u64 shoup4(u64 x, u64 w, u64 ws, u64 nq) {
    const unsigned x0 = lo32(x), x1 = hi32(x);
#if B200_SHOUP_ALU_SUM
    const u64 C = mulwide_u32(x1, hi32(ws));
    const unsigned a = hi32(mulwide_u32(x0, hi32(ws))), b = hi32(mulwide_u32(x1, lo32(ws)));
    unsigned t0, t1;
    asm("{\n\t.reg .u32 s, k;\n\tadd.cc.u32 s, %2, %3;\n\taddc.u32 k, 0, 0;\n\tadd.cc.u32 %0, %4, s;\n\taddc.u32 %1, %5, k;\n\t}"
        : "=r"(t0), "=r"(t1) : "r"(a), "r"(b), "r"(lo32(C)), "r"(hi32(C)));
#else
#if B200_SHOUP_MULHI
    const u64 t = mulwide_u32(x1, hi32(ws)) + __umulhi(x0, hi32(ws)) + __umulhi(x1, lo32(ws));
#else
    const u64 t = mulwide_u32(x1, hi32(ws)) + hi32(mulwide_u32(x0, hi32(ws))) + hi32(mulwide_u32(x1, lo32(ws)));
#endif
    const unsigned t0 = lo32(t), t1 = hi32(t);
#endif
    const u64 r = madwide_u32(t0, lo32(nq), mulwide_u32(x0, lo32(w)));
    unsigned r1 = hi32(r);
    r1 = madlo_u32(x0, hi32(w), r1);
    r1 = madlo_u32(x1, lo32(w), r1);
    r1 = madlo_u32(t0, hi32(nq), r1);
    r1 = madlo_u32(t1, lo32(nq), r1);
    return ((u64)r1 << 32) | lo32(r);
}
// x holds a two's-complement value in [-m, m): add m back if it is negative.  One sign test on the high word and two
// predicated adds (3 instructions) instead of the 64-bit compare, two selects and the subtraction of csub (6).
__device__ __forceinline__ u64 addback_neg(u64 x, u64 m) {
    unsigned lo = lo32(x), hi = hi32(x);
    asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, 0;\n\t@p add.cc.u32 %0, %0, %2;\n\t@p addc.u32 %1, %1, %3;\n\t}"
        : "+r"(lo), "+r"(hi) : "r"(lo32(m)), "r"(hi32(m)));
    return ((u64)hi << 32) | lo;
}

// x*w mod q in [0, 2q) for ANY 64-bit x at the price of the approximate quotient: the multiply-add chain of shoup4 starts
// from -2q instead of 0 (free), which leaves the value of shoup4 minus 2q in [-2q, 2q), and the sign-based add-back
// (3 instructions) brings it to [0, 2q).  Same residue as the exact Shoup product, possibly the other representative.
__device__ __forceinline__ u64 shoup2q(u64 x, u64 w, u64 ws, u64 nq, u64 twoq) {
    const unsigned x0 = lo32(x), x1 = hi32(x);
    const u64 t = mulwide_u32(x1, hi32(ws)) + hi32(mulwide_u32(x0, hi32(ws))) + hi32(mulwide_u32(x1, lo32(ws)));
    const unsigned t0 = lo32(t), t1 = hi32(t);
    const u64 r = madwide_u32(t0, lo32(nq), madwide_u32(x0, lo32(w), 0 - twoq));
    unsigned r1 = hi32(r);
    r1 = madlo_u32(x0, hi32(w), r1);
    r1 = madlo_u32(x1, lo32(w), r1);
    r1 = madlo_u32(t0, hi32(nq), r1);
    r1 = madlo_u32(t1, lo32(nq), r1);
    return addback_neg(((u64)r1 << 32) | lo32(r), twoq);
}
// x >= m ? x - m : x with the subtraction done as a multiply-add by nm = 2^64 - m on the FMA pipe:
// 3 ALU (compare + select) + 2 FMA instructions instead of 6 ALU.
__device__ __forceinline__ u64 csub_fma(u64 x, u64 m, u64 nm) {
    const unsigned k = x >= m ? 1u : 0u;
    const u64 r = madwide_u32(k, lo32(nm), x);
    return ((u64)madlo_u32(k, hi32(nm), hi32(r)) << 32) | lo32(r);
}

// hi-word-only variant of csub_fma (2 ALU + 2 FMA): subtracts when hi32(x) > hi32(m).  The result is
// < m + 2^32 for x < 2m + 2^32, which the lazy butterflies tolerate when 8q + 2^33 <= 2^64.
__device__ __forceinline__ u64 csub_fma_hi(u64 x, u64 m, u64 nm) {
    unsigned k;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(k) : "r"(hi32(x)), "r"(hi32(m)));
    const u64 r = madwide_u32(k, lo32(nm), x);
    return ((u64)madlo_u32(k, hi32(nm), hi32(r)) << 32) | lo32(r);
}
// x < 16q  ->  [0, q) through a 32-bit quotient estimate (1 shift + 1 mul.hi + 2 multiply-adds) and one
// conditional subtraction, instead of a chain of three or four conditional subtractions.
// FINAL_ALU: the last conditional subtraction as compare/select/subtract (ALU pipe) instead of a multiply-add.
// FINAL: last conditional subtraction as 0 = multiply-add, 1 = compare/select/subtract, 2 = one more q subtracted up front
// (the +1 rides on the mad.hi that computes the estimate) and added back if the result is negative (sign test + two
// predicated adds): 7 instructions for the whole reduction.
template <int FINAL = 0>
__device__ __forceinline__ u64 canon_est(u64 x, u64 q, u64 nq, unsigned sh, unsigned mul) {
    unsigned k;
    if (FINAL == 2) asm("mad.hi.u32 %0, %1, %2, 1;" : "=r"(k) : "r"((unsigned)(x >> sh)), "r"(mul));
    else k = __umulhi((unsigned)(x >> sh), mul);
    const u64 r = madwide_u32(k, lo32(nq), x);
    const u64 y = ((u64)madlo_u32(k, hi32(nq), hi32(r)) << 32) | lo32(r);
    return FINAL == 3 ? y : FINAL == 2 ? addback_neg(y, q) : FINAL == 1 ? csub(y, q) : csub_fma(y, q, nq);   // 3: lazy, [0, 2q)
}
