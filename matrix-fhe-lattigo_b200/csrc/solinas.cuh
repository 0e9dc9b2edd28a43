// solinas.cuh -- constant-operand modular multiply for primes of the form q = 2^k + c with small |c|
// (what the reference's generators emit: ring/primes.go NTTFriendlyPrimesGenerator walks 2^logQ +- j*NthRoot + 1,
// ring/primes_3n.go:11-43 Find3NRNSPrimes walks 2^logQ + j*3N + 1), on SIGNED lazy words.
//
// Why: the multiply pipe is the unit that saturates in the transforms (profiles/r01_*, r02_*): a 32x32->64 multiply
// occupies it twice as long as a 32-bit one.  The Shoup/Harvey butterfly of modarith.cuh needs 5 wide + 4 narrow
// multiplies, a carry chain for the quotient and a conditional subtraction per butterfly.  For q = 2^k + c no quotient
// is needed: the high part of the product is folded back with 2^k == -c (mod q).
//
//   y = y1*2^32 + y0   with y0 = (int)low word in [-2^31, 2^31), y1 = high word + (low word >> 31)      (one LEA.HI)
//   w = w_hi*2^s + w_lo,  V = w*2^32 mod q = v_hi*2^s + v_lo      (s = k-32; BALANCED residues, lo parts in [-2^(s-1), 2^(s-1)))
//   A = y0*w_lo + y1*v_lo,   B = y0*w_hi + y1*v_hi                (four signed 32x32->64 multiply-adds)
//   y*w == A + 2^s*B = A + 2^s*B.lo + 2^k*B.hi == A + 2^s*B.lo - c*B.hi   (mod q)     (one more wide multiply, two shifts)
//
// Five wide multiplies, no quotient, no carry between A and B, no conditional subtraction: the result lies in about
// (-0.6q, 1.6q) for ANY |y| < 2^62.9, so the butterflies are plain 64-bit additions on signed words and a word is
// brought back ("fold", below) only every four or five layers.  Values are exact field elements; after the final
// canonical reduction the transforms are bit-identical to the reference's MRed-based ones (ring/ntt.go:155-257) --
// checked by the parity tests in tests/.  Range bookkeeping: tools/fold_ranges.py (run by tests/test_fold_ranges.py).
#pragma once
#include <cstdint>

typedef unsigned long long u64;
typedef long long s64;

// twiddle entry {w, V = w*2^32 mod q}: one 128-bit load
struct __align__(16) stw_t { int w_lo, w_hi, v_lo, v_hi; };

__device__ __forceinline__ s64 madwide_s32(int a, int b, s64 c) {
    s64 r; asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
__device__ __forceinline__ s64 pack64(unsigned lo, unsigned hi) { return (s64)(((u64)hi << 32) | lo); }

// per-prime constants (host: fold_setup in b200ring.cu)
struct SolConst {
    u64 q;
    u64 bias;          // the multiple of q next below 2^63: added before the final fold so that the quotient word is >= 0
    int negc;          // -c            (q = 2^k + c)
    unsigned two_s;    // 2^(k-32)
    unsigned sh;       // k - 32
    unsigned mask;     // 2^(k-32) - 1
};
__device__ __forceinline__ s64 mulwide_s32(int a, int b) {
    s64 r; asm("mul.wide.s32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ s64 madwide_us(unsigned a, unsigned b, s64 c) {
    s64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
__device__ __forceinline__ s64 sol_mul(s64 y, const stw_t w, const SolConst &c) {
    const unsigned ylo = (unsigned)y, yhi = (unsigned)((u64)y >> 32);
    const int y0 = (int)ylo, y1 = (int)(yhi + (ylo >> 31));
    s64 A = mulwide_s32(y0, w.w_lo); A = madwide_s32(y1, w.v_lo, A);
    s64 B = mulwide_s32(y0, w.w_hi); B = madwide_s32(y1, w.v_hi, B);
    const s64 T = madwide_us((unsigned)B, c.two_s, A);
    return madwide_s32((int)(B >> 32), c.negc, T);
}
// the same with the shift 2^s*B.lo done by two ALU shifts instead of a wide multiply
__device__ __forceinline__ s64 sol_mul_sh(s64 y, const stw_t w, const SolConst &c) {
    const unsigned ylo = (unsigned)y, yhi = (unsigned)((u64)y >> 32);
    const int y0 = (int)ylo, y1 = (int)(yhi + (ylo >> 31));
    s64 A = mulwide_s32(y0, w.w_lo); A = madwide_s32(y1, w.v_lo, A);
    s64 B = mulwide_s32(y0, w.w_hi); B = madwide_s32(y1, w.v_hi, B);
    const unsigned blo = (unsigned)B;
    const s64 sh = pack64(blo << c.sh, blo >> (32 - c.sh));
    return mulwide_s32((int)(B >> 32), c.negc) + sh + A;
}
// fold: signed word -> [-|c|*2^(63-k), 2^k + |c|*2^(63-k)], same residue
__device__ __forceinline__ s64 sol_fold(s64 v, const SolConst &c) {
    const unsigned lo = (unsigned)v, hi = (unsigned)((u64)v >> 32);
    return madwide_s32((int)hi >> c.sh, c.negc, pack64(lo, hi & c.mask));
}
// the same for an UNSIGNED 64-bit word (any value): result in (-2^(64-k)|c|, 2^k)  for c > 0
__device__ __forceinline__ s64 sol_fold_u(u64 v, const SolConst &c) {
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    return madwide_s32((int)(hi >> c.sh), c.negc, pack64(lo, hi & c.mask));
}
// canonical representative in [0, q) of a signed word with |v| < 2^63 - q.  v + bias is positive, its quotient word
// h = (v + bias) >> k is in [0, 2^(64-k)), so the fold lands in (-2^(64-k)|c|, 2^k + 2^(64-k)|c|): one add-back of q
// when negative and one subtraction when >= q (the latter only for c < 0).
template <bool QBELOW>   // QBELOW: q = 2^k - |c| (the fold can land in [q, 2^k + ..) and needs one more subtraction)
__device__ __forceinline__ u64 sol_canon_t(s64 v, const SolConst &c) {
    s64 r = sol_fold_u((u64)v + c.bias, c);
    unsigned lo = (unsigned)r, hi = (unsigned)((u64)r >> 32);
    asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, 0;\n\t@p add.cc.u32 %0, %0, %2;\n\t@p addc.u32 %1, %1, %3;\n\t}"
        : "+r"(lo), "+r"(hi) : "r"((unsigned)c.q), "r"((unsigned)(c.q >> 32)));
    u64 u = ((u64)hi << 32) | lo;
    if (QBELOW) u = u >= c.q ? u - c.q : u;
    return u;
}
__device__ __forceinline__ u64 sol_canon(s64 v, const SolConst &c) {
    return c.negc > 0 ? sol_canon_t<true>(v, c) : sol_canon_t<false>(v, c);
}
// canonical reduction of N words behind ONE warp-uniform branch on the sign of c (the per-word form above if-converts the
// extra subtraction into compares and selects for every word: 13 instead of 8 instructions per word)
template <int N>
__device__ __forceinline__ void sol_canon_n(u64 (&y)[N], const s64 (&x)[N], const SolConst &c) {
    if (c.negc > 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) y[k] = sol_canon_t<true>(x[k], c);
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) y[k] = sol_canon_t<false>(x[k], c);
    }
}
// v/2 mod q on a signed lazy word: add q when v is odd (q is odd), then an exact arithmetic shift
__device__ __forceinline__ s64 sol_half(s64 v, const SolConst &c) {
    const s64 m = -(v & 1);
    return (v + ((s64)c.q & m)) >> 1;
}
