// solinas.cuh -- constant-operand modular multiply for primes of the form q = 2^k + c with small |c|
// (every prime the reference's generators emit: ring/primes.go NTTFriendlyPrimesGenerator walks 2^logQ +- j*NthRoot + 1,
// ring/primes_3n.go:11-43 Find3NRNSPrimes walks 2^logQ + j*3N + 1), on SIGNED lazy words.
//
// Why: the Shoup/Harvey butterfly of modarith.cuh needs 5 wide + 4 narrow 32-bit multiplies, a carry chain for the
// quotient and a conditional subtraction per butterfly (about 24 instructions, 14 of them on the multiply pipe, which
// is the unit that saturates: profiles/r01_*).  For q = 2^k + c the quotient is not needed at all:
//
//   y = y1*2^31 + y0  (y0 = low 31 bits, y1 = y >> 31 arithmetic, |y| < 2^62)
//   S = y0*w + y1*W1  with W1 = w*2^31 mod q,  both twiddle words stored as BALANCED residues in (-q/2, q/2]
//     = A + 2^32*B   (four signed 32x32 multiply-adds on 32-bit halves of w and W1; |S| < 2^31*q)
//   S = Sh*2^k + Sl  (one funnel shift, one mask)          2^k = q - c   =>   S == Sl - Sh*c  (mod q)
//   T = Sl - Sh*c    (one signed wide multiply-add)        T in [-2^31|c|, 2^k + 2^31|c|]
//
// i.e. 5 wide multiply-adds + 2 + 2 + 1 single-word ALU/move instructions and NO conditional subtraction: the result is
// within about one q of zero on either side whatever the size of y, so butterflies are plain 64-bit adds and
// subtractions on signed words and a word only has to be brought back ("fold", 3 instructions) every few layers.
// Results are exact field elements; after the final canonical reduction they are bit-identical to the reference's
// MRed-based transform (ring/ntt.go:155-257) -- checked against the oracle in tests/.
//
// The low product A is kept unsigned by starting its multiply-add chain from BIAS, a multiple of q next to 2^63, so
// that its high word can be fed to the B chain as a plain (hi, 0) register pair.
#pragma once
#include <cstdint>

typedef unsigned long long u64;
typedef long long s64;

// per-prime constants of the fold arithmetic (host: solinas_setup in b200ring.cu)
struct SolAConst {
    u64 q;
    u64 bias;        // m*q next to 2^63 (added to the low product)
    int negc;        // -c            (q = 2^k + c)
    unsigned sh;     // k - 32
    unsigned mask;   // 2^(k-32) - 1
    unsigned pad_;
};

// twiddle entry: {w, W1 = w*2^31 mod q} as balanced residues, each split into (lo: signed 32 bits, hi) with
// value = hi*2^32 + lo, lo in [-2^31, 2^31); packed little-endian so that one 128-bit load brings w_lo, w_hi, W1_lo, W1_hi
struct __align__(16) stw_t { int w_lo, w_hi, v_lo, v_hi; };


__device__ __forceinline__ s64 madwide_s32(int a, int b, s64 c) {
    s64 r; asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
__device__ __forceinline__ unsigned funnel_r(unsigned lo, unsigned hi, unsigned s) {
    unsigned r; asm("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(s)); return r;
}
__device__ __forceinline__ s64 pack64(unsigned lo, unsigned hi) { return (s64)(((u64)hi << 32) | lo); }

// y*w mod q as a signed word in [-2^31|c|, 2^k + 2^31|c|]; requires |y| <= the prime's y_max (2^62 - 2^40 or 2^61)
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
// fold: any signed word -> [-|c|*2^(63-k), 2^k + |c|*2^(63-k)], same residue
__device__ __forceinline__ s64 sola_fold(s64 v, const SolAConst &c) {
    const unsigned lo = (unsigned)v, hi = (unsigned)((u64)v >> 32);
    const int h = (int)hi >> c.sh;
    return madwide_s32(h, c.negc, pack64(lo, hi & c.mask));
}
// canonical representative in [0, q) of a folded word (|v| < 2q)
__device__ __forceinline__ u64 sola_canon(s64 v, const SolAConst &c) {
    s64 r = sola_fold(v, c);
    if (r < 0) r += (s64)c.q;
    if (r >= (s64)c.q) r -= (s64)c.q;
    return (u64)r;
}

// ---- design B: twiddle words split at bit k-32, no quotient, no carry between the partial sums ---------------------
//   y = y1*2^32 + y0 with y0 = (int)low word in [-2^31, 2^31), y1 = high word + (low word >> 31)        (one LEA.HI)
//   w = w_hi*2^s + w_lo, V = w*2^32 mod q = v_hi*2^s + v_lo  (s = k-32; balanced residues, lo parts in [-2^(s-1), 2^(s-1)))
//   A = y0*w_lo + y1*v_lo,  B = y0*w_hi + y1*v_hi,  y*w == A + 2^s*B = A + 2^s*B.lo + 2^k*B.hi == A + 2^s*B.lo - c*B.hi
// six wide multiply-adds, one address-class instruction for the split, no shift, no mask, no carry between A and B.
// |y| < 2^62  =>  result in (-0.48q, 1.48q) for k = 60, c < 2^27.
struct SolConst {
    u64 q;
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
// fold: signed word -> [-|c|*2^(63-k), 2^k + |c|*2^(63-k)], same residue
__device__ __forceinline__ s64 sol_fold(s64 v, const SolConst &c) {
    const unsigned lo = (unsigned)v, hi = (unsigned)((u64)v >> 32);
    return madwide_s32((int)hi >> c.sh, c.negc, pack64(lo, hi & c.mask));
}
__device__ __forceinline__ u64 sol_canon(s64 v, const SolConst &c) {
    s64 r = sol_fold(v, c);
    if (r < 0) r += (s64)c.q;
    if (r >= (s64)c.q) r -= (s64)c.q;
    return (u64)r;
}
