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
