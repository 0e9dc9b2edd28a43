// hostmath.hpp -- host-side number theory for context creation (one-time, per prime).
//
// Produces what the reference's ring/subring.go:129-214 (generateNTTConstants), :218-251
// (PrimitiveRoot), :255-290 (Find3NPrimitiveRoot), ring/modular_reduction.go:68-75,99-107
// (GenMRedConstant / GenBRedConstant) and ring/primes.go:11-13 (IsPrime) produce, so that the
// uploaded tables are bit-identical to a Go ring built from the same moduli.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace hostmath {

typedef unsigned long long u64;   // same type as the device code (and __umul64hi)
typedef unsigned __int128 u128;

inline u64 mulhi(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }
inline u64 mulmod(u64 a, u64 b, u64 m) { return (u64)(((u128)a * b) % m); }
inline u64 addmod(u64 a, u64 b, u64 m) { u64 s = a + b; return (s >= m || s < a) ? s - m : s; }
inline u64 submod(u64 a, u64 b, u64 m) { return a >= b ? a - b : a + m - b; }
inline u64 powmod(u64 b, u64 e, u64 m) {
    u64 r = 1 % m;
    b %= m;
    for (; e; e >>= 1) {
        if (e & 1) r = mulmod(r, b, m);
        b = mulmod(b, b, m);
    }
    return r;
}
inline u64 invmod_prime(u64 a, u64 p) { return powmod(a, p - 2, p); }

// Miller-Rabin with the first 12 prime bases: deterministic below 3.3e24, so for every u64 it gives
// the verdict of big.Int.ProbablyPrime(0) (Baillie-PSW, exact below 2^64) used by ring.IsPrime.
inline bool is_prime(u64 n) {
    static const u64 bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return false;
    for (u64 p : bases)
        if (n % p == 0) return n == p;
    u64 d = n - 1;
    int s = 0;
    while ((d & 1) == 0) d >>= 1, ++s;
    for (u64 a : bases) {
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int r = 1; r < s && witness; ++r) {
            x = mulmod(x, x, n);
            if (x == n - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

inline u64 gcd(u64 a, u64 b) {
    while (b) { u64 t = a % b; a = b; b = t; }
    return a;
}

// Brent's variant of Pollard rho; n composite and odd.
inline u64 rho_brent(u64 n) {
    for (u64 c = 1;; ++c) {
        u64 y = 2, x = 2, ys = 2, g = 1, qq = 1;
        const u64 m = 128;
        for (u64 r = 1; g == 1; r <<= 1) {
            x = y;
            for (u64 i = 0; i < r; ++i) y = addmod(mulmod(y, y, n), c, n);
            for (u64 k = 0; k < r && g == 1; k += m) {
                ys = y;
                for (u64 i = 0; i < std::min(m, r - k); ++i) {
                    y = addmod(mulmod(y, y, n), c, n);
                    qq = mulmod(qq, x > y ? x - y : y - x, n);
                }
                g = gcd(qq, n);
            }
        }
        if (g == n) {
            g = 1;
            while (g == 1) {
                ys = addmod(mulmod(ys, ys, n), c, n);
                g = gcd(x > ys ? x - ys : ys - x, n);
            }
        }
        if (g != n) return g;
    }
}

inline void split(u64 n, std::vector<u64> &out) {
    if (n == 1) return;
    if (is_prime(n)) { out.push_back(n); return; }
    u64 d = (n % 2 == 0) ? 2 : rho_brent(n);
    split(d, out);
    split(n / d, out);
}

// Unique prime factors of m, ascending (the set utils/factorization.GetFactors returns).
inline std::vector<u64> unique_factors(u64 m) {
    std::vector<u64> f;
    for (u64 p = 2; p < 1024 && p * p <= m; ++p)
        if (m % p == 0) {
            f.push_back(p);
            while (m % p == 0) m /= p;
        }
    split(m, f);
    std::sort(f.begin(), f.end());
    f.erase(std::unique(f.begin(), f.end()), f.end());
    return f;
}

inline bool is_primitive_root(u64 g, u64 q, const std::vector<u64> &factors) {
    for (u64 f : factors)
        if (powmod(g, (q - 1) / f, q) == 1) return false;
    return true;
}

// Smallest primitive root >= 3 (the reference never tries 2: subring.go:236-239).
inline u64 primitive_root(u64 q, const std::vector<u64> &factors) {
    for (u64 g = 3;; ++g)
        if (is_primitive_root(g, q, factors)) return g;
}

inline u64 mred_constant(u64 q) {   // q^-1 mod 2^64 by Newton iteration (== q^(2^63-1))
    u64 x = q;                      // correct to 3 bits for odd q
    for (int i = 0; i < 6; ++i) x *= 2 - q * x;
    return x;
}
inline void bred_constant(u64 q, u64 out[2]) {   // floor(2^128 / q), q odd > 1
    u128 r = ~(u128)0 / q;
    out[0] = (u64)(r >> 64);
    out[1] = (u64)r;
}
inline u64 mform(u64 a, u64 q) { return (u64)((((u128)a) << 64) % q); }              // a*2^64 mod q
inline u64 mred(u64 x, u64 y, u64 q, u64 qinv) {                                    // x*y*2^-64 mod q
    u128 m = (u128)x * y;
    u64 r = (u64)(m >> 64) - mulhi((u64)m * qinv, q) + q;
    return r >= q ? r - q : r;
}
inline u64 shoup(u64 w, u64 q) { return (u64)((((u128)w) << 64) / q); }             // floor(w*2^64/q)

inline u64 bitrev(u64 x, int bits) {
    u64 r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

}  // namespace hostmath
