// primes.cu -- host-side modulus and root-of-unity search of the 3N / power-of-two ring constructors.
//
// Restates the behaviour of ring/primes_3n.go (Find3NRNSPrimes, OrderMod, HasPrimitiveMthRoot,
// FindPrimitiveRootOfUnity, Derive3NRoots), ring/primes.go (IsPrime, Find3NFriendlyPrime,
// NTTFriendlyPrimesGenerator) and ring/subring.go:218-290 (PrimitiveRoot, Find3NPrimitiveRoot) behind the C ABI, so a
// caller can build a ring (b200_ctx_create) from (N, bit size, count) exactly as core/rlwe/params.go:924,966,998 and
// core/rlwe/params_3n.go:804-805 do.  No device work: these run without a GPU.  Factoring uses Pollard rho instead of
// the reference's trial division / big.Int arithmetic; the results are the same integers.
#include <cmath>
#include <string>
#include <vector>

#include "../../include/b200ring.h"
#include "hostmath.hpp"
#include "internal.hpp"

namespace hm = hostmath;
typedef hm::u64 u64;

static int fail(int code, const std::string &msg) {
    b200_set_error(msg);
    return code;
}

static u64 splitmix64(u64 &s) {
    u64 z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// multiplicative order of g modulo the prime p (ring/primes_3n.go:78-91)
static u64 order_mod(u64 g, u64 p) {
    u64 order = p - 1;
    for (u64 f : hm::unique_factors(p - 1))
        while (order % f == 0 && hm::powmod(g, order / f, p) == 1) order /= f;
    return order;
}

extern "C" int b200_is_prime(uint64_t x) { return hm::is_prime(x) ? 1 : 0; }

extern "C" int b200_find_3n_rns_primes(int N, int bitSize, int count, int searchBudget, uint64_t *out) {
    if (N <= 0 || bitSize <= 2 || count <= 0) return fail(B200_EINVAL, "invalid args");   // primes_3n.go:12-14
    if (bitSize > 63 || !out) return fail(B200_EINVAL, "b200_find_3n_rns_primes: bitSize > 63 or out is NULL");
    const u64 threeN = 3ULL * (u64)N, base = 1ULL << bitSize;
    u64 candidate = ((base - 1) / threeN + 1) * threeN + 1;   // first value = 1 mod 3N above 2^bitSize
    int found = 0;
    for (int steps = 0; steps < searchBudget && found < count; ++steps) {
        if (hm::is_prime(candidate)) out[found++] = candidate;   // strictly increasing, so always distinct
        if (candidate > ~0ULL - threeN) break;
        candidate += threeN;
    }
    if (found < count)
        return fail(B200_EINVAL, "could not find enough 3N-friendly primes (found " + std::to_string(found) + "/" + std::to_string(count) + ")");
    return B200_OK;
}

extern "C" int b200_find_3n_friendly_prime(int N, int bitSize, uint64_t *out) {
    if (N <= 0 || bitSize < 0 || bitSize > 63 || !out) return fail(B200_EINVAL, "b200_find_3n_friendly_prime: invalid args");
    const u64 threeN = 3ULL * (u64)N, base = 1ULL << bitSize;
    const u64 first = ((base - 1) / threeN + 1) * threeN;
    for (u64 k = 0; k < 1000; ++k) {   // upwards: primes.go:28-43
        const u64 c = first + k * threeN + 1;
        if (c > ~0ULL - threeN) break;
        if (hm::is_prime(c)) { *out = c; return B200_OK; }
    }
    if (base > threeN) {               // then downwards: primes.go:46-64
        const u64 last = ((base - 1) / threeN) * threeN;
        for (u64 k = 0; k < 1000 && last >= k * threeN; ++k) {
            const u64 c = last - k * threeN + 1;
            if (c <= 1) break;
            if (hm::is_prime(c)) { *out = c; return B200_OK; }
        }
    }
    return fail(B200_EINVAL, "could not find 3N-friendly prime for N=" + std::to_string(N) + ", bitSize=" + std::to_string(bitSize));
}

extern "C" int b200_order_mod(uint64_t g, uint64_t p, uint64_t *out) {
    if (!out || p < 3 || !hm::is_prime(p) || g % p == 0) return fail(B200_EINVAL, "b200_order_mod: p must be an odd prime and g a unit");
    *out = order_mod(g % p, p);
    return B200_OK;
}

extern "C" int b200_has_primitive_mth_root(uint64_t q, uint64_t m) { return (m != 0 && (q - 1) % m == 0) ? 1 : 0; }

extern "C" int b200_primitive_root(uint64_t q, uint64_t *g_out) {
    if (!g_out || q < 5 || !hm::is_prime(q)) return fail(B200_EINVAL, "b200_primitive_root: q must be a prime >= 5");
    *g_out = hm::primitive_root(q, hm::unique_factors(q - 1));   // smallest primitive root >= 3, subring.go:236-248
    return B200_OK;
}

extern "C" int b200_find_primitive_root_of_unity(uint64_t q, uint64_t m, uint64_t seed, uint64_t *out) {
    if (!out || q < 5 || !hm::is_prime(q)) return fail(B200_EINVAL, "b200_find_primitive_root_of_unity: q must be a prime >= 5");
    if (!b200_has_primitive_mth_root(q, m)) return fail(B200_EINVAL, "no primitive m-th root exists: (q-1) not divisible by m");
    const u64 exp = (q - 1) / m;
    if (seed == 0) {   // the deterministic root b200_ctx_create uses: Find3NPrimitiveRoot, subring.go:255-290
        *out = hm::powmod(hm::primitive_root(q, hm::unique_factors(q - 1)), exp, q);
        return B200_OK;
    }
    u64 s = seed;
    for (int i = 0; i < 256; ++i) {   // primes_3n.go:135-147 with a seeded generator in place of crypto/rand
        const u64 x = 2 + splitmix64(s) % (q - 3);   // [2, q-2]
        const u64 g = hm::powmod(x, exp, q);
        if (g == 0 || g == 1) continue;
        if (order_mod(g, q) == m) { *out = g; return B200_OK; }
    }
    return fail(B200_EINVAL, "failed to find primitive m-th root after 256 tries");
}

extern "C" int b200_derive_3n_roots(uint64_t q, int a, int b, uint64_t omega, uint64_t out[3]) {
    if (a <= 0 || b < 0 || !out) return fail(B200_EINVAL, "invalid a,b");
    if (a > 40 || b > 24 || q < 5 || !hm::is_prime(q)) return fail(B200_EINVAL, "b200_derive_3n_roots: a, b out of range or q not prime");
    u64 p3 = 1;
    for (int i = 0; i < b; ++i) p3 *= 3;
    const u64 n = (1ULL << a) * p3, m = 3 * n;
    if (!b200_has_primitive_mth_root(q, m)) return fail(B200_EINVAL, "q=" + std::to_string(q) + " does not support primitive 3N-th root");
    if (omega % q == 0 || order_mod(omega % q, q) != m) return fail(B200_EINVAL, "omega does not have order 3N");
    out[0] = omega;
    out[1] = hm::powmod(omega, 3 * p3, q);      // order 2^a
    out[2] = hm::powmod(omega, 1ULL << a, q);   // order 3^(b+1)
    return B200_OK;
}

// ---- NTTFriendlyPrimesGenerator (ring/primes.go:69-277): primes 2^BitSize +/- k*NthRoot + 1, the state in a plain struct
static bool up_exhausted(const b200_prime_gen *g, u64 next) { return std::log2((double)next) - g->Size >= 0.5; }
static bool down_exhausted(const b200_prime_gen *g, u64 prev) { return g->Size - std::log2((double)prev) >= 0.5 || prev < g->NthRoot; }

extern "C" int b200_prime_gen_init(b200_prime_gen *g, uint64_t BitSize, uint64_t NthRoot) {
    if (!g || BitSize < 1 || BitSize > 63 || NthRoot == 0) return fail(B200_EINVAL, "b200_prime_gen_init: invalid args");
    g->Size = (double)BitSize;
    g->NthRoot = NthRoot;
    g->NextPrime = g->PrevPrime = (1ULL << BitSize) + 1;
    g->CheckNextPrime = !(g->NextPrime > ~0ULL - NthRoot);
    g->CheckPrevPrime = !(g->PrevPrime < NthRoot);
    g->PrevPrime -= NthRoot;
    return B200_OK;
}

static int next_upstream(b200_prime_gen *g, u64 *out) {
    u64 next = g->NextPrime;
    // the Go loop spins forever once CheckNextPrime is false (primes.go:147-149); here that is the same error again
    while (g->CheckNextPrime) {
        if (up_exhausted(g, next) || next > ~0ULL - g->NthRoot) break;
        if (hm::is_prime(next)) { g->NextPrime = next + g->NthRoot; *out = next; return B200_OK; }
        next += g->NthRoot;
    }
    g->CheckNextPrime = 0;
    return fail(B200_EINVAL, "cannot NextUpstreamPrime: prime list for upstream primes is exhausted (overlap with next bit-size or prime > 2^{64})");
}

static int next_downstream(b200_prime_gen *g, u64 *out) {
    u64 prev = g->PrevPrime;
    while (g->CheckPrevPrime) {
        if (down_exhausted(g, prev)) break;
        if (hm::is_prime(prev)) { g->PrevPrime = prev - g->NthRoot; *out = prev; return B200_OK; }
        prev -= g->NthRoot;
    }
    g->CheckPrevPrime = 0;
    return fail(B200_EINVAL, "cannot NextDownstreamPrime: prime list for downstream primes is exhausted (overlap with previous bit-size or prime < NthRoot");
}

static int next_alternating(b200_prime_gen *g, u64 *out) {   // primes.go:211-277, one upward and one downward probe per round
    u64 next = g->NextPrime, prev = g->PrevPrime;
    int cn = g->CheckNextPrime, cp = g->CheckPrevPrime;
    for (;;) {
        if (!(cn || cp))
            return fail(B200_EINVAL, "cannot NextAlternatingPrime: prime list for both upstream and downstream primes is exhausted (overlap with previous/next bit-size or NthRoot > prime > 2^{64} ");
        if (cn) {
            if (up_exhausted(g, next) || next > ~0ULL - g->NthRoot) cn = 0;
            else if (hm::is_prime(next)) {
                g->NextPrime = next + g->NthRoot; g->PrevPrime = prev; g->CheckNextPrime = cn; g->CheckPrevPrime = cp;
                *out = next;
                return B200_OK;
            } else next += g->NthRoot;
        }
        if (cp) {
            if (down_exhausted(g, prev)) cp = 0;
            else if (hm::is_prime(prev)) {
                g->NextPrime = next; g->PrevPrime = prev - g->NthRoot; g->CheckNextPrime = cn; g->CheckPrevPrime = cp;
                *out = prev;
                return B200_OK;
            } else prev -= g->NthRoot;
        }
    }
}

extern "C" int b200_prime_gen_next(b200_prime_gen *g, int direction, int k, uint64_t *out) {
    if (!g || !out || k < 0 || direction < -1 || direction > 1) return fail(B200_EINVAL, "b200_prime_gen_next: invalid args");
    for (int i = 0; i < k; ++i) {
        u64 p = 0;
        const int rc = direction > 0 ? next_upstream(g, &p) : direction < 0 ? next_downstream(g, &p) : next_alternating(g, &p);
        if (rc) return rc;
        out[i] = p;
    }
    return B200_OK;
}
