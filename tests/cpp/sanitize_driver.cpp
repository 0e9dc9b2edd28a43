// Driver for compute-sanitizer (memcheck / racecheck): runs every kernel family once at small-but-real shapes
// through the C ABI and checks round trips.  No oracle, no Python: `compute-sanitizer --tool memcheck ./sanitize_driver`.
#include <cstdio>
#include <random>
#include <vector>

#include "b200ring.hpp"

static int fails = 0;
#define EXPECT(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

static std::vector<uint64_t> random_poly(std::mt19937_64 &g, int batch, const std::vector<uint64_t> &q, int N) {
    std::vector<uint64_t> v((size_t)batch * q.size() * N);
    for (int b = 0; b < batch; ++b)
        for (size_t l = 0; l < q.size(); ++l)
            for (int i = 0; i < N; ++i) v[((size_t)b * q.size() + l) * N + i] = g() % q[l];
    return v;
}

static void exercise(int N, const std::vector<uint64_t> &Q, int LQ, const char *name) {
    std::mt19937_64 g(42);
    ring::Ring r = ring::Ring::NewRing(N, Q);
    const int batch = 2;
    auto ha = random_poly(g, batch, Q, N), hb = random_poly(g, batch, Q, N);
    ring::Poly a = r.NewPoly(batch), b = r.NewPoly(batch), c = r.NewPoly(batch), d = r.NewPoly(batch), e = r.NewPoly(batch);
    a.Set(ha); b.Set(hb);
    r.NTT(a, c); r.INTT(c, c);
    EXPECT(c.Get() == ha);
    r.NTTLazy(a, c); r.Reduce(c, c); r.INTTLazy(c, c); r.Reduce(c, c);
    EXPECT(c.Get() == ha);
    r.MulPoly(a, b, c);                                   // fused product
    r.NTT(a, d); r.NTT(b, e); r.MForm(d, d); r.MulCoeffsMontgomery(d, e, d); r.INTT(d, d);
    EXPECT(c.Get() == d.Get());
    ring::check(b200_tensor(r.handle(), a.handle(), b.handle(), b.handle(), a.handle(), c.handle(), d.handle(), e.handle(), r.MaxLevel(), 1));
    ring::check(b200_div_by_last_modulus_ntt(r.handle(), a.handle(), c.handle(), r.MaxLevel(), 1));
    std::vector<uint64_t> out(ha.size());
    const int ops[2] = {0, 1};
    ring::check(b200_transform_host(r.handle(), ha.data(), out.data(), batch, r.MaxLevel(), ops, 2, 0));
    EXPECT(out == ha);
    if ((N & (N - 1)) == 0) {                                  // automorphisms and the monomial product (power-of-two rings)
        const uint64_t gen = 5;
        r.AutomorphismNTT(a, gen, c);
        r.AutomorphismNTTWithIndex(a, ring::Ring::AutomorphismNTTIndex(N, 2 * (uint64_t)N, gen), d);
        EXPECT(c.Get() == d.Get());
        r.NTT(a, c); r.AutomorphismNTT(c, gen, d);             // AutomorphismNTT(NTT p) == NTT(Automorphism p)
        r.Automorphism(a, gen, c); r.Reduce(c, c); r.NTT(c, c);
        EXPECT(c.Get() == d.Get());
        r.MultByMonomial(a, 1, c); r.MultByMonomial(c, 8, c); r.MultByMonomial(a, 9, d);
        r.Reduce(c, c); r.Reduce(d, d);
        EXPECT(c.Get() == d.Get());
        r.Shift(a, 3, c); r.Shift(c, -3, c);
        EXPECT(c.Get() == ha);
    }
    if (LQ > 0) {
        b200_keyswitch *ks = nullptr;
        ring::check(b200_keyswitch_create(r.handle(), LQ, &ks));
        const int LP = (int)Q.size() - LQ, beta = (LQ + LP - 1) / LP;
        ring::Poly key = r.NewPoly(2 * beta);
        key.Set(random_poly(g, 2 * beta, Q, N));
        ring::check(b200_gadget_product(ks, LQ - 1, LP - 1, a.handle(), key.handle(), c.handle(), d.handle()));
        ring::check(b200_mul_relin(ks, LQ - 1, LP - 1, a.handle(), b.handle(), b.handle(), a.handle(), key.handle(), c.handle(), d.handle(), 1));
        ring::check(b200_mod_up(ks, 0, LQ - 1, LP - 1, a.handle(), c.handle()));
        ring::check(b200_mod_down_ntt(ks, LQ - 1, LP - 1, a.handle(), c.handle()));
        b200_keyswitch_destroy(ks);
    }
    r.Sync();
    std::printf("%s N=%d L=%zu ok\n", name, N, Q.size());
}

int main() {
    const std::vector<uint64_t> qi = {0x1fffffffffe00001ULL, 0x1fffffffffc80001ULL, 0x1fffffffffb40001ULL, 0x1fffffffff500001ULL, 0x1fffffffff380001ULL};
    // 3N-friendly primes = 1 mod 3*N for N up to 3*2^13 (Find3NRNSPrimes(3*2^13, 60, 5))
    const std::vector<uint64_t> q3 = {0x10000000000a1001ULL, 0x100000000014c001ULL, 0x100000000015e001ULL, 0x10000000001a6001ULL, 0x10000000001e5001ULL};   // Find3NRNSPrimes(3*2^12, 60, 5)
    exercise(1 << 12, qi, 3, "standard/big-path");
    exercise(64, qi, 3, "standard/resident");
    exercise(1 << 17, {qi[0], qi[1]}, 0, "standard/2^17");
    exercise(3 << 12, q3, 3, "matrix/big-path");
    exercise(24, q3, 3, "matrix/resident");
    {   // conjugate-invariant ring: round trip through the folding layer on both kernel families
        for (int N : {64, 1 << 11}) {
            ring::Ring r = ring::Ring::NewRingConjugateInvariant(N, {qi[0], qi[1]});
            std::mt19937_64 g(7);
            auto h = random_poly(g, 2, {qi[0], qi[1]}, N);
            ring::Poly a = r.NewPoly(2), b = r.NewPoly(2);
            a.Set(h);
            r.NTT(a, b); r.INTT(b, b);
            EXPECT(b.Get() == h);
            std::printf("conjugate-invariant N=%d ok\n", N);
        }
    }
    std::printf(fails ? "FAILED\n" : "ok\n");
    return fails != 0;
}
