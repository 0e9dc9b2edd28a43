// C++ host-side test over include/b200ring.hpp: mirrors ring/ntt_test.go TestNTT (KAT N=16, first limb pair)
// and the constructor error cases of ring/ring_test.go:128-167.  Exit code 0 = pass.  Needs a GPU for the
// compute part; with --no-gpu only the constructor validation (which precedes device access) is run.
#include <cstdio>
#include <cstring>
#include <vector>

#include "b200ring.hpp"

static int fails = 0;
#define EXPECT(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

template <typename F> static bool throws_invalid(F f) {
    try { f(); } catch (const std::invalid_argument &) { return true; } catch (...) { return false; }
    return false;
}

int main(int argc, char **argv) {
    const bool no_gpu = argc > 1 && !std::strcmp(argv[1], "--no-gpu");
    const std::vector<uint64_t> Q = {576460752303439873ULL, 576460752303702017ULL};
    EXPECT(throws_invalid([&] { ring::Ring::NewRing(15, Q); }));                       // invalid ring degree
    EXPECT(throws_invalid([&] { ring::Ring::NewRing(16, {}); }));                      // empty ModuliChain
    EXPECT(throws_invalid([&] { ring::Ring::NewRing(16, {Q[0], Q[0]}); }));            // not distinct
    EXPECT(throws_invalid([&] { ring::Ring::NewRing(16, {Q[0] + 2}); }));              // not prime
    EXPECT(throws_invalid([&] { ring::Ring::NewRingFromType(16, Q, ring::Matrix); })); // Matrix needs 3N degree
    EXPECT(ring::IsValidRingDegreeFor3N(12) && ring::IsValidRingDegreeFor3N(16) && !ring::IsValidRingDegreeFor3N(20));
    // ring/primes_3n.go, ring/primes.go: host-only modulus and root search (ring/ring_3n_test.go:42-73, ring/ring_test.go:195-245)
    {
        const std::vector<uint64_t> c3 = ring::Find3NRNSPrimes(98304, 60, 24, 1000);
        EXPECT(c3.size() == 24 && c3.front() == 0x1000000000248001ULL && c3.back() == 0x10000000042b0001ULL);
        EXPECT(ring::Find3NFriendlyPrime(12288, 60) == 0x10000000000a1001ULL && ring::IsPrime(c3[7]) && !ring::IsPrime(c3[7] + 2));
        EXPECT(throws_invalid([&] { ring::Find3NRNSPrimes(98304, 60, 24, 10); }));
        const uint64_t omega = ring::Find3NPrimitiveRoot(1117, 36);
        const ring::Roots3N w = ring::Derive3NRoots(1117, 2, 1, omega);
        EXPECT(ring::OrderMod(omega, 1117) == 36 && ring::OrderMod(w.omega2Pow, 1117) == 4 && ring::OrderMod(w.omega3Pow, 1117) == 9);
        EXPECT(ring::OrderMod(ring::FindPrimitiveRootOfUnity(c3[0], 3 * 98304, 99), c3[0]) == 3 * 98304);
        EXPECT(throws_invalid([&] { ring::Derive3NRoots(1117, 2, 1, w.omega2Pow); }));
        ring::NTTFriendlyPrimesGenerator g = ring::NewNTTFriendlyPrimesGenerator(55, 1 << 17);
        const std::vector<uint64_t> alt = g.NextAlternatingPrimes(3);
        EXPECT(alt.size() == 3 && alt[0] == 0x80000000080001ULL && alt[1] == 0x80000000440001ULL && alt[2] == 0x7fffffffba0001ULL);
        EXPECT(g.NextUpstreamPrime() > alt[1] && g.NextDownstreamPrime() < alt[2]);
        EXPECT(throws_invalid([&] { ring::Ring::NewRing(98304, {c3[0], c3[0] + 3 * 98304}); }));   // second value is not prime
    }
    if (no_gpu) { std::printf(fails ? "FAILED\n" : "ok (no gpu)\n"); return fails != 0; }

    // ring/ntt_test.go:17-32, N = 16, limb 0
    const std::vector<uint64_t> poly = {29335002291498019ULL, 74733314878908829ULL, 345757914625392883ULL, 424592696763883150ULL,
        305098757618029540ULL, 315880659253740539ULL, 566291353020324899ULL, 381879490285643315ULL, 34642655966258078ULL,
        436368737741273744ULL, 422320479487058982ULL, 251503834452711492ULL, 379754966293786644ULL, 266993967580766257ULL,
        265441209649369663ULL};
    ring::Ring r = ring::Ring::NewRing(16, Q);
    EXPECT(r.N() == 16 && r.MaxLevel() == 1 && r.GetType() == ring::Standard);
    ring::Poly x = r.NewPoly(), z = r.NewPoly();
    std::vector<uint64_t> l0(16, 7);
    for (size_t i = 0; i < poly.size(); ++i) l0[i] = poly[i];
    x.SetLimb(0, 0, l0);
    x.SetLimb(0, 1, l0);
    r.NTT(x, z);
    r.INTT(z, z);
    EXPECT(z.Limb(0, 0, 16) == l0 && z.Limb(0, 1, 16) == l0);
    // plugin granularity: SubRing transformer on host slices equals the batched path
    std::vector<uint64_t> h(16), d;
    r.SubRingNTT(0, l0, h);
    r.NTT(x, z);
    d = z.Limb(0, 0, 16);
    EXPECT(h == d);
    // product pipeline of ring/ring_test.go:106-110 equals the fused product
    ring::Poly a = r.NewPoly(), b = r.NewPoly(), c = r.NewPoly(), c2 = r.NewPoly();
    a.SetLimb(0, 0, l0); a.SetLimb(0, 1, l0); b.SetLimb(0, 0, h); b.SetLimb(0, 1, h);
    r.MulPoly(a, b, c2);
    r.NTT(a, a); r.NTT(b, b); r.MForm(a, a); r.MulCoeffsMontgomery(a, b, c); r.INTT(c, c);
    EXPECT(c.Get() == c2.Get());
    // wire format (ring/poly.go:97-160): rows | (len | words)*, little-endian; round trip through a second poly
    {
        const std::vector<unsigned char> w = c.MarshalBinary(0, 1);
        EXPECT(w.size() == 8 + 2 * (8 + 16 * 8) && w[0] == 2 && w[8] == 16);
        uint64_t first = 0;
        for (int i = 0; i < 8; ++i) first |= (uint64_t)w[16 + i] << (8 * i);
        EXPECT(first == c.Limb(0, 0, 16)[0]);
        ring::Poly c3 = r.NewPoly();
        size_t used = 0;
        EXPECT(c3.UnmarshalBinary(0, w, &used) == 1 && used == w.size() && c3.Get() == c.Get());
        bool bad = false;
        try { c3.UnmarshalBinary(0, std::vector<unsigned char>(w.begin(), w.end() - 1)); } catch (const ring::Panic &) { bad = true; }
        EXPECT(bad);
    }
    {   // inverse transform with extra destinations (the key switch's fused all-gather; here the "peer" is local)
        ring::Poly g1 = r.NewPoly(), g2 = r.NewPoly(), peer = r.NewPoly();
        r.NTT(x, z);
        r.INTT(z, g1);
        r.INTTLimbsGather(z, g2, 0, 2, {peer.DevicePtr()});
        EXPECT(g1.Get() == g2.Get() && peer.Get() == g1.Get());
    }
    bool panicked = false;
    try { r.AtLevel(5); } catch (const ring::Panic &) { panicked = true; }
    EXPECT(panicked);
    std::printf(fails ? "FAILED\n" : "ok\n");
    return fails != 0;
}
