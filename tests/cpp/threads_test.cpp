// Concurrency test over include/b200ring.hpp: the reference documents Ring/SubRing as safe for concurrent use
// (ring/ring.go:192-194); here several host threads drive ONE ring context at once, each on its own polynomials,
// through the entry points that share the context's scratch buffers (two-pass NTT, fused product, rescale).
// Every thread must reproduce the results computed single-threaded beforehand.  Exit code 0 = pass.
#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>

#include "b200ring.hpp"

static uint64_t mix(uint64_t &s) {   // splitmix64
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

int main() {
    const std::vector<uint64_t> Q = {0x1fffffffffe00001ULL, 0x1fffffffffc80001ULL, 0x1fffffffffb40001ULL};
    const int N = 1 << 12, L = (int)Q.size(), T = 4, ITERS = 25;
    ring::Ring r = ring::Ring::NewRing(N, Q);
    std::vector<std::vector<uint64_t>> a(T), b(T), want_ntt(T), want_mul(T), want_div(T);
    for (int t = 0; t < T; ++t) {
        uint64_t seed = 1000 + t;
        a[t].resize((size_t)L * N); b[t].resize((size_t)L * N);
        for (int l = 0; l < L; ++l)
            for (int i = 0; i < N; ++i) { a[t][(size_t)l * N + i] = mix(seed) % Q[l]; b[t][(size_t)l * N + i] = mix(seed) % Q[l]; }
        ring::Poly pa = r.NewPoly(), pb = r.NewPoly(), pc = r.NewPoly();
        pa.Set(a[t]); pb.Set(b[t]);
        r.NTT(pa, pc); want_ntt[t] = pc.Get();
        r.DivRoundByLastModulusNTT(pc, pc); want_div[t] = pc.Get();
        r.MulPoly(pa, pb, pc); want_mul[t] = pc.Get();
    }
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            try {
                ring::Poly pa = r.NewPoly(), pb = r.NewPoly(), pc = r.NewPoly();
                for (int it = 0; it < ITERS; ++it) {
                    pa.Set(a[t]); pb.Set(b[t]);
                    r.NTT(pa, pc);
                    if (pc.Get() != want_ntt[t]) ++bad;
                    r.DivRoundByLastModulusNTT(pc, pc);
                    if (pc.Get() != want_div[t]) ++bad;
                    r.MulPoly(pa, pb, pc);
                    if (pc.Get() != want_mul[t]) ++bad;
                    r.NTT(pa, pa); r.INTT(pa, pa);
                    if (pa.Get() != a[t]) ++bad;
                }
            } catch (const std::exception &e) {
                std::printf("thread %d: %s\n", t, e.what());
                ++bad;
            }
        });
    for (auto &x : th) x.join();
    std::printf(bad.load() ? "FAILED (%d mismatches)\n" : "ok\n", bad.load());
    return bad.load() != 0;
}
