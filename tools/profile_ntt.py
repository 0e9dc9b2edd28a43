#!/usr/bin/env python3
"""Small driver for ncu / sweeps: builds the headline ring, runs NTT+INTT a few times.

    python tools/profile_ntt.py [--ring 3n|pow2] [--polys 32] [--iters 2] [--time]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import b200ring  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ring", default="3n")
ap.add_argument("--polys", type=int, default=32)
ap.add_argument("--limbs", type=int, default=24)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--time", action="store_true")
ap.add_argument("--logn", type=int, default=16)
ap.add_argument("--a", type=int, default=15)
ap.add_argument("--fold-primes", action="store_true", help="power-of-two ring on primes 2^60 + c (fold arithmetic) instead of the 61-bit Qi60 chain")
ap.add_argument("--tree", action="store_true", help="3N ring: keep NTT-domain slots in tree order (no slot permutation)")
args = ap.parse_args()

QI60 = [0x1fffffffffe00001, 0x1fffffffffc80001, 0x1fffffffffb40001, 0x1fffffffff500001,
        0x1fffffffff380001, 0x1fffffffff000001, 0x1ffffffffef00001, 0x1ffffffffee80001,
        0x1ffffffffeb40001, 0x1ffffffffe780001, 0x1ffffffffe600001, 0x1ffffffffe4c0001,
        0x1ffffffffdf40001, 0x1ffffffffdac0001, 0x1ffffffffda40001, 0x1ffffffffc680001,
        0x1ffffffffc000001, 0x1ffffffffb880001, 0x1ffffffffb7c0001, 0x1ffffffffb300001,
        0x1ffffffffb1c0001, 0x1ffffffffadc0001, 0x1ffffffffa400001, 0x1ffffffffa140001]
if args.ring == "3n":
    N = 3 << args.a
    if args.a == 15:
        moduli = bench.MODULI[: args.limbs]
    else:
        moduli = b200ring.Find3NRNSPrimes(N, 60, args.limbs, 100000)
else:
    N, moduli = 1 << args.logn, QI60[: args.limbs]
    if args.fold_primes:
        gen = b200ring.NTTFriendlyPrimesGenerator(60, 2 * N)
        moduli = gen.NextUpstreamPrimes(args.limbs)
TREE = {"tree_order": True} if args.tree else {}
ring = b200ring.NewRing(N, moduli)
ring.SetStream(torch.cuda.current_stream().cuda_stream)
data = torch.empty((args.polys, len(moduli), N), dtype=torch.int64, device="cuda")
for l, q in enumerate(moduli):
    data[:, l, :] = torch.randint(0, q, (args.polys, N), dtype=torch.int64, device="cuda")
poly = b200ring.Poly.from_torch(ring, data)
ref = data.clone()
for _ in range(args.iters):
    ring.NTT(poly, poly, **TREE)
    ring.INTT(poly, poly, **TREE)
torch.cuda.synchronize()
assert os.environ.get("B200_SKIP_CHECK") or torch.equal(ref, data)
if args.time:
    for _ in range(3):
        ring.NTT(poly, poly, **TREE); ring.INTT(poly, poly, **TREE)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ring.NTT(poly, poly, **TREE); ring.INTT(poly, poly, **TREE)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    units = 2 * args.polys * len(moduli)
    rate = units / (ms * 1e-3)
    print("ring=%s polys=%d limbs=%d chunkMB=%s: %.3f ms/step  %.0f limb-transforms/s  %.1f%% of HBM roofline" % (
        args.ring + ":N=%d" % N, args.polys, len(moduli), os.environ.get("B200_CHUNK_MB", "dflt"), ms, rate, 100 * rate * 2 * N * 8 / 6.553e12))
    prof = ring.Profile(lambda: (ring.NTT(poly, poly, **TREE), ring.INTT(poly, poly, **TREE)))
    print("   ", {k: round(v["ms"], 3) for k, v in prof.items()})
if os.environ.get("B200_PROFILE_MUL"):
    b = torch.empty_like(data)
    for l, q in enumerate(moduli):
        b[:, l, :] = torch.randint(0, q, (args.polys, N), dtype=torch.int64, device="cuda")
    c = torch.empty_like(data)
    pb, pc = b200ring.Poly.from_torch(ring, b), b200ring.Poly.from_torch(ring, c)
    for _ in range(2):
        ring.MulPoly(poly, pb, pc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        ring.MulPoly(poly, pb, pc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    units = args.polys * len(moduli)
    print("polymul ring=%s: %.3f ms  %.0f limb-products/s  %.1f%% of HBM roofline (24N bytes/limb-product)" % (
        args.ring, ms, units / (ms * 1e-3), 100 * units / (ms * 1e-3) * 3 * N * 8 / 6.553e12))
    prof = ring.Profile(lambda: ring.MulPoly(poly, pb, pc))
    print("   ", {k: round(v["ms"], 3) for k, v in prof.items()})
if os.environ.get("B200_PROFILE_VEC"):
    b = torch.empty_like(data); c = torch.empty_like(data)
    for l, q in enumerate(moduli):
        b[:, l, :] = torch.randint(0, q, (args.polys, N), dtype=torch.int64, device="cuda")
    pb, pc = b200ring.Poly.from_torch(ring, b), b200ring.Poly.from_torch(ring, c)
    for name, nbuf in (("MulCoeffsMontgomery", 3), ("MulCoeffsMontgomeryThenAdd", 4), ("MulCoeffsBarrett", 3), ("Add", 3), ("MForm", 2), ("Reduce", 2)):
        f = getattr(ring, name)
        call = (lambda: f(poly, pc)) if nbuf == 2 else (lambda: f(poly, pb, pc))
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = nbuf * data.numel() * 8 / (ms * 1e-3) / 1e9
        print("vec_op %-28s %.3f ms  %.0f GB/s  %.1f%% of HBM peak" % (name, ms, gbs, 100 * gbs / 6553))
if os.environ.get("B200_PROFILE_NEXT"):
    # survey rows f-2 / f-3: fused tensor product and RNS rescale
    bufs = [torch.empty_like(data) for _ in range(6)]
    for t in bufs[:3]:
        for l, q in enumerate(moduli):
            t[:, l, :] = torch.randint(0, q, (args.polys, N), dtype=torch.int64, device="cuda")
    a1, b0, b1, c0, c1, c2 = (b200ring.Poly.from_torch(ring, t) for t in bufs)

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms = timeit(lambda: ring.TensorProduct(poly, a1, b0, b1, c0, c1, c2))
    gbs = 7 * data.numel() * 8 / (ms * 1e-3) / 1e9
    print("tensor product (4 in, 3 out): %.3f ms  %.0f GB/s  %.1f%% of HBM peak" % (ms, gbs, 100 * gbs / 6553))

    def unfused():
        ring.MForm(poly, c0); ring.MForm(a1, c2)
        ring.MulCoeffsMontgomery(c0, b1, c1); ring.MulCoeffsMontgomeryThenAdd(c2, b0, c1)
        ring.MulCoeffsMontgomery(c0, b0, c0); ring.MulCoeffsMontgomery(c2, b1, c2)
    print("  same through separate operators: %.3f ms" % timeit(unfused))
    ms = timeit(lambda: ring.DivRoundByLastModulusNTT(poly, c0))
    Lq = len(moduli)
    print("rescale DivRoundByLastModulusNTT (L=%d -> %d): %.3f ms for %d polys = %.0f polys/s, %d limb-transforms -> %.0f limb-transforms/s"
          % (Lq, Lq - 1, ms, args.polys, args.polys / (ms * 1e-3), args.polys * Lq, args.polys * Lq / (ms * 1e-3)))
if os.environ.get("B200_PROFILE_KS"):
    # key-switch gadget product at the shape of BASELINE config 5: LQ ciphertext primes + LP special primes
    LQ, LP, B = int(os.environ.get("KS_LQ", 24)), int(os.environ.get("KS_LP", 6)), int(os.environ.get("KS_B", 8))
    primes = b200ring.Find3NRNSPrimes(N, 60, LQ + LP, 100000) if args.ring == "3n" else QI60[:LQ] + [0x1ffffffff6c80001, 0x1ffffffff6140001, 0x1ffffffff5f40001, 0x1ffffffff5700001, 0x1ffffffff4bc0001, 0x1ffffffff4380001][:LP]
    rqp = b200ring.NewRing(N, primes)
    rqp.SetStream(torch.cuda.current_stream().cuda_stream)
    ks = b200ring.KeySwitcher(rqp, LQ)
    beta = (LQ + LP - 1) // LP
    mk = lambda n: torch.stack([torch.stack([torch.randint(0, q, (N,), dtype=torch.int64, device="cuda") for q in primes]) for _ in range(n)])
    cx_t, key_t, o0_t, o1_t = mk(B), mk(2 * beta), mk(B), mk(B)
    cx, key, o0, o1 = (b200ring.Poly.from_torch(rqp, t) for t in (cx_t, key_t, o0_t, o1_t))
    for _ in range(2):
        ks.GadgetProduct(LQ - 1, LP - 1, cx, key, o0, o1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        ks.GadgetProduct(LQ - 1, LP - 1, cx, key, o0, o1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nt = LQ + beta * (LQ + LP) - sum(min(LP, LQ - d * LP) for d in range(beta)) + 2 * (LP + LQ)
    print("gadget product ring=%s N=%d LQ=%d LP=%d beta=%d batch=%d: %.3f ms = %.1f key-switches/s, %d limb-transforms each -> %.0f limb-transforms/s"
          % (args.ring, N, LQ, LP, beta, B, ms, B / (ms * 1e-3), nt, B * nt / (ms * 1e-3)))
    prof = rqp.Profile(lambda: ks.GadgetProduct(LQ - 1, LP - 1, cx, key, o0, o1))
    print("   ", {k: round(v["ms"], 3) for k, v in prof.items()})
