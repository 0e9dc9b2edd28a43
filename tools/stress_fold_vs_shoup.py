#!/usr/bin/env python3
"""Stress comparison of the two independent transform implementations (fold arithmetic / Shoup arithmetic) on the GPU:
random and structured inputs, all 24 primes of BASELINE config 3 at N = 3*2^15 plus a power-of-two ring on fold primes,
forward, inverse, lazy and tree-order variants, bit for bit.  The Shoup ring is built with B200_NO_FOLD=1.

    python tools/stress_fold_vs_shoup.py [--rounds 6]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import b200ring  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rounds", type=int, default=6)
args = ap.parse_args()


def rings(N, moduli):
    os.environ.pop("B200_NO_FOLD", None)
    fold = b200ring.NewRing(N, moduli)
    os.environ["B200_NO_FOLD"] = "1"
    shoup = b200ring.NewRing(N, moduli)
    os.environ.pop("B200_NO_FOLD", None)
    return fold, shoup


def check(name, N, moduli, polys):
    fold, shoup = rings(N, moduli)
    gen = torch.Generator(device="cuda"); gen.manual_seed(0xF01D)
    qs = torch.tensor(moduli, dtype=torch.int64, device="cuda").view(1, -1, 1)
    bad = 0
    for rnd in range(args.rounds):
        x = torch.empty((polys, len(moduli), N), dtype=torch.int64, device="cuda")
        if rnd == 0:
            x[:] = qs - 1                                            # all q-1
        elif rnd == 1:
            x.zero_(); x[:, :, ::2] = (qs - 1).expand(polys, -1, N // 2)   # alternating
        elif rnd == 2:
            x[:] = qs + (qs - 1)                                      # 2q-1: the top of the input contract
        else:
            for l, q in enumerate(moduli):
                x[:, l, :] = torch.randint(0, q, (polys, N), dtype=torch.int64, device="cuda", generator=gen)
        for tree in ((False, True) if name.startswith("3n") else (False,)):
            kw = {"tree_order": True} if tree else {}
            a, b = x.clone(), x.clone()
            pa, pb = b200ring.Poly.from_torch(fold, a), b200ring.Poly.from_torch(shoup, b)
            fold.NTT(pa, pa, **kw); shoup.NTT(pb, pb, **kw)
            torch.cuda.synchronize()
            bad += int((a != b).sum().item())
            fold.INTT(pa, pa, **kw); shoup.INTT(pb, pb, **kw)
            torch.cuda.synchronize()
            bad += int((a != b).sum().item())
            bad += int((a != x % qs).sum().item())                   # round trip returns the canonical input
            # the inverse of arbitrary words taken as NTT-domain data (any representative below 2q)
            a, b = x.clone(), x.clone()
            pa, pb = b200ring.Poly.from_torch(fold, a), b200ring.Poly.from_torch(shoup, b)
            fold.INTT(pa, pa, **kw); shoup.INTT(pb, pb, **kw)
            torch.cuda.synchronize()
            bad += int((a != b).sum().item())
            fold.NTTLazy(pa, pa, **kw); shoup.NTTLazy(pb, pb, **kw)
            torch.cuda.synchronize()
            bad += int(((a % qs) != (b % qs)).sum().item())
    print("%-28s rounds=%d polys=%d limbs=%d: %s" % (name, args.rounds, polys, len(moduli), "ok" if bad == 0 else "%d MISMATCHES" % bad))
    return bad


bad = check("3n N=3*2^15 (24 C3 primes)", 3 << 15, bench.MODULI, 4)
bad += check("3n N=3*2^12 (generic fold)", 3 << 12, b200ring.Find3NRNSPrimes(3 << 12, 60, 4, 100000), 4)
g = b200ring.NTTFriendlyPrimesGenerator(60, 2 << 16)
bad += check("pow2 N=2^16 (fold primes)", 1 << 16, g.NextUpstreamPrimes(6), 4)
g = b200ring.NTTFriendlyPrimesGenerator(60, 2 << 13)
bad += check("pow2 N=2^13 (generic fold)", 1 << 13, g.NextUpstreamPrimes(4), 4)
sys.exit(1 if bad else 0)
