#!/usr/bin/env python3
"""All five BASELINE.json configs at their full sizes on one GPU, one JSON line each (CUDA-event times, 3 warm-up
repetitions; inputs resident in HBM and larger than L2 except config 1, which is a single 96 KiB limb).

    python tools/bench_configs.py [--out profiles/r01_configs.json]

  C1  3N ring N = 3*2^12, one prime, one polynomial: forward + inverse round trip (latency; the reference's own test shape)
  C2  N = 2^16, L = 24, 256 polynomials: NTT + INTT in place
  C3  N = 3*2^15, L = 24, 128 polynomials: NTT + INTT in place (the bench.py headline)
  C4  N = 3*2^15, L = 16, 1024 (a, b) pairs: c = INTT(MulCoeffsMontgomery(MForm(NTT a), NTT b)) through b200_polymul
  C5  N = 3*2^15, LQ = 24, LP = 6: ciphertext x ciphertext multiply with relinearisation (b200_mul_relin), batch 16
Inputs are uniform residues (torch generator on the device; parity is the tests' job, here only round trips are checked).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import b200ring  # noqa: E402
import bench  # noqa: E402

PEAK = bench.peaks()[0] * 1e9
QI60 = [0x1fffffffffe00001, 0x1fffffffffc80001, 0x1fffffffffb40001, 0x1fffffffff500001,
        0x1fffffffff380001, 0x1fffffffff000001, 0x1ffffffffef00001, 0x1ffffffffee80001,
        0x1ffffffffeb40001, 0x1ffffffffe780001, 0x1ffffffffe600001, 0x1ffffffffe4c0001,
        0x1ffffffffdf40001, 0x1ffffffffdac0001, 0x1ffffffffda40001, 0x1ffffffffc680001,
        0x1ffffffffc000001, 0x1ffffffffb880001, 0x1ffffffffb7c0001, 0x1ffffffffb300001,
        0x1ffffffffb1c0001, 0x1ffffffffadc0001, 0x1ffffffffa400001, 0x1ffffffffa140001]   # ring/test_params.go:15-20


def uniform(ring, moduli, polys, N, gen):
    t = torch.empty((polys, len(moduli), N), dtype=torch.int64, device="cuda")
    for l, q in enumerate(moduli):
        t[:, l, :] = torch.randint(0, q, (polys, N), dtype=torch.int64, device="cuda", generator=gen)
    return t


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def transform_config(name, N, moduli, polys, reps, what):
    ring = b200ring.NewRing(N, moduli)
    ring.SetStream(torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device="cuda"); gen.manual_seed(0xB200)
    data = uniform(ring, moduli, polys, N, gen)
    ref = data.clone()
    p = b200ring.Poly.from_torch(ring, data)
    ms = timed(lambda: (ring.NTT(p, p), ring.INTT(p, p)), reps)
    assert torch.equal(ref, data), name + ": round trip"
    units = 2 * polys * len(moduli)
    rate = units / (ms * 1e-3)
    return {"config": name, "workload": what, "ms_per_step": ms, "value": rate, "unit": "limb-transforms/s",
            "roofline_frac": rate * 2 * N * 8 / PEAK, "resident_bytes": data.numel() * 8}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    want = set(args.only.split(",")) if args.only else {"C1", "C2", "C3", "C4", "C5"}
    rows = []
    if "C1" in want:
        N = 3 << 12
        r = transform_config("C1", N, b200ring.Find3NRNSPrimes(N, 60, 1, 100000), 1, 50, "3N ring N=3*2^12, 1 prime, 1 polynomial, NTT+INTT round trip")
        r["note"] = "latency of two launches on one 96 KiB limb, not a throughput shape"
        rows.append(r)
    if "C2" in want:
        rows.append(transform_config("C2", 1 << 16, QI60, 256, 10, "N=2^16, L=24 (Qi60[0:24]), 256 polynomials, NTT+INTT in place"))
    if "C3" in want:
        rows.append(transform_config("C3", bench.N, bench.MODULI, 128, 10, "N=3*2^15, L=24, 128 polynomials, NTT+INTT in place"))
    if "C4" in want:
        N, moduli, polys = bench.N, bench.MODULI[:16], 1024
        ring = b200ring.NewRing(N, moduli)
        ring.SetStream(torch.cuda.current_stream().cuda_stream)
        gen = torch.Generator(device="cuda"); gen.manual_seed(0xB204)
        a, b = uniform(ring, moduli, polys, N, gen), uniform(ring, moduli, polys, N, gen)
        c = torch.empty_like(a)
        pa, pb, pc = (b200ring.Poly.from_torch(ring, t) for t in (a, b, c))
        ms = timed(lambda: ring.MulPoly(pa, pb, pc), 3, warm=2)
        # X * b = b shifted: a := X (monomial) on the first polynomial checks the product cheaply
        a[0].zero_(); a[0, :, 1] = 1
        ring.MulPoly(pa, pb, pc)
        torch.cuda.synchronize()
        bb, cc = b[0], c[0]
        q = torch.tensor(moduli, dtype=torch.int64, device="cuda").view(-1, 1)
        # X * sum b_j X^j mod X^N - X^(N/2) + 1:  c_0 = -b_{N-1}, c_{N/2} = b_{N/2-1} + b_{N-1}, c_j = b_{j-1} otherwise
        exp = torch.roll(bb, 1, dims=1).clone()
        last = bb[:, N - 1:N]
        exp[:, 0:1] = (q - last) % q
        exp[:, N // 2:N // 2 + 1] = (bb[:, N // 2 - 1:N // 2] + last) % q
        assert torch.equal(exp, cc), "C4: X * b"
        units = polys * len(moduli)
        rate = units / (ms * 1e-3)
        rows.append({"config": "C4", "workload": "N=3*2^15, L=16, 1024 (a,b) pairs, fused NTT,NTT,mul,INTT (b200_polymul)", "ms_per_step": ms,
                     "value": rate, "unit": "limb-products/s", "roofline_frac": rate * 3 * N * 8 / PEAK, "resident_bytes": 3 * a.numel() * 8,
                     "limb_transform_equivalent_per_s": 3 * rate})
        del a, b, c, pa, pb, pc, ring
        torch.cuda.empty_cache()
    if "C5" in want:
        N, LQ, LP, B = bench.N, 24, 6, 16
        primes = b200ring.Find3NRNSPrimes(N, 60, LQ + LP, 100000)
        ring = b200ring.NewRing(N, primes)
        ring.SetStream(torch.cuda.current_stream().cuda_stream)
        ks = b200ring.KeySwitcher(ring, LQ)
        levelQ, levelP = LQ - 1, LP - 1
        beta = (levelQ + levelP + 1) // (levelP + 1)
        gen = torch.Generator(device="cuda"); gen.manual_seed(0xB205)
        t = [uniform(ring, primes, B, N, gen) for _ in range(4)]
        rlk = uniform(ring, primes, 2 * beta, N, gen)
        o = [torch.zeros_like(t[0]) for _ in range(2)]
        P = lambda x: b200ring.Poly.from_torch(ring, x)
        polys = [P(x) for x in t] + [P(rlk)] + [P(x) for x in o]
        ms = timed(lambda: ks.MulRelin(levelQ, levelP, *polys), 5, warm=2)
        ms_ks = timed(lambda: ks.GadgetProduct(levelQ, levelP, polys[0], polys[4], polys[5], polys[6]), 5, warm=2)
        dec = ring.NewPoly(beta * B)
        ms_dec = timed(lambda: ks.DecomposeNTT(levelQ, levelP, polys[0], dec), 5, warm=2)
        ms_hoist = timed(lambda: ks.GadgetProductHoisted(levelQ, levelP, dec, polys[4], polys[5], polys[6]), 5, warm=2)
        rows.append({"config": "C5", "workload": "N=3*2^15, LQ=24, LP=6 (beta=%d): ct x ct multiply + relinearisation (b200_mul_relin), batch %d, one GPU" % (beta, B),
                     "ms_per_step": ms, "value": B / (ms * 1e-3), "unit": "ciphertext multiplies/s",
                     "gadget_product_ms": ms_ks, "key_switches_per_s": B / (ms_ks * 1e-3),
                     "decompose_ntt_ms": ms_dec, "gadget_product_hoisted_ms": ms_hoist,
                     "hoisted_key_switches_per_s": B / (ms_hoist * 1e-3),
                     "limb_transforms_per_key_switch": beta * (LQ + LP) + 2 * (LQ + LP), "note": "limb-sharded 2/4-GPU numbers: tools/bench_keyswitch.py"})
    for r in rows:
        print(json.dumps(r))
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"peak_hbm_gbs": PEAK / 1e9, "gpu": torch.cuda.get_device_name(0), "configs": rows}, f, indent=1)


if __name__ == "__main__":
    main()
