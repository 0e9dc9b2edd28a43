#!/usr/bin/env python3
"""BASELINE config 5: key-switch gadget product at N = 3*2^15, LQ = 24, LP = 6 (beta = 4), limb-sharded over G GPUs
with NCCL all-gathers (one of the coefficient-domain input limbs, one of the accumulators' P limbs).

    python tools/bench_keyswitch.py                       # one GPU, b200_gadget_product
    torchrun --nproc-per-node G tools/bench_keyswitch.py  # G ranks, sharding.gadget_product_sharded
Prints one JSON line on rank 0 (key-switches/s for the whole job; the batch of ciphertexts is fixed = strong scaling).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import b200ring  # noqa: E402
from b200ring import sharding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--lq", type=int, default=24)
ap.add_argument("--lp", type=int, default=6)
ap.add_argument("--a", type=int, default=15)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--profile", action="store_true", help="per-kernel CUDA-event times of one call on every rank")
ap.add_argument("--verify", action="store_true", help="compare this rank's Q limbs with the single-GPU b200_gadget_product on the same data")
ap.add_argument("--fused", action="store_true", help="exchanges fused into the inverse transforms (peer stores over NVLink)")
args = ap.parse_args()
rank, local, world = sharding.init_process_group()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
N, LQ, LP, B = 3 << args.a, args.lq, args.lp, args.batch
primes = b200ring.Find3NRNSPrimes(N, 60, LQ + LP, 100000)
ring = b200ring.NewRing(N, primes, device=local)
ring.SetStream(torch.cuda.current_stream().cuda_stream)
ks = b200ring.KeySwitcher(ring, LQ)
levelQ, levelP = LQ - 1, LP - 1
beta = (levelQ + levelP + 1) // (levelP + 1)
gen = torch.Generator(device=dev); gen.manual_seed(1234)          # same data on every rank
mk = lambda n: torch.stack([torch.stack([torch.randint(0, q, (N,), dtype=torch.int64, device=dev, generator=gen) for q in primes]) for _ in range(n)])
cx_t, key_t = mk(B), mk(2 * beta)
key = b200ring.Poly.from_torch(ring, key_t)
o0, o1 = torch.zeros_like(cx_t), torch.zeros_like(cx_t)


fused = None
if args.fused:
    fused = sharding.ShardedKeySwitch(ks, ring, B, rank, world)
    cx_p, o0_p, o1_p = b200ring.Poly.from_torch(ring, cx_t), b200ring.Poly.from_torch(ring, o0), b200ring.Poly.from_torch(ring, o1)


def run():
    if fused is not None:
        fused.gadget_product(levelQ, levelP, cx_p, key, o0_p, o1_p)
    elif world == 1:
        ks.GadgetProduct(levelQ, levelP, b200ring.Poly.from_torch(ring, cx_t), key, b200ring.Poly.from_torch(ring, o0), b200ring.Poly.from_torch(ring, o1))
    else:
        sharding.gadget_product_sharded(ks, ring, levelQ, levelP, cx_t, key, o0, o1, rank, world)


def barrier():
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


for _ in range(2):
    run()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    run()
e1.record()
barrier()
ms = sharding.max_over_ranks(e0.elapsed_time(e1) / args.reps, device=dev)
q0, q1 = sharding.block_of_rank(levelQ + 1, world, rank)
chk = int(o0[:, q0:q1].sum().item()) & 0xFFFFFFFF
verified = None
if args.verify:   # every rank holds the whole input (same seed): the unsharded product on this GPU is the check
    r0, r1 = torch.zeros_like(cx_t), torch.zeros_like(cx_t)
    ks.GadgetProduct(levelQ, levelP, b200ring.Poly.from_torch(ring, cx_t.clone()), key, b200ring.Poly.from_torch(ring, r0), b200ring.Poly.from_torch(ring, r1))
    ring.Sync()
    ok = bool(torch.equal(r0[:, q0:q1], o0[:, q0:q1]) and torch.equal(r1[:, q0:q1], o1[:, q0:q1]) and (q1 == q0 or int(r0[:, q0:q1].abs().sum().item()) != 0))
    verified = sharding.max_over_ranks(0.0 if ok else 1.0, device=dev) == 0.0
if rank == 0:
    print(json.dumps({"workload": "gadget product N=%d LQ=%d LP=%d beta=%d batch=%d" % (N, LQ, LP, beta, B), "n_gpus": world, "fused_exchanges": bool(args.fused),
                      "ms": ms, "key_switches_per_s": B / (ms * 1e-3), "limb_transforms_per_s": B * 180 / (ms * 1e-3) if (LQ, LP) == (24, 6) else None,
                      "allgather_bytes_per_step": B * (LQ + 2 * LP) * N * 8, "checksum_rank0": chk, "verified_against_single_gpu": verified}))
if args.profile:
    prof = ring.Profile(run)
    barrier()
    for r in range(world):
        if r == rank:
            sys.stderr.write("rank %d: %s\n" % (rank, {k: round(v["ms"], 3) for k, v in prof.items()}))
        barrier()
if fused is not None:
    fused.close()
if world > 1:
    torch.distributed.destroy_process_group()
