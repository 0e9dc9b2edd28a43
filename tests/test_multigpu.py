"""Limb-sharded gadget product (BASELINE config 5): the range-limited device steps emulated for G ranks on one GPU,
and -- when the box has >= 2 GPUs -- the real thing under torchrun with NCCL all-gathers."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, QI60
import b200ring
from b200ring import sharding
from oracle import oracle as O


def test_block_partition():
    for n in (1, 6, 24, 30):
        for G in (1, 2, 4, 8):
            spans = [sharding.block_of_rank(n, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))


def _setup(kind, N, LQ, LP, batch, seed=5):
    if kind == "std":
        primes = QI60[: LQ + LP]
    else:
        primes = O.find_3n_rns_primes(N, 60, LQ + LP)
    ring = b200ring.NewRing(N, primes)
    ks = b200ring.KeySwitcher(ring, LQ)
    rng = np.random.default_rng(seed)
    beta = (LQ + LP - 1) // LP
    mk = lambda n: np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in primes]) for _ in range(n)])
    return ring, ks, primes, beta, mk(batch), mk(2 * beta)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,LQ,LP,G", [("std", 1 << 12, 6, 2, 2), ("m3", 3 * 2**10, 7, 3, 4), ("m3", 3 * 2**12, 5, 2, 8)])
def test_sharded_steps_emulated_on_one_gpu(kind, N, LQ, LP, G):
    """Each emulated rank runs only the range-limited steps on its own limbs; the all-gathers are plain copies."""
    batch = 2
    ring, ks, primes, beta, hcx, hkey = _setup(kind, N, LQ, LP, batch)
    levelQ, levelP = LQ - 1, LP - 1
    cx, key = ring.NewPoly(batch).set(hcx), ring.NewPoly(2 * beta).set(hkey)
    w0, w1 = ring.NewPoly(batch), ring.NewPoly(batch)
    ks.GadgetProduct(levelQ, levelP, cx, key, w0, w1)
    want0, want1 = w0.get(), w1.get()
    qb = [sharding.block_of_rank(levelQ + 1, G, r) for r in range(G)]
    pb = [sharding.block_of_rank(levelP + 1, G, r) for r in range(G)]
    # step 1 + gather
    cxinv_full = np.zeros_like(hcx)
    for r in range(G):
        t = ring.NewPoly(batch)
        ring.NTTLimbs(cx, t, qb[r][0], qb[r][1] - qb[r][0], inverse=True)
        cxinv_full[:, qb[r][0]: qb[r][1]] = t.get()[:, qb[r][0]: qb[r][1]]
    cxinv = ring.NewPoly(batch).set(cxinv_full)
    # step 2 + 3 per rank, then gather of the P limbs
    acc_full = [np.zeros_like(hcx), np.zeros_like(hcx)]
    for r in range(G):
        only = np.zeros_like(hcx)
        only[:, qb[r][0]: qb[r][1]] = hcx[:, qb[r][0]: qb[r][1]]          # a rank sees only its own NTT-domain limbs
        a0, a1 = ring.NewPoly(batch), ring.NewPoly(batch)
        ks.DecomposeMac(levelQ, levelP, ring.NewPoly(batch).set(only), cxinv, key, a0, a1, qb[r][0], qb[r][1], pb[r][0], pb[r][1])
        for a, full in ((a0, acc_full[0]), (a1, acc_full[1])):
            ring.NTTLimbs(a, a, LQ + pb[r][0], pb[r][1] - pb[r][0], inverse=True)
            g = a.get()
            full[:, qb[r][0]: qb[r][1]] = g[:, qb[r][0]: qb[r][1]]
            full[:, LQ + pb[r][0]: LQ + pb[r][1]] = g[:, LQ + pb[r][0]: LQ + pb[r][1]]
    # step 4 per rank
    for c, want in ((0, want0), (1, want1)):
        for r in range(G):
            mine = np.zeros_like(hcx)
            mine[:, qb[r][0]: qb[r][1]] = acc_full[c][:, qb[r][0]: qb[r][1]]
            mine[:, LQ:] = acc_full[c][:, LQ:]
            out = ring.NewPoly(batch)
            ks.ModDownLimbs(levelQ, levelP, ring.NewPoly(batch).set(mine), out, qb[r][0], qb[r][1])
            got = out.get()
            assert (got[:, qb[r][0]: qb[r][1]] == want[:, qb[r][0]: qb[r][1]]).all(), (c, r)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N", [("m3", 3 * 2**15), ("std", 1 << 16), ("std", 1 << 13), ("m3", 3 * 2**12), ("std", 64), ("m3", 9 * 2**6)])
def test_gather_inverse_writes_every_destination(kind, N):
    """b200_intt_limbs_gather with the 'peers' on the same GPU: every destination receives the transform of the
    requested limbs (tuned, generic and fallback kernel families) and nothing else is touched."""
    L, batch = 4, 2
    primes = QI60[:L] if kind == "std" else O.find_3n_rns_primes(N, 60, L)
    ring = b200ring.NewRing(N, primes)
    rng = np.random.default_rng(9)
    host = np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in primes]) for _ in range(batch)])
    src, want = ring.NewPoly(batch).set(host), ring.NewPoly(batch)
    ring.INTT(src, want)
    w = want.get()
    for limb0, nl, npeers in ((1, 2, 3), (0, 4, 1), (3, 1, 7), (2, 2, 0)):
        sentinel = np.full_like(host, 12345)
        out = ring.NewPoly(batch).set(sentinel)
        peers = [ring.NewPoly(batch).set(sentinel) for _ in range(npeers)]
        ring.INTTLimbsGather(src, out, limb0, nl, [p.DevicePtr() for p in peers])
        for dst in [out] + peers:
            g = dst.get()
            assert (g[:, limb0:limb0 + nl] == w[:, limb0:limb0 + nl]).all()
            mask = np.ones(L, dtype=bool); mask[limb0:limb0 + nl] = False
            assert (g[:, mask] == 12345).all()
    # in place (the accumulators of the key switch are transformed where they are)
    peer = ring.NewPoly(batch)
    ring.INTTLimbsGather(src, src, 0, L, [peer.DevicePtr()])
    assert (src.get() == w).all() and (peer.get() == w).all()
    with pytest.raises(b200ring.B200Error):
        ring.INTTLimbsGather(src, src, 0, L, [peer.DevicePtr()] * 8)


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch, torch.distributed as dist
import b200ring
from b200ring import sharding
from test_multigpu import _setup
rank, local, world = sharding.init_process_group("nccl")
torch.cuda.set_device(local)
kind, N, LQ, LP, batch = "m3", 3 * 2**13, 8, 4, 3
ring, ks, primes, beta, hcx, hkey = _setup(kind, N, LQ, LP, batch)     # same seed on every rank
# the ring was created on device 0 by default: rebuild on this rank's device
ring = b200ring.NewRing(N, primes, device=local); ks = b200ring.KeySwitcher(ring, LQ)
ring.SetStream(torch.cuda.current_stream().cuda_stream)
levelQ, levelP = LQ - 1, LP - 1
key = ring.NewPoly(2 * beta).set(hkey)
full = ring.NewPoly(batch).set(hcx); w0, w1 = ring.NewPoly(batch), ring.NewPoly(batch)
ks.GadgetProduct(levelQ, levelP, full, key, w0, w1)                      # single-GPU result as the expectation
q0, q1 = sharding.block_of_rank(levelQ + 1, world, rank)
mine = np.zeros_like(hcx); mine[:, q0:q1] = hcx[:, q0:q1]
cx_t = torch.from_numpy(mine.view(np.int64)).cuda()
o0, o1 = torch.zeros_like(cx_t), torch.zeros_like(cx_t)
sharding.gadget_product_sharded(ks, ring, levelQ, levelP, cx_t, key, o0, o1, rank, world)
g0, g1 = o0.cpu().numpy().view(np.uint64), o1.cpu().numpy().view(np.uint64)
assert (g0[:, q0:q1] == w0.get()[:, q0:q1]).all() and (g1[:, q0:q1] == w1.get()[:, q0:q1]).all(), "sharded result differs"
# the same with both exchanges fused into the producing kernels (peer stores over NVLink + one-word barriers)
fused = sharding.ShardedKeySwitch(ks, ring, batch, rank, world)
cxp = ring.NewPoly(batch).set(mine); f0, f1 = ring.NewPoly(batch), ring.NewPoly(batch)
for _ in range(3):                                                        # repeated calls reuse the mapped buffers
    fused.gadget_product(levelQ, levelP, cxp, key, f0, f1)
assert (f0.get()[:, q0:q1] == w0.get()[:, q0:q1]).all() and (f1.get()[:, q0:q1] == w1.get()[:, q0:q1]).all(), "fused sharded result differs"
# under a non-default torch stream: the barrier's all-reduce is ordered against THAT stream, so the fused path must move
# the ring onto it (ShardedKeySwitch._bind_stream); back-to-back calls with no host synchronisation in between
side = torch.cuda.Stream()
f0.set(np.zeros_like(hcx)); f1.set(np.zeros_like(hcx)); ring.Sync()
with torch.cuda.stream(side):
    for _ in range(4):
        fused.gadget_product(levelQ, levelP, cxp, key, f0, f1)
side.synchronize()
assert (f0.get()[:, q0:q1] == w0.get()[:, q0:q1]).all() and (f1.get()[:, q0:q1] == w1.get()[:, q0:q1]).all(), "fused result differs on a side stream"
ring.SetStream(torch.cuda.current_stream().cuda_stream)
fused.close()
dist.barrier(); dist.destroy_process_group()
os.write(1, ("rank" + str(rank) + "-ok\n").encode())
'''


@pytest.mark.gpu
def test_sharded_gadget_product_nccl(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)]
    r = subprocess.run(cmd, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for k in range(world):
        assert ("rank%d-ok" % k) in r.stdout
