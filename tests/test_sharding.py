"""Limb sharding: partition logic + a world_size-2 gloo run of the N>1 host path."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from b200ring import sharding


def test_limb_partition_is_a_disjoint_cover():
    for L in (1, 3, 16, 24):
        for G in (1, 2, 4, 8):
            owned = [sharding.limbs_of_rank(L, G, r) for r in range(G)]
            flat = sorted(i for o in owned for i in o)
            assert flat == list(range(L))
            for r, o in enumerate(owned):
                assert all(sharding.owner_of_limb(i, G) == r for i in o)
    idx, mod, roots = sharding.shard_moduli([11, 13, 17, 19, 23], 2, 1, roots=[1, 2, 3, 4, 5])
    assert idx == [1, 3] and mod == [13, 19] and roots == [2, 4]
    with pytest.raises(ValueError):
        sharding.limbs_of_rank(4, 2, 2)


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from b200ring import sharding
rank, local, world = sharding.init_process_group("gloo")
assert world == 2 and dist.get_backend() == "gloo"
L, N, B = 6, 8, 3
idx = sharding.limbs_of_rank(L, world, rank)
full = torch.arange(B * L * N, dtype=torch.int64).reshape(B, L, N)
local_t = full[:, idx, :].contiguous()
got = sharding.gather_limbs(local_t, L, world, rank)
assert torch.equal(got, full), "gather_limbs mismatch"
m = sharding.max_over_ranks(1.0 + rank, device="cpu")
assert m == 2.0, m
dist.barrier()
dist.destroy_process_group()
os.write(1, ("rank" + str(rank) + "-ok\n").encode())   # one atomic write per rank
'''


def test_two_rank_gloo_gather_and_timing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank0-ok" in r.stdout and "rank1-ok" in r.stdout
