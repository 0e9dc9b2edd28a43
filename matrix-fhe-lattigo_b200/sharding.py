"""Limb-wise sharding of an RNS ring over the GPUs of one box.

Every (polynomial x limb) transform and every element-wise multiply is independent
(ring/ntt.go:127-131 is a loop over limbs with no cross-limb data), so limb l goes to rank
l mod G and each rank builds a Ring over its own primes only: no collective on the NTT /
MulCoeffs path.  torch.distributed is used for rendezvous, barriers and the max-over-ranks timing.
"""
import os


def limbs_of_rank(L, world_size, rank):
    """Indices of the limbs rank `rank` owns: l = rank, rank+G, rank+2G, ..."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return list(range(rank, L, world_size))


def owner_of_limb(limb, world_size):
    return limb % world_size


def shard_moduli(moduli, world_size, rank, roots=None):
    idx = limbs_of_rank(len(moduli), world_size, rank)
    return idx, [moduli[i] for i in idx], ([roots[i] for i in idx] if roots is not None else None)


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when not launched by it."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun env (nccl on GPUs, gloo otherwise)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def max_over_ranks(value, device=None):
    """All-reduce(MAX) of a python float (the contract's 'max over ranks' timing)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_limbs(local, L, world_size, rank):
    """All-gather limb-sharded data [batch][L_local][N] -> [batch][L][N] on every rank (the exchange
    step the key-switch basis extension needs).  Requires L % world_size == 0."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return local
    assert L % world_size == 0, "gather_limbs needs the same number of limbs on every rank"
    parts = [torch.empty_like(local) for _ in range(world_size)]
    dist.all_gather(parts, local.contiguous())
    out = torch.empty((local.shape[0], L, local.shape[2]), dtype=local.dtype, device=local.device)
    for r in range(world_size):
        out[:, r::world_size, :] = parts[r]
    return out
