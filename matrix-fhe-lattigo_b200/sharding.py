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


def block_of_rank(n, world_size, rank):
    """Contiguous share [lo, hi) of n limbs for `rank` (balanced; empty when n < world_size for the last ranks)."""
    return (rank * n) // world_size, ((rank + 1) * n) // world_size


def all_gather_limbs(t, lo_hi_of_rank, world_size):
    """In-place all-gather of limb blocks of a [batch][L][N] tensor: rank r contributes t[:, lo_r:hi_r, :].
    Blocks may have different sizes (padded to the largest for the collective)."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return
    spans = [lo_hi_of_rank(r) for r in range(world_size)]
    width = max(hi - lo for lo, hi in spans)
    if width == 0:
        return
    rank = dist.get_rank()
    lo, hi = spans[rank]
    send = torch.zeros((t.shape[0], width, t.shape[2]), dtype=t.dtype, device=t.device)
    send[:, : hi - lo] = t[:, lo:hi]
    recv = [torch.empty_like(send) for _ in range(world_size)]
    dist.all_gather(recv, send)
    for r, (a, b) in enumerate(spans):
        if r != rank and b > a:
            t[:, a:b] = recv[r][:, : b - a]


def gadget_product_sharded(ks, ring, levelQ, levelP, cx_t, key, out0_t, out1_t, rank, world_size):
    """rlwe.Evaluator.GadgetProduct with the limbs sharded over `world_size` GPUs (BASELINE config 5).

    Every rank holds the QP ring; rank r owns the contiguous Q block block_of_rank(levelQ+1, G, r) and P block
    block_of_rank(levelP+1, G, r).  cx_t / out*_t: [batch][LQ+LP][N] int64 CUDA tensors; on entry the owned Q limbs
    of cx_t are valid (NTT domain), on exit the owned Q limbs of out0_t / out1_t hold the result.  `key`: b200 Poly
    with the evaluation key (only the owned limbs are read).  Two NCCL all-gathers: the coefficient-domain input
    limbs (levelQ+1 limbs per polynomial) and the P limbs of both accumulators."""
    import torch
    from .ring import Poly
    LQ = ks.LQ
    q0, q1 = block_of_rank(levelQ + 1, world_size, rank)
    p0, p1 = block_of_rank(levelP + 1, world_size, rank)
    cxinv_t, acc0_t, acc1_t = torch.empty_like(cx_t), torch.empty_like(cx_t), torch.empty_like(cx_t)
    cx, cxinv, acc0, acc1, out0, out1 = (Poly.from_torch(ring, t) for t in (cx_t, cxinv_t, acc0_t, acc1_t, out0_t, out1_t))
    ring.NTTLimbs(cx, cxinv, q0, q1 - q0, inverse=True)
    ring.Sync()
    all_gather_limbs(cxinv_t, lambda r: block_of_rank(levelQ + 1, world_size, r), world_size)
    torch.cuda.current_stream().synchronize()
    ks.DecomposeMac(levelQ, levelP, cx, cxinv, key, acc0, acc1, q0, q1, p0, p1)
    for acc in (acc0, acc1):
        ring.NTTLimbs(acc, acc, LQ + p0, p1 - p0, inverse=True)
    ring.Sync()
    for t in (acc0_t, acc1_t):
        all_gather_limbs(t, lambda r: tuple(LQ + x for x in block_of_rank(levelP + 1, world_size, r)), world_size)
    torch.cuda.current_stream().synchronize()
    ks.ModDownLimbs(levelQ, levelP, acc0, out0, q0, q1)
    ks.ModDownLimbs(levelQ, levelP, acc1, out1, q0, q1)
    ring.Sync()
    return q0, q1


class ShardedKeySwitch:
    """gadget_product_sharded with both exchanges fused into the producing kernels.

    Every rank keeps three library-allocated work polynomials (coefficient-domain input, two accumulators), exports
    them with CUDA IPC once, and maps the peers' copies.  The inverse transforms that precede the two exchanges of the
    path then store their output into every rank's copy over NVLink from inside their last kernel
    (b200_intt_limbs_gather); what is left of the collectives is a one-word all-reduce per exchange that orders the
    peers' reads after the writes.  Those two barriers also protect the buffers against the next call (a rank cannot
    pass barrier k before every rank has finished the work it queued ahead of barrier k).

    Stream contract: the barrier is a torch.distributed all-reduce, which torch orders against torch's CURRENT stream,
    while the library's kernels (gather transforms, DecomposeMac, ModDownLimbs) run on the ring's stream.  Every call of
    gadget_product therefore binds the ring to torch.cuda.current_stream() first (draining the stream the ring was on
    if it differs), so the peer stores of one stream are what the barrier waits for -- also under
    `with torch.cuda.stream(s):`."""

    def __init__(self, ks, ring, batch, rank, world_size):
        import torch
        import torch.distributed as dist
        self.ks, self.ring, self.rank, self.world = ks, ring, rank, world_size
        self.cxinv, self.acc0, self.acc1 = ring.NewPoly(batch), ring.NewPoly(batch), ring.NewPoly(batch)
        ring.Sync()
        mine = [p.IpcHandle() for p in (self.cxinv, self.acc0, self.acc1)]
        everyone = [None] * world_size
        if world_size > 1:
            dist.all_gather_object(everyone, mine)
        else:
            everyone[0] = mine
        self._mapped = [[ring.OpenPeer(everyone[r][i]) for r in range(world_size) if r != rank] for i in range(3)]
        self._flag = torch.zeros(1, dtype=torch.int32, device="cuda:%d" % ring.device)
        self._stream = None

    def _bind_stream(self):
        """Run the library's kernels on the stream the collectives are ordered against."""
        import torch
        cur = torch.cuda.current_stream(self.ring.device).cuda_stream
        if cur != self._stream:
            self.ring.Sync()               # whatever was queued on the ring's previous stream is complete before the switch
            self.ring.SetStream(cur)
            self._stream = cur

    def _barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self._flag)   # stream-ordered: later kernels on this stream wait for every rank's earlier ones

    def gadget_product(self, levelQ, levelP, cx, key, out0, out1):
        """cx: owned Q limbs valid (NTT domain); out0/out1: the owned Q limbs receive the result."""
        ks, ring, LQ = self.ks, self.ring, self.ks.LQ
        self._bind_stream()
        q0, q1 = block_of_rank(levelQ + 1, self.world, self.rank)
        p0, p1 = block_of_rank(levelP + 1, self.world, self.rank)
        ring.INTTLimbsGather(cx, self.cxinv, q0, q1 - q0, self._mapped[0])
        self._barrier()
        ks.DecomposeMac(levelQ, levelP, cx, self.cxinv, key, self.acc0, self.acc1, q0, q1, p0, p1)
        ring.INTTLimbsGather(self.acc0, self.acc0, LQ + p0, p1 - p0, self._mapped[1])
        ring.INTTLimbsGather(self.acc1, self.acc1, LQ + p0, p1 - p0, self._mapped[2])
        self._barrier()
        ks.ModDownLimbs(levelQ, levelP, self.acc0, out0, q0, q1)
        ks.ModDownLimbs(levelQ, levelP, self.acc1, out1, q0, q1)
        return q0, q1

    def close(self):
        self.ring.Sync()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        for group in self._mapped:
            for ptr in group:
                self.ring.ClosePeer(ptr)
        self._mapped = [[], [], []]
