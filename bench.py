#!/usr/bin/env python3
"""bench.py -- headline benchmark: 3N-NTT limb-transforms/s at N = 3*2^15, L = 24 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the CPU path, all host threads
    torchrun ... bench.py --gpus N ...                             # one rank per GPU, limb-sharded

One *step* = forward NTT then inverse NTT, in place, of the rank's whole resident batch
([polys][L_local][N] uint64, 2.25 GiB, far larger than the 126 MB L2, so every step streams from
HBM).  A limb-transform is one length-N transform of one RNS limb of one polynomial; a step is
2 * polys * L_local of them per rank.  Limb l lives on rank l mod G (no collective on this path);
the polynomial count grows with G so per-GPU work is fixed (weak scaling).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 3 * 2**15
MODULI = [0x1000000000248001, 0x1000000000b00001, 0x1000000000cf8001, 0x1000000000ef0001,
          0x1000000001568001, 0x10000000019a0001, 0x1000000001be0001, 0x1000000001cb8001,
          0x1000000001eb0001, 0x10000000026d8001, 0x1000000002720001, 0x1000000002888001,
          0x1000000002c30001, 0x1000000002c78001, 0x1000000002eb8001, 0x10000000030b0001,
          0x1000000003260001, 0x1000000003530001, 0x1000000003650001, 0x1000000003698001,
          0x1000000003e30001, 0x1000000003e78001, 0x1000000003ec0001, 0x10000000042b0001]   # Find3NRNSPrimes(N, 60, 24)
L = len(MODULI)
POLYS_PER_GPU = 128
BYTES_PER_UNIT = 2 * N * 8            # algorithmic: read the limb once, write it once
MODMULS_PER_UNIT = N // 2 + N + (N // 2) * 14       # cyclotomic layer + radix-3 layer + 14 radix-2 layers = 835 584
BUTTERFLY_LOOP_PEAK = 1083.0                          # G butterflies/s per B200, measured (see integer_roofline below)
METRIC = "3N-NTT limb-transforms/sec at N=3*2^15, L=24"
UNIT = "limb-transforms/s"
WORKLOAD = "config3: 3N ring N=98304 (a=15,b=1), L=24 61-bit primes, batched NTT+INTT in place, limb-sharded"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line).  The sampler is started
    early (nvidia-smi needs a few hundred ms to come up) and samples are filtered to [mark_start, mark_end]."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        samples = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                samples.append((ts, float(r[1]), float(r[2]), [v.strip().lower().startswith("active") for v in r[5:9]]))
            except Exception:
                continue
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [x for x in samples if self.t0 is not None and self.t0 - 0.02 <= x[0] <= self.t1 + 0.02]
        where = "timed region"
        if not inside:   # region shorter than the sampling period: take the samples closest to it
            mid = 0.5 * ((self.t0 or samples[-1][0]) + (self.t1 or samples[-1][0]))
            inside = sorted(samples, key=lambda x: abs(x[0] - mid))[:3]
            where = "nearest to timed region"
        sm = sorted(x[1] for x in inside)
        reasons = set()
        for x in inside:
            for name, on in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), x[3]):
                if on:
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(x[2] for x in inside), "reasons": sorted(reasons),
                "samples": len(inside), "window": where}


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_run(min_seconds, steps=None, warmup=0, threads=0):
    """The oracle port (oracle/ref_ring.c fast 3N transform; the reference's own O(N^2)/O(N^3) path cannot
    run at this size) on the host cores: NTT+INTT over a bounded sample."""
    import numpy as np
    from oracle import oracle as O
    O.build()
    T = threads or O.num_threads()
    rings = [O.ThreeNSubRing(N, q) for q in MODULI]
    polys = max(1, (16 * T + L - 1) // L)                     # >= 16 units per thread and direction: a step is ~0.1 s, long enough to amortise the thread start-up
    data = np.empty((polys * L, N), dtype=np.uint64)
    for u in range(polys * L):
        data[u] = O.fill_uniform(0xB2000003, u // L, u % L, MODULI[u % L], N)
    ref = data[:2].copy()
    units_per_step = 2 * polys * L

    def step():
        O.batch_transform(rings, data, inverse=False, threads=T)
        O.batch_transform(rings, data, inverse=True, threads=T)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    n = 0
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if steps is not None and n >= steps:
            break
        if steps is None and dt >= min_seconds and n >= 2:
            break
    assert (data[:2] == ref).all(), "CPU round trip mismatch"
    return {"value": n * units_per_step / dt, "unit": UNIT, "cores": T, "kind": "port",
            "sample": "%d steps of NTT+INTT over %d polys x %d limbs (%d limb-transforms/step), %.1f s" % (n, polys, L, units_per_step, dt),
            "ms_per_step": 1e3 * dt / n, "steps": n}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    r = cpu_run(0, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": {"workload": WORKLOAD, "arm": "CPU oracle port, all host threads"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
def run_e2e(args, ring, data, polys, Ll, world, units_per_step, barrier, dev):
    import torch
    import b200ring
    from b200ring import sharding
    e2e_steps = max(1, min(args.steps, 3))
    host_in = torch.empty((polys, Ll, N), dtype=torch.int64).pin_memory()
    host_out = torch.empty((polys, Ll, N), dtype=torch.int64).pin_memory()
    host_in.copy_(data.cpu())
    nwords = polys * Ll * N

    def e2e_step():   # the public host-buffer call: chunked H2D -> NTT -> INTT -> D2H, copies overlapped with compute
        ring.TransformHost(host_in, host_out, ["NTT", "INTT"])

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = sharding.max_over_ranks(time.perf_counter() - t0, device=dev)
    assert torch.equal(host_out, host_in), "e2e round trip mismatch"
    return {"value": world * e2e_steps * units_per_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nwords * 8,
           "d2h_bytes_per_step": nwords * 8, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps}



def run_gpu(args):
    import numpy as np
    import torch
    import b200ring
    from b200ring import sharding

    rank, local_rank, world = sharding.init_process_group()
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: b200ring has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    idx, moduli, _ = sharding.shard_moduli(MODULI, world, rank)
    Ll = len(moduli)
    polys = POLYS_PER_GPU * world
    sampler = ClockSampler(local_rank) if rank == 0 else None   # started early; samples filtered to the timed region
    ring = b200ring.NewRing(N, moduli, device=local_rank)
    ring.SetStream(torch.cuda.current_stream().cuda_stream)

    # synthetic uniform coefficients, generated in place on the device (limb l < q_l)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0xB2000003 + rank)
    data = torch.empty((polys, Ll, N), dtype=torch.int64, device=dev)
    for l, q in enumerate(moduli):
        data[:, l, :] = torch.randint(0, q, (polys, N), dtype=torch.int64, device=dev, generator=gen)
    poly = b200ring.Poly.from_torch(ring, data)
    checksum0 = int(data.sum().item())
    first = data[0, 0].clone()
    units_per_step = 2 * polys * Ll

    def step():
        ring.NTT(poly, poly)
        ring.INTT(poly, poly)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # parity evidence inside the run: forward transform of one limb against the oracle (rank 0 only)
    verified = None
    if rank == 0 and not args.no_verify:
        from oracle import oracle as O
        ring.NTT(poly, poly)
        got = data[0, 0].cpu().numpy().astype(np.uint64)
        ring.INTT(poly, poly)
        want = O.ThreeNSubRing(N, moduli[0]).NTT(first.cpu().numpy().astype(np.uint64))
        verified = bool((got == want).all())
        if not verified:
            raise SystemExit("bench: CUDA NTT differs from the oracle")

    for _ in range(args.warmup):
        step()
    barrier()
    if sampler:
        sampler.mark_start()
    launches0 = ring.LaunchCount()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    if sampler:
        sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = ring.LaunchCount() - launches0
    clocks = sampler.stop() if sampler else None
    ms = sharding.max_over_ranks(ms, device=dev)
    assert int(data.sum().item()) == checksum0 and torch.equal(data[0, 0], first), "round trip changed the data"
    value = world * args.steps * units_per_step / (ms * 1e-3)

    # per-kernel durations (CUDA events on the launching stream) for the roofline of the dominant kernel
    prof = ring.Profile(lambda: [step() for _ in range(max(1, min(args.steps, 3)))])
    peak, peak_src = peaks()
    dom = max(prof, key=lambda k: prof[k]["ms"]) if prof else None
    roofline = None
    if dom:
        d = prof[dom]
        per_launch_bytes = d["units"] / d["launches"] * BYTES_PER_UNIT      # each pass reads + writes every word once
        achieved = per_launch_bytes / (d["ms"] / d["launches"] * 1e-3) / 1e9
        traffic = None
        try:   # DRAM bytes per launch from the committed ncu --set full capture of this command (profiles/)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            traffic = tj["dram_bytes_per_unit"][dom] * d["units"] / d["launches"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "algorithmic_bytes": per_launch_bytes, "peak_source": peak_src, "launch_us": 1e3 * d["ms"] / d["launches"],
                    "share_of_step": d["ms"] / sum(v["ms"] for v in prof.values()),
                    "kernels": {k: {"ms": v["ms"], "launches": v["launches"]} for k, v in prof.items()}}
    path_gbs = value / world * BYTES_PER_UNIT / 1e9
    path = {"achieved": path_gbs, "peak": peak, "unit": "GB/s per GPU", "frac": path_gbs / peak,
            "note": "whole transform (both passes) against 16*N algorithmic bytes per limb-transform"}
    # the bound that actually binds: 64-bit modular butterflies on the integer pipes.  The ceiling is
    # the register-resident loop of the same fold butterflies measured with tools/solbench on this pool's B200s
    # (profiles/r02_solbench.txt: 1 083 G butterflies/s; the Shoup/Harvey butterfly of round 1 reaches 919), not a datasheet number.
    gbf = value / world * MODMULS_PER_UNIT / 1e9
    integer = {"bound": "integer pipes", "achieved": gbf, "peak": BUTTERFLY_LOOP_PEAK, "unit": "G modular butterflies/s per GPU",
               "frac": gbf / BUTTERFLY_LOOP_PEAK, "modmuls_per_limb_transform": MODMULS_PER_UNIT,
               "peak_source": "tools/solbench register-resident loop of the fold butterfly, profiles/r02_solbench.txt"}

    # end to end through the C ABI with HOST buffers: pinned H2D + NTT + INTT + D2H inside the timed region
    e2e = None
    nwords = polys * Ll * N
    if not args.no_e2e:
        e2e = run_e2e(args, ring, data, polys, Ll, world, units_per_step, barrier, dev)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_run(args.cpu_seconds)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "polys_per_gpu": POLYS_PER_GPU, "polys": polys, "limbs_per_gpu": Ll,
                           "limb_transforms_per_step": world * units_per_step, "resident_bytes_per_gpu": nwords * 8,
                           "l2": "inputs (2.25 GiB per GPU) are larger than L2, no explicit flush",
                           "parallelism": "limb-sharded x%d, no collective" % world},
                "roofline": roofline, "path_roofline": path, "integer_roofline": integer, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "clocks": clocks, "verified_against_oracle": verified}
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL prints its version banner on stdout at communicator creation) must not add lines to the ONE
    JSON line of the contract: route fd 1 to stderr for the run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", 1)))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (ncu runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
