"""bench.py's reference arm (the CPU oracle port timed on the host cores) runs without a GPU: check the JSON contract of
its line, alone and launched like the driver does for N > 1 (torchrun: rank 0 alone prints, the other ranks exit 0)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def _check(line, n):
    assert KEYS <= set(line), KEYS - set(line)
    assert line["impl"] == "reference" and line["n_gpus"] == n and line["steps"] == 2 and line["warmup"] == 1
    assert line["unit"] == "limb-transforms/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["dtype"] == "u64" and line["vs_baseline"] is None and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "limb-transforms/step" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(line["ms_per_step"] * 1e-3 * line["value"] - int(cb["sample"].split("(")[1].split()[0])) < 1e-6 * line["value"] + 1


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _lines(r.stdout)
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_prints_once():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29619", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = _lines(r.stdout)
    assert len(lines) == 1, r.stdout
    _check(lines[0], 2)
