"""CPU tests of bench.py's contract: the reference arm (the oracle port on the host cores) prints one valid JSON line with the
keys the driver reads, and the product arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("collocation points/sec per train step")
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("C5 ")                       # same workload as the product arm's default
    assert d["config"]["points_per_gpu"] == 8388608                        # the per-GPU share of the 64M-point config
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_uses_every_core_and_the_same_sample_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must still use all host cores and the same
    bounded sample at every N (round 1: the N > 1 lines ran on one thread with a different sample)."""
    def line(gpus, env):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus), "--steps", "1",
                            "--warmup", "3", "--config", "2"], capture_output=True, text=True, timeout=300, env=e)
        assert r.returncode == 0, r.stderr[-2000:]
        return json.loads([ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")][0])
    a = line(1, {})
    b = line(8, {"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "8"})
    ncpu = len(os.sched_getaffinity(0))
    assert a["cpu_baseline"]["cores"] == ncpu and b["cpu_baseline"]["cores"] == ncpu
    assert a["cpu_baseline"]["sample"].split(" x ")[0] == b["cpu_baseline"]["sample"].split(" x ")[0]
    # every other rank exits 0 without work and prints nothing
    e = dict(os.environ); e.update({"RANK": "3", "WORLD_SIZE": "8"})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8"], capture_output=True, text=True,
                       timeout=120, env=e)
    assert r.returncode == 0 and not r.stdout.strip()


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "2", "--warmup", "3")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]   # no number is printed
