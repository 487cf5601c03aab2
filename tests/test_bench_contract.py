"""bench.py's CPU-runnable legs: the reference arm prints one JSON line with the contract's keys (the driver
runs `bench.py --impl reference` next to the GPU arm), ranks other than 0 stay silent, and the GPU arm refuses to run
without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          env=dict(os.environ, **(env or {})), timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _bench("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "Gcell-updates/s" and line["unit"] == "Gcell/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "Life 16384x16384" in line["config"]["workload"] and line["dtype"] == "i32"


def test_reference_arm_other_ranks_print_nothing():
    r = _bench("--impl", "reference", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CUDA present")
    r = _bench("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
