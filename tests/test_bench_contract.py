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
    assert line["warmup"] == 1      # the driver's K / W are honoured (same_steps), on the full 16384^2 grid (same_config)
    assert line["config"]["global_grid"] == "16384x16384" and "16384x16384" in line["cpu_baseline"]["sample"]
    hyd = line["workloads"]["hydro"]
    assert hyd["value"] > 0 and hyd["dtype"] == "f64" and "1024x1024" in hyd["cpu_baseline"]["sample"]


def test_reference_arm_forces_the_openmp_thread_count():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm must still use every core it may run on and
    report the count OpenMP really uses."""
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "0", env={"OMP_NUM_THREADS": "1", "OM_BENCH_TEST_SAMPLE": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.strip().split("\n") if l.startswith("{")][0])
    want = len(os.sched_getaffinity(0))
    assert line["omp_threads"] == want == line["cpu_baseline"]["cores"]


def test_config_is_the_same_dict_for_both_arms():
    sys.path.insert(0, ROOT)
    import bench
    for n in (1, 2, 8):
        assert bench.config_for("life", n)["global_grid"] == f"16384x{16384 * n}"
        assert bench.config_for("hydro32k", max(n, 2))["global_grid"] == "32768x32768"


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


def test_traffic_stamps_belong_to_the_committed_kernel_sources():
    """`roofline.traffic` is reported only while the ncu capture's source hash matches the loaded kernel: the committed table must
    match the committed generated sources (a kernel change without a new capture turns the field into null, not into a stale number)."""
    import hashlib
    gen = os.path.join(ROOT, "paraiso_b200", "_generated")
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        table = json.load(f)
    for key, d, name in (("om_Life_proceed_stage0", "Life_CC", "Life"), ("om_Hydro_proceed_stage1_fast", "Hydro_OO_Double_fast", "Hydro"),
                         ("om_Hydro_proceed_stage1_exact", "Hydro_OO_Double", "Hydro")):
        with open(os.path.join(gen, d, f"{name}_kernels.cu"), "rb") as f, open(os.path.join(gen, d, "om_runtime.cuh"), "rb") as r:
            h = hashlib.sha1(f.read() + r.read()).hexdigest()[:16]
        assert table[key]["kernel_source_sha1_16"] == h, key
        assert 0.95 < table[key]["ratio"] < 1.1            # DRAM traffic = algorithmic bytes: nothing is read twice
        assert os.path.exists(os.path.join(ROOT, table[key]["source"]))
