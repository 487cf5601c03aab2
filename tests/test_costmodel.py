"""paraiso_b200.costmodel: the static estimate that prunes the schedule search.  The recorded points
(profiles/r1j_costmodel.json: static inputs from nvcc / cuobjdump, times measured on the B200) pin the formula; the SASS
parser is exercised on the built Hydro library when cuobjdump is present."""
import json
import os
import shutil

import pytest

from paraiso_b200 import costmodel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _points():
    with open(os.path.join(ROOT, "profiles", "r1j_costmodel.json")) as f:
        return json.load(f)["points"]


def test_model_ranks_the_measured_sweep():
    pts = _points()
    pred = [costmodel.cycles_per_cell(p["instructions"], p["fp64"], p["local"], p["threads"], p["ctas_per_sm"],
                                      p["cells_per_thread"], p["overhead"]) for p in pts]
    meas = [p["measured_ms"] for p in pts]
    assert costmodel.spearman(pred, meas) >= 0.75
    top = lambda v, k: set(sorted(range(len(v)), key=lambda i: v[i])[:k])
    assert top(pred, 3) == top(meas, 3)                      # pruning to 3 candidates keeps the three measured best
    assert min(range(len(meas)), key=lambda i: meas[i]) in top(pred, 3)
    # absolute scale: cycles per cell -> milliseconds for 4096^2 on 148 SMs x 4 schedulers at 1.965 GHz
    for c, p in zip(pred, pts):
        ms = 4096 * 4096 * c / (costmodel.SMS * 4) / costmodel.SM_CLOCK_HZ * 1e3
        assert abs(ms - p["predicted_ms"]) < 1e-6 * ms
        assert 0.75 < ms / p["measured_ms"] < 1.25


def test_resident_ctas():
    assert costmodel.resident_ctas(134, 128, 54560) == 3      # the Hydro fast build (profiles/r1i_hydro_fast_ncu.txt: 3 / 3)
    assert costmodel.resident_ctas(128, 256, 114 * 1024) == 1 or costmodel.resident_ctas(128, 256, 113 * 1024) == 2
    assert costmodel.resident_ctas(56, 128, 20 * 1024) == 9   # Life: register-limited to 9 (profiles/r1c_life_ncu.txt)


def test_spearman():
    assert costmodel.spearman([1, 2, 3, 4], [10, 20, 30, 40]) == pytest.approx(1.0)
    assert costmodel.spearman([1, 2, 3, 4], [4, 3, 2, 1]) == pytest.approx(-1.0)


@pytest.mark.skipif(not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")), reason="cuobjdump not available")
def test_estimate_of_the_built_hydro_library():
    from paraiso_b200.machines import build_hydro
    desc, so = build_hydro(fast=True, verbose=True)
    e = costmodel.estimate_stage(desc, so, size=(4096, 4096))
    assert e.symbol == "om_Hydro_proceed_stage1" and e.threads == 128 and e.ctas_per_sm == 3
    assert 1300 < e.instructions < 1700 and 600 < e.fp64 < 800 and e.local == 0      # round 2: 1 469 (716) after the ghost-write code left
    assert 0.85 < e.ms / 1.15 < 1.2                          # measured kernel time: 1.15-1.16 ms (profiles/r2s_bench_n1.json)
