"""The generated CUDA kernels, compiled for host threads (tests/emu), against the oracle.  This checks the
schedule logic (rings, lags, phases, ghost writes, reductions, both skeletons) on a machine without a GPU;
the GPU parity tests proper are in test_gpu_parity.py."""
import os

import numpy as np
import pytest

from oracle.cpu import OracleMachine
from paraiso_b200.examples.helloworld import helloworld_om, helloworld_setup
from paraiso_b200.examples.hydro import hydro_om, hydro_setup
from paraiso_b200.examples.life import life_om, life_setup
from paraiso_b200.examples.shiftexample import shiftexample_om, shiftexample_setup
from paraiso_b200.runtime import Machine
from tests.emu.build_emu import build_emulated


def _life(size, steps, mode, window=True, staging="cp_async", cold=False, clean=False, warp=False):
    setup = life_setup("master", size=size)
    setup.tuning.skeleton = mode
    setup.tuning.row_window = window
    setup.tuning.staging = staging
    setup.tuning.cold_rare = cold
    setup.tuning.clean_ctas = clean
    setup.tuning.warp_rings = warp
    desc, so = build_emulated(setup, life_om("master"), tag=f"Life_{mode}_{int(window)}_{staging}" + ("_cold" if cold else "") + ("_clean" if clean else "") + ("_warp" if warp else ""))
    with open(os.path.join(os.path.dirname(so), "Life_kernels.cu")) as f:
        src = f.read()
    assert ("register streaming" in src) == (mode == "stream")
    assert ("stencil window (rotates by renaming)" in src) == (mode == "ring" and window)
    assert ("om_bulk_g2s" in src) == (staging == "bulk")
    assert ("const OmRare om_rr = [=]() __attribute__((noinline))" in src) == cold
    assert ("const bool cta_rare = __syncthreads_or(rare);" in src) == clean
    assert ("__syncwarp();   // the row's segment was staged by this warp's own lanes" in src) == warp
    m = Machine(desc, so, size=size, device="cpu", _emulated=True)
    o = OracleMachine(life_setup("master", size=size), life_om("master"))
    init = (np.random.default_rng(7).random((size[1], size[0])) < 0.35).astype(np.int32)
    m.call("init"); o.call("init")
    m.set("cell", init); o.interior("cell")[...] = init
    for t in range(steps):
        m.call("proceed"); o.call("proceed")
        assert np.array_equal(m.get("cell"), o.interior("cell")), t
        assert int(m.scalar("population")) == int(o.scalar("population")[0])
    assert int(m.scalar("generation")) == steps


@pytest.mark.parametrize("size", [(80, 48), (5, 3), (513, 40), (1030, 37), (1, 1), (2, 2), (1, 7), (7, 1), (80, 100), (4, 90), (600, 130)])
def test_life_ring_skeleton(size):
    """(includes domains narrower than the ghost width: every neighbour of a 1x1 Cyclic grid is the cell itself; the tall ones
    have chunks without a y wrap, where the edge strips' ghost columns go through the lean form of the rare block)"""
    _life(size, 4, "ring")


@pytest.mark.parametrize("size", [(80, 48), (1030, 37), (2, 2)])
def test_life_with_the_rare_block_out_of_line(size):
    """Tuning.cold_rare: partial vectors and ghost copies of a row go through a noinline closure that returns the accumulators."""
    _life(size, 4, "ring", cold=True)


@pytest.mark.parametrize("size", [(80, 48), (1600, 70), (2, 2)])
def test_life_with_row_bodies_for_ctas_without_a_rare_block(size):
    """Tuning.clean_ctas: CTAs that are neither an edge strip nor a chunk with a y wrap (the 1600-wide grid has them) run copies of
    the row bodies without the rarely taken block."""
    _life(size, 4, "ring", clean=True)


@pytest.mark.parametrize("size", [(80, 48), (1030, 37), (2, 2), (600, 130), (1600, 70)])
def test_life_with_warp_private_rings(size):
    """Tuning.warp_rings: every warp stages its own segment of a ring row (its own copy of the pads included) and the row loop
    has no CTA barrier."""
    _life(size, 4, "ring", warp=True)


@pytest.mark.parametrize("size", [(80, 48), (513, 40)])
def test_life_register_streaming_skeleton(size):
    _life(size, 4, "stream")


@pytest.mark.parametrize("size,window", [((80, 48), True), ((513, 40), True), ((1030, 37), False)])
def test_life_ring_skeleton_with_tma_bulk_staging(size, window):
    _life(size, 5, "ring", window=window, staging="bulk")


@pytest.mark.parametrize("size", [(80, 48), (513, 40)])
def test_life_ring_skeleton_without_row_window(size):
    _life(size, 4, "ring", window=False)


def test_hydro_master_double_bit_identical():
    size = (64, 48)
    desc, so = build_emulated(hydro_setup(size), hydro_om("master"))
    m = Machine(desc, so, size=size, device="cpu", _emulated=True)
    o = OracleMachine(hydro_setup(size), hydro_om("master"))
    for k, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / size[0], dR1=1.0 / size[1]).items():
        m.set_scalar(k, v)
        o.scalar(k)[0] = v
    m.call("init"); o.call("init")
    names = ["density", "velocity0", "velocity1", "pressure"]
    for n in names:
        assert np.array_equal(m.get(n, with_margin=True).view(np.uint64), o.array(n).view(np.uint64))
    for t in range(2):
        m.call("proceed"); o.call("proceed")
    for n in names:
        assert np.array_equal(m.get(n, with_margin=True).view(np.uint64), o.array(n).view(np.uint64)), n
    assert m.scalar("time") == o.scalar("time")[0]


def _hydro_emulated(size, steps, fast=False, prefetch=None, tag=None):
    setup = hydro_setup(size, fast=fast)
    if prefetch is not None:
        setup.tuning.direct_prefetch = prefetch
    desc, so = build_emulated(setup, hydro_om("master"), tag=tag)
    m = Machine(desc, so, size=size, device="cpu", _emulated=True)
    o = OracleMachine(hydro_setup(size), hydro_om("master"))
    for k, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / size[0], dR1=1.0 / size[1]).items():
        m.set_scalar(k, v)
        o.scalar(k)[0] = v
    m.call("init"); o.call("init")
    for t in range(steps):
        m.call("proceed"); o.call("proceed")
    return m, o, so


@pytest.mark.parametrize("size", [(1, 1), (129, 5)])
def test_hydro_ragged_and_tiny_grids_bit_identical(size):
    """Grids narrower than the margin (3), than one vector (2 doubles) and than one CTA strip, odd widths: every array
    including its margin cells and the CFL time step equal the oracle bit for bit (offline also (5,3) (2,3) (3,70) (260,33)
    (4,4) (7,1) (1,9))."""
    m, o, _so = _hydro_emulated(size, 2)
    for n in ["density", "velocity0", "velocity1", "pressure"]:
        assert np.array_equal(m.get(n, with_margin=True).view(np.uint64), o.array(n).view(np.uint64)), n
    assert m.scalar("time") == o.scalar("time")[0]


def test_hydro_register_prefetch_of_unstaged_inputs_bit_identical():
    """Tuning.direct_prefetch: inputs read at column offset 0 are loaded one row ahead at the loop top."""
    m, o, so = _hydro_emulated((70, 37), 2, prefetch=True, tag="Hydro_pf")
    with open(os.path.join(os.path.dirname(so), "Hydro_kernels.cu")) as f:
        assert "next row to prefetch" in f.read()
    for n in ["density", "velocity0", "velocity1", "pressure"]:
        assert np.array_equal(m.get(n, with_margin=True).view(np.uint64), o.array(n).view(np.uint64)), n
    assert m.scalar("time") == o.scalar("time")[0]


def test_hydro_flipped_materialisation_genes_bit_identical():
    """Tuning.mat_flip (the per-node Manifest/Delayed genes the schedule search flips): recomputing a value at every
    cursor instead of keeping it in a shared-memory ring must not change a single bit."""
    from paraiso_b200.generator.b200.emit import describe_only
    size = (70, 37)
    genes = describe_only(hydro_setup(size), hydro_om("master"), "proceed")
    cheap = sorted([g for g in genes if g["chosen"] and g["cost"] <= 40], key=lambda g: g["vid"])
    assert len(cheap) >= 3
    setup = hydro_setup(size)
    setup.tuning.mat_flip = tuple(("proceed", g["vid"]) for g in cheap[:3])
    desc, so = build_emulated(setup, hydro_om("master"), tag="Hydro_flip")
    base_desc, _ = build_emulated(hydro_setup(size), hydro_om("master"))
    rings = lambda d: [st["rings"] for k in d["kernels"] if k["name"] == "proceed" for st in k["stages"]]
    assert rings(desc) != rings(base_desc) or desc != base_desc
    m = Machine(desc, so, size=size, device="cpu", _emulated=True)
    o = OracleMachine(hydro_setup(size), hydro_om("master"))
    for k, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / size[0], dR1=1.0 / size[1]).items():
        m.set_scalar(k, v)
        o.scalar(k)[0] = v
    m.call("init"); o.call("init")
    for t in range(2):
        m.call("proceed"); o.call("proceed")
    for n in ["density", "velocity0", "velocity1", "pressure"]:
        assert np.array_equal(m.get(n, with_margin=True).view(np.uint64), o.array(n).view(np.uint64)), n


def test_hydro_carried_dt_reduce_bit_identical_and_invalidated_by_host_writes():
    """Tuning.carry_reduces: the last stage of `proceed` also reduces dt for the next call, which then replaces its
    level-0 stage by a slot copy.  A host write to an array or a scalar in between must bring the stage back."""
    size = (70, 37)
    setup = hydro_setup(size)
    setup.tuning.carry_reduces = True
    desc, so = build_emulated(setup, hydro_om("master"), tag="Hydro_carry")
    k = [k for k in desc["kernels"] if k["name"] == "proceed"][0]
    assert k["carry"] and k["carry"]["skip_stage"] == 0 and k["carry"]["arrays"] == [7, 8, 9, 10]
    m = Machine(desc, so, size=size, device="cpu", _emulated=True)
    o = OracleMachine(hydro_setup(size), hydro_om("master"))
    for kk, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / size[0], dR1=1.0 / size[1]).items():
        m.set_scalar(kk, v)
        o.scalar(kk)[0] = v
    m.call("init"); o.call("init")
    names = ["density", "velocity0", "velocity1", "pressure"]
    launches = []
    for t in range(6):
        if t == 3:      # host write to an array: the carried dt is stale
            a = m.get("pressure", with_margin=True)
            a[20:25, 30:40] *= 4.0
            m.set("pressure", a, with_margin=True)
            o.array("pressure")[20:25, 30:40] *= 4.0
        if t == 5:      # host write to a scalar the reduce depends on
            m.set_scalar("dR0", 0.5 / size[0])
            o.scalar("dR0")[0] = 0.5 / size[0]
        before = m.launches
        m.call("proceed"); o.call("proceed")
        launches.append(m.launches - before)
        for n in names:
            assert np.array_equal(m.get(n, with_margin=True).view(np.uint64), o.array(n).view(np.uint64)), (t, n)
        assert m.scalar("time") == o.scalar("time")[0], t
    assert launches == [3, 2, 2, 3, 2, 3]     # dt stage + flux stage + scalar kernel, or without the dt stage


def test_hydro_fast_math_schedule_within_tolerance():
    """Setup.fast_math: shared reciprocals (a * (1/b)), std::max/min helpers, Goldschmidt sqrt entry points — here with
    the emulation's exact 1/b and sqrt, so the difference to the reference arithmetic is the re-association only."""
    m, o, so = _hydro_emulated((64, 48), 3, fast=True, tag="Hydro_fastmath")
    with open(os.path.join(os.path.dirname(so), "Hydro_kernels.cu")) as f:
        src = f.read()
    assert "om_frcp(" in src and "om_fmax_std(" in src and "om_fsqrt(" in src
    for n in ["density", "velocity0", "velocity1", "pressure"]:
        a, b = m.get(n), o.interior(n)
        assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)), n
    assert abs(m.scalar("time") - o.scalar("time")[0]) <= 1e-12 * o.scalar("time")[0]


def test_helloworld_known_answer():
    """examples/HelloWorld: table(x,y) = x*y on 10x20, total = 45*190 = 8550."""
    desc, so = build_emulated(helloworld_setup(), helloworld_om(), tag="Hello")
    m = Machine(desc, so, device="cpu", _emulated=True)
    m.call("create")
    t = m.get("table")
    assert t.shape == (20, 10) and all(t[y, x] == x * y for x in range(10) for y in range(20))
    assert int(m.scalar("total")) == 8550


@pytest.mark.parametrize("cyclic", [False, True])
def test_shiftexample_known_answer(cyclic):
    """examples/ShiftExample: init; increment; calculate -> table[i] = 10000*t[i-1] + 100*t[i] + t[i+1], t[i] = i+1."""
    s = shiftexample_setup(cyclic)
    desc, so = build_emulated(s, shiftexample_om(), tag=f"Shift{cyclic}")
    m = Machine(desc, so, device="cpu", _emulated=True)
    for k in ("init", "increment", "calculate"):
        m.call(k)
    t = np.arange(1, 9)
    if cyclic:
        want = 10000 * np.roll(t, 1) + 100 * t + np.roll(t, -1)
        got = m.get("table").ravel()
    else:   # Open: margins hold loadIndex+1 too (index -1 -> 0, index 8 -> 9); only the interior is valid
        tt = np.arange(0, 10)
        want = 10000 * tt[:-2] + 100 * tt[1:-1] + tt[2:]
        got = m.get("table").ravel()
        assert list(m.get("table", with_margin=True).ravel()[[0, -1]]) == [0, 0]   # never written: stays 0
    assert np.array_equal(got, want)
    assert int(m.scalar("total")) == int(want.sum())
