"""Generator fuzzing: seeded random OM programs (random expression trees over loads, shifts of up to three cells in either
axis, integer arithmetic, comparisons, select, max / min, loadIndex, with a reduce feeding a second stage in some of them,
on random Open / Cyclic boundary mixes and ragged sizes) run on the emulated kernels and compared with the oracle bit for
bit — every array of the memory box, margins included, and every scalar.  The backend is a generator: whatever the Builder
can say must come out right, not only Life and Hydro."""
import random

import numpy as np
import pytest

from oracle.cpu import OracleMachine
from paraiso_b200.annotation import CYCLIC, OPEN
from paraiso_b200.generator.native import Setup
from paraiso_b200.om.builder import (StaticValue, bind, broadcast, ge, imm, load, loadIndex, lt, makeOM, max_, min_, reduce,
                                     select, shift, store)
from paraiso_b200.om.graph import ARRAY, SCALAR, Named
from paraiso_b200.runtime import Machine
from tests.emu.build_emu import build_emulated


def random_program(seed: int):
    rng = random.Random(seed)
    a = Named("a", StaticValue(ARRAY, "Int"))
    b = Named("b", StaticValue(ARRAY, "Int"))
    s = Named("s", StaticValue(SCALAR, "Int"))
    t = Named("t", StaticValue(SCALAR, "Int"))
    use_reduce = rng.random() < 0.6
    two_stage = use_reduce and rng.random() < 0.5
    plan = [rng.random() for _ in range(400)]         # the same decisions every time the builder re-runs
    red_op = rng.choice(["Sum", "Max", "Min"])

    def kernel():
        it = iter(plan)
        nxt = lambda: next(it)
        leaves = [bind(load(a)), bind(load(b))]

        def expr(depth):
            r = nxt()
            if depth == 0 or r < 0.15:
                c = nxt()
                if c < 0.6:
                    return leaves[int(nxt() * 2)]
                if c < 0.8:
                    return bind(loadIndex(int(nxt() * 2)))
                return imm(int(nxt() * 9) - 4, ARRAY, "Int")
            if r < 0.40:
                v = (int(nxt() * 7) - 3, int(nxt() * 7) - 3)
                return bind(shift(v, expr(depth - 1)))
            if r < 0.75:
                x, y = expr(depth - 1), expr(depth - 1)
                op = int(nxt() * 5)
                return bind([x + y, x - y, x * (int(nxt() * 5) - 2), max_(x, y), min_(x, y)][op])
            x, y, z = expr(depth - 1), expr(depth - 1), expr(depth - 1)
            cond = lt(x, y) if nxt() < 0.5 else ge(x, imm(int(nxt() * 20) - 10, ARRAY, "Int"))
            return bind(select(cond, y, z))
        e1 = bind(expr(3))
        e2 = bind(expr(3))
        if use_reduce:
            r = bind(reduce(red_op, e1))
            store(s, r)
            if two_stage:
                e2 = bind(e2 + broadcast(r) / 1000)
        store(t, load(t) + 1)
        store(a, e2 - (e2 / 4096) * 4096)             # keep the values small: no int overflow over the steps
        store(b, e1 - (e1 / 4096) * 4096)
    om = lambda: makeOM("Fuzz", [], [a, b, s, t], [("k", kernel)], dim=2)
    size = (rng.choice([17, 33, 70, 130]), rng.choice([5, 9, 14]))
    bnd = (rng.choice([OPEN, CYCLIC]), rng.choice([OPEN, CYCLIC]))
    return om, Setup(local_size=size, boundary=bnd)


# 76: a shifted immediate is reduced (regression, see below); 1057: Shift (-3,-3) of Shift (2,-3) of loadIndex 1 on a Cyclic
# axis of 5 rows reads row y + 6, two periods out of range (om_wrap_far: the reference's (i + n) % n, PlanTrans.hs:459-462)
@pytest.mark.parametrize("seed", list(range(12)) + [76, 1057])
def test_random_program_matches_oracle(seed):
    om, setup = random_program(seed)
    desc, so = build_emulated(setup, om(), tag=f"fuzz_{seed}")
    m = Machine(desc, so, device="cpu", _emulated=True)
    o = OracleMachine(setup, om())
    rng = np.random.default_rng(seed)
    for name in ("a", "b"):
        arr = rng.integers(-30, 30, o.array(name).shape).astype(np.int32)
        o.array(name)[...] = arr
        m.set(name, arr.reshape(m.get(name, with_margin=True).shape), with_margin=True)
    for step in range(3):
        m.call("k"); o.call("k")
        for st in desc["statics"]:
            if st["realm"] == "Array":
                got, want = m.get(st["name"], with_margin=True), o.array(st["name"])
                assert np.array_equal(got.reshape(want.shape), want), (seed, step, st["name"], setup.boundary, setup.local_size)
            else:
                assert int(m.scalar(st["name"])) == int(o.scalar(st["name"])[0]), (seed, step, st["name"])


def test_shifted_immediate_keeps_its_valid_region():
    """`shift v (imm c)` has the value c everywhere but the shrunk Valid region of a Shift (BoundaryAnalysis.hs:85-94): the
    reference writes it — and reduces it — only there; the cells outside stay 0.  (Found by seed 76: hash-consing used to
    drop the Shift of a position-independent value together with its region.)"""
    a = Named("a", StaticValue(ARRAY, "Int"))
    s = Named("s", StaticValue(SCALAR, "Int"))

    def k():
        x = bind(shift((-1, 0), shift((2, 1), imm(-2, ARRAY, "Int") * 1)))
        store(s, reduce("Sum", x))
        store(a, x + load(a) * 0)
    om = lambda: makeOM("ShiftImm", [], [a, s], [("k", k)], dim=2)
    setup = Setup(local_size=(17, 5), boundary=(OPEN, OPEN))
    desc, so = build_emulated(setup, om(), tag="fuzz_shiftimm")
    m = Machine(desc, so, device="cpu", _emulated=True)
    o = OracleMachine(setup, om())
    m.call("k"); o.call("k")
    want = o.array("a")
    assert np.array_equal(m.get("a", with_margin=True).reshape(want.shape), want)
    assert int(m.scalar("s")) == int(o.scalar("s")[0]) == int(want.sum()) != -2 * want.size


# ---- the same for Double (bit-exact: the default build has no FMA contraction and IEEE division / sqrt) and for rank 1 / 3 ----------
def random_float_program(seed: int, dim: int):
    from paraiso_b200.om.builder import abs_, sqrt
    rng = random.Random(1000 + seed)
    a = Named("a", StaticValue(ARRAY, "Double"))
    b = Named("b", StaticValue(ARRAY, "Double"))
    s = Named("s", StaticValue(SCALAR, "Double"))
    use_reduce = rng.random() < 0.6
    two_stage = use_reduce and rng.random() < 0.5
    red_op = rng.choice(["Max", "Min"])            # (a floating-point Sum is folded in another order on the device)
    plan = [rng.random() for _ in range(400)]
    reach = 3 if dim < 3 else 1

    def kernel():
        it = iter(plan)
        nxt = lambda: next(it)
        leaves = [bind(load(a)), bind(load(b))]

        def expr(depth):
            r = nxt()
            if depth == 0 or r < 0.15:
                c = nxt()
                if c < 0.75:
                    return leaves[int(nxt() * 2)]
                return imm(round(nxt() * 4 - 2, 3), ARRAY, "Double")
            if r < 0.40:
                v = tuple(int(nxt() * (2 * reach + 1)) - reach for _ in range(dim))
                return bind(shift(v, expr(depth - 1)))
            if r < 0.80:
                x, y = expr(depth - 1), expr(depth - 1)
                op = int(nxt() * 7)
                return bind([x + y, x - y, x * y * 0.25, max_(x, y), min_(x, y), x / (abs_(y) + 1.5), sqrt(abs_(x) + 0.5)][op])
            x, y, z = expr(depth - 1), expr(depth - 1), expr(depth - 1)
            return bind(select(lt(x, y), y, z))
        e1 = bind(expr(3))
        e2 = bind(expr(3))
        if use_reduce:
            r = bind(reduce(red_op, e1))
            store(s, r)
            if two_stage:
                e2 = bind(e2 + broadcast(r) * 0.125)
        store(a, max_(min_(e2, imm(8.0, ARRAY, "Double")), imm(-8.0, ARRAY, "Double")))
        store(b, e1 * 0.5)
    om = lambda: makeOM("FuzzF", [], [a, b, s], [("k", kernel)], dim=dim)
    sizes = {1: [(257,), (1000,)], 2: [(33, 9), (70, 14), (130, 5)], 3: [(20, 9, 6), (33, 5, 7)]}[dim]
    return om, Setup(local_size=rng.choice(sizes), boundary=tuple(rng.choice([OPEN, CYCLIC]) for _ in range(dim)))


@pytest.mark.parametrize("dim,seed", [(1, 0), (1, 1), (2, 0), (2, 1), (2, 2), (2, 3), (3, 0), (3, 1), (3, 2)])
def test_random_float_program_matches_oracle(dim, seed):
    om, setup = random_float_program(seed, dim)
    desc, so = build_emulated(setup, om(), tag=f"fuzzf_{dim}_{seed}")
    m = Machine(desc, so, device="cpu", _emulated=True)
    o = OracleMachine(setup, om())
    rng = np.random.default_rng(seed)
    for name in ("a", "b"):
        arr = rng.uniform(-2.0, 2.0, o.array(name).shape)
        o.array(name)[...] = arr
        m.set(name, arr.reshape(m.get(name, with_margin=True).shape), with_margin=True)
    for step in range(3):
        m.call("k"); o.call("k")
        for st in desc["statics"]:
            if st["realm"] == "Array":
                got, want = m.get(st["name"], with_margin=True), o.array(st["name"])
                assert np.array_equal(got.reshape(want.shape).view(np.uint64), want.view(np.uint64)), (dim, seed, step, st["name"], setup.boundary)
            else:
                assert np.float64(m.scalar(st["name"])).view(np.uint64) == o.scalar(st["name"]).view(np.uint64)[0], (dim, seed, step)


# ---- random programs through the generated C++ class on several emulated devices -------------------------------------------------
@pytest.mark.parametrize("seed", [3, 14, 17, 21])
def test_random_program_on_several_devices_equals_one(seed, tmp_path):
    """Slab decomposition of arbitrary stencils (asymmetric reach of up to six rows, Open and Cyclic cuts, reduces feeding a
    second stage): the generated class on 2 and 3 emulated devices prints what one device prints."""
    import os
    from tests.emu import hostclass
    from tests.generic_driver import driver_source
    om, setup = random_program(seed)
    setup.local_size = (setup.local_size[0], 24)            # tall enough for three slabs of any reach
    tag = f"fuzzdev_{seed}"
    desc, _so = build_emulated(setup, om(), tag=tag)
    drv = str(tmp_path / "driver.cpp")
    with open(drv, "w") as f:
        f.write(driver_source(desc, ["k", "k", "k"]))
    exe = str(tmp_path / "drv")
    hostclass.link_emulated(setup, om(), tag, drv, exe)
    want = hostclass.run(exe)
    assert want.strip()
    for devices in (2, 3):
        assert hostclass.run(exe, devices=devices) == want, devices


# ---- Cyclic grids narrower than their two ghost zones together (seed 1069: 130 x 5 with reach 3 + 3) ---------------------------
def _oracle_exe(setup, om, drv, tmp_path):
    """The generic driver against the oracle's reference-style class (header-only text of oracle/plantrans.py)."""
    import os
    import subprocess

    from oracle import plantrans
    from paraiso_b200.generator.plan import translate
    hdr = tmp_path / "oracle_hdr"
    hdr.mkdir(exist_ok=True)
    (hdr / f"{om.name}.hpp").write_text(plantrans.emit(translate(setup, om)))
    exe = str(tmp_path / "oracle_drv")
    subprocess.run([os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-w", "-ffp-contract=off", f"-I{hdr}", "-x", "c++", drv,
                    "-o", exe], check=True)
    return exe


def test_narrow_cyclic_grid_python_host():
    """A cell of a 5-row Cyclic axis with ghost zones of 3 + 3 rows has two images along that axis and four more in the
    corners; the kernels' fused ghost writes produce one per direction and the corner, the host redoes the wrap."""
    test_random_program_matches_oracle(1069)


def test_narrow_cyclic_grid_generated_host_class_equals_reference_style_class(tmp_path):
    import subprocess

    from tests.emu import hostclass
    from tests.generic_driver import driver_source
    om, setup = random_program(1069)
    assert setup.boundary == (CYCLIC, CYCLIC) and setup.local_size == (130, 5)
    desc, _so = build_emulated(setup, om(), tag="fuzzdev_narrow")
    assert max(desc["radius_lo"][1], desc["radius_hi"][1]) <= 5 < desc["radius_lo"][1] + desc["radius_hi"][1]
    drv = str(tmp_path / "driver.cpp")
    with open(drv, "w") as f:
        f.write(driver_source(desc, ["k", "k", "k"]))
    exe = str(tmp_path / "drv")
    hostclass.link_emulated(setup, om(), "fuzzdev_narrow", drv, exe)
    want = subprocess.run([_oracle_exe(setup, om(), drv, tmp_path)], capture_output=True, text=True, check=True).stdout
    assert want.strip() and hostclass.run(exe) == want
