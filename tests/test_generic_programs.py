"""The backend is a generator, not two hand-written kernels: synthetic OM programs that exercise corners the
two headline programs do not (several reduces in one stage, wide and asymmetric shifts, mixed boundary kinds,
float data and a float Sum, an intermediate read through shifts in both axes, a reduce feeding a later stage),
run on the emulated kernels and compared with the oracle."""
import numpy as np
import pytest

from oracle.cpu import OracleMachine
from paraiso_b200.annotation import CYCLIC, OPEN
from paraiso_b200.generator.native import Setup
from paraiso_b200.om.builder import (StaticValue, bind, broadcast, cast, imm, load, loadIndex, makeOM, max_, min_, reduce,
                                     select, shift, store, lt, sqrt)
from paraiso_b200.om.graph import ARRAY, SCALAR, Named
from paraiso_b200.runtime import Machine
from tests.emu.build_emu import build_emulated


def run_both(om_fn, setup, kernels, tag, fill, rtol=0.0):
    desc, so = build_emulated(setup, om_fn(), tag=tag)
    size = tuple(setup.local_size) + (1,) * (2 - len(setup.local_size))
    m = Machine(desc, so, device="cpu", _emulated=True)
    o = OracleMachine(setup, om_fn())
    for name, arr in fill.items():
        m.set(name, arr.reshape(m.get(name, with_margin=True).shape), with_margin=True)
        o.array(name)[...] = arr.reshape(o.array(name).shape)
    for k in kernels:
        m.call(k); o.call(k)
    for i, s in enumerate(desc["statics"]):
        if s["realm"] == "Array":
            a, b = m.get(s["name"], with_margin=True), o.array(s["name"])
            a = a.reshape(b.shape)
            if rtol:
                assert np.allclose(a, b, rtol=rtol, atol=0), s["name"]
            else:
                assert np.array_equal(a, b), s["name"]
        else:
            a, b = m.scalar(s["name"]), o.scalar(s["name"])[0]
            assert (abs(a - b) <= rtol * abs(b)) if rtol else (a == b), s["name"]
    return m, o


def mem_shape(setup, om_fn):
    from paraiso_b200.generator.plan import translate
    p = translate(setup, om_fn())
    ms = p.memory_size + (1,) * (2 - len(p.memory_size))
    return (ms[1], ms[0])


def test_three_reduces_in_one_stage_and_reduce_feeding_a_later_stage():
    a = Named("a", StaticValue(ARRAY, "Int"))
    s = Named("s", StaticValue(SCALAR, "Int"))
    mx = Named("mx", StaticValue(SCALAR, "Int"))
    mn = Named("mn", StaticValue(SCALAR, "Int"))

    def k():
        x = bind(load(a))
        y = bind(x * 3 - shift((1, 0), x))
        tot = bind(reduce("Sum", y))
        store(s, tot)
        store(mx, reduce("Max", y))
        store(mn, reduce("Min", y + 7))
        store(a, y - broadcast(tot) / 1000)        # second stage: depends on the reduce
    om = lambda: makeOM("Multi", [], [a, s, mx, mn], [("k", k)], dim=2)
    setup = Setup(local_size=(70, 9), boundary=(CYCLIC, CYCLIC))
    fill = {"a": np.random.default_rng(3).integers(-50, 50, mem_shape(setup, om)).astype(np.int32)}
    run_both(om, setup, ["k", "k"], "gen_multi", fill)


@pytest.mark.parametrize("bnd", [(CYCLIC, OPEN), (OPEN, CYCLIC), (OPEN, OPEN)])
def test_wide_asymmetric_stencil_mixed_boundaries(bnd):
    a = Named("a", StaticValue(ARRAY, "Int"))
    b = Named("b", StaticValue(ARRAY, "Int"))

    def k():
        x = bind(load(a))
        store(b, shift((5, 0), x) + 2 * shift((-6, 2), x) - shift((0, -3), x) + loadIndex(0) * 100 + loadIndex(1))
        store(a, x + 1)
    om = lambda: makeOM("Wide", [], [a, b], [("k", k)], dim=2)
    setup = Setup(local_size=(61, 23), boundary=bnd)
    fill = {"a": np.random.default_rng(4).integers(0, 1000, mem_shape(setup, om)).astype(np.int32)}
    run_both(om, setup, ["k", "k", "k"], f"gen_wide_{bnd[0][0]}{bnd[1][0]}", fill)


def test_intermediate_shifted_in_both_axes_float_with_float_sum():
    """A non-trivial intermediate (materialised in a shared-memory ring) read at x and y offsets; float arithmetic;
    a float Sum (folded in a different but fixed order on the device: compared to 1e-5)."""
    u = Named("u", StaticValue(ARRAY, "Float"))
    e = Named("e", StaticValue(SCALAR, "Float"))

    def k():
        x = bind(load(u))
        g = bind(sqrt(x * x + 1.5) / (x + 3.0))                 # expensive: becomes a MAT node
        lap = bind(shift((1, 0), g) + shift((-1, 0), g) + shift((0, 1), g) + shift((0, -1), g) - 4 * g)
        new = bind(x + 0.1 * lap)
        store(u, select(lt(new, imm(0, ARRAY, "Float")), imm(0, ARRAY, "Float"), new))
        store(e, reduce("Sum", new * new))
    om = lambda: makeOM("Diff", [], [u, e], [("k", k)], dim=2)
    setup = Setup(local_size=(300, 17), boundary=(CYCLIC, OPEN))
    fill = {"u": np.random.default_rng(5).random(mem_shape(setup, om)).astype(np.float32)}
    desc, so = build_emulated(setup, om(), tag="gen_diff")
    assert desc["kernels"][0]["stages"][0]["rings"] >= 1
    m, o = run_both(om, setup, ["k", "k"], "gen_diff", fill, rtol=1e-5)


def test_one_dimensional_open_chain_with_cast():
    t = Named("t", StaticValue(ARRAY, "Double"))
    c = Named("c", StaticValue(SCALAR, "Double"))

    def k():
        x = bind(load(t))
        store(t, 0.25 * shift((1,), x) + 0.5 * x + 0.25 * shift((-1,), x) + cast(loadIndex(0), "Double") * 1e-3)
        store(c, reduce("Max", max_(x, min_(x * 2, x + 1))))
    om = lambda: makeOM("Chain", [], [t, c], [("k", k)], dim=1)
    setup = Setup(local_size=(1500,), boundary=(OPEN,))
    fill = {"t": np.random.default_rng(6).random(mem_shape(setup, om))}
    run_both(om, setup, ["k", "k", "k"], "gen_chain", fill)


def test_scalar_only_kernel_and_reduce_of_a_load():
    a = Named("a", StaticValue(ARRAY, "Int"))
    n = Named("n", StaticValue(SCALAR, "Int"))
    tot = Named("tot", StaticValue(SCALAR, "Int"))

    def tick():
        store(n, load(n) * 2 + 1)

    def total():
        store(tot, reduce("Sum", load(a)) + load(n))
    om = lambda: makeOM("Tick", [], [a, n, tot], [("tick", tick), ("total", total)], dim=2)
    setup = Setup(local_size=(33, 5), boundary=(OPEN, OPEN))
    fill = {"a": np.random.default_rng(8).integers(0, 9, mem_shape(setup, om)).astype(np.int32)}
    m, o = run_both(om, setup, ["tick", "tick", "total"], "gen_tick", fill)
    assert int(m.scalar("n")) == 3 and int(m.scalar("tot")) == int(fill["a"].sum()) + 3


def test_initialcondition_example_operators():
    """examples/InitialCondition: cast, ^ by squaring, ** = exp(log x * y), atan — identical to the oracle when both
    sides use the same libm (emulated kernels); the CUDA math library is compared with a tolerance in the GPU tests."""
    from paraiso_b200.examples.initialcondition import initialcondition_om, initialcondition_setup
    setup = initialcondition_setup((60, 50))
    desc, so = build_emulated(setup, initialcondition_om(), tag="gen_heart")
    m = Machine(desc, so, device="cpu", _emulated=True)
    o = OracleMachine(setup, initialcondition_om())
    m.call("create"); o.call("create")
    a, b = m.get("table"), o.array("table")
    assert np.array_equal(a, b, equal_nan=True)
    assert np.isfinite(a).sum() > 0.9 * a.size and np.ptp(a[np.isfinite(a)]) > 2.0      # the heart: atan saturates both ways
