"""Select sinking (paraiso_b200/generator/b200/selectsink.py): `select k (f a..) (f b..)` becomes `f (select k a b ..)`.
The rewrite must be exact (every OM instruction is pure), so the emulated kernels still equal the oracle bit for bit,
and it must actually fire on the pattern it was written for: the four-way select that ends Hydro's HLLC solver
(examples/Hydro/HydroMain.hs:252-258)."""
from collections import Counter

import numpy as np

from paraiso_b200.annotation import OPEN
from paraiso_b200.examples.hydro import hydro_om, hydro_setup
from paraiso_b200.generator.b200.schedule import fold_ops
from paraiso_b200.generator.b200.selectsink import sink_selects
from paraiso_b200.generator.native import Setup
from paraiso_b200.generator.plan import translate
from paraiso_b200.om.builder import StaticValue, bind, imm, load, lt, makeOM, select, shift, sqrt, store
from paraiso_b200.om.graph import ARRAY, Named
from tests.test_generic_programs import mem_shape, run_both


def _hist(ops):
    return Counter(o.inst.arg if o.kind == "Arith" else o.kind for o in ops.values() if o.realm == "Array")


def _proceed_dag(setup, om):
    plan = translate(setup, om)
    k = [k for k in plan.om.kernels if k.name == "proceed"][0]
    return fold_ops(k.dataflow, plan.om.dim)


def test_hydro_hllc_computes_one_star_state_per_wall():
    ops, stores = _proceed_dag(hydro_setup(), hydro_om("master"))
    new_ops, new_stores, stats = sink_selects(ops, stores)
    assert stats["accepted"] >= 4 and stats["rewritten"] >= 16          # 4 walls x 4 flux components
    assert stats["cost_after"] < 0.85 * stats["cost_before"]
    h0, h1 = _hist(ops), _hist(new_ops)
    assert h1["Div"] <= h0["Div"] - 24 and h1["Mul"] <= h0["Mul"] - 60 and h1["Add"] <= h0["Add"] - 40
    assert h1["Select"] <= h0["Select"] + 20                             # the selects move to the leaves; few are added
    assert len(new_stores) == len(stores)
    for v, o in new_ops.items():                                        # ascending ids are still a topological order
        assert all(a < v for a in o.args)
    # store targets keep their Valid annotation (they are never rewritten)
    assert [new_ops[v].valid for (_s, v) in new_stores] == [ops[v].valid for (_s, v) in stores]


def test_programs_without_mergeable_selects_are_untouched():
    from paraiso_b200.examples.life import life_om, life_setup
    ops, stores = _proceed_dag(life_setup("master"), life_om("master"))
    new_ops, new_stores, stats = sink_selects(ops, stores)
    assert stats["accepted"] == 0 and new_ops is ops and new_stores is stores


def test_knob_switches_the_pass_off():
    from paraiso_b200.generator.b200.schedule import schedule_kernel
    plan = translate(hydro_setup(), hydro_om("master"))
    k = [k for k in plan.om.kernels if k.name == "proceed"][0]
    on = schedule_kernel(plan.om, k, 11)
    off = schedule_kernel(plan.om, k, 11, sink_selects=False)
    assert off.sink_stats is None and on.sink_stats["accepted"] >= 4
    assert len(on.ops) < len(off.ops)


def _riemann_like_om():
    """A four-way select over mirrored formulas (plain / star on the left and right state), with commutative operands
    written in different orders on the two sides."""
    q = Named("q", StaticValue(ARRAY, "Double"))
    u = Named("u", StaticValue(ARRAY, "Double"))
    out = Named("out", StaticValue(ARRAY, "Double"))

    def k():
        qr, ur = bind(load(q)), bind(load(u))
        ql, ul = bind(shift((1, 0), qr)), bind(shift((1, 0), ur))
        sl = bind(ul - sqrt(ql))
        sr = bind(ur + sqrt(qr))
        sm = bind((qr * ur - ql * ul) / (ql + qr))
        plain_l = bind(ql * ul + ul * ul * ql)
        plain_r = bind(ur * qr + qr * ur * ur)          # same formula, commutative operands swapped
        star_l = bind(ql * (sl - ul) / (sl - sm) * (sm + ul / ql))
        star_r = bind(qr * (sr - ur) / (sr - sm) * (sm + ur / qr))
        zero = imm(0, ARRAY, "Double")
        f = bind(select(lt(zero, sl), plain_l, select(lt(zero, sm), star_l, select(lt(zero, sr), star_r, plain_r))))
        store(out, f - shift((-1, 0), f))
    return makeOM("Riemann", [], [q, u, out], [("k", k)], dim=2)


def test_four_way_select_is_sunk_and_stays_bit_identical():
    setup = Setup(local_size=(67, 11), boundary=(OPEN, OPEN))
    plan = translate(setup, _riemann_like_om())
    ops, stores = fold_ops(plan.om.kernels[0].dataflow, 2)
    new_ops, _st, stats = sink_selects(ops, stores)
    assert stats["accepted"] == 1
    h0, h1 = _hist(ops), _hist(new_ops)
    assert h1["Div"] == h0["Div"] - 2          # one star state (2 divisions) instead of two
    rng = np.random.default_rng(11)
    shp = mem_shape(setup, _riemann_like_om)
    fill = {"q": rng.uniform(0.5, 2.0, shp), "u": rng.uniform(-2.0, 2.0, shp)}   # all four branches occur
    run_both(_riemann_like_om, setup, ["k"], "gen_riemann", fill)


def test_fast_algebra_removes_zero_and_unit_terms_only_in_fast_builds():
    """selectsink.simplify_fast: x*0, x+0, x*1, (a*b)/b -> a.  Ids are kept, store targets are never replaced, and the
    exact build does not run the pass (schedule_kernel(fast_algebra=False) is the default)."""
    from paraiso_b200.generator.b200.schedule import schedule_kernel
    from paraiso_b200.generator.b200.selectsink import simplify_fast
    ops, stores = _proceed_dag(hydro_setup(fast=True), hydro_om("master"))
    new_ops, new_stores, removed = simplify_fast(ops, stores)
    assert removed >= 60 and set(new_ops) <= set(ops)
    assert [v for (_s, v) in new_stores] == [v for (_s, v) in stores]
    h0, h1 = _hist(ops), _hist(new_ops)
    assert h1["Mul"] <= h0["Mul"] - 24 and h1["Add"] <= h0["Add"] - 24 and h1["Div"] <= h0["Div"] - 12
    for v, o in new_ops.items():
        assert all(a in new_ops and a < v for a in o.args)
    plan = translate(hydro_setup(), hydro_om("master"))
    k = [k for k in plan.om.kernels if k.name == "proceed"][0]
    exact = schedule_kernel(plan.om, k, 11)
    fast = schedule_kernel(plan.om, k, 11, fast_algebra=True)
    assert len(fast.ops) < len(exact.ops)
