"""Synthetic OM programs for schedule corners the two headline programs do not reach (the same shapes as
tests/test_generic_programs.py): name -> (make_om, setup, kernel call sequence)."""
from paraiso_b200.annotation import CYCLIC, OPEN
from paraiso_b200.generator.native import Setup
from paraiso_b200.om.builder import (StaticValue, bind, broadcast, cast, imm, load, loadIndex, lt, makeOM, max_, min_, reduce,
                                     select, shift, sqrt, store)
from paraiso_b200.om.graph import ARRAY, SCALAR, Named


def multi_reduce():
    a = Named("a", StaticValue(ARRAY, "Int"))
    s, mx, mn = (Named(n, StaticValue(SCALAR, "Int")) for n in ("s", "mx", "mn"))

    def k():
        x = bind(load(a))
        y = bind(x * 3 - shift((1, 0), x))
        tot = bind(reduce("Sum", y))
        store(s, tot)
        store(mx, reduce("Max", y))
        store(mn, reduce("Min", y + 7))
        store(a, y - broadcast(tot) / 1000)        # second stage: depends on the reduce
    return makeOM("Multi", [], [a, s, mx, mn], [("k", k)], dim=2)


def wide_stencil():
    a = Named("a", StaticValue(ARRAY, "Int"))
    b = Named("b", StaticValue(ARRAY, "Int"))

    def k():
        x = bind(load(a))
        store(b, shift((5, 0), x) + 2 * shift((-6, 2), x) - shift((0, -3), x) + loadIndex(0) * 100 + loadIndex(1))
        store(a, x + 1)
    return makeOM("Wide", [], [a, b], [("k", k)], dim=2)


def ring_float():
    u = Named("u", StaticValue(ARRAY, "Float"))
    e = Named("e", StaticValue(SCALAR, "Float"))

    def k():
        x = bind(load(u))
        g = bind(sqrt(x * x + 1.5) / (x + 3.0))                 # expensive: becomes a shared-memory ring
        lap = bind(shift((1, 0), g) + shift((-1, 0), g) + shift((0, 1), g) + shift((0, -1), g) - 4 * g)
        new = bind(x + 0.1 * lap)
        store(u, select(lt(new, imm(0, ARRAY, "Float")), imm(0, ARRAY, "Float"), new))
        store(e, reduce("Sum", new * new))
    return makeOM("Diff", [], [u, e], [("k", k)], dim=2)


def chain_1d():
    t = Named("t", StaticValue(ARRAY, "Double"))
    c = Named("c", StaticValue(SCALAR, "Double"))

    def k():
        x = bind(load(t))
        store(t, 0.25 * shift((1,), x) + 0.5 * x + 0.25 * shift((-1,), x) + cast(loadIndex(0), "Double") * 1e-3)
        store(c, reduce("Max", max_(x, min_(x * 2, x + 1))))
    return makeOM("Chain", [], [t, c], [("k", k)], dim=1)


PROGRAMS = {
    "multi_reduce": (multi_reduce, Setup(local_size=(70, 9), boundary=(CYCLIC, CYCLIC)), ["k", "k"]),
    "wide_CO": (wide_stencil, Setup(local_size=(61, 23), boundary=(CYCLIC, OPEN)), ["k", "k", "k"]),
    "wide_OC": (wide_stencil, Setup(local_size=(61, 23), boundary=(OPEN, CYCLIC)), ["k", "k", "k"]),
    "wide_OO": (wide_stencil, Setup(local_size=(61, 23), boundary=(OPEN, OPEN)), ["k", "k", "k"]),
    "ring_float": (ring_float, Setup(local_size=(300, 17), boundary=(CYCLIC, OPEN)), ["k", "k"]),
    "chain_1d": (chain_1d, Setup(local_size=(1500,), boundary=(OPEN,)), ["k", "k", "k"]),
}
