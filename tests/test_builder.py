"""Builder / annotation semantics (mirrors Test/Paraiso/Annotation.hs:22-35 and OM/Builder/Internal.hs)."""
import pytest

from paraiso_b200 import annotation as A
from paraiso_b200.om.builder import StaticValue, bind, build_kernel, imm, load, makeOM, reduce, shift, store
from paraiso_b200.om.graph import ARRAY, SCALAR, Named, Setup, DynValue


def test_annotation_set_is_unique_per_type():
    a = A.add(A.Manifest, A.add(A.Delayed, A.add(A.Alive(True), [])))
    s = A.set_(A.Existing, a)
    assert A.to_list(A.Allocation, s) == [A.Existing]
    assert A.to_list(A.Alive, s) == [A.Alive(True)]          # no cross-type contamination
    assert A.to_maybe(A.Allocation, A.add(A.Manifest, s)) == A.Manifest   # add puts in front, toMaybe takes the first
    assert A.weak_set(A.Delayed, s) == s


def test_unbound_builder_is_rerun_and_bind_shares():
    x = Named("x", StaticValue(ARRAY, "Int"))
    setup = Setup([Named("x", DynValue(ARRAY, "Int"))], [])

    def unbound():
        v = load(x)            # not bound: every use re-runs the Load
        store(x, v + v)

    def bound():
        v = bind(load(x))
        store(x, v + v)
    assert len(build_kernel(setup, "k", unbound).dataflow.nodes) == 7   # 2x(Load, value) + Add, value + Store
    assert len(build_kernel(setup, "k", bound).dataflow.nodes) == 5


def test_shift_sign_and_imm_order():
    x = Named("x", StaticValue(ARRAY, "Int"))
    setup = Setup([Named("x", DynValue(ARRAY, "Int"))], [])

    def k():
        store(x, 10 * shift((1,), load(x)))
    g = build_kernel(setup, "k", k).dataflow
    ops = [(n.inst.op, n.inst.arg) for n in g.nodes if n.inst is not None]
    # mkOp2 runs both operand builders first, then materialises immediates (Internal.hs:302-317)
    assert ops == [("Load", 0), ("Shift", (1,)), ("Imm", 10), ("Arith", "Mul"), ("Store", 0)]


def test_type_mismatch_and_unknown_static():
    x = Named("x", StaticValue(ARRAY, "Int"))
    setup = Setup([Named("x", DynValue(ARRAY, "Int"))], [])
    with pytest.raises(TypeError):
        build_kernel(setup, "k", lambda: store(x, imm(1.5, ARRAY, "Double")))
    with pytest.raises(KeyError):
        build_kernel(setup, "k", lambda: store(Named("y", StaticValue(ARRAY, "Int")), imm(1, ARRAY, "Int")))
