"""The Builder EDSL: the eight reserved words plus overloaded arithmetic.

Mirrors Language/Paraiso/OM/Builder/Internal.hs and OM/Builder/Boolean.hs.  A Haskell
`Builder v g a (Value r c)` is a *computation* that appends nodes to the graph when it is
run, and running it twice appends the nodes twice; `bind` (Internal.hs:163-164) runs it once
and wraps the result.  `B` below keeps exactly that semantics (a thunk plus its static realm
and content type), so a program transcribed statement by statement produces the same node
numbering as the reference (checked against examples-old/*-exampled/dist/*.cpp in tests).

Reserved words (Paraiso.cabal:40-45): load store imm loadIndex loadSize shift reduce broadcast.
"""
from __future__ import annotations

import math
from fractions import Fraction
from typing import Callable, List, Sequence

from .graph import ARRAY, SCALAR, DynValue, Graph, Inst, Kernel, Named, Node, OM, Setup

# ---------------------------------------------------------------------------------------------
# builder state (Internal.hs:64-83)
# ---------------------------------------------------------------------------------------------


class BuilderState:
    def __init__(self, setup: Setup):
        self.setup = setup
        self.current_annotation = list(setup.global_annotation)
        self.target = Graph()


_state: List[BuilderState] = []


def _st() -> BuilderState:
    if not _state:
        raise RuntimeError("Builder action used outside buildKernel")
    return _state[-1]


class Value:
    """OM/Value.hs:19-22: FromNode realm content node | FromImm realm content."""
    __slots__ = ("realm", "ctype", "node", "content")

    def __init__(self, realm, ctype, node=None, content=None):
        self.realm, self.ctype, self.node, self.content = realm, ctype, node, content


class StaticValue:
    """OM/Value.hs:24 — a typed handle on a static variable."""

    def __init__(self, realm: str, ctype: str):
        self.realm, self.ctype = realm, ctype

    def dyn(self) -> DynValue:
        return DynValue(self.realm, self.ctype)


def add_node(froms: List[int], node: Node) -> int:  # Internal.hs:112-119
    return _st().target.add(froms, node)


def _add_inst(froms, inst: Inst) -> int:  # addNodeE with NInst
    return add_node(froms, Node(inst=inst, anot=list(_st().current_annotation)))


def _add_value(froms, dyn: DynValue) -> int:  # addNodeE with NValue
    return add_node(froms, Node(value=dyn, anot=list(_st().current_annotation)))


def value_to_node(val: Value) -> int:  # Internal.hs:135-146
    if val.node is not None:
        return val.node
    n0 = _add_inst([], Inst("Imm", val.content, imm_type=val.ctype))
    return _add_value([n0], DynValue(val.realm, val.ctype))


def lookup_static(name: str, dyn: DynValue) -> int:  # Internal.hs:150-166
    vs = _st().setup.static_values
    matches = [(i, v) for i, v in enumerate(vs) if v.name == name]
    if len(matches) != 1:
        raise KeyError(f"{len(matches)} match found for '{name}'")
    i, v = matches[0]
    if v.namee != dyn:
        raise TypeError(f"type mismatch; expected: {v.namee}; actual: {name}::{dyn}")
    return i


# ---------------------------------------------------------------------------------------------
# B: the builder computation with operator overloading
# ---------------------------------------------------------------------------------------------

_FLOATING = ("Float", "Double")


def _rational(x):
    if isinstance(x, bool) or isinstance(x, int) or isinstance(x, Fraction):
        return x
    if isinstance(x, float):
        # a literal written in program text: keep the decimal exactly as written
        # (Haskell: fromRational of the literal's exact value)
        return Fraction(repr(x))
    raise TypeError(f"cannot make an immediate from {x!r}")


class B:
    """Builder v g a (Value realm ctype)."""
    __slots__ = ("_run", "realm", "ctype")

    def __init__(self, run: Callable[[], Value], realm: str, ctype: str):
        self._run, self.realm, self.ctype = run, realm, ctype

    def run(self) -> Value:
        return self._run()

    # -- coercion of Python literals (fromInteger / fromRational', Internal.hs:325,366)
    def _lift(self, x) -> "B":
        if isinstance(x, B):
            return x
        return imm(x, self.realm, self.ctype)

    # Additive / Ring / Field (Internal.hs:320-368)
    def __add__(self, o): return mkOp2("Add", self, self._lift(o))
    def __radd__(self, o): return mkOp2("Add", self._lift(o), self)
    def __sub__(self, o): return mkOp2("Sub", self, self._lift(o))
    def __rsub__(self, o): return mkOp2("Sub", self._lift(o), self)
    def __mul__(self, o): return mkOp2("Mul", self, self._lift(o))
    def __rmul__(self, o): return mkOp2("Mul", self._lift(o), self)
    def __truediv__(self, o): return mkOp2("Div", self, self._lift(o))
    def __rtruediv__(self, o): return mkOp2("Div", self._lift(o), self)
    def __floordiv__(self, o): return mkOp2("Div", self, self._lift(o))  # IntegralDomain div
    def __mod__(self, o): return mkOp2("Mod", self, self._lift(o))
    def __neg__(self): return mkOp1("Neg", self)

    # Boolean (Internal.hs:377-382); Python's & | ~ stand in for && || not
    def __and__(self, o): return mkOp2("And", self, self._lift(o))
    def __or__(self, o): return mkOp2("Or", self, self._lift(o))
    def __invert__(self): return mkOp1("Not", self)

    def __pow__(self, n: int):  # Ring (^) by repeated squaring, Internal.hs:335-347
        if not isinstance(n, int) or n < 0:
            raise TypeError("only non-negative integer powers; use pow_ for ^/")
        if n == 0:
            return imm(1, self.realm, self.ctype)
        if n == 1:
            return self
        a = self

        def run():
            ba = bind(a)

            def f(x: B, n2: int) -> B:
                if n2 == 1:
                    return x
                n3 = n2 // 2

                def run2():
                    bx = bind(f(x, n3))
                    sq = bx * bx
                    return (x * sq if n2 - 2 * n3 > 0 else sq).run()
                return B(run2, x.realm, x.ctype)
            return f(ba, n).run()
        return B(run, self.realm, self.ctype)

    def __bool__(self):
        raise TypeError("a Builder value has no truth value; use select()")


def ret(v: Value) -> B:
    """`return v` in the Builder monad."""
    return B(lambda: v, v.realm, v.ctype)


def bind(b: B) -> B:
    """bind = fmap return (Internal.hs:163-164): run now, reuse the resulting value."""
    return ret(b.run())


def imm(c, realm: str, ctype: str) -> B:  # Internal.hs:274-277
    c = _rational(c)
    return B(lambda: Value(realm, ctype, content=c), realm, ctype)


def imm_exact(c: float, realm: str, ctype: str) -> B:
    """An immediate given as a machine float (e.g. pi), not as program text."""
    return B(lambda: Value(realm, ctype, content=float(c)), realm, ctype)


def pi(realm: str, ctype: str) -> B:  # Transcendental pi = imm pi (Internal.hs:408)
    import numpy as np
    return imm_exact(float(np.float32(math.pi)) if ctype == "Float" else math.pi, realm, ctype)


def load(named: Named) -> B:  # Internal.hs:169-179
    sv: StaticValue = named.namee
    dyn = sv.dyn()

    def run():
        idx = lookup_static(named.name, dyn)
        n0 = _add_inst([], Inst("Load", idx))
        n1 = _add_value([n0], dyn)
        return Value(sv.realm, sv.ctype, node=n1)
    return B(run, sv.realm, sv.ctype)


def store(named: Named, b) -> None:  # Internal.hs:182-194 (an action: runs immediately)
    sv: StaticValue = named.namee
    if not isinstance(b, B):
        b = imm(b, sv.realm, sv.ctype)
    val = b.run()
    idx = lookup_static(named.name, DynValue(val.realm, val.ctype))
    n0 = value_to_node(val)
    _add_inst([n0], Inst("Store", idx))


def reduce(op: str, b: B) -> B:  # Internal.hs:200-212
    def run():
        val = b.run()
        n1 = value_to_node(val)
        n2 = _add_inst([n1], Inst("Reduce", op))
        n3 = _add_value([n2], DynValue(SCALAR, val.ctype))
        return Value(SCALAR, val.ctype, node=n3)
    return B(run, SCALAR, b.ctype)


def broadcast(b: B) -> B:  # Internal.hs:216-227
    def run():
        val = b.run()
        n1 = value_to_node(val)
        n2 = _add_inst([n1], Inst("Broadcast"))
        n3 = _add_value([n2], DynValue(ARRAY, val.ctype))
        return Value(ARRAY, val.ctype, node=n3)
    return B(run, ARRAY, b.ctype)


def loadIndex(axis: int, gauge: str = "Int") -> B:  # Internal.hs:231-241
    def run():
        n0 = _add_inst([], Inst("LoadIndex", axis))
        n1 = _add_value([n0], DynValue(ARRAY, gauge))
        return Value(ARRAY, gauge, node=n1)
    return B(run, ARRAY, gauge)


def loadSize(axis: int, gauge: str = "Int", realm: str = SCALAR) -> B:  # Internal.hs:244-254
    # `realm=ARRAY` reproduces the older API (`loadSize TLocal`) used by examples-old/*-exampled
    def run():
        n0 = _add_inst([], Inst("LoadSize", axis))
        n1 = _add_value([n0], DynValue(realm, gauge))
        return Value(realm, gauge, node=n1)
    return B(run, realm, gauge)


def shift(vec: Sequence[int], b: B) -> B:  # Internal.hs:257-269
    vec = tuple(int(x) for x in vec)

    def run():
        val = b.run()
        n1 = value_to_node(val)
        n2 = _add_inst([n1], Inst("Shift", vec))
        n3 = _add_value([n2], DynValue(val.realm, val.ctype))
        return Value(ARRAY, val.ctype, node=n3)
    return B(run, ARRAY, b.ctype)


def mkOp1(op: str, b1: B) -> B:  # Internal.hs:287-299
    def run():
        v1 = b1.run()
        n1 = value_to_node(v1)
        n0 = _add_inst([n1], Inst("Arith", op))
        n01 = _add_value([n0], DynValue(v1.realm, v1.ctype))
        return Value(v1.realm, v1.ctype, node=n01)
    return B(run, b1.realm, b1.ctype)


def mkOp2(op: str, b1: B, b2: B, out_type: str = None) -> B:  # Internal.hs:302-317, Boolean.hs:21-35
    def run():
        v1 = b1.run()
        v2 = b2.run()
        n1 = value_to_node(v1)
        n2 = value_to_node(v2)
        n0 = _add_inst([n1, n2], Inst("Arith", op))
        t = out_type or v1.ctype
        n01 = _add_value([n0], DynValue(v1.realm, t))
        return Value(v1.realm, t, node=n01)
    return B(run, b1.realm, out_type or b1.ctype)


def _lift2(a, b):
    if isinstance(a, B):
        return a, a._lift(b)
    if isinstance(b, B):
        return b._lift(a), b
    raise TypeError("at least one operand must be a Builder value")


# comparison (Boolean.hs:41-58)
def eq(a, b): a, b = _lift2(a, b); return mkOp2("EQ", a, b, "Bool")
def ne(a, b): a, b = _lift2(a, b); return mkOp2("NE", a, b, "Bool")
def lt(a, b): a, b = _lift2(a, b); return mkOp2("LT", a, b, "Bool")
def le(a, b): a, b = _lift2(a, b); return mkOp2("LE", a, b, "Bool")
def gt(a, b): a, b = _lift2(a, b); return mkOp2("GT", a, b, "Bool")
def ge(a, b): a, b = _lift2(a, b); return mkOp2("GE", a, b, "Bool")


def select(bb: B, b1, b2) -> B:  # Boolean.hs:61-78
    if not isinstance(b1, B) and not isinstance(b2, B):
        raise TypeError("select needs at least one typed branch")
    b1, b2 = _lift2(b1, b2)

    def run():
        vb = bb.run()
        v1 = b1.run()
        v2 = b2.run()
        nb = value_to_node(vb)
        n1 = value_to_node(v1)
        n2 = value_to_node(v2)
        n0 = _add_inst([nb, n1, n2], Inst("Arith", "Select"))
        n01 = _add_value([n0], DynValue(v1.realm, v1.ctype))
        return Value(v1.realm, v1.ctype, node=n01)
    return B(run, b1.realm, b1.ctype)


# Lattice (up/dn), Absolute, Algebraic, Transcendental (Internal.hs:384-417)
def max_(a, b): a, b = _lift2(a, b); return mkOp2("Max", a, b)
def min_(a, b): a, b = _lift2(a, b); return mkOp2("Min", a, b)
def abs_(a: B): return mkOp1("Abs", a)
def signum(a: B): return mkOp1("Signum", a)
def sqrt(a: B): return mkOp1("Sqrt", a)
def exp(a: B): return mkOp1("Exp", a)
def log(a: B): return mkOp1("Log", a)
def sin(a: B): return mkOp1("Sin", a)
def cos(a: B): return mkOp1("Cos", a)
def tan(a: B): return mkOp1("Tan", a)
def asin(a: B): return mkOp1("Asin", a)
def acos(a: B): return mkOp1("Acos", a)
def atan(a: B): return mkOp1("Atan", a)
def recip(a: B): return mkOp1("Inv", a)
def pow_(a: B, y): return mkOp2("Pow", a, a._lift(y))  # x ^/ y


def cast(b1: B, ctype2: str) -> B:  # Internal.hs:422-432
    def run():
        v1 = b1.run()
        n1 = value_to_node(v1)
        n0 = _add_inst([n1], Inst("Arith", "Cast", cast_to=ctype2))
        n01 = _add_value([n0], DynValue(v1.realm, ctype2))
        return Value(v1.realm, ctype2, node=n01)
    return B(run, b1.realm, ctype2)


def annotate(f: Callable[[list], list], b1: B) -> B:  # Internal.hs:465-479
    def run():
        v1 = b1.run()
        n1 = value_to_node(v1)
        nd = _st().target.nodes[n1]
        nd.anot = f(nd.anot)
        return Value(v1.realm, v1.ctype, node=n1)
    return B(run, b1.realm, b1.ctype)


def with_annotation(f, thunk: Callable[[], object]):  # Internal.hs:451-461
    st = _st()
    a0 = st.current_annotation
    st.current_annotation = f(list(a0))
    try:
        return thunk()
    finally:
        st.current_annotation = a0


# ---------------------------------------------------------------------------------------------
# numeric-prelude / typelevel-tensor helpers used by the example programs
# ---------------------------------------------------------------------------------------------

def zero(realm, ctype) -> B:
    return imm(0, realm, ctype)


def sum_(xs: Sequence[B]) -> B:
    """NumericPrelude.sum = foldl (+) zero — visible as `0 + a + b` in generated code
    (examples-old/Hydro-exampled/dist/Hydro.cpp:204-205)."""
    acc = zero(xs[0].realm, xs[0].ctype)
    for x in xs:
        acc = acc + x
    return acc


def foldl1(f, xs):
    acc = xs[0]
    for x in xs[1:]:
        acc = f(acc, x)
    return acc


def contract(dim: int, f: Callable[[int], B]) -> B:
    """Data.Tensor.TypeLevel.contract: sum over the axes, starting from zero."""
    return sum_([f(i) for i in range(dim)])


def unit_vector(dim: int, axis: int):
    return tuple(1 if i == axis else 0 for i in range(dim))


# ---------------------------------------------------------------------------------------------
# buildKernel / makeOM (Internal.hs:53-61, OM.hs:28-41)
# ---------------------------------------------------------------------------------------------

def build_kernel(setup: Setup, name: str, builder: Callable[[], None]) -> Kernel:
    _state.append(BuilderState(setup))
    try:
        builder()
        g = _state[-1].target
    finally:
        _state.pop()
    return Kernel(name, g)


def makeOM(name: str, anot: list, vars_: List[Named], kernels, dim: int = 2) -> OM:
    """kernels: list of (kernel name, zero-argument Python function running Builder actions)."""
    setup = Setup(static_values=[Named(v.name, v.namee.dyn() if isinstance(v.namee, StaticValue) else v.namee)
                                 for v in vars_], global_annotation=list(anot))
    ks = [build_kernel(setup, n, b) for (n, b) in kernels]
    return OM(name=name, setup=setup, kernels=ks, dim=dim)
