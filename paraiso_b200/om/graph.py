"""Orthotope Machine IR: a bipartite dataflow graph of value and instruction nodes.

Mirrors Language/Paraiso/OM/Graph.hs:82-123 (Node, Edge, Inst), OM/Arithmetic.hs:20-110
(Operator + arity), OM/Reduce.hs:9, OM/Realm.hs:37-40, OM/DynValue.hs:17 and OM.hs:16-41.

Node numbering follows the reference's FGL usage: a node's id is the number of
nodes present when it was added (OM/Builder/Internal.hs:104-119), edges into a
node are ordered (EOrd i) by argument position.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from fractions import Fraction
from typing import Any, List, Optional, Tuple

# ---- realms and content types -------------------------------------------------------------
SCALAR = "Scalar"
ARRAY = "Array"

# Haskell type name -> C++ type (ClarisTrans.hs:189-199)
CPP_TYPE = {
    "()": "void",
    "Bool": "bool",
    "Int": "int",
    "Integer": "long long int",
    "Float": "float",
    "Double": "double",
}
TYPE_BYTES = {"Bool": 1, "Int": 4, "Integer": 8, "Float": 4, "Double": 8}


@dataclass(frozen=True)
class DynValue:
    """OM/DynValue.hs:17 — realm + content type of a value node."""
    realm: str
    type: str


# ---- arithmetic operators (OM/Arithmetic.hs:20-66) ------------------------------------------
ARITY = {
    "Identity": (1, 1), "Add": (2, 1), "Sub": (2, 1), "Neg": (1, 1), "Mul": (2, 1),
    "Div": (2, 1), "Mod": (2, 1), "DivMod": (2, 2), "Inv": (1, 1), "Not": (1, 1),
    "And": (2, 1), "Or": (2, 1), "EQ": (2, 1), "NE": (2, 1), "LT": (2, 1), "LE": (2, 1),
    "GT": (2, 1), "GE": (2, 1), "Max": (2, 1), "Min": (2, 1), "Abs": (1, 1),
    "Signum": (1, 1), "Select": (3, 1), "Ipow": (2, 1), "Pow": (2, 1), "Madd": (3, 1),
    "Msub": (3, 1), "Nmadd": (3, 1), "Nmsub": (3, 1), "Sqrt": (1, 1), "Exp": (1, 1),
    "Log": (1, 1), "Sin": (1, 1), "Cos": (1, 1), "Tan": (1, 1), "Asin": (1, 1),
    "Acos": (1, 1), "Atan": (1, 1), "Atan2": (2, 1), "Sincos": (1, 2), "Cast": (1, 1),
}

REDUCE_OPS = ("Max", "Min", "Sum")  # OM/Reduce.hs:9


@dataclass(frozen=True)
class Inst:
    """OM/Graph.hs:113-123.  `op` is one of Load Store Reduce Broadcast LoadIndex LoadSize
    Shift Imm Arith; `arg` carries the payload (static index, reduce operator, axis,
    shift vector, immediate, arithmetic operator); `cast_to` is the target type of Arith Cast."""
    op: str
    arg: Any = None
    cast_to: Optional[str] = None
    imm_type: Optional[str] = None  # content type of an Imm

    def arity(self) -> Tuple[int, int]:
        return {
            "Load": (0, 1), "Store": (1, 0), "Reduce": (1, 1), "Broadcast": (1, 1),
            "LoadIndex": (0, 1), "LoadSize": (0, 1), "Shift": (1, 1), "Imm": (0, 1),
        }.get(self.op) or ARITY[self.arg]


@dataclass
class Node:
    """NValue DynValue anot | NInst Inst anot (OM/Graph.hs:82-89)."""
    value: Optional[DynValue] = None
    inst: Optional[Inst] = None
    anot: list = field(default_factory=list)   # Annotation = [Dynamic] (Annotation.hs:17)
    pre: List[int] = field(default_factory=list)  # ordered by EOrd
    suc: List[int] = field(default_factory=list)

    @property
    def is_value(self) -> bool:
        return self.value is not None


class Graph:
    """FGL.Gr (Node v g a) Edge restricted to what the reference uses."""

    def __init__(self):
        self.nodes: List[Node] = []

    def no_nodes(self) -> int:
        return len(self.nodes)

    def add(self, froms: List[int], node: Node) -> int:
        n = len(self.nodes)
        node.pre = list(froms)
        node.suc = []
        self.nodes.append(node)
        for f in froms:
            self.nodes[f].suc.append(n)
        return n

    def lab(self, i: int) -> Node:
        return self.nodes[i]

    def pre_inst(self, i: int) -> Tuple[int, Inst]:
        """The instruction node defining value node i."""
        for p in self.nodes[i].pre:
            if not self.nodes[p].is_value:
                return p, self.nodes[p].inst
        raise ValueError(f"value node {i} has no defining instruction")

    def operands(self, i: int) -> List[int]:
        """Ordered value-node operands of the instruction that defines value node i."""
        p, _ = self.pre_inst(i)
        return list(self.nodes[p].pre)


@dataclass
class Named:
    name: str
    namee: Any


@dataclass
class Setup:
    """OM/Graph.hs:30-36."""
    static_values: List[Named]
    global_annotation: list


@dataclass
class Kernel:
    name: str
    dataflow: Graph


@dataclass
class OM:
    """OM.hs:16-22."""
    name: str
    setup: Setup
    kernels: List[Kernel]
    dim: int = 2


def imm_value(content, ctype: str):
    """Concrete value of an immediate in its content type.  Fractions are rounded once,
    correctly, as Haskell's fromRational does (Builder/Internal.hs:366, 374)."""
    import numpy as np
    if ctype == "Bool":
        return bool(content)
    if ctype in ("Int", "Integer"):
        return int(content)
    if ctype == "Double":
        return float(Fraction(content)) if not isinstance(content, float) else content
    if ctype == "Float":
        if isinstance(content, float):
            return np.float32(content)
        fr = Fraction(content)
        c = np.float32(float(fr))
        best, bestd = c, abs(Fraction(float(c)) - fr)
        for cand in (np.nextafter(c, np.float32(np.inf)), np.nextafter(c, np.float32(-np.inf))):
            d = abs(Fraction(float(cand)) - fr)
            if d < bestd:
                best, bestd = cand, d
        return best
    raise ValueError(ctype)
