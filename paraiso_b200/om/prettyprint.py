"""Text dump of an annotated OM (the reference's interchange / debug format).

Mirrors Language/Paraiso/OM/PrettyPrint.hs:34-106 (`prettyPrintA1`): one line per node
`<idx> <label> [<- (ord)pred ...] [-> (ord)succ ...]` followed by its annotations (allocation, Valid
intervals, Depend.Direct / Indirect, Depend.Calc for Manifest nodes, write group).  With `legacy=True`
the realm names and the immediate format of the older generator revision are used, which makes the
output comparable byte for byte with the dumps checked in under examples-old/*-exampled/output/OM.txt.
"""
from __future__ import annotations

from typing import List

from .. import annotation as A
from .graph import ARRAY, SCALAR, OM, Graph, Node, imm_value

HS_TYPE = {"Int": "Int", "Integer": "Integer", "Float": "Float", "Double": "Double", "Bool": "Bool"}


def _realm(r: str, legacy: bool) -> str:
    if legacy:
        return {ARRAY: "Local", SCALAR: "Global"}[r]
    return r


def _dyn(dv, legacy) -> str:
    return f"DynValue {{realm = {_realm(dv.realm, legacy)}, typeRep = {HS_TYPE[dv.type]}}}"


def _vec(v) -> str:
    s = "Vec"
    for i, x in enumerate(v):
        s = f"{s} :~ {x}" if i == 0 else f"({s}) :~ {x}"
    return s


def _imm(inst, legacy) -> str:
    if legacy:
        return f"Imm <<{HS_TYPE[inst.imm_type]}>>"
    v = imm_value(inst.arg, inst.imm_type)
    if inst.imm_type == "Bool":
        return "Imm " + ("true" if v else "false")
    if inst.imm_type in ("Int", "Integer"):
        return f"Imm {int(v)}"
    return "Imm " + (repr(float(v)) + ("f" if inst.imm_type == "Float" else ""))


def _node(nd: Node, legacy: bool) -> str:
    if nd.is_value:
        return _dyn(nd.value, legacy)
    i = nd.inst
    if i.op in ("Load", "Store"):
        return f"{i.op} static[{i.arg}]"
    if i.op == "Reduce":
        return f"Reduce {i.arg}"
    if i.op == "Broadcast":
        return "Broadcast"
    if i.op in ("LoadIndex", "LoadSize"):
        return f"{i.op} (Axis {{axisIndex = {i.arg}}})"
    if i.op == "Shift":
        return f"Shift ({_vec(i.arg)})"
    if i.op == "Imm":
        return _imm(i, legacy)
    if i.op == "Arith":
        if i.arg == "Cast":
            return f"Arith (Cast {HS_TYPE[i.cast_to]})"
        return f"Arith {i.arg}"
    raise ValueError(i.op)


def _nb(x) -> str:
    if x == A.NEGA_INF:
        return "[-inf"
    if x == A.POSI_INF:
        return "+inf]"
    return f"[{x[1]}" if x[0] == 1 else f"{x[1]}]"


def _anot(anots: list, legacy: bool, alive: bool = True) -> List[str]:
    out: List[str] = []
    allocs = A.to_list(A.Allocation, anots)
    out += [a.kind for a in allocs]
    for v in A.to_list(A.Valid, anots):
        out.append(" ".join("[empty]" if iv.lower is None else f"{_nb(iv.lower)}..{_nb(iv.upper)}" for iv in v.intervals))
    out += ["Depend.Direct [" + ",".join(map(str, d.nodes)) + "]" for d in A.to_list(A.Direct, anots)]
    out += ["Depend.Indirect [" + ",".join(map(str, d.nodes)) + "]" for d in A.to_list(A.Indirect, anots)]
    if allocs == [A.Manifest]:
        out += ["Depend.Calc [" + ",".join(map(str, sorted(c.nodes))) + "]" for c in A.to_list(A.Calc, anots)]
    if alive:
        out += [f"Alive {a.alive}" for a in A.to_list(A.Alive, anots)]
    out += [f"KernelWriteGroup {{getKernelGroupID = {g.gid}}}" for g in A.to_list(A.KernelWriteGroup, anots)]
    out += [f"OMWriteGroup {{getOMGroupID = {g.gid}}}" for g in A.to_list(A.OMWriteGroup, anots)]
    return ["  " + l for l in out]


def _edges(symbol: str, xs) -> str:
    if not xs:
        return ""
    return " ".join([symbol] + [f"({o}){i}" for (o, i) in sorted(xs)])


def pretty_print_a1(om: OM, legacy: bool = False, alive: bool = True) -> str:
    lines = [f"OM name: {om.name}", "** Static Variables"]
    static = "".join(f'Named (Name "{sv.name}") ({_dyn(sv.namee, legacy)})\n' for sv in om.setup.static_values)
    lines.append(static)
    lines.append("** Kernels")
    kerns = []
    for k in om.kernels:
        g: Graph = k.dataflow
        kl = [f"*** Kernel name: {k.name}"]
        for idx, nd in enumerate(g.nodes):
            ins = [(o, p) for o, p in enumerate(nd.pre)]
            outs = [(g.nodes[s].pre.index(idx), s) for s in nd.suc]
            kl.append(" ".join([str(idx), _node(nd, legacy), _edges("<-", ins), _edges("->", outs)]))
            kl += _anot(nd.anot, legacy, alive)
        kerns.append("\n".join(kl) + "\n")
    lines.append("\n".join(kerns) + "\n")
    return "\n".join(lines) + "\n"
