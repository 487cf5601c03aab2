"""OM interchange: read an annotated-graph dump back into an OM.

The text format is the one `OM/PrettyPrint.hs:34-106` writes (`prettyPrintA1`, e.g. examples/Life/Generator.hs:30
writes output/OM.txt) and `om/prettyprint.py` mirrors.  With it a real Haskell Paraiso can drive this backend
without GHC being present here: dump the (optimised or not) OM upstream, `parse_om` it, `generate` from it.
Dumps of the current printer carry immediate values (`Imm 3`, `Imm 1.6666666666666667`, `Imm 0.47f`); the older
revision printed `Imm <<Int>>`, which cannot be rebuilt and is rejected.
"""
from __future__ import annotations

import re
from typing import List

from .. import annotation as A
from .graph import ARRAY, SCALAR, DynValue, Graph, Inst, Kernel, Named, Node, OM, Setup

_REALM = {"Array": ARRAY, "Scalar": SCALAR, "Local": ARRAY, "Global": SCALAR}
_DYN = re.compile(r"DynValue \{realm = (\w+), typeRep = (\w+)\}")
_EDGE = re.compile(r"\((\d+)\)(\d+)")


def _parse_label(label: str):
    m = _DYN.fullmatch(label)
    if m:
        return Node(value=DynValue(_REALM[m.group(1)], m.group(2)))
    if label.startswith("Load static["):
        return Node(inst=Inst("Load", int(label[12:-1])))
    if label.startswith("Store static["):
        return Node(inst=Inst("Store", int(label[13:-1])))
    if label.startswith("Reduce "):
        return Node(inst=Inst("Reduce", label.split()[1]))
    if label == "Broadcast":
        return Node(inst=Inst("Broadcast"))
    m = re.fullmatch(r"(LoadIndex|LoadSize) \(Axis \{axisIndex = (\d+)\}\)", label)
    if m:
        return Node(inst=Inst(m.group(1), int(m.group(2))))
    if label.startswith("Shift "):
        return Node(inst=Inst("Shift", tuple(int(x) for x in re.findall(r":~ (-?\d+)", label))))
    if label.startswith("Imm "):
        txt = label[4:]
        if txt.startswith("<<"):
            raise ValueError("legacy dump without immediate values cannot be imported: " + label)
        return Node(inst=Inst("Imm", txt))          # typed once the output value node is known
    m = re.fullmatch(r"Arith \(Cast (\w+)\)", label)
    if m:
        return Node(inst=Inst("Arith", "Cast", cast_to=m.group(1)))
    if label.startswith("Arith "):
        return Node(inst=Inst("Arith", label.split()[1]))
    raise ValueError(f"cannot parse node label: {label!r}")


def _imm_content(txt: str, ctype: str):
    if ctype == "Bool":
        return txt.strip().lower() == "true"
    if ctype in ("Int", "Integer"):
        return int(txt)
    return float(txt.rstrip("f"))


def parse_om(text: str, dim: int = None) -> OM:
    lines = text.split("\n")
    name = None
    statics: List[Named] = []
    kernels: List[Kernel] = []
    cur_nodes = None     # list of (label node, ordered preds, annotation lines)
    kname = None

    def finish():
        if cur_nodes is None:
            return
        g = Graph()
        for (nd, ins, anots) in cur_nodes:
            for a in anots:
                if a in ("Manifest",):      # user `Anot.add Alloc.Manifest <?>` marks survive re-analysis (weakSet)
                    nd.anot = A.add(A.Manifest, nd.anot)
            g.add([p for (_o, p) in sorted(ins)], nd)
        for i, nd in enumerate(g.nodes):     # type the immediates from the value node they define
            if nd.inst is not None and nd.inst.op == "Imm":
                out = g.nodes[nd.suc[0]].value
                g.nodes[i].inst = Inst("Imm", _imm_content(nd.inst.arg, out.type), imm_type=out.type)
        # Manifest marks that the analysis would set anyway are not user annotations
        for nd in g.nodes:
            sucs = [g.nodes[s] for s in nd.suc]
            pre0 = g.nodes[nd.pre[0]] if nd.pre else None
            auto = any(s.inst is not None and s.inst.op in ("Store", "Reduce", "Broadcast") for s in sucs) or \
                (pre0 is not None and pre0.inst is not None and pre0.inst.op in ("Reduce", "Broadcast"))
            if auto:
                nd.anot = [a for a in nd.anot if a != A.Manifest]
        kernels.append(Kernel(kname, g))

    for ln in lines:
        if ln.startswith("OM name: "):
            name = ln[9:].strip()
        elif ln.startswith("Named (Name "):
            m = re.fullmatch(r'Named \(Name "(.*)"\) \((DynValue \{.*\})\)', ln.strip())
            d = _DYN.fullmatch(m.group(2))
            statics.append(Named(m.group(1), DynValue(_REALM[d.group(1)], d.group(2))))
        elif ln.startswith("*** Kernel name: "):
            finish()
            kname = ln[17:].strip()
            cur_nodes = []
        elif ln.startswith("  ") and cur_nodes:
            cur_nodes[-1][2].append(ln.strip())
        elif ln and ln[0].isdigit() and cur_nodes is not None:
            idx_s, rest = ln.split(" ", 1)
            left, _, _outs = rest.partition(" -> ") if " -> " in rest else (rest, "", "")
            label, _, ins = left.partition(" <- ") if " <- " in left else (left, "", "")
            ins_l = [(int(o), int(i)) for (o, i) in _EDGE.findall(ins)]
            assert int(idx_s) == len(cur_nodes), "node ids must be dense and ordered"
            cur_nodes.append((_parse_label(label.strip()), ins_l, []))
    finish()
    if dim is None:
        dim = 2
        for k in kernels:
            for nd in k.dataflow.nodes:
                if nd.inst is not None and nd.inst.op == "Shift":
                    dim = len(nd.inst.arg)
    return OM(name=name, setup=Setup(static_values=statics, global_annotation=[]), kernels=kernels, dim=dim)
