"""OM interchange: read an annotated-graph dump back into an OM.

The text format is the one `OM/PrettyPrint.hs:34-106` writes (`prettyPrintA1`, e.g. examples/Life/Generator.hs:30
writes output/OM.txt) and `om/prettyprint.py` mirrors.  With it a real Haskell Paraiso can drive this backend
without GHC being present here: dump the (optimised or not) OM upstream, `parse_om` it, `generate` from it.
Dumps of the current printer carry immediate values (`Imm 3`, `Imm 1.6666666666666667`, `Imm 0.47f`); the older
revision printed `Imm <<Int>>` (`examples-old/*/output/OM.txt`).  Such a dump alone cannot be rebuilt and is rejected;
together with the C++ the same generator run wrote next to it (`dist/<Name>.cpp`) it can: every immediate is printed
there as a literal initialiser of the value node it defines (`int a87_0_0 = 2;`, `(a105) = (0);` —
PlanTrans.hs:527-544,553-555 with ClarisTrans.hs:189-199 for the literal), so `recover_immediates` reads the table
{kernel: {value node id: literal}} from that text and `parse_om(dump, immediates=table)` fills it in.  This makes the
reference's checked-in graph dumps direct inputs of this backend.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional

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
            return Node(inst=Inst("Imm", None))     # legacy dump: the value comes from `immediates` (see parse_om)
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


_LIT = r"-?(?:\d+\.?\d*(?:[eE][-+]?\d+)?f?|true|false)"
_FUNC = re.compile(r"^void\s+\w+::(\w+)\s*\(.*\)\s*\{\s*$")
_LOCAL_IMM = re.compile(r"^\s*(?:const\s+)?\w+\s+a(\d+)(?:_m?\d+)*\s*=\s*\(?(" + _LIT + r")\)?;\s*$")
_MANIFEST_IMM = re.compile(r"^\s*\(+a(\d+)\)(?:\[\w+\]\))?\s*=\s*\((" + _LIT + r")\);\s*$")
_SUBCALL = re.compile(r"^\s*(\w+_sub_\d+)\s*\(")


def recover_immediates(cpp_text: str) -> Dict[str, Dict[int, str]]:
    """{kernel name: {id of the value node an Imm defines: literal text}} from reference-generated C++.

    A subkernel function (`<Name>_sub_<g>` in the old revision, `om_<kernel>_sub_<g>` on master, PlanTrans.hs:261-263)
    initialises each Delayed immediate it needs as a local (`<type> a<id>_<cursor> = <literal>;`, one copy per cursor,
    all equal) and stores a Manifest one directly (`(a<id>) = (<literal>);` / `((a<id>)[addr_origin]) = (<literal>);`).
    Node ids are per kernel; the kernel a subkernel belongs to is read off the kernel functions' bodies
    (`void <Name>::proceed () { <Name>_sub_2(...); ... }`, PlanTrans.hs:225-258)."""
    per_func: Dict[str, Dict[int, str]] = {}
    calls: Dict[str, List[str]] = {}
    cur = None
    for ln in cpp_text.split("\n"):
        m = _FUNC.match(ln)
        if m:
            cur = m.group(1)
            per_func[cur] = {}
            calls[cur] = []
            continue
        if cur is None:
            continue
        m = _LOCAL_IMM.match(ln) or _MANIFEST_IMM.match(ln)
        if m:
            nid, lit = int(m.group(1)), m.group(2)
            old = per_func[cur].setdefault(nid, lit)
            if old != lit:
                raise ValueError(f"{cur}: node {nid} initialised with {old} and {lit}")
            continue
        m = _SUBCALL.match(ln)
        if m:
            calls[cur].append(m.group(1))
    table: Dict[str, Dict[int, str]] = {}
    for kname, subs in calls.items():
        if not subs:
            continue
        table[kname] = {}
        for sname in subs:
            for nid, lit in per_func.get(sname, {}).items():
                old = table[kname].setdefault(nid, lit)
                if old != lit:
                    raise ValueError(f"{kname}: node {nid} initialised with {old} and {lit}")
    return table


def parse_om(text: str, dim: int = None, immediates: Optional[Dict[str, Dict[int, str]]] = None) -> OM:
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
                if nd.inst.arg is None:
                    lit = (immediates or {}).get(kname, {}).get(nd.suc[0])
                    if lit is None:
                        raise ValueError(f"legacy dump without immediate values cannot be imported: kernel {kname}, "
                                         f"node {i} (Imm <<{out.type}>>) has no entry in `immediates`")
                else:
                    lit = nd.inst.arg
                g.nodes[i].inst = Inst("Imm", _imm_content(lit, out.type), imm_type=out.type)
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
