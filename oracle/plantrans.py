"""ORACLE (test infrastructure, not product): CPU restatement of the reference generator's
native C++ backend.

Restates Language/Paraiso/Generator/PlanTrans.hs for Native.language == CPlusPlus:
  * class skeleton, storages, size/margin/variable accessors   PlanTrans.hs:50-215
  * per-kernel driver: subkernel calls in group order, then
    `static = manifest` whole-container copies                  PlanTrans.hs:225-276
  * subkernel body: flat loop over prod(boundarySize), index
    decode, Open `addr+const` / Cyclic `((x+c+n)%n)` addressing,
    `shift v` moves the cursor by -v, loadIndex/loadSize, every
    Delayed value re-evaluated per requested cursor            PlanTrans.hs:406-596
  * operator table                                            PlanTrans.hs:670-710
  * serial om_reduce_*/om_broadcast library                   PlanTrans.hs:719 (draft.cpp:4-21)
  * C++ types and literal formatting                          ClarisTrans.hs:189-199
Deviations, all semantics-free or pinned by SURVEY §8c: canonical OpenMP loop headers (g++ 13
rejects the parenthesised form), `Abs` emitted as std::abs of the value type (g++ 13 binds the
reference's unqualified abs(double) to ::abs(int)), plus an extern "C" shim for ctypes.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use this module.
"""
from __future__ import annotations

from typing import Dict, List, Set, Tuple

import numpy as np

from paraiso_b200 import annotation as A
from paraiso_b200.generator.plan import Plan, SubKernelRef, translate as om_translate
from paraiso_b200.om.graph import ARRAY, CPP_TYPE, SCALAR, Graph, imm_value

CPU_LIB = r"""
template <class T> T om_broadcast (const T& x) {
  return x;
}
template <class T> T om_reduce_sum (const std::vector<T> &xs) {
  T ret = 0;
  for (int i = 0; i < (int)xs.size(); ++i) ret+=xs[i];
  return ret;
}
template <class T> T om_reduce_min (const std::vector<T> &xs) {
  T ret = xs[0];
  for (int i = 1; i < (int)xs.size(); ++i) ret=std::min(ret,xs[i]);
  return ret;
}
template <class T> T om_reduce_max (const std::vector<T> &xs) {
  T ret = xs[0];
  for (int i = 1; i < (int)xs.size(); ++i) ret=std::max(ret,xs[i]);
  return ret;
}
"""


def fmt_imm(content, ctype: str) -> str:
    v = imm_value(content, ctype)
    if ctype == "Bool":
        return "true" if v else "false"
    if ctype in ("Int", "Integer"):
        return str(int(v))
    if ctype == "Double":
        return repr(float(v))
    if ctype == "Float":
        return repr(float(np.float32(v))) + "f"
    raise ValueError(ctype)


def _cursor_text(c: Tuple[int, ...]) -> str:  # PlanTrans.hs:627-640
    return "_".join(str(x).replace("-", "m") for x in c)


def _prod(xs):
    r = 1
    for x in xs:
        r *= x
    return r


class SubEmitter:
    def __init__(self, plan: Plan, sub: SubKernelRef):
        self.plan, self.sub = plan, sub
        self.setup = plan.setup
        self.g: Graph = plan.om.kernels[sub.kernel_idx].dataflow
        self.dim = self.setup.dim
        self.memory_size = list(plan.memory_size)
        self.boundary_size = [m - (sub.lower_boundary[ax] + sub.upper_boundary[ax] if self.setup.boundary[ax] == A.OPEN else 0)
                              for ax, m in enumerate(self.memory_size)]
        self.inputs = set(sub.input_idxs)
        self.outputs = set(sub.output_idxs)
        self.all = sorted(self.inputs | self.outputs | set(sub.calc_idxs))
        self.zero = tuple(0 for _ in range(self.dim))

    # ---- index codecs (PlanTrans.hs:440-490) ------------------------------------------------
    def codec_mod(self) -> List[str]:
        out = []
        for idx in range(self.dim):
            x = "i" if idx == 0 else f"(i / {_prod(self.boundary_size[:idx])})"
            if idx != self.dim - 1:
                x = f"({x} % {self.boundary_size[idx]})"
            out.append(x)
        return out

    def codec_mod_add(self) -> List[str]:
        # DEVIATION: the reference adds the *plan's* lowerMargin here
        # (PlanTrans.hs:446-448) while looping over the subkernel's own boundary box and decoding
        # loadIndex with the subkernel's lowerBoundary (:462).  The two agree whenever
        # lowerBoundary is 0 (shortcut at :450) or equals lowerMargin — all that the checked-in
        # samples exercise — and disagree for partially shrunk boxes (master's Hydro marks HLLC
        # outputs Manifest).  The evident intent, the subkernel's own lower boundary, is used.
        off = [self.sub.lower_boundary[idx] if self.setup.boundary[idx] == A.OPEN else 0 for idx in range(self.dim)]
        return [f"({x} + {off[idx]})" for idx, x in enumerate(self.codec_mod())]

    def codec_addr(self) -> str:
        if self.memory_size == self.boundary_size:
            return "i"
        return " + ".join(f"({x} * {_prod(self.memory_size[:idx])})" for idx, x in enumerate(self.codec_mod_add()))

    def _protect(self, idx, x):
        if self.setup.boundary[idx] == A.OPEN:
            return x
        n = self.memory_size[idx]
        return f"(({x} + {n}) % {n})"

    def codec_load_index(self, cursor, ax) -> str:
        off = self.plan.lower_margin[ax] - self.sub.lower_boundary[ax] - cursor[ax]
        return self._protect(ax, f"({self.codec_mod()[ax]} - ({off}))")

    def codec_cursor(self, cursor) -> str:
        if all(b == A.OPEN for b in self.setup.boundary):
            hard = sum(cursor[idx] * _prod(self.memory_size[:idx]) for idx in range(self.dim))
            return f"addr_origin + ({hard})"
        terms = []
        for idx, x in enumerate(self.codec_mod_add()):
            stride = _prod(self.memory_size[:idx])
            terms.append(f"{stride} * " + self._protect(idx, f"({x} + ({cursor[idx]}))"))
        return " + ".join(terms)

    # ---- rhs and cursor requests (PlanTrans.hs:546-570) -------------------------------------
    def nm(self, idx, cursor) -> str:
        return f"a{idx}_{_cursor_text(cursor)}"

    def rhs_and_request(self, idx, cursor):
        g = self.g
        inst_idx, inst = g.pre_inst(idx)
        prepre = list(g.nodes[inst_idx].pre)
        realm = self.sub.realm
        if idx in self.inputs:
            if realm == ARRAY and g.nodes[idx].value.realm == ARRAY:
                return f"a{idx}[{self.codec_cursor(cursor)}]", []
            return f"a{idx}", []
        op = inst.op
        if op == "Imm":
            return fmt_imm(inst.arg, inst.imm_type), []
        if op == "Arith":
            args = [self.nm(p, cursor) for p in prepre]
            return self.rhs_arith(inst, args, idx), [(p, cursor) for p in prepre]
        if op == "Shift":
            c2 = tuple(c - v for c, v in zip(cursor, inst.arg))
            return self.nm(prepre[0], c2), [(prepre[0], c2)]
        if op == "LoadIndex":
            return self.codec_load_index(cursor, inst.arg), []
        if op == "LoadSize":
            return str(self.setup.local_size[inst.arg]), []
        if op == "Reduce":
            return f"om_reduce_{inst.arg.lower()}(a{prepre[0]})", []
        if op == "Broadcast":
            return f"om_broadcast(a{prepre[0]})", []
        raise ValueError(op)

    def rhs_arith(self, inst, a, idx) -> str:  # PlanTrans.hs:670-710
        op = inst.arg
        infix = {"Add": "+", "Sub": "-", "Mul": "*", "Div": "/", "Mod": "%", "And": "&&", "Or": "||",
                 "EQ": "==", "NE": "!=", "LT": "<", "LE": "<=", "GT": ">", "GE": ">="}
        if op == "Identity":
            return a[0]
        if op in infix:
            return f"({a[0]}) {infix[op]} ({a[1]})"
        if op == "Neg":
            return f"-({a[0]})"
        if op == "Inv":
            return f"1/({a[0]})"
        if op == "Not":
            return f"!({a[0]})"
        if op == "Select":
            return f"({a[0]}) ? ({a[1]}) : ({a[2]})"
        if op == "Max":
            return f"std::max({a[0]}, {a[1]})"
        if op == "Min":
            return f"std::min({a[0]}, {a[1]})"
        if op == "Abs":
            return f"std::abs({a[0]})"   # pinned semantics, see module docstring
        if op in ("Sqrt", "Exp", "Log", "Sin", "Cos", "Tan", "Asin", "Acos", "Atan", "Atan2"):
            return f"{op.lower()}({', '.join(a)})"
        if op == "Cast":
            return f"({CPP_TYPE[inst.cast_to]})({a[0]})"
        return f"{op.lower()}({', '.join(a)})"

    def ctype(self, idx) -> str:
        return CPP_TYPE[self.g.nodes[idx].value.type]

    def loop_content(self) -> List[str]:
        g = self.g
        vals = [i for i in self.all if g.nodes[i].is_value]
        # lhsCursors: outputs at the origin; everything else at the cursors requested by later nodes
        cursors: Dict[int, Set[Tuple[int, ...]]] = {i: set() for i in vals}
        for i in vals:
            if i in self.outputs:
                cursors[i].add(self.zero)
        for j in sorted(vals, reverse=True):
            for cur in list(cursors[j]):
                for (p, c2) in self.rhs_and_request(j, cur)[1]:
                    if p in cursors:
                        cursors[p].add(c2)
        lines = []
        for i in vals:
            for cur in sorted(cursors[i], key=lambda c: tuple(reversed(c))):
                expr, _ = self.rhs_and_request(i, cur)
                if i in self.outputs:
                    if self.sub.realm == ARRAY:
                        lines.append(f"(a{i})[addr_origin] = ({expr});")
                    else:
                        lines.append(f"(a{i}) = ({expr});")
                else:
                    lines.append(f"{self.ctype(i)} {self.nm(i, cur)} = {expr};")
        return lines

    def args(self) -> str:
        g = self.g
        out = []
        for idx in self.sub.input_idxs:
            dv = g.nodes[idx].value
            t = CPP_TYPE[dv.type] if dv.realm == SCALAR else f"std::vector<{CPP_TYPE[dv.type]}> "
            out.append(f"const {t} & a{idx}")
        for idx in self.sub.output_idxs:
            dv = g.nodes[idx].value
            t = CPP_TYPE[dv.type] if dv.realm == SCALAR else f"std::vector<{CPP_TYPE[dv.type]}> "
            out.append(f"{t} & a{idx}")
        return ", ".join(out)

    def body(self) -> str:
        hdr = (f"/*\nlowerMargin = {self.sub.lower_boundary}\nupperMargin = {self.sub.upper_boundary}\n*/\n")
        if self.sub.realm == SCALAR:
            return "\n".join(self.loop_content()) + "\n"
        n = _prod(self.boundary_size)
        s = hdr + "#pragma omp parallel for\n"
        s += f"for (int i = 0; i < {n}; i += 1) {{\n"
        s += f"int addr_origin = {self.codec_addr()};\n"
        s += "\n".join(self.loop_content()) + "\n}\n"
        return s


def emit(plan: Plan) -> str:
    """One self-contained C++ translation unit: the reference-style class + a C shim."""
    name = plan.name
    om = plan.om
    mem = plan.memory_size
    dim = plan.setup.dim
    out: List[str] = ["// GENERATED by oracle/plantrans.py — reference-style native C++ (test oracle / CPU baseline)",
                      "#include <algorithm>", "#include <cmath>", "#include <cstdlib>", "#include <vector>", "#include <cstring>",
                      CPU_LIB]

    def cpptype(dv):
        return CPP_TYPE[dv.type] if dv.realm == SCALAR else f"std::vector<{CPP_TYPE[dv.type]}> "

    out.append(f"/*\nlowerMargin = {plan.lower_margin}\nupperMargin = {plan.upper_margin}\n*/")
    out.append(f"class {name} {{")
    for st in plan.storages:
        out.append(f"public: {cpptype(st.dyn)} {st.name};")
    inits = [f"{st.name}(om_memory_size())" for st in plan.storages if st.dyn.realm == ARRAY]
    sinits = [f"{st.name}()" for st in plan.storages if st.dyn.realm == SCALAR]
    out.append(f"public: {name} () : " + ",".join(inits + sinits) + " {}")

    def size_funcs(prefix, vec):
        out.append(f"public: int {prefix} () {{ return {_prod(vec)}; }}")
        for i, v in enumerate(vec):
            out.append(f"public: int {prefix}_{i} () {{ return {v}; }}")
    size_funcs("om_size", plan.setup.local_size)
    size_funcs("om_memory_size", mem)
    for i in range(dim):
        out.append(f"public: int om_lower_margin_{i} () {{ return {plan.lower_margin[i]}; }}")
        out.append(f"public: int om_upper_margin_{i} () {{ return {plan.upper_margin[i]}; }}")
    # accessors (PlanTrans.hs:117-157)
    for st in plan.storages:
        if st.static_idx is None:
            continue
        out.append(f"public: {cpptype(st.dyn)} & {st.static_name} () {{ return {st.name}; }}")
        if st.dyn.realm == ARRAY:
            args = ", ".join(f"int i{k}" for k in range(dim))
            expr = ""
            for k in reversed(range(dim)):
                term = f"(om_lower_margin_{k}() + i{k})"
                expr = term if expr == "" else f"{term} + om_memory_size_{k}() * ({expr})"
            out.append(f"public: {CPP_TYPE[st.dyn.type]} & {st.static_name} ({args}) {{ return {st.name}[{expr}]; }}")
    subs = [SubEmitter(plan, s) for s in plan.sub_kernels]
    for se in subs:
        out.append(f"public: void {se.sub.name} ({se.args()});")
    for k in om.kernels:
        out.append(f"public: void {k.name} ();")
    out.append("};")
    for se in subs:
        out.append(f"void {name}::{se.sub.name} ({se.args()}) {{\n{se.body()}}}")
    # kernel drivers (PlanTrans.hs:225-276)
    for kidx, k in enumerate(om.kernels):
        g = k.dataflow
        body = []

        def find_var(idx):
            load_idx = None
            for j in g.nodes[idx].pre:
                nd = g.nodes[j]
                if nd.inst is not None and nd.inst.op == "Load":
                    load_idx = nd.inst.arg
            for st in plan.storages:
                if st.manifest == (kidx, idx) or (load_idx is not None and st.static_idx == load_idx):
                    return st.name
            raise KeyError(idx)
        for se in subs:
            if se.sub.kernel_idx != kidx:
                continue
            body.append(f"{se.sub.name}(" + ", ".join(find_var(i) for i in se.sub.input_idxs + se.sub.output_idxs) + ");")
        for idx, nd in enumerate(g.nodes):
            if nd.inst is not None and nd.inst.op == "Store":
                pre = nd.pre[0]
                st = [s for s in plan.storages if s.static_idx == nd.inst.arg][0]
                ma = [s for s in plan.storages if s.manifest == (kidx, pre)][0]
                body.append(f"({st.name}) = ({ma.name});")
        out.append(f"void {name}::{k.name} () {{\n" + "\n".join(body) + "\n}")
    # C shim
    out.append('extern "C" {')
    out.append(f"void* om_new() {{ return new {name}(); }}")
    out.append(f"void om_delete(void* p) {{ delete ({name}*)p; }}")
    out.append(f"int om_memory_size_k(void* p, int k) {{ {name}* s=({name}*)p; " +
               " ".join(f"if (k=={i}) return s->om_memory_size_{i}();" for i in range(dim)) + " return s->om_memory_size(); }")
    out.append(f"int om_lower_margin_k(void* p, int k) {{ {name}* s=({name}*)p; " +
               " ".join(f"if (k=={i}) return s->om_lower_margin_{i}();" for i in range(dim)) + " return -1; }")
    out.append(f"int om_upper_margin_k(void* p, int k) {{ {name}* s=({name}*)p; " +
               " ".join(f"if (k=={i}) return s->om_upper_margin_{i}();" for i in range(dim)) + " return -1; }")
    cases = []
    for st in plan.storages:
        if st.static_idx is None:
            continue
        ptr = f"(void*)s->{st.name}.data()" if st.dyn.realm == ARRAY else f"(void*)&s->{st.name}"
        cases.append(f"if (idx=={st.static_idx}) return {ptr};")
    out.append(f"void* om_static_ptr(void* p, int idx) {{ {name}* s=({name}*)p; " + " ".join(cases) + " return 0; }")
    kc = " ".join(f'if (!strcmp(k,"{k.name}")) {{ s->{k.name}(); return 0; }}' for k in om.kernels)
    out.append(f"int om_call(void* p, const char* k) {{ {name}* s=({name}*)p; {kc} return -1; }}")
    out.append("}")
    return "\n".join(out) + "\n"


def generate(setup, om) -> str:
    return emit(om_translate(setup, om))
