"""Emit sm_100a CUDA for the stages planned by schedule.py.

What the reference emits per subkernel is a flat grid-stride loop that re-reads global memory
for every (input, cursor) pair (PlanTrans.hs:295-314, 406-428, 468-484).  What is emitted here,
per stage, is one `__global__` kernel in which a CTA owns a strip of NT*V columns and streams
along axis 1:

    for each row iteration j:
        cp.async the next input rows into shared-memory rings        (LDGSTS, zero fill off-array)
        wait + barrier
        phase 1: every MAT scope of level 1 computes its row j+lag   -> shared-memory rings
        barrier
        phase 2: ...
        OUT scope: stores (128-bit where aligned) + reduce accumulation in registers
    block reduce -> per-CTA partial -> last CTA folds partials -> device scalar slot

plus a one-thread kernel for the Scalar-realm part of the OM kernel, and `extern "C"` launchers.
Per-cell arithmetic is the OM's SSA DAG verbatim (operator table of PlanTrans.hs:670-710 with
std::max/std::min operand order spelled out), so with -fmad=false results are bit-identical to
the reference's C++.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from ... import annotation as A
from ...om.graph import ARRAY, CPP_TYPE, SCALAR, TYPE_BYTES, OM, imm_value
from ..native import Setup
from ..plan import Plan
from .schedule import KernelSchedule, Op, Stage, schedule_kernel

PF = 2  # cp.async prefetch distance in rows


def c_imm(content, ctype: str) -> str:
    v = imm_value(content, ctype)
    if ctype == "Bool":
        return "true" if v else "false"
    if ctype in ("Int", "Integer"):
        return f"({int(v)})"
    if ctype == "Double":
        return f"({float(v)!r})" if np.isfinite(v) else str(v)
    if ctype == "Float":
        return f"({float(np.float32(v))!r}f)"
    raise ValueError(ctype)


def _ru(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _cur(c) -> str:
    return "_".join(str(x).replace("-", "m") for x in c)


def arith_expr(op: Op, a: List[str]) -> str:
    """C expression for an Arith node (PlanTrans.hs:670-710)."""
    o = op.inst.arg
    infix = {"Add": "+", "Sub": "-", "Mul": "*", "Div": "/", "Mod": "%", "And": "&&", "Or": "||",
             "EQ": "==", "NE": "!=", "LT": "<", "LE": "<=", "GT": ">", "GE": ">="}
    T = CPP_TYPE[op.ctype]
    if o == "Identity":
        return a[0]
    if o in infix:
        return f"({a[0]} {infix[o]} {a[1]})"
    if o == "Neg":
        return f"(-{a[0]})"
    if o == "Inv":
        return f"(({T})1 / {a[0]})"
    if o == "Not":
        return f"(!{a[0]})"
    if o == "Select":
        return f"({a[0]} ? {a[1]} : {a[2]})"
    if o == "Max":   # std::max(a, b) == (a < b) ? b : a
        return f"(({a[0]} < {a[1]}) ? {a[1]} : {a[0]})"
    if o == "Min":   # std::min(a, b) == (b < a) ? b : a
        return f"(({a[1]} < {a[0]}) ? {a[1]} : {a[0]})"
    if o == "Abs":
        return f"abs({a[0]})" if op.ctype in ("Int", "Integer") else f"fabs({a[0]})" if op.ctype == "Double" else f"fabsf({a[0]})"
    if o == "Signum":
        return f"(({T})(({a[0]} > 0) - ({a[0]} < 0)))"
    if o == "Cast":
        return f"(({CPP_TYPE[op.inst.cast_to]}){a[0]})"
    if o in ("Sqrt", "Exp", "Log", "Sin", "Cos", "Tan", "Asin", "Acos", "Atan", "Atan2", "Pow"):
        f = o.lower()
        if op.ctype == "Float":
            f += "f"
        return f"{f}({', '.join(a)})"
    if o == "Ipow":
        return f"(({T})pow((double){a[0]}, (double){a[1]}))"
    if o == "Madd":
        return f"(({a[0]} * {a[1]}) + {a[2]})"
    if o == "Msub":
        return f"(({a[0]} * {a[1]}) - {a[2]})"
    if o == "Nmadd":
        return f"(-(({a[0]} * {a[1]}) + {a[2]}))"
    if o == "Nmsub":
        return f"(-(({a[0]} * {a[1]}) - {a[2]}))"
    raise NotImplementedError(o)


class StageEmitter:
    def __init__(self, om: OM, plan: Plan, ks: KernelSchedule, st: Stage, stage_idx: int, V: int, NT: int):
        self.om, self.plan, self.ks, self.st, self.idx = om, plan, ks, st, stage_idx
        self.ops = ks.ops
        self.V, self.NT = V, NT
        self.HL, self.HR = _ru(st.halo_x[0], V), _ru(st.halo_x[1], V)
        self.PL, self.PR = _ru(st.pad_x[0], V), _ru(st.pad_x[1], V)
        self.W_OUT = NT * V - self.HL - self.HR
        assert self.W_OUT > 0
        self.RW = NT * V + self.PL + self.PR
        self.name = f"om_{om.name}_{ks.name}_stage{stage_idx}"
        self.lines: List[str] = []
        self.ring_inputs = [i for i in st.inputs.values() if i.via_smem]
        self.direct_inputs = [i for i in st.inputs.values() if not i.via_smem]
        # ring depths: inputs get PF + 1 extra rows for the in-flight async copies
        self.depth: Dict[int, int] = {}
        for i in self.ring_inputs:
            self.depth[i.vid] = i.depth + PF + 1
        for m in st.mats.values():
            self.depth[m.vid] = m.depth
        self.lag: Dict[int, int] = {i.vid: i.lag for i in st.inputs.values()}
        self.lag.update({m.vid: m.lag for m in st.mats.values()})
        self.static_of = {i.vid: i.static_idx for i in st.inputs.values()}
        self.margin_lo = plan.lower_margin + (0,) * (2 - len(plan.lower_margin))
        self.margin_hi = plan.upper_margin + (0,) * (2 - len(plan.upper_margin))

    # ------------------------------------------------------------------------------------------
    def T(self, v) -> str:
        return CPP_TYPE[self.ops[v].ctype]

    def smem_bytes(self) -> int:
        tot = 0
        for v, d in self.depth.items():
            tot = _ru(tot, 16) + d * self.RW * TYPE_BYTES[self.ops[v].ctype]
        return _ru(tot, 16)

    def emit(self, s=""):
        self.lines.append(s)

    # ---- scalar (uniform) values ------------------------------------------------------------
    def scalar_code(self, roots: List[int]) -> List[str]:
        """Evaluate Scalar-realm values from the device scalar table, in id order."""
        ops = self.ops
        need: List[int] = []
        seen = set()
        stack = list(roots)
        while stack:
            v = stack.pop()
            if v in seen:
                continue
            seen.add(v)
            op = ops[v]
            if op.kind in ("Reduce",):
                continue
            stack.extend(a for a in op.args)
        out = []
        for v in sorted(seen):
            op = ops[v]
            T = CPP_TYPE[op.ctype]
            if op.kind == "Load":
                out.append(f"const {T} s{v} = om_slot_load<{T}>(sc, {op.inst.arg});")
            elif op.kind == "Reduce":
                out.append(f"const {T} s{v} = om_slot_load<{T}>(sc, {self.ks.reduce_slots[v]});")
            elif op.kind == "Imm":
                out.append(f"const {T} s{v} = {c_imm(op.inst.arg, op.ctype)};")
            elif op.kind == "LoadSize":
                out.append(f"const {T} s{v} = ({T}){'g.nx' if op.inst.arg == 0 else 'g.ny'};")
            elif op.kind == "Arith":
                out.append(f"const {T} s{v} = {arith_expr(op, ['s%d' % a for a in op.args])};")
            else:
                raise NotImplementedError(f"scalar {op.kind}")
        return out

    # ---- array values inside one scope ------------------------------------------------------
    def scope(self, targets: List[int], row: str, is_out: bool) -> Tuple[List[str], Dict]:
        """SSA statements computing `targets` at cursor 0 for the V lanes of this thread at device
        row `row`.  Returns (lines, {(vid, lane): expr})."""
        ops, V = self.ops, self.V
        mats = self.st.mats
        lines: List[str] = []
        memo: Dict[Tuple[int, Tuple[int, int], int], str] = {}
        ring_rd: Dict[Tuple[int, int, int], str] = {}
        slot_rd: Dict[Tuple[int, int], str] = {}
        local_mats: Dict[int, None] = {}   # MAT targets computed in this scope (usable at cursor 0)
        target_set = set(targets)

        def slot(b, cy):
            key = (b, cy)
            if key not in slot_rd:
                nm = f"sl{b}_{str(cy).replace('-', 'm')}"
                lines.append(f"const int {nm} = ({row} + {cy} + {1 << 20} * {self.depth[b]}) % {self.depth[b]};")
                slot_rd[key] = nm
            return slot_rd[key]

        def ring_read(b, cur, k):
            o = k + cur[0]
            key = (b, cur[1], o)
            if key in ring_rd:
                return ring_rd[key]
            T = self.T(b)
            sl = slot(b, cur[1])
            nm = f"r{b}_{str(cur[1]).replace('-', 'm')}_{str(o).replace('-', 'm')}"
            bytes_ = TYPE_BYTES[ops[b].ctype]
            vt = {("int", 4): "int4", ("float", 4): "float4", ("double", 2): "double2",
                  ("int", 2): "int2", ("float", 2): "float2"}.get((T, V))
            if 0 <= o < V and vt:
                vn = f"rv{b}_{str(cur[1]).replace('-', 'm')}"
                lines.append(f"const {vt} {vn} = *reinterpret_cast<const {vt}*>(&ring{b}[{sl} * RW + PL + tid * V]);")
                for kk in range(V):
                    n2 = f"r{b}_{str(cur[1]).replace('-', 'm')}_{kk}"
                    lines.append(f"const {T} {n2} = {vn}.{'xyzw'[kk]};")
                    ring_rd[(b, cur[1], kk)] = n2
                return ring_rd[key]
            lines.append(f"const {T} {nm} = ring{b}[{sl} * RW + PL + tid * V + ({o})];")
            ring_rd[key] = nm
            return nm

        def val(v, cur, k) -> str:
            key = (v, cur, k)
            if key in memo:
                return memo[key]
            op = ops[v]
            T = CPP_TYPE[op.ctype]
            if op.realm == SCALAR:
                memo[key] = f"s{v}"
                return memo[key]
            if v in mats and not (v in target_set and cur == (0, 0)):
                if v in local_mats and cur == (0, 0):
                    e = memo[(v, (0, 0), k)]
                else:
                    e = ring_read(v, cur, k)
                memo[key] = e
                return e
            if op.kind == "Load":
                inp = self.st.inputs[v]
                if inp.via_smem:
                    e = ring_read(v, cur, k)
                else:
                    assert cur == (0, 0)
                    e = f"d{v}_{k}"
                    if (v, "direct") not in memo:
                        memo[(v, "direct")] = "1"
                        lines.extend(self.direct_load(v, row))
                memo[key] = e
                return e
            if op.kind == "Imm":
                e = c_imm(op.inst.arg, op.ctype)
            elif op.kind == "Broadcast":
                e = f"s{op.args[0]}"
            elif op.kind == "LoadIndex":
                ax = op.inst.arg
                nm = f"ix{ax}_{_cur(cur)}_{k}"
                if (nm, "def") not in memo:
                    memo[(nm, "def")] = "1"
                    if ax == 0:
                        lines.append(f"const int {nm} = g.cyc_x ? om_wrap(tc + {k} + ({cur[0]}) - g.xorg, g.nx) : (tc + {k} + ({cur[0]}) - g.xorg);")
                    else:
                        lines.append(f"const int {nm} = g.cyc_y ? om_wrap({row} + ({cur[1]}) - g.yorg + g.y0, g.ny) : ({row} + ({cur[1]}) - g.yorg + g.y0);")
                e = f"(({T}){nm})"
            elif op.kind == "LoadSize":
                e = f"(({T}){'g.nx' if op.inst.arg == 0 else 'g.ny'})"
            elif op.kind == "Shift":
                s = tuple(op.inst.arg) + (0,) * (2 - len(op.inst.arg))
                e = val(op.args[0], (cur[0] - s[0], cur[1] - s[1]), k)
                memo[key] = e
                return e
            elif op.kind == "Arith":
                args = [val(a, cur, k) for a in op.args]
                e = arith_expr(op, args)
            else:
                raise NotImplementedError(op.kind)
            nm = f"v{v}_{_cur(cur)}_{k}"
            lines.append(f"const {T} {nm} = {e};")
            memo[key] = nm
            return nm

        result = {}
        for t in sorted(targets):
            for k in range(V):
                result[(t, k)] = val(t, (0, 0), k)
            if t in mats:
                local_mats[t] = None
        return lines, result

    def direct_load(self, v, row) -> List[str]:
        V = self.V
        T = self.T(v)
        sidx = self.static_of[v]
        ls = [f"{T} " + ", ".join(f"d{v}_{k} = 0" for k in range(V)) + ";"]
        ls.append(f"if ({row} >= 0 && {row} < g.rows) {{")
        ls.append(f"  const {T}* __restrict__ p = in{sidx} + (size_t){row} * g.pitch;")
        for k in range(V):
            ls.append(f"  if (tc + {k} >= 0 && tc + {k} < g.pitch) d{v}_{k} = __ldg(p + tc + {k});")
        ls.append("}")
        return ls

    # ---- whole kernel ---------------------------------------------------------------------------
    def kernel(self) -> str:
        st, V, NT = self.st, self.V, self.NT
        om = self.om
        E = self.emit
        in_statics = sorted({i.static_idx for i in st.inputs.values()})
        out_statics = [s for (s, _v) in st.store_targets]
        sv = om.setup.static_values
        params = ["const OmGeom g"]
        for s in in_statics:
            params.append(f"const {CPP_TYPE[sv[s].namee.type]}* __restrict__ in{s}")
        for s in out_statics:
            params.append(f"{CPP_TYPE[sv[s].namee.type]}* __restrict__ out{s}")
        params += ["om_slot_t* __restrict__ sc", "unsigned* __restrict__ red_counter", "om_slot_t* __restrict__ red_partials"]
        E(f"// stage {self.idx} of kernel `{self.ks.name}` (reduce level {st.level}): "
          f"{len(st.mats)} shared-memory rings for intermediates, {len(self.ring_inputs)} for inputs, "
          f"{len(st.phases)} phase(s), warm-up {st.warmup} rows")
        E(f"__global__ void __launch_bounds__({NT}) {self.name}_kernel({', '.join(params)}) {{")
        E(f"  constexpr int V = {V}, NT = {NT}, HL = {self.HL}, PL = {self.PL}, RW = {self.RW}, W_OUT = {self.W_OUT};")
        E("  const int tid = threadIdx.x;")
        E("  OM_DYNAMIC_SMEM(om_smem);")
        off = 0
        for v, d in self.depth.items():
            off = _ru(off, 16)
            E(f"  {self.T(v)}* const ring{v} = reinterpret_cast<{self.T(v)}*>(om_smem + {off});  // {d} rows")
            off += d * self.RW * TYPE_BYTES[self.ops[v].ctype]
        # column geometry: memory box [cx0, cx1), strips start at a V-aligned column
        mlx, mhx = self.margin_lo[0], self.margin_hi[0]
        mly, mhy = self.margin_lo[1], self.margin_hi[1]
        E(f"  const int cx0 = g.xorg - {mlx}, cx1 = g.xorg + g.nx + {mhx};")
        E("  const int cA = (cx0 / V) * V;")
        E("  const int strip_lo = cA + blockIdx.x * W_OUT;          // first output column of this CTA")
        E("  const int tc = strip_lo - HL + tid * V;                  // first column of this thread")
        E("  const int r0 = g.own_r0 + blockIdx.y * g.chunk_rows;")
        E("  const int r1 = min(r0 + g.chunk_rows, g.own_r1);")
        # uniform scalars
        roots = list(dict.fromkeys(st.scalar_roots))
        for l in self.scalar_code(roots):
            E("  " + l)
        # reduce accumulators
        for (v, rop, slot) in st.reduce_targets:
            T = self.T(v)
            ident = {"Sum": f"({T})0", "Min": self.type_max(v), "Max": self.type_min(v)}[rop]
            E(f"  {T} acc{v} = {ident};")
        warm = st.warmup
        has_ring_in = bool(self.ring_inputs)
        lead = warm + (PF if has_ring_in else 0)
        E(f"  for (int j = r0 - {lead}; j < r1; ++j) {{")
        if has_ring_in:
            E("    // ---- stage the next input rows (LDGSTS); off-array cells are zero-filled")
            for i in self.ring_inputs:
                v = i.vid
                T = self.T(v)
                B = TYPE_BYTES[i.ctype] * V
                E(f"    {{ const int rr = j + {i.lag + PF};")
                E(f"      const int sl = (rr + {1 << 20} * {self.depth[v]}) % {self.depth[v]};")
                E(f"      const bool rok = (rr >= 0) && (rr < g.rows);")
                E(f"      const {T}* __restrict__ src = in{i.static_idx} + (size_t)(rok ? rr : 0) * g.pitch;")
                E(f"      {{ const bool ok = rok && (tc >= 0) && (tc + V <= g.pitch);")
                E(f"        om_cp_async<{B}>(&ring{v}[sl * RW + PL + tid * V], src + (ok ? tc : 0), ok ? {B} : 0); }}")
                if self.PL:
                    E(f"      if (tid < {self.PL // V}) {{ const int c = strip_lo - HL - PL + tid * V; const bool ok = rok && (c >= 0) && (c + V <= g.pitch);")
                    E(f"        om_cp_async<{B}>(&ring{v}[sl * RW + tid * V], src + (ok ? c : 0), ok ? {B} : 0); }}")
                if self.PR:
                    E(f"      if (tid < {self.PR // V}) {{ const int c = strip_lo - HL + NT * V + tid * V; const bool ok = rok && (c >= 0) && (c + V <= g.pitch);")
                    E(f"        om_cp_async<{B}>(&ring{v}[sl * RW + PL + NT * V + tid * V], src + (ok ? c : 0), ok ? {B} : 0); }}")
                E("    }")
            E("    om_cp_async_commit();")
            E(f"    om_cp_async_wait<{PF}>();")
        E("    __syncthreads();")
        nph = max(len(st.phases), st.out_level)
        for lvl in range(1, nph + 1):
            if lvl > 1:
                E("    __syncthreads();")
            mats_here = st.phases[lvl - 1] if lvl - 1 < len(st.phases) else []
            # scopes: MAT nodes sharing a lag share SSA values
            lags = sorted({st.mats[m].lag for m in mats_here}, key=lambda a: min(m for m in mats_here if st.mats[m].lag == a))
            for a in lags:
                grp = [m for m in mats_here if st.mats[m].lag == a]
                early = min(st.mats[m].early for m in grp)
                E(f"    if (j >= r0 - {-early}) {{   // phase {lvl}: rows j+{a} of {len(grp)} intermediate(s)")
                E(f"      const int row = j + {a};")
                lines, res = self.scope(grp, "row", False)
                for l in lines:
                    E("      " + l)
                for m in grp:
                    E(f"      {{ const int sl = (row + {1 << 20} * {self.depth[m]}) % {self.depth[m]};")
                    for k in range(V):
                        E(f"        ring{m}[sl * RW + PL + tid * V + {k}] = {res[(m, k)]};")
                    E("      }")
                E("    }")
            if lvl == st.out_level:
                self.emit_out()
        E("  }")
        self.emit_reduce_epilogue()
        E("}")
        return "\n".join(self.lines)

    def type_max(self, v) -> str:
        return {"Int": "2147483647", "Float": "__int_as_float(0x7f800000)", "Double": "__longlong_as_double(0x7ff0000000000000LL)",
                "Integer": "9223372036854775807LL"}[self.ops[v].ctype]

    def type_min(self, v) -> str:
        return {"Int": "(-2147483647-1)", "Float": "__int_as_float(0xff800000)", "Double": "__longlong_as_double(0xfff0000000000000LL)",
                "Integer": "(-9223372036854775807LL-1)"}[self.ops[v].ctype]

    def valid_box(self, v) -> Tuple[int, int, int, int]:
        """(lb_x, ub_x, lb_y, ub_y) of the node's Valid region in memory-box coordinates
        (OMTrans.hs:103-116)."""
        from ..plan import _valid_to_lower, _valid_to_upper
        valid = self.ops[v].valid
        lo = _valid_to_lower(self.plan.setup, valid)
        hi = _valid_to_upper(self.plan.setup, valid)
        lo = tuple(lo) + (0,) * (2 - len(lo))
        hi = tuple(hi) + (0,) * (2 - len(hi))
        return lo[0], hi[0], lo[1], hi[1]

    def emit_out(self):
        st, V, E = self.st, self.V, self.emit
        targets = [v for (_s, v) in st.store_targets] + [v for (v, _o, _k) in st.reduce_targets]
        targets = list(dict.fromkeys(targets))
        mlx, mhx = self.margin_lo[0], self.margin_hi[0]
        mly, mhy = self.margin_lo[1], self.margin_hi[1]
        E("    if (j >= r0) {   // OUT: stores and reduce accumulation for row j")
        E("      const int row = j;")
        lines, res = self.scope(targets, "row", True)
        for l in lines:
            E("      " + l)
        # global memory-box row of this device row (reference memory coordinates)
        E(f"      const int gmy = row - g.yorg + g.y0 + {mly};")
        E(f"      const int memy = g.ny + {mly + mhy};")
        for k in range(V):
            E(f"      const bool in{k} = (tc + {k} >= max(cx0, strip_lo)) && (tc + {k} < min(cx1, strip_lo + W_OUT));")
        all_in = " && ".join(f"in{k}" for k in range(V))
        for v in targets:
            lbx, ubx, lby, uby = self.valid_box(v)
            T = self.T(v)
            E(f"      const bool vy{v} = (gmy >= {lby}) && (gmy < memy - {uby});")
            for k in range(V):
                E(f"      const {T} o{v}_{k} = (vy{v} && (tc + {k} >= cx0 + {lbx}) && (tc + {k} < cx1 - {ubx})) ? {res[(v, k)]} : ({T})0;")
        for (s, v) in st.store_targets:
            T = self.T(v)
            bytes_ = TYPE_BYTES[self.ops[v].ctype] * V
            E(f"      {{ {T}* __restrict__ p = out{s} + (size_t)row * g.pitch + tc;")
            vt = None
            if V > 1 and bytes_ == 16:
                vt = {"int": "int4", "float": "float4", "double": "double2"}.get(T)
            elif V > 1 and bytes_ == 8:
                vt = {"int": "int2", "float": "float2"}.get(T)
            if vt:
                comps = ", ".join(f"o{v}_{k}" for k in range(V))
                E(f"        if ({all_in}) {{ *reinterpret_cast<{vt}*>(p) = make_{vt}({comps}); }}")
                E("        else {")
                for k in range(V):
                    E(f"          if (in{k}) p[{k}] = o{v}_{k};")
                E("        }")
            else:
                for k in range(V):
                    E(f"        if (in{k}) p[{k}] = o{v}_{k};")
            # fused ghost-cell writes for Cyclic axes (the wrap the reference computes with % per read,
            # PlanTrans.hs:477-484, is materialised once per written cell here)
            E("        const bool ex = g.cyc_x && (strip_lo < g.xorg + g.gx_hi || strip_lo + W_OUT > g.xorg + g.nx - g.gx_lo);")
            E("        const bool ey = g.wrap_y_local && (row < g.yorg + g.gy_hi || row >= g.yorg + g.nyl - g.gy_lo);")
            E("        if (ex || ey) {")
            for k in range(V):
                E(f"          if (in{k}) {{ const int c = tc + {k} - g.xorg; const int r = row - g.yorg;")
                E("            const int dc = !g.cyc_x ? 0 : (c < g.gx_hi ? g.nx : (c >= g.nx - g.gx_lo ? -g.nx : 0));")
                E("            const int dr = !g.wrap_y_local ? 0 : (r < g.gy_hi ? g.nyl : (r >= g.nyl - g.gy_lo ? -g.nyl : 0));")
                E(f"            if (dc) p[{k} + dc] = o{v}_{k};")
                E(f"            if (dr) p[{k} + (ptrdiff_t)dr * g.pitch] = o{v}_{k};")
                E(f"            if (dc && dr) p[{k} + dc + (ptrdiff_t)dr * g.pitch] = o{v}_{k};")
                # a domain narrower than the ghost width wraps from both sides
                E("            if (g.cyc_x && c < g.gx_hi && c >= g.nx - g.gx_lo) p[%d - g.nx] = o%d_%d;" % (k, v, k))
                E("            if (g.wrap_y_local && r < g.gy_hi && r >= g.nyl - g.gy_lo) p[%d - (ptrdiff_t)g.nyl * g.pitch] = o%d_%d;" % (k, v, k))
                E("          }")
            E("        }")
            E("      }")
        for (v, rop, slot) in st.reduce_targets:
            cls = {"Sum": "OmSum", "Min": "OmMin", "Max": "OmMax"}[rop]
            for k in range(V):
                E(f"      if (in{k}) acc{v} = {cls}::op(acc{v}, o{v}_{k});")
        E("    }")

    def emit_reduce_epilogue(self):
        st, E, NT = self.st, self.emit, self.NT
        if not st.reduce_targets:
            return
        E("  // ---- block reduce -> per-CTA partial -> last CTA folds all partials (om_runtime.cuh)")
        for t, (v, rop, slot) in enumerate(st.reduce_targets):
            T = self.T(v)
            cls = {"Sum": "OmSum", "Min": "OmMin", "Max": "OmMax"}[rop]
            ident = {"Sum": f"({T})0", "Min": self.type_max(v), "Max": self.type_min(v)}[rop]
            E(f"  {{ __shared__ {T} red{v}[32]; {T} result;")
            E(f"    {T}* partials = reinterpret_cast<{T}*>(red_partials + (size_t){t} * gridDim.x * gridDim.y);")
            E(f"    if (om_block_reduce_finalize<{cls}, {T}, NT>(acc{v}, {ident}, partials, red_counter + {t}, red{v}, result)) {{")
            E(f"      om_slot_store<{T}>(sc, {slot}, result);")
            E(f"      red_counter[{t}] = 0u;")
            E("    }")
            E("    __syncthreads();")
            E("  }")

    def launcher(self) -> str:
        st = self.st
        sv = self.om.setup.static_values
        in_statics = sorted({i.static_idx for i in st.inputs.values()})
        out_statics = [s for (s, _v) in st.store_targets]
        args = ["*g"]
        for s in in_statics:
            args.append(f"(const {CPP_TYPE[sv[s].namee.type]}*)cur[{s}]")
        for s in out_statics:
            args.append(f"({CPP_TYPE[sv[s].namee.type]}*)alt[{s}]")
        args += ["(om_slot_t*)sc", "(unsigned*)scratch", "(om_slot_t*)((char*)scratch + 256)"]
        smem = self.smem_bytes()
        mlx, mhx = self.margin_lo[0], self.margin_hi[0]
        L = []
        L.append(f'extern "C" int {self.name}(const OmGeom* g, void* const* cur, void* const* alt, void* sc, void* scratch, void* stream) {{')
        L.append(f"  const int cx0 = g->xorg - {mlx}, cx1 = g->xorg + g->nx + {mhx};")
        L.append(f"  const int cA = (cx0 / {self.V}) * {self.V};")
        L.append(f"  const int strips = (cx1 - cA + {self.W_OUT} - 1) / {self.W_OUT};")
        L.append("  const int nrows = g->own_r1 - g->own_r0;")
        L.append("  if (nrows <= 0 || strips <= 0) return 0;")
        L.append("  const int chunks = (nrows + g->chunk_rows - 1) / g->chunk_rows;")
        L.append(f"  static bool attr_set = false;")
        L.append(f"  if (!attr_set) {{ cudaError_t e = cudaFuncSetAttribute({self.name}_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, {max(smem, 1)}); if (e != cudaSuccess) return (int)e; attr_set = true; }}")
        L.append(f"  OM_LAUNCH({self.name}_kernel, dim3(strips, chunks), {self.NT}, {smem}, (cudaStream_t)stream, {', '.join(args)});")
        L.append("  OM_CUDA_CHECK_LAUNCH();")
        L.append("  return 0;")
        L.append("}")
        return "\n".join(L)


def emit_scalar_stage(om: OM, ks: KernelSchedule) -> Optional[str]:
    """One-thread kernel for the Scalar-realm stores of an OM kernel (the reference's Scalar
    subkernels, PlanTrans.hs:417,566-567), run after the array stages."""
    if not ks.scalar_stores:
        return None
    name = f"om_{om.name}_{ks.name}_scalars"
    dummy = StageEmitter.__new__(StageEmitter)
    dummy.ops, dummy.ks = ks.ops, ks
    lines = StageEmitter.scalar_code(dummy, [v for (_s, v) in ks.scalar_stores])
    L = [f"__global__ void {name}_kernel(const OmGeom g, om_slot_t* __restrict__ sc) {{"]
    L.append("  if (threadIdx.x != 0 || blockIdx.x != 0) return;")
    L += ["  " + l for l in lines]
    for (s, v) in ks.scalar_stores:
        T = CPP_TYPE[ks.ops[v].ctype]
        L.append(f"  om_slot_store<{T}>(sc, {s}, s{v});")
    L.append("}")
    L.append(f'extern "C" int {name}(const OmGeom* g, void* sc, void* stream) {{')
    L.append(f"  OM_LAUNCH({name}_kernel, dim3(1, 1), 32, 0, (cudaStream_t)stream, *g, (om_slot_t*)sc);")
    L.append("  OM_CUDA_CHECK_LAUNCH();")
    L.append("  return 0;")
    L.append("}")
    return "\n".join(L)


def pick_vnt(om: OM, st: Stage, ks: KernelSchedule) -> Tuple[int, int]:
    """Cells per thread and threads per CTA.  16-byte vectors for 4-byte cells; wide double-
    precision DAGs (register-bound) use one cell per thread."""
    sv = om.setup.static_values
    types = [sv[i.static_idx].namee.type for i in st.inputs.values()] + [sv[s].namee.type for (s, _v) in st.store_targets]
    types += [ks.ops[v].ctype for (v, _o, _k) in st.reduce_targets]
    width = max([TYPE_BYTES[t] for t in types] + [4])
    if st.mats:
        return (1, 256)
    return (16 // width if width <= 8 else 1, 128)
