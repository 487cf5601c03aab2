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

import re
from typing import Dict, List, Optional, Tuple

import numpy as np

from ...annotation import CYCLIC, OPEN
from ...om.graph import ARRAY, CPP_TYPE, SCALAR, TYPE_BYTES, OM, imm_value
from ..plan import Plan
from .schedule import KernelSchedule, Op, Stage



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


APRON_ROWS = 32   # allocated (zero-initialised) rows the host must provide above and below every array
VEC_TYPE = {("int", 4): "int4", ("float", 4): "float4", ("double", 2): "double2", ("int", 2): "int2", ("float", 2): "float2"}


def _m(x: int) -> str:
    return str(x).replace("-", "m")


class StageEmitter:
    """Emits one fused stage.  Everything that does not depend on the row (column predicates, ghost
    edge tests, thread-uniform arithmetic) is hoisted out of the row loop, and ring slots are kept as
    loop-carried element offsets (one per (depth, row offset) pair) instead of being recomputed with
    a modulo per access."""

    def __init__(self, om: OM, plan: Plan, ks: KernelSchedule, st: Stage, stage_idx: int, V: int, NT: int):
        self.om, self.plan, self.ks, self.st, self.idx = om, plan, ks, st, stage_idx
        self.ops = ks.ops
        self.V, self.NT = V, NT
        self.fast = bool(getattr(plan.setup, "fast_math", False))
        self.tuning = plan.setup.tuning
        self.newton = (not self.fast) and self.tuning.exact_divsqrt == "newton"   # branch-free IEEE-correct Double division / sqrt
        self.uses_range_flag = False
        self._steady = False          # emitting the steady-state copy of a row-window body (no start / end tests)
        self._clean = False           # ... for a CTA without partial vectors or ghost copies (no rarely taken block)
        self.has_rare = False         # some row body has a rarely taken block
        self._ieee_scope = False      # emitting the cold clone of a scope: the compiler's own IEEE division / sqrt
        self._guarded_ops = 0         # guarded divisions / square roots emitted by the scope() call in progress
        self.PF = self.tuning.prefetch_rows     # cp.async prefetch distance in rows
        self.HL, self.HR = _ru(st.halo_x[0], V), _ru(st.halo_x[1], V)
        self.PL, self.PR = _ru(st.pad_x[0], V), _ru(st.pad_x[1], V)
        self.W_OUT = NT * V - self.HL - self.HR
        assert self.W_OUT > 0
        self.RW = NT * V + self.PL + self.PR
        self.name = f"om_{om.name}_{ks.name}_stage{stage_idx}"
        self.ring_inputs = [i for i in st.inputs.values() if i.via_smem]
        self.depth: Dict[int, int] = {}
        for i in self.ring_inputs:        # self.PF + 1 extra rows for the in-flight async copies
            self.depth[i.vid] = i.depth + self.PF + 1
        for m in st.mats.values():
            self.depth[m.vid] = m.depth
        self.static_of = {i.vid: i.static_idx for i in st.inputs.values()}
        self.margin_lo = tuple(plan.lower_margin) + (0,) * (3 - len(plan.lower_margin))
        # loadIndex on an axis generated Cyclic wraps with the reference's full modulo (om_wrap_far): index shifts compose past
        # the stencil radius and the grid may be narrower than the ghost width.  Axes generated Open keep the one-period form
        # behind their (never taken) run-time test.
        bnd = tuple(plan.setup.boundary) + (OPEN,) * (3 - len(plan.setup.boundary))
        self.wrap_fn = tuple("om_wrap_far" if b == CYCLIC else "om_wrap" for b in bnd)
        self.margin_hi = tuple(plan.upper_margin) + (0,) * (3 - len(plan.upper_margin))
        self.dim3 = plan.setup.dim == 3                    # rank 3: one plane of axis 2 per blockIdx.z (schedule.lower_z)
        self.zoff_of = {i.vid: i.zoff for i in st.inputs.values()}
        self.zdefs: Dict[str, str] = {}                     # plane-shifted base pointers, defined at the kernel top
        self.slotvars: Dict[Tuple[int, int], str] = {}     # (depth, row offset c) -> element-offset variable
        self.uniform_used: Dict[int, None] = {}
        self.pre: List[str] = []                            # hoisted, before the row loop
        self.uniform = self._uniform_nodes()
        self.window_u = None
        self.loop_top: List[str] = []                       # first statements of every row iteration
        self.direct_pf_tags: set = set()
        # TMA bulk staging: one cp.async.bulk per input row issued by thread 0, completion on an mbarrier; needs
        # 16-byte aligned row segments, i.e. 16-byte vectors per thread
        self.bulk = (self.tuning.staging == "bulk" and bool(self.ring_inputs) and
                     all(TYPE_BYTES[i.ctype] * V == 16 for i in self.ring_inputs))
        self.NBAR = self.PF + 2      # mbarriers in flight: one per staged iteration
        # row-window mode: MAT-free stages whose inputs are staged in rings keep the stencil window in registers
        self.window = (not st.mats) and bool(self.ring_inputs) and self.tuning.row_window
        if self.window:
            self.wcmin = min(i.lag - i.depth + 1 for i in self.ring_inputs)
            self.wcmax = max(i.lag for i in self.ring_inputs)
            self.U = self.wcmax - self.wcmin + 1
        # grouped staging (Tuning.barrier_group): the rows of one window rotation are staged together behind one CTA barrier;
        # the bodies of a group only read the row that enters the window, so 2 U ring rows suffice (readers of group g use rows
        # [j + lag, j + lag + U), the staging of group g + 1 behind the same barrier writes [j + lag + U, j + lag + 2 U))
        self.grouped = self.window and not self.bulk and bool(getattr(self.tuning, "barrier_group", False))
        if self.grouped:
            for i in self.ring_inputs:
                self.depth[i.vid] = 2 * self.U
        # warp-private rings (Tuning.warp_rings): every warp stages its own 32 V columns plus its own copy of the pads, so a row
        # that enters the window only has to be ordered among the lanes of one warp — __syncwarp instead of a CTA barrier per row
        self.warp_rings = (self.window and not self.bulk and not self.grouped and bool(getattr(self.tuning, "warp_rings", False))
                           and NT % 32 == 0 and NT > 32)
        if self.warp_rings:
            self.RWW = self.PL + 32 * V + self.PR      # one warp's segment of a ring row
            self.RW = (NT // 32) * self.RWW
        # inputs read without staging (column offset 0 only) are prefetched one row ahead into registers
        self.direct_pf = bool(self.tuning.direct_prefetch) and not self.window
        # rows touched outside [own_r0, own_r1): must stay inside the apron the ABI requires
        lags = [i.lag for i in st.inputs.values()] + [0]
        lows = [i.lag - i.depth + 1 for i in st.inputs.values()] + [0]
        reach = max(st.warmup + max(0, -min(lows)), max(lags) + self.PF) + 1
        if reach > APRON_ROWS:
            raise ValueError(f"stage needs {reach} apron rows, the ABI provides OM_APRON_ROWS = {APRON_ROWS}: lower "
                             "Tuning.prefetch_rows or raise Tuning.mat_threshold (fewer shared-memory rings, shorter warm-up)")

    # ------------------------------------------------------------------------------------------
    def T(self, v) -> str:
        return CPP_TYPE[self.ops[v].ctype]

    def inp(self, v: int) -> str:
        """Base pointer of the static array behind Load value v — for rank 3, of the plane that virtual input reads."""
        sidx = self.static_of[v]
        if not self.dim3:
            return f"in{sidx}"
        nm = f"inz{v}"
        z = self.zoff_of[v]
        # planes beyond the stack are only asked for by cells outside the Valid region (Open axis 2), whose results
        # are masked: clamp instead of reading outside the allocation
        zexpr = "zp" if z == 0 else f"min(max(zp + ({z}), 0), g.nzl + g.gz_lo + g.gz_hi - 1)"
        self.zdefs[nm] = f"const {self.T(v)}* __restrict__ {nm} = in{sidx} + (ptrdiff_t)({zexpr}) * g.plane;"
        return nm

    def outp(self, s: int, T: str) -> str:
        if not self.dim3:
            return f"out{s}"
        nm = f"outz{s}"
        self.zdefs[nm] = f"{T}* __restrict__ {nm} = out{s} + (ptrdiff_t)zp * g.plane;"
        return nm

    def is_light(self) -> bool:
        """Streaming stage: one phase, rings below 48 KB — launched in many short chunks (Tuning.chunk_rows_light); a heavy
        stage gets one full wave of equally long CTAs."""
        return not (len(self.st.phases) > 1 or self.smem_bytes() > 48 * 1024)

    def smem_bytes(self) -> int:
        tot = 0
        for v, d in self.depth.items():
            tot = _ru(tot, 16) + d * self.RW * TYPE_BYTES[self.ops[v].ctype]
        if getattr(self, "bulk", False):
            tot = _ru(tot, 16) + 8 * self.NBAR
        return _ru(tot, 16)

    def _uniform_nodes(self):
        """Array-realm values that are the same for every cell (built from Broadcast / Imm only)."""
        uni = set()
        for v in sorted(self.ops):
            op = self.ops[v]
            if op.realm != ARRAY:
                continue
            if op.kind in ("Imm", "Broadcast"):
                uni.add(v)
            elif op.kind == "Arith" and all((a in uni) or self.ops[a].realm == SCALAR for a in op.args):
                uni.add(v)
            elif op.kind == "Shift" and op.args[0] in uni:
                uni.add(v)
        return uni

    def slot_off(self, depth: int, c: int) -> str:
        if depth == 1:
            return "0"
        key = (depth, c % depth)
        if key not in self.slotvars:
            self.slotvars[key] = f"so{depth}_{key[1]}"
        return self.slotvars[key]

    # ---- scalar (uniform) values ------------------------------------------------------------
    def scalar_code(self, roots: List[int]) -> List[str]:
        """Evaluate Scalar-realm values from the device scalar table, in id order."""
        ops = self.ops
        seen = set()
        stack = list(roots)
        while stack:
            v = stack.pop()
            if v in seen:
                continue
            seen.add(v)
            op = ops[v]
            if op.kind == "Reduce":
                continue
            stack.extend(op.args)
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
                out.append(f"const {T} s{v} = ({T}){('g.nx', 'g.ny', 'g.nz')[op.inst.arg]};")
            elif op.kind == "Arith":
                out.append(f"const {T} s{v} = {self.arith(op, ['s%d' % a for a in op.args])};")
            else:
                raise NotImplementedError(f"scalar {op.kind}")
        return out

    def arith(self, op: Op, args: List[str]) -> str:
        """arith_expr plus one exact strength reduction: x / 2^k  ==  x * 2^-k in IEEE arithmetic."""
        if op.inst.arg == "Div" and op.ctype in ("Float", "Double"):
            d = self.ops[op.args[1]]
            if d.kind == "Imm":
                val = float(imm_value(d.inst.arg, d.ctype))
                if val != 0.0 and np.isfinite(val):
                    mant, _e = np.frexp(abs(val))
                    if mant == 0.5 and np.isfinite(1.0 / val) and abs(1.0 / val) > 1e-300:
                        return f"({args[0]} * {c_imm(1.0 / val, op.ctype)})"
        return arith_expr(op, args)

    def uniform_code(self) -> List[str]:
        ops = self.ops
        need = set()
        stack = list(self.uniform_used)
        while stack:
            v = stack.pop()
            if v in need:
                continue
            need.add(v)
            stack.extend(a for a in ops[v].args if ops[a].realm == ARRAY)
        out = []
        for v in sorted(need):
            op = ops[v]
            T = CPP_TYPE[op.ctype]
            nm = lambda a: f"s{a}" if ops[a].realm == SCALAR else f"u{a}"
            if op.kind == "Imm":
                out.append(f"const {T} u{v} = {c_imm(op.inst.arg, op.ctype)};")
            elif op.kind == "Broadcast":
                out.append(f"const {T} u{v} = s{op.args[0]};")
            elif op.kind == "Shift":
                out.append(f"const {T} u{v} = u{op.args[0]};")
            else:
                out.append(f"const {T} u{v} = {self.arith(op, [nm(a) for a in op.args])};")
        return out

    def staged_read(self, lines, ring_rd, lag, b, cur, k) -> str:
        """Read staged value `b` (input or MAT) at cursor `cur` for lane k: shared-memory ring, or — in
        row-window mode — the register copy of that ring row (each ring row is read from shared memory
        once, when it enters the stencil window, and reused by the following rows)."""
        V = self.V
        o = k + cur[0]
        if self.window_u is not None:
            return f"w{b}_{(self.window_u + cur[1] - self.wcmin) % self.U}_{_m(o)}"
        key = (b, cur[1], o)
        if key in ring_rd:
            return ring_rd[key]
        T = self.T(b)
        so = self.slot_off(self.depth[b], lag + cur[1])
        vt = VEC_TYPE.get((T, V))
        if 0 <= o < V and vt:
            vn = f"rv{b}_{_m(cur[1])}"
            lines.append(f"const {vt} {vn} = *reinterpret_cast<const {vt}*>(&ring{b}[{so} + tb]);")
            for kk in range(V):
                n2 = f"r{b}_{_m(cur[1])}_{kk}"
                lines.append(f"const {T} {n2} = {vn}.{'xyzw'[kk]};")
                ring_rd[(b, cur[1], kk)] = n2
            return ring_rd[key]
        nm = f"r{b}_{_m(cur[1])}_{_m(o)}"
        lines.append(f"const {T} {nm} = ring{b}[{so} + tb + ({o})];")
        ring_rd[key] = nm
        return nm

    # ---- array values inside one scope ------------------------------------------------------
    def scope(self, targets: List[int], lag: int) -> Tuple[List[str], Dict]:
        """SSA statements computing `targets` at cursor 0 for the V lanes of this thread at device row
        `row` (= j + lag).  Returns (lines, {(vid, lane): expr})."""
        ops, V = self.ops, self.V
        mats = self.st.mats
        lines: List[str] = []
        memo: Dict[tuple, str] = {}
        ring_rd: Dict[Tuple[int, int, int], str] = {}
        target_set = set(targets)

        def ring_read(b, cur, k):
            return self.staged_read(lines, ring_rd, lag, b, cur, k)

        def direct_read(v, cur, k):
            assert cur[0] == 0, "direct global reads are only scheduled for column offset 0"
            if self.direct_pf and not self._ieee_scope:      # (the cold IEEE clone re-loads instead of keeping the prefetched row alive)
                # software pipelining: the row this iteration needs was loaded one iteration ago (loop top), so the
                # HBM latency overlaps a whole row of arithmetic instead of stalling the few resident warps
                off = lag + cur[1]
                tag = f"{v}_{_m(off)}"
                if tag not in self.direct_pf_tags:
                    self.direct_pf_tags.add(tag)
                    T = self.T(v)
                    sidx = self.static_of[v]
                    vt = VEC_TYPE.get((T, V))
                    self.pre.append(f"const {T}* __restrict__ pn{tag} = {self.inp(v)} + (ptrdiff_t)(jbeg + ({off})) * g.pitch + tc;   // next row to prefetch")
                    if vt:
                        self.pre.append(f"{vt} nq{tag} = __ldg(reinterpret_cast<const {vt}*>(pn{tag})); pn{tag} += g.pitch;")
                        self.loop_top.append(f"const {vt} qp{tag} = nq{tag}; nq{tag} = __ldg(reinterpret_cast<const {vt}*>(pn{tag})); pn{tag} += g.pitch;")
                        for kk in range(V):
                            self.loop_top.append(f"const {T} dp{tag}_{kk} = qp{tag}.{'xyzw'[kk]};")
                    else:
                        for kk in range(V):
                            self.pre.append(f"{T} nd{tag}_{kk} = __ldg(pn{tag} + {kk});")
                            self.loop_top.append(f"const {T} dp{tag}_{kk} = nd{tag}_{kk}; nd{tag}_{kk} = __ldg(pn{tag} + g.pitch + {kk});")
                        self.loop_top.append(f"pn{tag} += g.pitch;")
                return f"dp{tag}_{k}"
            key = ("direct", v, cur[1])
            if key not in memo:
                memo[key] = "1"
                T = self.T(v)
                sidx = self.static_of[v]
                cy = cur[1]
                tag = f"{v}_{_m(cy)}"
                vt = VEC_TYPE.get((T, V))
                # no bounds predicates: the ABI requires OM_APRON_ROWS allocated rows around every array
                lines.append(f"const {T}* __restrict__ pd{tag} = {self.inp(v)} + (ptrdiff_t)(row + ({cy})) * g.pitch + tc;")
                if vt:
                    lines.append(f"const {vt} qd{tag} = __ldg(reinterpret_cast<const {vt}*>(pd{tag}));")
                    for kk in range(V):
                        lines.append(f"const {T} d{tag}_{kk} = qd{tag}.{'xyzw'[kk]};")
                else:
                    for kk in range(V):
                        lines.append(f"const {T} d{tag}_{kk} = __ldg(pd{tag} + {kk});")
            return f"d{v}_{_m(cur[1])}_{k}"

        def val(v, cur, k) -> str:
            key = (v, cur, k)
            if key in memo:
                return memo[key]
            op = ops[v]
            T = CPP_TYPE[op.ctype]
            if op.realm == SCALAR:
                memo[key] = f"s{v}"
                return memo[key]
            if v in self.uniform:
                self.uniform_used[v] = None
                memo[key] = f"u{v}"
                return memo[key]
            if v in mats and not (v in target_set and cur == (0, 0)):
                e = ring_read(v, cur, k)
                memo[key] = e
                return e
            if op.kind == "Load":
                inp = self.st.inputs[v]
                e = ring_read(v, cur, k) if inp.via_smem else direct_read(v, cur, k)
                memo[key] = e
                return e
            if op.kind == "StoredValue":      # carried reduce: the (masked) value this row stores, defined by emit_out
                memo[key] = f"o{op.args[0]}_{k}"
                return memo[key]
            if op.kind == "LoadIndex":
                ax = op.inst.arg
                nm = f"ix{ax}_{_cur(cur)}_{k}" + (f"_z{_m(op.zoff)}" if ax == 2 else "")
                if (nm, "def") not in memo:
                    memo[(nm, "def")] = "1"
                    if ax == 0:
                        hn = f"hix_{_m(cur[0])}_{k}"
                        if hn not in self.pre_names:
                            self.pre_names.add(hn)
                            self.pre.append(f"const int {hn} = g.cyc_x ? {self.wrap_fn[0]}(tc + {k} + ({cur[0]}) - g.xorg, g.nx) : (tc + {k} + ({cur[0]}) - g.xorg);")
                        lines.append(f"const int {nm} = {hn};")
                    elif ax == 1:
                        lines.append(f"const int {nm} = g.cyc_y ? {self.wrap_fn[1]}(row + ({cur[1]}) - g.yorg + g.y0, g.ny) : (row + ({cur[1]}) - g.yorg + g.y0);")
                    else:      # axis 2: the CTA's plane plus the offset lower_z gave this node
                        lines.append(f"const int {nm} = g.cyc_z ? {self.wrap_fn[2]}(zp + ({op.zoff}) - g.zorg + g.z0, g.nz) : (zp + ({op.zoff}) - g.zorg + g.z0);")
                e = f"(({T}){nm})"
            elif op.kind == "LoadSize":
                e = f"(({T}){('g.nx', 'g.ny', 'g.nz')[op.inst.arg]})"
            elif op.kind == "Shift":
                s = tuple(op.inst.arg) + (0,) * (2 - len(op.inst.arg))
                e = val(op.args[0], (cur[0] - s[0], cur[1] - s[1]), k)
                memo[key] = e
                return e
            elif op.kind == "Arith":
                args = [val(a, cur, k) for a in op.args]
                e = self.arith(op, args)
                if self.newton and op.ctype == "Double" and op.inst.arg in ("Max", "Min"):
                    e = f"{'om_fmax_std' if op.inst.arg == 'Max' else 'om_fmin_std'}({args[0]}, {args[1]})"   # the same compare + select, 3 instructions
                if (self.newton and not self._ieee_scope and op.ctype == "Double" and op.inst.arg in ("Div", "Inv", "Sqrt")
                        and e.count("*") == 0):
                    self.uses_range_flag = True
                    self._guarded_ops += 1
                    if op.inst.arg == "Sqrt":
                        e = f"om_sqrt_rn({args[0]}, om_slow)"
                    else:
                        den = args[-1]
                        rk = ("rcp", den)
                        if rk not in memo:     # one reciprocal refinement per distinct denominator (and cursor / lane)
                            memo[rk] = f"rc_{len([1 for q in memo if isinstance(q, tuple) and q and q[0] == 'rcp'])}_{k}"
                            lines.append(f"const double {memo[rk]} = om_rcp_rn_seq({den}, om_slow);")
                        e = f"om_div_rn({'1.0' if op.inst.arg == 'Inv' else args[0]}, {den}, {memo[rk]}, om_slow)"
                if self.fast and op.ctype == "Double" and op.inst.arg in ("Max", "Min"):
                    e = f"{'om_fmax_std' if op.inst.arg == 'Max' else 'om_fmin_std'}({args[0]}, {args[1]})"   # same result, 3 instructions
                if self.fast and op.ctype == "Double" and op.inst.arg in ("Div", "Inv", "Sqrt") and e.count("*") == 0:
                    if op.inst.arg == "Sqrt":
                        e = f"om_fsqrt({args[0]})"
                    else:
                        den = args[-1]
                        rk = ("rcp", den)
                        if rk not in memo:     # one reciprocal per distinct denominator (and cursor / lane)
                            memo[rk] = f"rc_{len([1 for q in memo if isinstance(q, tuple) and q and q[0] == 'rcp'])}_{k}"
                            lines.append(f"const double {memo[rk]} = om_frcp({den});")
                        e = memo[rk] if op.inst.arg == "Inv" else f"om_fdiv_r({args[0]}, {den}, {memo[rk]})"
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
        return lines, result

    def scope_guarded(self, targets: List[int], lag: int, tag: str = "") -> Tuple[List[str], Dict]:
        """scope() for the bit-exact build with the branch-free division / sqrt (Tuning.exact_divsqrt = "newton"): the scope's
        code once with om_div_rn / om_sqrt_rn, which only raise the cell's `om_slow` flag on a tiny non-zero operand, and —
        behind ONE branch — a cold clone of the same code with the compiler's IEEE expansions (slow paths included) that
        re-evaluates the cell when the flag is up.  The scope's results are hoisted into variables both versions assign."""
        self._guarded_ops = 0
        lines, res = self.scope(targets, lag)
        if not (self.newton and self._guarded_ops):
            return lines, res
        if self.tuning.exact_guard != "redo":      # measurement variants: the flag only, or no guard at all (dead code to the compiler)
            return (["bool om_slow = false;"] + lines + (["om_bad |= om_slow ? 1u : 0u;"] if self.tuning.exact_guard == "flag" else [])), res
        self._ieee_scope = True
        try:
            slow_lines, slow_res = self.scope(targets, lag)
        finally:
            self._ieee_scope = False
        names = {key: f"g{tag}{key[0]}_{key[1]}" for key in res}
        out = [f"{self.T(key[0])} {nm};" for key, nm in names.items()]
        out.append("bool om_slow = false;   // a tiny non-zero operand of a division / square root: outside the branch-free sequences' range")
        out.append("{")
        out += ["  " + l for l in lines]
        out += [f"  {names[key]} = {e};" for key, e in res.items()]
        out.append("}")
        # the cold clone lives in a function of its own (a noinline closure that captures what it reads by value), so that it
        # costs the hot path neither registers nor scheduling freedom: inlined, it took the flux kernel from 10.0 to 7.4 Gcell/s
        out.append("if (om_slow) {   // (cold) this cell again with the compiler's IEEE division / sqrt: every stored bit is IEEE's either way")
        out.append("  struct OmRedo { " + " ".join(f"{self.T(key[0])} v{n};" for n, key in enumerate(res)) + " };")
        out.append("  auto om_redo = [=]() __attribute__((noinline)) -> OmRedo {")
        out += ["    " + l for l in slow_lines]
        out.append("    OmRedo r;")
        out += [f"    r.v{n} = {slow_res[key]};" for n, key in enumerate(res)]
        out.append("    return r;")
        out.append("  };")
        out.append("  const OmRedo om_r = om_redo();")
        out += [f"  {names[key]} = om_r.v{n};" for n, key in enumerate(res)]
        out.append("  atomicAdd(&red_counter[OM_SIG_SLOW], 1u);")
        out.append("}")
        return out, names

    # ---- whole kernel ---------------------------------------------------------------------------
    def kernel(self) -> str:
        st, V, NT = self.st, self.V, self.NT
        om = self.om
        self.pre_names = set()
        in_statics = sorted({i.static_idx for i in st.inputs.values()})
        out_statics = list(dict.fromkeys(s for (s, _v) in st.store_targets))
        sv = om.setup.static_values
        params = ["const OmGeom g"]
        for s in in_statics:
            params.append(f"const {CPP_TYPE[sv[s].namee.type]}* __restrict__ in{s}")
        for s in out_statics:
            params.append(f"{CPP_TYPE[sv[s].namee.type]}* __restrict__ out{s}")
        params += ["om_slot_t* __restrict__ sc", "unsigned* __restrict__ red_counter", "om_slot_t* __restrict__ red_partials"]
        mlx, mhx = self.margin_lo[0], self.margin_hi[0]
        warm = st.warmup
        has_ring_in = bool(self.ring_inputs)
        lead = warm + (self.PF if (has_ring_in and not self.grouped) else 0)
        self.post_init: List[str] = []      # statements between the declarations and the row loop (grouped staging: the first group)

        # ---- loop body first (it registers slot counters, hoisted values, uniform nodes) -------
        def stage_inputs(B, row_shift: int):
            if self.bulk:
                B.append("// stage the next input rows: one TMA bulk copy (cp.async.bulk / UBLKCP) per row, issued by thread 0,")
                B.append("// completing on this iteration's mbarrier; the consumer side waits for the barrier armed PF iterations ago")
                tot = sum(self.RW * TYPE_BYTES[i.ctype] for i in self.ring_inputs)
                B.append("if (tid == 0) {")
                B.append(f"  om_mbar_expect_tx(&mbar[bar_i], {tot});")
                for i in self.ring_inputs:
                    v = i.vid
                    T = self.T(v)
                    so = self.slot_off(self.depth[v], i.lag + self.PF + row_shift)
                    ln = f"const {T}* __restrict__ src{v} = {self.inp(v)} + (ptrdiff_t)(jbeg + {i.lag + self.PF}) * g.pitch + strip_lo - HL - PL;   // advances one row per staged row"
                    if ln not in self.pre:
                        self.pre.append(ln)
                    B.append(f"  om_bulk_g2s(&ring{v}[{so}], src{v}, {self.RW * TYPE_BYTES[i.ctype]}, &mbar[bar_i]);")
                    B.append(f"  src{v} += g.pitch;")
                B.append("}")
                B.append(f"if (it >= {self.PF}) om_mbar_wait(&mbar[bar_w], bar_wp);")
                return
            B.append("// stage the next input rows (LDGSTS); the apron rows around every array make bounds checks unnecessary")
            # (rows beyond the chunk's last needed row are not fetched: without the test every chunk would read
            #  prefetch_rows rows it never uses — 6 % more DRAM reads for 32-row chunks)
            # (short chunks only, i.e. row-window stages: in the heavy stages' one-wave chunks of hundreds of rows the test
            #  saves nothing and costs the flux kernel 4 %)
            jx = "j" if self.window_u is None else f"j + {self.window_u}"
            B.append(f"if ({jx} + {self.PF} < r1) {{" if (self.window and not self._steady) else "{")
            for i in self.ring_inputs:
                v = i.vid
                T = self.T(v)
                nb = TYPE_BYTES[i.ctype] * V
                so = self.slot_off(self.depth[v], i.lag + self.PF + row_shift)
                ln = f"const {T}* __restrict__ src{v} = {self.inp(v)} + (ptrdiff_t)(jbeg + {i.lag + self.PF}) * g.pitch + tc;   // advances one row per staged row"
                if ln not in self.pre:
                    self.pre.append(ln)
                B.append(f"om_cp_async<{nb}>(&ring{v}[{so} + tb], src{v}, {nb});")
                if self.warp_rings:      # the pads of this warp's own segment
                    if self.PL:
                        B.append(f"if (lane < {self.PL // V}) om_cp_async<{nb}>(&ring{v}[{so} + wb + lane * V], src{v} - PL, {nb});")
                    if self.PR:
                        B.append(f"if (lane < {self.PR // V}) om_cp_async<{nb}>(&ring{v}[{so} + wb + PL + 32 * V + lane * V], src{v} + 32 * V, {nb});")
                else:
                    if self.PL:
                        B.append(f"if (tid < {self.PL // V}) om_cp_async<{nb}>(&ring{v}[{so} + tid * V], src{v} - PL, {nb});")
                    if self.PR:
                        B.append(f"if (tid < {self.PR // V}) om_cp_async<{nb}>(&ring{v}[{so} + PL + NT * V + tid * V], src{v} + NT * V, {nb});")
                B.append(f"src{v} += g.pitch;")
            B.append("}")
            B.append("om_cp_async_commit();")
            B.append(f"om_cp_async_wait<{self.PF}>();")

        nph = max(len(st.phases), st.out_level)
        bodies: List[List[str]] = []
        def stage_row(out, c: int, guard: str, ind: str):
            """cp.async of one row of every ring input into the slot of row (group base + c)."""
            out.append(f"{ind}if ({guard}) {{")
            for i in self.ring_inputs:
                v, T = i.vid, self.T(i.vid)
                nb = TYPE_BYTES[i.ctype] * V
                so = self.slot_off(self.depth[v], i.lag + c)
                ln = f"const {T}* __restrict__ src{v} = {self.inp(v)} + (ptrdiff_t)(jbeg + {i.lag}) * g.pitch + tc;   // advances one row per staged row"
                if ln not in self.pre:
                    self.pre.append(ln)
                out.append(f"{ind}  om_cp_async<{nb}>(&ring{v}[{so} + tb], src{v}, {nb});")
                if self.PL:
                    out.append(f"{ind}  if (tid < {self.PL // V}) om_cp_async<{nb}>(&ring{v}[{so} + tid * V], src{v} - PL, {nb});")
                if self.PR:
                    out.append(f"{ind}  if (tid < {self.PR // V}) om_cp_async<{nb}>(&ring{v}[{so} + PL + NT * V + tid * V], src{v} + NT * V, {nb});")
                out.append(f"{ind}  src{v} += g.pitch;")
            out.append(f"{ind}}}")

        group_top: List[str] = []
        if self.window and self.grouped:
            # the first group is staged before the loop; every group then waits for its own rows, synchronises the CTA once and
            # stages the next group (rows beyond the chunk's last needed row are not fetched)
            for u in range(self.U):
                stage_row(self.post_init, u, f"jbeg + {u} < r1", "")
            self.post_init.append("om_cp_async_commit();")
            group_top.append("om_cp_async_wait<0>();")
            group_top.append("__syncthreads();")
            for u in range(self.U):
                stage_row(group_top, self.U + u, f"j + {self.U + u} < r1", "")
            group_top.append("om_cp_async_commit();")
        steady_bodies: List[List[str]] = []
        clean_bodies: List[List[str]] = []
        if self.window:
            # one unrolled body per window row: the register sets rotate by renaming
            wregs: Dict[str, str] = {}
            peel = self.tuning.peel_fill and not self.grouped
            for u in list(range(self.U)) * ((3 if self.tuning.clean_ctas else 2) if peel else 1):
                # (second pass: the steady-state copies — every row of the group is stored and every staged row is needed, so the
                #  per-row tests go and a group is tested once, CTA-uniformly; third pass: the same for a CTA none of whose threads
                #  ever takes the rarely taken block of a row — it is not emitted at all)
                self._steady = len(bodies) == self.U
                self._clean = self._steady and len(steady_bodies) == self.U
                self.window_u = u
                B: List[str] = []
                if not self.grouped:
                    stage_inputs(B, 0)
                    B.append("__syncwarp();   // the row's segment was staged by this warp's own lanes" if self.warp_rings else "__syncthreads();")
                B.append("// the row that enters the stencil window: shared memory -> registers, once")
                for i in self.ring_inputs:
                    b, T = i.vid, self.T(i.vid)
                    so = self.slot_off(self.depth[b], i.lag + (u if self.grouped else 0))
                    sset = (u + i.lag - self.wcmin) % self.U
                    vt = VEC_TYPE.get((T, V))
                    names = [f"w{b}_{sset}_{_m(o)}" for o in range(-i.rd_xlo, V + i.rd_xhi)]
                    for n_ in names:
                        wregs[n_] = T
                    if vt:
                        B.append(f"{{ const {vt} q = *reinterpret_cast<const {vt}*>(&ring{b}[{so} + tb]); " +
                                 " ".join(f"w{b}_{sset}_{kk} = q.{'xyzw'[kk]};" for kk in range(V)) + " }")
                    else:
                        for kk in range(V):
                            B.append(f"w{b}_{sset}_{kk} = ring{b}[{so} + tb + {kk}];")
                    for o in list(range(-i.rd_xlo, 0)) + list(range(V, V + i.rd_xhi)):
                        B.append(f"w{b}_{sset}_{_m(o)} = ring{b}[{so} + tb + ({o})];")
                B += self.emit_out(row_expr=f"j + {u}", guard="true" if self._steady else f"j + {u} >= r0")
                (clean_bodies if self._clean else steady_bodies if self._steady else bodies).append(B)
            self._steady = self._clean = False
            self.window_u = None
            self.wregs = wregs
        else:
            def build(mode: str) -> List[str]:
                """One row iteration.  "steady" (j >= r0, all but the first rows of a chunk): every scope runs and the row is
                stored — no tests, so the code between two barriers is one basic block the compiler can schedule across the
                scopes; "fill": the first rows of a chunk, each scope behind its own start test, nothing stored yet;
                "both": one body for all rows (every scope and the stores behind their tests)."""
                steady = mode == "steady"
                B: List[str] = []
                if has_ring_in:
                    stage_inputs(B, 0)
                B.append("__syncthreads();")
                for lvl in range(1, nph + 1):
                    if lvl > 1:
                        B.append("__syncthreads();")
                    mats_here = st.phases[lvl - 1] if lvl - 1 < len(st.phases) else []
                    lags = sorted({st.mats[m].lag for m in mats_here}, key=lambda a: min(m for m in mats_here if st.mats[m].lag == a))
                    for a in lags:
                        grp = [m for m in mats_here if st.mats[m].lag == a]
                        early = min(st.mats[m].early for m in grp)
                        head = "{" if steady else f"if (j >= r0 - {-early}) {{"
                        B.append(f"{head}   // phase {lvl}: row j+{a} of {len(grp)} intermediate(s)")
                        B.append(f"  const int row = j + {a};")
                        lines, res = self.scope_guarded(grp, a)
                        B += ["  " + l for l in lines]
                        for m in grp:
                            so = self.slot_off(self.depth[m], a)
                            T = self.T(m)
                            vt = VEC_TYPE.get((T, V))
                            if vt:
                                B.append(f"  *reinterpret_cast<{vt}*>(&ring{m}[{so} + tb]) = make_{vt}({', '.join(res[(m, k)] for k in range(V))});")
                            else:
                                for k in range(V):
                                    B.append(f"  ring{m}[{so} + tb + {k}] = {res[(m, k)]};")
                        B.append("}")
                    if lvl == st.out_level and mode != "fill":
                        B += self.emit_out(guard="true" if steady else "j >= r0")
                return B
            fill_body = None
            if st.mats and self.tuning.peel_fill:
                # two row loops: the first rows of a chunk (pipeline fill), then the steady rows — instead of a start test per scope
                steady = build("steady")
                fill_body = self.loop_top + build("fill")
                bodies.append(self.loop_top + steady)
            else:
                bodies.append(self.loop_top + build("both"))

        # ---- assemble -----------------------------------------------------------------------------
        L: List[str] = []
        E = L.append
        E(f"// stage {self.idx} of kernel `{self.ks.name}` (reduce level {st.level}): "
          f"{len(st.mats)} shared-memory rings for intermediates, {len(self.ring_inputs)} for inputs, "
          f"{nph} phase(s), warm-up {st.warmup} rows, {V} cell(s) per thread")
        minb = self.tuning.min_blocks if not st.mats else self.tuning.min_blocks_heavy
        lb = f"__launch_bounds__({NT}, {minb})" if minb else f"__launch_bounds__({NT})"
        E(f"__global__ void {lb} {self.name}_kernel({', '.join(params)}) {{")
        E(f"  constexpr int V = {V}, NT = {NT}, HL = {self.HL}, PL = {self.PL}, RW = {self.RW}, W_OUT = {self.W_OUT};")
        E("  const int tid = threadIdx.x;")
        if self.warp_rings:
            E(f"  constexpr int RWW = {self.RWW};                          // one warp's segment of a ring row: its 32 V columns between its own pads")
            E("  const int lane = tid & 31, wb = (tid >> 5) * RWW;")
            E("  const int tb = wb + PL + lane * V;                     // this thread's element offset inside a ring row")
        else:
            E("  const int tb = PL + tid * V;                           // this thread's element offset inside a ring row")
        E("  OM_DYNAMIC_SMEM(om_smem);")
        off = 0
        for v, d in self.depth.items():
            off = _ru(off, 16)
            E(f"  {self.T(v)}* const ring{v} = reinterpret_cast<{self.T(v)}*>(om_smem + {off});  // {d} rows")
            off += d * self.RW * TYPE_BYTES[self.ops[v].ctype]
        if self.bulk:
            off = _ru(off, 16)
            E(f"  uint64_t* const mbar = reinterpret_cast<uint64_t*>(om_smem + {off});   // {self.NBAR} mbarriers")
            E(f"  if (tid == 0) {{ for (int i = 0; i < {self.NBAR}; ++i) om_mbar_init(&mbar[i], 1); om_mbar_init_fence(); }}")
            E("  __syncthreads();")
            E(f"  int it = 0, bar_i = 0, bar_w = {self.NBAR - self.PF}, bar_wp = 1;   // iteration, barrier armed now, barrier awaited now + its parity")
        E(f"  const int cx0 = g.xorg - {mlx}, cx1 = g.xorg + g.nx + {mhx};   // columns of the reference memory box")
        E("  const int cA = (cx0 / V) * V;")
        E("  const int strip_lo = cA + blockIdx.x * W_OUT;            // first output column of this CTA")
        E("  const int tc = strip_lo - HL + tid * V;                  // first column of this thread")
        E("  // several ranks (g.bfirst): the chunk with the slab's top rows runs first, then chunks 0, 1, ... — both chunks a neighbour")
        E("  // reads from are in the first wave and signal the host's communication stream once their rows are stored")
        # light (streaming) stages of rank-1 / rank-2 machines understand the boundary-first chunk order; heavy stages fill the
        # GPU with one wave of long chunks, so there is nothing to reorder (and their register budget is tight)
        self.bfirst = self.is_light() and not self.dim3
        cy = "(g.bfirst ? (blockIdx.y == 0 ? (int)gridDim.y - 1 : (int)blockIdx.y - 1) : (int)blockIdx.y)" if self.bfirst else "(int)blockIdx.y"
        E("  // chunk c of gridDim.y covers rows own_r0 + [c, c + 1) * nrows / gridDim.y: heights differ by at most one row, whatever the count")
        E("  // (32-bit unsigned arithmetic: the host keeps (chunks + 1) * rows below 2^32)")
        E(f"  const unsigned cy_rows = (unsigned){cy} * (unsigned)(g.own_r1 - g.own_r0);")
        E("  const int r0 = g.own_r0 + (int)(cy_rows / gridDim.y);")
        E("  const int r1 = g.own_r0 + (int)((cy_rows + (unsigned)(g.own_r1 - g.own_r0)) / gridDim.y);")
        E(f"  const int jbeg = r0 - {lead};")
        if self.dim3:
            E(f"  const int zp = g.own_z0 + blockIdx.z * {st.zplanes};   // this CTA's (first) plane of axis 2 (device plane index)")
            for l in self.zdefs.values():
                E("  " + l)
        for l in self.scalar_code(list(dict.fromkeys(st.scalar_roots))):
            E("  " + l)
        for l in self.uniform_code():
            E("  " + l)
        for l in self.pre:
            E("  " + l)
        for (v, rop, slot) in self.reduce_slots_once():
            T = self.T(v)
            ident = {"Sum": f"({T})0", "Min": self.type_max(v), "Max": self.type_min(v)}[rop]
            E(f"  {T} acc{slot} = {ident};   // reduce slot {slot}")
        if self.uses_range_flag:
            E("  unsigned om_bad = 0u;   // sticky: this thread stored a NaN / Inf (om_div_rn / om_sqrt_rn, om_runtime.cuh)")
            if self.depth:
                # While a chunk's pipeline fills, scopes read ring rows nobody has written yet.  Nothing computed from them reaches a
                # stored cell, but whatever the previous kernel left in shared memory is then an operand: after an NCCL kernel that is
                # integer data, i.e. denormals as doubles, and every such lane sent its cell to the IEEE slow path (10 000 cells per
                # launch on two GPUs, 2x the kernel time).  Zeros are benign operands.
                E(f"  for (int i = tid; i < {self.smem_bytes() // 16}; i += NT) reinterpret_cast<int4*>(om_smem)[i] = make_int4(0, 0, 0, 0);")
                E("  __syncthreads();")
        for (d, c), nm in sorted(self.slotvars.items()):
            E(f"  int {nm} = ((((jbeg + {c}) % {d}) + {d}) % {d}) * RW;")
        if self.window:
            byT: Dict[str, List[str]] = {}
            for n_, T in self.wregs.items():
                byT.setdefault(T, []).append(n_)
            for T, ns in byT.items():
                E(f"  {T} " + ", ".join(f"{n_} = 0" for n_ in sorted(ns)) + ";   // stencil window (rotates by renaming)")
        for l in self.post_init:
            E("  " + l)
        if clean_bodies and self.has_rare:
            E("  const bool cta_rare = __syncthreads_or(rare);   // CTA-uniform: an edge strip, a chunk with a y wrap, a strip with partial vectors")
        nb_ = len(bodies)
        if not self.window and fill_body is not None:
            E("  int j = jbeg;")
            E("  for (; j < min(r0, r1); ++j) {   // pipeline fill: every scope behind its own start test, nothing stored")
            L += ["    " + l for l in fill_body]
            for (d, c), nm in sorted(self.slotvars.items()):
                E(f"    {nm} += RW; if ({nm} == {d} * RW) {nm} = 0;")
            E("  }")
            E("  for (; j < r1; ++j) {   // steady rows: one basic block between barriers")
        else:
            E(f"  for (int j = jbeg; j < r1; j += {nb_}) {{")
        L += ["    " + l for l in group_top]

        def emit_bodies(bs, tested: bool, ind0: str):
            for u, B in enumerate(bs):
                wrap = nb_ > 1 and tested
                if wrap:
                    E(f"{ind0}if (j + {u} < r1) {{")
                ind = ind0 + ("  " if wrap else "")
                L.extend(ind + l for l in B)
                # slot offsets are relative to the body's own row, so they advance after every body (grouped staging: relative to the
                # group's first row; they advance by U rows after the group)
                for (d, c), nm in sorted(self.slotvars.items()):
                    if not self.grouped:
                        E(ind + f"{nm} += RW; if ({nm} == {d} * RW) {nm} = 0;")
                if self.bulk:
                    E(ind + f"++it; if (++bar_i == {self.NBAR}) bar_i = 0; if (++bar_w == {self.NBAR}) {{ bar_w = 0; bar_wp ^= 1; }}")
                if wrap:
                    E(f"{ind0}}}")
        if steady_bodies:
            E(f"    if (j >= r0 && j + {self.U - 1 + self.PF} < r1) {{   // steady group: every row is stored, every staged row is needed")
            if clean_bodies and self.has_rare:
                E("      if (!cta_rare) {   // no thread of this CTA has a partial vector or a ghost copy: the row bodies without that block")
                emit_bodies(clean_bodies, False, "        ")
                E("      } else {")
                emit_bodies(steady_bodies, False, "        ")
                E("      }")
            else:
                emit_bodies(steady_bodies, False, "      ")
            E("    } else {   // the first and the last rows of a chunk: each row behind its tests")
            emit_bodies(bodies, True, "      ")
            E("    }")
        else:
            emit_bodies(bodies, True, "    ")
        if self.grouped:
            for (d, c), nm in sorted(self.slotvars.items()):
                E(f"    {nm} += {self.U} * RW; if ({nm} >= {d} * RW) {nm} -= {d} * RW;")
        E("  }")
        if self.uses_range_flag:
            E("  if (om_bad) red_counter[OM_SIG_RANGE] = 1u;   // the host raises at its next synchronisation point")
        if self.bfirst:
            E("  // (block indices are re-read here instead of being kept in registers across the row loop)")
            E("  if (g.bfirst && ((g.sig_lo && blockIdx.y == 1u % gridDim.y) || (g.sig_hi && blockIdx.y == 0u))) {   // this CTA's rows are what a neighbour waits for")
            E("    __syncthreads();")
            E("    om_signal_boundary(red_counter, gridDim.x * ((g.sig_lo ? 1u : 0u) + ((g.sig_hi && (gridDim.y > 1 || !g.sig_lo)) ? 1u : 0u)));")
            E("  }")
        L += self.emit_reduce_epilogue()
        E("}")
        return "\n".join(L)

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

    def valid_box_z(self, v) -> Tuple[int, int]:
        """(lb_z, ub_z) of the node's Valid region along axis 2 (rank-3 machines)."""
        from ..plan import _valid_to_lower, _valid_to_upper
        if not self.dim3:
            return 0, 0
        valid = self.ops[v].valid
        return _valid_to_lower(self.plan.setup, valid)[2], _valid_to_upper(self.plan.setup, valid)[2]

    def emit_out(self, row_expr: str = "j", guard: str = "j >= r0") -> List[str]:
        """Stores + reduce accumulation for one output row.  The common case (all V lanes of the thread
        inside the strip's output range, no ghost cell to mirror) is a single predicated vector store;
        per-lane predicates are only evaluated on the rare partial-vector path."""
        st, V = self.st, self.V
        B: List[str] = []
        Z = st.zplanes
        splane = lambda s_, v_: st.store_plane.get((s_, v_), 0)
        rplane = lambda n: st.reduce_plane[n] if n < len(st.reduce_plane) else 0
        # (value, plane offset inside the CTA's group of Z planes) of everything this row stores or reduces
        entries = list(dict.fromkeys([(v, splane(s_, v)) for (s_, v) in st.store_targets] +
                                     [(v, rplane(n)) for n, (v, _o, _k) in enumerate(st.reduce_targets)]))
        targets = list(dict.fromkeys(v for (v, _zo) in entries))
        oname = (lambda v_, zo_, k_: f"o{v_}_{k_}") if Z == 1 else (lambda v_, zo_, k_: f"o{v_}z{zo_}_{k_}")
        zlive = lambda zo_: None if zo_ == 0 else f"zlive{zo_}"
        mly, mhy = self.margin_lo[1], self.margin_hi[1]

        def P(line):
            if line not in self.pre:
                self.pre.append(line)
        P("const int out_lo = max(cx0, strip_lo), out_hi = min(cx1, strip_lo + W_OUT);   // this CTA's output columns")
        P("const bool li_all = (tc >= out_lo) && (tc + V <= out_hi);                     // whole vector inside")
        P("const bool li_any = (tc + V > out_lo) && (tc < out_hi);")
        # ghost-cell writes exist only in machines generated with a Cyclic axis (the boundary kinds are baked into the kernels:
        # margins, Valid masks); Open / Open machines (Hydro) carry neither the predicates nor the branch per store
        cyclic = any(b == CYCLIC for b in self.plan.setup.boundary)
        if cyclic:
            P("const bool edge_x = g.cyc_x && (strip_lo < g.xorg + g.gx_hi || strip_lo + W_OUT > g.xorg + g.nx - g.gx_lo);")
            P("const bool edge_y = g.wrap_y_local && (r0 < g.yorg + g.gy_hi || r1 > g.yorg + g.nyl - g.gy_lo);")
            P("const bool edge_any = edge_x || edge_y;   // this CTA writes cells that have a ghost copy")
        # vector stages of rank-1 / rank-2 machines: the full-vector store and accumulate are predicated on li_all; partial vectors at the
        # strip's edge and cells with a ghost copy go through ONE rarely taken block per row instead of a branch per store and reduce
        compact = Z == 1 and V > 1 and all(VEC_TYPE.get((self.T(v), V)) for (_s, v) in st.store_targets)
        if compact:
            if cyclic:
                # (in a strip at the x edge only the vectors that hold a column with a ghost copy take the block — one warp of the CTA,
                #  not all of them: 2 of Life 16384^2's 32 strips are edge strips)
                P("const bool ghost_x = edge_x && li_any && (tc - g.xorg < g.gx_hi || tc + V - g.xorg > g.nx - g.gx_lo);")
                P("const bool ghost_only = ghost_x && li_all && !edge_y;   // ... and nothing else to do: the lean form of the block")
            P("const bool rare = (li_any && !li_all)" + (" || edge_y || ghost_x" if cyclic else "") + ";   // this thread has more to do than one full-vector store per row")
        rare_lines: List[str] = []
        lean_lines: List[str] = []     # the same block for a whole vector inside the strip whose only extra work is the x ghost copy
        B.append(f"if ({guard}) {{   // OUT: stores and reduce accumulation for one row")
        B.append(f"  const int row = {row_expr};")
        lines, res = self.scope_guarded(targets, 0)
        B += ["  " + l for l in lines]
        need_gmy = False
        for zo in sorted({zo_ for (_v, zo_) in entries if zo_ > 0}):
            P(f"const bool zlive{zo} = zp + {zo} < g.own_z1;   // the last group of planes may be incomplete")
        for (v, zo) in entries:
            lbx, ubx, lby, uby = self.valid_box(v)
            T = self.T(v)
            conds_row = []
            if lby or uby:
                need_gmy = True
                conds_row.append(f"(gmy >= {lby}) && (gmy < memy - {uby})")
            lbz, ubz = self.valid_box_z(v)
            if lbz or ubz:
                P(f"const int gmz = zp - g.zorg + g.z0 + {self.margin_lo[2]}, memz = g.nz + {self.margin_lo[2] + self.margin_hi[2]};   // plane in the reference memory box")
                conds_row.append(f"(gmz + {zo} >= {lbz}) && (gmz + {zo} < memz - {ubz})")
            for k in range(V):
                conds = list(conds_row)
                if lbx or ubx:
                    hn = f"vx{v}_{k}"
                    P(f"const bool {hn} = (tc + {k} >= cx0 + {lbx}) && (tc + {k} < cx1 - {ubx});")
                    conds.append(hn)
                if conds:   # cells of the memory box outside the Valid region are never written by the reference: they stay 0
                    B.append(f"  const {T} {oname(v, zo, k)} = ({' && '.join(conds)}) ? {res[(v, k)]} : ({T})0;")
                else:
                    B.append(f"  const {T} {oname(v, zo, k)} = {res[(v, k)]};")
        if self.uses_range_flag:      # the state guard of the bit-exact build: only cells this thread really stores count
            for (s_, v_) in st.store_targets:
                if self.ops[v_].ctype == "Double":
                    B.append("  if (li_any) om_bad |= " + " | ".join(
                        (f"om_state_bad({oname(v_, splane(s_, v_), k)})" if V == 1 else
                         f"((tc + {k} >= out_lo && tc + {k} < out_hi) ? om_state_bad({oname(v_, splane(s_, v_), k)}) : 0u)") for k in range(V)) + ";")
        if need_gmy:
            idx = B.index(f"  const int row = {row_expr};")
            B.insert(idx + 1, f"  const int gmy = row - g.yorg + g.y0 + {mly}; const int memy = g.ny + {mly + mhy};   // row in the reference memory box")
        for (s, v) in st.store_targets:
            T = self.T(v)
            vt = VEC_TYPE.get((T, V))
            zo = splane(s, v)
            ps = f"po{s}" if zo == 0 else f"po{s}z{zo}"
            on = [oname(v, zo, k) for k in range(V)]
            P(f"{T}* __restrict__ {ps} = {self.outp(s, T)} + (ptrdiff_t)r0 * g.pitch + tc" + (f" + (ptrdiff_t){zo} * g.plane" if zo else "") +
              ";   // advances one row per output row")
            dst = B
            if compact:
                mk = f"make_{vt}({', '.join(on)})"
                st_fn = {"cs": "__stcs(reinterpret_cast<%s*>(p), %s)", "cg": "__stcg(reinterpret_cast<%s*>(p), %s)"}.get(
                    self.tuning.store_hint, "*reinterpret_cast<%s*>(p) = %s") % (vt, mk)
                B.append(f"  {{ {T}* __restrict__ p = {ps}; {ps} += g.pitch; if (li_all) {{ {st_fn}; }} }}")
                dst = rare_lines
                dst.append(f"  {{ {T}* __restrict__ p = {ps} - g.pitch;")
                if cyclic:
                    lean_lines.append(f"  {{ {T}* __restrict__ p = {ps} - g.pitch;")
                    for k in range(V):
                        lean_lines.append(f"    {{ const int c = tc + {k} - g.xorg; if (c < g.gx_hi) p[{k} + g.nx] = {on[k]}; if (c >= g.nx - g.gx_lo) p[{k} - g.nx] = {on[k]}; }}")
                    lean_lines.append("  }")
                dst.append("    if (!li_all) {")
                for k in range(V):
                    dst.append(f"      if (tc + {k} >= out_lo && tc + {k} < out_hi) p[{k}] = {on[k]};")
                dst.append("    }")
            else:
                B.append(f"  {{ {T}* __restrict__ p = {ps}; {ps} += g.pitch;" + (f" if ({zlive(zo)}) {{" if zo else ""))
            if compact:
                pass
            elif vt:
                mk = f"make_{vt}({', '.join(on)})"
                if self.tuning.store_hint == "cs":      # streaming store (evict-first): the row is not read again before the next step
                    B.append(f"    if (li_all) {{ __stcs(reinterpret_cast<{vt}*>(p), {mk}); }}")
                elif self.tuning.store_hint == "cg":
                    B.append(f"    if (li_all) {{ __stcg(reinterpret_cast<{vt}*>(p), {mk}); }}")
                else:
                    B.append(f"    if (li_all) {{ *reinterpret_cast<{vt}*>(p) = {mk}; }}")
                B.append("    else if (li_any) {")
                for k in range(V):
                    B.append(f"      if (tc + {k} >= out_lo && tc + {k} < out_hi) p[{k}] = {on[k]};")
                B.append("    }")
            elif V == 1:
                B.append(f"    if (li_any) p[0] = {on[0]};")
            else:
                for k in range(V):
                    B.append(f"    if (tc + {k} >= out_lo && tc + {k} < out_hi) p[{k}] = {on[k]};")
            # fused ghost-cell writes for Cyclic axes: the wrap the reference evaluates with % on every
            # read (PlanTrans.hs:477-484) is materialised once per written cell
            if cyclic:
                dst.append("    if (edge_any) { if (li_any && (edge_x || row < g.yorg + g.gy_hi || row >= g.yorg + g.nyl - g.gy_lo)) {")
                for k in range(V):
                    dst.append(f"      if (tc + {k} >= out_lo && tc + {k} < out_hi) {{ const int c = tc + {k} - g.xorg; const int r = row - g.yorg;")
                    dst.append("        const int dc = !g.cyc_x ? 0 : (c < g.gx_hi ? g.nx : (c >= g.nx - g.gx_lo ? -g.nx : 0));")
                    dst.append("        const int dr = !g.wrap_y_local ? 0 : (r < g.gy_hi ? g.nyl : (r >= g.nyl - g.gy_lo ? -g.nyl : 0));")
                    dst.append(f"        if (dc) p[{k} + dc] = {on[k]};")
                    dst.append(f"        if (dr) p[{k} + (ptrdiff_t)dr * g.pitch] = {on[k]};")
                    dst.append(f"        if (dc && dr) p[{k} + dc + (ptrdiff_t)dr * g.pitch] = {on[k]};")
                    dst.append(f"        if (g.cyc_x && c < g.gx_hi && c >= g.nx - g.gx_lo) p[{k} - g.nx] = {on[k]};   // domain narrower than the ghost width")
                    dst.append(f"        if (g.wrap_y_local && r < g.gy_hi && r >= g.nyl - g.gy_lo) p[{k} - (ptrdiff_t)g.nyl * g.pitch] = {on[k]};")
                    dst.append("      }")
                dst.append("    } }")
            dst.append("  }" + ("}" if zo else ""))
        carried_now = [False]     # the carried scope's values live in its own block: it keeps the plain form

        def accumulate(v, rop, slot, names, ind):
            cls = {"Sum": "OmSum", "Min": "OmMin", "Max": "OmMax"}[rop]
            chain = f"acc{slot}"
            for k in range(V):
                chain = f"{cls}::op({chain}, {names[k]})"
            if V == 1:
                B.append(f"{ind}if (li_any) {{ acc{slot} = {chain}; }}")
                return
            if compact and not carried_now[0]:
                B.append(f"{ind}if (li_all) {{ acc{slot} = {chain}; }}")
                rare_lines.append("  if (!li_all) {")
                for k in range(V):
                    rare_lines.append(f"    if (tc + {k} >= out_lo && tc + {k} < out_hi) acc{slot} = {cls}::op(acc{slot}, {names[k]});")
                rare_lines.append("  }")
                return
            B.append(f"{ind}if (li_all) {{ acc{slot} = {chain}; }}")
            B.append(f"{ind}else if (li_any) {{")
            for k in range(V):
                B.append(f"{ind}  if (tc + {k} >= out_lo && tc + {k} < out_hi) acc{slot} = {cls}::op(acc{slot}, {names[k]});")
            B.append(f"{ind}}}")
        for n, (v, rop, slot) in enumerate(st.reduce_targets):
            zo = rplane(n)
            if zo:
                B.append(f"  if ({zlive(zo)}) {{")
            accumulate(v, rop, slot, [oname(v, zo, k) for k in range(V)], "    " if zo else "  ")
            if zo:
                B.append("  }")
        if rare_lines:
            self.has_rare = True
        if self._clean:
            rare_lines = []
        if rare_lines and self.tuning.cold_rare:
            # the block as a function of its own (a noinline closure that captures what it reads by value and returns the accumulators):
            # inline, ptxas lays its ~100 instructions per vector lane out in the middle of every row body, and the hot path of a
            # row is two short runs with a 5 KB jump between them (Life: a 38 KB loop of which 8 KB execute)
            slots = sorted({int(m) for l in rare_lines for m in re.findall(r"\bacc(\d+)\b", l)})
            styp = {slot: self.T(v) for (v, _rop, slot) in self.st.reduce_targets + self.st.carried}
            body = [re.sub(r"\bacc(\d+)\b", lambda m: f"om_a.s{m.group(1)}", l) for l in rare_lines]
            B.append("  if (rare) {   // (cold) partial vectors at the strip's edge, cells with a ghost copy")
            if slots:
                B.append("    struct OmRare { " + " ".join(f"{styp[s_]} s{s_};" for s_ in slots) + " };")
                B.append("    const OmRare om_rr = [=]() __attribute__((noinline)) -> OmRare {")
                B.append("      OmRare om_a; " + " ".join(f"om_a.s{s_} = acc{s_};" for s_ in slots))
                B += ["    " + l for l in body]
                B.append("      return om_a;")
                B.append("    }();")
                B.append("    " + " ".join(f"acc{s_} = om_rr.s{s_};" for s_ in slots))
            else:
                B.append("    [=]() __attribute__((noinline)) {")
                B += ["    " + l for l in body]
                B.append("    }();")
            B.append("  }")
        elif rare_lines and lean_lines:
            # an edge strip's ghost column is written on every row by one warp of the CTA, and the CTA waits for it at the row's
            # barrier: that warp gets 4 compare + store pairs per vector instead of the general block's ~300 instructions
            B.append("  if (rare) {   // partial vectors at the strip's edge, cells with a ghost copy")
            B.append("    if (ghost_only) {   // a whole vector whose only extra is the ghost copy across x")
            B += ["    " + l for l in lean_lines]
            B.append("    } else {")
            B += ["    " + l for l in rare_lines]
            B.append("    }")
            B.append("  }")
        elif rare_lines:
            B.append("  if (rare) {   // partial vectors at the strip's edge, cells with a ghost copy")
            B += ["  " + l for l in rare_lines]
            B.append("  }")
        if st.carried:
            carried_now[0] = True
            # the level-0 reduce of the NEXT call of this kernel, evaluated on the values just stored (schedule.find_carry)
            B.append("  {   // carried reduce: next call's level-0 stage becomes an 8-byte copy")
            lines, res = self.scope_guarded([v for (v, _o, _k) in st.carried], 0, tag="c")
            B += ["    " + l for l in lines]
            for (v, rop, slot) in st.carried:
                accumulate(v, rop, slot, [res[(v, k)] for k in range(V)], "    ")
            B.append("  }")
        B.append("}")
        return B

    def reduce_slots_once(self) -> List[Tuple[int, str, int]]:
        """One (value, op, slot) per reduce slot of the stage (a rank-3 CTA that computes several planes accumulates
        several values into the same slot)."""
        seen, out = set(), []
        for (v, rop, slot) in self.st.reduce_targets + self.st.carried:
            if slot not in seen:
                seen.add(slot)
                out.append((v, rop, slot))
        return out

    def emit_reduce_epilogue(self) -> List[str]:
        st = self.st
        L: List[str] = []
        if not (st.reduce_targets or st.carried):
            return L
        L.append("  // block reduce -> per-CTA partial -> the last CTA folds all partials (om_runtime.cuh)")
        for t, (v, rop, slot) in enumerate(self.reduce_slots_once()):
            T = self.T(v)
            cls = {"Sum": "OmSum", "Min": "OmMin", "Max": "OmMax"}[rop]
            ident = {"Sum": f"({T})0", "Min": self.type_max(v), "Max": self.type_min(v)}[rop]
            L.append(f"  {{ __shared__ {T} red{slot}[32]; {T} result;")
            L.append(f"    {T}* partials = reinterpret_cast<{T}*>(red_partials + (size_t){t} * gridDim.x * gridDim.y * gridDim.z);")
            L.append(f"    if (om_block_reduce_finalize<{cls}, {T}, NT>(acc{slot}, {ident}, partials, red_counter + {t}, red{slot}, result)) {{")
            L.append(f"      om_slot_store<{T}>(sc, {slot}, g.red_accumulate ? {cls}::op(om_slot_load<{T}>(sc, {slot}), result) : result);")
            L.append(f"      red_counter[{t}] = 0u;")
            L.append("    }")
            L.append("    __syncthreads();")
            L.append("  }")
        return L

    def launcher(self) -> str:
        st = self.st
        sv = self.om.setup.static_values
        in_statics = sorted({i.static_idx for i in st.inputs.values()})
        out_statics = list(dict.fromkeys(s for (s, _v) in st.store_targets))
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
        L.append("  const int chunks = g->nchunks > 0 ? (g->nchunks < nrows ? g->nchunks : nrows) : (nrows + g->chunk_rows - 1) / g->chunk_rows;")
        L.append("  static bool attr_set[64] = {};   // function attributes are per device")
        L.append("  int dev = 0; cudaGetDevice(&dev);")
        L.append(f"  if (dev >= 64 || !attr_set[dev]) {{ cudaError_t e = cudaFuncSetAttribute({self.name}_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, {max(smem, 1)}); if (e != cudaSuccess) return (int)e; if (dev < 64) attr_set[dev] = true; }}")
        L.append(f"  const int planes = (g->own_z1 - g->own_z0 + {st.zplanes - 1}) / {st.zplanes};   // rank 3: one layer of CTAs per group of {st.zplanes} plane(s) of axis 2")
        L.append("  if (planes <= 0) return 0;")
        L.append(f"  OM_LAUNCH({self.name}_kernel, dim3(strips, chunks, planes), {self.NT}, {smem}, (cudaStream_t)stream, {', '.join(args)});")
        L.append("  OM_CUDA_CHECK_LAUNCH();")
        L.append("  return 0;")
        L.append("}")
        L.append(f"// resident CTAs per SM for this stage (the host sizes the grid to whole waves with it)")
        L.append(f'extern "C" int {self.name}_occupancy(void) {{')
        L.append(f"  int n = 0;")
        L.append(f"  if (cudaFuncSetAttribute({self.name}_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, {max(smem, 1)}) != cudaSuccess) return -1;")
        L.append(f"  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, {self.name}_kernel, {self.NT}, {smem}) != cudaSuccess) return -1;")
        L.append("  return n;")
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


def pick_vnt(om: OM, st: Stage, ks: KernelSchedule, tuning) -> Tuple[int, int]:
    """Cells per thread and threads per CTA.  16-byte vectors for 4-byte cells; wide double-
    precision DAGs (register-bound) use one cell per thread."""
    sv = om.setup.static_values
    types = [sv[i.static_idx].namee.type for i in st.inputs.values()] + [sv[s].namee.type for (s, _v) in st.store_targets]
    types += [ks.ops[v].ctype for (v, _o, _k) in st.reduce_targets + st.carried]
    width = max([TYPE_BYTES[t] for t in types] + [4])
    if st.mats:
        return (tuning.cells_heavy, tuning.threads_heavy)
    return (16 // width if width <= 8 else 1, tuning.threads_light)
