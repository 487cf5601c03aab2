"""B200 schedule: re-cut an OM kernel into fused, row-streaming GPU stages.

This replaces the reference's subkernel cut (OMTrans.hs:145-161: one flat loop per OMWriteGroup,
every Manifest value a full HBM array, PlanTrans.hs:406-596) with:

  * one GPU *stage* per reduce level.  A `Reduce` is the only global barrier in an OM kernel
    (OM/Graph.hs:116), so stage L computes the inputs of the level-L reduces and the array
    stores of level L; anything it needs from earlier levels is recomputed from the static
    arrays instead of being written to HBM.  Reduce results live in device scalar slots.
  * inside a stage a CTA owns a strip of columns and streams along axis 1.  Values that are
    read through a non-trivial `Shift` and are not cheap to recompute are *materialised* once
    per cell in a shared-memory ring of rows ("MAT" nodes); everything else is evaluated in
    registers at the cursor it is requested at (the reference's Delayed semantics,
    PlanTrans.hs:527-544).  Static input arrays are staged in shared-memory rings too.
  * phases (groups of MAT nodes of equal depth) are separated by one CTA barrier inside a row
    iteration; row lags A(m) and ring depths D(m) follow from the stencil's axis-1 offsets.

The numerical result of every node is the reference's: same SSA DAG, no reassociation.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Set, Tuple

from ... import annotation as A
from ...om.graph import ARRAY, SCALAR, Graph, Inst, Kernel, OM

Cursor = Tuple[int, int]

OP_COST = {"Div": 6, "Sqrt": 8, "Inv": 6, "Mod": 6, "Exp": 10, "Log": 10, "Sin": 12, "Cos": 12, "Tan": 14,
           "Asin": 14, "Acos": 14, "Atan": 14, "Atan2": 16, "Pow": 20, "Identity": 0, "Cast": 0}
MAT_THRESHOLD = 3


@dataclass
class Op:
    """A value node with its defining instruction folded in."""
    vid: int
    kind: str                  # Load Imm LoadIndex LoadSize Shift Arith Reduce Broadcast
    inst: Inst
    args: List[int]
    realm: str
    ctype: str
    valid: Optional[A.Valid] = None
    zoff: int = 0              # rank-3 machines after lower_z: axis-2 offset of a Load / LoadIndex(2) relative to the CTA's plane


@dataclass
class MatNode:
    vid: int
    level: int = 0        # phase level (1..); inputs are level 0
    lag: int = 0          # A(m): computed for row j + lag at iteration j
    depth: int = 1        # ring rows
    xlo: int = 0          # needed columns beyond the output strip (>= 0 each side)
    xhi: int = 0
    early: int = 0        # E(m) <= 0: first iteration (relative to the chunk start) it must run
    rd_xlo: int = 0       # most negative / positive column offset it is read at (ring padding)
    rd_xhi: int = 0


@dataclass
class InputArr:
    static_idx: int
    vid: int              # the Load value node
    ctype: str
    lag: int = 0
    depth: int = 1
    via_smem: bool = False
    xlo: int = 0
    xhi: int = 0
    early: int = 0
    rd_xlo: int = 0
    rd_xhi: int = 0
    zoff: int = 0         # rank 3: the plane this virtual input reads, relative to the CTA's plane


@dataclass
class Stage:
    """One fused GPU kernel."""
    kernel: str
    level: int
    store_targets: List[Tuple[int, int]] = field(default_factory=list)    # (static idx, value id)
    reduce_targets: List[Tuple[int, str, int]] = field(default_factory=list)  # (value id, op, slot)
    inputs: Dict[int, InputArr] = field(default_factory=dict)              # by Load value id
    mats: Dict[int, MatNode] = field(default_factory=dict)
    phases: List[List[int]] = field(default_factory=list)                  # MAT vids per level
    scalar_roots: List[int] = field(default_factory=list)                   # scalar-realm values needed (uniform)
    warmup: int = 0
    halo_x: Tuple[int, int] = (0, 0)     # thread coverage beyond the output strip
    pad_x: Tuple[int, int] = (0, 0)      # ring padding beyond thread coverage
    out_level: int = 1                   # phase in which the OUT scope (stores / reduces) runs
    mat_candidates: List[dict] = field(default_factory=list)   # shifted values: {vid, op, cost, default, chosen}
    carried: List[Tuple[int, str, int]] = field(default_factory=list)   # (value id, op, slot): next call's level-0 reduces
    zplanes: int = 1                                                    # rank 3: planes of axis 2 one CTA computes
    store_plane: Dict[Tuple[int, int], int] = field(default_factory=dict)   # (static, value id) -> plane offset inside the CTA's group
    reduce_plane: List[int] = field(default_factory=list)                   # plane offset of every reduce_targets entry


@dataclass
class KernelSchedule:
    name: str
    ops: Dict[int, Op]
    stages: List[Stage]
    scalar_stores: List[Tuple[int, int]]          # (static idx, value id), Scalar realm
    array_stores: List[Tuple[int, int]]
    reduce_slots: Dict[int, int]                  # Reduce result value id -> slot
    loaded_arrays: List[int]                      # static idx of arrays read by any stage
    carry: Optional[dict] = None                  # reduce carried to the next call (find_carry)
    extra_slots: int = 0                          # scalar slots used beyond the reduce slots
    sink_stats: Optional[dict] = None             # selectsink.sink_selects report (weighted DAG size before / after)


COMMUTATIVE = {"Add", "Mul", "And", "Or", "EQ", "NE"}


def fold_ops(g: Graph, dim: int, normalize: bool = True, pull_shifts: bool = False) -> Tuple[Dict[int, Op], List[Tuple[int, int]]]:
    """Fold (inst, value) node pairs into ops and hash-cons them.

    The reference never merges structurally identical nodes (an un-`bind`-ed Builder expression
    is re-run at every use, OM/Builder/Internal.hs:163-164, so e.g. Hydro's boundary-condition
    selects exist dozens of times).  All OM instructions are pure, so merging identical
    (op, operands) pairs, composing chained Shifts and dropping zero Shifts is exact.

    `normalize` orders the operands of commutative IEEE / boolean operators (a + b and b + a are the same bits).
    `pull_shifts` rewrites f(shift_s a, shift_s b, c) with position-independent c into shift_s f(a, b, c) — the value of
    a pure per-cell function read at cell x - s — which merges e.g. the "left cell" / "right cell" sound speeds of
    Hydro's first-order walls (980 -> 936 ops).  It is off: with the present materialisation rule the merged values
    become 9 more shared-memory rings and a third phase (134 KB, one CTA per SM), a net loss on the B200."""
    ops: Dict[int, Op] = {}
    stores: List[Tuple[int, int]] = []
    canon: Dict[int, int] = {}
    table: Dict[tuple, int] = {}
    indep: Set[int] = set()          # position-independent values (same for every cell)

    def intern(i: int, inst: Inst, args: List[int], realm, ctype, valid) -> int:
        payload = inst.arg if inst.op != "Imm" else (repr(inst.arg), inst.imm_type)
        if normalize and inst.op == "Arith" and inst.arg in COMMUTATIVE and len(args) == 2:
            args = sorted(args)
        key = (inst.op, payload, inst.cast_to, tuple(args), realm, ctype,
               repr(valid) if (inst.op == "Arith" and inst.arg == "Identity") else None)   # an Identity exists to carry its Valid region
        if key in table:
            return table[key]
        table[key] = i
        ops[i] = Op(i, inst.op, inst, list(args), realm, ctype, valid)
        if realm == SCALAR or inst.op in ("Imm", "Broadcast", "LoadSize") or (inst.op in ("Arith", "Shift") and all(a in indep for a in args)):
            indep.add(i)
        return i

    for i, nd in enumerate(g.nodes):
        if nd.is_value:
            p, inst = g.pre_inst(i)
            args = [canon[a] for a in g.nodes[p].pre]
            valid = A.to_maybe(A.Valid, nd.anot)
            realm, ctype = nd.value.realm, nd.value.type
            if inst.op == "Shift":
                vec = tuple(inst.arg)
                src = args[0]
                if ops[src].kind == "Shift":
                    vec = tuple(a + b for a, b in zip(vec, ops[src].inst.arg))
                    src = ops[src].args[0]
                if all(x == 0 for x in vec) or src in indep:
                    if valid is not None and ops[src].valid is not None and valid != ops[src].valid:
                        # a shifted position-independent value has the same value everywhere but a smaller Valid region
                        # (BoundaryAnalysis.hs:85-94): where it is stored or reduced, the cells outside that region stay 0
                        # in the reference.  Keep the region on an Identity node instead of dropping the Shift.
                        canon[i] = intern(i, Inst("Arith", "Identity"), [src], realm, ctype, valid)
                    else:
                        canon[i] = src
                    continue
                inst = Inst("Shift", vec)
                args = [src]
            elif pull_shifts and inst.op == "Arith" and realm == ARRAY:
                vecs = {tuple(ops[a].inst.arg) for a in args if ops[a].kind == "Shift"}
                if len(vecs) == 1 and all(ops[a].kind == "Shift" or a in indep for a in args):
                    inner_args = [ops[a].args[0] if ops[a].kind == "Shift" else a for a in args]
                    # the new per-cell op takes the id of this value's instruction node (unused as an op id, and
                    # between the operands' ids and i, so ascending ids stay a topological order)
                    inner = intern(p, inst, inner_args, realm, ctype, None)
                    inst = Inst("Shift", next(iter(vecs)))
                    args = [inner]
            canon[i] = intern(i, inst, args, realm, ctype, valid)
        elif nd.inst.op == "Store":
            stores.append((nd.inst.arg, canon[nd.pre[0]]))
    return ops, stores


def lower_z(ops: Dict[int, Op], stores: List[Tuple[int, int]], zplanes: int = 1):
    """Rank-3 machines: turn the op DAG into a rank-2 DAG per plane of axis 2.

    A CTA of a rank-3 machine works inside one plane z of axis 2 (blockIdx.z), streaming along axis 1 exactly like a
    rank-2 kernel.  The axis-2 component of every Shift is pushed down to the leaves: the value of v needed at plane
    offset cz is a copy of v's expression whose array Loads (and LoadIndex 2) carry `zoff = cz` — a "virtual input"
    that is simply the same static array addressed one or more planes away — and whose Shifts keep their (axis 0,
    axis 1) part only.  This is the reference's own evaluation rule (every Delayed value is recomputed at the cursor
    it is requested at, PlanTrans.hs:527-544) applied to axis 2; along axes 0 and 1 the shared-memory rings of the
    rank-2 schedule still remove the recomputation.  Neighbouring planes are re-read by the CTAs of the planes next to
    them, i.e. from L2.

    With `zplanes` = Z > 1 a CTA computes Z consecutive planes: every store exists once per plane offset zo in [0, Z)
    and a Reduce folds its argument at all Z offsets, so the rows of Z + 2r planes are staged for Z planes of output
    instead of 1 + 2r for one (the virtual inputs of neighbouring offsets coincide and are merged).
    Returns (ops, [(static, value, zo)], {value: zo})."""
    def zc(a: int, cz: int) -> int:      # position-independent values exist once
        o = ops[a]
        if o.realm == SCALAR or o.kind in ("Imm", "Broadcast", "LoadSize") or (o.kind == "LoadIndex" and o.inst.arg != 2):
            return 0
        return cz
    need: Dict[int, Set[int]] = {v: set() for v in ops}
    for (_s, v) in stores:
        if ops[v].realm == ARRAY:
            need[v].update(zc(v, zo) for zo in range(zplanes))
        else:
            need[v].add(0)
    for v in sorted(ops):
        if ops[v].kind == "Reduce":
            need[v].add(0)
    for v in sorted(ops, reverse=True):
        o = ops[v]
        for cz in need[v]:
            if o.kind == "Shift":
                a = o.args[0]
                need[a].add(zc(a, cz - tuple(o.inst.arg)[2]))
            elif o.kind == "Reduce":
                for a in o.args:
                    need[a].update(zc(a, zo) for zo in range(zplanes))
            elif o.kind == "Broadcast" or o.realm == SCALAR:
                for a in o.args:
                    need[a].add(0)
            else:
                for a in o.args:
                    need[a].add(zc(a, cz))
    out: Dict[int, Op] = {}
    m: Dict[Tuple[int, int], int] = {}
    table: Dict[tuple, int] = {}
    for v in sorted(ops):
        o = ops[v]
        for cz in sorted(need[v]):
            inst, kind, zoff = o.inst, o.kind, 0
            if kind == "Shift":
                vec = tuple(o.inst.arg)
                a = o.args[0]
                src = m[(a, zc(a, cz - vec[2]))]
                if vec[0] == 0 and vec[1] == 0:
                    m[(v, cz)] = src
                    continue
                if out[src].kind == "Shift":     # compose with a Shift that survived below
                    vec = (vec[0] + out[src].inst.arg[0], vec[1] + out[src].inst.arg[1], 0)
                    src = out[src].args[0]
                    if vec[0] == 0 and vec[1] == 0:
                        m[(v, cz)] = src
                        continue
                inst, args = Inst("Shift", (vec[0], vec[1])), [src]
            elif kind == "Reduce":       # the argument at every plane offset of the CTA's group, in offset order
                args = [m[(a, zc(a, zo))] for a in o.args for zo in range(zplanes)]
            elif kind == "Broadcast" or o.realm == SCALAR:
                args = [m[(a, 0)] for a in o.args]
            else:
                args = [m[(a, zc(a, cz))] for a in o.args]
                if kind == "Load" or (kind == "LoadIndex" and o.inst.arg == 2):
                    zoff = cz
            payload = inst.arg if inst.op != "Imm" else (repr(inst.arg), inst.imm_type)
            key = (inst.op, payload, inst.cast_to, tuple(args), o.realm, o.ctype, zoff)
            if key not in table:
                nid = len(out)
                out[nid] = Op(nid, kind, inst, args, o.realm, o.ctype, o.valid, zoff)
                table[key] = nid
            m[(v, cz)] = table[key]
    stores_z = []
    for (s_, v) in stores:
        if ops[v].realm == ARRAY:
            stores_z += [(s_, m[(v, zc(v, zo))], zo) for zo in range(zplanes)]
        else:
            stores_z.append((s_, m[(v, 0)], 0))
    return out, stores_z


def _cost(op: Op) -> int:
    if op.kind in ("Load", "Imm", "LoadIndex", "LoadSize", "Broadcast", "Shift"):
        return 0
    return OP_COST.get(op.inst.arg, 1)


def _pad2(c, dim) -> Cursor:
    c = tuple(c)
    return (c[0], c[1] if dim > 1 else 0)


class StageBuilder:
    def __init__(self, ops: Dict[int, Op], dim: int, stage: Stage, mat_threshold: int = MAT_THRESHOLD, mat_flip=()):
        self.ops, self.dim, self.stage = ops, dim, stage
        self.mat_threshold = mat_threshold
        self.mat_flip = set(mat_flip)          # value ids whose materialise / recompute decision is inverted
        self.stage_z_inputs = True
        self.candidates: List[dict] = []       # every shifted value with its cost and decision (for the schedule search)

    # -- closure of array nodes needed by the stage (through shifts), and scalar roots
    def closure(self, roots: List[int]) -> Set[int]:
        seen: Set[int] = set()
        stack = list(roots)
        while stack:
            v = stack.pop()
            if v in seen:
                continue
            seen.add(v)
            op = self.ops[v]
            if op.realm == SCALAR:
                continue  # handled by scalar closure
            if op.kind == "Broadcast":
                self.stage.scalar_roots.append(op.args[0])
                continue
            stack.extend(op.args)
        return seen

    def choose_mats(self, needed: Set[int]) -> Set[int]:
        """A value read through a non-zero Shift is materialised when recomputing it from the
        nearest materialised values / inputs costs more than MAT_THRESHOLD weighted ops."""
        ops = self.ops
        shifted: Set[int] = set()
        for v in needed:
            op = ops[v]
            if op.realm == ARRAY and op.kind == "Shift" and any(x != 0 for x in op.inst.arg):
                src = op.args[0]
                while ops[src].kind == "Shift":   # chained shifts compose
                    src = ops[src].args[0]
                shifted.add(src)
        mats: Set[int] = set()
        cost: Dict[int, int] = {}
        for v in sorted(needed):
            op = ops[v]
            if op.realm != ARRAY:
                cost[v] = 0
                continue
            c = _cost(op) + sum(cost.get(a, 0) for a in op.args if ops[a].realm == ARRAY)
            if v in shifted and op.kind not in ("Load", "Imm", "LoadIndex", "Broadcast") and c > 0:
                default = c > self.mat_threshold
                chosen = default != ((self.stage.kernel, v) in self.mat_flip)
                self.candidates.append(dict(vid=v, op=op.inst.arg, cost=c, default=default, chosen=chosen))
                if chosen:
                    mats.add(v)
                    c = 0
            cost[v] = c
        return mats

    def reads(self, root: int, mats: Set[int]) -> Dict[Tuple[int, Cursor], None]:
        """(boundary node, cursor) pairs read when `root` is evaluated at cursor 0 with every
        non-materialised value recomputed inline (PlanTrans.hs:546-570 restricted to one phase)."""
        ops, dim = self.ops, self.dim
        out: Dict[Tuple[int, Cursor], None] = {}
        seen: Set[Tuple[int, Cursor]] = set()
        stack = [(root, (0, 0), True)]
        while stack:
            v, cur, is_root = stack.pop()
            if (v, cur) in seen:
                continue
            seen.add((v, cur))
            op = ops[v]
            if op.realm == SCALAR:
                continue
            if not is_root and (v in mats or op.kind == "Load"):
                out[(v, cur)] = None
                continue
            if op.kind == "Load":      # a root that is itself a Load (store x <- load y)
                out[(v, cur)] = None
                continue
            if op.kind == "Shift":
                s = _pad2(op.inst.arg, dim)
                stack.append((op.args[0], (cur[0] - s[0], cur[1] - s[1]), False))
            elif op.kind in ("Arith", "StoredValue"):
                for a in op.args:
                    stack.append((a, cur, False))
        return out

    def build(self, roots: List[int]):
        st, ops = self.stage, self.ops
        needed = self.closure(roots)
        mat_set = self.choose_mats(needed)
        # reads per phase root
        rd: Dict[int, Dict[Tuple[int, Cursor], None]] = {}
        for m in mat_set:
            rd[m] = self.reads(m, mat_set)
        OUT = -1
        out_reads: Dict[Tuple[int, Cursor], None] = {}
        for r in roots:
            if r in mat_set or ops[r].kind == "Load":
                out_reads[(r, (0, 0))] = None
            else:
                out_reads.update(self.reads(r, mat_set))
        rd[OUT] = out_reads
        # prune MAT nodes not reachable from OUT
        live: Set[int] = set()
        stack = [OUT]
        while stack:
            n = stack.pop()
            for (b, _c) in rd[n]:
                if b in mat_set and b not in live:
                    live.add(b)
                    stack.append(b)
        mat_set = live
        # consumers first: OUT, then MAT nodes by descending id (ids are topologically ordered)
        order = [OUT] + sorted(mat_set, reverse=True)
        lag = {OUT: 0}
        xlo = {OUT: 0}
        xhi = {OUT: 0}
        early = {OUT: 0}
        info: Dict[int, dict] = {}
        for n in order:
            for (b, c) in rd[n]:
                d = info.setdefault(b, dict(rd_xlo=0, rd_xhi=0, uses=[]))
                d["uses"].append((n, c))
        producers = sorted(info.keys(), reverse=True)
        for b in producers:
            uses = info[b]["uses"]
            lag[b] = max(lag[n] + c[1] for (n, c) in uses)
            xlo[b] = max([0] + [xlo[n] - c[0] for (n, c) in uses])
            xhi[b] = max([0] + [xhi[n] + c[0] for (n, c) in uses])
            lo_row = min(lag[n] + c[1] for (n, c) in uses)
            info[b]["depth"] = lag[b] - lo_row + 1
            early[b] = min(early[n] + lag[n] + c[1] - lag[b] for (n, c) in uses)
            info[b]["rd_xlo"] = max([0] + [-c[0] for (_n, c) in uses])
            info[b]["rd_xhi"] = max([0] + [c[0] for (_n, c) in uses])
        # phase levels.  Rows written in earlier iterations are already behind the loop-top
        # barrier; only same-iteration reads constrain the order, and only reads from another
        # thread's column (c0 != 0) need a CTA barrier in between.
        level: Dict[int, int] = {}
        for n in sorted(mat_set) + [OUT]:
            l = 1
            for (b, c) in rd[n]:
                if b in mat_set and lag[n] + c[1] == lag[b]:
                    l = max(l, level[b] + (1 if c[0] != 0 else 0))
            level[n] = l
        st.out_level = level[OUT]
        for b in producers:
            d = info[b]
            if b in mat_set:
                st.mats[b] = MatNode(vid=b, level=level[b], lag=lag[b], depth=d["depth"], xlo=xlo[b], xhi=xhi[b],
                                     early=early[b], rd_xlo=d["rd_xlo"], rd_xhi=d["rd_xhi"])
            else:
                op = ops[b]
                # through shared memory only when another thread's columns are read; same-column reads
                # at several rows are plain (L1/L2-resident) global loads
                # (rank 3: the neighbouring planes' rows are staged too — they come from L2 / HBM with the same latency
                #  as the centre plane's, and the staging pipeline is what hides it)
                via = any(c[0] != 0 for (_n, c) in d["uses"]) or (op.zoff != 0 and self.stage_z_inputs)
                st.inputs[b] = InputArr(static_idx=op.inst.arg, vid=b, ctype=op.ctype, lag=lag[b], depth=d["depth"],
                                        via_smem=via, xlo=xlo[b], xhi=xhi[b], early=early[b],
                                        rd_xlo=d["rd_xlo"], rd_xhi=d["rd_xhi"], zoff=op.zoff)
        nlev = max([m.level for m in st.mats.values()] + [st.out_level])
        st.phases = [[m.vid for m in sorted(st.mats.values(), key=lambda m: m.vid) if m.level == l] for l in range(1, nlev + 1)]
        st.warmup = -min([0] + [m.early for m in st.mats.values()] + [i.early for i in st.inputs.values() if i.via_smem])
        st.halo_x = (max([0] + [m.xlo for m in st.mats.values()]), max([0] + [m.xhi for m in st.mats.values()]))
        ringed = list(st.mats.values()) + [i for i in st.inputs.values() if i.via_smem]
        st.pad_x = (max([0] + [r.rd_xlo for r in ringed]), max([0] + [r.rd_xhi for r in ringed]))
        self.reads_of = rd
        return mat_set


def find_carry(ops: Dict[int, Op], rl: Dict[int, int], reduce_slots: Dict[int, int], array_stores, scalar_stores,
               levels: List[int]) -> Optional[dict]:
    """Reduces that can be carried from one call of the kernel to the next.

    Hydro's `proceed` starts with a pass over the four state arrays whose only result is dt = cfl * min(...) — 32 of the
    96 bytes per cell the step moves.  That reduce reads the arrays at cursor 0 only, and the arrays it
    reads are exactly the ones the kernel's last stage stores.  So the last stage can evaluate the same expression on
    the values it is about to store and reduce them for the *next* call; the next call then replaces its level-0 stage
    by an 8-byte copy, as long as nothing else wrote the arrays or the scalars in between (the host side tracks that).

    Conditions: the level-0 stage stores nothing and each of its reduce arguments (a) contains no Shift, (b) loads only
    arrays that the kernel stores, all at its last level, (c) loads no scalar the kernel stores and uses no reduce
    result.  Returns the cloned expression roots (Load s -> StoredValue(value stored to s)) or None."""
    if len(levels) < 2 or levels[0] != 0:
        return None
    if any(rl[v] == 0 for (_s, v) in array_stores):
        return None
    level0 = [r for r in sorted(reduce_slots) if rl[ops[r].args[0]] == 0]
    if not level0:
        return None
    last = levels[-1]
    stored = {s: v for (s, v) in array_stores}
    if any(rl[v] != last for v in stored.values()):
        return None
    scalar_stored = {s for (s, _v) in scalar_stores}
    clone: Dict[int, int] = {}
    new_ops: Dict[int, Op] = {}
    fresh = [max(ops) + 1]
    arrays: Set[int] = set()
    scalars: Set[int] = set()

    def scalar_ok(v: int) -> bool:
        stack, seen = [v], set()
        while stack:
            x = stack.pop()
            if x in seen:
                continue
            seen.add(x)
            o = ops[x]
            if o.kind == "Reduce":
                return False
            if o.kind == "Load":
                if o.inst.arg in scalar_stored:
                    return False
                scalars.add(o.inst.arg)
            stack.extend(o.args)
        return True

    def go(v: int) -> Optional[int]:
        """Clone of v with loads replaced (v itself when nothing below it loads an array); None if not carriable."""
        if v in clone:
            return clone[v]
        o = ops[v]
        if o.realm == SCALAR:
            r = v if scalar_ok(v) else None
        elif o.kind == "Shift":
            r = None
        elif o.kind == "Load":
            if o.inst.arg not in stored:
                r = None
            else:
                arrays.add(o.inst.arg)
                r = fresh[0]
                fresh[0] += 1
                new_ops[r] = Op(r, "StoredValue", Inst("StoredValue", o.inst.arg), [stored[o.inst.arg]], o.realm, o.ctype, None)
        elif o.kind in ("Imm", "LoadIndex", "LoadSize"):
            r = v
        elif o.kind in ("Arith", "Broadcast"):
            args = [go(a) for a in o.args]
            if any(a is None for a in args):
                r = None
            elif args == o.args:
                r = v
            else:
                r = fresh[0]
                fresh[0] += 1
                new_ops[r] = Op(r, o.kind, o.inst, args, o.realm, o.ctype, o.valid)
        else:
            r = None
        clone[v] = r
        return r

    roots = []
    for r in level0:
        c = go(ops[r].args[0])
        if c is None or c == ops[r].args[0]:
            return None
        roots.append((r, c, ops[r].inst.arg))
    return dict(roots=roots, new_ops=new_ops, arrays=sorted(arrays), scalars=sorted(scalars), level=last)


def schedule_kernel(om: OM, kernel: Kernel, slot_base: int, mat_threshold: int = MAT_THRESHOLD, mat_flip=(),
                    carry_reduces: bool = True, zplanes: int = 1, sink_selects: bool = True,
                    fast_algebra: bool = False, pull_shifts: bool = False) -> KernelSchedule:
    g = kernel.dataflow
    dim = om.dim
    ops, stores = fold_ops(g, dim, pull_shifts=pull_shifts)
    sink_stats = None
    if fast_algebra:
        from .selectsink import simplify_fast
        ops, stores, _removed = simplify_fast(ops, stores)
    if sink_selects:
        from .selectsink import sink_selects as _sink
        ops, stores, sink_stats = _sink(ops, stores)
    plane_of: Dict[Tuple[int, int], int] = {}
    if dim == 3:
        ops, stores_z = lower_z(ops, stores, zplanes)
        stores = [(s_, v) for (s_, v, _zo) in stores_z]
        plane_of = {(s_, v): zo for (s_, v, zo) in stores_z}
    else:
        zplanes = 1
    # reduce levels
    rl: Dict[int, int] = {}
    for v in sorted(ops):
        op = ops[v]
        base = max([rl[a] for a in op.args] + [0])
        rl[v] = base + 1 if op.kind == "Reduce" else base
    reduce_slots: Dict[int, int] = {}
    for v in sorted(ops):
        if ops[v].kind == "Reduce":
            reduce_slots[v] = slot_base + len(reduce_slots)
    array_stores = [(s, v) for (s, v) in stores if ops[v].realm == ARRAY]
    scalar_stores = [(s, v) for (s, v) in stores if ops[v].realm == SCALAR]
    levels = sorted({rl[v] for (_s, v) in array_stores} |
                    {rl[ops[v].args[0]] for v in reduce_slots})
    stages: List[Stage] = []
    loaded: Set[int] = set()
    found = find_carry(ops, rl, reduce_slots, array_stores, scalar_stores, levels) if (carry_reduces and zplanes == 1) else None
    carry = None
    if found:
        ops.update(found["new_ops"])
        nxt = slot_base + len(reduce_slots)
        carry = dict(skip_stage=0, level=found["level"], arrays=found["arrays"], scalars=found["scalars"],
                     pairs=[(reduce_slots[r], nxt + n) for n, (r, _c, _o) in enumerate(found["roots"])])
    for L in levels:
        st = Stage(kernel=kernel.name, level=L)
        st.store_targets = [(s, v) for (s, v) in array_stores if rl[v] == L]
        st.reduce_targets = [(a, ops[v].inst.arg, reduce_slots[v]) for v in sorted(reduce_slots)
                             if rl[ops[v].args[0]] == L for a in ops[v].args]
        st.reduce_plane = [zo for v in sorted(reduce_slots) if rl[ops[v].args[0]] == L for zo in range(len(ops[v].args))]
        st.zplanes = zplanes
        st.store_plane = {(s_, v): plane_of.get((s_, v), 0) for (s_, v) in st.store_targets}
        roots = [v for (_s, v) in st.store_targets] + [v for (v, _o, _k) in st.reduce_targets]
        if found and L == found["level"]:
            st.carried = [(c, rop, carry["pairs"][n][1]) for n, (_r, c, rop) in enumerate(found["roots"])]
            roots += [c for (c, _o, _k) in st.carried]
        sb = StageBuilder(ops, dim, st, mat_threshold, mat_flip)
        sb.build(list(dict.fromkeys(roots)))
        st.mat_candidates = sb.candidates
        loaded |= {i.static_idx for i in st.inputs.values()}
        stages.append(st)
    return KernelSchedule(name=kernel.name, ops=ops, stages=stages, scalar_stores=scalar_stores,
                          array_stores=array_stores, reduce_slots=reduce_slots, loaded_arrays=sorted(loaded),
                          carry=carry, extra_slots=len(carry["pairs"]) if carry else 0, sink_stats=sink_stats)
