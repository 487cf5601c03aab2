"""Select sinking: evaluate a formula once on selected operands instead of once per alternative.

The reference evaluates every alternative of a `select` and picks one afterwards (PlanTrans.hs:670-710 prints
`c ? a : b` on two fully computed values).  Hydro's HLLC solver (examples/Hydro/HydroMain.hs:237-276) ends in

    select (0 < shockLeft) left (select (0 < shockStar) leftStar (select (0 < shockRight) rightStar right))

applied to every flux component, where `left`/`right` and `leftStar`/`rightStar` are the same formulas evaluated on
the left and on the right state: both star states and both plain fluxes are computed for every wall although one of
them is used.  All OM instructions are pure, so

    select k (f a1 b1 ..) (f a2 b2 ..)  ==  f (select k a1 a2) (select k b1 b2) ..          (bit for bit)

whenever the two alternatives are the same operator tree over different leaves.  This pass finds such pairs by
anti-unification of the two operand DAGs, after rotating a select chain so that corresponding alternatives face
each other:

    select c1 A (select c2 B (select c3 C D))  ==  select (c1 || c2) (select c1 A B) (select c3 C D)
    select c3 C D                              ==  select (!c3) D C

and keeps a rewrite only if the weighted size of the live DAG shrinks.  Operands of commutative IEEE operators may be
matched in either order (a + b and b + a are the same bits; fold_ops already relies on that).  Nothing is
reassociated and no operation changes: the value of every surviving node is the reference's.
"""
from __future__ import annotations

from collections import Counter
from typing import Dict, List, Optional, Tuple

from ...om.graph import ARRAY, Inst

COMMUTATIVE = {"Add", "Mul", "And", "Or", "EQ", "NE"}
OP_COST = {"Div": 6, "Sqrt": 8, "Inv": 6, "Mod": 6, "Exp": 10, "Log": 10, "Sin": 12, "Cos": 12, "Tan": 14,
           "Asin": 14, "Acos": 14, "Atan": 14, "Atan2": 16, "Pow": 20, "Identity": 0, "Cast": 0}


def _cost(op) -> int:
    if op.kind != "Arith" or op.realm != ARRAY:      # scalar-realm arithmetic is thread-uniform and hoisted out of the row loop
        return 0
    return OP_COST.get(op.inst.arg, 1)


class _Sinker:
    def __init__(self, ops: Dict[int, "Op"], roots: List[int], protected: set):
        from .schedule import Op
        self.Op = Op
        self.ops: Dict[int, "Op"] = dict(ops)
        self.roots = list(roots)
        self.protected = set(protected)
        self.fresh = max(ops) + 1 if ops else 0
        self.order: Dict[int, Tuple[int, int, int]] = {v: (v, 0, 0) for v in ops}   # topological sort key
        self.table: Dict[tuple, int] = {}
        for v in sorted(ops):
            self.table.setdefault(self._key(ops[v].kind, ops[v].inst, ops[v].args, ops[v].realm, ops[v].ctype, ops[v].zoff), v)
        self.repl: Dict[int, int] = {}
        self.votes: Counter = Counter()        # (x, y) pairs that faced each other at unambiguous operand positions
        self._tc: Dict[int, int] = {}
        self._mc: Dict[Tuple[int, int], Tuple[int, Optional[bool]]] = {}

    # ---- hash-consed node creation ---------------------------------------------------------------
    @staticmethod
    def _key(kind, inst: Inst, args, realm, ctype, zoff=0):
        payload = inst.arg if inst.op != "Imm" else (repr(inst.arg), inst.imm_type)
        return (kind, inst.op, payload, inst.cast_to, tuple(args), realm, ctype, zoff)

    def mk(self, opname: str, args: List[int], ctype: str, cast_to=None) -> int:
        if opname in COMMUTATIVE and len(args) == 2:
            args = sorted(args)
        inst = Inst("Arith", opname, cast_to)
        key = self._key("Arith", inst, args, ARRAY, ctype)
        v = self.table.get(key)
        if v is not None:
            return v
        v = self.fresh
        self.fresh += 1
        self.ops[v] = self.Op(v, "Arith", inst, list(args), ARRAY, ctype, None)
        anchor = max(self.order[a] for a in args)
        self.order[v] = (anchor[0], 1, v)
        self.table[key] = v
        return v

    # ---- additive (tree) cost estimates that steer the matching ------------------------------------
    def tree_cost(self, v: int) -> int:
        c = self._tc.get(v)
        if c is None:
            op = self.ops[v]
            c = _cost(op) + (sum(self.tree_cost(a) for a in op.args) if op.kind == "Arith" and op.realm == ARRAY else 0)
            self._tc[v] = c
        return c

    def descendable(self, x: int, y: int) -> bool:
        ox, oy = self.ops[x], self.ops[y]
        return (ox.kind == "Arith" and oy.kind == "Arith" and ox.realm == ARRAY and oy.realm == ARRAY
                and ox.inst.arg == oy.inst.arg and ox.inst.cast_to == oy.inst.cast_to and ox.ctype == oy.ctype
                and len(ox.args) == len(oy.args))

    def selectable(self, x: int, y: int) -> bool:
        ox, oy = self.ops[x], self.ops[y]
        return ox.ctype == oy.ctype

    def pairings(self, x: int, y: int) -> List[Tuple[Optional[bool], List[Tuple[int, int]]]]:
        """Ways of matching the operands of two nodes with the same operator: [(tag, [(a, b), ...])]."""
        ox, oy = self.ops[x], self.ops[y]
        out = [(False, list(zip(ox.args, oy.args)))]
        if ox.inst.arg in COMMUTATIVE and len(ox.args) == 2:
            out.append((True, [(ox.args[0], oy.args[1]), (ox.args[1], oy.args[0])]))
        return out

    def merged_cost(self, x: int, y: int) -> Tuple[int, Optional[bool]]:
        """(tree cost of the best unification of x and y, chosen pairing tag or None for a leaf select)."""
        if x == y:
            return self.tree_cost(x), None
        r = self._mc.get((x, y))
        if r is not None:
            return r
        leaf = 1 + self.tree_cost(x) + self.tree_cost(y)
        best: Tuple[int, Optional[bool]] = (leaf, None)
        if self.descendable(x, y):
            cands = []
            for tag, prs in self.pairings(x, y):
                if not all(self.selectable(a, b) for a, b in prs):
                    continue
                c = _cost(self.ops[x]) + sum(self.merged_cost(a, b)[0] for a, b in prs)
                cands.append((c, -sum(self.votes[(a, b)] for a, b in prs if a != b), tag, prs))
            if cands:
                cands.sort(key=lambda t: (t[0], t[1], t[2]))
                c, _v, tag, prs = cands[0]
                if c < leaf:
                    best = (c, tag)
                    unambiguous = len(cands) == 1 or cands[1][0] > c
                    if unambiguous:
                        self._pending_votes.extend((a, b) for a, b in prs if a != b)
        self._mc[(x, y)] = best
        return best

    def vote(self, candidates: List[Tuple[int, int]]):
        """Two dry passes: operand pairs seen at unambiguous positions break the ties of commutative operators."""
        for _ in range(3):
            self._mc.clear()
            self._pending_votes: List[Tuple[int, int]] = []
            for x, y in candidates:
                self.merged_cost(x, y)
            new = Counter(self._pending_votes)
            if new == self.votes:
                break
            self.votes = new
        self._mc.clear()
        self._pending_votes = []
        for x, y in candidates:
            self.merged_cost(x, y)

    # ---- the rewrite -------------------------------------------------------------------------------
    def unify(self, k: int, x: int, y: int, memo: Dict[Tuple[int, int], int]) -> int:
        """Node computing `k ? x : y` with the select pushed towards the leaves."""
        if x == y:
            return x
        r = memo.get((x, y))
        if r is not None:
            return r
        _c, tag = self.merged_cost(x, y)
        if tag is None:
            r = self.mk("Select", [k, x, y], self.ops[x].ctype)
        else:
            prs = dict(self.pairings(x, y))[tag]
            ox = self.ops[x]
            r = self.mk(ox.inst.arg, [self.unify(k, a, b, memo) for a, b in prs], ox.ctype, ox.inst.cast_to)
        memo[(x, y)] = r
        return r

    def chain(self, s: int):
        """conds [c1..cn], values [V1..Vn], else E of the select chain rooted at s (else-branches followed)."""
        conds, vals = [], []
        v = s
        while True:
            op = self.ops[v]
            if not (op.kind == "Arith" and op.inst.arg == "Select" and op.realm == ARRAY and op.ctype == self.ops[s].ctype):
                break
            conds.append(op.args[0])
            vals.append(op.args[1])
            v = op.args[2]
        return conds, vals, v

    def build_chain(self, conds, vals, els, ctype) -> int:
        r = els
        for c, v in reversed(list(zip(conds, vals))):
            r = self.mk("Select", [c, v, r], ctype)
        return r

    def alternatives(self, s: int):
        """[(k, first, second)]: s == k ? first : second, for every split of the chain and both orientations of `second`."""
        conds, vals, els = self.chain(s)
        ctype = self.ops[s].ctype
        n = len(conds)
        out = []
        for m in range(1, n + 1):
            k = conds[0]
            for c in conds[1:m]:
                k = self.mk("Or", [k, c], "Bool")
            first = self.build_chain(conds[:m - 1], vals[:m - 1], vals[m - 1], ctype)
            second = self.build_chain(conds[m:], vals[m:], els, ctype)
            out.append((k, first, second))
            if m < n and m >= 2 and n - m == 1:
                # second == select c D E == select (!c) E D: lets `first`'s value order face the mirrored one
                nc = self.mk("Not", [conds[m]], "Bool")
                out.append((k, first, self.mk("Select", [nc, els, vals[m]], ctype)))
        return out

    # ---- live DAG bookkeeping ------------------------------------------------------------------------
    def resolve(self, v: int) -> int:
        while v in self.repl:
            v = self.repl[v]
        return v

    def live_cost(self) -> int:
        seen, total = set(), 0
        stack = [self.resolve(r) for r in self.roots]
        while stack:
            v = stack.pop()
            if v in seen:
                continue
            seen.add(v)
            op = self.ops[v]
            total += _cost(op)
            stack.extend(self.resolve(a) for a in op.args)
        return total

    def select_roots(self) -> Dict[tuple, List[int]]:
        """Live select-chain roots (selects that are not the else-branch of a live select), grouped by their conditions."""
        live, inner = set(), set()
        stack = [self.resolve(r) for r in self.roots]
        while stack:
            v = stack.pop()
            if v in live:
                continue
            live.add(v)
            stack.extend(self.resolve(a) for a in self.ops[v].args)
        is_sel = lambda v: self.ops[v].kind == "Arith" and self.ops[v].inst.arg == "Select" and self.ops[v].realm == ARRAY
        for v in live:
            if is_sel(v):
                e = self.ops[v].args[2]
                if is_sel(e) and self.ops[e].ctype == self.ops[v].ctype:
                    inner.add(e)
        groups: Dict[tuple, List[int]] = {}
        for v in sorted(live, reverse=True):
            if is_sel(v) and v not in inner and v not in self.protected:
                groups.setdefault(tuple(self.chain(v)[0]), []).append(v)
        return groups

    def run(self) -> dict:
        stats = dict(groups=0, accepted=0, rewritten=0, cost_before=self.live_cost())
        groups = self.select_roots()
        stats["groups"] = len(groups)
        for conds, members in groups.items():
            alts = {s: self.alternatives(s) for s in members}
            self.vote([(f, sec) for s in members for (_k, f, sec) in alts[s]])
            before = self.live_cost()
            trial: Dict[int, int] = {}
            memos: Dict[Tuple[int, int], Dict[Tuple[int, int], int]] = {}
            for s in members:
                best = None
                for idx, (k, f, sec) in enumerate(alts[s]):
                    c, tag = self.merged_cost(f, sec)
                    if tag is None:
                        continue            # the alternatives do not even share their top operator
                    gain = 1 + self.tree_cost(f) + self.tree_cost(sec) - c
                    if best is None or gain > best[0]:
                        best = (gain, idx)
                if best is None or best[0] <= 0:
                    continue
                k, f, sec = alts[s][best[1]]
                trial[s] = self.unify(k, f, sec, memos.setdefault((k, best[1]), {}))
            if not trial:
                continue
            self.repl.update(trial)
            after = self.live_cost()
            if after < before:
                stats["accepted"] += 1
                stats["rewritten"] += len(trial)
            else:
                for s in trial:
                    del self.repl[s]
        stats["cost_after"] = self.live_cost()
        return stats

    def result(self, stores):
        """Dense renumbering of the live DAG (ascending ids stay a topological order; old nodes keep their relative order)."""
        live = set()
        stack = [self.resolve(r) for r in self.roots]
        while stack:
            v = stack.pop()
            if v in live:
                continue
            live.add(v)
            stack.extend(self.resolve(a) for a in self.ops[v].args)
        # Kahn's algorithm with the old ids as priorities: a topological order in which old nodes keep their relative
        # order wherever the rewrites allow it
        import heapq
        args_of = {v: [self.resolve(a) for a in self.ops[v].args] for v in live}
        users: Dict[int, List[int]] = {v: [] for v in live}
        indeg = {}
        for v in live:
            distinct = set(args_of[v])
            indeg[v] = len(distinct)
            for a in distinct:
                users[a].append(v)
        heap = [(self.order[v], v) for v in live if indeg[v] == 0]
        heapq.heapify(heap)
        new_id: Dict[int, int] = {}
        while heap:
            _k, v = heapq.heappop(heap)
            new_id[v] = len(new_id)
            for u in users[v]:
                indeg[u] -= 1
                if indeg[u] == 0:
                    heapq.heappush(heap, (self.order[u], u))
        assert len(new_id) == len(live), "select sinking produced a cyclic DAG"
        out = {}
        for v, n in new_id.items():
            o = self.ops[v]
            args = [new_id[self.resolve(a)] for a in o.args]
            assert all(a < n for a in args), "select sinking broke the topological order"
            out[n] = self.Op(n, o.kind, o.inst, args, o.realm, o.ctype, o.valid, o.zoff)
        return out, [(s_, new_id[self.resolve(v)]) for (s_, v) in stores]


def sink_selects(ops: Dict[int, "Op"], stores: List[Tuple[int, int]]):
    """Returns (ops, stores, stats).  The DAG is returned unchanged (same ids) when no rewrite pays."""
    roots = [v for (_s, v) in stores] + [v for v in sorted(ops) if ops[v].kind == "Reduce"]
    protected = set(v for (_s, v) in stores)
    for v in ops:
        if ops[v].kind in ("Reduce", "Broadcast"):
            protected.update(ops[v].args)
    sk = _Sinker(ops, roots, protected)
    stats = sk.run()
    if not stats["accepted"]:
        return ops, stores, stats
    new_ops, new_stores = sk.result(stores)
    return new_ops, new_stores, stats


def simplify_fast(ops: Dict[int, "Op"], stores: List[Tuple[int, int]]):
    """Algebraic clean-up for `Setup.fast_math` builds only (results within rounding of the reference, not bit-identical).

    The Builder's `sum` / `contract` fold from an explicit zero and dot products with unit vectors multiply by literal
    0 and 1 (visible in the reference's own output, Hydro.cpp:204-205), and the HLLC star state divides a product by
    one of its factors (`mome / dens` with `mome = dens * v`, HydroMain.hs:243-248):
        x * 0 -> 0      x * 1 -> x      x + 0 -> x      x - 0 -> x      x / 1 -> x      select c a a -> a
        (a * b) / b -> a                                                   (one rounding instead of two)
    The first group is exact for finite x up to the sign of a zero; the last differs from the reference by <= 1 ulp.
    Ids are kept (every replacement is an earlier node), duplicates created by the rewrites are merged.
    Returns (ops, stores, number of nodes removed)."""
    from .schedule import Op
    from ...om.graph import imm_value
    protected = set(v for (_s, v) in stores)
    for v in ops:
        if ops[v].kind in ("Reduce", "Broadcast"):
            protected.update(ops[v].args)
    canon: Dict[int, int] = {}
    out: Dict[int, Op] = {}
    table: Dict[tuple, int] = {}

    def const(v: int):
        o = out[v]
        if o.kind == "Imm" and o.ctype in ("Double", "Float"):
            return float(imm_value(o.inst.arg, o.ctype))
        return None

    for v in sorted(ops):
        o = ops[v]
        args = [canon[a] for a in o.args]
        r = None
        if o.kind == "Arith" and o.ctype in ("Double", "Float") and v not in protected:
            t = o.inst.arg
            c = [const(a) for a in args]
            if t == "Mul":
                if 0.0 in c:
                    r = args[c.index(0.0)]
                elif c[0] == 1.0:
                    r = args[1]
                elif c[1] == 1.0:
                    r = args[0]
            elif t == "Add":
                if c[0] == 0.0:
                    r = args[1]
                elif c[1] == 0.0:
                    r = args[0]
            elif t == "Sub" and c[1] == 0.0:
                r = args[0]
            elif t == "Div":
                num = out[args[0]]
                if c[1] == 1.0:
                    r = args[0]
                elif num.kind == "Arith" and num.inst.arg == "Mul" and args[1] in num.args and num.args[0] != num.args[1]:
                    r = num.args[1 - num.args.index(args[1])]
            elif t == "Select" and args[1] == args[2]:
                r = args[1]
            if r is not None and (out[r].ctype != o.ctype or out[r].realm != o.realm):
                r = None
        if r is None:
            if o.kind == "Arith" and o.inst.arg in COMMUTATIVE and len(args) == 2:
                args = sorted(args)
            key = _Sinker._key(o.kind, o.inst, args, o.realm, o.ctype, o.zoff)
            if o.kind in ("Arith", "Shift") and key in table and v not in protected:
                r = table[key]
            else:
                table.setdefault(key, v)
                out[v] = Op(v, o.kind, o.inst, args, o.realm, o.ctype, o.valid, o.zoff)
                r = v
        canon[v] = r
    # drop what is no longer reachable
    live, stack = set(), [canon[v] for (_s, v) in stores] + [v for v in out if out[v].kind == "Reduce"]
    while stack:
        v = stack.pop()
        if v not in live:
            live.add(v)
            stack.extend(out[v].args)
    res = {v: out[v] for v in sorted(live)}
    return res, [(s_, canon[v]) for (s_, v) in stores], len(ops) - len(res)
