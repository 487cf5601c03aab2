"""Register-streaming skeleton for stages without shared-memory intermediates.

For a stage whose stores / reduces depend on the static arrays only through small stencils (no MAT
nodes; column offsets no larger than the per-thread vector width), the shared-memory rings and CTA
barriers of cuda.StageEmitter are unnecessary.  Here every warp is independent:

  * a thread owns V consecutive columns and streams along axis 1; the rows of the stencil window
    live in a rotating set of registers (the row loop is unrolled by window + prefetch depth, so the
    rotation is a renaming, not a copy);
  * each row is read from HBM exactly once per thread with one 128-bit load, issued self.PREFETCH rows
    ahead of its use;
  * x-neighbours come from the adjacent lanes with __shfl_up/down; only lane 0 and lane 31 fetch
    their halo cell from global memory (an L1/L2 hit: the neighbouring warp streams that column);
  * stores are 128-bit; reduces accumulate in registers and finish with the shared block/grid fold.

This is the path Life takes (one int32 array, 3x3 stencil): ~15 issued instructions per cell update
instead of the ~47 of the first shared-memory version (profiles/r1_life_*.txt).
"""
from __future__ import annotations

from typing import List

from ...om.graph import CPP_TYPE
from .cuda import VEC_TYPE, StageEmitter



def eligible(st, V: int, tuning) -> bool:
    # Measured on B200 (profiles/r1_life_sweep.txt): for Life the shared-memory ring skeleton with
    # cp.async staging sustains more bytes in flight per SM (5.8 TB/s) than register streaming
    # (4.3 TB/s, register-limited occupancy), so streaming is opt-in (Tuning.skeleton = "stream").
    if tuning.skeleton != "stream" or st.mats or st.carried:
        return False
    for i in st.inputs.values():
        if i.rd_xlo > V or i.rd_xhi > V:
            return False
    return True


class WarpStreamEmitter(StageEmitter):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        st = self.st
        self.depth = {}                      # no shared memory at all
        self.ring_inputs = []
        self.PL = self.PR = 0
        self.RW = self.NT * self.V
        ins = list(st.inputs.values())
        for i in ins:
            i.via_smem = True                # "staged" = streamed through the register window here
        # window of rows (relative to the output row) that the stencil touches, per stage
        self.cmin = min([i.lag - i.depth + 1 for i in ins] + [0])
        self.cmax = max([i.lag for i in ins] + [0])
        self.PREFETCH = self.tuning.stream_prefetch     # rows in flight per thread ahead of the stencil window
        self.U = self.cmax - self.cmin + 1 + self.PREFETCH
        self.cur_u = 0

    def smem_bytes(self) -> int:
        return 0

    # registers of row-set s of input b
    def qn(self, b, s, k): return f"q{b}_{s}_{k}"
    def hm(self, b, s, i): return f"h{b}_{s}_m{i}"
    def hp(self, b, s, i): return f"h{b}_{s}_p{i}"

    def staged_read(self, lines, ring_rd, lag, b, cur, k) -> str:
        s = (self.cur_u + cur[1]) % self.U
        o = k + cur[0]
        if 0 <= o < self.V:
            return self.qn(b, s, o)
        if o < 0:
            return self.hm(b, s, -o)
        return self.hp(b, s, o - self.V + 1)

    # direct reads do not exist in this skeleton: every input streams through registers
    def load_row(self, u_set: int, row_expr: str) -> List[str]:
        """Issue the global loads of one row into register set `u_set` for every input (no bounds
        predicates: the ABI requires OM_APRON_ROWS allocated rows around every array)."""
        V = self.V
        L = [f"{{ const int rr = {row_expr};"]
        for i in self.st.inputs.values():
            b, T = i.vid, self.T(i.vid)
            vt = VEC_TYPE.get((T, V))
            L.append(f"  const {T}* __restrict__ p{b} = in{i.static_idx} + (ptrdiff_t)rr * g.pitch + tc;")
            if vt:
                L.append(f"  {{ const {vt} q = __ldg(reinterpret_cast<const {vt}*>(p{b})); " +
                         " ".join(f"{self.qn(b, u_set, k)} = q.{'xyzw'[k]};" for k in range(V)) + " }")
            else:
                L.append("  " + " ".join(f"{self.qn(b, u_set, k)} = __ldg(p{b} + {k});" for k in range(V)))
            for h in range(1, i.rd_xlo + 1):
                L.append(f"  if (lane == 0) {self.hm(b, u_set, h)} = __ldg(p{b} - {h});")
            for h in range(1, i.rd_xhi + 1):
                L.append(f"  if (lane == 31) {self.hp(b, u_set, h)} = __ldg(p{b} + {V - 1 + h});")
        L.append("}")
        return L

    def finish_row(self, u_set: int) -> List[str]:
        """Neighbour exchange for a row that has arrived: lanes 1..31 / 0..30 take their halo cells
        from the adjacent lane's vector; the warp-edge lanes keep what they loaded themselves."""
        V = self.V
        L = []
        for i in self.st.inputs.values():
            b, T = i.vid, self.T(i.vid)
            for h in range(1, i.rd_xlo + 1):
                L.append(f"{{ const {T} t = __shfl_up_sync(0xffffffffu, {self.qn(b, u_set, V - h)}, 1); if (lane != 0) {self.hm(b, u_set, h)} = t; }}")
            for h in range(1, i.rd_xhi + 1):
                L.append(f"{{ const {T} t = __shfl_down_sync(0xffffffffu, {self.qn(b, u_set, h - 1)}, 1); if (lane != 31) {self.hp(b, u_set, h)} = t; }}")
        return L

    def kernel(self) -> str:
        st, V, NT, U = self.st, self.V, self.NT, self.U
        om = self.om
        self.pre_names = set()
        in_statics = sorted({i.static_idx for i in st.inputs.values()})
        out_statics = [s for (s, _v) in st.store_targets]
        sv = om.setup.static_values
        params = ["const OmGeom g"]
        for s in in_statics:
            params.append(f"const {CPP_TYPE[sv[s].namee.type]}* __restrict__ in{s}")
        for s in out_statics:
            params.append(f"{CPP_TYPE[sv[s].namee.type]}* __restrict__ out{s}")
        params += ["om_slot_t* __restrict__ sc", "unsigned* __restrict__ red_counter", "om_slot_t* __restrict__ red_partials"]
        mlx, mhx = self.margin_lo[0], self.margin_hi[0]
        cmin, cmax = self.cmin, self.cmax
        # unrolled bodies
        bodies: List[List[str]] = []
        for u in range(U):
            self.cur_u = u
            B: List[str] = [f"if (j + {u} < r1) {{"]
            B += ["  " + l for l in self.load_row((u + cmax + self.PREFETCH) % U, f"j + {u + cmax + self.PREFETCH}")]
            B += ["  " + l for l in self.finish_row((u + cmax) % U)]
            B += ["  " + l for l in self.emit_out(row_expr=f"j + {u}", guard="true")]
            B.append("}")
            bodies.append(B)
        L: List[str] = []
        E = L.append
        E(f"// stage {self.idx} of kernel `{self.ks.name}` (reduce level {st.level}): register streaming, no shared memory;")
        E(f"// stencil rows {cmin}..{cmax}, {self.PREFETCH} rows prefetched, row loop unrolled x{U}, {V} cell(s) per thread")
        E(f"__global__ void __launch_bounds__({NT}) {self.name}_kernel({', '.join(params)}) {{")
        E(f"  constexpr int V = {V}, NT = {NT}, HL = 0, W_OUT = {self.W_OUT};")
        E("  const int tid = threadIdx.x;")
        E("  const int lane = tid & 31;")
        E(f"  const int cx0 = g.xorg - {mlx}, cx1 = g.xorg + g.nx + {mhx};   // columns of the reference memory box")
        E("  const int cA = (cx0 / V) * V;")
        E("  const int strip_lo = cA + blockIdx.x * W_OUT;            // first output column of this CTA")
        E("  const int tc = strip_lo + tid * V;                       // first column of this thread")
        E("  const int r0 = g.own_r0 + blockIdx.y * g.chunk_rows;")
        E("  const int r1 = min(r0 + g.chunk_rows, g.own_r1);")
        for l in self.scalar_code(list(dict.fromkeys(st.scalar_roots))):
            E("  " + l)
        for l in self.uniform_code():
            E("  " + l)
        for l in self.pre:
            E("  " + l)
        for (v, rop, slot) in st.reduce_targets:
            T = self.T(v)
            ident = {"Sum": f"({T})0", "Min": self.type_max(v), "Max": self.type_min(v)}[rop]
            E(f"  {T} acc{slot} = {ident};   // reduce slot {slot}")
        # register sets
        for i in st.inputs.values():
            T = self.T(i.vid)
            names = []
            for s in range(U):
                names += [self.qn(i.vid, s, k) for k in range(V)]
                names += [self.hm(i.vid, s, h) for h in range(1, i.rd_xlo + 1)]
                names += [self.hp(i.vid, s, h) for h in range(1, i.rd_xhi + 1)]
            E(f"  {T} " + ", ".join(f"{n} = 0" for n in names) + ";")
        E("  // prologue: fill the stencil window and the prefetch queue of the first row")
        for c in range(cmin, cmax + self.PREFETCH):
            for l in self.load_row(c % U, f"r0 + ({c})"):
                E("  " + l)
        for c in range(cmin, cmax):
            for l in self.finish_row(c % U):
                E("  " + l)
        E(f"  for (int j = r0; j < r1; j += {U}) {{")
        for B in bodies:
            L += ["    " + l for l in B]
        E("  }")
        L += self.emit_reduce_epilogue()
        E("}")
        return "\n".join(L)
