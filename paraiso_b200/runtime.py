"""Host side of a generated machine: the Python mirror of the emitted C++ class.

The reference's emitted class owns every array as a value member and exposes
`<kernel>()`, `om_size[_k]()`, `om_memory_size[_k]()`, `om_lower/upper_margin_k()`, `name()` and
`name(i0, i1)` (PlanTrans.hs:50-215).  `Machine` keeps that surface (same names, same argument
meaning) over the C ABI of lib<Name>.so:

  * arrays live on the device in a padded, pitched layout with ghost zones; a
    `store` is a pointer swap between two buffers instead of the reference's whole-array copy
    (PlanTrans.hs:243-258);
  * scalars live in device slots; host reads synchronise, like the reference's thrust_vector
    mirror (Generator/draft.cpp:53-132);
  * with world_size > 1 the grid is slab-decomposed along the outermost axis; ghost rows travel
    with torch.distributed send/recv (NCCL over NVLink on GPUs), reduce results with all_reduce.

PyTorch supplies device memory, streams and the process group only; every cell update runs in
the generated CUDA kernels.  There is no CPU fallback: `device` must be a CUDA device unless a
test passes `_emulated=True` together with a library built by tests/emu.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import numpy as np
import torch

from .annotation import CYCLIC

TORCH_TYPE = {"Int": torch.int32, "Float": torch.float32, "Double": torch.float64, "Bool": torch.bool,
              "Integer": torch.int64}
NP_TYPE = {"Int": np.int32, "Float": np.float32, "Double": np.float64, "Bool": np.bool_, "Integer": np.int64}
APRON = 32   # = APRON_ROWS of generator/b200/cuda.py


class OmGeom(ctypes.Structure):
    """Mirror of `struct OmGeom` in csrc/om_runtime.cuh."""
    _fields_ = [(n, ctypes.c_int) for n in (
        "nx", "ny", "pitch", "rows", "xorg", "yorg", "y0", "nyl",
        "gx_lo", "gx_hi", "gy_lo", "gy_hi", "cyc_x", "cyc_y", "wrap_y_local",
        "own_r0", "own_r1", "chunk_rows", "red_accumulate",
        "nz", "plane", "zorg", "gz_lo", "gz_hi", "cyc_z", "own_z0", "own_z1", "z0", "nzl",
        "bfirst", "sig_lo", "sig_hi", "nchunks")]


_HP_GROUP: Dict[int, object] = {}


def high_priority_group():
    """One NCCL process group over all ranks whose internal streams are high-priority: the ghost-row send/recv and the
    scalar all-reduces must get SM slots while a stage kernel still has CTAs waiting to be scheduled."""
    import torch.distributed as dist
    key = dist.get_world_size()
    if key not in _HP_GROUP:
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            _HP_GROUP[key] = dist.new_group(backend="nccl", pg_options=opts)
        except Exception:
            _HP_GROUP[key] = None      # the default group works too, at default priority
    return _HP_GROUP[key]


def _ru(x, m):
    return (x + m - 1) // m * m


def slab_rows(ny: int, nranks: int, rank: int):
    """Rows [y0, y0+nyl) of the global axis 1 owned by `rank` (remainder rows go to the first ranks)."""
    base, rem = divmod(ny, nranks)
    nyl = base + (1 if rank < rem else 0)
    y0 = rank * base + min(rank, rem)
    return y0, nyl


class Machine:
    def __init__(self, desc: dict, lib_path: str, size=None, device="cuda", rank: int = 0, nranks: int = 1,
                 group=None, overlap: bool = True, _emulated: bool = False):
        self.desc = desc
        self.name = desc["name"]
        self.device = torch.device(device)
        if self.device.type != "cuda" and not _emulated:
            raise RuntimeError("paraiso_b200 machines run on CUDA devices only (no CPU fallback)")
        self.emulated = _emulated
        self.lib = ctypes.CDLL(lib_path)   # raises OSError if the extension is missing
        if getattr(self.lib, f"om_{self.name}_abi_version")() != 3:
            raise RuntimeError("ABI version mismatch")
        self.rank, self.nranks, self.group = rank, nranks, group
        size = list(size) if size is not None else list(desc["local_size"])
        size = size + [1] * (3 - len(size))
        self.nx, self.ny, self.nz = int(size[0]), int(size[1]), int(size[2])
        self.dim3 = desc["dim"] == 3
        if not self.dim3 and self.nz != 1:
            raise ValueError("a rank-2 machine takes a (nx, ny) size")
        pad3 = lambda t, fill: list(t) + [fill] * (3 - len(t))
        self.boundary = pad3(desc["boundary"], "Open")
        self.mlo, self.mhi = pad3(desc["lower_margin"], 0), pad3(desc["upper_margin"], 0)
        rlo, rhi = pad3(desc["radius_lo"], 0), pad3(desc["radius_hi"], 0)
        self.cyc = [b == CYCLIC for b in self.boundary]
        for ax in range(3):
            if not self.cyc[ax]:
                assert self.mlo[ax] == rlo[ax] and self.mhi[ax] == rhi[ax], "Open margins must equal the stencil radius"
        # the slab decomposition cuts the outermost axis: axis 1 of rank-1 / rank-2 machines (rows are contiguous), axis 2 of
        # rank-3 machines (whole planes are contiguous)
        self.y0, self.nyl = (0, self.ny) if self.dim3 else slab_rows(self.ny, nranks, rank)
        self.z0, self.nzl = slab_rows(self.nz, nranks, rank) if self.dim3 else (0, 1)
        if nranks > 1 and (self.nzl < max(rlo[2], rhi[2]) if self.dim3 else self.nyl < max(rlo[1], rhi[1])):
            raise ValueError("slab thinner than the stencil radius")
        self.gx_lo, self.gx_hi, self.gy_lo, self.gy_hi = rlo[0], rhi[0], rlo[1], rhi[1]
        self.xorg = _ru(max(self.gx_lo, 1), 32)
        self.pitch = _ru(self.xorg + self.nx + self.gx_hi, 32)
        self.yorg = self.gy_lo
        self.rows = self.nyl + self.gy_lo + self.gy_hi
        first, last = (rank == 0, rank == nranks - 1) if not self.dim3 else (True, True)
        if self.cyc[1]:
            self.own_r0, self.own_r1 = self.yorg, self.yorg + self.nyl
        else:
            self.own_r0 = 0 if first else self.yorg
            self.own_r1 = self.rows if last else self.yorg + self.nyl
        self.cx0, self.cx1 = self.xorg - self.mlo[0], self.xorg + self.nx + self.mhi[0]
        # rank 3: planes of rows * pitch elements stacked along axis 2 (ghost planes included); rank 2: one plane
        self.gz_lo, self.gz_hi = rlo[2], rhi[2]
        self.zorg = self.gz_lo
        self.planes = self.nzl + self.gz_lo + self.gz_hi
        if self.cyc[2]:
            self.own_z0, self.own_z1 = self.zorg, self.zorg + self.nzl
        else:      # Open: the end ranks also own the reference's margin planes
            self.own_z0 = 0 if rank == 0 else self.zorg
            self.own_z1 = self.planes if rank == nranks - 1 else self.zorg + self.nzl
        # storage
        self.statics = desc["statics"]
        self.index = {s["name"]: i for i, s in enumerate(self.statics)}
        self._storage: List[torch.Tensor] = []

        self.cur: List[Optional[torch.Tensor]] = []
        self.alt: List[Optional[torch.Tensor]] = []
        for s in self.statics:
            if s["realm"] == "Array":
                t = TORCH_TYPE[s["type"]]
                # APRON rows of zero-initialised slack above and below: the kernels read a few rows
                # beyond the slab while filling their pipelines and do not bounds-check (C-ABI contract)
                for lst in (self.cur, self.alt):
                    full = torch.zeros((self.planes * self.rows + 2 * APRON, self.pitch), dtype=t, device=self.device)
                    self._storage.append(full)
                    lst.append(full[APRON:APRON + self.planes * self.rows])

            else:
                self.cur.append(None)
                self.alt.append(None)
        self.nslots = desc["nslots"]
        self.sc = torch.zeros(self.nslots, dtype=torch.int64, device=self.device)
        self.max_blocks = 1 << (18 if self.dim3 else 16)
        nred = max([len(st["reduces"]) for k in desc["kernels"] for st in k["stages"]] + [1])
        self.scratch = torch.zeros(256 + 8 * self.max_blocks * nred, dtype=torch.uint8, device=self.device)
        self.kernels = {k["name"]: k for k in desc["kernels"]}
        self._ptr_cur = (ctypes.c_void_p * len(self.statics))()
        self._ptr_alt = (ctypes.c_void_p * len(self.statics))()
        self._refresh_ptrs()
        self._fn = {}
        for k in desc["kernels"]:
            for st in k["stages"]:
                f = getattr(self.lib, st["symbol"])
                f.argtypes = [ctypes.POINTER(OmGeom), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_void_p]
                f.restype = ctypes.c_int
                self._fn[st["symbol"]] = f
            if k["scalars"]:
                f = getattr(self.lib, k["scalars"])
                f.argtypes = [ctypes.POINTER(OmGeom), ctypes.c_void_p, ctypes.c_void_p]
                f.restype = ctypes.c_int
                self._fn[k["scalars"]] = f
        self.launches = 0
        self.wave_round = True       # light stages: round the chunk count to whole waves of resident CTAs (see _geom)
        self.force_chunks = 0        # tuning tools: this many chunks per strip for light stages (0: the rule of _geom)
        self.early_exchanges = 0     # stage launches whose ghost-row exchange was triggered by the in-kernel boundary signal
        self._geom_cache: Dict[str, OmGeom] = {}
        self._partial: Dict[int, dict] = {}     # static scalar index -> pending all_reduce description
        self.overlap = overlap
        self._carry_kernel: Optional[str] = None   # kernel whose carried reduces are valid for its next call
        self._comm_event = None
        self._comm_stream = torch.cuda.Stream(self.device, priority=-1) if (self.device.type == "cuda" and nranks > 1) else None
        self._ar_group = group
        if self._comm_stream is not None and group is None:
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_backend() == "nccl" and dist.get_world_size() == nranks:
                self.group = high_priority_group()
        self._wait_boundary = getattr(self.lib, f"om_{self.name}_wait_boundary")
        self._wait_boundary.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        self._wait_boundary.restype = ctypes.c_int

    # ---- reference size accessors (PlanTrans.hs:160-215) ------------------------------------
    def om_size(self, k=None):
        return self.nx * self.ny * self.nz if k is None else (self.nx, self.ny, self.nz)[k]

    def om_memory_size(self, k=None):
        m = (self.nx + self.mlo[0] + self.mhi[0], self.ny + self.mlo[1] + self.mhi[1], self.nz + self.mlo[2] + self.mhi[2])
        return m[0] * m[1] * m[2] if k is None else m[k]

    def om_lower_margin(self, k): return self.mlo[k]
    def om_upper_margin(self, k): return self.mhi[k]

    # ---- plumbing ---------------------------------------------------------------------------
    def _refresh_ptrs(self):
        for i, s in enumerate(self.statics):
            self._ptr_cur[i] = self.cur[i].data_ptr() if self.cur[i] is not None else None
            self._ptr_alt[i] = self.alt[i].data_ptr() if self.alt[i] is not None else None

    def _stream(self):
        if self.device.type == "cuda":
            return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return ctypes.c_void_p(0)

    def _geom(self, st: dict, rows=None, accumulate: bool = False, bfirst: bool = False) -> Optional[OmGeom]:
        """Launch geometry of a stage over the rank's owned rows, or over the sub-range `rows` of them.  `bfirst`: boundary-first
        chunk order with the in-kernel signal (None when the chunks do not hold a neighbour's rows whole)."""
        key = (st["symbol"], rows, accumulate, bfirst)
        if key in self._geom_cache:
            return self._geom_cache[key]
        r0, r1 = rows if rows is not None else (self.own_r0, self.own_r1)
        nrows = r1 - r0
        strips = max(1, -(-(self.cx1 - (self.cx0 // st["V"]) * st["V"]) // st["w_out"]))
        occ = getattr(self.lib, st["symbol"] + "_occupancy")()
        if occ <= 0:
            raise RuntimeError(f"{st['symbol']}: occupancy query failed ({occ})")
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count if self.device.type == "cuda" else 4
        layers = self.own_z1 - self.own_z0
        if st.get("chunk_rows", 0) <= 0:
            # heavy (shared-memory) stages: few, long chunks — every chunk pays the pipeline fill (warm-up rows) again — whose
            # CTAs fill whole waves of resident CTAs: the count that maximises (fill of the last wave) x (useful rows per
            # chunk).  Hydro 4096^2: 34 strips x 13 chunks = 442 of 444 slots, one wave; a 32768-wide slab has 269 strips,
            # where "one wave" meant ONE chunk and 269 CTAs on 444 slots (round 1) — now 13 chunks, 3497 CTAs in 8 waves.
            wave, fill = sms * occ, st["warmup"] + 2
            best, chunks = -1.0, 1
            for c in range(1, max(1, nrows // max(32, 8 * fill)) + 1):
                ctas = strips * c * layers
                score = ctas / (-(-ctas // wave) * wave) * (nrows / c) / (nrows / c + fill)
                if score > best + 1e-9:
                    best, chunks = score, c
        else:
            # light streaming stages: short chunks (Tuning.chunk_rows_light, measured in profiles/r1_life_sweep.txt)
            # keep the set of concurrently streamed rows compact; warm-up rows are L2 hits.  The count is then rounded to
            # whole waves of resident CTAs (sms * occ) when that moves it by less than 8 % (Life 16384^2: 1024 -> 1040 chunks =
            # 25.0 waves).  Measured, the kernel is hardly wave-quantised — flat within 1.5 % from 13 to 24 waves
            # (profiles/r2o_life_chunks.jsonl); on the final kernel 1040 chunks run 0.6 % faster than the unrounded 1024 and
            # 0.4-1.2 % faster than 915 ... 1165 (profiles/r2ar_life_chunkcount.jsonl): the height itself is what matters most
            chunks = max(1, nrows // st["chunk_rows"])
            wave = sms * occ
            k = max(1, round(strips * chunks * layers / wave))
            whole = (k * wave) // (strips * layers)
            if self.wave_round and whole >= 1 and abs(whole - chunks) <= 0.08 * chunks:
                chunks = whole
            if self.force_chunks:
                chunks = self.force_chunks
        # the reduction scratch holds one partial per CTA: very large slabs get fewer, taller chunks instead of more CTAs
        if strips * chunks * layers > self.max_blocks and strips * layers <= self.max_blocks:
            chunks = self.max_blocks // (strips * layers)
        chunks = max(1, min(chunks, max(nrows, 1), (1 << 32) // max(nrows, 1) - 2))      # (32-bit row arithmetic in the kernels)
        chunk_rows = -(-max(nrows, 1) // chunks)
        g = OmGeom(nx=self.nx, ny=self.ny, pitch=self.pitch, rows=self.rows, xorg=self.xorg, yorg=self.yorg,
                   y0=self.y0, nyl=self.nyl, gx_lo=self.gx_lo, gx_hi=self.gx_hi, gy_lo=self.gy_lo, gy_hi=self.gy_hi,
                   cyc_x=int(self.cyc[0]), cyc_y=int(self.cyc[1]),
                   wrap_y_local=int(self.cyc[1] and (self.nranks == 1 or self.dim3)),
                   own_r0=r0, own_r1=r1, chunk_rows=max(1, chunk_rows), red_accumulate=int(accumulate),
                   nz=self.nz, plane=(self.rows * self.pitch if self.dim3 else 0), zorg=self.zorg, gz_lo=self.gz_lo,
                   gz_hi=self.gz_hi, cyc_z=int(self.cyc[2]), own_z0=self.own_z0, own_z1=self.own_z1, z0=self.z0, nzl=self.nzl,
                   nchunks=chunks)
        if strips * chunks * layers > self.max_blocks:
            raise ValueError("grid too large for the reduction scratch")
        if bfirst:
            has_up = self.cyc[1] or self.rank < self.nranks - 1
            has_down = self.cyc[1] or self.rank > 0
            nch, shortest = chunks, nrows // chunks
            # the rows a neighbour reads must lie inside the first / the last chunk
            if nch < 2 or shortest < max(self.gy_hi, self.gy_lo) or not st.get("bfirst"):
                self._geom_cache[key] = None
                return None
            g.bfirst, g.sig_lo, g.sig_hi = 1, int(has_down and self.gy_hi > 0), int(has_up and self.gy_lo > 0)
            if not (g.sig_lo or g.sig_hi):
                self._geom_cache[key] = None
                return None
        self._geom_cache[key] = g
        return g

    # ---- kernels ------------------------------------------------------------------------------
    def _launch(self, st: dict, stream, rows=None, accumulate=False, bfirst=False):
        g = self._geom(st, rows, accumulate, bfirst)
        if g.own_r1 <= g.own_r0:
            return
        rc = self._fn[st["symbol"]](ctypes.byref(g), self._ptr_cur, self._ptr_alt, self.sc.data_ptr(),
                                    self.scratch.data_ptr(), stream)
        if rc != 0:
            raise RuntimeError(f"{st['symbol']} failed with CUDA error {rc}")
        self.launches += 1
        if self._narrow():
            for o in st["outputs"]:
                self._fill_ghosts_modular(self.alt[o])

    def _join_comm(self):
        """Make the compute stream wait for the ghost-row exchange that is still in flight on the side stream."""
        if self._comm_event is not None:
            torch.cuda.current_stream(self.device).wait_event(self._comm_event)
            self._comm_event = None

    def call(self, kernel: str):
        """Run one OM kernel (`init`, `proceed`, ...) — the emitted member function of that name.

        Several ranks: the stage that writes the kernel's array stores is followed by the ghost-row exchange of what it
        wrote, on the high-priority communication stream.  A light (streaming) stage is launched ONCE in boundary-first
        chunk order: the CTAs of the first wave compute the rows the neighbours need and raise a flag, a one-thread kernel
        on the communication stream waits for it, and the NCCL send/recv overlaps the remaining waves of the same launch.
        A heavy stage fills the GPU with one wave of CTAs, so its exchange starts when the launch ends — concurrently
        with the all-reduce of its reduce results on the compute stream.  The next kernel call waits for the exchange."""
        k = self.kernels[kernel]
        stream = self._stream()
        self._join_comm()
        stores = k["array_stores"]
        cuda = self.device.type == "cuda"
        # carried reduces (schedule.find_carry): the previous call of this same kernel already reduced the arrays it
        # stored, and nothing has written them (or the scalars involved) since -> the level-0 stage is an 8-byte copy
        carry = k.get("carry")
        use_carry = bool(carry) and self._carry_kernel == kernel
        self._carry_kernel = None
        exchanged = False
        for si, st in enumerate(k["stages"]):
            if use_carry and si == carry["skip_stage"]:
                for (slot, cslot) in carry["pairs"]:
                    self.sc[slot:slot + 1].copy_(self.sc[cslot:cslot + 1])
                continue
            last_storing = self.nranks > 1 and stores and sorted(st["outputs"]) == sorted(stores) and not self._narrow()
            early = (last_storing and self.overlap and st.get("chunk_rows", 0) > 0 and not self.dim3
                     and self._geom(st, bfirst=True) is not None)
            if early and cuda:
                # the communication stream is ordered after everything enqueued BEFORE the launch, not after the launch itself
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self._launch(st, stream, bfirst=True)
                with torch.cuda.stream(self._comm_stream):
                    self._comm_stream.wait_event(ev)
                    rc = self._wait_boundary(self.scratch.data_ptr(), ctypes.c_void_p(self._comm_stream.cuda_stream))
                    if rc != 0:
                        raise RuntimeError(f"om_{self.name}_wait_boundary failed with CUDA error {rc}")
                    self.launches += 1
                    self._exchange_rows([self.alt[s_] for s_ in stores])       # alt becomes cur at the swap below
                    self._comm_event = torch.cuda.Event()
                    self._comm_event.record(self._comm_stream)
                self.early_exchanges += 1
                exchanged = True
            elif early:      # emulated kernels (tests): same launch geometry, everything in order
                self._launch(st, stream, bfirst=True)
                self._wait_boundary(self.scratch.data_ptr(), None)
                if int(self.scratch[136:140].view(torch.int32)[0]) != 0:
                    raise RuntimeError("the boundary CTAs did not signal")
                self._exchange_rows([self.alt[s_] for s_ in stores])
                self.early_exchanges += 1
                exchanged = True
            else:
                self._launch(st, stream)
                if last_storing:
                    if cuda:
                        ev = torch.cuda.Event()
                        ev.record(torch.cuda.current_stream(self.device))
                        with torch.cuda.stream(self._comm_stream):
                            self._comm_stream.wait_event(ev)
                            self._exchange_rows([self.alt[s_] for s_ in stores])
                            self._comm_event = torch.cuda.Event()
                            self._comm_event.record(self._comm_stream)
                    else:
                        self._exchange_rows([self.alt[s_] for s_ in stores])
                    exchanged = True
            if self.nranks > 1:
                for r in st["reduces"]:
                    if r.get("deferred"):
                        # nothing on the device consumes this value: each rank keeps its partial result and the
                        # all_reduce happens when the host reads the scalar (a collective read, see scalar())
                        for sx in r["stored_to"]:
                            self._partial[sx] = r
                    else:
                        self._allreduce_slot(r)
        if k["scalars"]:
            g = self._geom(k["stages"][0]) if k["stages"] else OmGeom(nx=self.nx, ny=self.ny, nz=self.nz)   # scalar code only reads the sizes
            rc = self._fn[k["scalars"]](ctypes.byref(g), self.sc.data_ptr(), stream)
            if rc != 0:
                raise RuntimeError(f"{k['scalars']} failed with CUDA error {rc}")
            self.launches += 1
        for s in stores:
            self.cur[s], self.alt[s] = self.alt[s], self.cur[s]
            if self.dim3 and self.cyc[2] and self.nranks == 1:
                self._fill_z_ghosts(self.cur[s])     # whole planes (their x / y ghost cells were written by the kernel)
        self._refresh_ptrs()
        if self.nranks > 1 and stores and not exchanged:
            self._exchange_rows([self.cur[s] for s in stores])
        if carry:
            self._carry_kernel = kernel

    def call_stage(self, kernel: str, idx: int):
        """Launch a single array stage without the pointer swap (benchmark / profiling hook)."""
        st = self.kernels[kernel]["stages"][idx]
        g = self._geom(st)
        rc = self._fn[st["symbol"]](ctypes.byref(g), self._ptr_cur, self._ptr_alt, self.sc.data_ptr(),
                                    self.scratch.data_ptr(), self._stream())
        if rc != 0:
            raise RuntimeError(f"{st['symbol']} failed with CUDA error {rc}")
        self.launches += 1

    def __getattr__(self, item):
        ks = self.__dict__.get("kernels", {})
        if item in ks:
            return lambda: self.call(item)
        raise AttributeError(item)

    # ---- multi-rank pieces ------------------------------------------------------------------------
    def _allreduce_slot(self, r: dict):
        import torch.distributed as dist
        op = {"Sum": dist.ReduceOp.SUM, "Min": dist.ReduceOp.MIN, "Max": dist.ReduceOp.MAX}[r["op"]]
        view = self.sc[r["slot"]:r["slot"] + 1].view(TORCH_TYPE[r["type"]])[:1]
        # (the default group when the ghost rows travel on the high-priority one: two communicators, so that the exchange on the
        #  communication stream and this all-reduce on the compute stream run side by side instead of queueing behind each other)
        dist.all_reduce(view, op=op, group=self._ar_group)   # 4-byte types reduce the low half of the slot; the rest stays zero

    def _exchange_rows(self, arrays):
        """Ghost rows <- neighbours' boundary interior rows (full pitch, so x ghosts travel too), for one array or a list
        of arrays in ONE NCCL group (one communication kernel per exchange).  Rank-3 machines are cut along axis 2:
        whole ghost planes travel (their x / y ghost cells included)."""
        import torch.distributed as dist
        if isinstance(arrays, torch.Tensor):
            arrays = [arrays]
        n, r = self.nranks, self.rank
        up, down = (r + 1) % n, (r - 1) % n
        ops = []
        for a in arrays:
            if self.dim3:
                a, g_lo, g_hi, y0, y1, cyc = self._v3(a), self.gz_lo, self.gz_hi, self.zorg, self.zorg + self.nzl, self.cyc[2]
            else:
                g_lo, g_hi, y0, y1, cyc = self.gy_lo, self.gy_hi, self.yorg, self.yorg + self.nyl, self.cyc[1]
            has_up = cyc or r < n - 1
            has_down = cyc or r > 0
            if g_lo and has_up:     # my top interior rows are the upper neighbour's lower ghost rows
                ops.append(dist.P2POp(dist.isend, a[y1 - g_lo:y1], up, group=self.group))
            if g_hi and has_down:   # my bottom interior rows are the lower neighbour's upper ghost rows
                ops.append(dist.P2POp(dist.isend, a[y0:y0 + g_hi], down, group=self.group))
            if g_lo and has_down:
                ops.append(dist.P2POp(dist.irecv, a[0:g_lo], down, group=self.group))
            if g_hi and has_up:
                ops.append(dist.P2POp(dist.irecv, a[y1:y1 + g_hi], up, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _v3(self, a: torch.Tensor) -> torch.Tensor:
        """[plane, row, column] view of an array (one plane for rank-2 machines)."""
        return a.view(self.planes, self.rows, self.pitch)

    def _fill_z_ghosts(self, a: torch.Tensor):
        """Rank 3, Cyclic axis 2: ghost planes <- the interior planes they wrap to (contiguous device copies)."""
        a3 = self._v3(a)
        z0, z1 = self.zorg, self.zorg + self.nzl
        if self.gz_lo:
            a3[0:self.gz_lo] = a3[z1 - self.gz_lo:z1]
        if self.gz_hi:
            a3[z1:z1 + self.gz_hi] = a3[z0:z0 + self.gz_hi]

    def _narrow(self) -> bool:
        """A Cyclic axis this rank holds whole is narrower than its two ghost zones together.  A cell then has more images
        than the kernels' fused ghost writes produce (one per direction and axis, plus the corner), so the host redoes the
        wrap after every launch — grids of a few cells only, never the measured path."""
        cached = self.__dict__.get("_narrow_cached")
        if cached is not None:
            return cached
        whole_y = self.nranks == 1 or self.dim3
        self._narrow_cached = bool((self.cyc[0] and self.nx < self.gx_lo + self.gx_hi) or
                    (self.cyc[1] and whole_y and self.nyl < self.gy_lo + self.gy_hi) or
                    (self.dim3 and self.cyc[2] and self.nranks == 1 and self.nzl < self.gz_lo + self.gz_hi))
        return self._narrow_cached

    @staticmethod
    def _wrap_axis(v: torch.Tensor, dim: int, org: int, n: int, g_lo: int, g_hi: int):
        """Ghost cells of one axis <- interior cell (j mod n), for any ghost width (also wider than the interior)."""
        if g_lo or g_hi:
            idx = torch.remainder(torch.arange(-g_lo, n + g_hi, device=v.device), n) + org
            v.narrow(dim, org - g_lo, g_lo + n + g_hi).copy_(v.index_select(dim, idx))

    def _fill_ghosts_modular(self, a: torch.Tensor):
        v = self._v3(a) if self.dim3 else a
        d = 1 if self.dim3 else 0
        if self.cyc[0]:
            self._wrap_axis(v, d + 1, self.xorg, self.nx, self.gx_lo, self.gx_hi)
        if self.cyc[1] and (self.nranks == 1 or self.dim3):
            self._wrap_axis(v, d, self.yorg, self.nyl, self.gy_lo, self.gy_hi)
        if self.dim3 and self.cyc[2] and self.nranks == 1:
            self._wrap_axis(v, 0, self.zorg, self.nzl, self.gz_lo, self.gz_hi)

    def _fill_ghosts(self, a: torch.Tensor):
        """Host-initiated ghost refresh after the host wrote an array (not on the step path)."""
        if self._narrow():
            self._fill_ghosts_modular(a)
            if self.nranks > 1:
                self._exchange_rows(a)
            return
        if self.dim3:
            a3 = self._v3(a)
            x0, x1 = self.xorg, self.xorg + self.nx
            y0, y1 = self.yorg, self.yorg + self.nyl
            if self.cyc[0]:
                if self.gx_lo:
                    a3[:, :, x0 - self.gx_lo:x0] = a3[:, :, x1 - self.gx_lo:x1]
                if self.gx_hi:
                    a3[:, :, x1:x1 + self.gx_hi] = a3[:, :, x0:x0 + self.gx_hi]
            if self.cyc[1]:
                if self.gy_lo:
                    a3[:, 0:self.gy_lo] = a3[:, y1 - self.gy_lo:y1]
                if self.gy_hi:
                    a3[:, y1:y1 + self.gy_hi] = a3[:, y0:y0 + self.gy_hi]
            if self.nranks > 1:
                self._exchange_rows(a)
            elif self.cyc[2]:
                self._fill_z_ghosts(a)
            return
        x0, x1 = self.xorg, self.xorg + self.nx
        if self.cyc[0]:
            if self.gx_lo:
                a[:, x0 - self.gx_lo:x0] = a[:, x1 - self.gx_lo:x1]
            if self.gx_hi:
                a[:, x1:x1 + self.gx_hi] = a[:, x0:x0 + self.gx_hi]
        if self.nranks > 1:
            self._exchange_rows(a)
        elif self.cyc[1]:
            y0, y1 = self.yorg, self.yorg + self.nyl
            if self.gy_lo:
                a[0:self.gy_lo] = a[y1 - self.gy_lo:y1]
            if self.gy_hi:
                a[y1:y1 + self.gy_hi] = a[y0:y0 + self.gy_hi]

    # ---- host accessors (PlanTrans.hs:117-157) ------------------------------------------------------
    def _box(self, with_margin: bool):
        """(plane, row, column) slices of the interior, or of the part of the reference's memory box this rank owns."""
        if with_margin:
            return slice(self.own_z0, self.own_z1), slice(self.own_r0, self.own_r1), slice(self.cx0, self.cx1)
        return (slice(self.zorg, self.zorg + self.nzl), slice(self.yorg, self.yorg + self.nyl),
                slice(self.xorg, self.xorg + self.nx))

    def _check_flags(self):
        """Sticky error words of the scratch header (om_runtime.cuh), read where the host synchronises anyway."""
        w = self.scratch[128:144].view(torch.int32).cpu().numpy()
        if w[2]:
            raise RuntimeError(f"{self.name}: a boundary-first launch never signalled its boundary rows (om_wait_boundary timed out)")
        if w[3]:
            raise RuntimeError(f"{self.name}: a stage of the bit-exact build stored a NaN, Inf or denormal value.  Its branch-free division / "
                               "square root are IEEE-correct for normal operands only, so from here on the state may differ from the "
                               "reference's — rebuild with Tuning.exact_divsqrt = 'ieee' (the compiler's expansion with slow paths)")

    def slow_path_cells(self) -> int:
        """Cells the bit-exact build re-evaluated with the compiler's IEEE division / sqrt so far (tiny non-zero operands;
        diagnostic counter in the scratch header, OM_SIG_SLOW)."""
        return int(self.scratch[144:148].view(torch.int32).cpu().numpy()[0])

    def get(self, name: str, with_margin: bool = False) -> np.ndarray:
        """Local slab of a static Array as [i1, i0] (axis 0 fastest), optionally with the margins
        of the reference's memory box that this rank owns; synchronises with the device."""
        i = self.index[name]
        rz, ry, rx = self._box(with_margin)
        self._join_comm()
        out = self._v3(self.cur[i])[rz, ry, rx].contiguous().cpu().numpy()
        self._check_flags()
        return out if self.dim3 else out[0]

    def set(self, name: str, values: np.ndarray, with_margin: bool = False):
        i = self.index[name]
        rz, ry, rx = self._box(with_margin)
        t = torch.as_tensor(np.ascontiguousarray(values), dtype=self.cur[i].dtype)
        self._join_comm()
        self._carry_kernel = None
        dst = self._v3(self.cur[i])[rz, ry, rx]
        dst[...] = t.to(self.device).reshape(dst.shape)
        self._fill_ghosts(self.cur[i])

    def set_from_host(self, name: str, host: torch.Tensor):
        """Asynchronous upload of the local interior from a (pinned) host tensor."""
        i = self.index[name]
        rz, ry, rx = self._box(False)
        self._join_comm()
        self._carry_kernel = None
        dst = self._v3(self.cur[i])[rz, ry, rx]
        dst.copy_(host.view(dst.shape), non_blocking=True)
        self._fill_ghosts(self.cur[i])

    def set_from_device(self, name: str, src: torch.Tensor):
        """Replace the local interior by a device tensor (stream-ordered device copy + ghost refresh)."""
        i = self.index[name]
        rz, ry, rx = self._box(False)
        self._join_comm()
        self._carry_kernel = None
        dst = self._v3(self.cur[i])[rz, ry, rx]
        dst.copy_(src.view(dst.shape))
        self._fill_ghosts(self.cur[i])

    def interior_into(self, name: str, dst: torch.Tensor):
        """Stream-ordered copy of the local interior into a device tensor (no host synchronisation)."""
        rz, ry, rx = self._box(False)
        src = self._v3(self.cur[self.index[name]])[rz, ry, rx]
        dst.view(src.shape).copy_(src)

    def capture(self, kernel: str, calls: int = 2) -> "GraphedCalls":
        """`calls` consecutive calls of an OM kernel captured into one CUDA graph (launches, ghost-row exchange,
        all-reduces and stream hops included).  `calls` must be even so that the current / alternate buffers are back in
        place after a replay.  Warm the kernel up with a few eager calls first (NCCL connections, function attributes)."""
        return GraphedCalls(self, kernel, calls)

    def scalar(self, name: str):
        """Host read of a static scalar (synchronises).  With several ranks the read of a reduce-derived scalar
        that no kernel consumes is collective: every rank must call it (the all_reduce was deferred to here)."""
        s = self.statics[self.index[name]]
        if self.nranks > 1 and self.index[name] in self._partial:
            r = self._partial.pop(self.index[name])
            self._allreduce_slot(dict(r, slot=self.index[name]))
        v = self.sc[self.index[name]:self.index[name] + 1].cpu().numpy()
        self._check_flags()
        return v.view(NP_TYPE[s["type"]])[0]

    def set_scalar(self, name: str, value):
        s = self.statics[self.index[name]]
        z = np.zeros(1, dtype=np.int64)
        z.view(NP_TYPE[s["type"]])[0] = value
        self._carry_kernel = None
        self.sc[self.index[name]] = int(z[0])

    def synchronize(self):
        self._join_comm()
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        self._check_flags()


class GraphedCalls:
    """An even number of consecutive calls of one OM kernel as a CUDA graph ("whole step captured in a CUDA graph").

    The reference's per-kernel driver is a serial list of subkernel calls (PlanTrans.hs:225-258); here a step is a handful of
    launches, stream hops and NCCL operations issued by the host, and on several GPUs that issue cost (not NVLink) is what
    separates N ranks from one.  Replaying the captured pair costs one graph launch.  The buffers swap twice per replay, so
    the baked-in pointers stay valid; host-side bookkeeping (pending partial reduces, carried-reduce validity, the launch
    counter) is replayed alongside."""

    def __init__(self, m: Machine, kernel: str, calls: int = 2):
        if calls < 2 or calls % 2:
            raise ValueError("capture an even number of calls: the current / alternate buffers must be back in place")
        if m.device.type != "cuda":
            raise RuntimeError("CUDA graphs need a CUDA device")
        if m._narrow():
            raise RuntimeError("narrow Cyclic grids redo their ghost wrap on the host side: not capturable")
        self.m, self.kernel, self.calls = m, kernel, calls
        k = m.kernels[kernel]
        self.needs_carry = bool(k.get("carry")) and m._carry_kernel == kernel
        m._join_comm()
        torch.cuda.synchronize(m.device)
        ptrs = [t.data_ptr() for t in m.cur if t is not None]
        l0, partial0 = m.launches, dict(m._partial)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            for _ in range(calls):
                m.call(kernel)
            m._join_comm()
        assert ptrs == [t.data_ptr() for t in m.cur if t is not None]
        self.launches_per_replay = m.launches - l0
        # the capture itself executed nothing: the machine's state is what it was, and so is the carried-reduce validity
        m._carry_kernel = kernel if self.needs_carry else None
        self._after = dict(partial=dict(m._partial), carry=kernel if k.get("carry") else None)
        m.launches = l0
        m._partial.clear()
        m._partial.update(partial0)

    def replay(self):
        m = self.m
        if self.needs_carry and m._carry_kernel != self.kernel:
            # something wrote the state since the last call: the carried reduce baked into the graph's first call is stale
            for _ in range(self.calls):
                m.call(self.kernel)
            return
        m._join_comm()
        self.graph.replay()
        m.launches += self.launches_per_replay
        m._partial.update(self._after["partial"])
        m._carry_kernel = self._after["carry"]
