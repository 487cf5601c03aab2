"""Static cost model for schedule candidates: rank them after nvcc, before any GPU time is spent.

The reference's tuner scores every individual by running it (Tuning/Genetic.hs:43,172 regenerate + the benchmark of
examples-old/GA/main-kh.cu:63-104); `tuning.grid_search` / `local_search` / `genetic_search` do the same here.  This module
prunes that: a candidate is generated and compiled (no GPU needed — nvcc cross-compiles), and its dominant stage is
scored from three things the build leaves behind:

  * the SASS of the stage's row loop (`cuobjdump -sass`): issue slots per warp-row and how many of them go to the FP64
    pipe, which on sm_100a accepts one warp instruction every two cycles per scheduler (64 FP64 lanes per SM);
  * registers per thread (`ptxas -v`) and dynamic shared memory -> resident CTAs per SM -> warps per scheduler;
  * the schedule's geometric overhead: halo columns a CTA computes but does not store, warm-up rows per chunk.

    cycles per warp-row  =  max(issue slots, 2 * FP64 instructions) / u(w)
    u(w)                 =  w / (w + W_HALF)          w = resident warps per scheduler
    issue slots         +=  SPILL_SLOTS per LDL / STL (register-capped candidates spill)
    CTAs of W warps, W % 4 != 0, load the four schedulers unevenly (ceil(W/4) vs W/4 warps per CTA):
                         *=  1 + RAGGED * (ceil(W/4) / (W/4) - 1)

`u` is the fraction of issue slots a scheduler fills with `w` warps of one long dependent FP64 chain each; W_HALF = 1.6
reproduces the measured B200 points (three warps: 62 % issue-active in profiles/r1i_hydro_fast_ncu.txt); SPILL_SLOTS and
RAGGED are fitted on the twelve measured CTA shapes of profiles/r1g_sweep_fast.jsonl.  On those the model's three best
candidates are the three measured best, the rank correlation is 0.79 and the worst prediction is 24 % off
(profiles/r1j_costmodel.json; tests/test_costmodel.py re-checks the formula against the recorded points without compiling).  The model is for *heavy* (compute-bound) stages; streaming stages are
bandwidth-bound and all their candidates cost the same to it, so `prune` leaves those to the measured search.
"""
from __future__ import annotations

import dataclasses
import os
import re
import shutil
import subprocess
from typing import Callable, Dict, List, Optional

from .build import build_machine
from .generator.native import Setup, Tuning

W_HALF = 1.6
SPILL_SLOTS = 8          # issue slots charged per local-memory instruction (LDL / STL) in the row loop
RAGGED = 0.5             # weight of the CTA-local round-robin imbalance for CTAs whose warp count is not a multiple of 4
SMS, SM_CLOCK_HZ = 148, 1.965e9          # B200; only used to turn cycles into milliseconds
REGS_PER_SM, SMEM_PER_SM, MAX_WARPS_PER_SM, SMEM_CTA_RESERVE = 65536, 227 * 1024, 64, 1024
FP64_OPS = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def sass_functions(so: str) -> Dict[str, List[tuple]]:
    """{mangled function name: [(address, opcode, text)]} from `cuobjdump -sass`."""
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-sass", so], capture_output=True, text=True, check=True).stdout
    funcs: Dict[str, List[tuple]] = {}
    cur = None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur is not None:
            text = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
            cur.append((int(m.group(1), 16), text.split()[0], text))
    return funcs


def loop_stats(insts: List[tuple]) -> dict:
    """Instruction mix of the largest loop (longest span closed by a backward branch) of one function."""
    best = (0, 0)
    for addr, op, text in insts:
        if op.startswith("BRA"):
            t = re.findall(r"0x([0-9a-f]+)", text)
            if t and int(t[-1], 16) < addr and addr - int(t[-1], 16) > best[1] - best[0]:
                best = (int(t[-1], 16), addr)
    body = [op for addr, op, _ in insts if best[0] <= addr <= best[1]] if best[1] else [op for _, op, _ in insts]
    fp64 = sum(1 for op in body if op.split(".")[0] in FP64_OPS)
    return dict(instructions=len(body), fp64=fp64, local=sum(1 for op in body if op[:3] in ("LDL", "STL")), mufu=sum(1 for op in body if op.startswith("MUFU")),
                smem=sum(1 for op in body if op.split(".")[0] in ("LDS", "STS")), barriers=sum(1 for op in body if op.startswith("BAR")))


def ptxas_registers(log_path: str, symbol: str) -> Optional[int]:
    """Registers per thread of `<symbol>_kernel` from the `ptxas -v` log that build_machine(verbose=True) writes."""
    if not os.path.exists(log_path):
        return None
    with open(log_path) as f:
        text = f.read()
    m = re.search(rf"Compiling entry function '\w*{re.escape(symbol)}_kernel\w*'.*?Used (\d+) registers", text, flags=re.S)
    return int(m.group(1)) if m else None


def resident_ctas(regs: int, threads: int, smem: int, min_blocks: int = 0) -> int:
    """CTAs of this shape one SM holds (register file in 8-register / warp granules, shared memory, warp slots)."""
    regs_alloc = -(-regs // 8) * 8
    by_regs = REGS_PER_SM // (regs_alloc * threads)
    by_smem = SMEM_PER_SM // (smem + SMEM_CTA_RESERVE) if smem else 32
    return max(0, min(by_regs, by_smem, MAX_WARPS_PER_SM * 32 // threads, 32))


@dataclasses.dataclass
class Estimate:
    tuning: dict
    symbol: str
    instructions: int           # issue slots per warp-row of the stage's row loop
    fp64: int
    local: int                  # LDL / STL among them (spills)
    registers: int
    smem: int
    threads: int
    ctas_per_sm: int
    warps_per_scheduler: float
    overhead: float             # (computed columns / stored columns) * (rows incl. warm-up / rows)
    cycles_per_cell: float
    ms: Optional[float] = None  # for `size`, when given

    def as_dict(self) -> dict:
        return dataclasses.asdict(self)


def cycles_per_cell(instructions: int, fp64: int, local: int, threads: int, ctas: int, cells_per_thread: int, overhead: float) -> float:
    """The model's formula (see the module docstring): scheduler cycles per stored cell."""
    warps = threads // 32
    w = ctas * warps / 4.0
    slots = max(instructions + SPILL_SLOTS * local, 2 * fp64)
    util = w / (w + W_HALF) if w > 0 else 1e-9
    ragged = 1.0 + RAGGED * (-(-warps // 4) / (warps / 4.0) - 1.0)
    return slots / util * ragged / (32 * cells_per_thread) * overhead


def estimate_stage(desc: dict, so: str, kernel: str = "proceed", stage: int = -1, size=None, tuning: Tuning = None,
                   funcs: Dict[str, List[tuple]] = None) -> Estimate:
    """Score one built machine's stage (default: the last stage of `proceed`)."""
    k = [k for k in desc["kernels"] if k["name"] == kernel][0]
    st = k["stages"][stage]
    funcs = funcs or sass_functions(so)
    name = [f for f in funcs if st["symbol"] + "_kernel" in f]
    if not name:
        raise KeyError(st["symbol"])
    ls = loop_stats(funcs[name[0]])
    log = os.path.join(os.path.dirname(so), "ptxas_" + os.path.basename(so)[len("libom_"):-len(".so")] + ".log")
    regs = ptxas_registers(log, st["symbol"]) or 128
    nt = st["NT"]
    ctas = resident_ctas(regs, nt, st["smem"])
    w = ctas * nt / 32 / 4.0
    cols = nt * st["V"]
    over = cols / max(1, st["w_out"])
    rows = None
    if size is not None:
        # one wave of equal CTAs (runtime.Machine._geom): chunks = SMs * CTAs/SM / strips
        strips = max(1, -(-size[0] // st["w_out"]))
        chunks = max(1, min((SMS * max(ctas, 1)) // strips, size[1] // max(32, 8 * (st["warmup"] + 2))))
        rows = -(-size[1] // chunks)
        over *= (rows + st["warmup"]) / rows
    cyc_cell = cycles_per_cell(ls["instructions"], ls["fp64"], ls["local"], nt, ctas, st["V"], over)
    est = Estimate(tuning=dataclasses.asdict(tuning) if tuning else {}, symbol=st["symbol"], instructions=ls["instructions"],
                   fp64=ls["fp64"], local=ls["local"], registers=regs, smem=st["smem"], threads=nt, ctas_per_sm=ctas, warps_per_scheduler=w,
                   overhead=over, cycles_per_cell=cyc_cell)
    if size is not None:
        est.ms = size[0] * size[1] * cyc_cell / (SMS * 4) / SM_CLOCK_HZ * 1e3
    return est


def estimate(make_setup: Callable[[], Setup], make_om: Callable, cands: List[Tuning], size, kernel: str = "proceed",
             stage: int = -1, fmad: bool = False, tag_prefix: str = "cost") -> List[Estimate]:
    """Generate + nvcc every candidate (no GPU) and return the estimates, cheapest first.  A candidate that does not
    build, or whose CTA does not fit an SM, is dropped."""
    from .tuning import tag_of
    out = []
    for t in cands:
        setup = make_setup()
        setup.tuning = t
        try:
            desc, so = build_machine(setup, make_om(), tag=f"{tag_prefix}_{make_om().name}_{tag_of(t)}", fmad=fmad, verbose=True)
            e = estimate_stage(desc, so, kernel=kernel, stage=stage, size=size, tuning=t)
        except Exception:
            continue
        if e.ctas_per_sm >= 1:
            out.append(e)
    return sorted(out, key=lambda e: e.cycles_per_cell)


def prune(make_setup: Callable[[], Setup], make_om: Callable, cands: List[Tuning], size, keep: int = 4, **kw) -> List[Tuning]:
    """The `keep` candidates the model likes best — what `tuning.grid_search` should still time on the GPU."""
    best = estimate(make_setup, make_om, cands, size, **kw)[:keep]
    by_key = {repr(sorted(dataclasses.asdict(t).items())): t for t in cands}
    return [by_key[repr(sorted(e.tuning.items()))] for e in best]


def spearman(xs: List[float], ys: List[float]) -> float:
    def ranks(v):
        order = sorted(range(len(v)), key=lambda i: v[i])
        r = [0.0] * len(v)
        for k, i in enumerate(order):
            r[i] = float(k)
        return r
    rx, ry = ranks(xs), ranks(ys)
    n = len(xs)
    mx, my = sum(rx) / n, sum(ry) / n
    cov = sum((a - mx) * (b - my) for a, b in zip(rx, ry))
    vx, vy = sum((a - mx) ** 2 for a in rx), sum((b - my) ** 2 for b in ry)
    return cov / (vx * vy) ** 0.5 if vx and vy else 0.0
