"""Measured schedule search for the B200 backend (first step).

The reference tunes generated code with a genetic algorithm over a genome of CUDA grid x block, Manifest/Delayed
choices and __syncthreads placements, scoring each individual by measured cell updates per second with a sanity gate
(Tuning/Genetic.hs:138-172; examples-old/GA/main-kh.cu:20-104).  Here the genome is `generator.native.Tuning`
(skeleton, CTA width, prefetch depth, register row window, chunk height, launch bounds); `grid_search` generates,
compiles and times every candidate on the GPU and returns them best first.  The committed defaults of `Tuning` are the
winners recorded in profiles/r1_life_sweep.txt.  Candidates must pass the same parity checks as the default build
before being adopted (tests/test_gpu_parity.py runs against whatever `Tuning()` says).
"""
from __future__ import annotations

import dataclasses
import itertools
from typing import Callable, Dict, Iterable, List

from .build import build_machine
from .generator.native import Setup, Tuning
from .runtime import Machine


def candidates(space: Dict[str, Iterable], base: Tuning = None) -> List[Tuning]:
    """Cartesian product of the given field -> values space applied to `base`."""
    base = base or Tuning()
    keys = list(space)
    out = []
    for vals in itertools.product(*[list(space[k]) for k in keys]):
        out.append(dataclasses.replace(base, **dict(zip(keys, vals))))
    return out


def tag_of(t: Tuning) -> str:
    import hashlib
    parts = []
    for f in dataclasses.fields(t):
        v = getattr(t, f.name)
        if f.name == "mat_flip":
            v = hashlib.sha1(repr(sorted(v)).encode()).hexdigest()[:8] if v else "0"
        parts.append(f"{f.name[:2]}{v}")
    return "_".join(parts).replace(" ", "")


def measure(m: Machine, kernel: str, steps: int = 30, warmup: int = 5, stage: int = None) -> float:
    """Milliseconds per call of `kernel` (or of one of its stages), CUDA events on the current stream."""
    import torch
    fn = (lambda: m.call_stage(kernel, stage)) if stage is not None else (lambda: m.call(kernel))
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(m.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(m.device)
    return e0.elapsed_time(e1) / steps


def grid_search(make_setup: Callable[[], Setup], make_om: Callable, cands: List[Tuning], size, kernel: str = "proceed",
                stage: int = None, prepare: Callable[[Machine], None] = None, fmad: bool = False, steps: int = 30) -> List[dict]:
    """Generate + nvcc + time every candidate; returns [{tuning, ms, cells_per_s, occupancy, smem}] best first."""
    results = []
    for t in cands:
        setup = make_setup()
        setup.tuning = t
        try:
            desc, so = build_machine(setup, make_om(), tag=f"tune_{make_om().name}_{tag_of(t)}", fmad=fmad)
            m = Machine(desc, so, size=size)
            if prepare:
                prepare(m)
            ms = measure(m, kernel, steps=steps, stage=stage)
            st = m.kernels[kernel]["stages"][stage if stage is not None else -1]
            results.append(dict(tuning=dataclasses.asdict(t), ms=ms, cells_per_s=m.nx * m.nyl / ms * 1e3,
                                occupancy=getattr(m.lib, st["symbol"] + "_occupancy")(), smem=st["smem"]))
        except Exception as e:   # a candidate that does not build or launch is simply not adopted
            results.append(dict(tuning=dataclasses.asdict(t), error=repr(e)[:300]))
    return sorted(results, key=lambda r: r.get("ms", float("inf")))


def local_search(make_setup: Callable[[], Setup], make_om: Callable, size, kernel: str = "proceed",
                 prepare: Callable[[Machine], None] = None, fmad: bool = False, steps: int = 10, passes: int = 1,
                 min_gain: float = 0.005, budget_s: float = 600.0, max_recompute_cost: int = 400,
                 log: Callable[[dict], None] = None) -> dict:
    """Per-node materialise / recompute search — the Manifest/Delayed genes of the reference's genome
    (Tuning/Genetic.hs:150-160; DecideAllocation.hs:41-43 takes the choice from the annotation the GA wrote).

    Every value that is read through a Shift is a gene: kept in a shared-memory ring (computed once per cell) or
    recomputed at each cursor.  The threshold rule gives the start individual; this is a first-improvement coordinate
    descent over single-gene flips, each candidate generated, compiled and timed on the GPU like `grid_search` does
    (a candidate whose rings do not fit, or that does not build, is skipped).  Results are unaffected by the schedule:
    every node is still evaluated from the same SSA expression (tests/test_emulated.py runs a flipped schedule
    against the oracle)."""
    import time
    t_end = time.time() + budget_s
    base = make_setup().tuning

    def run(flips):
        t = dataclasses.replace(base, mat_flip=tuple(sorted(flips)))
        r = grid_search(make_setup, make_om, [t], size, kernel=kernel, prepare=prepare, fmad=fmad, steps=steps)[0]
        r["mat_flip"] = [list(f) for f in sorted(flips)]
        if log:
            log(r)
        return r

    best = run(set(base.mat_flip))
    if "ms" not in best:
        raise RuntimeError(f"the start individual does not run: {best.get('error')}")
    flips = set(base.mat_flip)
    setup = make_setup()
    from .generator.b200.emit import describe_only
    # (un-materialising a value whose expression tree has thousands of operations only produces an enormous kernel)
    genes = [(c["kernel"], c["vid"], c["cost"], c["chosen"]) for c in describe_only(setup, make_om(), kernel)
             if not (c["chosen"] and c["cost"] > max_recompute_cost)]
    # cheap materialised values first (recomputing them may free a ring), then expensive recomputed ones
    genes.sort(key=lambda g: (not g[3], g[2] if g[3] else -g[2]))
    for _ in range(passes):
        improved = False
        for (kn, vid, _cost, _chosen) in genes:
            if time.time() > t_end:
                break
            trial = set(flips) ^ {(kn, vid)}
            r = run(trial)
            if "ms" in r and r["ms"] < best["ms"] * (1.0 - min_gain):
                best, flips, improved = r, trial, True
        if not improved:
            break
    return best
