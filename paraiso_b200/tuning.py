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
    return "_".join(f"{f.name[:2]}{getattr(t, f.name)}" for f in dataclasses.fields(t)).replace(" ", "")


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
