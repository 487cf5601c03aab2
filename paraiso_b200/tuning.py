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
                stage: int = None, prepare: Callable[[Machine], None] = None, fmad: bool = False, steps: int = 30,
                prune_to: int = 0) -> List[dict]:
    """Generate + nvcc + time every candidate; returns [{tuning, ms, cells_per_s, occupancy, smem}] best first.
    `prune_to` > 0: only the candidates the static cost model (costmodel.py: SASS instruction mix, registers, shared
    memory -> predicted cycles per cell; no GPU time) ranks among its best `prune_to` are timed."""
    if prune_to and len(cands) > prune_to:
        from .costmodel import prune
        cands = prune(make_setup, make_om, cands, size, keep=prune_to, kernel=kernel, stage=-1 if stage is None else stage, fmad=fmad)
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


# ---- genetic search: the reference's operators over this backend's genome -----------------------------------------
# Tuning/Genetic.hs encodes an individual as a bit string (CUDA grid x block, one Manifest/Delayed bit per node with an
# AllocationChoice, two sync bits per value) and offers three operators: `mutate` (a geometric number of point
# mutations, :93-108), `cross` (multi-point crossover, :110-131) and `triangulate` (take `left` where it differs from
# `base`, else `right`, :133-137).  The evolution loop itself lives in scripts outside the library (examples-old/GA).
# Here a locus is either a `Tuning` knob with a finite domain or a materialisation gene (kernel, value id).

def _loci(space: Dict[str, Iterable], genes) -> List[tuple]:
    return [("knob", k) for k in space] + [("gene", tuple(g)) for g in genes]


def genome_of(t: Tuning, space: Dict[str, Iterable], genes) -> tuple:
    flips = set(tuple(f) for f in t.mat_flip)
    return tuple(getattr(t, l[1]) if l[0] == "knob" else (l[1] in flips) for l in _loci(space, genes))


def tuning_of(genome: tuple, base: Tuning, space: Dict[str, Iterable], genes) -> Tuning:
    kn, flips = {}, []
    for l, v in zip(_loci(space, genes), genome):
        if l[0] == "knob":
            kn[l[1]] = v
        elif v:
            flips.append(l[1])
    return dataclasses.replace(base, mat_flip=tuple(sorted(flips)), **kn)


def mutate(genome: tuple, space: Dict[str, Iterable], genes, rng) -> tuple:
    """At least one point mutation, one more with probability 1/2 each time (Genetic.hs:93-108)."""
    loci = _loci(space, genes)
    g = list(genome)
    while True:
        i = rng.randrange(len(loci))
        if loci[i][0] == "knob":
            dom = [v for v in space[loci[i][1]] if v != g[i]]
            if dom:
                g[i] = rng.choice(dom)
        else:
            g[i] = not g[i]
        if rng.random() < 0.5:
            return tuple(g)


def cross(a: tuple, b: tuple, rng) -> tuple:
    """Multi-point crossover with a geometric number of cut points (Genetic.hs:110-131)."""
    n = len(a)
    cuts = []
    while True:
        cuts.append(rng.randint(-1, n + 1))
        if rng.random() < 0.5:
            break
    return tuple(a[i] if sum(1 for c in cuts if c < i) % 2 else b[i] for i in range(n))


def triangulate(base: tuple, left: tuple, right: tuple) -> tuple:
    """Genetic.hs:133-137: what `left` changed relative to `base`, on top of `right`."""
    return tuple(l if b != l else r for b, l, r in zip(base, left, right))


def genetic_search(base: Tuning, space: Dict[str, Iterable], genes, evaluate: Callable[[Tuning], float], population: int = 8,
                   generations: int = 4, seed: int = 1, log: Callable[[dict], None] = None, budget_s: float = None) -> dict:
    """Evolve `population` individuals for `generations` rounds; `evaluate(tuning)` returns the cost (ms per step;
    inf for an individual that does not build, run or pass the sanity gate — examples-old/GA/main-kh.cu:20-61 plays that
    role in the reference).  The better half survives; children come from mutate / cross / triangulate of survivors.
    Every distinct genome is evaluated once.  Returns {tuning, ms, evaluated, history}."""
    import random
    import time
    rng = random.Random(seed)
    t_end = time.time() + budget_s if budget_s else None
    cache: Dict[tuple, float] = {}

    def cost(g):
        if g not in cache:
            if t_end and time.time() > t_end:
                return float("inf")
            cache[g] = evaluate(tuning_of(g, base, space, genes))
            if log:
                log(dict(genome=[str(v) for v in g], ms=cache[g]))
        return cache[g]

    start = genome_of(base, space, genes)
    pop = [start]
    while len(pop) < population:
        m = mutate(start, space, genes, rng)
        if m not in pop:
            pop.append(m)
    history = []
    for _gen in range(generations):
        pop.sort(key=cost)
        history.append(cost(pop[0]))
        survivors = pop[:max(2, population // 2)]
        children = []
        tries = 0
        while len(survivors) + len(children) < population and tries < 50 * population:
            tries += 1
            op = rng.random()
            if op < 0.5:
                c = mutate(rng.choice(survivors), space, genes, rng)
            elif op < 0.8:
                c = cross(rng.choice(survivors), rng.choice(survivors), rng)
            else:
                c = triangulate(start, rng.choice(survivors), rng.choice(survivors))
            if c not in survivors and c not in children:
                children.append(c)
        pop = survivors + children
        if t_end and time.time() > t_end:
            break
    pop.sort(key=cost)
    history.append(cost(pop[0]))
    return dict(tuning=tuning_of(pop[0], base, space, genes), ms=cost(pop[0]), evaluated=len(cache), history=history)


def gpu_evaluator(make_setup: Callable[[], Setup], make_om: Callable, size, kernel: str = "proceed",
                  prepare: Callable[[Machine], None] = None, fmad: bool = False, steps: int = 10,
                  gate: dict = None, log: Callable[[dict], None] = None) -> Callable[[Tuning], float]:
    """evaluate() for genetic_search: generate + nvcc + time one individual on the GPU; inf if it fails to build or run,
    or if the sanity gate rejects it.

    `gate` = dict(size, steps, arrays={name: expected ndarray of the local interior}, scalars={name: expected value},
    prepare=callable(machine) or None, rtol=0.0): before an individual is timed it is run for gate["steps"] calls of
    `kernel` on a gate["size"] grid and must reproduce the expected state — bit for bit with rtol = 0 (every schedule of
    a bit-exact build evaluates the same SSA DAG), within rtol for fast_math builds.  This is the role of `isWorking` in
    the reference's benchmark (examples-old/GA/main-kh.cu:20-61,96-102: a wrong individual scores zero)."""
    import numpy as np

    def passes(t: Tuning) -> bool:
        setup = make_setup()
        setup.tuning = t
        desc, so = build_machine(setup, make_om(), tag=f"tune_{make_om().name}_{tag_of(t)}", fmad=fmad)
        m = Machine(desc, so, size=gate["size"])
        (gate.get("prepare") or prepare)(m)
        for _ in range(gate["steps"]):
            m.call(kernel)
        rtol = gate.get("rtol", 0.0)
        for n, want in gate["arrays"].items():
            got = m.get(n)
            if rtol == 0.0:
                if not np.array_equal(got.view(np.uint8), np.ascontiguousarray(want).view(np.uint8)):
                    return False
            elif not np.max(np.abs(got - want)) <= rtol * max(float(np.max(np.abs(want))), 1e-300):
                return False
        for n, want in gate.get("scalars", {}).items():
            got = m.scalar(n)
            if (got != want) if rtol == 0.0 else (abs(got - want) > rtol * abs(want)):
                return False
        return True

    def evaluate(t: Tuning) -> float:
        try:
            if gate is not None and not passes(t):
                if log:
                    log(dict(tuning=dataclasses.asdict(t), rejected="sanity gate"))
                return float("inf")
        except Exception as e:
            if log:
                log(dict(tuning=dataclasses.asdict(t), error=repr(e)[:300]))
            return float("inf")
        r = grid_search(make_setup, make_om, [t], size, kernel=kernel, prepare=prepare, fmad=fmad, steps=steps)[0]
        if log:
            log(r)
        if "ms" not in r:
            return float("inf")
        return r["ms"]
    return evaluate
