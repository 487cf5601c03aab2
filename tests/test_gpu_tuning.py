"""SURVEY §8 f1 on the device: individuals of the schedule search — a flipped Manifest/Delayed gene (Tuning.mat_flip),
other CTA shapes, the children of one genetic_search generation — are generated, compiled, run on the B200 and must stay
bit-identical to the oracle (the sanity gate inside tuning.gpu_evaluator; the reference's GA gates with `isWorking`,
examples-old/GA/main-kh.cu:20-61)."""
import dataclasses

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NAMES = ("density", "velocity0", "velocity1", "pressure")
SIZE, STEPS = (96, 80), 4
SPACE = dict(threads_heavy=[128, 256], prefetch_rows=[1, 2])


def _setup():
    from paraiso_b200.examples.hydro import hydro_setup
    return hydro_setup(SIZE)


def _om():
    from paraiso_b200.examples.hydro import hydro_om
    return hydro_om("master")


def _prepare(m):
    from paraiso_b200.machines import hydro_set_params
    hydro_set_params(m, (m.nx, m.ny))
    m.call("init")
    if getattr(_prepare, "ic", None) is not None and (m.nx, m.ny) == SIZE:
        for n in NAMES:
            m.set(n, _prepare.ic[n], with_margin=True)


def _oracle_gate():
    from oracle.cpu import OracleMachine
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    o = OracleMachine(hydro_setup(SIZE), hydro_om("master"), openmp=True, opt="-O2")
    for k, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / SIZE[0], dR1=1.0 / SIZE[1]).items():
        o.scalar(k)[0] = v
    o.call("init")
    _prepare.ic = {n: o.array(n).copy() for n in NAMES}      # share the initial condition (CUDA sin vs libm sin)
    for _ in range(STEPS):
        o.call("proceed")
    return dict(size=SIZE, steps=STEPS, arrays={n: o.interior(n).copy() for n in NAMES}, scalars={"time": float(o.scalar("time")[0])})


def _genes():
    from paraiso_b200.generator.b200.emit import describe_only
    return [(c["kernel"], c["vid"]) for c in describe_only(_setup(), _om(), "proceed") if c["cost"] <= 40][:6]


def first_population():
    """(base tuning, the genomes genetic_search evaluates first with seed 1) — prebuilt by __graft_entry__.build()."""
    import random
    from paraiso_b200 import tuning
    base, genes = _setup().tuning, _genes()
    rng = random.Random(1)
    start = tuning.genome_of(base, SPACE, genes)
    pop = [start]
    while len(pop) < 4:
        g = tuning.mutate(start, SPACE, genes, rng)
        if g not in pop:
            pop.append(g)
    return base, genes, [tuning.tuning_of(g, base, SPACE, genes) for g in pop]


def prebuild():
    from paraiso_b200.build import build_machine
    from paraiso_b200.tuning import tag_of
    base, genes, pop = first_population()
    flipped = dataclasses.replace(base, mat_flip=(genes[0],))
    for t in pop + [flipped]:
        s = _setup()
        s.tuning = t
        build_machine(s, _om(), tag=f"tune_{_om().name}_{tag_of(t)}")


def test_flipped_gene_individual_is_bit_identical_to_the_oracle():
    from paraiso_b200 import tuning
    gate = _oracle_gate()
    base, genes, _pop = first_population()
    log = []
    ev = tuning.gpu_evaluator(_setup, _om, SIZE, prepare=_prepare, steps=3, gate=gate, log=log.append)
    assert np.isfinite(ev(base))
    assert np.isfinite(ev(dataclasses.replace(base, mat_flip=(genes[0],)))), log
    # negative control: the same individual against a gate that expects something else scores inf
    wrong = dict(gate, arrays=dict(gate["arrays"], density=gate["arrays"]["density"] + 1e-9))
    assert ev.__closure__ is not None
    assert tuning.gpu_evaluator(_setup, _om, SIZE, prepare=_prepare, steps=3, gate=wrong)(base) == float("inf")


def test_one_genetic_generation_stays_in_parity():
    from paraiso_b200 import tuning
    gate = _oracle_gate()
    base, genes, _pop = first_population()
    log = []
    ev = tuning.gpu_evaluator(_setup, _om, SIZE, prepare=_prepare, steps=3, gate=gate, log=log.append)
    best = tuning.genetic_search(base, SPACE, genes, ev, population=4, generations=1, seed=1)
    assert np.isfinite(best["ms"]) and best["evaluated"] >= 4
    rejected = [r for r in log if "rejected" in r]
    assert not rejected, rejected      # every schedule evaluates the same SSA DAG: none may leave parity
    assert sum(1 for r in log if "ms" in r) >= 3
