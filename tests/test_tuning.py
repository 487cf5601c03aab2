"""The schedule-search machinery that does not need a GPU: the genome mapping, the three operators of the reference's
Tuning/Genetic.hs (mutate / cross / triangulate) and the evolution loop on a synthetic cost function."""
import random

from paraiso_b200.generator.native import Tuning
from paraiso_b200.tuning import candidates, cross, genetic_search, genome_of, mutate, tag_of, triangulate, tuning_of

SPACE = {"threads_heavy": [96, 128, 192, 256], "min_blocks_heavy": [0, 2, 3, 4], "carry_reduces": [False, True]}
GENES = [("proceed", 322), ("proceed", 362), ("proceed", 1906)]


def test_genome_round_trip_and_tags():
    t = Tuning(threads_heavy=128, min_blocks_heavy=3, mat_flip=(("proceed", 362),))
    g = genome_of(t, SPACE, GENES)
    assert g == (128, 3, False, False, True, False)
    assert tuning_of(g, Tuning(), SPACE, GENES) == t
    assert len(candidates({"prefetch_rows": [1, 2], "threads_light": [64, 128, 256]})) == 6
    assert tag_of(t) != tag_of(Tuning(threads_heavy=128, min_blocks_heavy=3))          # flips are part of the build tag


def test_operators_follow_the_reference_semantics():
    rng = random.Random(3)
    base = genome_of(Tuning(), SPACE, GENES)
    for _ in range(200):
        m = mutate(base, SPACE, GENES, rng)
        assert len(m) == len(base)
        for locus, v in zip(list(SPACE) + GENES, m):
            assert (v in SPACE[locus]) if locus in SPACE else isinstance(v, bool)
    assert any(mutate(base, SPACE, GENES, rng) != base for _ in range(20))
    a, b = (96, 0, False, False, False, False), (256, 4, True, True, True, True)
    for _ in range(100):
        c = cross(a, b, rng)
        assert all(ci in (ai, bi) for ci, ai, bi in zip(c, a, b))
    # triangulate: what `left` changed relative to `base`, applied on top of `right` (Genetic.hs:133-137)
    assert triangulate((1, 1, 1, 1), (1, 2, 1, 3), (5, 6, 7, 8)) == (5, 2, 7, 3)


def test_evolution_finds_the_optimum_of_a_synthetic_cost_and_evaluates_each_genome_once():
    calls = []

    def evaluate(t: Tuning) -> float:
        calls.append(t)
        if t.threads_heavy == 192:
            return float("inf")                        # an individual that "does not build" is never adopted
        return (abs(t.threads_heavy - 128) / 32 + abs(t.min_blocks_heavy - 3) + (0 if t.carry_reduces else 0.5) +
                (0.25 if ("proceed", 362) in t.mat_flip else 0) + (0 if ("proceed", 1906) in t.mat_flip else 0.3) + 1.0)
    best = genetic_search(Tuning(), SPACE, GENES, evaluate, population=10, generations=12, seed=7)
    assert best["ms"] == 1.0
    assert (best["tuning"].threads_heavy, best["tuning"].min_blocks_heavy, best["tuning"].carry_reduces) == (128, 3, True)
    assert best["tuning"].mat_flip == (("proceed", 1906),)
    assert best["history"] == sorted(best["history"], reverse=True)      # elitism: the best cost never gets worse
    keys = [(t.threads_heavy, t.min_blocks_heavy, t.carry_reduces, t.mat_flip) for t in calls]
    assert len(keys) == len(set(keys)) == best["evaluated"]
