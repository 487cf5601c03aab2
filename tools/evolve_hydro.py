"""Hydro proceed: genetic search over CTA shape, launch bounds, prefetch, carried reduce and the per-node
materialisation genes (paraiso_b200.tuning.genetic_search), every individual built and timed on the GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.examples.hydro import hydro_om, hydro_setup  # noqa: E402
from paraiso_b200.generator.b200.emit import describe_only  # noqa: E402
from paraiso_b200.machines import hydro_set_params  # noqa: E402
from paraiso_b200.tuning import genetic_search, gpu_evaluator  # noqa: E402

if __name__ == "__main__":
    fast = len(sys.argv) < 2 or sys.argv[1] != "exact"
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 240.0
    size = (4096, 4096)
    mk = lambda: hydro_setup(fast=fast)
    om = lambda: hydro_om("master")
    space = {"threads_heavy": [96, 128, 160, 192, 256], "min_blocks_heavy": [0, 2, 3, 4], "prefetch_rows": [1, 2],
             "carry_reduces": [False, True], "direct_prefetch": [False, True]}
    genes = [(c["kernel"], c["vid"]) for c in describe_only(mk(), om(), "proceed") if c["cost"] <= 400]

    def prepare(m):
        hydro_set_params(m, size)
        m.call("init")
    log = lambda r: print(json.dumps(r), flush=True)
    best = genetic_search(mk().tuning, space, genes, gpu_evaluator(mk, om, size, prepare=prepare, fmad=fast, steps=8),
                          population=8, generations=6, seed=20261017, log=log, budget_s=budget)
    import dataclasses
    print("BEST", json.dumps(dict(tuning=dataclasses.asdict(best["tuning"]), ms=best["ms"], evaluated=best["evaluated"], history=best["history"])), flush=True)
