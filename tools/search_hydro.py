"""Hydro proceed: per-node materialise / recompute search on the GPU box (paraiso_b200.tuning.local_search)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.examples.hydro import hydro_om, hydro_setup  # noqa: E402
from paraiso_b200.machines import hydro_set_params  # noqa: E402
from paraiso_b200.tuning import local_search  # noqa: E402

if __name__ == "__main__":
    fast = len(sys.argv) < 2 or sys.argv[1] != "exact"
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
    size = (4096, 4096)

    def prepare(m):
        hydro_set_params(m, size)
        m.call("init")

    def log(r):
        r = dict(r)
        r.pop("tuning", None)
        if "ms" in r:
            r["Gcell_per_s"] = size[0] * size[1] / r["ms"] / 1e6
        print(json.dumps(r), flush=True)
    def mk():
        from paraiso_b200.generator.native import Tuning
        s = hydro_setup(fast=fast)
        s.tuning = Tuning.from_env(s.tuning)        # OM_* overrides of the start individual's knobs
        return s
    best = local_search(mk, lambda: hydro_om("master"), size, prepare=prepare, fmad=fast,
                        steps=10, passes=2, budget_s=budget, log=log)
    print("BEST", json.dumps(dict(mat_flip=best["mat_flip"], ms=best["ms"])), flush=True)
