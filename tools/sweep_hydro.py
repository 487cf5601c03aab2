"""Hydro proceed: schedule search on the GPU box (paraiso_b200.tuning.grid_search)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.examples.hydro import hydro_om, hydro_setup  # noqa: E402
from paraiso_b200.machines import hydro_set_params  # noqa: E402
from paraiso_b200.tuning import candidates, grid_search  # noqa: E402

if __name__ == "__main__":
    fast = len(sys.argv) < 2 or sys.argv[1] != "exact"
    space = json.loads(sys.argv[2]) if len(sys.argv) > 2 else dict(threads_heavy=[128, 192, 256, 320])
    size = (4096, 4096)

    flips = tuple(tuple(f) for f in json.loads(sys.argv[3])) if len(sys.argv) > 3 else ()

    def mk():
        s = hydro_setup(fast=fast)
        s.tuning.mat_flip = flips
        return s

    def prepare(m):
        hydro_set_params(m, size)
        m.call("init")
    for r in grid_search(mk, lambda: hydro_om("master"), candidates(space, mk().tuning), size, prepare=prepare, fmad=fast, steps=10):
        if "ms" in r:
            r["Gcell_per_s"] = size[0] * size[1] / r["ms"] / 1e6
        print(json.dumps(r), flush=True)
