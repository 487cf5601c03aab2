"""Hydro bit-exact build: where the time of the guarded branch-free division goes.  Variants of Tuning.exact_guard / exact_divsqrt
at the default CTA shape; `--prebuild` compiles them here, without it they are timed on the GPU (4096^2, 10 steps from init)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.build import build_machine  # noqa: E402
from paraiso_b200.examples.hydro import hydro_om, hydro_setup  # noqa: E402
from paraiso_b200.machines import hydro_set_params  # noqa: E402

VARIANTS = [("redo", "newton", 128, 3), ("flag", "newton", 128, 3), ("none", "newton", 128, 3), ("redo", "ieee", 256, 0), ("redo", "ieee", 128, 3),
            ("redo", "newton", 256, 2), ("flag", "newton", 256, 2), ("none", "newton", 256, 2)]
SIZE = (4096, 4096)

if __name__ == "__main__":
    for guard, div, nt, minb in VARIANTS:
        s = hydro_setup()
        s.tuning.exact_guard, s.tuning.exact_divsqrt, s.tuning.threads_heavy, s.tuning.min_blocks_heavy = guard, div, nt, minb
        desc, so = build_machine(s, hydro_om("master"), tag=f"variant_Hydro_{guard}_{div}_{nt}_{minb}", verbose=True)
        if "--prebuild" in sys.argv:
            continue
        import torch
        from paraiso_b200.runtime import Machine
        from paraiso_b200.tuning import measure
        m = Machine(desc, so, size=SIZE)
        hydro_set_params(m, SIZE)
        m.call("init")
        ms = measure(m, "proceed", steps=10)
        for _ in range(300):
            m.call("proceed")
        ms_late = measure(m, "proceed", steps=10)
        torch.cuda.synchronize()
        print(json.dumps(dict(guard=guard, divsqrt=div, threads=nt, min_blocks=minb, ms=ms, Gcell_per_s=SIZE[0] * SIZE[1] / ms / 1e6,
                              ms_after_300_steps=ms_late, slow_cells=m.slow_path_cells() if guard == "redo" and div == "newton" else None)), flush=True)
        del m
