"""Hydro flux stage with and without the peeled pipeline-fill loop (Tuning.peel_fill), both builds, same box.  `--prebuild` compiles here."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.build import build_machine  # noqa: E402
from paraiso_b200.examples.hydro import hydro_om, hydro_setup  # noqa: E402
from paraiso_b200.machines import hydro_set_params  # noqa: E402

SIZE = (4096, 4096)
if __name__ == "__main__":
    for fast in (True, False):
        for peel in (True, False):
            s = hydro_setup(fast=fast)
            s.tuning.peel_fill = peel
            desc, so = build_machine(s, hydro_om("master"), tag=f"variant_Hydro_peel{int(peel)}_{'fast' if fast else 'exact'}", fmad=fast)
            if "--prebuild" in sys.argv:
                continue
            import torch
            from paraiso_b200.runtime import Machine
            from paraiso_b200.tuning import measure
            m = Machine(desc, so, size=SIZE)
            hydro_set_params(m, SIZE)
            m.call("init")
            for _ in range(20):
                m.call("proceed")
            ms = sorted(measure(m, "proceed", steps=20) for _ in range(5))
            print(json.dumps(dict(build="fast" if fast else "exact", peel_fill=peel, ms_best=ms[0], ms_median=ms[2],
                                  Gcell_per_s=SIZE[0] * SIZE[1] / ms[2] / 1e6)), flush=True)
            del m
