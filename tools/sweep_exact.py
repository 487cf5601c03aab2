"""CTA-shape sweep of Hydro's bit-exact build (branch-free IEEE-correct division / sqrt): `--prebuild` compiles every candidate
here (nvcc cross-compiles without a GPU; the libraries travel with the snapshot), without it the candidates are timed on
the GPU, each one behind the parity gate of tuning.gpu_evaluator (bit-identical to the default build's state)."""
import dataclasses
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.build import build_machine  # noqa: E402
from paraiso_b200.examples.hydro import hydro_om, hydro_setup  # noqa: E402
from paraiso_b200.machines import hydro_set_params  # noqa: E402
from paraiso_b200.tuning import candidates, gpu_evaluator, tag_of  # noqa: E402

FAST = "--fast" in sys.argv
SPACE = dict(threads_heavy=[128, 192, 256], min_blocks_heavy=[0, 2, 3, 4], direct_prefetch=[False, True], carry_reduces=[False, True])
if FAST:
    SPACE = dict(threads_heavy=[96, 128, 160], min_blocks_heavy=[3, 4, 5], prefetch_rows=[1, 2])
for a in sys.argv[1:]:
    if a.startswith("{"):
        SPACE = json.loads(a)
SIZE = (4096, 4096)


def mk():
    return hydro_setup(fast=FAST)


def prepare(m):
    hydro_set_params(m, (m.nx, m.ny))
    m.call("init")


if __name__ == "__main__":
    cands = candidates(SPACE, mk().tuning)
    if "--prebuild" in sys.argv:
        for t in cands:
            s = mk()
            s.tuning = t
            try:
                build_machine(s, hydro_om("master"), tag=f"tune_Hydro_{tag_of(t)}", fmad=FAST)
            except Exception as e:
                print("build failed", tag_of(t), repr(e)[:200])
        sys.exit(0)
    from paraiso_b200.machines import hydro_machine
    ref = hydro_machine((512, 384), fast=FAST)
    prepare(ref)
    for _ in range(4):
        ref.call("proceed")
    gate = dict(size=(512, 384), steps=4, arrays={n: ref.get(n) for n in ("density", "velocity0", "velocity1", "pressure")},
                scalars={"time": ref.scalar("time")})
    ev = gpu_evaluator(mk, lambda: hydro_om("master"), SIZE, prepare=prepare, fmad=FAST, steps=10, gate=gate,
                       log=lambda r: print(json.dumps({k: v for k, v in r.items()}), flush=True))
    for t in cands:
        ev(t)
