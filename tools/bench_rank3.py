"""Rank-3 Life (26 neighbours, Cyclic) on the GPU: ms per step, Gcell/s and algorithmic GB/s (8 B per cell update)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from paraiso_b200.build import build_machine  # noqa: E402
from paraiso_b200.examples.rank3 import life3d_om  # noqa: E402
from paraiso_b200.generator.native import Setup, Tuning  # noqa: E402
from paraiso_b200.runtime import Machine  # noqa: E402
from paraiso_b200.tuning import measure  # noqa: E402

def heat(n, z):
    from paraiso_b200.examples.rank3 import heat3d_om
    size = (n, n, n)
    setup = Setup(local_size=size, boundary=("Cyclic", "Cyclic", "Cyclic"))
    setup.tuning = Tuning.from_env(setup.tuning)       # OM_* overrides (sweeps)
    setup.tuning.planes_per_cta = z
    desc, so = build_machine(setup, heat3d_om(), tag=f"Heat3_bench_z{z}")
    m = Machine(desc, so, size=size)
    m.set("u", torch.rand((n, n, n), dtype=torch.float32).numpy())
    ms = measure(m, "proceed", steps=20, warmup=3)
    st = m.kernels["proceed"]["stages"][0]
    print(json.dumps(dict(program="heat3d float 7-point", planes_per_cta=z, size=size, ms=ms, Gcell_per_s=n ** 3 / ms / 1e6, alg_GB_per_s=8 * n ** 3 / ms / 1e6,
                          rings=st["rings"], smem=st["smem"], V=st["V"], NT=st["NT"])))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    z = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    if len(sys.argv) > 2 and sys.argv[2] == "heat":
        heat(n, z)
        sys.exit(0)
    size = (n, n, n)
    setup = Setup(local_size=size, boundary=("Cyclic", "Cyclic", "Cyclic"))
    setup.tuning.planes_per_cta = z
    desc, so = build_machine(setup, life3d_om(), tag=f"Life3_bench_z{z}")
    m = Machine(desc, so, size=size)
    g = torch.Generator(device="cuda").manual_seed(1)
    init = (torch.rand((n, n, n), device="cuda", generator=g) < 0.3).to(torch.int32)
    m.set("cell", init.cpu().numpy())
    ms = measure(m, "proceed", steps=20, warmup=3)
    cells = n ** 3
    st = m.kernels["proceed"]["stages"][0]
    print(json.dumps(dict(program="life3d 26 neighbours", planes_per_cta=z, size=size, ms=ms, Gcell_per_s=cells / ms / 1e6, alg_GB_per_s=8 * cells / ms / 1e6, population=int(m.scalar("population")),
                          rings=st["rings"], smem=st["smem"], V=st["V"], NT=st["NT"])))
