"""Life 16384^2: CTA width (Tuning.threads_light = columns per CTA / 4) and resident CTAs, same box.  `--prebuild` compiles here."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.build import build_machine  # noqa: E402
from paraiso_b200.examples.life import life_om, life_setup  # noqa: E402

CASES = [(128, 9, 3, 20), (256, 4, 3, 20), (256, 4, 3, 32), (256, 4, 4, 24), (512, 2, 3, 24), (512, 2, 4, 32), (64, 18, 3, 20), (192, 6, 3, 20)]
size = (16384, 16384)
if __name__ == "__main__":
    built = []
    for nt, minb, pf, cr in CASES:
        s = life_setup("master")
        s.tuning.threads_light, s.tuning.min_blocks, s.tuning.prefetch_rows, s.tuning.chunk_rows_light = nt, minb, pf, cr
        built.append(((nt, minb, pf, cr), build_machine(s, life_om("master"), tag=f"variant_Life_nt{nt}_b{minb}_pf{pf}_c{cr}", verbose=True)))
    if "--prebuild" in sys.argv:
        sys.exit(0)
    import torch
    from paraiso_b200.machines import life_seed
    from paraiso_b200.runtime import Machine
    from paraiso_b200.tuning import measure
    seed = torch.from_numpy(life_seed(size[0], 0, size[1])).pin_memory()
    for (nt, minb, pf, cr), (desc, so) in built:
        m = Machine(desc, so, size=size)
        m.call("init")
        m.set_from_host("cell", seed)
        st = m.kernels["proceed"]["stages"][0]
        ms = min(measure(m, "proceed", steps=20, stage=0) for _ in range(3))
        print(json.dumps(dict(threads=nt, min_blocks=minb, prefetch_rows=pf, chunk_rows=cr, occupancy=getattr(m.lib, st["symbol"] + "_occupancy")(),
                              chunks=m._geom(st).nchunks, ms=ms, GBs=2 * 4 * size[0] * size[1] / ms / 1e6)), flush=True)
        del m
