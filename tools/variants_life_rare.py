"""Life 16384^2: the rarely taken block of a row (partial vectors, ghost copies) inline or as a noinline closure (Tuning.cold_rare),
next to a library built before the per-thread `ghost_x` predicate if one was kept (_generated/variant_Life_old), same box.
`--prebuild` compiles here."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.build import build_machine  # noqa: E402
from paraiso_b200.examples.life import life_om, life_setup  # noqa: E402

size = (16384, 16384)
if __name__ == "__main__":
    built = []
    gen = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "paraiso_b200", "_generated")
    for name, d in (("before ghost_x", "variant_Life_old"), ("ghost_x, general block only", "variant_Life_ghostx")):   # kept libraries of earlier generators
        if os.path.exists(os.path.join(gen, d, "libom_Life.so")):
            with open(os.path.join(gen, d, "Life_abi.json")) as f:
                built.append((name, (json.load(f), os.path.join(gen, d, "libom_Life.so"))))
    for cold in (False, True):
        s = life_setup("master")
        s.tuning.cold_rare = cold
        built.append((f"lean ghost block, cold_rare={cold}", build_machine(s, life_om("master"), tag=f"variant_Life_cold{int(cold)}", verbose=True)))
    # timing bound only (wrong ghost columns): no thread of an edge strip takes the block for its ghost copy
    from paraiso_b200.build import compile_kernels, generate_to
    s = life_setup("master")
    desc, d = generate_to(s, life_om("master"), tag="variant_Life_noghost")
    with open(os.path.join(d, "Life_kernels.cu")) as f:
        src = f.read()
    with open(os.path.join(d, "Life_kernels.cu"), "w") as f:
        f.write(src.replace("const bool ghost_x = edge_x &&", "const bool ghost_x = false &&"))
    built.append(("no x ghost copies (bound, wrong result)", (desc, compile_kernels(d, desc["name"], verbose=True))))
    if "--prebuild" in sys.argv:
        sys.exit(0)
    import numpy as np
    import torch
    from paraiso_b200.machines import life_seed
    from paraiso_b200.runtime import Machine
    from paraiso_b200.tuning import measure
    seed = torch.from_numpy(life_seed(size[0], 0, size[1])).pin_memory()
    ref = None
    for rep in range(2):
        for name, (desc, so) in built:
            m = Machine(desc, so, size=size)
            m.call("init")
            m.set_from_host("cell", seed)
            st = m.kernels["proceed"]["stages"][0]
            ms = min(measure(m, "proceed", steps=20, stage=0) for _ in range(3))
            m.set_from_host("cell", seed)
            for _ in range(3):
                m.call("proceed")
            got = (m.get("cell").astype(np.int64).sum(), int(m.scalar("population")))
            ref = ref or got
            print(json.dumps(dict(variant=name, occupancy=getattr(m.lib, st["symbol"] + "_occupancy")(), chunks=m._geom(st).nchunks, ms=ms,
                                  GBs=2 * 4 * size[0] * size[1] / ms / 1e6, same_result=(got == ref))), flush=True)
            del m
