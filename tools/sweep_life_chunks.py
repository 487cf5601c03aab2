"""Life 16384^2: kernel time against the number of row chunks per strip (launch geometry only — one build).  The default rule
(runtime.Machine._geom) rounds rows / Tuning.chunk_rows_light to whole waves of resident CTAs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from paraiso_b200.machines import life_machine, life_seed  # noqa: E402
from paraiso_b200.tuning import measure  # noqa: E402

from paraiso_b200.build import build_machine  # noqa: E402
from paraiso_b200.examples.life import life_om, life_setup  # noqa: E402
from paraiso_b200.runtime import Machine  # noqa: E402

size = (16384, 16384)
seed = None


def variant_setups():
    a = life_setup("master")
    b = life_setup("master")
    b.tuning.peel_fill = False
    c = life_setup("master")
    c.tuning.barrier_group = True
    return [("default (steady-state copy of the row loop)", None, a), ("one row loop, every row behind its tests", "Life_CC_tested", b),
            ("barrier per group of rows", "Life_CC_grouped", c)]


def variants():
    for label, tag, s in variant_setups():
        if tag is None:
            yield label, life_machine(size)
        else:
            desc, so = build_machine(s, life_om("master"), tag=tag)
            yield label, Machine(desc, so, size=size)


if __name__ == "__main__":
    if "--prebuild" in sys.argv:
        for _label, tag, s in variant_setups():
            if tag:
                build_machine(s, life_om("master"), tag=tag)
        sys.exit(0)
    seed = torch.from_numpy(life_seed(size[0], 0, size[1])).pin_memory()
    for label, m in variants():
        m.call("init")
        m.set_from_host("cell", seed)
        st = m.kernels["proceed"]["stages"][0]
        occ = getattr(m.lib, st["symbol"] + "_occupancy")()
        for chunks in [0, 541, 656, 749, 790, 832, 874, 915, 1000]:
            m.force_chunks = chunks
            m._geom_cache.clear()
            ms = min(measure(m, "proceed", steps=20, stage=0) for _ in range(3))
            g = m._geom(st)
            print(json.dumps(dict(variant=label, occupancy=occ, chunks=g.nchunks, forced=chunks, rows_per_chunk=16384 / g.nchunks, ctas=g.nchunks * 32,
                                  waves=g.nchunks * 32 / (148 * occ), ms=ms, GBs=2 * 4 * 16384 * 16384 / ms / 1e6)), flush=True)
        del m
