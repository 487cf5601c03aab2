"""Life 16384^2, final kernel: the number of chunks around the wave-rounded 1040 (Machine.force_chunks), same box."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from paraiso_b200.machines import life_machine, life_seed  # noqa: E402
from paraiso_b200.tuning import measure  # noqa: E402

size = (16384, 16384)
seed = torch.from_numpy(life_seed(size[0], 0, size[1])).pin_memory()
for rep in range(2):
    for chunks in (0, 1024, 999, 1040, 1082, 1124, 1165, 915, 957):
        m = life_machine(size)
        if chunks:
            m.force_chunks = chunks
        m.call("init")
        m.set_from_host("cell", seed)
        st = m.kernels["proceed"]["stages"][0]
        ms = min(measure(m, "proceed", steps=20, stage=0) for _ in range(3))
        print(json.dumps(dict(force_chunks=chunks, chunks=m._geom(st).nchunks, ms=ms, GBs=2 * 4 * size[0] * size[1] / ms / 1e6)), flush=True)
        del m
