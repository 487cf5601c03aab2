"""Life proceed kernel: schedule search on the GPU box (paraiso_b200.tuning.grid_search)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.examples.life import life_om, life_setup  # noqa: E402
from paraiso_b200.machines import life_seed  # noqa: E402
from paraiso_b200.tuning import candidates, grid_search  # noqa: E402

SPACES = {
    "default": dict(skeleton=["ring"], threads_light=[64, 128, 256], prefetch_rows=[2, 3], row_window=[True, False], chunk_rows_light=[32]),
    "stream": dict(skeleton=["stream"], threads_light=[64, 128], stream_prefetch=[2, 4], chunk_rows_light=[32, 128]),
    "chunks": dict(chunk_rows_light=[16, 24, 32, 48, 64, 128]),
}

if __name__ == "__main__":
    space = SPACES[sys.argv[1]] if len(sys.argv) > 1 and sys.argv[1] in SPACES else json.loads(sys.argv[1])
    size = (16384, 16384)

    def prepare(m):
        m.call("init")
        m.set("cell", life_seed(size[0], 0, size[1]))
    for r in grid_search(lambda: life_setup("master"), lambda: life_om("master"), candidates(space), size, stage=0, prepare=prepare):
        if "ms" in r:
            r["GB_per_s"] = size[0] * size[1] * 8 / r["ms"] / 1e6
        print(json.dumps(r), flush=True)
