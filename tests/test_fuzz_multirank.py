"""Random OM programs (tests/test_fuzz_programs.py) on the Python host with several gloo ranks: every rank's slab of every
array and every scalar equals the one-rank run — decomposition invariance for arbitrary stencils, Open and Cyclic cuts,
reduces that feed a second stage (all_reduce between the stages) and reduces only the host reads (deferred all_reduce)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _machine(seed, rank, world, width=None):
    sys.path.insert(0, ROOT)
    from paraiso_b200.runtime import Machine
    from tests.emu.build_emu import build_emulated
    from tests.test_fuzz_programs import random_program
    om, setup = random_program(seed)
    setup.local_size = (width or setup.local_size[0], 24)
    desc, so = build_emulated(setup, om(), tag=f"fuzzdev_{seed}" + (f"_w{width}" if width else ""))
    m = Machine(desc, so, size=setup.local_size, device="cpu", rank=rank, nranks=world, _emulated=True)
    rng = np.random.default_rng(seed)
    full = {n: rng.integers(-30, 30, (24, setup.local_size[0])).astype(np.int32) for n in ("a", "b")}
    for n in ("a", "b"):
        m.set(n, full[n][m.y0:m.y0 + m.nyl])
    for _ in range(3):
        m.call("k")
    return m


def _worker(rank, world, port, seed, ret, width=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _machine(seed, rank, world, width)
    ret[rank] = (m.y0, m.nyl, m.get("a"), m.get("b"), int(m.scalar("s")), int(m.scalar("t")))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("seed,world", [(3, 2), (14, 3), (17, 2), (76, 3)])
def test_random_program_ranks_equal_one_rank(seed, world, width=None):
    one = _machine(seed, 0, 1, width)
    want = (one.get("a"), one.get("b"), int(one.scalar("s")), int(one.scalar("t")))
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29650 + seed % 200 + world, seed, ret, width), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        y0, nyl, a, b, s, t = ret[rank]
        assert np.array_equal(a, want[0][y0:y0 + nyl]) and np.array_equal(b, want[1][y0:y0 + nyl]), (seed, rank)
        assert (s, t) == want[2:], (seed, rank)


def test_cyclic_axis_0_narrower_than_its_ghost_zones_on_two_ranks():
    """Seed 1069 (Cyclic / Cyclic, reach 3 + 3) on a grid 5 cells wide: every cell has two images along axis 0; the host
    redoes the wrap after each launch and before the ghost-row exchange (runtime.Machine._fill_ghosts_modular)."""
    test_random_program_ranks_equal_one_rank(1069, 2, width=5)
