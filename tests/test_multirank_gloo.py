"""world_size-2 runs of the slab-decomposed host path on CPU (gloo backend, emulated kernels):
decomposition invariance — the 2-rank result is bit-identical to the 1-rank result — for the Cyclic ring
(Life: halo rows + population all_reduce(sum)) and the Open chain (Hydro: halo rows + dt all_reduce(min))."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, which, size, steps, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from paraiso_b200.runtime import Machine
    from tests.emu.build_emu import build_emulated
    if which == "life":
        from paraiso_b200.examples.life import life_om, life_setup
        from paraiso_b200.machines import life_seed
        desc, so = build_emulated(life_setup("master", size=size), life_om("master"), tag="Life_ring_1_cp_async")
        m = Machine(desc, so, size=size, device="cpu", rank=rank, nranks=world, _emulated=True)
        m.call("init")
        m.set("cell", life_seed(size[0], m.y0, m.nyl, nx_global=size[0]))
        for _ in range(steps):
            m.call("proceed")
        ret[rank] = (m.y0, m.get("cell"), int(m.scalar("population")), m.early_exchanges)
    else:
        from paraiso_b200.examples.hydro import hydro_om, hydro_setup
        from paraiso_b200.machines import hydro_set_params
        setup = hydro_setup(size)
        setup.tuning.carry_reduces = which == "hydro_carry"
        desc, so = build_emulated(setup, hydro_om("master"), tag="Hydro_carry" if which == "hydro_carry" else None)
        m = Machine(desc, so, size=size, device="cpu", rank=rank, nranks=world, _emulated=True)
        hydro_set_params(m, size)
        m.call("init")
        for _ in range(steps):
            m.call("proceed")
        ret[rank] = (m.y0, {n: m.get(n) for n in ("density", "velocity0", "velocity1", "pressure")}, float(m.scalar("time")))
    dist.barrier()
    dist.destroy_process_group()


def _run(which, size, steps, world, port):
    if world == 1:
        ret = {}
        # a 1-rank run needs no process group
        sys.path.insert(0, ROOT)
        from paraiso_b200.runtime import Machine
        from tests.emu.build_emu import build_emulated
        if which == "life":
            from paraiso_b200.examples.life import life_om, life_setup
            from paraiso_b200.machines import life_seed
            desc, so = build_emulated(life_setup("master", size=size), life_om("master"), tag="Life_ring_1_cp_async")
            m = Machine(desc, so, size=size, device="cpu", _emulated=True)
            m.call("init")
            m.set("cell", life_seed(size[0], 0, size[1]))
            for _ in range(steps):
                m.call("proceed")
            return m.get("cell"), int(m.scalar("population"))
        from paraiso_b200.examples.hydro import hydro_om, hydro_setup
        from paraiso_b200.machines import hydro_set_params
        setup = hydro_setup(size)
        setup.tuning.carry_reduces = which == "hydro_carry"
        desc, so = build_emulated(setup, hydro_om("master"), tag="Hydro_carry" if which == "hydro_carry" else None)
        m = Machine(desc, so, size=size, device="cpu", _emulated=True)
        hydro_set_params(m, size)
        m.call("init")
        for _ in range(steps):
            m.call("proceed")
        return {n: m.get(n) for n in ("density", "velocity0", "velocity1", "pressure")}, float(m.scalar("time"))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, which, size, steps, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


@pytest.mark.parametrize("world", [2, 3])      # 3 ranks: uneven slabs (17 + 17 + 16 rows)
def test_life_ranks_equal_one_rank(world):
    size, steps = (96, 50), 5
    cell1, pop1 = _run("life", size, steps, 1, 0)
    parts = _run("life", size, steps, world, 29611 + world)
    cell2 = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])], axis=0)
    assert len(parts) == world
    assert np.array_equal(cell1, cell2)
    assert all(p[2] == pop1 for p in parts)      # all_reduce(sum) gives every rank the global population


@pytest.mark.parametrize("world,size", [(2, (96, 120)), (3, (70, 170))])
def test_life_boundary_first_launch_signals_and_equals_one_rank(world, size):
    """Slabs tall enough for several chunks: the stage is launched once in boundary-first chunk order, the CTAs of the
    first and the last chunk raise the "boundary rows written" flag (checked by the emulated om_wait_boundary), and the
    result still equals the 1-rank run bit for bit (3 ranks: uneven slabs 57 + 57 + 56)."""
    steps = 4
    cell1, pop1 = _run("life", size, steps, 1, 0)
    parts = _run("life", size, steps, world, 29631 + world)
    assert all(p[3] >= steps for p in parts), "the boundary-first path was not taken"
    cell2 = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])], axis=0)
    assert np.array_equal(cell1, cell2)
    assert all(p[2] == pop1 for p in parts)


@pytest.mark.parametrize("world", [2, 3])      # 3 ranks: slabs of 13 + 12 + 12 rows, a middle rank without physical margins
def test_hydro_ranks_equal_one_rank(world):
    size, steps = (40, 37), 3
    f1, t1 = _run("hydro", size, steps, 1, 0)
    parts = sorted(_run("hydro", size, steps, world, 29621 + world), key=lambda p: p[0])
    for n in f1:
        f2 = np.concatenate([p[1][n] for p in parts], axis=0)
        assert np.array_equal(f1[n].view(np.uint64), f2.view(np.uint64)), n
    assert all(p[2] == t1 for p in parts)        # all_reduce(min) of the CFL time step


def test_hydro_carried_dt_reduce_two_ranks():
    """Tuning.carry_reduces with several ranks: the carried dt is all_reduced right after the stage that produced it."""
    size, steps = (40, 37), 4
    f1, t1 = _run("hydro", size, steps, 1, 0)
    parts = sorted(_run("hydro_carry", size, steps, 2, 29641), key=lambda p: p[0])
    for n in f1:
        f2 = np.concatenate([p[1][n] for p in parts], axis=0)
        assert np.array_equal(f1[n].view(np.uint64), f2.view(np.uint64)), n
    assert all(p[2] == t1 for p in parts)


def _worker3(rank, world, port, which, size, steps, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ret[rank] = _rank3_run(which, size, steps, rank, world)
    dist.barrier()
    dist.destroy_process_group()


def _rank3_run(which, size, steps, rank, world):
    from paraiso_b200.examples.rank3 import diffusion3d_om, life3d_om
    from paraiso_b200.generator.native import Setup
    from paraiso_b200.runtime import Machine
    from tests.emu.build_emu import build_emulated
    if which == "life3":
        setup = Setup(local_size=size, boundary=("Cyclic", "Cyclic", "Cyclic"))
        desc, so = build_emulated(setup, life3d_om(), tag="life3d_CCC")
        m = Machine(desc, so, size=size, device="cpu", rank=rank, nranks=world, _emulated=True)
        full = (np.random.default_rng(3).random((size[2], size[1], size[0])) < 0.3).astype(np.int32)
        m.set("cell", full[m.z0:m.z0 + m.nzl])
        for _ in range(steps):
            m.call("proceed")
        return m.z0, m.get("cell"), int(m.scalar("population"))
    setup = Setup(local_size=size, boundary=("Open", "Open", "Open"))
    desc, so = build_emulated(setup, diffusion3d_om(), tag="diff3d_OOO")
    m = Machine(desc, so, size=size, device="cpu", rank=rank, nranks=world, _emulated=True)
    m.call("init")
    for _ in range(steps):
        m.call("proceed")
    return m.z0, m.get("u"), float(m.scalar("peak"))


@pytest.mark.parametrize("which,size", [("life3", (37, 9, 11)), ("diff3", (40, 10, 13))])
def test_rank3_slabs_along_axis2_equal_one_rank(which, size):
    """Rank-3 machines are cut along axis 2: ghost planes travel, loadIndex(2) carries the slab offset, the Max / Sum
    reduces are all_reduced (3 ranks: uneven slabs, a middle rank without physical margins)."""
    steps = 3
    _z, a1, s1 = _rank3_run(which, size, steps, 0, 1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker3, args=(3, 29651 + (which == "diff3"), which, size, steps, ret), nprocs=3, join=True)
    parts = sorted([ret[r] for r in range(3)], key=lambda p: p[0])
    a3 = np.concatenate([p[1] for p in parts], axis=0)
    assert np.array_equal(a1.view(np.uint8), a3.view(np.uint8))
    assert all(p[2] == s1 for p in parts)
