"""2-GPU decomposition invariance through NCCL (halo rows by send/recv, reduce results by all_reduce):
the 2-rank result is bit-identical to the 1-GPU result.  Skipped on a single-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, which, size, steps, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from paraiso_b200.machines import hydro_machine, hydro_set_params, life_machine, life_seed
    if which == "diff3":
        from paraiso_b200.build import build_machine
        from paraiso_b200.examples.rank3 import diffusion3d_om
        from paraiso_b200.generator.native import Setup
        from paraiso_b200.runtime import Machine
        desc, so = build_machine(Setup(local_size=size, boundary=("Cyclic", "Open", "Open")), diffusion3d_om(), tag="Diff3_COO")
        m = Machine(desc, so, size=size, device=dev, rank=rank, nranks=world)
        m.call("init")
        for _ in range(steps):
            m.call("proceed")
        ret[rank] = (m.z0, m.get("u"), float(m.scalar("peak")))
    elif which in ("life", "life_graph"):
        m = life_machine(size, device=dev, rank=rank, nranks=world)
        m.call("init")
        m.set("cell", life_seed(size[0], m.y0, m.nyl, nx_global=size[0]))
        _steps(m, steps, which.endswith("_graph"))
        ret[rank] = (m.y0, m.get("cell"), int(m.scalar("population")), m.early_exchanges)
    else:
        m = hydro_machine(size, device=dev, rank=rank, nranks=world, fast=which.startswith("hydro_fast"))
        hydro_set_params(m, size)
        m.call("init")
        _steps(m, steps, which.endswith("_graph"))
        ret[rank] = (m.y0, {n: m.get(n) for n in ("density", "velocity0", "velocity1", "pressure")}, float(m.scalar("time")))
    dist.barrier()
    dist.destroy_process_group()


def _steps(m, steps, graph):
    """`steps` proceed() calls; with `graph`: two eager calls, then replays of a captured pair (Machine.capture)."""
    if not graph:
        for _ in range(steps):
            m.call("proceed")
        return
    assert steps % 2 == 0 and steps >= 4
    m.call("proceed"); m.call("proceed")
    g = m.capture("proceed", 2)
    for _ in range((steps - 2) // 2):
        g.replay()


def _multi(which, size, steps, port):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, which, size, steps, ret), nprocs=2, join=True)
    return sorted([ret[r] for r in range(2)], key=lambda p: p[0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_life_two_gpus_equal_one_gpu():
    from paraiso_b200.machines import life_machine, life_seed
    size, steps = (4096, 3000), 20
    m = life_machine(size)
    m.call("init")
    m.set("cell", life_seed(size[0], 0, size[1]))
    for _ in range(steps):
        m.call("proceed")
    parts = _multi("life", size, steps, 29711)
    assert np.array_equal(m.get("cell"), np.concatenate([p[1] for p in parts], axis=0))
    assert all(p[2] == int(m.scalar("population")) for p in parts)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_life_two_gpus_boundary_first_and_graph_replay_equal_one_gpu():
    """The stage is launched once per step in boundary-first chunk order, the ghost rows leave when the in-kernel signal
    fires, and pairs of steps are replayed from a CUDA graph (NCCL send/recv captured): still bit-identical to one GPU."""
    from paraiso_b200.machines import life_machine, life_seed
    size, steps = (4096, 3000), 20
    m = life_machine(size)
    m.call("init")
    m.set("cell", life_seed(size[0], 0, size[1]))
    _steps(m, steps, True)       # the 1-GPU machine replays a graph too
    parts = _multi("life_graph", size, steps, 29717)
    assert all(p[3] >= 2 for p in parts), "the boundary-first path was not taken"
    assert np.array_equal(m.get("cell"), np.concatenate([p[1] for p in parts], axis=0))
    assert all(p[2] == int(m.scalar("population")) for p in parts)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("which", ["hydro_graph", "hydro_fast_graph"])
def test_hydro_two_gpus_graph_replay_equal_one_gpu(which):
    """Graph replay of step pairs on two GPUs (ghost rows + dt all-reduce inside the graph; the fast build also carries
    its dt reduce from call to call) equals eager stepping on one GPU bit for bit."""
    from paraiso_b200.machines import hydro_machine, hydro_set_params
    size, steps = (1024, 777), 10
    m = hydro_machine(size, fast=which.startswith("hydro_fast"))
    hydro_set_params(m, size)
    m.call("init")
    for _ in range(steps):
        m.call("proceed")
    parts = _multi(which, size, steps, 29719 + (which == "hydro_graph"))
    for n in ("density", "velocity0", "velocity1", "pressure"):
        two = np.concatenate([p[1][n] for p in parts], axis=0)
        assert np.array_equal(m.get(n).view(np.uint64), two.view(np.uint64)), n
    assert all(p[2] == float(m.scalar("time")) for p in parts)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_hydro_two_gpus_equal_one_gpu():
    from paraiso_b200.machines import hydro_machine, hydro_set_params
    size, steps = (1024, 777), 10
    m = hydro_machine(size)
    hydro_set_params(m, size)
    m.call("init")
    for _ in range(steps):
        m.call("proceed")
    parts = _multi("hydro", size, steps, 29713)
    for n in ("density", "velocity0", "velocity1", "pressure"):
        two = np.concatenate([p[1][n] for p in parts], axis=0)
        assert np.array_equal(m.get(n).view(np.uint64), two.view(np.uint64)), n
    assert all(p[2] == float(m.scalar("time")) for p in parts)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_rank3_two_gpus_equal_one_gpu():
    """Rank-3 machines are cut along axis 2: ghost planes by NCCL send/recv, the Max reduce by all_reduce."""
    from paraiso_b200.build import build_machine
    from paraiso_b200.examples.rank3 import diffusion3d_om
    from paraiso_b200.generator.native import Setup
    from paraiso_b200.runtime import Machine
    size, steps = (200, 48, 37), 5
    desc, so = build_machine(Setup(local_size=size, boundary=("Cyclic", "Open", "Open")), diffusion3d_om(), tag="Diff3_COO")
    m = Machine(desc, so, size=size)
    m.call("init")
    for _ in range(steps):
        m.call("proceed")
    parts = _multi("diff3", size, steps, 29715)
    two = np.concatenate([p[1] for p in parts], axis=0)
    assert np.array_equal(m.get("u").view(np.uint64), two.view(np.uint64))
    assert all(p[2] == float(m.scalar("peak")) for p in parts)
