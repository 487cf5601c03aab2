"""paraiso_b200.hostio.HostPipeline: stepping with the state uploaded from pinned host memory and the result grid downloaded
every step (three streams, two staging slots each way) returns, step by step, exactly what synchronous stepping returns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pipelined_steps_equal_synchronous_steps():
    import torch
    from paraiso_b200.hostio import HostPipeline
    from paraiso_b200.machines import life_machine, life_seed
    size, steps = (1000, 777), 7
    a, b = life_machine(size), life_machine(size)
    a.call("init"); b.call("init")
    states = [life_seed(size[0], 0, size[1], seed=100 + k) for k in range(steps)]      # a different input every step
    pipe = HostPipeline(a, "proceed", ["cell"])
    ins = [{"cell": torch.from_numpy(s).pin_memory()} for s in states]
    outs = [{"cell": torch.empty((size[1], size[0]), dtype=torch.int32).pin_memory()} for _ in range(steps)]
    for k in range(steps):
        pipe.submit(ins[k], outs[k])
    pipe.drain()
    torch.cuda.synchronize()
    assert pipe.h2d_bytes == pipe.d2h_bytes == 4 * size[0] * size[1]
    for k in range(steps):
        b.set("cell", states[k])
        b.call("proceed")
        assert np.array_equal(outs[k]["cell"].numpy(), b.get("cell")), k
    assert np.array_equal(a.get("cell"), b.get("cell"))
    assert int(a.scalar("population")) == int(b.scalar("population"))


def test_pipeline_with_hydro_state_arrays():
    import torch
    from paraiso_b200.hostio import HostPipeline
    from paraiso_b200.machines import hydro_machine, hydro_set_params
    size = (256, 192)
    names = ["density", "velocity0", "velocity1", "pressure"]
    a, b = hydro_machine(size, fast=True), hydro_machine(size, fast=True)
    for m in (a, b):
        hydro_set_params(m, size)
        m.call("init")
    host_in = {n: torch.from_numpy(np.ascontiguousarray(a.get(n))).pin_memory() for n in names}
    host_out = {n: torch.empty_like(host_in[n]).pin_memory() for n in names}
    pipe = HostPipeline(a, "proceed", names)
    for _ in range(3):                       # the same input three times: the output must be one step from it each time
        pipe.submit(host_in, host_out)
    pipe.drain()
    torch.cuda.synchronize()
    b.call("proceed")
    for n in names:
        assert np.array_equal(host_out[n].numpy().view(np.uint64), b.get(n).view(np.uint64)), n
