"""Size-independent properties at the BASELINE size on the B200: Life 16384^2 Cyclic commutes with
translations of the initial condition (no oracle needed at this size), and its population is the sum of its cells."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_life_16384_commutes_with_translations():
    import torch
    from paraiso_b200.machines import life_machine, life_seed
    size, steps, (dy, dx) = (16384, 16384), 6, (-4099, 8191)
    m = life_machine(size)
    init = torch.from_numpy(life_seed(size[0], 0, size[1]))

    def run(start):
        m.call("init")
        m.set("cell", start.numpy())
        for _ in range(steps):
            m.call("proceed")
        return torch.from_numpy(m.get("cell")), int(m.scalar("population"))
    a, pop_a = run(init)
    b, pop_b = run(torch.roll(init, shifts=(dy, dx), dims=(0, 1)))
    assert torch.equal(b, torch.roll(a, shifts=(dy, dx), dims=(0, 1)))
    assert pop_a == pop_b == int(a.sum(dtype=torch.int64))
