"""Rank-3 machines on the GPU through the C ABI, against the oracle (bit-exact: integer Life3, IEEE double Diff3)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(om_fn, setup, tag):
    from oracle.cpu import OracleMachine
    from paraiso_b200.build import build_machine
    from paraiso_b200.runtime import Machine
    desc, so = build_machine(setup, om_fn(), tag=tag)
    return Machine(desc, so), OracleMachine(setup, om_fn(), openmp=True, opt="-O2")


@pytest.mark.parametrize("bnd,size", [(("Cyclic", "Cyclic", "Cyclic"), (300, 70, 33)), (("Open", "Cyclic", "Open"), (129, 40, 17))])
def test_life3d_bit_exact(bnd, size):
    from paraiso_b200.examples.rank3 import life3d_om
    from paraiso_b200.generator.native import Setup
    setup = Setup(local_size=size, boundary=bnd)
    m, o = _pair(life3d_om, setup, f"Life3_t_{''.join(b[0] for b in bnd)}")
    init = (np.random.default_rng(5).random(o.array("cell").shape) < 0.3).astype(np.int32)
    m.set("cell", init, with_margin=True)
    o.array("cell")[...] = init
    for t in range(6):
        m.call("proceed"); o.call("proceed")
        assert np.array_equal(m.get("cell", with_margin=True), o.array("cell")), t
        assert int(m.scalar("population")) == int(o.scalar("population")[0])
    assert int(m.scalar("generation")) == 6


@pytest.mark.parametrize("bnd", [("Open", "Open", "Open"), ("Cyclic", "Open", "Cyclic")])
def test_diffusion3d_bit_identical(bnd):
    from paraiso_b200.examples.rank3 import diffusion3d_om
    from paraiso_b200.generator.native import Setup
    setup = Setup(local_size=(260, 41, 19), boundary=bnd)
    m, o = _pair(diffusion3d_om, setup, f"Diff3_t_{''.join(b[0] for b in bnd)}")
    m.call("init"); o.call("init")
    for t in range(4):
        m.call("proceed"); o.call("proceed")
        assert np.array_equal(m.get("u", with_margin=True).view(np.uint64), o.array("u").view(np.uint64)), t
        assert m.scalar("peak") == o.scalar("peak")[0]
