"""Rank-3 machines: the axis-2 part of every shift is lowered to plane-shifted virtual inputs
(schedule.lower_z), one layer of CTAs per plane.  Emulated kernels against the oracle, which is rank-generic like the
reference's PlanTrans."""
import numpy as np
import pytest

from oracle.cpu import OracleMachine
from paraiso_b200.annotation import CYCLIC, OPEN
from paraiso_b200.generator.native import Setup
from paraiso_b200.examples.rank3 import diffusion3d_om, life3d_om
from paraiso_b200.runtime import Machine
from tests.emu.build_emu import build_emulated


def mem_shape3(setup, om_fn):
    from paraiso_b200.generator.plan import translate
    p = translate(setup, om_fn())
    return tuple(reversed(p.memory_size))


def run_both3(om_fn, setup, kernels, tag, fill, rtol=0.0):
    desc, so = build_emulated(setup, om_fn(), tag=tag)
    assert desc["dim"] == 3
    m = Machine(desc, so, device="cpu", _emulated=True)
    o = OracleMachine(setup, om_fn())
    for name, arr in fill.items():
        m.set(name, arr, with_margin=True)
        o.array(name)[...] = arr
    for k in kernels:
        m.call(k); o.call(k)
        for s in desc["statics"]:
            if s["realm"] == "Array":
                a, b = m.get(s["name"], with_margin=True), o.array(s["name"])
                assert a.shape == b.shape
                if rtol:
                    assert np.allclose(a, b, rtol=rtol, atol=0), (k, s["name"])
                else:
                    assert np.array_equal(a, b), (k, s["name"])
            else:
                a, b = m.scalar(s["name"]), o.scalar(s["name"])[0]
                assert (abs(a - b) <= rtol * abs(b)) if rtol else (a == b), (k, s["name"])
    return m, o


@pytest.mark.parametrize("bnd,size", [((CYCLIC, CYCLIC, CYCLIC), (37, 11, 6)), ((OPEN, CYCLIC, OPEN), (20, 9, 5)),
                                      ((CYCLIC, OPEN, CYCLIC), (130, 7, 3)), ((CYCLIC, CYCLIC, CYCLIC), (5, 4, 1))])
def test_life3d_bit_exact(bnd, size):
    setup = Setup(local_size=size, boundary=bnd)
    fill = {"cell": (np.random.default_rng(11).random(mem_shape3(setup, life3d_om)) < 0.3).astype(np.int32)}
    run_both3(life3d_om, setup, ["proceed"] * 3, f"life3d_{''.join(b[0] for b in bnd)}", fill)


@pytest.mark.parametrize("bnd", [(OPEN, OPEN, OPEN), (CYCLIC, OPEN, CYCLIC)])
def test_diffusion3d_bit_identical(bnd):
    setup = Setup(local_size=(70, 12, 7), boundary=bnd)
    run_both3(diffusion3d_om, setup, ["init", "proceed", "proceed"], f"diff3d_{''.join(b[0] for b in bnd)}", {})


@pytest.mark.parametrize("zplanes,size", [(2, (37, 11, 6)), (2, (20, 9, 5)), (4, (33, 7, 10)), (3, (16, 5, 1))])
def test_life3d_several_planes_per_cta(zplanes, size):
    """Tuning.planes_per_cta: a CTA computes Z consecutive planes (Z + 2 planes staged for Z planes of output); the last
    group may be incomplete (6 = 3 x 2, 5 = 2 x 2 + 1, 10 = 2 x 4 + 2, 1 < 3)."""
    setup = Setup(local_size=size, boundary=(CYCLIC, CYCLIC, CYCLIC))
    setup.tuning.planes_per_cta = zplanes
    fill = {"cell": (np.random.default_rng(12).random(mem_shape3(setup, life3d_om)) < 0.3).astype(np.int32)}
    desc, so = build_emulated(setup, life3d_om(), tag=f"life3d_z{zplanes}")
    assert desc["kernels"][0]["stages"][0]["zplanes"] == zplanes
    run_both3(life3d_om, setup, ["proceed"] * 3, f"life3d_z{zplanes}", fill)


@pytest.mark.parametrize("bnd", [(OPEN, OPEN, OPEN), (CYCLIC, OPEN, CYCLIC)])
def test_diffusion3d_two_planes_per_cta(bnd):
    """Open axis 2 (valid masks per plane offset), a Max reduce over both planes of a group, loadIndex(2) per plane."""
    setup = Setup(local_size=(70, 12, 7), boundary=bnd)
    setup.tuning.planes_per_cta = 2
    run_both3(diffusion3d_om, setup, ["init", "proceed", "proceed"], f"diff3d_z2_{''.join(b[0] for b in bnd)}", {})
