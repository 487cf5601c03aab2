"""The reference's analytic exams (attic/GA.reproduce/massive-test.cu:58-241: entropy wave, sound wave, Sod tube against
riemann-solver.h) run on the DEVICE kernels of both builds: convergence to the exact solutions, the bit-exact build
bit-identical to the oracle on every flow, the fast_math build within 1e-12 of it."""
import numpy as np
import pytest

from tests.test_hydro_analytic import GAMMA, SOD_L, SOD_R, Orc, entropy_wave, riemann_exact, sod, sound_wave

pytestmark = pytest.mark.gpu
NAMES = ("density", "velocity0", "velocity1", "pressure")


def periodic_machine(setup, fast):
    from paraiso_b200.build import build_machine
    from paraiso_b200.examples.hydro import hydro_om
    from paraiso_b200.runtime import Machine
    desc, so = build_machine(setup, hydro_om("periodic"), tag="HydroPeriodic_CC" + ("_fast" if fast else ""), fmad=fast)
    return Machine(desc, so, size=setup.local_size)


def prebuild():
    from paraiso_b200.build import build_machine
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    for fast in (False, True):
        build_machine(hydro_setup((64, 64), periodic=True, fast=fast), hydro_om("periodic"), tag="HydroPeriodic_CC" + ("_fast" if fast else ""), fmad=fast)


def gpu_cls(fast):
    class Gpu:
        def __init__(self, setup):
            from paraiso_b200.examples.hydro import hydro_setup
            self.m = periodic_machine(hydro_setup(setup.local_size, periodic=True, fast=fast), fast)
        def setp(self, p):
            for k, v in p.items(): self.m.set_scalar(k, v)
        def seta(self, n, a): self.m.set(n, a)
        def geta(self, n): return self.m.get(n)
        def time(self): return float(self.m.scalar("time"))
        def step(self): self.m.call("proceed")
    return Gpu


def _same_as_oracle(m, o, fast):
    if not fast:
        assert m.time() == o.time()
        for n in NAMES:
            assert np.array_equal(m.geta(n).view(np.uint64), o.geta(n).view(np.uint64)), n
    else:
        # north-star tolerance: 1e-12 relative on the conserved variables; the scale is the flow's (a 1e-5 velocity perturbation
        # on an O(1) background carries the background's rounding noise)
        assert abs(m.time() - o.time()) <= 1e-12 * o.time()
        scale = max(float(np.max(np.abs(o.geta(n)))) for n in NAMES)
        for n in NAMES:
            assert np.max(np.abs(m.geta(n) - o.geta(n))) <= 1e-12 * scale, n


@pytest.mark.parametrize("fast", [False, True])
def test_entropy_wave_on_the_device(fast):
    errs = {}
    for n in (32, 64, 128):
        m, errs[n], steps = entropy_wave(n, 0.1, gpu_cls(fast))
        if n == 64:
            o, _e, steps_o = entropy_wave(n, 0.1, Orc)
            assert steps == steps_o
            _same_as_oracle(m, o, fast)
    assert errs[64] < errs[32] / 2.4 and errs[128] < errs[64] / 2.4, errs
    assert errs[128] < 6e-4, errs


@pytest.mark.parametrize("fast", [False, True])
def test_sound_wave_on_the_device(fast):
    errs = {}
    for n in (64, 128, 256):
        m, errs[n], steps = sound_wave(n, 0.25, gpu_cls(fast))
        assert np.max(np.abs(m.geta("velocity1"))) < 1e-12
        if n == 64:
            o, _e, steps_o = sound_wave(n, 0.25, Orc)
            assert steps == steps_o
            _same_as_oracle(m, o, fast)
    assert errs[128] < errs[64] / 2.4 and errs[256] < errs[128] / 2.4, errs
    assert errs[256] < 3e-3, errs


@pytest.mark.parametrize("fast", [False, True])
def test_sod_tube_on_the_device(fast):
    t_end, errs = 0.08, {}
    for n in (256, 512, 1024):
        m, xs, steps = sod(n, t_end, gpu_cls(fast))
        t = m.time()
        win = (xs > 0.55) & (xs < 0.95)
        rho, u, p = riemann_exact(*SOD_L, *SOD_R, GAMMA, (xs[win] - 0.75) / t)
        errs[n] = float(np.mean(np.abs(m.geta("density")[0, :][win] - rho)))
        assert np.ptp(m.geta("density"), axis=0).max() == 0.0
        if n == 256:
            o, _xs, steps_o = sod(n, t_end, Orc)
            assert steps == steps_o
            _same_as_oracle(m, o, fast)
    assert errs[512] < errs[256] / 1.6 and errs[1024] < errs[512] / 1.6, errs
    assert errs[1024] < 3.5e-3, errs
    # the tube along axis 1 is the transpose of the tube along axis 0
    a, _x, sa = sod(128, 0.05, gpu_cls(fast), axis=0)
    b, _x, sb = sod(128, 0.05, gpu_cls(fast), axis=1)
    assert sa == sb and np.allclose(a.geta("density"), b.geta("density").T, rtol=1e-12, atol=0)
