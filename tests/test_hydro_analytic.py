"""Physics-level acceptance (the reference's manual exam, attic/GA.reproduce/massive-test.cu:58-118):
an entropy wave — a density perturbation advected by a uniform flow at constant pressure — is an exact solution of
the Euler equations, rho(x, y, t) = rho0(x - u t, y - v t).  The generated solver (periodic variant of
examples/Hydro, Cyclic boundaries, emulated kernels) must converge to it when the mesh is refined, and must agree
with the oracle bit for bit."""
import numpy as np
import pytest

from oracle.cpu import OracleMachine
from paraiso_b200.examples.hydro import hydro_om, hydro_setup
from paraiso_b200.runtime import Machine
from tests.emu.build_emu import build_emulated


def entropy_wave(n, t_end, machine_cls):
    size = (n, n)
    setup = hydro_setup(size, periodic=True)
    m = machine_cls(setup)
    u, v, p0 = 1.0, 0.5, 1.0
    xs = (np.arange(n) + 0.5) / n
    X, Y = np.meshgrid(xs, xs)                       # [y, x]
    rho0 = lambda x, y: 1.0 + 0.2 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
    params = dict(time=0.0, cfl=0.4, extent0=1.0, extent1=1.0, dR0=1.0 / n, dR1=1.0 / n)
    m.setp(params)
    m.seta("density", rho0(X, Y)); m.seta("velocity0", np.full((n, n), u)); m.seta("velocity1", np.full((n, n), v))
    m.seta("pressure", np.full((n, n), p0))
    steps = 0
    while m.time() < t_end and steps < 10000:
        m.step()
        steps += 1
    t = m.time()
    exact = rho0(X - u * t, Y - v * t)
    return m, np.mean(np.abs(m.geta("density") - exact)), steps


class Emu:
    def __init__(self, setup):
        desc, so = build_emulated(setup, hydro_om("periodic"), tag="HydroPeriodic")
        self.m = Machine(desc, so, size=setup.local_size, device="cpu", _emulated=True)
    def setp(self, p):
        for k, v in p.items(): self.m.set_scalar(k, v)
    def seta(self, n, a): self.m.set(n, a)
    def geta(self, n): return self.m.get(n)
    def time(self): return float(self.m.scalar("time"))
    def step(self): self.m.call("proceed")


class Orc:
    def __init__(self, setup):
        self.o = OracleMachine(setup, hydro_om("periodic"))
    def setp(self, p):
        for k, v in p.items(): self.o.scalar(k)[0] = v
    def seta(self, n, a): self.o.interior(n)[...] = a
    def geta(self, n): return self.o.interior(n).copy()
    def time(self): return float(self.o.scalar("time")[0])
    def step(self): self.o.call("proceed")


def test_entropy_wave_converges_and_matches_oracle():
    t_end = 0.1
    errs = {}
    for n in (16, 32, 64):
        m, err, steps = entropy_wave(n, t_end, Emu)
        errs[n] = err
        if n == 32:
            o, err_o, steps_o = entropy_wave(n, t_end, Orc)
            assert steps == steps_o and m.time() == o.time()
            for name in ("density", "velocity0", "velocity1", "pressure"):
                assert np.array_equal(m.geta(name).view(np.uint64), o.geta(name).view(np.uint64)), name
    # second-order MUSCL with a limiter that clips extrema: the L1 error falls by clearly more than 2x per refinement
    assert errs[32] < errs[16] / 2.4 and errs[64] < errs[32] / 2.4, errs
    assert errs[64] < 2e-3, errs
