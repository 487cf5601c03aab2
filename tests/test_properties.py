"""Size-independent properties of the two programs, on emulated kernels (the GPU versions at BASELINE sizes are in
tests/test_gpu_properties.py): Life on a Cyclic grid commutes with translations; the periodic Hydro variant is a
finite-volume scheme, so total mass, momentum and energy change only by rounding."""
import numpy as np

from paraiso_b200.examples.hydro import hydro_om, hydro_setup
from paraiso_b200.examples.life import life_om, life_setup
from paraiso_b200.machines import life_seed
from paraiso_b200.runtime import Machine
from tests.emu.build_emu import build_emulated


def life_after(m, init, steps):
    m.call("init")
    m.set("cell", init)
    for _ in range(steps):
        m.call("proceed")
    return m.get("cell").copy(), int(m.scalar("population"))


def test_life_commutes_with_translations():
    size = (96, 40)
    desc, so = build_emulated(life_setup("master", size=size), life_om("master"), tag="Life_prop")
    m = Machine(desc, so, size=size, device="cpu", _emulated=True)
    init = life_seed(size[0], 0, size[1])
    a, pop_a = life_after(m, init, 7)
    for (dy, dx) in ((0, 1), (3, 0), (-5, 17), (39, 95)):
        b, pop_b = life_after(m, np.roll(np.roll(init, dy, 0), dx, 1), 7)
        assert np.array_equal(b, np.roll(np.roll(a, dy, 0), dx, 1)), (dy, dx)
        assert pop_b == pop_a == int(a.sum())


def conserved_totals(get, gamma=5.0 / 3.0):
    rho, v0, v1, p = (get(n).astype(np.float64) for n in ("density", "velocity0", "velocity1", "pressure"))
    return np.array([rho.sum(), (rho * v0).sum(), (rho * v1).sum(), (0.5 * rho * (v0 * v0 + v1 * v1) + p / (gamma - 1.0)).sum()])


def test_periodic_hydro_conserves_mass_momentum_energy():
    n = 32
    setup = hydro_setup((n, n), periodic=True)
    desc, so = build_emulated(setup, hydro_om("periodic"), tag="HydroPeriodic")
    m = Machine(desc, so, size=(n, n), device="cpu", _emulated=True)
    for k, v in dict(time=0.0, cfl=0.4, extent0=1.0, extent1=1.0, dR0=1.0 / n, dR1=1.0 / n).items():
        m.set_scalar(k, v)
    xs = (np.arange(n) + 0.5) / n
    X, Y = np.meshgrid(xs, xs)
    m.set("density", 1.0 + 0.3 * np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y))
    m.set("velocity0", 0.8 + 0.2 * np.cos(2 * np.pi * Y)); m.set("velocity1", -0.3 + 0.1 * np.sin(4 * np.pi * X))
    m.set("pressure", 1.0 + 0.2 * np.cos(2 * np.pi * (X + Y)))
    c0 = conserved_totals(m.get)
    for _ in range(12):
        m.call("proceed")
    c1 = conserved_totals(m.get)
    scale = np.array([c0[0], c0[0], c0[0], c0[3]])          # momenta relative to the total mass (v ~ 1)
    assert np.all(np.abs(c1 - c0) <= 1e-13 * scale), (c1 - c0) / scale
    assert float(m.scalar("time")) > 0
