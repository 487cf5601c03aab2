"""GPU parity tests: the generated sm_100a kernels, driven through the C ABI, against the CPU oracle
(oracle/plantrans.py, pinned to the reference's generated C++ by tests/test_oracle_pin.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _life_pair(size):
    from oracle.cpu import OracleMachine
    from paraiso_b200.examples.life import life_om, life_setup
    from paraiso_b200.machines import life_machine, life_seed
    m = life_machine(size)
    o = OracleMachine(life_setup("master", size=size), life_om("master"), openmp=True, opt="-O3")
    init = life_seed(size[0], 0, size[1])
    m.call("init"); o.call("init")
    m.set("cell", init); o.interior("cell")[...] = init
    return m, o


@pytest.mark.parametrize("size,steps", [((80, 48), 40), ((2048, 2048), 100), ((1000, 777), 25), ((5, 3), 6), ((16384, 16384), 10)])
def test_life_bit_exact(size, steps):
    m, o = _life_pair(size)
    for t in range(steps):
        m.call("proceed"); o.call("proceed")
        if t in (0, 9, steps - 1):
            assert np.array_equal(m.get("cell"), o.interior("cell")), f"cells differ at step {t}"
            assert int(m.scalar("population")) == int(o.scalar("population")[0])
    assert int(m.scalar("generation")) == steps == int(o.scalar("generation")[0])


def test_life_tma_bulk_staging_variant_bit_exact():
    """Tuning.staging = "bulk": input rows staged by cp.async.bulk (TMA, UBLKCP) completing on mbarriers."""
    from oracle.cpu import OracleMachine
    from paraiso_b200.build import build_machine
    from paraiso_b200.examples.life import life_om, life_setup
    from paraiso_b200.machines import life_seed
    from paraiso_b200.runtime import Machine
    size, steps = (1000, 777), 25
    setup = life_setup("master")
    setup.tuning.staging = "bulk"
    setup.tuning.prefetch_rows = 3
    desc, so = build_machine(setup, life_om("master"), tag="Life_CC_bulk")
    with open(so.replace("libom_Life.so", "Life_kernels.cu")) as f:
        assert "om_bulk_g2s" in f.read()
    m = Machine(desc, so, size=size)
    o = OracleMachine(life_setup("master", size=size), life_om("master"), openmp=True, opt="-O3")
    init = life_seed(size[0], 0, size[1])
    m.call("init"); o.call("init")
    m.set("cell", init); o.interior("cell")[...] = init
    for t in range(steps):
        m.call("proceed"); o.call("proceed")
    assert np.array_equal(m.get("cell"), o.interior("cell"))
    assert int(m.scalar("population")) == int(o.scalar("population")[0])


def test_life_init_pattern_gosper():
    """examples/Life/main.cpp:21-28 seeds init-pat.txt through the element accessor on the 80x48 default grid."""
    from paraiso_b200.machines import life_machine
    pat = [(0, 4), (0, 5), (1, 4), (1, 5), (10, 4), (10, 5), (10, 6), (11, 3), (11, 7), (12, 2), (12, 8), (13, 2), (13, 8),
           (14, 5), (15, 3), (15, 7), (16, 4), (16, 5), (16, 6), (17, 5), (20, 2), (20, 3), (20, 4), (21, 2), (21, 3), (21, 4),
           (22, 1), (22, 5), (24, 0), (24, 1), (24, 5), (24, 6), (34, 2), (34, 3), (35, 2), (35, 3)]
    m = life_machine((80, 48))
    m.call("init")
    c = np.zeros((48, 80), np.int32)
    for x, y in pat:
        c[y, x] = 1
    m.set("cell", c)
    ref = c.copy()
    for _ in range(60):   # independent 5-line periodic Life
        n = sum(np.roll(np.roll(ref, dy, 0), dx, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0))
        ref = (((ref == 0) & (n == 3)) | ((ref == 1) & (n >= 2) & (n <= 3))).astype(np.int32)
        m.call("proceed")
    assert np.array_equal(m.get("cell"), ref)
    assert int(m.scalar("population")) == int(ref.sum())


def _hydro_pair(size, fmad=False, fast=False):
    from oracle.cpu import OracleMachine
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    from paraiso_b200.machines import hydro_machine, hydro_set_params
    m = hydro_machine(size, fmad=fmad, fast=fast)
    o = OracleMachine(hydro_setup(size), hydro_om("master"), openmp=True, opt="-O2")
    hydro_set_params(m, size)
    for k, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / size[0], dR1=1.0 / size[1]).items():
        o.scalar(k)[0] = v
    m.call("init"); o.call("init")
    return m, o


NAMES = ["density", "velocity0", "velocity1", "pressure"]


def _conserved(get):
    rho, v0, v1, p = (get(n).astype(np.float64) for n in NAMES)
    e = 0.5 * rho * (v0 * v0 + v1 * v1) + p / (5.0 / 3.0 - 1.0)
    return [rho, rho * v0, rho * v1, e]


@pytest.mark.parametrize("size,steps", [((64, 48), 5), ((512, 512), 20), ((1024, 1024), 20), ((700, 333), 8)])
def test_hydro_exact_build_is_bit_identical(size, steps):
    """-fmad=false build: every cell of every state array equals the oracle's bit for bit."""
    m, o = _hydro_pair(size)
    for n in NAMES:   # CUDA sin vs libm sin differ in the last ulp: compare init loosely, then share ICs
        a, b = m.get(n, with_margin=True), o.array(n)
        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) < 1e-14
        m.set(n, b, with_margin=True)
    for t in range(steps):
        m.call("proceed"); o.call("proceed")
    for n in NAMES:
        a, b = m.get(n, with_margin=True), o.array(n)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), n
    assert m.scalar("time") == o.scalar("time")[0]


def test_hydro_fma_build_within_tolerance():
    """FMA-contracted build: conserved variables within 1e-12 relative (north-star tolerance, double)."""
    size, steps = (512, 512), 20
    m, o = _hydro_pair(size, fmad=True)
    for n in NAMES:
        m.set(n, o.array(n), with_margin=True)
    for t in range(steps):
        m.call("proceed"); o.call("proceed")
    ca = _conserved(lambda n: m.get(n))
    cb = _conserved(lambda n: o.interior(n))
    for a, b in zip(ca, cb):
        scale = np.max(np.abs(b))
        assert np.max(np.abs(a - b)) / scale < 1e-12
    assert abs(m.scalar("time") - o.scalar("time")[0]) / o.scalar("time")[0] < 1e-12


@pytest.mark.parametrize("size,steps", [((512, 512), 20), ((1024, 1024), 20)])
def test_hydro_fast_math_build_within_tolerance(size, steps):
    """Setup.fast_math (FMA + MUFU-seeded division / sqrt with shared reciprocals): conserved variables and
    `time` within 1e-12 relative of the oracle after 20 steps (north-star tolerance for double)."""
    m, o = _hydro_pair(size, fast=True)
    for n in NAMES:
        m.set(n, o.array(n), with_margin=True)
    for t in range(steps):
        m.call("proceed"); o.call("proceed")
    ca = _conserved(lambda n: m.get(n))
    cb = _conserved(lambda n: o.interior(n))
    for a, b in zip(ca, cb):
        assert np.max(np.abs(a - b)) / np.max(np.abs(b)) < 1e-12
    assert abs(m.scalar("time") - o.scalar("time")[0]) / o.scalar("time")[0] < 1e-12


def test_hydro_4096_three_steps():
    size, steps = (4096, 4096), 3
    m, o = _hydro_pair(size)
    for n in NAMES:
        m.set(n, o.array(n), with_margin=True)
    for t in range(steps):
        m.call("proceed"); o.call("proceed")
    for n in NAMES:
        assert np.array_equal(m.get(n, with_margin=True).view(np.uint64), o.array(n).view(np.uint64)), n


@pytest.mark.parametrize("which", ["life", "hydro", "hydro_fast"])
def test_graph_replay_equals_eager_stepping(which):
    """Machine.capture: pairs of proceed() calls replayed from a CUDA graph give the state eager stepping gives, bit for bit;
    a host write between replays invalidates the carried dt reduce baked into the fast build's graph (falls back to eager calls)."""
    from paraiso_b200.machines import hydro_machine, hydro_set_params, life_machine, life_seed
    def make():
        if which == "life":
            m = life_machine((1000, 777))
            m.call("init")
            m.set("cell", life_seed(1000, 0, 777))
            return m, ["cell"], "population"
        m = hydro_machine((700, 333), fast=which == "hydro_fast")
        hydro_set_params(m, (700, 333))
        m.call("init")
        return m, NAMES, "time"
    a, names, sc = make()
    b, _, _ = make()
    for _ in range(2):
        a.call("proceed"); b.call("proceed")
    g = a.capture("proceed", 2)
    l0 = a.launches
    for _ in range(4):
        g.replay()
        b.call("proceed"); b.call("proceed")
    assert a.launches - l0 == 4 * g.launches_per_replay > 0
    for n in names:
        assert np.array_equal(a.get(n).view(np.uint8), b.get(n).view(np.uint8)), n
    assert a.scalar(sc) == b.scalar(sc)
    # a host write in between: replay must still equal eager stepping
    for m in (a, b):
        x = m.get(names[0]); x[5:9, 7:30] = x[10:14, 7:30]; m.set(names[0], x)
    g.replay(); g.replay()
    for _ in range(4):
        b.call("proceed")
    for n in names:
        assert np.array_equal(a.get(n).view(np.uint8), b.get(n).view(np.uint8)), n
    assert a.scalar(sc) == b.scalar(sc)


def test_hydro_exact_build_with_denormal_velocities_is_bit_identical():
    """Tiny and denormal transverse velocities (what the front of a spreading perturbation looks like after a few hundred
    steps): the branch-free division hands those cells to the compiler's IEEE slow path, so the state still equals the
    oracle's bit for bit — and the slow path really ran."""
    size, steps = (256, 192), 6
    m, o = _hydro_pair(size)
    rng = np.random.default_rng(11)
    for n in NAMES:
        b = o.array(n)
        if n == "velocity1":
            mag = 10.0 ** rng.uniform(-323, -280, size=b.shape)
            b[...] = np.where(rng.random(b.shape) < 0.3, mag * rng.choice([-1.0, 1.0], size=b.shape), 0.0)
            b[::7, ::5] = 5e-324
        m.set(n, b, with_margin=True)
    for _ in range(steps):
        m.call("proceed"); o.call("proceed")
    for n in NAMES:
        assert np.array_equal(m.get(n, with_margin=True).view(np.uint64), o.array(n).view(np.uint64)), n
    assert m.scalar("time") == o.scalar("time")[0]
    assert m.slow_path_cells() > 0
