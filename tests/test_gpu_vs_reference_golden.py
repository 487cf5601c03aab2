"""GPU results against the output of the reference's OWN generated C++ (examples-old/*-exampled/dist), through the
golden fixtures in tests/golden (made by tests/golden/make_golden.py from oracle/_ref): the `exampled` programs run on
the B200 backend must reproduce them bit for bit.  Plus master's Hydro in float against the oracle (1e-5 tolerance
of the north star; the -fmad=false build is in fact bit-identical)."""
import json
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["density", "velocity0", "velocity1", "pressure"]


def test_life_exampled_equals_reference_output():
    from paraiso_b200.machines import build_life_exampled
    from paraiso_b200.runtime import Machine
    g = np.load(os.path.join(GOLD, "life_exampled.npz"))
    desc, so = build_life_exampled()
    m = Machine(desc, so)                       # 128x128, Open, margin 1
    m.call("init")                               # R-pentomino via loadIndex / loadSize, population reduce
    assert int(m.scalar("population")) == int(g["populations"][0])
    shape = tuple(g["shape"])
    for t in range(1, 101):
        m.call("proceed")
        assert int(m.scalar("population")) == int(g["populations"][t]), t
        if t in (1, 10, 100):
            want = np.unpackbits(g[f"cell_gen{t}"])[:shape[0] * shape[1]].reshape(shape)
            assert np.array_equal(m.get("cell", with_margin=True), want), t
    assert int(m.scalar("generation")) == 100 and int(m.scalar("population")) == 121


def test_hydro_exampled_float_equals_reference_output():
    from oracle.cpu import OracleMachine
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    from paraiso_b200.machines import build_hydro_exampled
    from paraiso_b200.runtime import Machine
    with open(os.path.join(GOLD, "hydro_exampled.json")) as f:
        g = json.load(f)
    desc, so = build_hydro_exampled()
    m = Machine(desc, so)                       # float, 1024x1024, Open, margin 3
    one = np.float32(1.0)
    params = dict(time=np.float32(0), cfl=np.float32(0.5), extent0=one, extent1=one,
                  dR0=one / np.float32(1024), dR1=one / np.float32(1024))
    for k, v in params.items():
        m.set_scalar(k, v)
    # initial condition: the reference evaluates `sin` in libm, CUDA in its own library; take the (pinned) oracle's
    # init, check it is the reference's (CRC), and upload it
    o = OracleMachine(hydro_setup((1024, 1024)), hydro_om("exampled"), openmp=True)
    for k, v in params.items():
        o.scalar(k)[0] = v
    o.call("init")
    m.call("init")
    for n in NAMES:
        assert zlib.crc32(o.array(n).tobytes()) == g["init_crc32"][n]
        a = m.get(n, with_margin=True)
        assert np.max(np.abs(a - o.array(n)) / np.maximum(np.abs(o.array(n)), 1e-30)) < 1e-6    # GPU init: sinf within an ulp
        m.set(n, o.array(n), with_margin=True)
    for t in range(1, 11):
        m.call("proceed")
        if str(t) in g["steps"]:
            s = g["steps"][str(t)]
            assert int(np.float32(m.scalar("time")).view(np.uint32)) == s["time_bits"], t
            for n in NAMES:
                assert zlib.crc32(np.ascontiguousarray(m.get(n, with_margin=True)).tobytes()) == s["crc32"][n], (t, n)


def test_hydro_master_float_within_1e5():
    from oracle.cpu import OracleMachine
    from paraiso_b200.build import build_machine
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    from paraiso_b200.runtime import Machine
    size, steps = (256, 200), 20
    desc, so = build_machine(hydro_setup(), hydro_om("master", real="Float"), tag="Hydro_OO_Float")
    m = Machine(desc, so, size=size)
    o = OracleMachine(hydro_setup(size), hydro_om("master", real="Float"), openmp=True)
    for k, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / size[0], dR1=1.0 / size[1]).items():
        m.set_scalar(k, np.float32(v))
        o.scalar(k)[0] = np.float32(v)
    o.call("init")
    for n in NAMES:
        m.set(n, o.array(n), with_margin=True)
    for _ in range(steps):
        m.call("proceed"); o.call("proceed")
    for n in NAMES:
        a, b = m.get(n).astype(np.float64), o.interior(n).astype(np.float64)
        assert np.max(np.abs(a - b)) / np.max(np.abs(b)) < 1e-5


def test_initialcondition_transcendentals_on_gpu():
    """examples/InitialCondition (cast, ^, **, atan): CUDA's exp/log/atan against libm, 1e-12."""
    from oracle.cpu import OracleMachine
    from paraiso_b200.build import build_machine
    from paraiso_b200.examples.initialcondition import initialcondition_om, initialcondition_setup
    from paraiso_b200.runtime import Machine
    setup = initialcondition_setup()
    desc, so = build_machine(setup, initialcondition_om(), tag="Heart_OO")
    m = Machine(desc, so)
    o = OracleMachine(setup, initialcondition_om())
    m.call("create"); o.call("create")
    a, b = m.get("table"), o.array("table")
    ok = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), ok)
    assert np.max(np.abs(a[ok] - b[ok])) < 1e-12


def test_hydro_master_float_on_gpu_against_the_reference_output_directly():
    """The master program built in float for the B200, with its own init kernel, against cells sampled from the reference's
    compiled Hydro.cpp (tests/golden/hydro_exampled_samples.npz): conserved variables within 1e-5 after 3 and 10 steps —
    a pin of master's Hydro that does not pass through the oracle's front-end."""
    import os
    from paraiso_b200.build import build_machine
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    from paraiso_b200.runtime import Machine
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hydro_exampled_samples.npz"))
    size = (1024, 1024)
    desc, so = build_machine(hydro_setup(), hydro_om("master", real="Float"), tag="Hydro_OO_Float")
    m = Machine(desc, so, size=size)
    one = np.float32(1.0)
    for k, v in dict(time=np.float32(0), cfl=np.float32(0.5), extent0=one, extent1=one, dR0=one / np.float32(1024), dR1=one / np.float32(1024)).items():
        m.set_scalar(k, v)
    m.call("init")

    def cons(rho, u, v, p):
        rho, u, v, p = (x.astype(np.float64) for x in (rho, u, v, p))
        return [rho, rho * u, rho * v, p / (5.0 / 3.0 - 1.0) + 0.5 * rho * (u * u + v * v)]
    for t in range(1, 11):
        m.call("proceed")
        if t in (3, 10):
            got = cons(*[m.get(n)[::8, ::8] for n in NAMES])
            want = cons(*[g[f"{n}_step{t}"] for n in NAMES])
            for a, b in zip(got, want):
                assert np.max(np.abs(a - b)) <= 1e-5 * np.max(np.abs(b)), t
            assert abs(float(m.scalar("time")) - float(g[f"time_step{t}"][0])) <= 1e-5 * float(g[f"time_step{t}"][0])
