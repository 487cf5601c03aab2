"""Pin the oracle (oracle/plantrans.py, the restatement of the reference generator) against
(a) the committed golden fixtures produced from the reference's OWN generated C++ (tests/golden), and
(b) that code itself (oracle/_ref/*.so, built from /root/reference by oracle/Makefile) when present.
Reference: examples-old/Life-exampled/dist/Life.cpp, examples-old/Hydro-exampled/dist/Hydro.cpp."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle.cpu import REFDIR, OracleMachine, RefHydro, RefLife
from paraiso_b200.examples.hydro import hydro_om, hydro_setup
from paraiso_b200.examples.life import life_om, life_setup

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
have_ref = os.path.exists(os.path.join(REFDIR, "libref_life.so")) and os.path.exists(os.path.join(REFDIR, "libref_hydro_omp.so"))
NAMES = ["density", "velocity0", "velocity1", "pressure"]


def test_life_oracle_matches_golden():
    g = np.load(os.path.join(GOLD, "life_exampled.npz"))
    m = OracleMachine(life_setup("exampled"), life_om("exampled"))
    m.call("init")
    assert m.scalar("population")[0] == g["populations"][0]
    shape = tuple(g["shape"])
    for t in range(1, 101):
        m.call("proceed")
        assert m.scalar("population")[0] == g["populations"][t], t
        if t in (1, 10, 100):
            want = np.unpackbits(g[f"cell_gen{t}"])[:shape[0] * shape[1]].reshape(shape)
            assert np.array_equal(m.array("cell"), want), t
    # SURVEY §4 known answer: generation 100 -> population 121
    assert m.scalar("generation")[0] == 100 and m.scalar("population")[0] == 121


@pytest.mark.skipif(not have_ref, reason="oracle/_ref not built (needs /root/reference)")
def test_life_oracle_matches_reference_code_every_step():
    m = OracleMachine(life_setup("exampled"), life_om("exampled"))
    r = RefLife()
    m.call("init"); r.init()
    assert np.array_equal(m.array("cell"), r.cell())
    for t in range(300):
        m.call("proceed"); r.proceed()
        assert np.array_equal(m.array("cell"), r.cell()), t
        assert m.scalar("population")[0] == r.population()
    assert m.scalar("generation")[0] == r.generation() == 300


def _hydro_oracle():
    m = OracleMachine(hydro_setup((1024, 1024)), hydro_om("exampled"), openmp=True)
    m.scalar("time")[0] = 0
    m.scalar("cfl")[0] = 0.5
    m.scalar("extent0")[0] = 1.0
    m.scalar("extent1")[0] = 1.0
    m.scalar("dR0")[0] = np.float32(1.0) / np.float32(1024)
    m.scalar("dR1")[0] = np.float32(1.0) / np.float32(1024)
    return m


def test_hydro_oracle_matches_golden_bit_exact():
    """float, 1024^2, 10 steps: CRC32 of every state array and the bits of `time` equal the reference's."""
    with open(os.path.join(GOLD, "hydro_exampled.json")) as f:
        g = json.load(f)
    m = _hydro_oracle()
    m.call("init")
    for n in NAMES:
        assert zlib.crc32(m.array(n).tobytes()) == g["init_crc32"][n], n
    for t in range(1, 11):
        m.call("proceed")
        if str(t) in g["steps"]:
            s = g["steps"][str(t)]
            assert int(m.scalar("time").view(np.uint32)[0]) == s["time_bits"], t
            for n in NAMES:
                assert zlib.crc32(m.array(n).tobytes()) == s["crc32"][n], (t, n)
    # SURVEY §4 known answers after 10 steps
    assert abs(float(m.scalar("time")[0]) - 0.00073982001) < 1e-10
    assert abs(m.interior("density").astype(np.float64).sum() - 29127872.4) < 0.1
    assert abs(m.interior("pressure").astype(np.float64).sum() - 637099.418) < 0.01


@pytest.mark.skipif(not have_ref, reason="oracle/_ref not built (needs /root/reference)")
def test_hydro_oracle_matches_reference_code_bit_exact():
    m = _hydro_oracle()
    r = RefHydro(openmp=True)
    r.setup_kh()
    m.call("init"); r.init()
    for t in range(3):
        m.call("proceed"); r.proceed()
        for n in NAMES:
            assert np.array_equal(m.array(n).view(np.uint32), r.array(n).view(np.uint32)), (t, n)
        assert m.scalar("time")[0] == r.scalar("time")[0]


def test_master_and_exampled_hydro_agree_in_double():
    """The two program revisions differ only in typing/annotations (HydroMain.hs diff), not in arithmetic."""
    size = (48, 40)
    a = OracleMachine(hydro_setup(size), hydro_om("master"))
    b = OracleMachine(hydro_setup(size), hydro_om("exampled", real="Double"))
    for o in (a, b):
        for k, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / size[0], dR1=1.0 / size[1]).items():
            o.scalar(k)[0] = v
        o.call("init")
    for t in range(4):
        a.call("proceed"); b.call("proceed")
    for n in NAMES:
        assert np.array_equal(a.array(n).view(np.uint64), b.array(n).view(np.uint64)), n


def test_master_and_exampled_hydro_agree_in_double_256x256_20_steps():
    """The longer version of the check above (every array bit for bit, 256x256, 20 steps): the program the BASELINE
    configs run (master, double) is the pinned exampled program with other typing, not a second transcription that could
    drift from it."""
    size = (256, 256)
    a = OracleMachine(hydro_setup(size), hydro_om("master"), openmp=True)
    b = OracleMachine(hydro_setup(size), hydro_om("exampled", real="Double"), openmp=True)
    for o in (a, b):
        for k, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / size[0], dR1=1.0 / size[1]).items():
            o.scalar(k)[0] = v
        o.call("init")
    for t in range(20):
        a.call("proceed"); b.call("proceed")
        if t in (0, 9, 19):
            for n in NAMES:
                assert np.array_equal(a.array(n).view(np.uint64), b.array(n).view(np.uint64)), (t, n)
            assert a.scalar("time")[0] == b.scalar("time")[0]


def _conserved32(rho, u, v, p):
    rho, u, v, p = (x.astype(np.float64) for x in (rho, u, v, p))
    return [rho, rho * u, rho * v, p / (5.0 / 3.0 - 1.0) + 0.5 * rho * (u * u + v * v)]


def test_master_float_program_against_the_reference_output_directly():
    """Decoupled pin for master's Hydro: the MASTER transcription (examples/hydro.py "master", Real = Float), run through the
    oracle with its own init, against cells sampled from the reference's compiled Hydro.cpp (tests/golden/
    hydro_exampled_samples.npz, every 8th interior cell after 3 and 10 steps) — conserved variables within the north
    star's 1e-5.  A slip in the master program could not hide behind the front-end it shares with the oracle."""
    g = np.load(os.path.join(GOLD, "hydro_exampled_samples.npz"))
    size = (1024, 1024)
    o = OracleMachine(hydro_setup(size), hydro_om("master", real="Float"), openmp=True)
    one = np.float32(1.0)
    for k, v in dict(time=np.float32(0), cfl=np.float32(0.5), extent0=one, extent1=one, dR0=one / np.float32(1024), dR1=one / np.float32(1024)).items():
        o.scalar(k)[0] = v
    o.call("init")
    for t in range(1, 11):
        o.call("proceed")
        if t in (3, 10):
            got = _conserved32(*[o.interior(n)[::8, ::8] for n in NAMES])
            want = _conserved32(*[g[f"{n}_step{t}"] for n in NAMES])
            for a, b in zip(got, want):
                assert np.max(np.abs(a - b)) <= 1e-5 * np.max(np.abs(b)), t
            assert abs(float(o.scalar("time")[0]) - float(g[f"time_step{t}"][0])) <= 1e-5 * float(g[f"time_step{t}"][0])
