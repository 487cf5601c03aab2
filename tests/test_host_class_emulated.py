"""The generated C++ host class on a machine without a GPU: <Name>.cpp is linked against the emulated kernels and the
host-memory CUDA runtime / NCCL stand-ins (tests/emu/hostclass.py).  Checks against the oracle what tests/test_gpu_host_class.py
checks on the B200 — lazy mirrors with host writes between kernels, scalar accessors, the carried dt reduce of the
fast_math Hydro build — and that OM_B200_GPUS=N (slabs along axis 1, ghost-row send/recv, all-reduce of dt, deferred
all-reduce of the population) prints exactly what one device prints."""
import os

import numpy as np
import pytest

from tests.emu import hostclass

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


@pytest.fixture(scope="module")
def life_exe(tmp_path_factory):
    from paraiso_b200.examples.life import life_om, life_setup
    exe = str(tmp_path_factory.mktemp("life") / "life_driver")
    hostclass.link_emulated(life_setup("master"), life_om("master"), "Life_hostclass", os.path.join(CPP, "life_driver.cpp"), exe)
    return exe


def _hydro_exe(tmp_path_factory, fast, size=(64, 48)):
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    exe = str(tmp_path_factory.mktemp("hydro") / f"hydro_driver_{int(fast)}")
    setup = hydro_setup(size, fast=fast)
    hostclass.link_emulated(setup, hydro_om("master"), f"Hydro_hostclass_{int(fast)}", os.path.join(CPP, "hydro_driver.cpp"), exe)
    return exe


@pytest.fixture(scope="module")
def hydro_exe(tmp_path_factory):
    return _hydro_exe(tmp_path_factory, False)


@pytest.fixture(scope="module")
def hydro_fast_exe(tmp_path_factory):
    return _hydro_exe(tmp_path_factory, True)


def test_life_class_matches_oracle(life_exe):
    from oracle.cpu import OracleMachine
    from paraiso_b200.examples.life import life_om, life_setup
    steps = 10
    out = hostclass.run(life_exe, [steps]).split("\n")
    W, H, gen, total, hsh = (int(v) for v in out[0].split())
    o = OracleMachine(life_setup("master"), life_om("master"))
    o.call("init")
    c = o.interior("cell")
    s = 20261017
    for y in range(H):
        for x in range(W):
            s = (s * 6364136223846793005 + 1442695040888963407) % (1 << 64)
            if (s >> 33) % 100 < 35:
                c[y, x] = 1
    pop = None
    for t in range(steps):
        o.call("proceed")
        pop = int(o.scalar("population")[0])
        if t % 4 == 3:
            o.interior("cell")[t % H, t % W] = 1
    cells = o.interior("cell")
    h = 1469598103934665603
    for v in cells.ravel():
        h = ((h ^ int(v)) * 1099511628211) % (1 << 64)
    assert (gen, total, hsh) == (steps, int(cells.sum()), h)
    assert out[1] == f"population {pop}"


def _oracle_hydro(size, steps):
    from oracle.cpu import OracleMachine
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    o = OracleMachine(hydro_setup(size), hydro_om("master"), openmp=True, opt="-O2")
    for k, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / size[0], dR1=1.0 / size[1]).items():
        o.scalar(k)[0] = v
    o.call("init")
    for _ in range(steps):
        o.call("proceed")
    return o


def _row_sum(a):
    acc = 0.0
    for v in a.ravel():
        acc += float(v)
    return acc


def test_hydro_class_is_bit_identical_to_oracle(hydro_exe):
    steps = 4
    out = hostclass.run(hydro_exe, [steps]).split()
    W, H = int(out[0]), int(out[1])
    o = _oracle_hydro((W, H), steps)
    assert float(out[2]) == float(o.scalar("time")[0])
    for col, n in ((3, "density"), (4, "pressure"), (5, "velocity0")):
        assert float(out[col]) == _row_sum(o.interior(n)), n


def test_hydro_fast_class_with_carried_reduce_within_tolerance(hydro_fast_exe):
    """fast_math build: the C++ host skips the CFL pre-pass from the second call on (carried dt reduce); sums of the
    primitives stay within 1e-12 relative of the IEEE oracle (the emulated intrinsics divide exactly, so in fact closer)."""
    steps = 5
    out = hostclass.run(hydro_fast_exe, [steps]).split()
    W, H = int(out[0]), int(out[1])
    o = _oracle_hydro((W, H), steps)
    assert abs(float(out[2]) - float(o.scalar("time")[0])) <= 1e-12 * float(o.scalar("time")[0])
    for col, n in ((3, "density"), (4, "pressure"), (5, "velocity0")):
        want = _row_sum(o.interior(n))
        assert abs(float(out[col]) - want) <= 1e-12 * abs(want), n


@pytest.mark.parametrize("devices", [2, 3])
def test_several_devices_print_what_one_device_prints(devices, life_exe, hydro_exe, hydro_fast_exe):
    assert hostclass.run(life_exe, [9], devices=devices) == hostclass.run(life_exe, [9])
    assert hostclass.run(hydro_exe, [3], devices=devices) == hostclass.run(hydro_exe, [3])
    assert hostclass.run(hydro_fast_exe, [4], devices=devices) == hostclass.run(hydro_fast_exe, [4])


@pytest.mark.parametrize("devices", [2, 3])
def test_tall_life_slabs_use_the_boundary_first_launch(devices, tmp_path):
    """Slabs of several chunks: the class launches the stage once per device in boundary-first chunk order, waits for the
    in-kernel signal (om_Life_wait_boundary) before the ghost-row send/recv, and prints what one device prints."""
    import subprocess
    from paraiso_b200.examples.life import life_om, life_setup
    exe = str(tmp_path / "life_tall")
    hostclass.link_emulated(life_setup("master", size=(64, 170)), life_om("master"), "Life_hostclass_tall", os.path.join(CPP, "life_driver.cpp"), exe)
    one = hostclass.run(exe, [6])
    env = dict(os.environ, OM_EMU_DEVICES=str(devices), OM_B200_GPUS=str(devices), OM_PRINT_EARLY="1")
    r = subprocess.run([exe, "6"], check=True, capture_output=True, text=True, env=env, timeout=600)
    assert r.stdout == one
    assert int(r.stderr.split()[-1]) >= 6, r.stderr


def test_class_rejects_more_devices_than_visible(life_exe):
    import subprocess
    r = subprocess.run([life_exe, "1"], capture_output=True, text=True, env=dict(os.environ, OM_EMU_DEVICES="2", OM_B200_GPUS="5"))
    assert r.returncode != 0 and "OM_B200_GPUS" in r.stderr


def test_rank3_class_matches_oracle(tmp_path):
    """Rank-3 machine: `cell(x, y, z)` accessors, plane-wise mirror copies, ghost planes of the Cyclic axis 2."""
    from oracle.cpu import OracleMachine
    from paraiso_b200.examples.rank3 import life3d_om
    from paraiso_b200.generator.native import Setup
    setup = Setup(local_size=(24, 10, 6), boundary=("Cyclic", "Cyclic", "Cyclic"))
    exe = str(tmp_path / "life3_driver")
    hostclass.link_emulated(setup, life3d_om(), "Life3_hostclass", os.path.join(CPP, "life3_driver.cpp"), exe)
    steps = 4
    out = hostclass.run(exe, [steps]).split("\n")
    W, H, D, gen, total, hsh = (int(v) for v in out[0].split())
    o = OracleMachine(setup, life3d_om())
    c = o.interior("cell")
    s = 20261017
    for z in range(D):
        for y in range(H):
            for x in range(W):
                s = (s * 6364136223846793005 + 1442695040888963407) % (1 << 64)
                if (s >> 33) % 100 < 30:
                    c[z, y, x] = 1
    pop = None
    for t in range(steps):
        o.call("proceed")
        pop = int(o.scalar("population")[0])
        if t == 2:
            o.interior("cell")[t % D, t % H, t % W] = 1
    cells = o.interior("cell")
    h = 1469598103934665603
    for v in cells.ravel():
        h = ((h ^ int(v)) * 1099511628211) % (1 << 64)
    assert (gen, total, hsh) == (steps, int(cells.sum()), h)
    assert out[1] == f"population {pop}"


def _diff3_oracle(setup, steps):
    from oracle.cpu import OracleMachine
    from paraiso_b200.examples.rank3 import diffusion3d_om
    o = OracleMachine(setup, diffusion3d_om())
    W, H, D = setup.local_size
    o.call("init")
    u = o.interior("u")
    u[3, 2, 1] = 0.75
    u[D - 1, H - 1, W - 1] = -0.5
    for t in range(steps):
        o.call("proceed")
        if t == 0:
            o.interior("u")[t % D, 0, 0] += 0.125
    return o


@pytest.mark.parametrize("bnd", [("Open", "Open", "Open"), ("Cyclic", "Open", "Cyclic")])
def test_rank3_slabs_along_axis2_on_several_devices(bnd, tmp_path):
    """OM_B200_GPUS=N on a rank-3 machine cuts axis 2: whole ghost planes travel, loadIndex(2) carries the slab offset,
    the Max reduce is all-reduced between the two stages.  1, 2 and 3 emulated devices print the oracle's numbers
    (every cell of the memory box, margins included), bit for bit."""
    from paraiso_b200.examples.rank3 import diffusion3d_om
    from paraiso_b200.generator.native import Setup
    setup = Setup(local_size=(20, 9, 8), boundary=bnd)
    tag = "Diff3_hostclass_" + "".join(b[0] for b in bnd)
    exe = str(tmp_path / "diff3_driver")
    hostclass.link_emulated(setup, diffusion3d_om(), tag, os.path.join(CPP, "diff3_driver.cpp"), exe)
    steps = 2
    o = _diff3_oracle(setup, steps)
    want = [f"{v:.17g}" for v in o.array("u").ravel()]
    for devices in (1, 2, 3):
        out = hostclass.run(exe, [steps], devices=devices).split("\n")
        assert out[0] == f"20 9 8 {float(o.scalar('peak')[0]):.17g}", devices
        got = out[1:1 + len(want)]
        # cells outside the Valid region of the stored value are never written by the reference (zero-constructed
        # manifest buffers): the oracle's array holds exactly what the accessor must return
        assert got == want, (devices, [i for i, (a, b) in enumerate(zip(got, want)) if a != b][:5])


@pytest.mark.parametrize("devices", [2, 3])
def test_rank3_life_on_several_devices(devices, tmp_path):
    from paraiso_b200.examples.rank3 import life3d_om
    from paraiso_b200.generator.native import Setup
    setup = Setup(local_size=(24, 10, 6), boundary=("Cyclic", "Cyclic", "Cyclic"))
    exe = str(tmp_path / "life3_driver")
    hostclass.link_emulated(setup, life3d_om(), "Life3_hostclass", os.path.join(CPP, "life3_driver.cpp"), exe)
    assert hostclass.run(exe, [4], devices=devices) == hostclass.run(exe, [4])
