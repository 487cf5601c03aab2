"""The generated C++ host class on the GPU: drivers written like the reference's examples (accessor writes,
proceed loop, accessor reads) are compiled against <Name>.hpp/.cpp + libom_<Name>.so and run; results must equal
the Python host path (same C ABI) / the oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "paraiso_b200", "_generated")
CUDA = "/usr/local/cuda"


def build_driver(name, tag, src, lib=None):
    d = os.path.join(GEN, tag)
    lib = lib or f"om_{name}"
    out = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, f"{name.lower()}_driver_{tag}")
    cxx = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-w", f"-I{d}", f"-I{CUDA}/include", os.path.join(ROOT, "tests", "cpp", src),
           os.path.join(d, f"{name}.cpp"), f"-L{d}", f"-l{lib}", f"-L{CUDA}/lib64", "-lcudart", "-lnccl", f"-Wl,-rpath,{d}", "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_life_cpp_class_matches_oracle():
    from oracle.cpu import OracleMachine
    from paraiso_b200.examples.life import life_om, life_setup
    from paraiso_b200.machines import build_life
    build_life()
    steps = 10
    exe = build_driver("Life", "Life_CC", "life_driver.cpp")
    out = subprocess.run([exe, str(steps)], check=True, capture_output=True, text=True).stdout.split("\n")
    W, H, gen, total, hsh = out[0].split()
    W, H = int(W), int(H)
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
    assert (int(gen), int(total), int(hsh)) == (steps, int(cells.sum()), h)
    assert out[1] == f"population {pop}"


@pytest.mark.parametrize("fast", [False, True])     # fast: carried dt reduce + per-node schedule genes in both hosts
def test_hydro_cpp_class_matches_python_host(fast):
    from paraiso_b200.machines import build_hydro, hydro_machine, hydro_set_params
    build_hydro(fast=fast)
    steps = 5
    exe = build_driver("Hydro", "Hydro_OO_Double_fast" if fast else "Hydro_OO_Double", "hydro_driver.cpp",
                       lib="om_Hydro_fma" if fast else None)
    out = subprocess.run([exe, str(steps)], check=True, capture_output=True, text=True).stdout.split()
    W, H = int(out[0]), int(out[1])
    m = hydro_machine((W, H), fast=fast)
    hydro_set_params(m, (W, H))
    m.call("init")
    for _ in range(steps):
        m.call("proceed")
    assert float(out[2]) == float(m.scalar("time"))
    # same kernels, same inputs -> identical arrays; the driver sums row by row in double
    for col, n in ((3, "density"), (4, "pressure"), (5, "velocity0")):
        a = m.get(n)
        acc = 0.0
        for v in a.ravel():
            acc += float(v)
        assert float(out[col]) == acc, n


def _run(exe, steps, gpus):
    env = dict(os.environ, OM_B200_GPUS=str(gpus))
    out = subprocess.run([exe, str(steps)], check=True, capture_output=True, text=True, env=env, timeout=600).stdout
    return [l for l in out.split("\n") if not l.startswith("NCCL version")]     # (banner printed when NCCL_DEBUG=VERSION)


@pytest.mark.parametrize("gpus", [2, 3, 4])
def test_cpp_classes_on_several_gpus_equal_one_gpu(gpus):
    """OM_B200_GPUS=N: the same binaries slab-decompose axis 1 over N devices (NCCL ghost-row send/recv, all-reduce of
    Hydro's dt, deferred all-reduce of Life's population).  Per-cell SSA and min / integer-sum reductions do not
    depend on the decomposition, so every printed number must be identical to the one-GPU run."""
    import torch
    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    from paraiso_b200.machines import build_hydro, build_life
    build_life(); build_hydro()
    life = build_driver("Life", "Life_CC", "life_driver.cpp")
    hydro = build_driver("Hydro", "Hydro_OO_Double", "hydro_driver.cpp")
    assert _run(life, 25, gpus) == _run(life, 25, 1)
    assert _run(hydro, 6, gpus) == _run(hydro, 6, 1)
    build_hydro(fast=True)
    hydro_fast = build_driver("Hydro", "Hydro_OO_Double_fast", "hydro_driver.cpp", lib="om_Hydro_fma")
    assert _run(hydro_fast, 6, gpus) == _run(hydro_fast, 6, 1)


@pytest.mark.parametrize("gpus", [2, 4])
def test_rank3_cpp_class_on_several_gpus_equals_one_gpu(gpus):
    """Rank-3 machines are cut along axis 2 (whole ghost planes by ncclSend/ncclRecv); same printed numbers as one GPU."""
    import torch
    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    from paraiso_b200.build import build_machine
    from paraiso_b200.examples.rank3 import life3d_om
    from paraiso_b200.generator.native import Setup
    setup = Setup(local_size=(48, 20, 12), boundary=("Cyclic", "Cyclic", "Cyclic"))
    build_machine(setup, life3d_om(), tag="Life3_host")
    exe = build_driver("Life3", "Life3_host", "life3_driver.cpp")
    assert _run(exe, 6, gpus) == _run(exe, 6, 1)


def test_cpp_class_rejects_more_gpus_than_visible():
    from paraiso_b200.machines import build_life
    build_life()
    exe = build_driver("Life", "Life_CC", "life_driver.cpp")
    r = subprocess.run([exe, "1"], capture_output=True, text=True, env=dict(os.environ, OM_B200_GPUS="64"))
    assert r.returncode != 0 and "OM_B200_GPUS" in r.stderr


def test_rank3_cpp_class_matches_oracle():
    """The generated class of a rank-3 machine: `cell(x, y, z)` accessors, plane-wise mirror copies, ghost planes."""
    from oracle.cpu import OracleMachine
    from paraiso_b200.build import build_machine
    from paraiso_b200.examples.rank3 import life3d_om
    from paraiso_b200.generator.native import Setup
    setup = Setup(local_size=(48, 20, 12), boundary=("Cyclic", "Cyclic", "Cyclic"))
    build_machine(setup, life3d_om(), tag="Life3_host")
    steps = 5
    exe = build_driver("Life3", "Life3_host", "life3_driver.cpp")
    out = subprocess.run([exe, str(steps)], check=True, capture_output=True, text=True).stdout.split("\n")
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
