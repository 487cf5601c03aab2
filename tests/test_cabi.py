"""The C-ABI libraries load and export every symbol that include/*.h declares (no compute: no GPU needed),
and the generated host class keeps the reference's public surface: the reference's own drivers compile and
link against it unchanged (only where /root/reference is mounted)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "paraiso_b200", "_generated")
LIBS = {"Life": os.path.join(GEN, "Life_CC", "libom_Life.so"), "Hydro": os.path.join(GEN, "Hydro_OO_Double", "libom_Hydro.so")}


def declared(header):
    with open(os.path.join(ROOT, "include", header)) as f:
        return re.findall(r"^int (om_\w+)\(", f.read(), flags=re.M)


@pytest.mark.parametrize("name", ["Life", "Hydro"])
def test_library_exports_every_declared_symbol(name):
    if not os.path.exists(LIBS[name]):
        import __graft_entry__
        __graft_entry__.build()
    syms = declared(f"om_{name}_abi.h")
    assert len(syms) >= 5
    try:
        lib = ctypes.CDLL(LIBS[name])
    except OSError as e:   # libcudart missing on a machine without the CUDA runtime
        pytest.skip(f"cannot load {LIBS[name]}: {e}")
    for s in syms:
        assert hasattr(lib, s), s
    assert getattr(lib, f"om_{name}_abi_version")() == 3
    assert f"om_{name}_wait_boundary" in syms


def test_header_is_plain_c(tmp_path):
    """The umbrella header compiles as C99, and its constants are the ones the hosts use (a real executable, not -fsyntax-only)."""
    from paraiso_b200.runtime import APRON, OmGeom
    src = ('#include "paraiso_b200.h"\n#include <stdio.h>\n'
           'int main(void) { OmGeomC g; g.bfirst = 0; (void)g; printf("%d %d\\n", OM_APRON_ROWS, (int)sizeof(OmGeomC)); return 0; }\n')
    exe = str(tmp_path / "hdr.out")
    r = subprocess.run(["gcc", "-std=c99", "-x", "c", "-I", os.path.join(ROOT, "include"), "-", "-o", exe], input=src,
                       text=True, capture_output=True)
    assert r.returncode == 0, r.stderr
    apron, geom = map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split())
    import ctypes as ct
    assert apron == APRON == 32 and geom == ct.sizeof(OmGeom)


def test_machine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from paraiso_b200.machines import build_life
    from paraiso_b200.runtime import Machine
    desc, so = build_life()
    with pytest.raises(RuntimeError):
        Machine(desc, so, device="cpu")           # no CPU fallback
    with pytest.raises(Exception):
        Machine(desc, so, device="cuda")          # and no silent one either


REF = "/root/reference"
DRIVERS = [("Life", "Life_CC", "examples/Life/main.cpp"), ("Hydro", "Hydro_OO_Double", "examples/Hydro/main-kh.cpp")]


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not mounted")
@pytest.mark.parametrize("name,tag,driver", DRIVERS)
def test_reference_driver_compiles_and_links_unchanged(name, tag, driver, tmp_path):
    d = os.path.join(GEN, tag)
    cuda = "/usr/local/cuda"
    exe = str(tmp_path / "main.out")
    cmd = ["/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++", "-std=c++17", "-O1", "-w", f"-I{d}", f"-I{cuda}/include",
           os.path.join(REF, driver), os.path.join(d, f"{name}.cpp"), f"-L{d}", f"-lom_{name}", f"-L{cuda}/lib64", "-lcudart", "-lnccl",
           f"-Wl,-rpath,{d}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_graph_capture_and_host_pipeline_need_a_cuda_device():
    """Machine.capture / hostio.HostPipeline are CUDA-only plumbing: on an emulated (CPU) machine they refuse instead of degrading."""
    import torch
    from paraiso_b200.examples.life import life_om, life_setup
    from paraiso_b200.runtime import Machine
    from tests.emu.build_emu import build_emulated
    desc, so = build_emulated(life_setup("master", size=(64, 48)), life_om("master"), tag="Life_ring_1_cp_async")
    m = Machine(desc, so, size=(64, 48), device="cpu", _emulated=True)
    with pytest.raises(ValueError):
        m.capture("proceed", 3)                       # odd: the buffers would not be back in place
    with pytest.raises(RuntimeError):
        m.capture("proceed", 2)                       # no CUDA device
    assert m.slow_path_cells() == 0 and m.early_exchanges == 0
    if not torch.cuda.is_available():
        from paraiso_b200.hostio import HostPipeline
        with pytest.raises(Exception):
            HostPipeline(m, "proceed", ["cell"])
