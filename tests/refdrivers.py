"""The reference's own example drivers (examples/*/main.cpp), compiled UNCHANGED from where they lie under
/root/reference against (a) the generated B200 host class and (b) the oracle's reference-style class.

Test infrastructure: (b) produces the committed stdout goldens (tests/golden/make_driver_goldens.py), (a) is linked by
__graft_entry__.build() into tests/cpp/_build/ref_<key> so that the GPU box — where /root/reference does not exist —
only has to run the executables (tests/test_gpu_reference_drivers.py).  No reference source is copied into the repo.
"""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CUDA = "/usr/local/cuda"
OUT = os.path.join(ROOT, "tests", "cpp", "_build")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# examples/Life/init-pat.txt as (x, y) pairs (the Gosper-gun seed main.cpp:21-28 writes through sim.cell(x, y))
LIFE_PATTERN = [(0, 4), (0, 5), (1, 4), (1, 5), (10, 4), (10, 5), (10, 6), (11, 3), (11, 7), (12, 2), (12, 8), (13, 2), (13, 8),
                (14, 5), (15, 3), (15, 7), (16, 4), (16, 5), (16, 6), (17, 5), (20, 2), (20, 3), (20, 4), (21, 2), (21, 3), (21, 4),
                (22, 1), (22, 5), (24, 0), (24, 1), (24, 5), (24, 6), (34, 2), (34, 3), (35, 2), (35, 3)]
LIFE_FRAMES = 40       # frames of examples/Life/main.cpp's endless loop that are compared


def _machines():
    from paraiso_b200.examples.helloworld import helloworld_om, helloworld_setup
    from paraiso_b200.examples.life import life_om, life_setup
    from paraiso_b200.examples.shiftexample import shiftexample_om, shiftexample_setup
    from paraiso_b200.machines import build_helloworld, build_life, build_shiftexample
    # key: (class, reference driver, include prefix the driver expects, B200 builder, (setup, om) for the oracle class)
    return {
        "helloworld": ("TableMaker", "examples/HelloWorld/main.cpp", "dist", build_helloworld, (helloworld_setup, helloworld_om)),
        "hellogpu": ("TableMaker", "examples/HelloGPU/main.cu", "dist", build_helloworld, (helloworld_setup, helloworld_om)),
        "shift_open": ("TableMaker", "examples/ShiftExample/main.cpp", "", lambda: build_shiftexample(False),
                       (lambda: shiftexample_setup(False), shiftexample_om)),
        "shift_cyclic": ("TableMaker", "examples/ShiftExample/main.cpp", "", lambda: build_shiftexample(True),
                         (lambda: shiftexample_setup(True), shiftexample_om)),
        "life": ("Life", "examples/Life/main.cpp", "", build_life, (lambda: life_setup("master"), lambda: life_om("master"))),
    }


KEYS = ["helloworld", "hellogpu", "shift_open", "shift_cyclic", "life"]


def _cxx():
    return "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"


def _incdir(header_dir: str, prefix: str) -> str:
    """A directory from which `#include "<prefix>/<Name>.hpp"` resolves to header_dir/<Name>.hpp."""
    if not prefix:
        return header_dir
    t = os.path.join(OUT, "inc_" + os.path.basename(header_dir.rstrip("/")))
    os.makedirs(t, exist_ok=True)
    link = os.path.join(t, prefix)
    if os.path.islink(link) and os.readlink(link) != header_dir:
        os.unlink(link)
    if not os.path.islink(link):
        os.symlink(header_dir, link)
    return t


def exe_path(key: str) -> str:
    return os.path.join(OUT, f"ref_{key}")


def link_b200(key: str) -> str:
    """g++ <reference driver, unchanged> + generated <Name>.cpp + libom_<Name>.so -> tests/cpp/_build/ref_<key>."""
    name, driver, prefix, build, _ = _machines()[key]
    _desc, so = build()
    d = os.path.dirname(so)
    os.makedirs(OUT, exist_ok=True)
    exe = exe_path(key)
    cmd = [_cxx(), "-std=c++17", "-O1", "-w", f"-I{_incdir(d, prefix)}", f"-I{d}", f"-I{CUDA}/include", "-x", "c++",
           os.path.join(REF, driver), os.path.join(d, f"{name}.cpp"), "-x", "none", f"-L{d}", f"-lom_{name}", f"-L{CUDA}/lib64",
           "-lcudart", "-lnccl", f"-Wl,-rpath,{d}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{driver} does not compile against the generated class:\n{r.stderr[-3000:]}")
    return exe


def link_oracle(key: str, outdir: str) -> str:
    """The same driver against the oracle's reference-style class (oracle/plantrans.py), CPU only."""
    from oracle import plantrans
    from paraiso_b200.generator.plan import translate
    name, driver, prefix, _build, (mk_setup, mk_om) = _machines()[key]
    hdr = os.path.join(outdir, f"hdr_{key}")
    os.makedirs(hdr, exist_ok=True)
    with open(os.path.join(hdr, f"{name}.hpp"), "w") as f:
        f.write(plantrans.emit(translate(mk_setup(), mk_om())))
    exe = os.path.join(outdir, f"oracle_{key}")
    cmd = [_cxx(), "-std=c++17", "-O1", "-w", "-ffp-contract=off", f"-I{_incdir(hdr, prefix)}", f"-I{hdr}", "-x", "c++",
           os.path.join(REF, driver), "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def link_emulated(key: str, outdir: str) -> str:
    """The same driver + the generated host class, with the generated kernels on host threads (tests/emu/cuda_emu.h)
    and the CUDA runtime calls of the host class served from host memory (tests/emu/cudart): the whole drop-in path
    without a GPU.  Test infrastructure only."""
    from tests.emu.build_emu import build_emulated
    name, driver, prefix, _build, (mk_setup, mk_om) = _machines()[key]
    _desc, so = build_emulated(mk_setup(), mk_om(), tag=f"{name}_ref_{key}")
    d = os.path.dirname(so)
    exe = os.path.join(outdir, f"emu_{key}")
    cmd = [_cxx(), "-std=c++20", "-O1", "-w", "-DOM_B200_NO_NCCL", f"-I{os.path.join(ROOT, 'tests', 'emu', 'cudart')}",
           f"-I{_incdir(d, prefix)}", f"-I{d}", "-x", "c++", os.path.join(REF, driver), os.path.join(d, f"{name}.cpp"), "-x", "none",
           so, f"-Wl,-rpath,{d}", "-pthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-3000:])
    return exe


def nosleep_shim() -> str:
    """LD_PRELOAD library that turns usleep() into a no-op: examples/Life/main.cpp sleeps up to 1 s between frames."""
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libnosleep.so")
    if not os.path.exists(so):
        src = "int usleep(unsigned int usec) { (void)usec; return 0; }\n"
        subprocess.run(["gcc", "-shared", "-fPIC", "-x", "c", "-", "-o", so], input=src, text=True, check=True)
    return so


def life_frames(text: str, rows: int = 48) -> str:
    """The first LIFE_FRAMES complete frames of examples/Life/main.cpp's output (rows/2 picture lines, a blank line and
    the generation / population line per frame).  main.cpp sizes its frame buffer for all rows but fills every other
    one, so each frame carries (W+1)*H/2 NUL bytes: they are dropped here, on both sides of the comparison."""
    per = rows // 2 + 2
    lines = text.replace("\0", "").split("\n")
    n = min(LIFE_FRAMES, len(lines) // per)
    return "\n".join(lines[: n * per]) + "\n"


def run(key: str, exe: str, seconds: float = 120.0) -> str:
    """stdout of one driver run (Life's endless loop is stopped once LIFE_FRAMES frames are out, `seconds` at the latest)."""
    if key != "life":
        return subprocess.run([exe], check=True, capture_output=True, text=True, timeout=seconds).stdout
    import time
    with tempfile.TemporaryDirectory() as cwd:
        with open(os.path.join(cwd, "init-pat.txt"), "w") as f:
            f.write("".join(f"{x} {y}\n" for x, y in LIFE_PATTERN))
        env = dict(os.environ, LD_PRELOAD=nosleep_shim())
        path = os.path.join(cwd, "stdout.txt")
        need = LIFE_FRAMES * (48 // 2 + 2) + 1
        with open(path, "w") as out:
            p = subprocess.Popen([exe], cwd=cwd, env=env, stdout=out, stderr=subprocess.PIPE)
            t0 = time.time()
            try:
                while time.time() - t0 < seconds:
                    if p.poll() is not None:
                        raise RuntimeError(f"Life driver exited by itself ({p.returncode}): {p.stderr.read()[-2000:]}")
                    with open(path) as f:
                        if f.read().count("\n") >= need:
                            break
                    time.sleep(0.05)
            finally:
                if p.poll() is None:
                    p.kill()
                    p.wait()
        with open(path) as f:
            return life_frames(f.read())


# ---- examples/Hydro/main-kh.cpp: integrates 1024^2 to t = 1 and writes 101 snapshots; the first one is compared --------------
HYDRO_TIMES = 4        # lines of the driver's stderr (sim.time() before each proceed()) that are compared


def _hydro_setup_om():
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    return hydro_setup(), hydro_om("master")


def link_b200_hydro() -> str:
    from paraiso_b200.machines import build_hydro
    _desc, so = build_hydro()
    d = os.path.dirname(so)
    os.makedirs(OUT, exist_ok=True)
    exe = exe_path("hydro")
    cmd = [_cxx(), "-std=c++17", "-O1", "-w", f"-I{d}", f"-I{CUDA}/include", os.path.join(REF, "examples/Hydro/main-kh.cpp"),
           os.path.join(d, "Hydro.cpp"), f"-L{d}", "-lom_Hydro", f"-L{CUDA}/lib64", "-lcudart", "-lnccl", f"-Wl,-rpath,{d}", "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def link_oracle_hydro(outdir: str) -> str:
    from oracle import plantrans
    from paraiso_b200.generator.plan import translate
    setup, om = _hydro_setup_om()
    hdr = os.path.join(outdir, "hdr_hydro")
    os.makedirs(hdr, exist_ok=True)
    with open(os.path.join(hdr, "Hydro.hpp"), "w") as f:
        f.write(plantrans.emit(translate(setup, om)))
    exe = os.path.join(outdir, "oracle_hydro")
    subprocess.run([_cxx(), "-std=c++17", "-O2", "-fopenmp", "-w", "-ffp-contract=off", f"-I{hdr}",
                    os.path.join(REF, "examples/Hydro/main-kh.cpp"), "-o", exe], check=True)
    return exe


def run_hydro(exe: str, seconds: float = 600.0) -> dict:
    """Run main-kh.cpp until it has printed HYDRO_TIMES + 1 times (the first snapshot, written after the second proceed(),
    is complete by then), stop it, return the printed times and a digest of output1/snapshot0000.txt (1024^2 lines
    `x y density velocity0 velocity1 pressure`, six significant digits): column sums and the cells (64 k, 64 k)."""
    import time
    with tempfile.TemporaryDirectory() as cwd:
        err_path = os.path.join(cwd, "stderr.txt")
        with open(err_path, "w") as err:
            p = subprocess.Popen([exe], cwd=cwd, stdout=subprocess.DEVNULL, stderr=err)
            t0 = time.time()
            try:
                while time.time() - t0 < seconds:
                    if p.poll() is not None:
                        raise RuntimeError(f"Hydro driver exited by itself ({p.returncode}): {open(err_path).read()[-2000:]}")
                    with open(err_path) as f:
                        if f.read().count("\n") >= HYDRO_TIMES + 1:
                            break
                    time.sleep(0.05)
                else:
                    raise RuntimeError("Hydro driver too slow")
            finally:
                if p.poll() is None:
                    p.kill()
                    p.wait()
        with open(err_path) as f:
            times = f.read().split("\n")[:HYDRO_TIMES]
        import numpy as np
        with open(os.path.join(cwd, "output1", "snapshot0000.txt")) as f:
            a = np.fromstring(f.read(), sep=" ").reshape(-1, 6)      # x y density velocity0 velocity1 pressure
        n = int(round(len(a) ** 0.5))
        # (no hash of the text: the driver's init() evaluates sin on the device, <= 2 ulp from libm's, and the near-zero
        #  velocities are printed with six digits — the comparison is numeric, see tests/test_gpu_reference_drivers.py)
        return dict(times=times, cells=len(a), column_sums=[float(v) for v in a.sum(axis=0)],
                    abs_sums=[float(v) for v in np.abs(a).sum(axis=0)],
                    diagonal=[[float(v) for v in a[i * n + i]] for i in range(0, n, 64)])


# ---- examples/InitialCondition/main.cpp: writes heart.txt (`x y table(x,y)` for 500 x 500 cells) -------------------------------
def _heart():
    from paraiso_b200.examples.initialcondition import initialcondition_om, initialcondition_setup
    return "TableMaker", "examples/InitialCondition/main.cpp", "dist", initialcondition_setup, initialcondition_om


def link_heart(kind: str, outdir: str = None) -> str:
    """kind = "b200" (real class + libom), "oracle" (reference-style class) or "emulated" (generated class, emulated kernels)."""
    name, driver, prefix, mk_setup, mk_om = _heart()
    if kind == "b200":
        from paraiso_b200.build import build_machine
        _desc, so = build_machine(mk_setup(), mk_om(), tag="Heart_OO")
        d = os.path.dirname(so)
        os.makedirs(OUT, exist_ok=True)
        exe = exe_path("initialcondition")
        cmd = [_cxx(), "-std=c++17", "-O1", "-w", f"-I{_incdir(d, prefix)}", f"-I{d}", f"-I{CUDA}/include", os.path.join(REF, driver),
               os.path.join(d, f"{name}.cpp"), f"-L{d}", f"-lom_{name}", f"-L{CUDA}/lib64", "-lcudart", "-lnccl", f"-Wl,-rpath,{d}", "-o", exe]
    elif kind == "oracle":
        from oracle import plantrans
        from paraiso_b200.generator.plan import translate
        hdr = os.path.join(outdir, "hdr_heart")
        os.makedirs(os.path.join(hdr, prefix), exist_ok=True)
        with open(os.path.join(hdr, prefix, f"{name}.hpp"), "w") as f:
            f.write(plantrans.emit(translate(mk_setup(), mk_om())))
        exe = os.path.join(outdir, "oracle_heart")
        cmd = [_cxx(), "-std=c++17", "-O1", "-w", "-ffp-contract=off", f"-I{hdr}", os.path.join(REF, driver), "-o", exe]
    else:
        from tests.emu.build_emu import build_emulated
        _desc, so = build_emulated(mk_setup(), mk_om(), tag="Heart_ref")
        d = os.path.dirname(so)
        exe = os.path.join(outdir, "emu_heart")
        cmd = [_cxx(), "-std=c++20", "-O1", "-w", "-DOM_B200_NO_NCCL", f"-I{os.path.join(ROOT, 'tests', 'emu', 'cudart')}",
               f"-I{_incdir(d, prefix)}", f"-I{d}", "-x", "c++", os.path.join(REF, driver), os.path.join(d, f"{name}.cpp"), "-x", "none",
               so, f"-Wl,-rpath,{d}", "-pthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-3000:])
    return exe


def run_heart(exe: str) -> str:
    with tempfile.TemporaryDirectory() as cwd:
        subprocess.run([exe], cwd=cwd, check=True, timeout=300)
        with open(os.path.join(cwd, "heart.txt")) as f:
            return f.read()


def heart_digest(text: str) -> dict:
    import numpy as np
    a = np.fromstring(text, sep=" ").reshape(-1, 3)
    return dict(cells=len(a), column_sums=[float(v) for v in a.sum(axis=0)], abs_sum=float(np.abs(a[:, 2]).sum()),
                samples=[[float(v) for v in a[i]] for i in range(0, len(a), 5003)])


def golden_path(key: str) -> str:
    return os.path.join(GOLDEN, f"driver_{'helloworld' if key == 'hellogpu' else key}.txt")
