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
    t = tempfile.mkdtemp(prefix="om_inc_")
    os.symlink(header_dir, os.path.join(t, prefix))
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


def golden_path(key: str) -> str:
    return os.path.join(GOLDEN, f"driver_{'helloworld' if key == 'hellogpu' else key}.txt")
