"""TEST INFRASTRUCTURE ONLY: link a C++ driver + the generated host class <Name>.cpp against the emulated kernels
(tests/emu/cuda_emu.h) and the host-memory CUDA runtime / NCCL stand-ins (tests/emu/cudart/), so the host class —
mirrors, dirty tracking, stage geometry, carried reduces, slab decomposition with ghost-row exchange — runs on a
machine without a GPU.  OM_EMU_DEVICES=N makes N "devices" visible (OM_B200_GPUS=N then cuts N slabs)."""
import os
import subprocess

from .build_emu import build_emulated

HERE = os.path.dirname(os.path.abspath(__file__))


def link_emulated(setup, om, tag: str, driver: str, exe: str, include_dirs=()):
    desc, so = build_emulated(setup, om, tag=tag)
    d = os.path.dirname(so)
    cxx = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"
    cmd = [cxx, "-std=c++20", "-O1", "-w", f"-I{os.path.join(HERE, 'cudart')}", *[f"-I{i}" for i in include_dirs], f"-I{d}", "-x", "c++",
           driver, os.path.join(d, f"{desc['name']}.cpp"), "-x", "none", so, f"-Wl,-rpath,{d}", "-pthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-3000:])
    return desc, so


def run(exe: str, args=(), devices: int = 1, timeout: float = 600.0) -> str:
    env = dict(os.environ, OM_EMU_DEVICES=str(devices), OM_B200_GPUS=str(devices))
    return subprocess.run([exe, *map(str, args)], check=True, capture_output=True, text=True, env=env, timeout=timeout).stdout


def link_tsan(setup, om, tag: str, driver: str, exe: str, drop_barrier: int = None, kernel_marker: str = None,
              sanitizer: str = "thread", apron: int = None):
    """The same link with ThreadSanitizer over the emulated kernels (CUDA threads are host threads, __syncthreads is a
    std::barrier): a missing barrier between a shared-memory ring's writers and readers is a reported data race.
    `drop_barrier` = n removes the n-th `__syncthreads();` after `kernel_marker` in the generated kernel source — the
    negative control that shows the detector sees the rings.  `sanitizer="address"` builds the same executable with
    AddressSanitizer instead: "device" memory is malloc'ed by the CUDA runtime stand-in, so a kernel that reads or writes
    outside an array's allocation (beyond the OM_APRON_ROWS slack the ABI promises) is reported.
    `apron` overrides the number of slack rows the host class allocates (negative control of the memory check).
    Returns None when the toolchain has no sanitizer runtime."""
    desc, so = build_emulated(setup, om, tag=tag)
    d = os.path.dirname(so)
    name = desc["name"]
    cu = os.path.join(d, f"{name}_kernels.cu")
    work = os.path.dirname(exe)
    if drop_barrier is not None:
        with open(cu) as f:
            lines = f.read().split("\n")
        start = next(i for i, l in enumerate(lines) if kernel_marker in l)
        sites = [i for i in range(start, len(lines)) if lines[i].strip() == "__syncthreads();"]
        lines[sites[drop_barrier]] = "    /* barrier removed by the negative control */"
        cu = os.path.join(work, f"{name}_dropped{drop_barrier}.cu")
        with open(cu, "w") as f:
            f.write("\n".join(lines))
    cxx = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"
    common = [cxx, "-std=c++20", "-O1", "-g", f"-fsanitize={sanitizer}", "-w", "-pthread"]
    obj = exe + "_kernels.o"
    r = subprocess.run(common + ["-c", "-include", os.path.join(HERE, "cuda_emu.h"), "-I", d, "-x", "c++", cu, "-o", obj],
                       capture_output=True, text=True)
    if r.returncode != 0:
        if "tsan" in r.stderr.lower() or "sanitize" in r.stderr.lower():
            return None
        raise RuntimeError(r.stderr[-3000:])
    host_cpp = os.path.join(d, f"{name}.cpp")
    if apron is not None:
        with open(host_cpp) as f:
            text = f.read()
        assert "const int APRON = 32;" in text
        host_cpp = os.path.join(work, f"{name}_apron{apron}.cpp")
        with open(host_cpp, "w") as f:
            f.write(text.replace("const int APRON = 32;", f"const int APRON = {apron};"))
    r = subprocess.run(common + [f"-I{os.path.join(HERE, 'cudart')}", f"-I{d}", driver, host_cpp, obj, "-o", exe],
                       capture_output=True, text=True)
    if r.returncode != 0:
        if "tsan" in r.stderr.lower() or "sanitize" in r.stderr.lower():
            return None
        raise RuntimeError(r.stderr[-3000:])
    return exe


def run_tsan(exe: str, args=(), timeout: float = 900.0, devices: int = 1):
    """(stdout, number of ThreadSanitizer / AddressSanitizer reports)."""
    env = dict(os.environ, OM_EMU_DEVICES=str(devices), OM_B200_GPUS=str(devices),
               TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0", ASAN_OPTIONS="detect_leaks=0 exitcode=0 halt_on_error=0")
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, env=env, timeout=timeout)
    return r.stdout, r.stderr.count("WARNING: ThreadSanitizer") + r.stderr.count("ERROR: AddressSanitizer")
