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
