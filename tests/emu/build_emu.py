"""TEST INFRASTRUCTURE ONLY: compile a generated <Name>_kernels.cu with g++ against
tests/emu/cuda_emu.h (CUDA threads -> host threads) so schedule logic can be checked on a
machine without a GPU.  Never imported by paraiso_b200/."""
import hashlib
import os
import subprocess

from paraiso_b200.build import generate_to

HERE = os.path.dirname(os.path.abspath(__file__))


def build_emulated(setup, om, tag=None, vnt=None):
    desc, d = generate_to(setup, om, tag=(tag or f"{om.name}_{''.join(b[0] for b in setup.boundary)}") + "_emu", vnt=vnt)
    cu = os.path.join(d, f"{desc['name']}_kernels.cu")
    so = os.path.join(d, f"libemu_{desc['name']}.so")
    stamp = so + ".sha1"
    with open(cu, "rb") as f, open(os.path.join(HERE, "cuda_emu.h"), "rb") as e, open(os.path.join(d, "om_runtime.cuh"), "rb") as r:
        h = hashlib.sha1(f.read() + e.read() + r.read()).hexdigest()
    if not (os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == h):
        cxx = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"
        cmd = [cxx, "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-w",
               "-include", os.path.join(HERE, "cuda_emu.h"), "-I", d, "-x", "c++", cu, "-o", so]
        subprocess.run(cmd, check=True)
        with open(stamp, "w") as f:
            f.write(h)
    return desc, so
