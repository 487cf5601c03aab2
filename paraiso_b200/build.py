"""Generate + compile a machine: OM -> files -> nvcc -> lib<Name>.so (in-tree, cached by content).

The reference leaves compilation to a user Makefile (examples/Hydro/Makefile:5-10:
`nvcc -O3 -arch=sm_20`).  Here the generated kernel file is compiled for sm_100a only.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
from typing import Optional, Tuple

from .generator.b200.emit import generate
from .generator.native import Setup
from .om.graph import OM

GEN_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_generated")
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the B200 backend has no CPU fallback")
    return p


def nvcc_flags(fmad: bool) -> list:
    # -fmad=false keeps the per-cell arithmetic bit-identical to the reference's C++ (no FMA
    # contraction); IEEE division / sqrt are nvcc's defaults.
    return NVCC_ARCH + ["-O3", "-lineinfo", "-std=c++17", f"-fmad={'true' if fmad else 'false'}",
                        "-prec-div=true", "-prec-sqrt=true", "-shared", "-Xcompiler", "-fPIC"]


def generate_to(setup: Setup, om: OM, tag: Optional[str] = None, vnt=None) -> Tuple[dict, str]:
    """Write the generated files under paraiso_b200/_generated/<tag>/ and return (abi, dir)."""
    files = generate(setup, om, vnt=vnt)
    tag = tag or f"{om.name}_{''.join(b[0] for b in setup.boundary)}"
    d = os.path.join(GEN_ROOT, tag)
    os.makedirs(d, exist_ok=True)
    for fn, text in files:
        path = os.path.join(d, fn)
        old = None
        if os.path.exists(path):
            with open(path) as f:
                old = f.read()
        if old != text:
            with open(path, "w") as f:
                f.write(text)
    desc = json.loads(dict(files)[f"{om.name}_abi.json"])
    return desc, d


def compile_kernels(d: str, name: str, fmad: bool = False, verbose: bool = False) -> str:
    """nvcc <Name>_kernels.cu -> libom_<Name>[_fma].so next to it (rebuilt when the source changed)."""
    cu = os.path.join(d, f"{name}_kernels.cu")
    so = os.path.join(d, f"libom_{name}{'_fma' if fmad else ''}.so")
    stamp = so + ".sha1"
    flags = nvcc_flags(fmad)
    with open(cu, "rb") as f, open(os.path.join(d, "om_runtime.cuh"), "rb") as r:
        h = hashlib.sha1(f.read() + r.read() + " ".join(flags).encode()).hexdigest()
    if os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == h:
        return so
    # several ranks of one job may get here at once (torchrun on a tree that was not prebuilt): build into a private file
    # and rename, so that nobody ever dlopens a half-written library
    tmp = f"{so}.tmp{os.getpid()}"
    cmd = [nvcc_path()] + flags + (["-Xptxas", "-v"] if verbose else []) + [cu, "-o", tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, so)
    if verbose:
        with open(os.path.join(d, f"ptxas_{name}{'_fma' if fmad else ''}.log"), "w") as f:
            f.write(res.stderr)
    with open(stamp + f".tmp{os.getpid()}", "w") as f:
        f.write(h)
    os.replace(stamp + f".tmp{os.getpid()}", stamp)
    return so


def build_machine(setup: Setup, om: OM, tag: Optional[str] = None, fmad: bool = False, vnt=None,
                  verbose: bool = False) -> Tuple[dict, str]:
    desc, d = generate_to(setup, om, tag=tag, vnt=vnt)
    so = compile_kernels(d, desc["name"], fmad=fmad, verbose=verbose)
    return desc, so
