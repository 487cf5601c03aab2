"""ORACLE (test infrastructure, not product): build + drive the CPU checkers.

  * `OracleMachine`  — compiles the C++ emitted by oracle/plantrans.py (the restatement of the
                       reference generator) for one (program, setup) and drives it via ctypes.
  * `RefLife/RefHydro` — drive oracle/_ref/*.so, i.e. the reference's OWN checked-in generated
                       C++ (examples-old/*-exampled/dist), built by oracle/Makefile.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REFDIR = os.path.join(HERE, "_ref")
NP_TYPE = {"Int": np.int32, "Float": np.float32, "Double": np.float64, "Bool": np.bool_, "Integer": np.int64}


def _cxx() -> str:
    return "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"


def compile_cpp(src: str, tag: str, openmp: bool = False, opt: str = "-O2") -> str:
    """Compile a C++ source string into oracle/_build/<tag>-<hash>.so (cached)."""
    os.makedirs(BUILD, exist_ok=True)
    flags = [opt, "-fPIC", "-shared", "-w", "-ffp-contract=off"] + (["-fopenmp"] if openmp else [])
    h = hashlib.sha1((src + " ".join(flags)).encode()).hexdigest()[:16]
    so = os.path.join(BUILD, f"{tag}-{h}.so")
    if not os.path.exists(so):
        cpp = os.path.join(BUILD, f"{tag}-{h}.cpp")
        with open(cpp, "w") as f:
            f.write(src)
        tmp = so + f".tmp{os.getpid()}"
        subprocess.run([_cxx()] + flags + [cpp, "-o", tmp], check=True)
        os.replace(tmp, so)
    return so


class OracleMachine:
    """The emitted reference-style class, driven from Python."""

    def __init__(self, setup, om, openmp: bool = False, opt: str = "-O2"):
        from . import plantrans
        from paraiso_b200.generator.plan import translate
        self.plan = translate(setup, om)
        src = plantrans.emit(self.plan)
        tag = f"{self.plan.name}_{'x'.join(map(str, setup.local_size))}_{''.join(b[0] for b in setup.boundary)}"
        self.so = compile_cpp(src, tag, openmp=openmp, opt=opt)
        self.lib = ctypes.CDLL(self.so)
        self.lib.om_new.restype = ctypes.c_void_p
        self.lib.om_static_ptr.restype = ctypes.c_void_p
        self.lib.om_static_ptr.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self.lib.om_call.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        self.lib.om_delete.argtypes = [ctypes.c_void_p]
        self.h = ctypes.c_void_p(self.lib.om_new())
        self.memory_size = self.plan.memory_size
        self.lower_margin = self.plan.lower_margin
        self.upper_margin = self.plan.upper_margin
        self.local_size = setup.local_size
        self.statics = {sv.name: (i, sv.namee) for i, sv in enumerate(self.plan.om.setup.static_values)}

    def __del__(self):
        try:
            self.lib.om_delete(self.h)
        except Exception:
            pass

    def call(self, kernel: str):
        if self.lib.om_call(self.h, kernel.encode()) != 0:
            raise KeyError(kernel)

    def array(self, name: str) -> np.ndarray:
        """View of a static Array in memory layout [.., i1, i0] (axis 0 fastest), margins included."""
        idx, dv = self.statics[name]
        n = int(np.prod(self.memory_size))
        ptr = self.lib.om_static_ptr(self.h, idx)
        buf = (ctypes.c_char * (n * np.dtype(NP_TYPE[dv.type]).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=NP_TYPE[dv.type]).reshape(tuple(reversed(self.memory_size)))

    def interior(self, name: str) -> np.ndarray:
        a = self.array(name)
        sl = tuple(slice(l, l + n) for l, n in zip(reversed(self.lower_margin), reversed(self.local_size)))
        return a[sl]

    def scalar(self, name: str) -> np.ndarray:
        idx, dv = self.statics[name]
        ptr = self.lib.om_static_ptr(self.h, idx)
        buf = (ctypes.c_char * np.dtype(NP_TYPE[dv.type]).itemsize).from_address(ptr)
        return np.frombuffer(buf, dtype=NP_TYPE[dv.type])


def _ref(name: str) -> ctypes.CDLL:
    path = os.path.join(REFDIR, name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle` where /root/reference is mounted")
    lib = ctypes.CDLL(path)
    lib.ref_new.restype = ctypes.c_void_p
    for f in ("ref_delete", "ref_init", "ref_proceed"):
        getattr(lib, f).argtypes = [ctypes.c_void_p]
    for f in ("ref_memory_size", "ref_size", "ref_lower_margin"):
        getattr(lib, f).argtypes = [ctypes.c_void_p, ctypes.c_int]
    return lib


class RefLife:
    """The reference's generated Life (examples-old/Life-exampled/dist/Life.cpp)."""

    def __init__(self):
        self.lib = _ref("libref_life.so")
        self.lib.ref_cell.restype = ctypes.POINTER(ctypes.c_int)
        self.lib.ref_cell.argtypes = [ctypes.c_void_p]
        self.lib.ref_population.argtypes = [ctypes.c_void_p]
        self.lib.ref_generation.argtypes = [ctypes.c_void_p]
        self.h = ctypes.c_void_p(self.lib.ref_new())
        self.m0 = self.lib.ref_memory_size(self.h, 0)
        self.m1 = self.lib.ref_memory_size(self.h, 1)

    def init(self): self.lib.ref_init(self.h)
    def proceed(self): self.lib.ref_proceed(self.h)
    def cell(self): return np.ctypeslib.as_array(self.lib.ref_cell(self.h), shape=(self.m1, self.m0))
    def population(self): return self.lib.ref_population(self.h)
    def generation(self): return self.lib.ref_generation(self.h)


class RefHydro:
    """The reference's generated Hydro (examples-old/Hydro-exampled/dist/Hydro.cpp), float 1024^2."""
    SCALARS = {"time": 1, "cfl": 2, "dR0": 3, "dR1": 4, "extent0": 5, "extent1": 6}
    ARRAYS = {"density": 7, "velocity0": 8, "velocity1": 9, "pressure": 10}

    def __init__(self, openmp: bool = False):
        self.lib = _ref("libref_hydro_omp.so" if openmp else "libref_hydro.so")
        self.lib.ref_scalar.restype = ctypes.POINTER(ctypes.c_float)
        self.lib.ref_scalar.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self.lib.ref_array.restype = ctypes.POINTER(ctypes.c_float)
        self.lib.ref_array.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self.h = ctypes.c_void_p(self.lib.ref_new())
        self.m0 = self.lib.ref_memory_size(self.h, 0)
        self.m1 = self.lib.ref_memory_size(self.h, 1)
        self.n0 = self.lib.ref_size(self.h, 0)
        self.n1 = self.lib.ref_size(self.h, 1)
        self.margin = self.lib.ref_lower_margin(self.h, 0)

    def init(self): self.lib.ref_init(self.h)
    def proceed(self): self.lib.ref_proceed(self.h)
    def scalar(self, name): return np.ctypeslib.as_array(self.lib.ref_scalar(self.h, self.SCALARS[name]), shape=(1,))
    def array(self, name): return np.ctypeslib.as_array(self.lib.ref_array(self.h, self.ARRAYS[name]), shape=(self.m1, self.m0))

    def setup_kh(self):
        """Parameter block of examples-old/Hydro-exampled/main-kh.cpp:39-45."""
        self.scalar("time")[0] = 0
        self.scalar("cfl")[0] = 0.5
        self.scalar("extent0")[0] = 1.0
        self.scalar("extent1")[0] = 1.0
        self.scalar("dR0")[0] = self.scalar("extent0")[0] / self.n0
        self.scalar("dR1")[0] = self.scalar("extent1")[0] / self.n1
