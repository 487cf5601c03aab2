"""Ready-made machines for the two north-star programs (generate -> nvcc -> Machine)."""
from __future__ import annotations

import numpy as np

from .build import build_machine
from .examples.hydro import hydro_om, hydro_setup
from .examples.life import life_om, life_setup
from .runtime import Machine


def build_life(fmad: bool = False, verbose: bool = False):
    """examples/Life/Generator.hs (Cyclic, Int): returns (abi, path of libom_Life.so)."""
    return build_machine(life_setup("master"), life_om("master"), tag="Life_CC", fmad=fmad, verbose=verbose)


def build_hydro(real: str = "Double", fmad: bool = False, verbose: bool = False, fast: bool = False):
    """examples/Hydro/HydroMain.hs (Open, Real = Double upstream).  `fast` = Setup.fast_math (implies FMA)."""
    setup = hydro_setup(fast=fast)
    # (per-node genes: tuning.local_search finds no flip that pays for either build at their present CTA shapes —
    #  profiles/r1_hydro_search_*.jsonl; with two 256-thread CTAs per SM the fast build gained 4 % from flipping value 362)
    return build_machine(setup, hydro_om("master", real=real), tag=f"Hydro_OO_{real}{'_fast' if fast else ''}",
                         fmad=fmad or fast, verbose=verbose)


def build_life_exampled(verbose: bool = False):
    """examples-old/Life-exampled/LifeMain.hs (Open, 128x128, R-pentomino init): the program whose generated C++
    the reference ships; used to compare the GPU result with the reference's own output."""
    return build_machine(life_setup("exampled"), life_om("exampled"), tag="LifeExampled_OO", verbose=verbose)


def build_hydro_exampled(verbose: bool = False):
    """examples-old/Hydro-exampled (Real = Float, Open, 1024x1024)."""
    return build_machine(hydro_setup(), hydro_om("exampled"), tag="HydroExampled_OO_Float", verbose=verbose)


def build_helloworld(verbose: bool = False):
    """examples/HelloWorld/Generator.hs (and examples/HelloGPU, the same program with language = CUDA): class TableMaker."""
    from .examples.helloworld import helloworld_om, helloworld_setup
    return build_machine(helloworld_setup(), helloworld_om(), tag="TableMaker_Hello", verbose=verbose)


def build_shiftexample(cyclic: bool, verbose: bool = False):
    """examples/ShiftExample/Generator.hs (rank 1; dist-open / dist-cyclic): class TableMaker."""
    from .examples.shiftexample import shiftexample_om, shiftexample_setup
    return build_machine(shiftexample_setup(cyclic), shiftexample_om(), tag="TableMaker_Shift" + ("C" if cyclic else "O"),
                         verbose=verbose)


def life_machine(size, device="cuda", **kw) -> Machine:
    desc, so = build_life()
    return Machine(desc, so, size=size, device=device, **kw)


def hydro_machine(size, device="cuda", real: str = "Double", fmad: bool = False, fast: bool = False, **kw) -> Machine:
    desc, so = build_hydro(real=real, fmad=fmad, fast=fast)
    return Machine(desc, so, size=size, device=device, **kw)


def hydro_set_params(m, size, cfl=0.5, extent=(1.0, 1.0)):
    """Parameter block of examples/Hydro/main-kh.cpp:38-43."""
    m.set_scalar("time", 0.0)
    m.set_scalar("cfl", cfl)
    m.set_scalar("extent0", extent[0])
    m.set_scalar("extent1", extent[1])
    m.set_scalar("dR0", extent[0] / size[0])
    m.set_scalar("dR1", extent[1] / size[1])


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def life_seed(nx: int, y0: int, nyl: int, nx_global: int = None, seed: int = 20261017, density: float = 0.35) -> np.ndarray:
    """Counter-based synthetic fill: cell (x, y) = 1 iff splitmix64(seed ^ (y*N + x)) < density * 2^64,
    so every rank (and the CPU oracle) fills its own slab identically."""
    n = nx_global or nx
    with np.errstate(over="ignore"):
        yy, xx = np.meshgrid(np.arange(y0, y0 + nyl, dtype=np.uint64), np.arange(nx, dtype=np.uint64), indexing="ij")
        h = splitmix64(np.uint64(seed) ^ (yy * np.uint64(n) + xx))
    return (h < np.uint64(int(density * 2.0 ** 64))).astype(np.int32)
