"""Race detection for the generated kernels (SURVEY §5 lists none in the reference; its GA flips __syncthreads blindly,
Annotation/SyncThreads.hs, Tuning/Genetic.hs:279-282).  The emulated kernels run one host thread per CUDA thread with
__syncthreads as a std::barrier, so ThreadSanitizer checks what the schedule promises: every read of a shared-memory
ring is ordered after its write by a CTA barrier — two per row in Hydro's flux stage.  A negative
control removes one of them and must be reported."""
import os

import pytest

from tests.emu import hostclass

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


def _need(exe):
    if exe is None:
        pytest.skip("no ThreadSanitizer runtime in this toolchain")
    return exe


def test_life_kernels_are_race_free(tmp_path):
    from paraiso_b200.examples.life import life_om, life_setup
    exe = _need(hostclass.link_tsan(life_setup("master"), life_om("master"), "Life_hostclass",
                                    os.path.join(CPP, "life_driver.cpp"), str(tmp_path / "life_tsan")))
    out, reports = hostclass.run_tsan(exe, [6])
    assert reports == 0
    plain = str(tmp_path / "life_plain")
    hostclass.link_emulated(life_setup("master"), life_om("master"), "Life_hostclass", os.path.join(CPP, "life_driver.cpp"), plain)
    assert out == hostclass.run(plain, [6])


@pytest.mark.parametrize("fast", [False, True])
def test_hydro_kernels_are_race_free(fast, tmp_path):
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    setup = hydro_setup((64, 48), fast=fast)
    exe = _need(hostclass.link_tsan(setup, hydro_om("master"), f"Hydro_hostclass_{int(fast)}",
                                    os.path.join(CPP, "hydro_driver.cpp"), str(tmp_path / "hydro_tsan")))
    out, reports = hostclass.run_tsan(exe, [3])
    assert reports == 0 and len(out.split()) == 6


@pytest.mark.parametrize("which", [0, 1])      # the loop-top barrier and the one between the two phases of a row
def test_a_removed_barrier_is_reported(which, tmp_path):
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    setup = hydro_setup((64, 48), fast=True)
    exe = _need(hostclass.link_tsan(setup, hydro_om("master"), "Hydro_hostclass_1", os.path.join(CPP, "hydro_driver.cpp"),
                                    str(tmp_path / "hydro_bad"), drop_barrier=which, kernel_marker="om_Hydro_proceed_stage1_kernel("))
    _out, reports = hostclass.run_tsan(exe, [2])
    assert reports > 0


def test_rank3_kernels_are_race_free(tmp_path):
    from paraiso_b200.examples.rank3 import diffusion3d_om
    from paraiso_b200.generator.native import Setup
    setup = Setup(local_size=(20, 9, 8), boundary=("Cyclic", "Open", "Cyclic"))
    exe = _need(hostclass.link_tsan(setup, diffusion3d_om(), "Diff3_hostclass_COC", os.path.join(CPP, "diff3_driver.cpp"),
                                    str(tmp_path / "diff3_tsan")))
    out, reports = hostclass.run_tsan(exe, [2])
    assert reports == 0 and out.startswith("20 9 8 ")


# ---- memory safety: the same executables under AddressSanitizer ("device" arrays are malloc'ed by the runtime stand-in) ----------
def _asan_cases():
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    from paraiso_b200.examples.life import life_om, life_setup
    from paraiso_b200.examples.rank3 import diffusion3d_om
    from paraiso_b200.generator.native import Setup
    return {"life": (life_setup("master"), life_om("master"), "Life_hostclass", "life_driver.cpp", 5),
            "hydro_fast": (hydro_setup((64, 48), fast=True), hydro_om("master"), "Hydro_hostclass_1", "hydro_driver.cpp", 3),
            "diff3": (Setup(local_size=(20, 9, 8), boundary=("Open", "Open", "Open")), diffusion3d_om(), "Diff3_hostclass_OOO",
                      "diff3_driver.cpp", 2)}


@pytest.mark.parametrize("case", ["life", "hydro_fast", "diff3"])
def test_kernels_stay_inside_their_allocations(case, tmp_path):
    """No access outside an array's allocation: pipeline fill and halo reads stay within the OM_APRON_ROWS slack rows
    the ABI promises (include/paraiso_b200.h), on one device and on three slabs (thin slabs are the hard case)."""
    setup, om, tag, driver, steps = _asan_cases()[case]
    exe = _need(hostclass.link_tsan(setup, om, tag, os.path.join(CPP, driver), str(tmp_path / f"{case}_asan"), sanitizer="address"))
    for devices in (1, 3):
        out, reports = hostclass.run_tsan(exe, [steps], devices=devices)
        assert reports == 0 and out.strip(), devices


def test_missing_apron_rows_are_reported(tmp_path):
    """Negative control of the memory check: a host that allocates no slack rows (the ABI asks for OM_APRON_ROWS = 32)
    makes the Life kernel's pipeline fill read outside the allocation, and AddressSanitizer says so."""
    setup, om, tag, driver, steps = _asan_cases()["life"]
    exe = _need(hostclass.link_tsan(setup, om, tag, os.path.join(CPP, driver), str(tmp_path / "life_noapron"), sanitizer="address", apron=0))
    _out, reports = hostclass.run_tsan(exe, [2])
    assert reports > 0


# ---- synthetic programs (tests/programs.py) through the generated class with a generated driver, under both sanitizers ------------
@pytest.mark.parametrize("prog,sanitizer", [("multi_reduce", "thread"), ("wide_CO", "thread"), ("ring_float", "thread"),
                                            ("wide_OC", "address"), ("wide_OO", "address"), ("ring_float", "address"),
                                            ("chain_1d", "address")])
def test_synthetic_programs_under_sanitizers(prog, sanitizer, tmp_path):
    """Several reduces per stage and a reduce feeding a later stage, wide asymmetric stencils on mixed boundaries, a float ring
    read in both axes, a rank-1 chain: no race, no access outside an allocation, and the same printed results as the plain
    build — on one device and (rank 2) on two slabs."""
    from tests.emu.build_emu import build_emulated
    from tests.generic_driver import driver_source
    from tests.programs import PROGRAMS
    make_om, setup, kernels = PROGRAMS[prog]
    tag = f"prog_{prog}"
    desc, _so = build_emulated(setup, make_om(), tag=tag)
    drv = str(tmp_path / "driver.cpp")
    with open(drv, "w") as f:
        f.write(driver_source(desc, kernels))
    plain = str(tmp_path / "plain")
    hostclass.link_emulated(setup, make_om(), tag, drv, plain)
    want = hostclass.run(plain)
    assert want.strip()
    exe = _need(hostclass.link_tsan(setup, make_om(), tag, drv, str(tmp_path / f"san_{sanitizer}"), sanitizer=sanitizer))
    for devices in ((1, 2) if len(setup.local_size) > 1 else (1,)):
        out, reports = hostclass.run_tsan(exe, devices=devices)
        assert reports == 0, (devices, reports)
        if desc["kernels"][0]["stages"] and prog != "ring_float":        # (a float Sum over two slabs folds in another order)
            assert out == want, devices
