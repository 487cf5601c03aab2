"""OM interchange: dump (reference text format) -> parse -> identical graph, identical generated CUDA."""
import pytest

from paraiso_b200.examples.hydro import hydro_om, hydro_setup
from paraiso_b200.examples.life import life_om, life_setup
from paraiso_b200.examples.shiftexample import shiftexample_om, shiftexample_setup
from paraiso_b200.generator.b200.emit import generate
from paraiso_b200.om.interchange import parse_om
from paraiso_b200.om.prettyprint import pretty_print_a1
from paraiso_b200.optimization import optimize

CASES = {"life": (lambda: life_om("master"), lambda: life_setup("master")),
         "hydro": (lambda: hydro_om("master"), lambda: hydro_setup((1024, 1024))),
         "shift": (shiftexample_om, lambda: shiftexample_setup(True))}


@pytest.mark.parametrize("name", list(CASES))
def test_roundtrip_unoptimised_and_optimised(name):
    mk, mksetup = CASES[name]
    raw = pretty_print_a1(mk())                       # dump before any analysis
    om2 = parse_om(raw)
    assert pretty_print_a1(om2) == raw
    opt = pretty_print_a1(optimize("O3", mk()))       # dump after analysis; re-analysing the import reproduces it
    assert pretty_print_a1(optimize("O3", parse_om(opt))) == opt
    assert pretty_print_a1(optimize("O3", parse_om(raw))) == opt


@pytest.mark.parametrize("name", ["life", "hydro"])
def test_imported_graph_generates_identical_kernels(name):
    mk, mksetup = CASES[name]
    a = dict(generate(mksetup(), mk()))
    b = dict(generate(mksetup(), parse_om(pretty_print_a1(mk()))))
    key = [k for k in a if k.endswith("_kernels.cu")][0]
    assert a[key] == b[key]


def test_legacy_dump_is_rejected():
    with pytest.raises(ValueError):
        parse_om(pretty_print_a1(life_om("exampled"), legacy=True))


def test_legacy_dump_without_immediates_is_rejected_with_partial_table():
    with pytest.raises(ValueError, match="no entry"):
        parse_om(pretty_print_a1(life_om("exampled"), legacy=True), immediates={"init": {}, "proceed": {}})


# --- the reference's checked-in dumps (examples-old/*/output/OM.txt, old `Imm <<Int>>` format) as direct inputs -------
import json
import os

from paraiso_b200.om.interchange import recover_immediates

_GOLD = os.path.join(os.path.dirname(__file__), "golden", "legacy_immediates.json")
_REF = "/root/reference/examples-old"
_LEGACY = {"life": (lambda: life_om("exampled"), lambda: life_setup("exampled"), False,
                    "Life-exampled/output/OM.txt", "Life-exampled/dist/Life.cpp"),
           "hydro": (lambda: hydro_om("exampled"), lambda: hydro_setup((1024, 1024)), True,
                     "Hydro-exampled/output/OM.txt", "Hydro-exampled/dist/Hydro.cpp")}


def _golden_table(name):
    with open(_GOLD) as f:
        return {k: {int(i): lit for i, lit in v.items()} for k, v in json.load(f)[name].items()}


@pytest.mark.parametrize("name", list(_LEGACY))
def test_legacy_dump_plus_generated_cpp_rebuilds_the_program(name):
    """Our legacy-format dump is byte-identical to the reference's (tests/test_om_dump.py pins its SHA-256), the
    immediate table comes from the reference's generated C++ (committed fixture): the import re-dumps to the same
    text, carries the same immediate values as the builder's program and generates the same CUDA."""
    mk, mksetup, alive, _dump, _cpp = _LEGACY[name]
    dump = pretty_print_a1(optimize("O3", mk()), legacy=True, alive=alive)
    om = parse_om(dump, immediates=_golden_table(name))
    assert pretty_print_a1(optimize("O3", om), legacy=True, alive=alive) == dump
    assert pretty_print_a1(optimize("O3", om)) == pretty_print_a1(optimize("O3", mk()))     # values, not just shapes
    a, b = dict(generate(mksetup(), mk())), dict(generate(mksetup(), om))
    key = [k for k in a if k.endswith("_kernels.cu")][0]
    assert a[key] == b[key]


@pytest.mark.parametrize("name", list(_LEGACY))
def test_immediates_recovered_from_oracle_text_equal_those_of_the_reference_text(name):
    """The oracle's reference-style C++ (oracle/plantrans.py, master naming om_<kernel>_sub_<g>) prints the same
    literal values at the same node ids as the reference's own generated file (old naming <Name>_sub_<g>)."""
    import oracle.plantrans as P
    mk, mksetup, _alive, _dump, _cpp = _LEGACY[name]
    import numpy as np

    def value(lit):        # the oracle prints a float immediate with all its digits, the reference with Haskell's shortest
        return float(np.float32(lit[:-1])) if lit.endswith("f") else lit       # `show`: the same float either way

    ours, ref = recover_immediates(P.generate(mksetup(), mk())), _golden_table(name)
    assert {k: {i: value(x) for i, x in v.items()} for k, v in ours.items()} == \
        {k: {i: value(x) for i, x in v.items()} for k, v in ref.items()}


@pytest.mark.skipif(not os.path.isdir(_REF), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("name", list(_LEGACY))
def test_reference_files_are_direct_inputs(name):
    mk, mksetup, alive, dump_path, cpp_path = _LEGACY[name]
    with open(os.path.join(_REF, dump_path)) as f:
        dump = f.read()
    with open(os.path.join(_REF, cpp_path)) as f:
        table = recover_immediates(f.read())
    assert table == _golden_table(name)
    om = parse_om(dump, immediates=table)
    assert pretty_print_a1(optimize("O3", om), legacy=True, alive=alive) == dump
    a, b = dict(generate(mksetup(), mk())), dict(generate(mksetup(), om))
    for k in a:
        if k.endswith((".cu", ".cpp", ".hpp")):
            assert a[k] == b[k], k


def test_command_line_front_end_from_dump_and_cpp(tmp_path):
    """tools/om2b200.py: dump (old format) + the C++ of the same generator run -> the tracked LifeExampled_OO sources
    (the ones the GPU parity tests run against the reference's compiled golden).  The two input files are written by
    our mirror printer and by the oracle's reference-style emitter, so this runs without /root/reference."""
    import subprocess
    import sys

    import oracle.plantrans as P
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    (tmp_path / "OM.txt").write_text(pretty_print_a1(optimize("O3", life_om("exampled")), legacy=True, alive=False))
    (tmp_path / "Life.cpp").write_text(P.generate(life_setup("exampled"), life_om("exampled")))
    out = tmp_path / "dist-b200"
    cmd = [sys.executable, os.path.join(root, "tools", "om2b200.py"), str(tmp_path / "OM.txt"), "--cpp",
           str(tmp_path / "Life.cpp"), "--size", "128x128", "--boundary", "open,open", "--out", str(out),
           "--tune", "prefetch_rows=6", "--tune", "chunk_rows_light=16", "--tune", "min_blocks=9"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    tracked = os.path.join(root, "paraiso_b200", "_generated", "LifeExampled_OO")
    for fn in ("Life.hpp", "Life.cpp", "Life_kernels.cu", "Life_abi.h"):
        with open(os.path.join(tracked, fn)) as f:
            assert (out / fn).read_text() == f.read(), fn
    bad = subprocess.run(cmd[:3] + ["--size", "128x128", "--out", str(out)], capture_output=True, text=True)
    assert bad.returncode != 0 and "immediates" in bad.stderr
