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
