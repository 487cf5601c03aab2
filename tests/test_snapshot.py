"""paraiso_b200.snapshot: the snapshot text format of examples/Hydro/main-kh.cpp:16-31 and the density map of plot.rb."""
import numpy as np

from paraiso_b200 import snapshot


class FakeMachine:
    def __init__(self, w, h):
        rng = np.random.default_rng(3)
        self.a = {n: rng.uniform(0.5, 90.0, (h, w)) for n in snapshot.FIELDS}
        self.s = {"dR0": 1.0 / w, "dR1": 1.0 / h}
    def get(self, n): return self.a[n]
    def scalar(self, n): return self.s[n]


def test_dump_load_round_trip_and_format(tmp_path):
    m = FakeMachine(12, 7)
    p = str(tmp_path / "snapshot0000.txt")
    snapshot.dump(p, m)
    lines = open(p).read().split("\n")
    assert len(lines) == 7 * 13 + 1 and lines[12] == "" and len(lines[0].split()) == 6          # a blank line after each row
    x, y, f = snapshot.load(p)
    assert np.allclose(x, (np.arange(12) + 0.5) / 12, rtol=1e-5) and np.allclose(y, (np.arange(7) + 0.5) / 7, rtol=1e-5)
    for n in snapshot.FIELDS:
        assert np.allclose(f[n], m.a[n], rtol=1e-5)             # six significant digits


def test_matches_the_reference_driver_golden_head():
    """The first line main-kh.cpp writes for the KH initial condition (tests/golden/driver_hydro.json) parses to the
    same numbers through load()'s tokeniser."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "driver_hydro.json")))
    row = g["diagonal"][0]
    assert row[0] == row[1] == 0.000488281 and abs(row[2] - 27.7778) < 1e-9


def test_plot_writes_a_ppm(tmp_path):
    m = FakeMachine(16, 8)
    p = str(tmp_path / "s.txt")
    snapshot.dump(p, m)
    out = str(tmp_path / "s.ppm")
    snapshot.plot(p, out)
    raw = open(out, "rb").read()
    assert raw.startswith(b"P6 16 8 255\n") and len(raw) == len(b"P6 16 8 255\n") + 16 * 8 * 3
    assert (snapshot.colour(np.array([0.0, 1.0])) == np.array([[0, 0, 0], [255, 255, 0]], dtype=np.uint8)).all()
