"""the reference's example drivers compile UNCHANGED against the generated host class
(examples/HelloWorld/main.cpp, examples/HelloGPU/main.cu, examples/ShiftExample/main.cpp for dist-open and dist-cyclic,
examples/Life/main.cpp), and the committed stdout goldens — produced by the same drivers on the oracle's reference-style
class — hold the hand-derivable answers.  The GPU side is tests/test_gpu_reference_drivers.py."""
import os
import re

import numpy as np
import pytest

from tests import refdrivers

needs_reference = pytest.mark.skipif(not os.path.isdir(refdrivers.REF), reason="/root/reference not mounted")


@needs_reference
@pytest.mark.parametrize("key", refdrivers.KEYS)
def test_driver_compiles_and_links_unchanged(key):
    exe = refdrivers.link_b200(key)
    assert os.access(exe, os.X_OK)


@needs_reference
@pytest.mark.parametrize("key", refdrivers.KEYS)
def test_goldens_are_what_the_oracle_class_prints(key, tmp_path):
    exe = refdrivers.link_oracle(key, str(tmp_path))
    with open(refdrivers.golden_path(key)) as f:
        assert refdrivers.run(key, exe) == f.read()


@needs_reference
@pytest.mark.parametrize("key", refdrivers.KEYS)
def test_driver_on_the_generated_class_with_emulated_kernels(key, tmp_path):
    """The reference's driver + the generated host class + the generated kernels on host threads: stdout equals the
    reference-style class's, byte for byte (mirrors, dirty tracking, margin coordinates, scalar accessors)."""
    exe = refdrivers.link_emulated(key, str(tmp_path))
    with open(refdrivers.golden_path(key)) as f:
        assert refdrivers.run(key, exe) == f.read()


def test_helloworld_golden_known_answer():
    """examples/HelloWorld/Generator.hs:58-63: table(x, y) = x*y on 10x20, total = 45*190."""
    with open(refdrivers.golden_path("helloworld")) as f:
        lines = f.read().rstrip("\n").split("\n")
    assert lines[-1] == "total: 8550" and len(lines) == 21
    for y, line in enumerate(lines[:20]):
        assert line == "".join(f"{x * y:4d}" for x in range(10))


@pytest.mark.parametrize("cyclic", [False, True])
def test_shiftexample_golden_known_answer(cyclic):
    """examples/ShiftExample: `shift v x` at cell i reads x[i - v]; Open keeps one margin cell per side (printed, index -1
    and 8), Cyclic wraps.  After init / increment the table is i+1; calculate = 10000*left + 100*centre + right."""
    with open(refdrivers.golden_path("shift_cyclic" if cyclic else "shift_open")) as f:
        blocks = [b.split("\n") for b in f.read().split("\n\n") if b.strip()]
    idx = list(range(8)) if cyclic else list(range(-1, 9))
    t = np.arange(1, 9)
    if cyclic:
        calc = 10000 * np.roll(t, 1) + 100 * t + np.roll(t, -1)
        tables = [np.arange(8), t, calc]
    else:
        tt = np.arange(0, 10)
        calc = np.concatenate([[0], 10000 * tt[:-2] + 100 * tt[1:-1] + tt[2:], [0]])    # margins are not valid: left 0
        tables = [np.arange(-1, 9), tt, calc]
    for blk, want in zip(blocks[:3], tables):
        assert [int(v) for v in blk[0].split()[1:]] == idx
        assert [int(v) for v in blk[1].split()[1:]] == list(want)
    inner = calc if cyclic else calc[1:-1]
    assert blocks[3][0] == f"total: {int(inner.sum())}"


def test_life_golden_matches_an_independent_life():
    """examples/Life/main.cpp on the Gosper-gun seed: every printed frame (two rows per text line, ` .':` glyphs) and
    population equals a five-line periodic numpy Life."""
    with open(refdrivers.golden_path("life")) as f:
        lines = f.read().split("\n")
    W, H, per = 80, 48, 26
    c = np.zeros((H, W), np.int64)
    for x, y in refdrivers.LIFE_PATTERN:
        c[y, x] = 1
    pop = 0                      # init stores population 0; the first frame is printed before any proceed()
    glyph = " .':"
    for t in range(refdrivers.LIFE_FRAMES):
        frame = lines[t * per:(t + 1) * per]
        for k in range(H // 2):
            assert frame[k] == "".join(glyph[2 * c[2 * k, x] + c[2 * k + 1, x]] for x in range(W)), (t, k)
        assert frame[24] == ""
        m = re.fullmatch(r"generation:\s+(\d+)  population:\s+(\d+)", frame[25])
        assert m and (int(m.group(1)), int(m.group(2))) == (t, pop), (t, frame[25])
        n = sum(np.roll(np.roll(c, dy, 0), dx, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0))
        c = (((c == 0) & (n == 3)) | ((c == 1) & (n >= 2) & (n <= 3))).astype(np.int64)
        pop = int(c.sum())


@needs_reference
def test_hydro_golden_is_what_the_oracle_class_prints(tmp_path):
    """examples/Hydro/main-kh.cpp on the reference-style class: dt of the first steps (7.3982e-05: the value the reference's
    own float sample reaches, tests/golden/hydro_exampled.json: time 0.00073982 after ten steps) and the snapshot digest."""
    import json
    with open(os.path.join(refdrivers.GOLDEN, "driver_hydro.json")) as f:
        want = json.load(f)
    assert want["times"][:2] == ["0", "7.3982e-05"]
    got = refdrivers.run_hydro(refdrivers.link_oracle_hydro(str(tmp_path)))
    assert got["times"] == want["times"] and got["column_sums"] == want["column_sums"] and got["diagonal"] == want["diagonal"]


@needs_reference
def test_initialcondition_driver_compiles_and_prints_the_oracle_output(tmp_path):
    """examples/InitialCondition/main.cpp (cast, ^, **, atan on a 500 x 500 grid, written to heart.txt): compiles unchanged
    against the generated class; with emulated kernels the file is byte-identical to the reference-style class's (both
    sides use the host's libm), whose digest is the committed golden."""
    import json
    assert os.access(refdrivers.link_heart("b200"), os.X_OK)
    want_text = refdrivers.run_heart(refdrivers.link_heart("oracle", str(tmp_path)))
    with open(os.path.join(refdrivers.GOLDEN, "driver_initialcondition.json")) as f:
        want = json.load(f)
    got = refdrivers.heart_digest(want_text)
    assert got["cells"] == want["cells"] == 250000 and got["column_sums"] == want["column_sums"] and got["samples"] == want["samples"]
    assert refdrivers.run_heart(refdrivers.link_heart("emulated", str(tmp_path))) == want_text
