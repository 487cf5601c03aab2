"""Snapshot I/O of the Hydro example: the text format examples/Hydro/main-kh.cpp:16-31 writes
(`x y density velocity0 velocity1 pressure` per cell, a blank line after each row, cell centres at dR * (i + 0.5)) and the
density map examples/Hydro/plot.rb draws from it with gnuplot (`splot u 1:2:3`, pm3d map, cbrange [0:100]).  gnuplot is
not a dependency here: the map is written as a binary PPM."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

FIELDS = ("density", "velocity0", "velocity1", "pressure")


def dump(path: str, machine, anti_alias: int = 1) -> None:
    """Write a snapshot of a Hydro `Machine` (or anything with get(name) / scalar(name)) in main-kh.cpp's format."""
    arrays = [np.asarray(machine.get(n), dtype=np.float64) for n in FIELDS]
    h, w = arrays[0].shape
    dr0, dr1 = float(machine.scalar("dR0")), float(machine.scalar("dR1"))
    with open(path, "w") as f:
        for iy in range(anti_alias // 2, h, anti_alias):
            for ix in range(anti_alias // 2, w, anti_alias):
                vals = [dr0 * (ix + 0.5), dr1 * (iy + 0.5)] + [a[iy, ix] for a in arrays]
                f.write(" ".join(f"{v:.6g}" for v in vals) + "\n")        # std::ostream's default: six significant digits
            f.write("\n")


def load(path: str) -> Tuple[np.ndarray, np.ndarray, Dict[str, np.ndarray]]:
    """Read a snapshot back: (x of the columns, y of the rows, {field: array[y, x]})."""
    with open(path) as f:
        a = np.fromstring(f.read(), sep=" ").reshape(-1, 6)
    ys = np.unique(a[:, 1])
    w = len(a) // len(ys)
    a = a.reshape(len(ys), w, 6)
    return a[0, :, 0].copy(), a[:, 0, 1].copy(), {n: a[:, :, 2 + k].copy() for k, n in enumerate(FIELDS)}


def colour(t: np.ndarray) -> np.ndarray:
    """gnuplot's default pm3d palette (rgbformulae 7,5,15): r = sqrt t, g = t^3, b = sin 2 pi t, clipped."""
    t = np.clip(t, 0.0, 1.0)
    rgb = np.stack([np.sqrt(t), t ** 3, np.clip(np.sin(2 * np.pi * t), 0.0, 1.0)], axis=-1)
    return (rgb * 255.0 + 0.5).astype(np.uint8)


def plot(path: str, out: str, field: str = "density", cbrange=(0.0, 100.0)) -> None:
    """plot.rb's picture: the field as a colour map (row 0 at the bottom, colour range `cbrange`), as a binary PPM."""
    _x, _y, fields = load(path)
    a = fields[field][::-1]
    img = colour((a - cbrange[0]) / (cbrange[1] - cbrange[0]))
    with open(out, "wb") as f:
        f.write(f"P6 {img.shape[1]} {img.shape[0]} 255\n".encode())
        f.write(img.tobytes())


if __name__ == "__main__":       # python -m paraiso_b200.snapshot output1/snapshot*.txt   (what `plot.rb output1/*.txt` does)
    import sys
    for fn in sys.argv[1:]:
        plot(fn, fn[:-4] + ".ppm")
        print(fn, "->", fn[:-4] + ".ppm", file=sys.stderr)
