"""Rank-3 programs.  The reference's IR is rank-generic (Native.hs:21-22, Orthotope.hs:19-22) but its
examples are rank 1 and 2; these two exercise what a rank-3 program can contain: a 26-neighbour integer stencil with a
Sum reduce, and a floating-point program with loadIndex / loadSize of all axes, an intermediate worth a shared-memory
ring, an asymmetric axis-2 reach and a Max reduce that feeds a second stage."""
from __future__ import annotations

from ..om.builder import (StaticValue, bind, broadcast, cast, eq, ge, imm, le, load, loadIndex, loadSize, makeOM, reduce,
                          select, shift, sqrt, store, sum_)
from ..om.graph import ARRAY, SCALAR, Named


def life3d_om():
    """26-neighbour life on a rank-3 grid (rule 5..7 survive / 6 born), population reduce, generation counter."""
    cell = Named("cell", StaticValue(ARRAY, "Int"))
    pop = Named("population", StaticValue(SCALAR, "Int"))
    gen = Named("generation", StaticValue(SCALAR, "Int"))

    def proceed():
        c = bind(load(cell))
        nb = [shift((dx, dy, dz), c) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)]
        num = bind(sum_(nb))
        alive = bind((eq(c, 0) & eq(num, 6)) | (eq(c, 1) & ge(num, 5) & le(num, 7)))
        new = bind(select(alive, imm(1, ARRAY, "Int"), 0))
        store(cell, new)
        store(pop, reduce("Sum", new))
        store(gen, load(gen) + 1)
    return makeOM("Life3", [], [cell, pop, gen], [("proceed", proceed)], dim=3)


def diffusion3d_om():
    """7-point diffusion with a position-dependent source (loadIndex of all three axes, loadSize), an intermediate that
    is worth a shared-memory ring, an asymmetric axis-2 reach and a Max reduce feeding a second stage."""
    u = Named("u", StaticValue(ARRAY, "Double"))
    peak = Named("peak", StaticValue(SCALAR, "Double"))

    def init():
        x, y, z = (cast(loadIndex(a), "Double") for a in range(3))
        n2 = broadcast(cast(loadSize(2), "Double"))
        store(u, (x * 0.25 + y * y * 0.125 - z) / (n2 + 1.0))

    def proceed():
        x = bind(load(u))
        g = bind(sqrt(x * x + 2.0) / (3.0 + x * x))                      # materialised along axes 0 / 1, recomputed across planes
        lap = bind(shift((1, 0, 0), g) + shift((-1, 0, 0), g) + shift((0, 1, 0), g) + shift((0, -1, 0), g) +
                   shift((0, 0, 1), g) + shift((0, 0, -2), g) - 6 * g)
        new = bind(x + 0.05 * lap + 1e-3 * cast(loadIndex(2), "Double"))
        mx = bind(reduce("Max", new))
        store(peak, mx)
        store(u, new / (broadcast(mx) + 1.0))
    return makeOM("Diff3", [], [u, peak], [("init", init), ("proceed", proceed)], dim=3)


def heat3d_om(real: str = "Float"):
    """7-point explicit heat equation u += k * laplacian(u): the lightest rank-3 stencil (bandwidth test)."""
    u = Named("u", StaticValue(ARRAY, real))

    def proceed():
        x = bind(load(u))
        lap = bind(shift((1, 0, 0), x) + shift((-1, 0, 0), x) + shift((0, 1, 0), x) + shift((0, -1, 0), x) +
                   shift((0, 0, 1), x) + shift((0, 0, -1), x) - 6 * x)
        store(u, x + 0.1 * lap)
    return makeOM("Heat3", [], [u], [("proceed", proceed)], dim=3)
