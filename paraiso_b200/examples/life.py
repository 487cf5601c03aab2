"""Conway's Game of Life as an Orthotope Machine program.

`master`   : transcription of /root/reference examples/Life/Generator.hs:39-115
             (80x48 by default, Cyclic, statics cell/population/generation, kernels init/proceed).
`exampled` : transcription of examples-old/Life-exampled/LifeMain.hs:33-125, the program whose
             generated C++ is checked in under examples-old/Life-exampled/dist (128x128, Open,
             R-pentomino written by `init`); used to pin the oracle against the reference's own output.
"""
from __future__ import annotations

from ..annotation import CYCLIC, OPEN
from ..generator.native import Setup
from ..om.builder import (StaticValue, bind, eq, foldl1, ge, le, load, loadIndex, loadSize,
                          makeOM, reduce, select, shift, store, sum_, imm)
from ..om.graph import ARRAY, SCALAR, Named, OM

# adjacency vectors (Generator.hs:79-82)
ADJ_VECS = list(zip([-1, 0, 1, -1, 1, -1, 0, 1],
                    [-1, -1, -1, 0, 0, 1, 1, 1]))
# R-pentomino (LifeMain.hs:46-49)
R5MINO = list(zip([1, 2, 0, 1, 1], [0, 0, 1, 1, 2]))


def life_om(variant: str = "master") -> OM:
    cell = Named("cell", StaticValue(ARRAY, "Int"))
    population = Named("population", StaticValue(SCALAR, "Int"))
    generation = Named("generation", StaticValue(SCALAR, "Int"))

    def proceed():  # Generator.hs:85-115 / LifeMain.hs:58-87
        old_cell = bind(load(cell))
        gen = bind(load(generation))
        neighbours = [bind(shift(v, old_cell)) for v in ADJ_VECS]
        if variant == "master":
            num = bind(sum_(neighbours))                      # NumericPrelude.sum: 0 + n1 + ...
        else:
            num = bind(foldl1(lambda a, b: a + b, neighbours))  # foldl1 (+)
        # (c==0 && n==3) || (c==1 && (n>=2 && n<=3))   -- && is infixr 3, || infixr 2
        is_alive = bind((eq(old_cell, 0) & eq(num, 3)) |
                        (eq(old_cell, 1) & (ge(num, 2) & le(num, 3))))
        new_cell = bind(select(is_alive, imm(1, ARRAY, "Int"), 0))
        store(population, reduce("Sum", new_cell))
        store(generation, gen + 1)
        store(cell, new_cell)

    if variant == "master":
        def init():  # Generator.hs:71-76
            store(cell, 0)
            store(population, 0)
            store(generation, 0)
        vars_ = [cell, population, generation]
    elif variant == "exampled":
        def init():  # LifeMain.hs:90-110
            coord = [bind(loadIndex(ax)) for ax in range(2)]
            size = [bind(loadSize(ax, realm=ARRAY)) for ax in range(2)]
            half = [bind(size[ax] // 2) for ax in range(2)]

            def agree(point):
                return foldl1(lambda a, b: a & b,
                              [eq(coord[i] - half[i], imm(point[i], ARRAY, "Int")) for i in range(2)])
            alive = bind(foldl1(lambda a, b: a | b, [agree(p) for p in R5MINO]))
            c = bind(select(alive, imm(1, ARRAY, "Int"), 0))
            store(cell, c)
            store(population, reduce("Sum", c))
            store(generation, imm(0, SCALAR, "Int"))
        vars_ = [population, generation, cell]
    else:
        raise ValueError(variant)
    return makeOM("Life", [], vars_, [("init", init), ("proceed", proceed)], dim=2)


def life_setup(variant: str = "master", size=None) -> Setup:
    if variant == "master":  # Generator.hs:39-44
        s = Setup(local_size=size or (80, 48), boundary=(CYCLIC, CYCLIC), directory="./dist/")
    else:
        s = Setup(local_size=size or (128, 128), boundary=(OPEN, OPEN), directory="./dist/")  # LifeMain.hs:121-125
    # rows in flight per CTA and chunk height, re-swept whenever the kernel changed: round 1 (profiles/r1_life_sweep.txt) 3 rows;
    # round 2 with balanced chunks (r2o_life_chunks.jsonl) 3 rows, 20-row chunks; after the lean ghost block took the kernel from
    # issue-bound to latency-bound (ALU pipe 75 % -> 67 %), deeper staging and shorter chunks pay (r2ag / r2ah_life_pf.jsonl, same box:
    # 3 rows / 20-row chunks 0.3512 ms, 6 / 20 0.3470, 4 / 16 0.3455, 6 / 16 0.3425, 6 / 14 0.3451, 6 / 12 0.3540): 1040 chunks of
    # 15.75 rows = 25 waves of CTAs, 6 x 2 KB rows in flight per CTA
    s.tuning.prefetch_rows = 6
    s.tuning.chunk_rows_light = 16
    s.tuning.min_blocks = 9          # 9 CTAs of 128 threads per SM: at most 56 registers (the steady-state copy of the row loop asks for 60)
    return s
