"""examples/ShiftExample/Generator.hs:36-69 — 1-D shift under Open vs Cyclic; pins the shift sign
convention: `shift v x` at cell i reads x[i - v] (PlanTrans.hs:561-563)."""
from ..annotation import CYCLIC, OPEN
from ..generator.native import Setup
from ..om.builder import StaticValue, bind, load, loadIndex, makeOM, reduce, shift, store
from ..om.graph import ARRAY, SCALAR, Named, OM


def shiftexample_om() -> OM:
    table = Named("table", StaticValue(ARRAY, "Int"))
    total = Named("total", StaticValue(SCALAR, "Int"))

    def init():
        store(table, loadIndex(0))

    def increment():
        store(table, 1 + load(table))

    def calculate():
        center = bind(load(table))
        right = bind(shift((-1,), center))
        left = bind(shift((1,), center))
        ret = bind(10000 * left + 100 * center + right)
        store(table, ret)
        store(total, reduce("Sum", ret))
    return makeOM("TableMaker", [], [table, total],
                  [("init", init), ("increment", increment), ("calculate", calculate)], dim=1)


def shiftexample_setup(cyclic: bool = False) -> Setup:  # Generator.hs:36-41, 26-28
    return Setup(local_size=(8,), boundary=(CYCLIC if cyclic else OPEN,),
                 directory="./dist-cyclic/" if cyclic else "./dist-open/")
