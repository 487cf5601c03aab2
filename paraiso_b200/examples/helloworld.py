"""examples/HelloWorld/Generator.hs:31-63 — multiplication table on a 10x20 grid and its sum (45*190 = 8550)."""
from ..generator.native import Setup
from ..om.builder import StaticValue, bind, loadIndex, makeOM, reduce, store
from ..om.graph import ARRAY, SCALAR, Named, OM


def helloworld_om() -> OM:
    table = Named("table", StaticValue(ARRAY, "Int"))
    total = Named("total", StaticValue(SCALAR, "Int"))

    def create():  # Generator.hs:58-63
        x = bind(loadIndex(0))
        y = bind(loadIndex(1))
        z = bind(x * y)
        store(table, z)
        store(total, reduce("Sum", z))
    return makeOM("TableMaker", [], [table, total], [("create", create)], dim=2)


def helloworld_setup() -> Setup:  # Generator.hs:31-34
    return Setup(local_size=(10, 20), directory="./dist/")
