"""examples/InitialCondition/Generator.hs:31-60 — a heart-shaped initial condition on a 500x500 grid: exercises
`cast`, integer powers (`^`, repeated squaring, OM/Builder/Internal.hs:335-347), `**` (NumericPrelude's default
`x ** y = exp (log x * y)`), division of immediates and `atan`."""
from ..generator.native import Setup
from ..om.builder import StaticValue, atan, bind, broadcast, cast, exp, imm, loadIndex, loadSize, log, makeOM, store
from ..om.graph import ARRAY, Named, OM


def starstar(x, y):
    """Algebra.Transcendental default: x ** y = exp (log x * y)."""
    return exp(log(x) * y)


def initialcondition_om() -> OM:
    table = Named("table", StaticValue(ARRAY, "Double"))

    def create():  # Generator.hs:54-60
        x01 = bind(cast(loadIndex(0), "Double") / cast(broadcast(loadSize(0)), "Double"))
        y01 = bind(cast(loadIndex(1), "Double") / cast(broadcast(loadSize(1)), "Double"))
        x = bind(4 * (x01 - 0.5))
        y = bind(5 * (y01 - 0.5))
        third = imm(1, ARRAY, "Double") / 3
        z = bind(atan((1 - x ** 2 - (y - starstar(x ** 2, third)) ** 2) * 10))
        store(table, z)
    return makeOM("TableMaker", [], [table], [("create", create)], dim=2)


def initialcondition_setup(size=(500, 500)) -> Setup:  # Generator.hs:31-34
    return Setup(local_size=tuple(size), directory="./dist/")
