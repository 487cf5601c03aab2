"""2-D compressible Euler solver (MUSCL + HLLC, two-stage) as an Orthotope Machine program.

Transcription of /root/reference examples/Hydro/Hydro.hs:53-192 (the `Hydrable` class, `Hydro`
record, bindPrimitive/bindConserved) and examples/Hydro/HydroMain.hs:44-302 (statics, init,
boundaryCondition, buildProceed, proceedSingle, addFlux, interpolate, hllc).

variant="master"   : Real from `real` (Double upstream), loadIndex cast to Real, loadSize
                     cast+broadcast, HLLC selector outputs annotated Manifest, Abs operator.
variant="exampled" : the older revision whose generated C++ is checked in under
                     examples-old/Hydro-exampled/dist (Real = Float; loadIndex/loadSize typed Real
                     directly, loadSize in the Array realm; no Manifest on the HLLC selector;
                     abs x = max x (negate x), visible at dist/Hydro.cpp:390-391).
Every Haskell `Builder` expression that is not `bind`-ed is re-run at each use, exactly as
in the reference; `B` thunks reproduce that (om/builder.py).
"""
from __future__ import annotations

from fractions import Fraction
from typing import List

from .. import annotation as A
from ..annotation import OPEN
from ..generator.native import Setup
from ..om.builder import (B, StaticValue, annotate, bind, broadcast, cast, contract, foldl1, ge, gt, imm,
                          le, load, loadIndex, loadSize, lt, makeOM, max_, min_, mkOp1, pi, reduce,
                          select, shift, sin, sqrt, store, unit_vector, abs_)
from ..om.graph import ARRAY, SCALAR, Named, OM

DIM = 2
LEAVES = (["density"] + [f"velocity{i}" for i in range(DIM)] + ["pressure"] +
          [f"momentum{i}" for i in range(DIM)] + ["energy", "enthalpy"] +
          [f"densityFlux{i}" for i in range(DIM)] +
          [f"momentumFlux{i}{j}" for i in range(DIM) for j in range(DIM)] +
          [f"energyFlux{i}" for i in range(DIM)] +
          ["soundSpeed", "kineticEnergy", "internalEnergy"])


class _Ctx:
    real = "Double"
    variant = "master"


def R(x) -> B:
    """A literal of type BR (Builder (Value TArray Real))."""
    return imm(x, ARRAY, _Ctx.real)


def kGamma() -> B:  # Hydro.hs:53-54
    return R(Fraction(5, 3))


def _abs(x: B) -> B:
    if _Ctx.variant == "exampled":
        return max_(x, mkOp1("Neg", x))
    return abs_(x)


# ---------------------------------------------------------------------------------------------
# class Hydrable (Hydro.hs:99-127)
# ---------------------------------------------------------------------------------------------
class Hydrable:
    def density(self) -> B: raise NotImplementedError

    def velocity(self) -> List[B]:
        return [self.momentum()[i] / self.density() for i in range(DIM)]

    def pressure(self) -> B:
        return (kGamma() - 1) * self.internalEnergy()

    def momentum(self) -> List[B]:
        return [self.density() * self.velocity()[i] for i in range(DIM)]

    def energy(self) -> B:
        return self.kineticEnergy() + 1 / (kGamma() - 1) * self.pressure()

    def enthalpy(self) -> B:
        return self.energy() + self.pressure()

    def densityFlux(self) -> List[B]:
        return self.momentum()

    def momentumFlux(self) -> List[List[B]]:
        return [[self.momentum()[i] * self.velocity()[j] + self.pressure() * R(1 if i == j else 0)
                 for j in range(DIM)] for i in range(DIM)]

    def energyFlux(self) -> List[B]:
        return [self.enthalpy() * self.velocity()[i] for i in range(DIM)]

    def soundSpeed(self) -> B:
        return sqrt(kGamma() * self.pressure() / self.density())  # soundSpeed', Hydro.hs:57-58

    def kineticEnergy(self) -> B:
        return 0.5 * contract(DIM, lambda i: self.velocity()[i] * self.momentum()[i])

    def internalEnergy(self) -> B:
        return self.energy() - self.kineticEnergy()


class PrimitiveVar(Hydrable):  # Hydro.hs:179-183
    def __init__(self, d, v, p): self.d, self.v, self.p = d, v, p
    def density(self): return self.d
    def velocity(self): return self.v
    def pressure(self): return self.p


class ConservedVar(Hydrable):  # Hydro.hs:185-189
    def __init__(self, d, m, e): self.d, self.m, self.e = d, m, e
    def density(self): return self.d
    def momentum(self): return self.m
    def energy(self): return self.e


class Hydro(Hydrable):
    """data Hydro a (Hydro.hs:130-136) with Functor/Applicative/Traversable over LEAVES order."""

    def __init__(self, leaves: List):
        assert len(leaves) == len(LEAVES)
        self.leaves = list(leaves)

    def _g(self, name): return self.leaves[LEAVES.index(name)]
    def density(self): return self._g("density")
    def velocity(self): return [self._g(f"velocity{i}") for i in range(DIM)]
    def pressure(self): return self._g("pressure")
    def momentum(self): return [self._g(f"momentum{i}") for i in range(DIM)]
    def energy(self): return self._g("energy")
    def enthalpy(self): return self._g("enthalpy")
    def densityFlux(self): return [self._g(f"densityFlux{i}") for i in range(DIM)]
    def momentumFlux(self): return [[self._g(f"momentumFlux{i}{j}") for j in range(DIM)] for i in range(DIM)]
    def energyFlux(self): return [self._g(f"energyFlux{i}") for i in range(DIM)]
    def soundSpeed(self): return self._g("soundSpeed")
    def kineticEnergy(self): return self._g("kineticEnergy")
    def internalEnergy(self): return self._g("internalEnergy")

    def fmap(self, f): return Hydro([f(x) for x in self.leaves])

    def mapM(self, f):  # traverse in declaration order
        return Hydro([f(x) for x in self.leaves])


def liftA(f, *hs: Hydro) -> Hydro:  # f <$> h1 <*> h2 ...
    return Hydro([f(*xs) for xs in zip(*[h.leaves for h in hs])])


def bindHydro(x: Hydrable) -> Hydro:  # Hydro.hs:70-96
    density = bind(x.density())
    velocity = [bind(v) for v in x.velocity()]
    pressure = bind(x.pressure())
    momentum = [bind(m) for m in x.momentum()]
    energy = bind(x.energy())
    enthalpy = bind(x.enthalpy())
    dflux = [bind(f) for f in x.densityFlux()]
    mflux = [[bind(f) for f in row] for row in x.momentumFlux()]
    eflux = [bind(f) for f in x.energyFlux()]
    cs = bind(x.soundSpeed())
    kin = bind(x.kineticEnergy())
    internal = bind(x.internalEnergy())
    return Hydro([density] + velocity + [pressure] + momentum + [energy, enthalpy] + dflux +
                 [f for row in mflux for f in row] + eflux + [cs, kin, internal])


def bindPrimitive(d, v, p) -> Hydro: return bindHydro(PrimitiveVar(d, v, p))
def bindConserved(d, m, e) -> Hydro: return bindHydro(ConservedVar(d, m, e))


# ---------------------------------------------------------------------------------------------
# HydroMain.hs
# ---------------------------------------------------------------------------------------------
def _names():
    real = _Ctx.real
    sreal = lambda n: Named(n, StaticValue(SCALAR, real))
    areal = lambda n: Named(n, StaticValue(ARRAY, real))
    return dict(
        generation=Named("generation", StaticValue(SCALAR, "Int")),
        time=sreal("time"), cfl=sreal("cfl"),
        dR=[sreal(f"dR{i}") for i in range(DIM)],
        extent=[sreal(f"extent{i}") for i in range(DIM)],
        density=areal("density"),
        velocity=[areal(f"velocity{i}") for i in range(DIM)],
        pressure=areal("pressure"))


def hydro_vars():  # HydroMain.hs:44-53
    n = _names()
    return [n["generation"], n["time"], n["cfl"]] + n["dR"] + n["extent"] + [n["density"]] + n["velocity"] + [n["pressure"]]


def _icoord():
    if _Ctx.variant == "exampled":   # loadIndex (0::Real) axis
        return [bind(loadIndex(ax, gauge=_Ctx.real)) for ax in range(DIM)]
    return [bind(cast(loadIndex(ax), _Ctx.real)) for ax in range(DIM)]  # HydroMain.hs:85,111


def _region(coord, extent):  # HydroMain.hs:95,121
    ex, ey = 0, 1
    return bind(gt(coord[ey], 0.47 * extent[ey]) & (lt(coord[ey], 0.53 * extent[ey]) & lt(coord[ex], 0)))


def buildInit():  # HydroMain.hs:79-102
    n = _names()
    dRG = [bind(load(x)) for x in n["dR"]]
    extentG = [bind(load(x)) for x in n["extent"]]
    dR = [bind(broadcast(x)) for x in dRG]
    extent = [bind(broadcast(x)) for x in extentG]
    icoord = _icoord()
    coord = [bind(dR[i] * icoord[i]) for i in range(DIM)]
    ex = 0
    vplus = [R(6), R(0)]
    vminus = [R(0), R(0)]
    region = _region(coord, extent)
    velo = [bind(select(region, vplus[i], vminus[i])) for i in range(DIM)]
    factor = bind(1 + 1e-3 * sin(6 * pi(ARRAY, _Ctx.real) * coord[ex]))
    store(n["density"], factor * kGamma() * kGamma() * select(region, R(1), 10))
    for i in range(DIM):
        store(n["velocity"][i], velo[i])
    store(n["pressure"], factor * 0.6)


def boundaryCondition(cell: Hydro) -> Hydro:  # HydroMain.hs:105-131
    n = _names()
    dRG = [bind(load(x)) for x in n["dR"]]
    extentG = [bind(load(x)) for x in n["extent"]]
    dR = [bind(broadcast(x)) for x in dRG]
    extent = [bind(broadcast(x)) for x in extentG]
    icoord = _icoord()
    if _Ctx.variant == "exampled":   # loadSize TLocal (0::Real) axis
        isize = [bind(loadSize(ax, gauge=_Ctx.real, realm=ARRAY)) for ax in range(DIM)]
    else:
        isize = [bind(broadcast(cast(loadSize(ax), _Ctx.real))) for ax in range(DIM)]
    coord = [bind(dR[i] * icoord[i]) for i in range(DIM)]
    vplus = [R(6), R(0)]
    vminus = [R(0), R(0)]
    region = _region(coord, extent)
    cell0 = bindPrimitive(kGamma() * kGamma() * select(region, R(1), 10),
                          [select(region, vplus[i], vminus[i]) for i in range(DIM)],
                          R(0.6))
    outOf = bind(foldl1(lambda a, b: a | b, [lt(icoord[i], 0) for i in range(DIM)]) |
                 foldl1(lambda a, b: a | b, [ge(icoord[i], isize[i]) for i in range(DIM)]))
    manifest = lambda an: A.add(A.Manifest, an)
    return liftA(lambda a, b: annotate(manifest, select(outOf, a, b)), cell0, cell)


def interpolateSingle(order: int, x0: B, x1: B, x2: B, x3: B):  # HydroMain.hs:218-235
    if order == 1:
        return (x1, x2)
    if order == 2:
        d01 = bind(x1 - x0)
        d12 = bind(x2 - x1)
        d23 = bind(x3 - x2)

        def absmaller(a, b):
            return select(le(a * b, 0), R(0), select(lt(_abs(a), _abs(b)), a, b))
        d1 = bind(absmaller(d01, d12))
        d2 = bind(absmaller(d12, d23))
        l = bind(x1 + d1 / 2)
        r = bind(x2 - d2 / 2)
        return (l, r)
    raise ValueError(f"{order}th order spatial interpolation is not yet implemented")


def interpolate(order: int, i: int, cell: Hydro):  # HydroMain.hs:199-216
    def shifti(n):
        vec = tuple(n if i == j else 0 for j in range(DIM))
        return lambda b: shift(vec, b)
    a0 = cell.mapM(lambda b: bind(shifti(2)(b)))
    a1 = cell.mapM(lambda b: bind(shifti(1)(b)))
    a2 = cell.mapM(lambda b: bind(shifti(0)(b)))
    a3 = cell.mapM(lambda b: bind(shifti(-1)(b)))
    intp = [interpolateSingle(order, w, x, y, z) for w, x, y, z in zip(a0.leaves, a1.leaves, a2.leaves, a3.leaves)]
    l = Hydro([p[0] for p in intp])
    r = Hydro([p[1] for p in intp])

    def bp(x: Hydro) -> Hydro:
        dens1 = bind(x.density())
        velo1 = [bind(v) for v in x.velocity()]
        pres1 = bind(x.pressure())
        return bindPrimitive(dens1, velo1, pres1)
    lp = bp(l)
    rp = bp(r)
    return lp, rp


def hllc(i: int, left: Hydro, right: Hydro) -> Hydro:  # HydroMain.hs:237-276
    def hllcQ(sp, p):
        return select(le(p, sp), R(1), sqrt(1 + (kGamma() + 1) / (2 * kGamma()) * (sp / p - 1)))

    def starState(starShock, shock, x: Hydro) -> Hydro:
        speed = x.velocity()[i]
        dens = bind(x.density() * (shock - speed) / (shock - starShock))
        mome = [bind(dens * (starShock if i == j else x.velocity()[j])) for j in range(DIM)]
        enrg = bind(dens * (x.energy() / x.density() +
                            (starShock - speed) * (starShock + x.pressure() / x.density() / (shock - speed))))
        return bindConserved(dens, mome, enrg)

    densMid = bind((left.density() + right.density()) / 2)
    soundMid = bind((left.soundSpeed() + right.soundSpeed()) / 2)
    speedLeft = left.velocity()[i]
    speedRight = right.velocity()[i]
    presStar = bind(max_(R(0), (left.pressure() + right.pressure()) / 2 -
                         densMid * soundMid * (speedRight - speedLeft)))
    shockLeft = bind(left.velocity()[i] - left.soundSpeed() * hllcQ(presStar, left.pressure()))
    shockRight = bind(right.velocity()[i] + right.soundSpeed() * hllcQ(presStar, right.pressure()))
    shockStar = bind((right.pressure() - left.pressure()
                      + left.density() * speedLeft * (shockLeft - speedLeft)
                      - right.density() * speedRight * (shockRight - speedRight))
                     / (left.density() * (shockLeft - speedLeft) -
                        right.density() * (shockRight - speedRight)))
    lesta = starState(shockStar, shockLeft, left)
    rista = starState(shockStar, shockRight, right)

    def selector(a, b, c, d):
        s = select(lt(R(0), shockLeft), a, select(lt(R(0), shockStar), b, select(lt(R(0), shockRight), c, d)))
        if _Ctx.variant in ("master", "periodic"):  # HydroMain.hs:258
            return annotate(lambda an: A.add(A.Manifest, an), s)
        return s
    return liftA(selector, left, lesta, rista, right).mapM(bind)


def addFlux(dt: B, dR: List[B], wall: List[Hydro], ex: int, cell: Hydro) -> Hydro:  # HydroMain.hs:185-196
    dtdx = bind(dt / dR[ex])
    leftWall = wall[ex].mapM(bind)
    neg_unit = tuple(-u for u in unit_vector(DIM, ex))
    rightWall = wall[ex].mapM(lambda b: bind(shift(neg_unit, b)))
    dens1 = bind(cell.density() + dtdx * (leftWall.densityFlux()[ex] - rightWall.densityFlux()[ex]))
    mome1 = [bind(cell.momentum()[j] + dtdx * (leftWall.momentumFlux()[j][ex] - rightWall.momentumFlux()[j][ex]))
             for j in range(DIM)]
    enrg1 = bind(cell.energy() + dtdx * (leftWall.energyFlux()[ex] - rightWall.energyFlux()[ex]))
    return bindConserved(dens1, mome1, enrg1)


def proceedSingle(order: int, dt: B, dR: List[B], cellF: Hydro, cellS: Hydro) -> Hydro:  # HydroMain.hs:169-180
    wall = []
    for i in range(DIM):
        lp, rp = interpolate(order, i, cellF)
        wall.append(hllc(i, lp, rp))
    # foldl1 (.) [f0, f1] $ return cellS  ==  f0 (f1 (return cellS)): axis 1 is added first
    cellN = cellS
    for i in reversed(range(DIM)):
        cellN = addFlux(dt, dR, wall, i, cellN)
    return bindPrimitive(max_(R(1e-2), cellN.density()), cellN.velocity(), max_(R(1e-2), cellN.pressure()))


def buildProceed():  # HydroMain.hs:134-164
    n = _names()
    dens = bind(load(n["density"]))
    velo = [bind(load(v)) for v in n["velocity"]]
    pres = bind(load(n["pressure"]))
    timeG = bind(load(n["time"]))
    cflG = bind(load(n["cfl"]))
    dRG = [bind(load(x)) for x in n["dR"]]
    dR = [bind(broadcast(x)) for x in dRG]
    cell0 = bindPrimitive(dens, velo, pres)
    # variant "periodic": the same solver without the jet-inflow boundary condition, for Cyclic setups
    # (analytic tests: advected entropy wave; SURVEY §8 f4, attic/GA.reproduce/massive-test.cu:58-118)
    cell = cell0 if _Ctx.variant == "periodic" else boundaryCondition(cell0)
    timescale = lambda i: dR[i] / (cell.soundSpeed() + _abs(cell.velocity()[i]))
    dts = bind(foldl1(min_, [timescale(i) for i in range(DIM)]))
    dtG = bind(cflG * reduce("Min", dts))
    dt = bind(broadcast(dtG))
    cell2 = proceedSingle(1, dt / 2, dR, cell, cell)
    cell3 = proceedSingle(2, dt, dR, cell2, cell)
    store(n["time"], timeG + dtG)
    store(n["density"], cell3.density())
    for i in range(DIM):
        store(n["velocity"][i], cell3.velocity()[i])
    store(n["pressure"], cell3.pressure())


def hydro_om(variant: str = "master", real: str = None) -> OM:
    _Ctx.variant = variant
    _Ctx.real = real or ("Float" if variant == "exampled" else "Double")
    return makeOM("Hydro", [], hydro_vars(), [("init", buildInit), ("proceed", buildProceed)], dim=DIM)


def hydro_setup(size=(1024, 1024), periodic: bool = False, fast: bool = False) -> Setup:  # HydroMain.hs:290-294
    """`fast` = Setup.fast_math.  The schedule knobs are the winners of the sweeps on the B200 (profiles/r1_hydro_sweep.txt,
    profiles/r2e_sweep_exact.jsonl): both builds run three 128-thread CTAs per SM."""
    from ..annotation import CYCLIC
    s = Setup(local_size=tuple(size), boundary=(CYCLIC, CYCLIC) if periodic else (OPEN, OPEN), directory="./dist/")
    s.fast_math = fast
    # 128-thread CTAs, three per SM (up to 168 registers per thread), register prefetch of the unstaged inputs, and dt of the
    # next step reduced in this step's epilogue.  fast build: 12.15 Gcell/s vs 11.24 with two 256-thread CTAs (round 1);
    # bit-exact build with the branch-free IEEE-correct division (Tuning.exact_divsqrt = "newton"): the same shape wins the
    # 48-candidate sweep of profiles/r2e_sweep_exact.jsonl (10.0 Gcell/s; two 256-thread CTAs without prefetch: 9.6)
    s.tuning.threads_heavy = 128
    s.tuning.min_blocks_heavy = 3
    s.tuning.carry_reduces = True
    # separate pipeline-fill loop (profiles/r2t_variants_peel.jsonl, same box): bit-exact build 2.219 -> 2.199 ms, fast build 1.158 -> 1.171 ms
    s.tuning.peel_fill = not fast
    return s
