"""Typed annotation bag carried by every OM node.

Mirrors Language/Paraiso/Annotation.hs:17-49 (`Annotation = [Dynamic]` with add / set /
weakSet / toMaybe / toList / map keyed by the Haskell type of the element) and the
annotation types under Language/Paraiso/Annotation/*.hs.  An element here is a small
frozen dataclass; its Python class plays the role of the Haskell TypeRep.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import FrozenSet, Optional, Tuple, Type


def empty() -> list:
    return []


def add(x, ys: list) -> list:  # Annotation.hs:27-28
    return [x] + list(ys)


def set_(x, ys: list) -> list:  # Annotation.hs:31-32
    return [x] + [y for y in ys if type(y) is not type(x)]


def weak_set(x, ys: list) -> list:  # Annotation.hs:35-38
    if any(type(y) is type(x) for y in ys):
        return list(ys)
    return [x] + list(ys)


def to_list(cls: Type, ys: list) -> list:  # Annotation.hs:43-44
    return [y for y in ys if type(y) is cls]


def to_maybe(cls: Type, ys: list):  # Annotation.hs:47-48
    for y in ys:
        if type(y) is cls:
            return y
    return None


def map_(cls: Type, f, ys: list) -> list:  # Annotation.hs:51-58
    return [f(y) if type(y) is cls else y for y in ys]


# ---- Annotation/Allocation.hs:14-21 ----------------------------------------------------------
@dataclass(frozen=True)
class Allocation:
    kind: str  # "Existing" | "Manifest" | "Delayed"


Existing = Allocation("Existing")
Manifest = Allocation("Manifest")
Delayed = Allocation("Delayed")


@dataclass(frozen=True)
class AllocationChoice:
    choices: Tuple[Allocation, ...]


# ---- Annotation/Execution.hs:13 ------------------------------------------------------------------
@dataclass(frozen=True)
class Alive:
    alive: bool


# ---- Annotation/Boundary.hs:21-38, Interval.hs:18-33 ---------------------------------------
# NearBoundary g, ordered NegaInfinity < LowerBoundary _ < UpperBoundary _ < PosiInfinity
# (derived Ord: constructor order first, then the payload).
NEGA_INF = (0, 0)
POSI_INF = (3, 0)


def lower_boundary(x: int):
    return (1, x)


def upper_boundary(x: int):
    return (2, x)


@dataclass(frozen=True)
class Interval:
    """Half-open [lower, upper) over NearBoundary; None,None encodes Empty."""
    lower: Optional[tuple]
    upper: Optional[tuple]

    @property
    def is_empty(self) -> bool:
        return self.lower is None or self.lower >= self.upper

    def intersection(self, o: "Interval") -> "Interval":
        if self.lower is None or o.lower is None:
            return EMPTY_INTERVAL
        r = Interval(max(self.lower, o.lower), min(self.upper, o.upper))
        return EMPTY_INTERVAL if r.is_empty else r


EMPTY_INTERVAL = Interval(None, None)


@dataclass(frozen=True)
class Valid:
    """Annotation/Boundary.hs:21 — one interval per axis."""
    intervals: Tuple[Interval, ...]

    def intersection(self, o: "Valid") -> "Valid":
        if len(self.intervals) != len(o.intervals):
            raise ValueError("length mismatch in merging two Valid")
        return Valid(tuple(a.intersection(b) for a, b in zip(self.intervals, o.intervals)))


OPEN = "Open"      # Annotation/Boundary.hs:35-38
CYCLIC = "Cyclic"


# ---- Annotation/Dependency.hs:23-51 ------------------------------------------------------
@dataclass(frozen=True)
class Direct:
    nodes: Tuple[int, ...]


@dataclass(frozen=True)
class Indirect:
    nodes: Tuple[int, ...]


@dataclass(frozen=True)
class Calc:
    nodes: FrozenSet[int]


@dataclass(frozen=True)
class KernelWriteGroup:
    gid: int


@dataclass(frozen=True)
class OMWriteGroup:
    gid: int


@dataclass(frozen=True)
class OptLevel:
    """Optimization.hs:55-61, stored as a global annotation (Optimization.hs:38-45)."""
    level: int  # Unoptimized=-1, O0..O3 = 0..3
