"""Orthotope Machine front-end: graph IR (graph.py) and Builder EDSL (builder.py)."""
from .graph import ARRAY, SCALAR, DynValue, Graph, Inst, Kernel, Named, Node, OM, Setup  # noqa: F401
