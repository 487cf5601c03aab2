"""OM -> Plan: storages, subkernels and margins.

Mirrors Language/Paraiso/Generator/OMTrans.hs:39-198 and Generator/Plan.hs:33-92.
The Plan records the reference's own subkernel cut (one subkernel per OMWriteGroup).  The
B200 backend uses the margins and the storage naming from here and re-cuts the kernels
itself (generator/b200/schedule.py); the oracle emitter (oracle/plantrans.py) follows the
reference cut literally.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

from .. import annotation as A
from ..om.graph import DynValue, OM
from ..optimization import optimize
from .native import Setup


@dataclass
class StorageRef:  # Plan.hs:70-92
    static_idx: Optional[int]            # StaticRef i
    manifest: Optional[Tuple[int, int]]  # ManifestRef kernel node
    dyn: DynValue
    static_name: Optional[str] = None

    @property
    def name(self) -> str:
        if self.static_idx is not None:
            return f"om_s{self.static_idx}_{self.static_name}"
        return f"om_m{self.manifest[0]}_{self.manifest[1]}"


@dataclass
class SubKernelRef:  # Plan.hs:51-68
    kernel_idx: int
    om_write_group_idx: int
    input_idxs: List[int]
    calc_idxs: List[int]
    output_idxs: List[int]
    realm: str
    lower_boundary: Tuple[int, ...]
    upper_boundary: Tuple[int, ...]
    kernel_name: str = ""

    @property
    def name(self) -> str:
        return f"om_{self.kernel_name}_sub_{self.om_write_group_idx}"


@dataclass
class Plan:  # Plan.hs:33-42
    name: str
    om: OM
    setup: Setup
    storages: List[StorageRef]
    sub_kernels: List[SubKernelRef]
    lower_margin: Tuple[int, ...]
    upper_margin: Tuple[int, ...]

    @property
    def memory_size(self) -> Tuple[int, ...]:
        return tuple(n + l + u for n, l, u in zip(self.setup.local_size, self.lower_margin, self.upper_margin))


def _valid_to_lower(setup: Setup, valid: A.Valid) -> Tuple[int, ...]:  # OMTrans.hs:103-110
    out = []
    for ax, iv in enumerate(valid.intervals):
        if setup.boundary[ax] == A.CYCLIC:
            out.append(0)
        elif iv.lower == A.NEGA_INF:
            out.append(0)
        elif iv.lower is not None and iv.lower[0] == 1:
            out.append(iv.lower[1])
        else:
            raise ValueError("wrong lower Margin!")
    return tuple(out)


def _valid_to_upper(setup: Setup, valid: A.Valid) -> Tuple[int, ...]:  # OMTrans.hs:111-116
    out = []
    for ax, iv in enumerate(valid.intervals):
        if setup.boundary[ax] == A.CYCLIC:
            out.append(0)
        elif iv.upper == A.POSI_INF:
            out.append(0)
        elif iv.upper is not None and iv.upper[0] == 2:
            out.append(-iv.upper[1])
        else:
            raise ValueError("wrong upper Margin!")
    return tuple(out)


def stencil_radius(om: OM) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
    """Margins the OM needs when every axis is treated as Open.  The B200 backend sizes its
    ghost zones with this on Cyclic axes too (the wrap is materialised as ghost cells)."""
    om = optimize("O3", om)
    valids = [v for k in om.kernels for nd in k.dataflow.nodes
              if nd.inst is not None and nd.inst.op == "Store" for v in A.to_list(A.Valid, nd.anot)]
    united = valids[0]
    for v in valids[1:]:
        united = united.intersection(v)
    fake = Setup(local_size=tuple(1 for _ in range(om.dim)))
    return _valid_to_lower(fake, united), _valid_to_upper(fake, united)


def translate(setup: Setup, om0: OM) -> Plan:  # OMTrans.hs:39-51
    om = optimize(setup.opt_level, om0)
    statics = [StorageRef(i, None, sv.namee, sv.name) for i, sv in enumerate(om.setup.static_values)]
    manifest_nodes = []  # (kernel idx, node idx, node)
    for kidx, k in enumerate(om.kernels):
        for idx, nd in enumerate(k.dataflow.nodes):
            if A.to_maybe(A.Allocation, nd.anot) == A.Manifest:
                if not nd.is_value:
                    raise ValueError("a non-Value node is marked as Manifest")
                manifest_nodes.append((kidx, idx, nd))
    manifests = [StorageRef(None, (k, i), nd.value) for (k, i, nd) in manifest_nodes]

    store_valids = [v for k in om.kernels for nd in k.dataflow.nodes
                    if nd.inst is not None and nd.inst.op == "Store" for v in A.to_list(A.Valid, nd.anot)]
    united = store_valids[0]
    for v in store_valids[1:]:
        united = united.intersection(v)

    groups = []
    for (k, i, nd) in manifest_nodes:
        g = A.to_maybe(A.OMWriteGroup, nd.anot)
        if g is None:
            raise ValueError(f"OMWriteGroup missing : {(k, i)}")
        groups.append(g.gid)
    n_sub = 1 + max([-1] + groups)

    subs: List[SubKernelRef] = []
    for gid in range(n_sub):
        mine = [(k, i, nd) for (k, i, nd), g in zip(manifest_nodes, groups) if g == gid]
        kidx = mine[0][0]
        inputs = sorted({x for (_, _, nd) in mine for d in A.to_list(A.Direct, nd.anot)[:1] for x in d.nodes})
        calcs = sorted({x for (_, _, nd) in mine for c in A.to_list(A.Calc, nd.anot)[:1] for x in c.nodes})
        realms = {nd.value.realm for (_, _, nd) in mine}
        valids = {A.to_maybe(A.Valid, nd.anot) for (_, _, nd) in mine}
        if len(realms) != 1 or len(valids) != 1:
            raise ValueError("elements are different. mismatch.")
        valid = next(iter(valids))
        subs.append(SubKernelRef(
            kernel_idx=kidx, om_write_group_idx=gid, input_idxs=inputs, calc_idxs=calcs,
            output_idxs=[i for (_, i, _) in mine], realm=next(iter(realms)),
            lower_boundary=_valid_to_lower(setup, valid), upper_boundary=_valid_to_upper(setup, valid),
            kernel_name=om.kernels[kidx].name))
    return Plan(name=om.name, om=om, setup=setup, storages=statics + manifests, sub_kernels=subs,
                lower_margin=_valid_to_lower(setup, united), upper_margin=_valid_to_upper(setup, united))
