"""Top level of the B200 backend: OM + Setup -> generated files.

Counterpart of Language/Paraiso/Generator.hs:41-66 (`generate` returns [(FilePath, Text)]) for
`Native.language == B200`.  Emits
    <Name>_kernels.cu   sm_100a kernels + extern "C" stage launchers   (cuda.py)
    <Name>_abi.json     machine description consumed by the host side (statics, margins,
                        ghost widths, stage table) — what the host class needs to drive the C ABI
    <Name>.hpp/.cpp     plain C++ host class with the reference's public surface (host.py)
    om_runtime.cuh      shared device runtime (the analogue of `commonLibraries`, PlanTrans.hs:723)
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Tuple

from ... import annotation as A
from ...om.graph import SCALAR, OM
from ..native import Setup
from ..plan import Plan, stencil_radius, translate as om_translate
from .cuda import StageEmitter, emit_scalar_stage, pick_vnt
from .schedule import KernelSchedule, schedule_kernel

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "csrc")


def describe(om: OM, plan: Plan, schedules: List[KernelSchedule], emitters) -> dict:
    setup = plan.setup
    dim = setup.dim
    rad_lo, rad_hi = stencil_radius(om)
    pad2 = lambda t: list(t) + [0] * (2 - len(t))
    statics = [dict(name=sv.name, realm=sv.namee.realm, type=sv.namee.type) for sv in om.setup.static_values]
    nslots = len(statics) + sum(len(ks.reduce_slots) + ks.extra_slots for ks in schedules)
    kernels = []
    # static scalars that some kernel loads: a reduce result stored into any other scalar is only ever read by the
    # host, so with several ranks its all_reduce can wait until the host asks for the value
    loaded_scalars = set()
    for ks in schedules:
        for op in ks.ops.values():
            if op.kind == "Load" and op.realm == SCALAR:
                loaded_scalars.add(op.inst.arg)
    for ks in schedules:
        stages = []
        consumed = set()       # reduce results that feed arithmetic or array stages
        for op in ks.ops.values():
            if op.kind != "Reduce":
                for a in op.args:
                    if ks.ops[a].kind == "Reduce":
                        consumed.add(a)
        direct_store = {}      # reduce value id -> static it is stored into unchanged
        for (sidx, v) in ks.scalar_stores:
            if ks.ops[v].kind == "Reduce":
                direct_store.setdefault(v, []).append(sidx)
        slot_to_vid = {slot: v for v, slot in ks.reduce_slots.items()}
        for si, st in enumerate(ks.stages):
            em = emitters[(ks.name, si)]
            stages.append(dict(
                symbol=em.name, level=st.level, V=em.V, NT=em.NT, smem=em.smem_bytes(), w_out=em.W_OUT,
                inputs=sorted({i.static_idx for i in st.inputs.values()}),
                outputs=list(dict.fromkeys(s for (s, _v) in st.store_targets)), zplanes=st.zplanes,
                reduces=[dict(op=rop, slot=slot, type=ks.ops[v].ctype,
                              # deferred: only stored, unchanged, into scalars that no kernel loads
                              deferred=(slot_to_vid[slot] not in consumed and slot_to_vid[slot] in direct_store and
                                        all(sx not in loaded_scalars for sx in direct_store[slot_to_vid[slot]])),
                              stored_to=direct_store.get(slot_to_vid[slot], []))
                         for (v, rop, slot) in {t[2]: t for t in reversed(st.reduce_targets)}.values()][::-1] +
                        # carried: the next call's level-0 reduce, produced by this stage (consumed on the device)
                        [dict(op=rop, slot=slot, type=ks.ops[v].ctype, deferred=False, stored_to=[], carried=True)
                         for (v, rop, slot) in st.carried],
                rings=len(em.depth), phases=len(st.phases), warmup=st.warmup,
                # boundary-first chunk order + in-kernel "boundary rows written" signal (several ranks; rank-1 / rank-2 machines)
                bfirst=bool(getattr(em, "bfirst", False)),
                mat_candidates=[dict(c, kernel=ks.name) for c in st.mat_candidates],
                chunk_rows=(0 if (len(st.phases) > 1 or em.smem_bytes() > 48 * 1024) else setup.tuning.chunk_rows_light)))
            assert not stages[-1]["bfirst"] or stages[-1]["chunk_rows"] > 0
        kernels.append(dict(
            name=ks.name, stages=stages,
            scalars=(f"om_{om.name}_{ks.name}_scalars" if ks.scalar_stores else None),
            array_stores=list(dict.fromkeys(s for (s, _v) in ks.array_stores)),
            scalar_stores=[s for (s, _v) in ks.scalar_stores],
            loaded_arrays=ks.loaded_arrays,
            # carry: {skip_stage, pairs [(reduce slot, carry slot)], arrays, scalars}: when the previous call on this
            # machine was this kernel and nothing else wrote those statics since, stage `skip_stage` is replaced by
            # copying each carry slot into its reduce slot
            carry=ks.carry))
    return dict(
        name=om.name, dim=dim, local_size=(list(setup.local_size) + [1] * (2 - dim)),
        boundary=list(setup.boundary) + [A.OPEN] * (2 - dim),
        lower_margin=pad2(plan.lower_margin), upper_margin=pad2(plan.upper_margin),
        radius_lo=pad2(rad_lo), radius_hi=pad2(rad_hi),
        statics=statics, nslots=nslots, kernels=kernels)


def generate(setup: Setup, om0: OM, vnt: Dict[Tuple[str, int], Tuple[int, int]] = None) -> List[Tuple[str, str]]:
    if setup.dim > 3:
        raise NotImplementedError("the B200 backend emits rank-1, rank-2 and rank-3 machines")
    plan = om_translate(setup, om0)
    om = plan.om
    nstat = len(om.setup.static_values)
    schedules: List[KernelSchedule] = []
    slot = nstat
    for k in om.kernels:
        ks = schedule_kernel(om, k, slot, setup.tuning.mat_threshold, setup.tuning.mat_flip, setup.tuning.carry_reduces, setup.tuning.planes_per_cta,
                             setup.tuning.sink_selects, bool(setup.fast_math) and setup.tuning.fast_algebra, setup.tuning.pull_shifts)
        slot += len(ks.reduce_slots) + ks.extra_slots
        schedules.append(ks)
    cu: List[str] = [
        f"// GENERATED by paraiso_b200 (language = B200) for OM `{om.name}` — do not edit.",
        f"// boundary = {setup.boundary}, lowerMargin = {plan.lower_margin}, upperMargin = {plan.upper_margin}",
        '#include "om_runtime.cuh"',
        "",
    ]
    emitters = {}
    for ks in schedules:
        for si, st in enumerate(ks.stages):
            V, NT = (vnt or {}).get((ks.name, si), pick_vnt(om, st, ks, setup.tuning))
            from .warpstream import WarpStreamEmitter, eligible
            em = (WarpStreamEmitter if eligible(st, V, setup.tuning) else StageEmitter)(om, plan, ks, st, si, V, NT)
            cu.append(em.kernel())
            cu.append(em.launcher())
            cu.append("")
            emitters[(ks.name, si)] = em
        sc = emit_scalar_stage(om, ks)
        if sc:
            cu.append(sc)
            cu.append("")
    desc = describe(om, plan, schedules, emitters)
    cu.append("// one thread on the host's communication stream: returns once the boundary CTAs of a g.bfirst launch have stored their rows")
    cu.append(f'extern "C" int om_{om.name}_wait_boundary(void* scratch, void* stream) {{')
    cu.append("  OM_LAUNCH(om_wait_boundary_kernel, dim3(1, 1), 32, 0, (cudaStream_t)stream, (unsigned*)scratch);")
    cu.append("  OM_CUDA_CHECK_LAUNCH();")
    cu.append("  return 0;")
    cu.append("}")
    cu.append(f'extern "C" int om_{om.name}_abi_version(void) {{ return 3; }}')
    with open(os.path.join(CSRC, "om_runtime.cuh")) as f:
        runtime = f.read()
    files = [(f"{om.name}_kernels.cu", "\n".join(cu) + "\n"),
             (f"{om.name}_abi.json", json.dumps(desc, indent=1) + "\n"),
             ("om_runtime.cuh", runtime)]
    from .host import emit_host
    files += emit_host(desc)
    return files


def generateIO(setup: Setup, om: OM) -> List[Tuple[str, str]]:  # Generator.hs:29-36
    os.makedirs(setup.directory, exist_ok=True)
    out = []
    for fn, text in generate(setup, om):
        path = os.path.join(setup.directory, fn)
        with open(path, "w") as f:
            f.write(text)
        out.append((path, text))
    return out


def describe_only(setup: Setup, om: OM, kernel: str = None) -> List[dict]:
    """The materialisation genes ({kernel, vid, op, cost, default, chosen}) of a machine, for tuning.local_search."""
    import json
    files = dict(generate(setup, om))
    desc = json.loads(files[f"{om.name}_abi.json"])
    return [c for k in desc["kernels"] if kernel in (None, k["name"]) for st in k["stages"] for c in st["mat_candidates"]]
