"""Host <-> device plumbing around a Machine: NUMA placement of a rank and a double-buffered step pipeline.

The reference's drivers keep the state on the host between kernel calls only through the mirrored accessors
(examples/Hydro/main-kh.cpp:38-59 uploads once, steps, downloads per snapshot).  A caller that feeds every step
from host memory and wants the result grid back pays PCIe in both directions; `HostPipeline` overlaps the three
legs (upload of step k+1, kernels of step k, download of step k-1) on three streams with two staging slots each
way, so that the steady-state step costs max(upload, kernels, download) instead of their sum.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch


def gpu_numa_node(index: int) -> Optional[int]:
    """NUMA node the GPU hangs off (sysfs via the PCI bus id), or None when the platform does not say."""
    try:
        p = torch.cuda.get_device_properties(index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            n = int(f.read().strip())
        return n if n >= 0 else None
    except Exception:
        return None


def _cpulist(text: str) -> List[int]:
    out: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def pin_to_gpu_numa(index: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node (so that pinned host buffers allocated afterwards are
    first-touched there and the copy engines do not cross the socket interconnect).  Returns what was done."""
    node = gpu_numa_node(index)
    info = {"numa_node": node, "cpus": None, "source": None}
    cpus = None
    try:
        if node is not None:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus, info["source"] = set(_cpulist(f.read())), "sysfs"
        else:
            cpus, info["source"] = nvml_cpu_affinity(index), "nvml"
        allowed = set(os.sched_getaffinity(0))
        use = sorted((cpus or set()) & allowed)
        if use and len(use) < len(allowed):
            os.sched_setaffinity(0, use)
            info["cpus"] = len(use)
    except Exception as e:      # placement is an optimisation only
        info["error"] = repr(e)[:120]
    return info


def nvml_cpu_affinity(index: int):
    """CPUs NVML calls ideal for the GPU (the cores of its NUMA node), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        return {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
    except Exception:
        return None


class HostPipeline:
    """step k: pinned host arrays -> device, one OM kernel call, result arrays -> pinned host; three streams, two slots."""

    def __init__(self, m, kernel: str, arrays: List[str], slots: int = 2):
        self.m, self.kernel, self.arrays, self.n = m, kernel, list(arrays), slots
        dev = m.device
        self.h2d = torch.cuda.Stream(dev)
        self.d2h = torch.cuda.Stream(dev)
        shape = lambda a: (m.nzl, m.nyl, m.nx) if m.dim3 else (m.nyl, m.nx)
        dt = lambda a: m.cur[m.index[a]].dtype
        self.stage_in = [{a: torch.empty(shape(a), dtype=dt(a), device=dev) for a in arrays} for _ in range(slots)]
        self.stage_out = [{a: torch.empty(shape(a), dtype=dt(a), device=dev) for a in arrays} for _ in range(slots)]
        self.ev_in = [None] * slots       # upload of the slot finished
        self.ev_used = [None] * slots     # compute consumed stage_in[slot] and filled stage_out[slot]
        self.ev_out = [None] * slots      # download of the slot finished
        self.k = 0
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.stage_in[0].values())
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.stage_out[0].values())

    def submit(self, host_in: Dict[str, torch.Tensor], host_out: Dict[str, torch.Tensor]):
        """Queue one step; returns immediately.  `host_in` / `host_out`: pinned host tensors of the local slab per array."""
        m, b = self.m, self.k % self.n
        self.k += 1
        comp = torch.cuda.current_stream(m.device)
        with torch.cuda.stream(self.h2d):
            if self.ev_used[b] is not None:
                self.h2d.wait_event(self.ev_used[b])          # the step that last read this slot has consumed it
            for a in self.arrays:
                self.stage_in[b][a].copy_(host_in[a].view(self.stage_in[b][a].shape), non_blocking=True)
            self.ev_in[b] = torch.cuda.Event()
            self.ev_in[b].record(self.h2d)
        comp.wait_event(self.ev_in[b])
        if self.ev_out[b] is not None:
            comp.wait_event(self.ev_out[b])                   # stage_out[b] has left for the host
        for a in self.arrays:
            m.set_from_device(a, self.stage_in[b][a])
        m.call(self.kernel)
        m._join_comm()
        for a in self.arrays:
            m.interior_into(a, self.stage_out[b][a])
        self.ev_used[b] = torch.cuda.Event()
        self.ev_used[b].record(comp)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.ev_used[b])
            for a in self.arrays:
                host_out[a].view(self.stage_out[b][a].shape).copy_(self.stage_out[b][a], non_blocking=True)
            self.ev_out[b] = torch.cuda.Event()
            self.ev_out[b].record(self.d2h)

    def drain(self):
        """Wait until every queued step's result is in its host buffer (the compute stream waits; the host does not)."""
        comp = torch.cuda.current_stream(self.m.device)
        for e in self.ev_out:
            if e is not None:
                comp.wait_event(e)
