#!/usr/bin/env python
"""Headline benchmark: Gcell-updates/s of Paraiso-generated Life (default) or Hydro on B200.

Contract (see README / DESIGN.md §6): `python bench.py --gpus N --steps K --warmup W` — one process per
GPU (torchrun for N > 1), W warm-up steps, K timed steps between barrier + synchronize, CUDA events,
max over ranks, ONE JSON line on rank 0.  A step is one `proceed()` of the generated machine over this
rank's slab; scaling is weak (16384^2 Life cells, or 4096^2 Hydro cells, per GPU).
`--impl reference` times the reference-style native C++ (oracle/plantrans.py emission: flat OpenMP
loops, recompute per cursor, serial reduce, copy on store) on the host cores for the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (per-GPU size, algorithmic bytes per cell update, dtype)
    "life": ((16384, 16384), 8, "i32"),
    "hydro": ((4096, 4096), 64, "f64"),
    # BASELINE.json configs[3]: one 32768^2 grid slab-decomposed over the GPUs (strong scaling; needs >= 2 GPUs)
    "hydro32k": ((32768, 32768), 64, "f64"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_machine(workload: str, size):
    """Reference-style C++ for this workload, built with -O3 -fopenmp."""
    from oracle.cpu import OracleMachine
    if workload == "life":
        from paraiso_b200.examples.life import life_om, life_setup
        from paraiso_b200.machines import life_seed
        o = OracleMachine(life_setup("master", size=size), life_om("master"), openmp=True, opt="-O3")
        o.call("init")
        o.interior("cell")[...] = life_seed(size[0], 0, size[1])
    else:
        from paraiso_b200.examples.hydro import hydro_om, hydro_setup
        o = OracleMachine(hydro_setup(size), hydro_om("master"), openmp=True, opt="-O3")
        for k, v in dict(time=0.0, cfl=0.5, extent0=1.0, extent1=1.0, dR0=1.0 / size[0], dR1=1.0 / size[1]).items():
            o.scalar(k)[0] = v
        o.call("init")
    return o


def time_cpu(workload: str, size, steps: int, warmup: int):
    os.environ.setdefault("OMP_NUM_THREADS", str(cpu_threads()))
    o = oracle_machine(workload, size)
    for _ in range(warmup):
        o.call("proceed")
    t0 = time.perf_counter()
    for _ in range(steps):
        o.call("proceed")
    dt = time.perf_counter() - t0
    return size[0] * size[1] * steps / dt / 1e9, dt / steps * 1e3


def fp64_roofline(m, stage: int, kernel_ms: float, clocks):
    """{"achieved", "peak", "unit", "frac", ...} of the FP64 pipe for one launch of proceed's stage `stage`, or None when
    cuobjdump is not available.  Static instruction counts come from paraiso_b200.costmodel (the SASS of the loaded library)."""
    try:
        from paraiso_b200 import costmodel
        e = costmodel.estimate_stage(m.desc, m.lib._name, kernel="proceed", stage=stage, size=(m.nx, m.nyl))
        sm_hz = float((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
        sms = 148
        try:
            import torch
            sms = torch.cuda.get_device_properties(m.device).multi_processor_count
        except Exception:
            pass
        warp_rows = m.nx * m.nyl / 32.0 * e.overhead
        achieved = warp_rows * e.fp64 / (kernel_ms * 1e-3) / 1e9          # G warp-instructions / s
        peak = sms * 4 * 0.5 * sm_hz / 1e9
        return {"bound": "fp64_pipe", "achieved": achieved, "peak": peak, "unit": "G warp-instr/s", "frac": achieved / peak,
                "fp64_instr_per_warp_row": e.fp64, "instr_per_warp_row": e.instructions, "overhead": e.overhead,
                "registers": e.registers, "ctas_per_sm": e.ctas_per_sm,
                "source": "static SASS count of the row loop (cuobjdump) x measured kernel time; ncu: profiles/r1i_hydro_fast_ncu.txt"}
    except Exception as ex:     # measurement garnish only: never lose the bench line over it
        return {"bound": "fp64_pipe", "unavailable": repr(ex)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("OM_BENCH_WORKLOAD", "life"), choices=list(WORKLOADS))
    ap.add_argument("--fmad", action="store_true", help="Hydro: FMA-contracted build (within 1e-12, not bit-exact)")
    ap.add_argument("--fast", action="store_true", help="Hydro: Setup.fast_math build (FMA + fast division/sqrt; ~1e-15 relative after 20 steps, inside the 1e-12 north-star tolerance)")
    ap.add_argument("--exact", action="store_true", help="Hydro: bit-exact build (-fmad=false, IEEE division): the default for Hydro is --fast")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload.startswith("hydro") and not args.exact and not args.fmad:
        args.fast = True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    size1, alg_bytes, dtype = WORKLOADS[args.workload]
    cfg_name = {"life": "Life 16384x16384 Int32 periodic per GPU (examples/Life/Generator.hs, Cyclic)",
                "hydro": "Hydro 2D Euler KH 4096x4096 double per GPU (examples/Hydro/HydroMain.hs, Open)",
                "hydro32k": "Hydro 2D Euler KH 32768x32768 double, slab-decomposed over the GPUs (examples/Hydro/HydroMain.hs, Open)"}[args.workload]
    strong = args.workload == "hydro32k"

    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: Life 4096x4096 / Hydro 1024x1024 (per-cell cost is size independent once out of cache)
        sample = (4096, 4096) if args.workload == "life" else (1024, 1024)
        args.workload = "life" if args.workload == "life" else "hydro"
        steps = max(1, min(args.steps, 5 if args.workload == "life" else 3))
        warm = min(args.warmup, 1)
        v, ms = time_cpu(args.workload, sample, steps, warm)
        thr = cpu_threads()
        line = {"impl": "reference", "metric": "Gcell-updates/s", "value": v, "unit": "Gcell/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": {"workload": cfg_name, "sample": f"{sample[0]}x{sample[1]}"},
                "cpu_baseline": {"value": v, "unit": "Gcell/s", "cores": thr, "kind": "port",
                                 "sample": f"{steps} proceed() steps of {args.workload} {sample[0]}x{sample[1]}, reference-style C++ (-O3 -fopenmp)"},
                "e2e": {"value": v, "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from paraiso_b200.machines import hydro_machine, hydro_set_params, life_machine, life_seed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gsize = size1 if strong else (size1[0], size1[1] * world)      # weak scaling: slabs stacked along the outermost axis
    kw = dict(device=dev, rank=rank, nranks=world)
    if args.workload == "life":
        m = life_machine(gsize, **kw)
        m.call("init")
        host = torch.from_numpy(life_seed(gsize[0], m.y0, m.nyl, nx_global=gsize[0])).pin_memory()
        m.set_from_host("cell", host)
        state = ["cell"]
        result_scalar = "population"
    else:
        m = hydro_machine(gsize, fmad=args.fmad, fast=args.fast, **kw)
        hydro_set_params(m, gsize)
        m.call("init")
        state = ["density", "velocity0", "velocity1", "pressure"]
        result_scalar = "time"
    cells = m.nx * m.nyl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    step = lambda: m.call("proceed")
    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = m.launches
    ms = timed(step, args.steps)
    launches = m.launches - l0
    # keep the same load running until nvidia-smi has delivered a few samples (its period is 100 ms,
    # a timed region can be shorter); the clocks reported are those seen under this load
    extra = int(max(1, min(20000, 1200.0 / max(ms / args.steps, 1e-3))))   # ~1.2 s, the same count on every rank
    for _ in range(extra):
        step()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop() if rank == 0 else None
    value = cells * world * args.steps / (ms * 1e-3) / 1e9

    # dominant kernel alone (the last array stage of proceed): CUDA events around back-to-back launches
    kinfo = m.kernels["proceed"]
    dom = len(kinfo["stages"]) - 1
    for _ in range(3):
        m.call_stage("proceed", dom)
    kms = timed(lambda: m.call_stage("proceed", dom), args.steps) / args.steps
    peak, peak_src = peaks()
    achieved = cells * alg_bytes / (kms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from ncu --set full
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(kinfo["stages"][dom]["symbol"] + ("_fast" if args.fast else ""), {}).get("dram_bytes_per_launch")

    # Hydro's flux kernel is bound by the FP64 pipe, not by HBM — report that roof next to the HBM fraction.
    # FP64 warp instructions per launch = static count in the row loop (cuobjdump) x warp-rows x halo / warm-up overhead;
    # the pipe takes one warp instruction every two cycles per scheduler (64 FP64 lanes per SM).
    fp64_pipe = fp64_roofline(m, dom, kms, clocks) if (rank == 0 and args.workload.startswith("hydro")) else None

    # end to end through the public host API: pinned host state -> device, proceed(), result scalar -> host
    pinned = {n: torch.from_numpy(np.ascontiguousarray(m.get(n))).pin_memory() for n in state}
    h2d = sum(t.numel() * t.element_size() for t in pinned.values())

    def e2e_step():
        for n in state:
            m.set_from_host(n, pinned[n])
        m.call("proceed")
        m.scalar(result_scalar)
    for _ in range(2):
        e2e_step()
    esteps = max(3, min(args.steps, 10))
    ems = timed(e2e_step, esteps)
    e2e_value = cells * world * esteps / (ems * 1e-3) / 1e9

    line = None
    if rank == 0:
        line = {"metric": "Gcell-updates/s", "value": value, "unit": "Gcell/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": {"workload": cfg_name, "global_grid": f"{gsize[0]}x{gsize[1]}", "per_gpu_grid": f"{m.nx}x{m.nyl}",
                           "decomposition": f"slab{world}" if world > 1 else "single", "l2": "state arrays are larger than L2 (no flush needed)",
                           "build": ("fast_math (FMA, MUFU-seeded div/sqrt; within 1e-12 of the reference)" if args.fast else
                                     "fmad=true" if args.fmad else "fmad=false (bit-exact vs reference C++)")},
                "roofline": {"bound": "hbm", "kernel": kinfo["stages"][dom]["symbol"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_cell": alg_bytes, "kernel_ms": kms, "fp64_pipe": fp64_pipe},
                # (h2d_gbs: the state upload alone bounds the end-to-end step — PCIe, not the kernel)
                "e2e": {"value": e2e_value, "unit": "Gcell/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                        "ms_per_step": ems / esteps, "h2d_gbs": h2d / (ems / esteps * 1e-3) / 1e9},
                "gpu_launches": launches, "clocks": clocks}
        if not args.no_cpu_baseline and world == 1:   # the CPU leg runs at N = 1 only (ranks of a multi-GPU job do not wait for it)
            sample = (4096, 4096) if args.workload == "life" else (1024, 1024)
            csteps = 5 if args.workload == "life" else 3
            v, _ = time_cpu("life" if args.workload == "life" else "hydro", sample, csteps, 1)
            line["cpu_baseline"] = {"value": v, "unit": "Gcell/s", "cores": cpu_threads(), "kind": "port",
                                    "sample": f"{csteps} proceed() steps of {args.workload} {sample[0]}x{sample[1]}, reference-style C++ (-O3 -fopenmp)"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
