"""Hydro stage-1 sweep on the GPU box: CTA width x build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib, json, os, sys
import torch

def run(cfg, size=(4096, 4096), steps=10):
    for k, v in cfg.items():
        os.environ[k] = str(v)
    import paraiso_b200.generator.b200.cuda as C, paraiso_b200.generator.b200.warpstream as W, paraiso_b200.generator.b200.emit as E
    import paraiso_b200.build as Bd, paraiso_b200.machines as M
    for mod in (C, W, E, Bd, M):
        importlib.reload(mod)
    fast = bool(int(cfg.get("FAST", 1)))
    from paraiso_b200.examples.hydro import hydro_om, hydro_setup
    setup = hydro_setup(); setup.fast_math = fast
    tag = "Hydro_sweep_" + "_".join(f"{k}{v}" for k, v in sorted(cfg.items()))
    desc, so = Bd.build_machine(setup, hydro_om("master"), tag=tag, fmad=fast)
    from paraiso_b200.runtime import Machine
    m = Machine(desc, so, size=size)
    M.hydro_set_params(m, size)
    m.call("init")
    for _ in range(3):
        m.call("proceed")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m.call("proceed")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = m.kernels["proceed"]["stages"][1]
    return dict(cfg=cfg, ms=ms, gcells=size[0] * size[1] / ms / 1e6, occ=getattr(m.lib, st["symbol"] + "_occupancy")(), smem=st["smem"], chunk_rows=m._geom(st).chunk_rows)

if __name__ == "__main__":
    cfgs = json.loads(sys.argv[1])
    for c in cfgs:
        try:
            print(json.dumps(run(c)), flush=True)
        except Exception as e:
            print("FAIL", c, repr(e)[:400], flush=True)
