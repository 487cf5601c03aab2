"""Parameter sweep for the Life proceed kernel (run on the GPU box): skeleton, CTA width, prefetch, waves."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itertools, json, os, sys, time
import torch

def run(cfg, size=(16384, 16384), steps=30):
    for k, v in cfg.items():
        os.environ[k] = str(v)
    tag = "Life_sweep_" + "_".join(f"{k[3:]}{v}" for k, v in sorted(cfg.items()) if k not in ("OM_WAVES", "OM_CHUNK_ROWS"))
    os.environ["OM_LIFE_TAG"] = tag
    import importlib
    import paraiso_b200.generator.b200.cuda as C, paraiso_b200.generator.b200.warpstream as W
    importlib.reload(C); importlib.reload(W)
    import paraiso_b200.generator.b200.emit as E
    importlib.reload(E)
    import paraiso_b200.build as Bd
    importlib.reload(Bd)
    import paraiso_b200.machines as M
    importlib.reload(M)
    m = M.life_machine(size)
    m.call("init")
    m.set("cell", M.life_seed(size[0], 0, size[1]))
    for _ in range(5):
        m.call_stage("proceed", 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m.call_stage("proceed", 0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = m.kernels["proceed"]["stages"][0]
    g = m._geom(st)
    return dict(cfg=cfg, ms=ms, gbs=size[0] * size[1] * 8 / ms / 1e6, occ=getattr(m.lib, st["symbol"] + "_occupancy")(), chunk_rows=g.chunk_rows, smem=st["smem"])

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "a"
    cfgs = []
    if which == "a":
        for nt, pf, w in itertools.product([32, 64, 128], [2, 4], [4, 16]):
            cfgs.append(dict(OM_MODE="ring", OM_NT=nt, OM_PF=pf, OM_WAVES=w))
        for nt, pf, w in itertools.product([64, 128], [3, 5], [4, 16]):
            cfgs.append(dict(OM_MODE="stream", OM_NT=nt, OM_PREFETCH=pf, OM_WAVES=w))
    elif which == "b":
        for nt, pf, cr in itertools.product([128, 256, 512], [1, 2, 3], [8, 16, 32, 64]):
            cfgs.append(dict(OM_MODE="ring", OM_NT=nt, OM_PF=pf, OM_CHUNK_ROWS=cr))
    elif which == "c":
        for mb, pf in itertools.product([0, 10, 12, 16], [2, 3]):
            cfgs.append(dict(OM_MODE="ring", OM_NT=128, OM_PF=pf, OM_MINBLOCKS=mb))
        for mb in (0, 6, 8):
            cfgs.append(dict(OM_MODE="ring", OM_NT=256, OM_PF=2, OM_MINBLOCKS=mb))
    else:
        cfgs = json.loads(sys.argv[2])
    for c in cfgs:
        try:
            print(json.dumps(run(c)), flush=True)
        except Exception as e:
            print("FAIL", c, repr(e)[:300], flush=True)
