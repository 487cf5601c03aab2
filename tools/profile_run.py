"""Profiling driver (run under ncu on the GPU box): a few proceed() steps of one workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys
import torch
from paraiso_b200.machines import hydro_machine, hydro_set_params, life_machine, life_seed

wl = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if wl == "life":
    size = (16384, 16384)
    m = life_machine(size)
    m.call("init")
    m.set("cell", life_seed(size[0], 0, size[1]))
else:
    size = (4096, 4096)
    mode = sys.argv[3] if len(sys.argv) > 3 else "exact"
    m = hydro_machine(size, fmad=(mode == "fma"), fast=(mode == "fast"))
    hydro_set_params(m, size)
    m.call("init")
for _ in range(steps):
    m.call("proceed")
torch.cuda.synchronize()
print("done", wl, steps)
