import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, sys
from tests.test_gpu_parity import _hydro_pair, _conserved, NAMES
for size, steps in [((512,512),20), ((1024,1024),20), ((1024,1024),60)]:
    m, o = _hydro_pair(size, fast=True)
    for n in NAMES: m.set(n, o.array(n), with_margin=True)
    for t in range(steps):
        m.call("proceed"); o.call("proceed")
    ca = _conserved(lambda n: m.get(n)); cb = _conserved(lambda n: o.interior(n))
    print(size, steps, [float(np.max(np.abs(a-b))/np.max(np.abs(b))) for a,b in zip(ca,cb)], abs(m.scalar("time")-o.scalar("time")[0])/o.scalar("time")[0])
