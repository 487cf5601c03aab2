"""Generate the golden fixtures from the reference's OWN generated C++ (oracle/_ref, built by
oracle/Makefile from /root/reference/examples-old/*-exampled/dist).  Run where /root/reference is mounted:

    make -C oracle && python tests/golden/make_golden.py

The fixtures are small (bit-packed cells, CRC32 of whole arrays, a few scalars) and are committed, so the
oracle can be checked against the reference on machines that do not have /root/reference."""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.cpu import RefHydro, RefLife  # noqa: E402


def main():
    r = RefLife()
    r.init()
    pops, packed = [r.population()], {}
    for t in range(1, 101):
        r.proceed()
        pops.append(r.population())
        if t in (1, 10, 100):
            packed[f"cell_gen{t}"] = np.packbits(r.cell().astype(np.uint8))
    np.savez_compressed(os.path.join(HERE, "life_exampled.npz"), populations=np.array(pops, np.int32),
                        shape=np.array(r.cell().shape), **packed)
    h = RefHydro(openmp=True)
    h.setup_kh()
    h.init()
    out = {"program": "examples-old/Hydro-exampled (float, 1024x1024, Open, margin 3), main-kh.cpp parameter block",
           "steps": {}}
    names = ["density", "velocity0", "velocity1", "pressure"]
    out["init_crc32"] = {n: zlib.crc32(h.array(n).tobytes()) for n in names}
    for t in range(1, 11):
        h.proceed()
        if t in (1, 2, 3, 10):
            out["steps"][str(t)] = {
                "time_bits": int(h.scalar("time").view(np.uint32)[0]),
                "crc32": {n: zlib.crc32(h.array(n).tobytes()) for n in names},
                "sum_density_interior": float(h.array("density")[3:-3, 3:-3].astype(np.float64).sum()),
                "sum_pressure_interior": float(h.array("pressure")[3:-3, 3:-3].astype(np.float64).sum()),
            }
    with open(os.path.join(HERE, "hydro_exampled.json"), "w") as f:
        json.dump(out, f, indent=1)
    # every 8th interior cell of every state array after 3 and 10 steps (float32): lets a DIFFERENT transcription of the
    # program (master's HydroMain.hs, float) be compared with the reference's compiled output within the north-star's 1e-5
    h = RefHydro(openmp=True)
    h.setup_kh()
    h.init()
    samples = {}
    for t in range(1, 11):
        h.proceed()
        if t in (3, 10):
            for n in names:
                samples[f"{n}_step{t}"] = h.array(n)[3:-3, 3:-3][::8, ::8].copy()
            samples[f"time_step{t}"] = h.scalar("time").copy()
    np.savez_compressed(os.path.join(HERE, "hydro_exampled_samples.npz"), **samples)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
