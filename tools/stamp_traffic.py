"""profiles/traffic.json from `ncu --set full` captures: DRAM bytes per launch of the dominant kernels, stamped with the hash of the
generated kernel source (and the runtime header it includes) they were taken from (bench.py reports `roofline.traffic` only while that source is the one loaded).
usage: stamp_traffic.py <life.ncu-rep> <hydro_fast.ncu-rep> <hydro_exact.ncu-rep>"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "paraiso_b200", "_generated")
ENTRIES = [("om_Life_proceed_stage0", "Life_CC/Life_kernels.cu", 2 * 4 * 16384 * 16384),
           ("om_Hydro_proceed_stage1_fast", "Hydro_OO_Double_fast/Hydro_kernels.cu", 64 * 4096 * 4096),
           ("om_Hydro_proceed_stage1_exact", "Hydro_OO_Double/Hydro_kernels.cu", 64 * 4096 * 4096)]


def dram_bytes(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(k)
        tot += float(r[i]) * scale[units[i]]
    return int(tot)


if __name__ == "__main__":
    table = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of one launch at the bench size, from the ncu --set full captures "
                         "summarised in this directory; kernel_source_sha1_16 = sha1 of the generated <Name>_kernels.cu + om_runtime.cuh the capture ran"}
    for (key, src, alg), rep in zip(ENTRIES, sys.argv[1:4]):
        with open(os.path.join(GEN, src), "rb") as f, open(os.path.join(GEN, os.path.dirname(src), "om_runtime.cuh"), "rb") as r:
            h = hashlib.sha1(f.read() + r.read()).hexdigest()[:16]
        b = dram_bytes(rep)
        table[key] = {"dram_bytes_per_launch": b, "algorithmic_bytes": alg, "ratio": round(b / alg, 4), "kernel_source_sha1_16": h,
                      "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "_ncu.txt")}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(table, f, indent=1)
    print(json.dumps(table, indent=1))
