"""Opcode histogram of a generated library's kernels (whole kernel and its row loop) from `cuobjdump -sass`:
the static evidence kept next to each ncu summary under profiles/ (which pipes the instructions go to, LDGSTS / UBLKCP
staging, 128-bit accesses, FP64 share, spills).  usage: sass_histogram.py <lib.so> [kernel-substring ...]"""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.costmodel import FP64_OPS, sass_functions  # noqa: E402


def loop_body(insts):
    best = (0, 0)
    for addr, op, text in insts:
        if op.startswith("BRA"):
            t = re.findall(r"0x([0-9a-f]+)", text)
            if t and int(t[-1], 16) < addr and addr - int(t[-1], 16) > best[1] - best[0]:
                best = (int(t[-1], 16), addr)
    return [i for i in insts if best[0] <= i[0] <= best[1]] if best[1] else insts


def report(name, insts):
    for label, body in (("whole kernel", insts), ("row loop", loop_body(insts))):
        h = collections.Counter(op for _a, op, _t in body)
        fam = collections.Counter(op.split(".")[0] for _a, op, _t in body)
        fp64 = sum(n for op, n in fam.items() if op in FP64_OPS)
        print(f"{name}: {label}: {len(body)} instructions, FP64 pipe {fp64}, "
              f"LDS {fam['LDS']}, STS {fam['STS']}, LDG {fam['LDG']}, STG {fam['STG']}, LDGSTS {fam['LDGSTS']}, UBLKCP {fam['UBLKCP']}, "
              f"MUFU {fam['MUFU']}, BAR {fam['BAR']}, local (spill) {fam['LDL'] + fam['STL']}")
        print("   " + ", ".join(f"{op} {n}" for op, n in sorted(h.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    funcs = sass_functions(sys.argv[1])
    pats = sys.argv[2:] or [""]
    for name, insts in funcs.items():
        if any(p in name for p in pats):
            report(name, insts)
