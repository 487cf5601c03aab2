#!/usr/bin/env python
"""Drive the B200 backend from a dump written by the Haskell front-end.

    python tools/om2b200.py output/OM.txt --size 80x48 --boundary cyclic,cyclic --out dist-b200
    python tools/om2b200.py examples-old/Life-exampled/output/OM.txt --cpp examples-old/Life-exampled/dist/Life.cpp \
        --size 128x128 --out dist-b200 [--compile]

The dump is what `prettyPrintA1` writes (OM/PrettyPrint.hs:34-35; every example calls it, e.g.
examples/Life/Generator.hs:30).  Dumps of the old printer (`Imm <<Int>>`) carry no immediate values: pass the C++ the
same generator run wrote (`--cpp`) and they are read from there (`om.interchange.recover_immediates`).  `--size` and
`--boundary` are the `localSize` / `boundary` fields of `Native.Setup` (Generator/Native.hs:16-24), which the dump does not
hold.  Writes <Name>.hpp, <Name>.cpp, <Name>_kernels.cu, <Name>_abi.h, <Name>_abi.json, om_runtime.cuh; `--compile` also
runs nvcc for sm_100a (no GPU needed) and leaves libom_<Name>.so next to them."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from paraiso_b200.annotation import CYCLIC, OPEN  # noqa: E402
from paraiso_b200.generator.b200.emit import generateIO  # noqa: E402
from paraiso_b200.generator.native import Setup  # noqa: E402
from paraiso_b200.om.interchange import parse_om, recover_immediates  # noqa: E402


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("dump", help="OM.txt written by prettyPrintA1")
    ap.add_argument("--cpp", help="reference-generated <Name>.cpp of the same run (only for old dumps without immediates)")
    ap.add_argument("--size", required=True, help="localSize, e.g. 16384x16384 (axis 0 first)")
    ap.add_argument("--boundary", help="open|cyclic per axis, comma separated (default: open everywhere)")
    ap.add_argument("--out", default="./dist-b200", help="output directory (Native.directory)")
    ap.add_argument("--gpus", type=int, default=1, help="slabs along the outermost axis")
    ap.add_argument("--fast-math", action="store_true")
    ap.add_argument("--tune", action="append", default=[], metavar="KNOB=VALUE",
                    help="a field of generator.native.Tuning, e.g. --tune prefetch_rows=3 --tune chunk_rows_light=24")
    ap.add_argument("--compile", action="store_true", help="also build libom_<Name>.so with nvcc (sm_100a)")
    a = ap.parse_args(argv)

    size = tuple(int(x) for x in a.size.lower().split("x"))
    names = {"open": OPEN, "cyclic": CYCLIC}
    bnd = tuple(names[b.strip().lower()] for b in a.boundary.split(",")) if a.boundary else tuple(OPEN for _ in size)
    if len(bnd) != len(size):
        ap.error("--boundary needs one entry per axis of --size")
    table = None
    if a.cpp:
        with open(a.cpp) as f:
            table = recover_immediates(f.read())
    with open(a.dump) as f:
        om = parse_om(f.read(), dim=len(size), immediates=table)
    setup = Setup(local_size=size, boundary=bnd, directory=a.out, gpus=a.gpus, fast_math=a.fast_math)
    for kv in a.tune:
        k, _, v = kv.partition("=")
        if not hasattr(setup.tuning, k):
            ap.error(f"unknown tuning knob {k!r}")
        old = getattr(setup.tuning, k)
        setattr(setup.tuning, k, (v.lower() in ("1", "true", "yes")) if isinstance(old, bool) else type(old)(v) if old is not None else int(v))
    for path, _text in generateIO(setup, om):
        print(path)
    if a.compile:
        from paraiso_b200.build import compile_kernels
        print(compile_kernels(a.out, om.name, fmad=a.fast_math))
    return 0


if __name__ == "__main__":
    sys.exit(main())
