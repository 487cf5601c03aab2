"""Regenerate tests/golden/driver_*.txt: stdout of the reference's own example drivers (examples/HelloWorld/main.cpp,
examples/ShiftExample/main.cpp for dist-open and dist-cyclic, examples/Life/main.cpp's first frames) compiled unchanged
against the oracle's reference-style C++ class (oracle/plantrans.py) and run on the CPU.

    python tests/golden/make_driver_goldens.py        (needs /root/reference; run in the build container)

examples/Hydro/main-kh.cpp (1024^2 double, runs to t = 1) is stopped after its first snapshot: driver_hydro.json holds the
printed times and a numeric digest of the snapshot.
examples/HelloGPU/main.cu is the HelloWorld driver again (same program, language = CUDA): it shares driver_helloworld.txt.
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refdrivers  # noqa: E402


def main():
    with tempfile.TemporaryDirectory() as tmp:
        for key in refdrivers.KEYS:
            exe = refdrivers.link_oracle(key, tmp)
            text = refdrivers.run(key, exe)
            path = refdrivers.golden_path(key)
            if key == "hellogpu":
                assert text == open(path).read(), "HelloGPU driver output differs from HelloWorld's"
                continue
            with open(path, "w") as f:
                f.write(text)
            print(path, len(text), "bytes")
        import json
        got = refdrivers.run_hydro(refdrivers.link_oracle_hydro(tmp))
        with open(os.path.join(refdrivers.GOLDEN, "driver_hydro.json"), "w") as f:
            json.dump(dict(_comment="examples/Hydro/main-kh.cpp, unchanged, on the oracle's reference-style class (1024^2 double): "
                                    "first printed times, column sums and diagonal cells of output1/snapshot0000.txt", **got), f, indent=1)
        print(got)
        heart = refdrivers.heart_digest(refdrivers.run_heart(refdrivers.link_heart("oracle", tmp)))
        with open(os.path.join(refdrivers.GOLDEN, "driver_initialcondition.json"), "w") as f:
            json.dump(dict(_comment="examples/InitialCondition/main.cpp, unchanged, on the oracle's reference-style class: digest of heart.txt "
                                    "(500 x 500 lines `x y atan(...)`, six significant digits)", **heart), f, indent=1)
        print({k: v for k, v in heart.items() if k != "samples"})


if __name__ == "__main__":
    main()
