"""Writes tests/golden/legacy_immediates.json: the immediate tables `recover_immediates` reads from the reference's own
generated C++ (examples-old/Life-exampled/dist/Life.cpp, examples-old/Hydro-exampled/dist/Hydro.cpp), so that the
legacy-dump import test also runs where /root/reference is absent.  Run here: python tests/golden/make_legacy_immediates.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from paraiso_b200.om.interchange import recover_immediates  # noqa: E402

REF = "/root/reference/examples-old"
out = {}
for key, path in [("life", "Life-exampled/dist/Life.cpp"), ("hydro", "Hydro-exampled/dist/Hydro.cpp")]:
    with open(os.path.join(REF, path)) as f:
        tab = recover_immediates(f.read())
    out[key] = {k: {str(i): lit for i, lit in sorted(v.items())} for k, v in tab.items()}
with open(os.path.join(os.path.dirname(__file__), "legacy_immediates.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
print({k: {kk: len(vv) for kk, vv in v.items()} for k, v in out.items()})
