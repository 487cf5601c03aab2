#!/bin/bash
# A/B on one box: Life with the rare block inline / out of line (and the library from before the per-thread ghost predicate),
# the fast-math division / sqrt error bounds, and a short bench line of the rebuilt Hydro fast_math kernel.
set +e
T=${1:-r2ad}
mkdir -p gpurun_out
( time timeout 300 python tools/variants_life_rare.py ) > gpurun_out/${T}_life_rare.jsonl 2> gpurun_out/${T}_life_rare.err
cat gpurun_out/${T}_life_rare.jsonl
( time timeout 300 python -m pytest tests/test_gpu_divsqrt.py -q ) > gpurun_out/${T}_divsqrt.log 2>&1
tail -3 gpurun_out/${T}_divsqrt.log
( time timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e ) > gpurun_out/${T}_bench_short.json 2> gpurun_out/${T}_bench_short.err
python - <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/%s_bench_short.json" % sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2ad_bench_short.json").read().strip().splitlines()[-1])
    print("life", d["value"], d["ms_per_step"], d.get("verified"))
    for k, w in d.get("workloads", {}).items():
        print(k, w.get("value"), w.get("ms_per_step"), w.get("verified"), w.get("roofline", {}).get("frac"))
except Exception as e:
    print("bench parse failed", e)
PY
