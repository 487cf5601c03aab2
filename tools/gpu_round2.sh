#!/bin/bash
# One-call 1-GPU validation of the round-2 code: all GPU tests, the bench line (Life + both Hydro builds), the reference arm,
# ncu --set full captures of the three dominant kernels and the launch list of bench.py itself.  Outputs land in gpurun_out/.
set +e
T=${1:-r2n}
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -4 gpurun_out/${T}_gpu_tests.log
( time timeout 300 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -c 300 gpurun_out/${T}_bench_n1.err
( time timeout 200 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
tail -c 200 gpurun_out/${T}_bench_reference.err
timeout 150 ncu --set full --clock-control none --import-source on -k regex:om_Life_proceed_stage0 -s 3 -c 1 -f -o gpurun_out/${T}_life \
    python tools/profile_run.py life 5 > gpurun_out/${T}_ncu_life.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:om_Hydro_proceed_stage1 -s 3 -c 1 -f -o gpurun_out/${T}_hydro_fast \
    python tools/profile_run.py hydro 5 fast > gpurun_out/${T}_ncu_hydro_fast.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:om_Hydro_proceed_stage1 -s 3 -c 1 -f -o gpurun_out/${T}_hydro_exact \
    python tools/profile_run.py hydro 5 exact > gpurun_out/${T}_ncu_hydro_exact.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${T}_ncu_bench.log 2>&1
ls -la gpurun_out | grep ${T}
