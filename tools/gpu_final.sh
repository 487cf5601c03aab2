#!/bin/bash
# One-call GPU validation of the round's final code: GPU tests, bench lines, ncu capture + launch list of the Hydro flux kernel.
# Outputs land in gpurun_out/ (copied to profiles/ afterwards).  Every step has its own timeout.
set +e
T=${1:-r1i}
mkdir -p gpurun_out
( time timeout 260 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -3 gpurun_out/${T}_gpu_tests.log
timeout 120 python bench.py > gpurun_out/${T}_bench_life.json 2> gpurun_out/${T}_bench_life.err
cat gpurun_out/${T}_bench_life.json
timeout 120 python bench.py --workload hydro > gpurun_out/${T}_bench_hydro_fast.json 2> gpurun_out/${T}_bench_hydro_fast.err
cat gpurun_out/${T}_bench_hydro_fast.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:proceed_stage1 -s 3 -c 1 -f -o gpurun_out/${T}_hydro_fast \
    python tools/profile_run.py hydro 5 fast > gpurun_out/${T}_ncu_hydro.log 2>&1
tail -2 gpurun_out/${T}_ncu_hydro.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_hydro_fast.csv \
    python bench.py --workload hydro --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench_hydro.log 2>&1
timeout 90 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
cat gpurun_out/${T}_bench_reference.json
timeout 60 python bench.py --workload hydro --exact --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_hydro_exact.json 2>/dev/null
cat gpurun_out/${T}_bench_hydro_exact.json
