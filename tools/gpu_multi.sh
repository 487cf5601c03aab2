#!/bin/bash
# Multi-GPU validation of the final code on N GPUs of one box: NCCL parity tests, the C++ class on several GPUs, bench lines.
set +e
N=${1:-2}; T=${2:-r1i}
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_host_class.py -m gpu -x -q ) > gpurun_out/${T}_gpu_tests_${N}gpu.log 2>&1
tail -4 gpurun_out/${T}_gpu_tests_${N}gpu.log
for wl in life hydro; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $wl \
      > gpurun_out/${T}_bench_${wl}_${N}gpu.json 2> gpurun_out/${T}_bench_${wl}_${N}gpu.err
  tail -1 gpurun_out/${T}_bench_${wl}_${N}gpu.json
done
