#!/bin/bash
# Multi-GPU validation on N GPUs of one box: NCCL parity tests (incl. graph replay), the C++ class on several devices, the bench line.
set +e
N=${1:-2}; T=${2:-r2}
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_host_class.py -m gpu -q ) > gpurun_out/${T}_gpu_tests_${N}gpu.log 2>&1
tail -3 gpurun_out/${T}_gpu_tests_${N}gpu.log
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 ) \
    > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
tail -c 300 gpurun_out/${T}_bench_${N}gpu.err
