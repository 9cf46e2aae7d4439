#!/bin/bash
# 8 GPUs, second pass (charged 8x): final build, in-step bucket bandwidth with the fixed timing, communicator CTA cap A/B
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
# the LRN fast path (beta = 0.75) is new in this build: check it on one GPU first
timeout 300 python -m pytest tests/test_layers_extra_gpu.py tests/test_trainer_gpu.py -m gpu -q -k "lrn or inception" > gpurun_out/m8b_pretest.log 2>&1; echo "rc=$?" >> gpurun_out/m8b_pretest.log
run() { port=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port "$@"; }
run 29571 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/m8b_bench.json 2> gpurun_out/m8b_bench.err
B2C_NCCL_MAX_CTAS=8 run 29572 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/m8b_bench_cta8.json 2> gpurun_out/m8b_bench_cta8.err
run 29573 bench.py --gpus 8 --model alexnet --steps 5 --warmup 3 > gpurun_out/m8b_bench_alexnet.json 2> gpurun_out/m8b_bench_alexnet.err
echo done
