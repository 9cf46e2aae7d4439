#!/bin/bash
# 8 GPUs, second pass (charged 8x): final build, in-step bucket bandwidth with the fixed timing, communicator CTA cap A/B
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { port=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port "$@"; }
run 29571 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/m8b_bench.json 2> gpurun_out/m8b_bench.err
B2C_NCCL_MAX_CTAS=8 run 29572 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/m8b_bench_cta8.json 2> gpurun_out/m8b_bench_cta8.err
echo done
