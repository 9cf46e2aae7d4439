#!/bin/bash
# 8 GPUs (charged 8x: keep it short): ResNet-50 full graph with the default 6 buckets and with 2, allreduce sweep, VGG-16 (553 MB exchange)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/m8_smi.txt 2>&1
run() { port=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port "$@"; }
run 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/m8_bench.json 2> gpurun_out/m8_bench.err
run 29555 bench.py --gpus 8 --steps 10 --warmup 3 --buckets 2 > gpurun_out/m8_bench_b2.json 2> gpurun_out/m8_bench_b2.err
NCCL_DEBUG=INFO run 29552 tools/allreduce_sweep.py > gpurun_out/m8_allreduce.log 2>&1
grep -E "NVLS|^\{" gpurun_out/m8_allreduce.log | head -80 > gpurun_out/m8_allreduce_summary.log
run 29553 bench.py --gpus 8 --model vgg16 --steps 5 --warmup 3 > gpurun_out/m8_bench_vgg16.json 2> gpurun_out/m8_bench_vgg16.err
echo done
