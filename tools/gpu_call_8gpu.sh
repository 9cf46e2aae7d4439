#!/bin/bash
# 8 GPUs: scaling bench (ResNet-50 full graph), VGG-16 (553 MB exchange), GoogLeNet, allreduce sweep
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/m8_smi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/m8_bench.json 2> gpurun_out/m8_bench.err
NCCL_DEBUG=INFO timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/allreduce_sweep.py > gpurun_out/m8_allreduce.log 2>&1
grep -E "NVLS|^\{" gpurun_out/m8_allreduce.log | head -80 > gpurun_out/m8_allreduce_summary.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 8 --model vgg16 --steps 5 --warmup 3 > gpurun_out/m8_bench_vgg16.json 2> gpurun_out/m8_bench_vgg16.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus 8 --model googlenet --steps 5 --warmup 3 > gpurun_out/m8_bench_googlenet.json 2> gpurun_out/m8_bench_googlenet.err
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/m8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/m8_tests.log
echo done
