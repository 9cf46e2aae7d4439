#!/bin/bash
# BatchNorm launch size for the 14x14 layers (4 097 .. 16 384 values per channel): 512 (default) vs 256 vs 128 threads
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for v in 512 256 128; do echo "== B2C_BN_THREADS_MID=$v" >> gpurun_out/c15_bn.log; B2C_BN_THREADS_MID=$v timeout 200 python tools/bn_sweep.py >> gpurun_out/c15_bn.log 2>&1; done
for v in 256 128; do B2C_BN_THREADS_MID=$v timeout 300 python -m pytest tests/test_layers_gpu.py -m gpu -q -k batchnorm > gpurun_out/c15_tests_$v.log 2>&1; echo "rc=$?" >> gpurun_out/c15_tests_$v.log; B2C_BN_THREADS_MID=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c15_bench_$v.json 2> gpurun_out/c15_bench_$v.err; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c15_bench_512.json 2> gpurun_out/c15_bench_512.err
echo done
