#!/bin/bash
# 2 GPUs, final build: the multi-rank correctness tests only
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 240 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/m2c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/m2c_tests.log
echo done
