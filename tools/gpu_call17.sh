#!/bin/bash
# BatchNorm backward: cp.async prefetch of the parked streams (opt-in B2C_BN_PREFETCH_BWD=1) against the register-load path
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export B2C_BN_PREFETCH_BWD=1
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_trainer_gpu.py -m gpu -q > gpurun_out/c17_tests_bn.log 2>&1; echo "rc=$?" >> gpurun_out/c17_tests_bn.log
timeout 200 python tools/bn_sweep.py > gpurun_out/c17_bn.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c17_bench_pre.json 2> gpurun_out/c17_bench_pre.err
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c17_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c17_tests.log
export B2C_BN_PREFETCH_BWD=0
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c17_bench_nopre.json 2> gpurun_out/c17_bench_nopre.err
echo done
