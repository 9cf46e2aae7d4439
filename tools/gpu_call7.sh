#!/bin/bash
# staged weight-gradient kernel (and the NT GEMM that rides on it): first run, parity, sweep, bench -- all with B2C_WGRAD_STAGED=1
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export B2C_WGRAD_STAGED=1
timeout 300 python tools/wstg_debug.py > gpurun_out/c7_debug.log 2>&1; echo "rc=$?" >> gpurun_out/c7_debug.log
if grep -q "raised\|timeouts recorded\|nan=True" gpurun_out/c7_debug.log; then
  timeout 300 compute-sanitizer --tool memcheck python tools/wstg_debug.py 2 64 8 8 64 1 0 > gpurun_out/c7_sanitizer.log 2>&1
  timeout 300 compute-sanitizer --tool memcheck python tools/wstg_debug.py 2 32 12 12 40 3 1 >> gpurun_out/c7_sanitizer.log 2>&1
  echo "debug failed"; exit 0
fi
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c7_parity.log 2>&1; echo "rc=$?" >> gpurun_out/c7_parity.log
timeout 600 python tools/layer_sweep.py resnet50 64 > gpurun_out/c7_sweep.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c7_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c7_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_stg -s 2 -c 1 -o gpurun_out/c7_wstg_3x3 python tools/one_layer.py 128 28 128 3 1 1 64 wgrad > gpurun_out/c7_ncu.log 2>&1
echo done
