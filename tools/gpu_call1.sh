#!/bin/bash
# round 2, GPU call 1: full suite (incl. full-size parity), experimental variants, in-kernel counters, sweeps, bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c1_tests.log
B2C_RUN_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_experimental_gpu.py -m gpu -q > gpurun_out/c1_experimental.log 2>&1; echo "rc=$?" >> gpurun_out/c1_experimental.log
for L in "128 28 128 3 1 1 64 fwd" "64 56 256 1 1 0 64 fwd" "256 56 64 1 1 0 64 fwd" "256 14 256 3 1 1 64 fwd" "512 7 512 3 1 1 64 fwd" "64 56 64 3 1 1 64 fwd" "64 56 64 3 1 1 64 wgrad" "256 14 256 3 1 1 64 wgrad"; do
  echo "== $L" >> gpurun_out/c1_prof.log
  B2C_PROF=1 timeout 120 python tools/one_layer.py $L >> gpurun_out/c1_prof.log 2>&1
done
timeout 600 python tools/layer_sweep.py resnet50 64 > gpurun_out/c1_sweep_default.txt 2>&1
B2C_SWEEP_ONLY="s2" B2C_WGRAD_COMPACT=1 timeout 300 python tools/layer_sweep.py resnet50 64 > gpurun_out/c1_sweep_compact.txt 2>&1
if grep -q "B2C_WGRAD3_TMA\] PASSED\|2 passed" gpurun_out/c1_experimental.log; then
  B2C_SWEEP_ONLY="k3" B2C_WGRAD3_TMA=1 timeout 300 python tools/layer_sweep.py resnet50 64 > gpurun_out/c1_sweep_wgrad3.txt 2>&1
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo done
