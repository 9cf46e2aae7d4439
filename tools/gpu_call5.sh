#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 200 python tools/stg_debug.py > gpurun_out/c5_debug.log 2>&1
B2C_STG_DBG=1 timeout 200 python tools/stg_debug.py > gpurun_out/c5_debug_nostore.log 2>&1
timeout 200 python tools/stg_debug.py 2 32 12 12 32 3 1 >> gpurun_out/c5_debug.log 2>&1
timeout 200 python tools/stg_debug.py 5 32 10 14 48 5 2 >> gpurun_out/c5_debug.log 2>&1
timeout 200 python tools/stg_debug.py 4 128 28 28 128 3 1 >> gpurun_out/c5_debug.log 2>&1
if grep -q "raised\|timeouts recorded: [1-9-]" gpurun_out/c5_debug.log; then
  timeout 300 compute-sanitizer --tool memcheck python tools/stg_debug.py 2 32 12 12 32 3 1 > gpurun_out/c5_sanitizer.log 2>&1
  echo "debug failed"; exit 0
fi
timeout 300 python tools/stg_diag.py > gpurun_out/c5_diag.log 2>&1; echo "rc=$?" >> gpurun_out/c5_diag.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv or full_size or golden" > gpurun_out/c5_parity.log 2>&1; echo "rc=$?" >> gpurun_out/c5_parity.log
for L in "128 28 128 3 1 1 64 fwd" "64 56 256 1 1 0 64 fwd" "256 56 64 1 1 0 64 fwd" "256 14 256 3 1 1 64 fwd" "64 56 64 3 1 1 64 fwd" "256 14 1024 1 1 0 64 dgrad"; do
  echo "== $L" >> gpurun_out/c5_prof.log
  B2C_PROF=1 timeout 120 python tools/one_layer.py $L >> gpurun_out/c5_prof.log 2>&1
done
timeout 600 python tools/layer_sweep.py resnet50 64 > gpurun_out/c5_sweep.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_stg -s 2 -c 1 -o gpurun_out/c5_stg_3x3 python tools/one_layer.py 128 28 128 3 1 1 64 fwd > gpurun_out/c5_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_stg -s 2 -c 1 -o gpurun_out/c5_stg_1x1 python tools/one_layer.py 64 56 256 1 1 0 64 fwd >> gpurun_out/c5_ncu.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c5_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c5_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
echo done
