#!/bin/bash
# restructured staged weight gradient + plane mode of the staged forward kernel: debug runs first, then parity, sweep, tests, bench, ncu
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python tools/wstg_debug.py > gpurun_out/c8_wdebug.log 2>&1; echo "rc=$?" >> gpurun_out/c8_wdebug.log
if grep -q "raised\|timeouts recorded\|nan=True" gpurun_out/c8_wdebug.log; then
  timeout 300 compute-sanitizer --tool memcheck python tools/wstg_debug.py 2 32 12 12 40 3 1 > gpurun_out/c8_wsanitizer.log 2>&1
  export B2C_WGRAD_STAGED=0
  echo "wgrad debug failed: staged wgrad off for the rest" >> gpurun_out/c8_wdebug.log
fi
# plane mode (7x7 maps)
export B2C_CONV_STAGED_PLANE=1
for shape in "3 512 7 7 512 3 1" "11 96 7 7 160 1 0" "6 32 7 7 128 5 2" "64 2048 7 7 512 1 0"; do
  timeout 120 python tools/stg_debug.py $shape >> gpurun_out/c8_pdebug.log 2>&1; echo "rc=$?" >> gpurun_out/c8_pdebug.log
done
if grep -q "raised\|timeouts recorded: [1-9]" gpurun_out/c8_pdebug.log; then
  timeout 300 compute-sanitizer --tool memcheck python tools/stg_debug.py 3 512 7 7 512 3 1 > gpurun_out/c8_psanitizer.log 2>&1
  export B2C_CONV_STAGED_PLANE=0
  echo "plane debug failed: plane mode off for the rest" >> gpurun_out/c8_pdebug.log
fi
for v in "1 2" "1 1" "1 3" "1 4" "0 2"; do set -- $v; echo "== B2C_BN_ONEPASS=$1 B2C_BN_OCC=$2" >> gpurun_out/c8_bn.log; B2C_BN_ONEPASS=$1 B2C_BN_OCC=$2 timeout 200 python tools/bn_sweep.py >> gpurun_out/c8_bn.log 2>&1; done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c8_parity.log 2>&1; echo "rc=$?" >> gpurun_out/c8_parity.log
timeout 600 python tools/layer_sweep.py resnet50 64 > gpurun_out/c8_sweep.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c8_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_stg -s 2 -c 1 -o gpurun_out/c8_wstg_3x3 python tools/one_layer.py 128 28 128 3 1 1 64 wgrad > gpurun_out/c8_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_stg -s 2 -c 1 -o gpurun_out/c8_wstg_1x1 python tools/one_layer.py 256 14 1024 1 1 0 64 wgrad >> gpurun_out/c8_ncu.log 2>&1
echo done
