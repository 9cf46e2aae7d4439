#!/bin/bash
# plane-mode staged wgrad (opt-in), compacted strided wgrad on the staged kernel, split-K InnerProduct forward, pool backward v2, deferred fan-out add
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B2C_WGRAD_STAGED_PLANE=1 timeout 300 python tools/wstg_debug.py > gpurun_out/c10_wdebug.log 2>&1; echo "rc=$?" >> gpurun_out/c10_wdebug.log
PLANE=1
if grep -q "raised\|timeouts recorded\|nan=True" gpurun_out/c10_wdebug.log; then
  B2C_WGRAD_STAGED_PLANE=1 timeout 300 compute-sanitizer --tool memcheck python tools/wstg_debug.py 9 64 7 7 72 3 1 > gpurun_out/c10_wsanitizer.log 2>&1
  PLANE=0
fi
timeout 300 python tools/wstg_debug.py >> gpurun_out/c10_wdebug.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c10_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c10_tests.log
if [ $PLANE = 1 ]; then
  B2C_WGRAD_STAGED_PLANE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_trainer_gpu.py -m gpu -q > gpurun_out/c10_tests_plane.log 2>&1; echo "rc=$?" >> gpurun_out/c10_tests_plane.log
fi
timeout 600 python tools/layer_sweep.py resnet50 64 > gpurun_out/c10_sweep.txt 2>&1
B2C_WGRAD_STAGED_PLANE=$PLANE timeout 600 python tools/layer_sweep.py resnet50 64 > gpurun_out/c10_sweep_plane.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
B2C_WGRAD_STAGED_PLANE=$PLANE timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench_plane.json 2> gpurun_out/c10_bench_plane.err
B2C_FUSE_SPLIT=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench_nosplit.json 2> gpurun_out/c10_bench_nosplit.err
for mdl in googlenet vgg16 alexnet lenet; do B2C_WGRAD_STAGED_PLANE=$PLANE timeout 600 python bench.py --model $mdl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench_$mdl.json 2> gpurun_out/c10_bench_$mdl.err; done
B2C_WGRAD_STAGED_PLANE=$PLANE timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/c10_fullnet_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c10_ncu_bench.log 2>&1
echo done
