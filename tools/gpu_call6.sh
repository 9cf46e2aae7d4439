#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c6_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c6_tests.log
timeout 600 python tools/layer_sweep.py resnet50 64 > gpurun_out/c6_sweep.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
B2C_FUSE=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c6_bench_nofuse.json 2> gpurun_out/c6_bench_nofuse.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c6_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c6_ncu_bench.log 2>&1
for M in alexnet vgg16 googlenet lenet; do
  timeout 600 python bench.py --model $M --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c6_bench_$M.json 2> gpurun_out/c6_bench_$M.err
done
echo done
