#!/bin/bash
# BatchNorm launches with 128-thread CTAs on tiny channels (7x7 maps): BN tests first, then the full suite, sweep, bench, launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_trainer_gpu.py -m gpu -q > gpurun_out/c14_tests_bn.log 2>&1; echo "rc=$?" >> gpurun_out/c14_tests_bn.log
timeout 200 python tools/bn_sweep.py > gpurun_out/c14_bn.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c14_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c14_tests.log
timeout 600 python bench.py --model googlenet --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c14_bench_googlenet.json 2> gpurun_out/c14_bench_googlenet.err
timeout 600 python bench.py --model alexnet --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c14_bench_alexnet.json 2> gpurun_out/c14_bench_alexnet.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c14_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c14_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/c14_fullnet_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c14_ncu_bench.log 2>&1
echo done
