#!/bin/bash
# final single-GPU pass of the round: full tests, a subset of the non-default switch variants, benches of every BASELINE net, launch list, smoke
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c12_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c12_tests.log
B2C_RUN_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_experimental_gpu.py -m gpu -q -k "BN_ONEPASS or FUSE_RES or FUSE_SPLIT or BN_CACHE or WGRAD_STAGED_PLANE or CONV_STAGED_PLANE" > gpurun_out/c12_tests_switches.log 2>&1; echo "rc=$?" >> gpurun_out/c12_tests_switches.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
for mdl in googlenet vgg16 alexnet lenet; do timeout 600 python bench.py --model $mdl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c12_bench_$mdl.json 2> gpurun_out/c12_bench_$mdl.err; done
timeout 600 python tools/layer_sweep.py resnet50 64 > gpurun_out/c12_sweep.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c12_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c12_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/c12_fullnet_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c12_ncu_bench.log 2>&1
timeout 200 python tools/bn_sweep.py > gpurun_out/c12_bn.log 2>&1
echo done
