#!/bin/bash
# smem-parked one-launch BatchNorm, residual-tail fusion, quad max-pool backward, plane mode default: tests, BN sweep, A/B benches, launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c9_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c9_tests.log
for v in "1 110" "1 0" "0 110"; do set -- $v; echo "== B2C_BN_ONEPASS=$1 B2C_BN_CACHE_KB=$2" >> gpurun_out/c9_bn.log; B2C_BN_ONEPASS=$1 B2C_BN_CACHE_KB=$2 timeout 200 python tools/bn_sweep.py >> gpurun_out/c9_bn.log 2>&1; done
timeout 600 python tools/layer_sweep.py resnet50 64 > gpurun_out/c9_sweep.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
B2C_FUSE_RES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_nores.json 2> gpurun_out/c9_bench_nores.err
B2C_FUSE_RES=0 B2C_BN_ONEPASS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_bn2launch.json 2> gpurun_out/c9_bench_bn2launch.err
for mdl in googlenet vgg16 alexnet; do timeout 600 python bench.py --model $mdl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_$mdl.json 2> gpurun_out/c9_bench_$mdl.err; done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/c9_fullnet_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c9_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_fwd_onepass -s 12 -c 1 -o gpurun_out/c9_bn_fwd python tools/bn_sweep.py > gpurun_out/c9_ncu_bn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_stg -s 2 -c 1 -o gpurun_out/c9_wstg_1x1 python tools/one_layer.py 256 14 1024 1 1 0 64 wgrad >> gpurun_out/c9_ncu_bn.log 2>&1
echo done
