#!/bin/bash
# BatchNorm: 16-CTA clusters (50 KB slices, both backward streams parked) vs 8, gradient-stream parking; plane wgrad default; final single-GPU artefacts
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c11_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c11_tests.log
B2C_BN_CLUSTER=8 timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_trainer_gpu.py -m gpu -q > gpurun_out/c11_tests_c8.log 2>&1; echo "rc=$?" >> gpurun_out/c11_tests_c8.log
for v in 16 8; do echo "== B2C_BN_CLUSTER=$v" >> gpurun_out/c11_bn.log; B2C_BN_CLUSTER=$v timeout 200 python tools/bn_sweep.py >> gpurun_out/c11_bn.log 2>&1; done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
B2C_BN_CLUSTER=8 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench_c8.json 2> gpurun_out/c11_bench_c8.err
timeout 600 python tools/layer_sweep.py resnet50 64 > gpurun_out/c11_sweep.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c11_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c11_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c11_bench_reference.json 2> gpurun_out/c11_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/c11_fullnet_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c11_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_onepass -s 16 -c 1 -o gpurun_out/c11_bn_bwd python tools/bn_sweep.py > gpurun_out/c11_ncu_bn.log 2>&1
echo done
