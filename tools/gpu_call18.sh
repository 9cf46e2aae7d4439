#!/bin/bash
# final confirmation of the defaults: bench (with the CPU baseline), smoke, launch list, BN tests
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c18_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c18_smoke.log
timeout 300 python -m pytest tests/test_layers_gpu.py tests/test_trainer_gpu.py -m gpu -q > gpurun_out/c18_tests_bn.log 2>&1; echo "rc=$?" >> gpurun_out/c18_tests_bn.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/c18_fullnet_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c18_ncu_bench.log 2>&1
echo done
