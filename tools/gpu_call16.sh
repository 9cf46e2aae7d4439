#!/bin/bash
# BatchNorm forward: cp.async prefetch of the slice (on by default) vs register loads; per-direction launch sizes; final checks
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_trainer_gpu.py -m gpu -q > gpurun_out/c16_tests_bn.log 2>&1; echo "rc=$?" >> gpurun_out/c16_tests_bn.log
for v in 1 0; do echo "== B2C_BN_PREFETCH=$v" >> gpurun_out/c16_bn.log; B2C_BN_PREFETCH=$v timeout 200 python tools/bn_sweep.py >> gpurun_out/c16_bn.log 2>&1; done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err
B2C_BN_PREFETCH=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c16_bench_noprefetch.json 2> gpurun_out/c16_bench_noprefetch.err
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c16_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c16_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c16_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c16_smoke.log
echo done
