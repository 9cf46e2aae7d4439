#!/bin/bash
# prefetching data layer + InnerProduct dW threshold: trainer / data-path tests, e2e bench, AlexNet
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_layers_extra_gpu.py tests/test_host_gpu.py -m gpu -q > gpurun_out/c13_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c13_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
timeout 600 python bench.py --model alexnet --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench_alexnet.json 2> gpurun_out/c13_bench_alexnet.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c13_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c13_smoke.log
echo done
