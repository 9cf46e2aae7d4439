#!/bin/bash
# First hardware session after the round that ended without GPU budget (see profiles/README.md, last section).  One GPU, ~10 minutes:
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/gpu_call_next.sh'
# 1. the GPU tests written without hardware (DataLayer on an LMDB), then the whole GPU suite;
# 2. the benchmark on the synthetic source (the round's headline) and with an LMDB behind the Data layer: the difference in `e2e` is
#    what the parser threads and the page-cache read cost; raw datums first, then JPEG-encoded ones (parser threads decode);
# 3. a launch list of the LMDB-backed step (the transform kernel's share of the step must stay where it was).
set -x
mkdir -p gpurun_out
python -m pytest tests/test_zz_data_layer_gpu.py -x -q > gpurun_out/next_tests_data_layer.log 2>&1; echo "rc=$?" >> gpurun_out/next_tests_data_layer.log
python -m pytest tests -x -q -m gpu > gpurun_out/next_tests.log 2>&1; echo "rc=$?" >> gpurun_out/next_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/next_bench_synthetic.json 2> gpurun_out/next_bench_synthetic.err
python tools/make_lmdb.py /tmp/raw_db --random 4096 --shape 3x256x256 > gpurun_out/next_make_lmdb.log 2>&1
python bench.py --steps 20 --warmup 5 --lmdb /tmp/raw_db --no-cpu-baseline > gpurun_out/next_bench_lmdb_raw.json 2> gpurun_out/next_bench_lmdb_raw.err
python - <<'PY' > gpurun_out/next_make_encoded.log 2>&1
import os, sys, numpy as np, cv2
sys.path.insert(0, os.getcwd())
from caffe_mpi_b200 import data_api, lmdb_io
rng = np.random.default_rng(0)
env = data_api.LMDB("/tmp/enc_db", "NEW")
for i in range(2048):
    img = cv2.resize(rng.integers(0, 256, (32, 32, 3), dtype=np.uint8), (256, 256), interpolation=cv2.INTER_CUBIC)
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
    env.put(lmdb_io.caffe_key(i, "img%d.jpg" % i), data_api.datum_serialize(0, 0, 0, enc.tobytes(), int(rng.integers(0, 1000)), encoded=True))
    if i % 1000 == 999:
        env.commit()
env.commit()
print(env.stat())
PY
python bench.py --steps 20 --warmup 5 --lmdb /tmp/enc_db --no-cpu-baseline > gpurun_out/next_bench_lmdb_jpeg.json 2> gpurun_out/next_bench_lmdb_jpeg.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/next_lmdb_launches.csv python bench.py --steps 1 --warmup 1 --lmdb /tmp/raw_db --no-cpu-baseline > gpurun_out/next_ncu.log 2>&1
tail -3 gpurun_out/next_tests_data_layer.log gpurun_out/next_tests.log
cat gpurun_out/next_bench_synthetic.json gpurun_out/next_bench_lmdb_raw.json gpurun_out/next_bench_lmdb_jpeg.json
