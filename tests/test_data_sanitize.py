"""AddressSanitizer + UBSan run of the input pipeline's host code (tests/sim/data_stress.cpp): the db::LMDB writer against a
std::map model over random commits, the reader on hundreds of damaged copies of a valid database (byte flips in page headers, node
tables and meta pages; truncations) -- each must end in caffe::FatalError or a clean walk, never in a crash -- ParseDatum on random
and truncated bytes, DataReader objects destroyed with batches in flight, and the two parsers of files that come from outside:
prototxt text through the Net graph builder and the solver reader (1 500 mutated nets per run: it found a Pooling layer fed a 1-D
blob indexing past its shape, zero strides dividing by zero and a scalar `shape:` dereferenced as a message -- all now fatal
checks), and the .caffemodel / .solverstate / BlobProto wire format on flipped and truncated bytes."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("seed", [7])
def test_data_pipeline_host_code_is_clean_under_sanitizers(tmp_path, seed):
    from caffe_mpi_b200 import capi
    capi.lib()                                               # libb2c.so must be built (the driver links it for the host layer's other symbols)
    r = subprocess.run(["make", "-C", os.path.join(HERE, "sim"), "data_stress"], capture_output=True, text=True)
    if r.returncode != 0:
        if "sanitize" in r.stderr or "asan" in r.stderr.lower():
            pytest.skip("toolchain has no sanitizer runtime: " + r.stderr[-300:])
        pytest.fail("tests/sim/data_stress does not build:\n" + r.stdout[-1000:] + r.stderr[-3000:])
    try:                                                     # seed files for the JPEG-decoder section (skipped there when cv2 is absent)
        import cv2
        import numpy as np
        rng = np.random.default_rng(seed)
        for k, (h, w, extra) in enumerate([(40, 56, []), (33, 17, [cv2.IMWRITE_JPEG_RST_INTERVAL, 2]), (24, 24, [cv2.IMWRITE_JPEG_OPTIMIZE, 1])]):
            img = cv2.resize(rng.integers(0, 256, (6, 6, 3), dtype=np.uint8), (w, h), interpolation=cv2.INTER_CUBIC)
            ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 70] + extra)
            (tmp_path / ("seed%d.jpg" % k)).write_bytes(enc.tobytes())
        ok, enc = cv2.imencode(".jpg", img[:, :, 0])
        (tmp_path / "seed3.jpg").write_bytes(enc.tobytes())
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_QUALITY, 60])
        (tmp_path / "seed4.jpg").write_bytes(enc.tobytes())              # progressive: refinement scans on damaged input
        for k, im in ((5, img), (6, img[:, :, 0])):                       # PNG: inflate + unfilter on damaged input (CRC / Adler mostly refuse)
            ok, enc = cv2.imencode(".png", im, [cv2.IMWRITE_PNG_COMPRESSION, 3 * (k - 5)])
            (tmp_path / ("seed%d.png" % k)).write_bytes(enc.tobytes())
    except ImportError:
        pass
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    env.pop("LD_PRELOAD", None)
    r = subprocess.run([os.path.join(HERE, "sim", "data_stress"), str(tmp_path), str(seed)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "data_stress ok" in r.stdout, r.stdout[-1000:] + r.stderr[-4000:]
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-4000:]


def test_reader_threads_are_clean_under_thread_sanitizer(tmp_path):
    """DataReader's parser threads against ThreadSanitizer: hand-over of batches through the per-thread queues, the first-error
    channel, destruction with batches in flight (tests/sim/data_stress.cpp, section 4 alone)."""
    from caffe_mpi_b200 import capi
    capi.lib()
    r = subprocess.run(["make", "-C", os.path.join(HERE, "sim"), "data_stress_tsan"], capture_output=True, text=True)
    if r.returncode != 0:
        if "sanitize" in r.stderr or "tsan" in r.stderr.lower():
            pytest.skip("toolchain has no ThreadSanitizer runtime: " + r.stderr[-300:])
        pytest.fail("tests/sim/data_stress_tsan does not build:\n" + r.stdout[-1000:] + r.stderr[-3000:])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0:second_deadlock_stack=1")
    env.pop("LD_PRELOAD", None)
    for seed in ("3",):
        r = subprocess.run([os.path.join(HERE, "sim", "data_stress_tsan"), str(tmp_path / seed), seed, "threads"], capture_output=True, text=True, env=env,
                           timeout=600)
        if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
            pytest.skip("ThreadSanitizer cannot map its shadow in this container: " + r.stderr[-200:])
        assert r.returncode == 0 and "data_stress ok" in r.stdout, r.stdout[-1000:] + r.stderr[-4000:]
        assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-6000:]
