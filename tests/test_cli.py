"""tools/caffe.py (the `caffe train/time/device_query` shell): argument handling and the no-CPU-mode rule."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = [sys.executable, os.path.join(ROOT, "tools", "caffe.py")]


def run(*a, cwd=ROOT):
    return subprocess.run(CLI + list(a), capture_output=True, text=True, cwd=cwd, timeout=120)


def test_train_needs_a_solver():
    r = run("train")
    assert r.returncode != 0 and "Need a solver definition" in r.stderr      # tools/caffe.cpp:  CHECK_GT(FLAGS_solver.size(), 0)


def test_snapshot_and_weights_are_exclusive():
    r = run("train", "--solver=x", "--snapshot=a.solverstate", "--weights=b.caffemodel")
    assert r.returncode != 0 and "but not both" in r.stderr                   # tools/caffe.cpp:166-168


def test_time_needs_a_model():
    r = run("time")
    assert r.returncode != 0 and "Need a model definition" in r.stderr


def test_missing_net_file_is_reported(tmp_path):
    s = tmp_path / "solver.prototxt"
    s.write_text('net: "does/not/exist.prototxt" base_lr: 0.1 lr_policy: "fixed" max_iter: 3')
    r = run("train", "--solver=%s" % s)
    assert r.returncode != 0 and "does not exist" in r.stderr


def test_no_cpu_mode(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from caffe_mpi_b200 import models
    (tmp_path / "net.prototxt").write_text(models.resnet50_prototxt(2))
    s = tmp_path / "solver.prototxt"
    s.write_text('net: "%s" %s' % (tmp_path / "net.prototxt", models.RESNET50_SOLVER))
    r = run("train", "--solver=%s" % s, "--iterations=1")
    assert r.returncode != 0 and "CUDA" in r.stderr                         # fails loudly, no CPU fallback
    assert run("device_query").returncode != 0
