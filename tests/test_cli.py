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


def test_train_schedule_follows_display_and_snapshot_intervals():
    """solver.cpp:289-345: a loss line every `display` iterations, a snapshot whenever iter % snapshot == 0 after a step; the chunks
    between those points run without the host looking at the device."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("caffe_cli", os.path.join(ROOT, "tools", "caffe.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    plan = list(cli.schedule(0, 25, 10, 0))
    assert plan == [(10, 10, True, False), (10, 20, True, False), (5, 25, True, False)]
    plan = list(cli.schedule(0, 12, 5, 4))
    assert plan == [(4, 4, False, True), (1, 5, True, False), (3, 8, False, True), (2, 10, True, False), (2, 12, True, True)]
    assert list(cli.schedule(7, 20, 0, 0)) == [(13, 20, True, False)]                      # no display: one chunk, one closing line
    assert list(cli.schedule(18, 31, 10, 10)) == [(2, 20, True, True), (10, 30, True, True), (1, 31, True, False)]   # a resumed run
    assert list(cli.schedule(5, 5, 10, 10)) == []
    assert sum(n for n, *_ in cli.schedule(3, 1003, 7, 11)) == 1000


def test_textproto_scalar_reads_solver_fields(tmp_path):
    from caffe_mpi_b200 import host_api
    s = tmp_path / "solver.prototxt"
    s.write_text('net: "a/b.prototxt"\ndisplay: 100  # comment\nsnapshot: 2500\nsnapshot_prefix: "models/x/snap"\nsnapshot_after_train: false\nrandom_seed: 1\n'
                 'stepvalue: 10 stepvalue: 20 train_state { level: 3 }')
    g = lambda k, d=None: host_api.textproto_scalar(str(s), k, d)
    assert (g("display"), g("snapshot"), g("snapshot_prefix"), g("snapshot_after_train"), g("random_seed")) == ("100", "2500", "models/x/snap", "false", "1")
    assert g("stepvalue") == "10" and g("test_interval", "0") == "0" and g("train_state") is None
    assert host_api.textproto_scalar('display: 7', "display", is_text=True) == "7"
    with pytest.raises(host_api.HostError):
        host_api.textproto_scalar("display: {", "display", is_text=True)


def test_rank_batch_divides_the_prototxt_batch_like_p2psync(tmp_path):
    """parallel.cpp:284-316: the Data layer's batch_size is per node; each of its solvers takes 1/solver_count, rounded up."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("caffe_cli2", os.path.join(ROOT, "tools", "caffe.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    net = ('layer { name: "d" type: "Data" top: "data" top: "label" data_param { source: "nowhere" backend: LMDB batch_size: 50 } transform_param { crop_size: 8 } }\n'
           'layer { name: "ip" type: "InnerProduct" bottom: "data" top: "ip" inner_product_param { num_output: 4 } }\n'
           'layer { name: "loss" type: "SoftmaxWithLoss" bottom: "ip" bottom: "label" top: "loss" }\n')
    assert cli.rank_batch(net, True, 0, 1) == 0                 # one solver: the prototxt's own batch
    assert cli.rank_batch(net, True, 0, 2) == 25
    assert cli.rank_batch(net, True, 0, 8) == 7                 # 50 -> 56 -> 7 per solver
    assert cli.rank_batch(net, True, 512, 8) == 64              # --batch is the node's batch as well
    assert cli.rank_batch(net, True, 96, 1) == 96
    inp = 'layer { name: "in" type: "Input" top: "data" input_param { shape { dim: 6 dim: 3 dim: 8 dim: 8 } } }\n'
    assert cli.rank_batch(inp, True, 0, 4) == 0                 # no data_param: left alone


def test_time_table_has_the_reference_layout():
    import importlib.util
    spec = importlib.util.spec_from_file_location("caffe_cli3", os.path.join(ROOT, "tools", "caffe.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    layers = [("data", "Data"), ("conv1", "Convolution"), ("loss", "SoftmaxWithLoss")]
    prof = [(0, "fwd", 0.01), (1, "fwd", 0.5), (1, "bwd", 1.25), (1, "wgrad", 0.75), (1, "dgrad", 0.5), (2, "fwd", 0.02), (2, "bwd", 0.03)]
    lines, f, b = cli.layer_time_lines(layers, prof)
    assert lines[0] == "Average time per layer: " and len(lines) == 1 + 2 * len(layers)
    assert lines[3] == "     conv1\tforward: 0.5 ms." and lines[4] == "     conv1\tbackward: 1.25 ms. (dgrad 0.5) (wgrad 0.75)"
    assert lines[2] == "      data\tbackward: 0 ms."
    assert abs(f - 0.53) < 1e-12 and abs(b - 1.28) < 1e-12


def test_compute_image_mean_writes_what_the_data_layer_reads(tmp_path):
    import numpy as np
    from caffe_mpi_b200 import data_api, lmdb_io
    rng = np.random.default_rng(2)
    imgs = rng.integers(0, 256, (37, 3, 6, 5), dtype=np.uint8)
    db = str(tmp_path / "db")
    lmdb_io.write_datum_lmdb(db, imgs, rng.integers(0, 9, 37))
    out = str(tmp_path / "mean.binaryproto")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "compute_image_mean.py"), db, out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Number of channels: 3" in r.stdout and "Processed 37 files." in r.stdout, r.stdout + r.stderr
    got = data_api.blobproto_load(out)                               # the parser the Data layer's mean_file goes through
    want = np.zeros((3, 6, 5), np.float32)
    for im in imgs:                                                   # compute_image_mean.cpp:75-99: float running sum, then / count
        want += im.astype(np.float32)
    want /= np.float32(37)
    assert got.shape == (1, 3, 6, 5) and np.array_equal(got[0], want)
    assert ("mean_value channel [0]: %g" % float(want[0].sum(dtype=np.float32) / 30)) in r.stdout


def test_convert_imageset_raw_and_encoded(tmp_path):
    """tools/convert_imageset.py -> db::LMDB writer -> DataReader: raw datums hold cv2's B,G,R planes (resized), --encoded datums the
    re-encoded or original files, which the parser threads decode back to the same pixels cv2.imdecode gives."""
    import numpy as np
    cv2 = pytest.importorskip("cv2")
    from caffe_mpi_b200 import data_api
    rng = np.random.default_rng(6)
    root = tmp_path / "imgs"
    (root / "a").mkdir(parents=True)
    names = []
    for i in range(7):
        img = cv2.resize(rng.integers(0, 256, (5, 6, 3), dtype=np.uint8), (40 + i, 30 + i), interpolation=cv2.INTER_CUBIC)
        name = "a/im%d.jpg" % i
        assert cv2.imwrite(str(root / name), img, [cv2.IMWRITE_JPEG_QUALITY, 90])
        names.append(name)
    lst = tmp_path / "list.txt"
    lst.write_text("".join("%s %d\n" % (n, i % 3) for i, n in enumerate(names)) + "a/missing.jpg 1\n")
    tool = [sys.executable, os.path.join(ROOT, "tools", "convert_imageset.py")]
    raw_db = str(tmp_path / "raw_db")
    r = subprocess.run(tool + ["--resize_width=24", "--resize_height=20", "--check_size", str(root) + "/", str(lst), raw_db], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Processed 7 files." in r.stdout and "Could not open or find file" in r.stderr, r.stdout + r.stderr
    items = data_api.LMDB(raw_db).items()
    assert [k for k, _ in items] == [("%08d_%s" % (i, n)).encode() for i, n in enumerate(names)]
    rd = data_api.DataReader(raw_db, 7)
    data, label, _, _ = rd.next()
    rd.close()
    want = np.stack([cv2.resize(cv2.imread(str(root / n)), (24, 20)).transpose(2, 0, 1) for n in names])
    assert np.array_equal(data, want) and label.tolist() == [i % 3 for i in range(7)]
    enc_db = str(tmp_path / "enc_db")
    r = subprocess.run(tool + ["--encoded", "--resize_width=24", "--resize_height=20", str(root) + "/", str(lst), enc_db], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    rd = data_api.DataReader(enc_db, 7, parser_threads=2)
    data, label, _, _ = rd.next()
    rd.close()
    for i, n in enumerate(names):                                 # what the reference would train on: imdecode of the stored bytes
        d = data_api.datum_parse(data_api.LMDB(enc_db).items()[i][1])
        assert d["encoded"] and d["channels"] == 0
        assert np.array_equal(data[i], cv2.imdecode(np.frombuffer(d["data"], np.uint8), cv2.IMREAD_UNCHANGED).transpose(2, 0, 1))
    # no resize + matching extension: the original file is stored untouched (ReadFileToDatum)
    keep_db = str(tmp_path / "keep_db")
    lst.write_text("%s 2\n" % names[0])
    r = subprocess.run(tool + ["--encoded", str(root) + "/", str(lst), keep_db], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert data_api.datum_parse(data_api.LMDB(keep_db).items()[0][1])["data"] == (root / names[0]).read_bytes()
