"""DataLayer on the device: TrainNet reading an LMDB through parser threads -> pinned batch -> H2D on the
copy stream -> b2c_transform_u8 -> top blobs, checked against what the reference's pipeline would deliver for the same database:
the records CursorManager hands this solver (tests/test_data_cpu.py's line-by-line restatement of data_reader.cpp:206-310), the
crop / mirror draws of a DataTransformer seeded with transform_param.random_seed (mt19937, data_transformer.cpp:127-137,729-749)
and DataTransformer::Transform's arithmetic (oracle/layers_oracle.py::transform_u8, data_transformer.cpp:178-312).  Bit-exact.

(The file name sorts last on purpose: it was written after the round's GPU budget was spent, so its first hardware run is the
driver's; everything it composes -- the kernel, the copy-stream slots, the reader -- is covered by earlier tests.)"""
import numpy as np
import pytest

from caffe_mpi_b200 import data_api, host_api, lmdb_io
from oracle import layers_oracle as lo
from test_data_cpu import oracle_batches

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180, method="thread")]     # a protocol bug must end the run, not hang the box (the file sorts last)
SOLVER = 'base_lr: 0.01 lr_policy: "fixed" momentum: 0.9 weight_decay: 0.0005 max_iter: 100 solver_mode: GPU'
# crops and datum sizes below keep InnerProduct's K odd, i.e. on the exact-fp32 GEMM: this file is about the data layer
NET = ('name: "db_net"\n'
       'layer {{ name: "data" type: "Data" top: "data" top: "label" include {{ phase: TRAIN }}\n'
       '  data_param {{ source: "{src}" backend: LMDB batch_size: {B} {dp} }}\n'
       '  transform_param {{ {tp} }} }}\n'
       'layer {{ name: "ip" type: "InnerProduct" bottom: "data" top: "ip" inner_product_param {{ num_output: 10 weight_filler {{ type: "gaussian" std: 0.01 }} }} }}\n'
       'layer {{ name: "loss" type: "SoftmaxWithLoss" bottom: "ip" bottom: "label" top: "loss" }}\n')


def _db(tmp_path, n, c, h, w):
    rng = np.random.default_rng(100 + n)
    imgs = rng.integers(0, 256, (n, c, h, w), dtype=np.uint8)
    labels = rng.integers(0, 10, n)
    path = str(tmp_path / "train_lmdb")
    lmdb_io.write_datum_lmdb(path, imgs, labels)
    return path, imgs, labels


def _check_steps(t, imgs, labels, B, steps, crop, seed, mirror, mean_values=None, mean_image=None, scale=1.0, P=1):
    n, C, H, W = imgs.shape
    ch, cw = (crop, crop) if crop else (H, W)
    want_batches = oracle_batches(n, steps, B, 1, 0, P)
    ho, wo, mir = data_api.transform_draws(seed, mirror, crop, True, B * steps, H, W)
    for i in range(steps):
        t.step(1, copy_input=True)
        pos = [p for p, _ in want_batches[i]]
        want = lo.transform_u8(imgs[pos], (ch, cw), ho[B * i:B * i + B], wo[B * i:B * i + B], mir[B * i:B * i + B], mean_values, mean_image, scale)
        got = t.get_blob("data").reshape(B, C, ch, cw)
        assert np.array_equal(got, want), f"batch {i}"
        assert np.array_equal(t.get_blob("label"), labels[pos].astype(np.float32)), f"labels of batch {i}"
        assert t.database_batches() == i + 1
        assert np.isfinite(t.loss())


def test_data_layer_delivers_the_reference_batches(tmp_path, monkeypatch):
    monkeypatch.delenv("B2C_DATA", raising=False)
    path, imgs, labels = _db(tmp_path, 50, 3, 12, 10)
    B = 4
    net = NET.format(src=path, B=B, dp="", tp="crop_size: 7 mirror: true mean_value: 104 mean_value: 117 mean_value: 123 scale: 0.5 random_seed: 77")
    t = host_api.Trainer(net, SOLVER, num_classes=10)
    assert t.database_batches() == 0
    # 15 batches of 4 out of 50 records: batch 12 wraps around the end of the database
    _check_steps(t, imgs, labels, B, 15, 7, 77, True, mean_values=[104.0, 117.0, 123.0], scale=0.5)
    # a step without a host copy keeps the resident batch (bench.py's device-only timing)
    before = t.get_blob("data").copy()
    t.step(1, copy_input=False)
    assert t.database_batches() == 15 and np.array_equal(t.get_blob("data"), before)


def test_data_layer_mean_file_no_crop_two_parser_threads(tmp_path, monkeypatch):
    monkeypatch.delenv("B2C_DATA", raising=False)
    path, imgs, labels = _db(tmp_path, 23, 1, 9, 11)
    mean = np.random.default_rng(4).uniform(90, 130, (1, 1, 9, 11)).astype(np.float32)
    mean_path = str(tmp_path / "mean.binaryproto")
    data_api.blobproto_save(mean_path, mean)
    B = 3
    net = NET.format(src=path, B=B, dp="parser_threads: 2", tp=f'mean_file: "{mean_path}" scale: 0.00390625 mirror: true random_seed: 5')
    t = host_api.Trainer(net, SOLVER, num_classes=10, default_channels=1)
    # two parser threads: thread t assembles batches t, t + 2, ... from its own stripe of the database (P * B records per cycle)
    _check_steps(t, imgs, labels, B, 9, 0, 5, True, mean_image=mean[0], scale=0.00390625, P=2)


def test_first_forward_loads_a_batch_without_being_asked(tmp_path, monkeypatch):
    monkeypatch.delenv("B2C_DATA", raising=False)
    path, imgs, labels = _db(tmp_path, 16, 3, 7, 9)
    net = NET.format(src=path, B=4, dp="", tp="random_seed: 1")
    t = host_api.Trainer(net, SOLVER, num_classes=10)
    loss = t.forward_backward()                         # Forward(copy_input = false) on a net that has never seen a batch
    assert np.isfinite(loss) and t.database_batches() == 1
    assert np.array_equal(t.get_blob("data").reshape(4, 3, 7, 9), imgs[:4].astype(np.float32))
    assert np.array_equal(t.get_blob("label"), labels[:4].astype(np.float32))
