"""Whole-net parity on the GPU: TrainNet (prototxt -> layers -> ForwardBackward -> SGD) against tests/netoracle.py.
Tolerance: 1e-3 relative (max|a-ref| / max|ref| per tensor), the north_star's bar for fp32 mode."""
import numpy as np
import pytest

from caffe_mpi_b200 import capi, host_api, models
import netoracle as no

pytestmark = pytest.mark.gpu
TOL = 1e-3
SOLVER = 'base_lr: 0.05 lr_policy: "fixed" momentum: 0.9 weight_decay: 0.0005 max_iter: 100 solver_mode: GPU'


def rel(a, ref, floor=1e-20):
    return float(np.max(np.abs(a.reshape(-1) - ref.reshape(-1))) / max(np.max(np.abs(ref)), floor))


def make_trainer(spec, rng, classes=10, math=capi.MATH_FP32, fuse=None):
    """fuse: None = library default (fusion pass on), False = B2C_FUSE=0 (the prototxt's layers one by one)."""
    import os
    shapes = no.param_shapes(spec)
    old = os.environ.get("B2C_FUSE")
    if fuse is not None:
        os.environ["B2C_FUSE"] = "1" if fuse else "0"
    try:
        t = host_api.Trainer(no.to_prototxt(spec), SOLVER, num_classes=classes, math=math)
    finally:
        if fuse is not None:
            if old is None:
                os.environ.pop("B2C_FUSE", None)
            else:
                os.environ["B2C_FUSE"] = old
    assert t.num_params() == len(shapes)
    params = []
    for i, (layer, kind, shp) in enumerate(shapes):
        if kind == "w":
            p = rng.standard_normal(shp).astype(np.float32) * np.float32(np.sqrt(2.0 / np.prod(shp[1:])))
        elif kind == "scale":
            p = rng.uniform(0.5, 1.5, shp).astype(np.float32)
        else:
            p = rng.uniform(-0.2, 0.2, shp).astype(np.float32)
        assert t.get_param(i).size == p.size
        t.set_param(i, p)
        params.append(p)
    dshape = spec[0]["shape"]
    data = rng.standard_normal(dshape).astype(np.float32)
    label = rng.integers(0, classes, dshape[0]).astype(np.float32)
    t.set_blob("data", data)
    t.set_blob("label", label)
    return t, params, data, label


@pytest.mark.parametrize("fuse", [False, True], ids=["unfused", "fused"])
@pytest.mark.parametrize("conv_bias", [False, True])
def test_forward_backward_matches_oracle(rng, conv_bias, fuse):
    spec = no.mini_resnet(conv_bias=conv_bias)
    t, params, data, label = make_trainer(spec, rng, fuse=fuse)
    loss = t.forward_backward()
    ref_loss, grads, v, d = no.forward_backward(spec, params, data, label)
    assert abs(loss - ref_loss) <= TOL * abs(ref_loss)
    for name in ("conv1", "pool1", "resA.1.sum", "resA.2.sum", "pool2", "fc"):
        assert rel(t.get_blob(name), v[name]) <= TOL, name
    # with the fusion pass on, the diff of a blob that an in-place ReLU follows holds dL/d(post-ReLU): the ReLU mask is applied
    # inside the fused backward kernel instead of in place (the oracle's d[] is the masked one)
    for name in ("fc", "pool2", "resA.1.conv1", "pool1", "conv1") + (() if fuse else ("resA.2.sum",)):
        assert rel(t.get_blob(name, diff=True), d[name]) <= TOL, name
    # gradients that are mathematically zero (a conv bias in front of BatchNorm, whose mean subtraction cancels it) are
    # rounding noise on both sides: the denominator is floored at 1e-3 of the largest gradient in the net
    floor = 1e-3 * max(float(np.max(np.abs(g))) for g in grads)
    for i, g in enumerate(grads):
        assert rel(t.get_param(i, 1), g, floor) <= TOL, no.param_shapes(spec)[i]


def test_lenet_matches_oracle(rng):
    """BASELINE configs[0]: LeNet on MNIST-shaped input, batch 64 -- loss, every gradient, then 2 SGD steps."""
    spec = no.lenet()
    t, params, data, label = make_trainer(spec, rng)
    loss = t.forward_backward()
    ref_loss, grads, v, d = no.forward_backward(spec, params, data, label)
    assert abs(loss - ref_loss) <= TOL * abs(ref_loss)
    for name in ("conv1", "pool1", "conv2", "pool2", "ip1", "ip2"):
        assert rel(t.get_blob(name), v[name]) <= TOL, name
    floor = 1e-3 * max(float(np.max(np.abs(g))) for g in grads)
    for i, g in enumerate(grads):
        assert rel(t.get_param(i, 1), g, floor) <= TOL, no.param_shapes(spec)[i]


def test_sgd_steps_match_oracle(rng):
    spec = no.mini_resnet()
    t, params, data, label = make_trainer(spec, rng)
    ref_losses, ref_params, ref_hist = no.sgd_steps(spec, params, data, label, 3, 0.05, 0.9, 0.0005)
    losses = []
    for _ in range(3):
        t.step(1)
        losses.append(t.loss())
    np.testing.assert_allclose(losses, ref_losses, rtol=2 * TOL)
    assert ref_losses[-1] < ref_losses[0]
    hfloor = 1e-3 * max(float(np.max(np.abs(h))) for h in ref_hist)
    for i, (p, h) in enumerate(zip(ref_params, ref_hist)):
        assert rel(t.get_param(i, 0), p) <= 2 * TOL, i
        assert rel(t.get_param(i, 2), h, hfloor) <= 5 * TOL, i
        assert not t.get_param(i, 1).any()            # diffs cleared by the update (sgd_solver.cu / clear_grads)


def test_tf32_mode_within_its_tolerance(rng):
    spec = no.mini_resnet()
    t, params, data, label = make_trainer(spec, rng, math=capi.MATH_TF32)
    loss = t.forward_backward()
    ref_loss, grads, _, _ = no.forward_backward(spec, params, data, label)
    assert abs(loss - ref_loss) <= 1e-2 * abs(ref_loss)


def test_resnet50_one_step_runs_and_learns(rng):
    """Full ResNet-50 train graph at batch 8, 224x224: loss starts near ln(1000) and goes down on a fixed batch."""
    t = host_api.Trainer(models.resnet50_prototxt(8), 'base_lr: 0.01 lr_policy: "fixed" momentum: 0.9 weight_decay: 1e-4 max_iter: 10', batch=8)
    assert t.num_params() == 161
    t.step(1)
    first = t.loss()
    assert 5.0 < first < 9.5
    t.step(4)
    assert t.loss() < first
    assert t.timed_steps(2, copy_input=True, read_loss=True) > 0
    assert np.isfinite(t.get_param(0)).all()


def test_snapshot_restore_roundtrip_on_device(rng, tmp_path):
    """Solver::Snapshot / Restore (solver.cpp:447-604, sgd_solver.cpp:261-353) on the device: step, snapshot, step  ==  fresh
    solver, restore, step -- bitwise, including BatchNorm running statistics and momentum history; a SECOND snapshot of the
    same process holds the current weights (the fused update writes the arena behind the Blob's host mirror); the
    .solverstate history list has one blob per Net::learnable_params() entry (5 per BatchNorm layer)."""
    spec = no.mini_resnet()
    t, params, data, label = make_trainer(spec, rng)
    t.step(2)
    state = t.snapshot(str(tmp_path / "snap"))
    assert state.endswith("_iter_2.solverstate")
    t.step(2)
    want = [t.get_param(i, 0).copy() for i in range(t.num_params())]
    want_h = [t.get_param(i, 2).copy() for i in range(t.num_params())]
    want_loss = t.loss()
    state2 = t.snapshot(str(tmp_path / "snap"))                  # iter 4: must contain the iter-4 weights, not the iter-2 ones
    # fresh solver with different weights, same batch
    t2, _, _, _ = make_trainer(spec, np.random.default_rng(7))
    t2.set_blob("data", data)
    t2.set_blob("label", label)
    t2.restore(state)
    assert t2.iter() == 2
    t2.step(2)
    assert t2.loss() == want_loss
    for i in range(t.num_params()):
        assert np.array_equal(t2.get_param(i, 0).view(np.uint32), want[i].view(np.uint32)), i
        assert np.array_equal(t2.get_param(i, 2).view(np.uint32), want_h[i].view(np.uint32)), i
    t3, _, _, _ = make_trainer(spec, np.random.default_rng(8))
    t3.restore(state2)
    assert t3.iter() == 4
    for i in range(t.num_params()):
        assert np.array_equal(t3.get_param(i, 0).view(np.uint32), want[i].view(np.uint32)), ("second snapshot is stale", i)
    # history length = every layer blob, like the reference's SolverState
    from caffe_mpi_b200 import host_api as ha
    L = ha.lib()
    L.b2h_wire_load.restype = __import__("ctypes").c_void_p
    h = L.b2h_wire_load(state.encode(), 1)
    assert h
    L.b2h_wire_num_history.argtypes = [__import__("ctypes").c_void_p]
    assert L.b2h_wire_num_history(h) == t.num_learnable() > t.num_params()
    L.b2h_wire_destroy.argtypes = [__import__("ctypes").c_void_p]
    L.b2h_wire_destroy(h)


def test_fusion_pass_is_bitwise_neutral(rng):
    """BatchNorm+ReLU / Eltwise+ReLU fusion and x_norm recomputation (csrc/layers_fused.cu) against the unfused graph: loss,
    every parameter gradient and three SGD steps, bit for bit."""
    spec = no.mini_resnet()
    a, params, data, label = make_trainer(spec, rng, fuse=False)
    b, _, _, _ = make_trainer(spec, np.random.default_rng(3), fuse=True)
    for i, p in enumerate(params):
        b.set_param(i, p)
    b.set_blob("data", data)
    b.set_blob("label", label)
    assert a.forward_backward() == b.forward_backward()
    for i in range(a.num_params()):
        assert np.array_equal(a.get_param(i, 1).view(np.uint32), b.get_param(i, 1).view(np.uint32)), i
    a.step(3); b.step(3)
    assert a.loss() == b.loss()
    for i in range(a.num_params()):
        assert np.array_equal(a.get_param(i, 0).view(np.uint32), b.get_param(i, 0).view(np.uint32)), i


def test_inception_style_net_matches_oracle(rng):
    """The layer kinds AlexNet / GoogLeNet / VGG-16 add (grouped conv, LRN, Concat, Dropout, two losses with loss_weight 0.3 / 1)
    through TrainNet against the net oracle: both losses, blobs on every branch, every parameter gradient (the auxiliary
    classifier's 0.3 included), then two SGD steps with fresh dropout masks per iteration."""
    spec = no.mini_inception()
    t, params, data, label = make_trainer(spec, rng)
    loss = t.forward_backward()
    ref_loss, grads, v, d = no.forward_backward(spec, params, data, label)
    assert abs(loss - ref_loss) <= TOL * abs(ref_loss)
    assert abs(float(t.get_blob("aux/loss1")[0]) - float(v["aux/loss1"])) <= TOL * abs(float(v["aux/loss1"]))
    for name in ("conv1", "norm1", "pool1", "inc/3x3", "inc/5x5", "inc/pool_proj", "inc/output", "aux/fc", "pool5", "cls"):
        assert rel(t.get_blob(name), v[name]) <= TOL, name
    for name in ("cls", "aux/cls", "inc/output", "pool1", "norm1"):
        assert rel(t.get_blob(name, diff=True), d[name]) <= TOL, name
    floor = 1e-3 * max(float(np.max(np.abs(g))) for g in grads)
    for i, g in enumerate(grads):
        assert rel(t.get_param(i, 1), g, floor) <= TOL, no.param_shapes(spec)[i]
    # two iterations of Solver::Step: the dropout streams advance by the blob size per iteration
    t.clear_param_diffs()                       # Step() accumulates into the diffs and relies on the update clearing them (solver.cpp:237-239)
    p, h = [q.copy() for q in params], [np.zeros_like(q) for q in params]
    import oracle
    for it in range(2):
        _, g, _, _ = no.forward_backward(spec, p, data, label, iteration=it + 1)     # iteration 0 was the forward_backward above
        for i in range(len(p)):
            _, w, hh = oracle.sgd_update(g[i], p[i], h[i], 0.9, 0.05, 0.0005)
            p[i], h[i] = w.reshape(p[i].shape), hh.reshape(p[i].shape)
        t.step(1)
    for i in range(len(p)):
        assert rel(t.get_param(i, 0), p[i]) <= 2 * TOL, i
