"""CPU tests pinning oracle/layers_oracle.py (the non-conv layer oracle) with closed forms and finite differences,
the way the reference pins these layers with GradientChecker (test_gradient_check_util.hpp, step 1e-2 / thr 1e-3)."""
import numpy as np
import pytest

from oracle import layers_oracle as lo


def fd_check(f, x, grad, rng, n=10, step=1e-3, thr=2e-3):
    flat = x.reshape(-1)
    for idx in rng.choice(flat.size, size=min(n, flat.size), replace=False):
        old = flat[idx]
        flat[idx] = old + step; lp = f()
        flat[idx] = old - step; lm = f()
        flat[idx] = old
        est = (lp - lm) / (2 * step)
        got = float(grad.reshape(-1)[idx])
        assert abs(est - got) <= thr * max(abs(est), abs(got), 1.0), (idx, est, got)


def test_relu(rng):
    x = rng.standard_normal((2, 3, 4, 5)).astype(np.float32)
    dy = rng.standard_normal(x.shape).astype(np.float32)
    assert np.array_equal(lo.relu_forward(x), np.where(x > 0, x, 0))
    assert np.array_equal(lo.relu_backward(dy, x), np.where(x > 0, dy, 0))


def test_batchnorm_forward_stats_and_running_average(rng):
    x = (rng.standard_normal((4, 3, 5, 5)) * 2 + 1).astype(np.float32)
    g, b = rng.standard_normal(3).astype(np.float32), rng.standard_normal(3).astype(np.float32)
    y, xn, mean, invstd, rm, rv = lo.bn_forward_train(x, g, b, 1e-4, 0.9, np.zeros(3, np.float32), np.zeros(3, np.float32), True)
    assert np.allclose(xn.mean(axis=(0, 2, 3)), 0, atol=1e-5) and np.allclose(xn.var(axis=(0, 2, 3)), 1, atol=2e-3)
    assert np.allclose(rm, x.mean(axis=(0, 2, 3)), atol=1e-5) and np.allclose(rv, x.var(axis=(0, 2, 3)) + 1e-4, rtol=1e-4)
    y2, *_, rm2, rv2 = lo.bn_forward_train(x * 2, g, b, 1e-4, 0.9, rm, rv, False)
    assert np.allclose(rm2, 0.1 * (2 * x).mean(axis=(0, 2, 3)) + 0.9 * rm, rtol=1e-5, atol=1e-6)
    assert np.allclose(y, xn * g.reshape(1, 3, 1, 1) + b.reshape(1, 3, 1, 1), atol=1e-6)


def test_batchnorm_backward_finite_differences(rng):
    x = rng.standard_normal((3, 2, 4, 4)).astype(np.float64)
    g, b = rng.standard_normal(2), rng.standard_normal(2)
    w = rng.standard_normal(x.shape)

    def loss():
        y, *_ = lo.bn_forward_train(x.astype(np.float32), g.astype(np.float32), b.astype(np.float32), 1e-4, 0.9, None, None, True)
        return float((y.astype(np.float64) * w).sum())
    y, xn, mean, invstd, *_ = lo.bn_forward_train(x.astype(np.float32), g.astype(np.float32), b.astype(np.float32), 1e-4, 0.9, None, None, True)
    dg, db, dx = lo.bn_backward(w.astype(np.float32), xn, g.astype(np.float32), invstd)
    fd_check(loss, x, dx, rng, step=1e-2, thr=5e-3)
    fd_check(loss, g, dg, rng, step=1e-2, thr=5e-3)
    fd_check(loss, b, db, rng, step=1e-2, thr=5e-3)


@pytest.mark.parametrize("H,k,s,p,method", [(112, 3, 2, 0, 0), (7, 7, 1, 0, 1), (8, 3, 2, 1, 0), (9, 2, 2, 0, 1), (13, 3, 2, 0, 0)])
def test_pooling(rng, H, k, s, p, method):
    x = rng.standard_normal((2, 3, H, H)).astype(np.float32)
    y, mask = lo.pool_forward(x, method, (k, k), (s, s), (p, p))
    Ho = lo.pooled_extent(H, k, s, p)
    assert y.shape == (2, 3, Ho, Ho)
    if (H, k, s, p) == (112, 3, 2, 0):
        assert Ho == 56                           # ResNet-50 pool1: ceil((112-3)/2)+1
    dy = rng.standard_normal(y.shape).astype(np.float32)
    dx = lo.pool_backward(dy, mask, x.shape, method, (k, k), (s, s), (p, p))
    xd = x.astype(np.float64)

    def loss():
        yy, _ = lo.pool_forward(xd.astype(np.float32), method, (k, k), (s, s), (p, p))
        return float((yy.astype(np.float64) * dy).sum())
    fd_check(loss, xd, dx, rng, step=1e-3, thr=5e-3)
    assert abs(float(dx.sum()) - float(dy.sum())) < 1e-2 * max(1.0, abs(float(dy.sum()))) or p > 0


def test_softmax_loss(rng):
    z = rng.standard_normal((5, 7)).astype(np.float64)
    lab = rng.integers(0, 7, 5).astype(np.float32)
    p, loss = lo.softmax_loss_forward(z.astype(np.float32), lab)
    assert np.allclose(p.sum(axis=1), 1, atol=1e-6)
    dx = lo.softmax_loss_backward(p, lab)
    fd_check(lambda: float(lo.softmax_loss_forward(z.astype(np.float32), lab)[1]), z, dx, rng, step=1e-2, thr=5e-3)
    # uniform logits -> loss = log(C)
    _, l0 = lo.softmax_loss_forward(np.zeros((4, 10), np.float32), np.array([1, 2, 3, 4], np.float32))
    assert l0 == pytest.approx(np.log(10), rel=1e-6)


def test_inner_product(rng):
    x = rng.standard_normal((4, 2, 3, 3)).astype(np.float32)
    w = rng.standard_normal((5, 18)).astype(np.float32)
    b = rng.standard_normal(5).astype(np.float32)
    y = lo.ip_forward(x, w, b)
    assert np.allclose(y[1, 2], x[1].reshape(-1) @ w[2] + b[2], rtol=1e-5, atol=1e-5)
    dy = rng.standard_normal(y.shape).astype(np.float32)
    dw, db, dx = lo.ip_backward(x, w, dy)
    assert np.allclose(db, dy.sum(0), atol=1e-5) and dw.shape == w.shape and dx.shape == x.shape
    assert np.allclose((dx * x).sum(), (dw * w).sum(), rtol=1e-4)     # <dx,x> == <dw,w> == <dy, y - b>
