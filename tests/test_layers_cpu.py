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


# ---------------------------------------------------------------------------------------------------------------------
# Known answers restated from the reference's own layer tests (numbers only): they pin oracle/layers_oracle.py, which
# the GPU kernels are in turn compared with on random data (tests/test_layers_gpu.py).
MAGIC6 = np.array([[35, 1, 6, 26, 19, 24], [3, 32, 7, 21, 23, 25], [31, 9, 2, 22, 27, 20],
                   [8, 28, 33, 17, 10, 15], [30, 5, 34, 12, 14, 16], [4, 36, 29, 13, 18, 11]], np.float32)


def _tile(plane, num=2, channels=2):
    return np.broadcast_to(plane, (num, channels) + plane.shape).astype(np.float32).copy()


def test_pool_max_square_known_answer():
    """test_pooling_layer.cpp:53-124 (TestForwardSquare): 3x5 input, kernel 2 -> values and argmax mask."""
    x = _tile(np.array([[1, 2, 5, 2, 3], [9, 4, 1, 4, 8], [1, 2, 5, 2, 3]], np.float32))
    y, mask = lo.pool_forward(x, 0, (2, 2), (1, 1), (0, 0))
    assert y.shape == (2, 2, 2, 4)
    assert np.array_equal(y, _tile(np.array([[9, 5, 5, 8], [9, 5, 5, 8]], np.float32)))
    assert np.array_equal(mask, np.broadcast_to(np.array([[5, 2, 2, 9], [5, 12, 12, 9]]), mask.shape))


def test_pool_max_rect_high_known_answer():
    """test_pooling_layer.cpp:126-250 (TestForwardRectHigh): magic(6), kernel 3x2."""
    y, mask = lo.pool_forward(_tile(MAGIC6), 0, (3, 2), (1, 1), (0, 0))
    want = np.array([[35, 32, 26, 27, 27], [32, 33, 33, 27, 27], [31, 34, 34, 27, 27], [36, 36, 34, 18, 18]], np.float32)
    wmask = np.array([[0, 7, 3, 16, 16], [7, 20, 20, 16, 16], [12, 26, 26, 16, 16], [31, 31, 26, 34, 34]])
    assert np.array_equal(y, _tile(want))
    assert np.array_equal(mask, np.broadcast_to(wmask, mask.shape))


def test_pool_max_rect_wide_known_answer():
    """test_pooling_layer.cpp:252-380 (TestForwardRectWide): magic(6), kernel 2x3."""
    y, mask = lo.pool_forward(_tile(MAGIC6), 0, (2, 3), (1, 1), (0, 0))
    want = np.array([[35, 32, 26, 26], [32, 32, 27, 27], [33, 33, 33, 27], [34, 34, 34, 17], [36, 36, 34, 18]], np.float32)
    wmask = np.array([[0, 7, 3, 3], [7, 7, 16, 16], [20, 20, 20, 16], [26, 26, 26, 21], [31, 31, 26, 34]])
    assert np.array_equal(y, _tile(want))
    assert np.array_equal(mask, np.broadcast_to(wmask, mask.shape))


def test_pool_max_padded_known_answer():
    """test_pooling_layer.cpp:483-527 (TestForwardMaxPadded): kernel 3, stride 2, pad 2 on 3x3."""
    x = np.array([[1, 2, 4], [2, 3, 2], [4, 2, 1]], np.float32).reshape(1, 1, 3, 3)
    y, _ = lo.pool_forward(x, 0, (3, 3), (2, 2), (2, 2))
    assert np.array_equal(y.reshape(3, 3), np.array([[1, 4, 4], [4, 4, 4], [4, 4, 1]], np.float32))


def test_pool_ave_known_answer():
    """test_pooling_layer.cpp:547-578 (TestForwardAve): constant 2, kernel 3, stride 1, pad 1 -> the padded-window divisor."""
    y, _ = lo.pool_forward(np.full((1, 1, 3, 3), 2, np.float32), 1, (3, 3), (1, 1), (1, 1))
    want = np.array([[8 / 9, 4 / 3, 8 / 9], [4 / 3, 2.0, 4 / 3], [8 / 9, 4 / 3, 8 / 9]])
    np.testing.assert_allclose(y.reshape(3, 3), want, atol=1e-5)


def test_pool_setup_shapes():
    """test_pooling_layer.cpp:382-425 (TestSetup / TestSetupPadded / TestSetupGlobalPooling) on the 2x3x6x5 fixture."""
    assert (lo.pooled_extent(6, 3, 2, 0), lo.pooled_extent(5, 3, 2, 0)) == (3, 2)
    assert (lo.pooled_extent(6, 3, 2, 1), lo.pooled_extent(5, 3, 2, 1)) == (4, 3)
    assert (lo.pooled_extent(6, 6, 1, 0), lo.pooled_extent(5, 5, 1, 0)) == (1, 1)


def test_batchnorm_forward_statistics():
    """test_batch_norm_layer.cpp TestForward: without scale/bias every channel of the output has mean 0 and variance 1
    (up to eps) over N*H*W."""
    rng = np.random.default_rng(1701)
    x = (rng.standard_normal((5, 2, 3, 4)) * 3 + 7).astype(np.float32)
    y = lo.bn_forward_train(x, None, None, 1e-5, 0.999, None, None, True)[0]
    for c in range(2):
        v = y[:, c].astype(np.float64)
        assert abs(v.mean()) < 1e-5 and abs(v.var() - 1.0) < 1e-3


def test_relu_known_answer():
    """test_neuron_layer.cpp TestReLU / TestReLUWithNegativeSlope (slope 0.01)."""
    x = np.array([-2.0, -0.5, 0.0, 0.5, 3.0], np.float32)
    assert np.array_equal(lo.relu_forward(x), np.array([0, 0, 0, 0.5, 3.0], np.float32))
    np.testing.assert_allclose(lo.relu_forward(x, 0.01), np.array([-0.02, -0.005, 0, 0.5, 3.0], np.float32), rtol=1e-6)


def test_inner_product_known_answer():
    """inner_product_layer.cpp semantics on a hand-computed case: y = x W^T + b with W [num_output x K] (transpose false)."""
    x = np.array([[1, 2, 3], [0, -1, 4]], np.float32)
    w = np.array([[1, 0, -1], [2, 1, 0]], np.float32)
    b = np.array([0.5, -1], np.float32)
    assert np.array_equal(lo.ip_forward(x, w, b), np.array([[-1.5, 3], [-3.5, -2]], np.float32))
    dw, db, dx = lo.ip_backward(x, w, np.ones((2, 2), np.float32))
    assert np.array_equal(dw, np.array([[1, 1, 7], [1, 1, 7]], np.float32)) and np.array_equal(db, np.array([2, 2], np.float32))
    assert np.array_equal(dx, np.array([[3, 1, -1], [3, 1, -1]], np.float32))


def test_softmax_loss_known_answer():
    """softmax_loss_layer.cpp:96-160: uniform logits over C classes give loss ln C; gradient (p - onehot) / N."""
    z = np.zeros((4, 10), np.float32)
    lab = np.array([0, 3, 9, 5], np.float32)
    p, loss = lo.softmax_loss_forward(z, lab)
    assert abs(float(loss) - np.log(10)) < 1e-6
    d = lo.softmax_loss_backward(p, lab)
    assert abs(d[0, 0] - (0.1 - 1) / 4) < 1e-7 and abs(d[0, 1] - 0.1 / 4) < 1e-7


def test_lrn_matches_the_reference_tests_formula_and_its_gradient(rng):
    """Forward against a literal per-element restatement of test_lrn_layer.cpp:61-91; backward by finite differences
    (the reference uses GradientChecker: TestGradientAcrossChannels)."""
    x = rng.standard_normal((2, 7, 3, 3)).astype(np.float32)
    for size, alpha, beta in ((5, 1.0, 0.75), (3, 1e-4, 0.75), (15, 1.0, 0.75)):      # 15 > channels: TestForwardAcrossChannelsLargeRegion
        y, scale = lo.lrn_forward(x, size, alpha, beta)
        for (n, c, h, w) in ((0, 0, 0, 0), (1, 3, 2, 1), (0, 6, 1, 2)):
            c0 = max(c - (size - 1) // 2, 0); c1 = min(c - (size - 1) // 2 + size, 7)
            s = 1.0 + sum(float(x[n, i, h, w]) ** 2 * alpha / size for i in range(c0, c1))
            assert abs(y[n, c, h, w] - x[n, c, h, w] / s ** beta) < 1e-5
    size, alpha, beta = 5, 1.0, 0.75
    y, scale = lo.lrn_forward(x, size, alpha, beta)
    dy = rng.standard_normal(x.shape).astype(np.float32)
    dx = lo.lrn_backward(x, y, scale, dy, size, alpha, beta)
    for idx in ((0, 0, 0, 0), (1, 3, 1, 1), (0, 6, 2, 2), (1, 5, 0, 2)):
        e = 1e-3
        xp, xm = x.copy(), x.copy()
        xp[idx] += e; xm[idx] -= e
        fd = ((lo.lrn_forward(xp, size, alpha, beta)[0].astype(np.float64) - lo.lrn_forward(xm, size, alpha, beta)[0]) * dy).sum() / (2 * e)
        assert abs(fd - dx[idx]) < 2e-3 * max(1.0, abs(fd))


def test_dropout_mask_statistics_and_scale():
    """dropout_layer.cpp: kept with probability 1 - ratio, kept values scaled by 1/(1 - ratio); reproducible per (seed, offset)."""
    m = lo.dropout_mask(200000, 0.5, 1701)
    assert set(np.unique(m)) == {0.0, 2.0} and abs((m > 0).mean() - 0.5) < 0.01
    m3 = lo.dropout_mask(200000, 0.3, 7, offset=10)
    assert abs((m3 > 0).mean() - 0.7) < 0.01 and abs(m3.max() - 1 / 0.7) < 1e-6
    assert np.array_equal(m3[5:100], lo.dropout_mask(95, 0.3, 7, offset=15))        # counter based: offset + i
    assert abs(m3.mean() - 1.0) < 0.01                                              # expectation preserved


def test_net_oracle_inception_style_finite_differences():
    """tests/netoracle.py with the AlexNet / GoogLeNet layer kinds (grouped conv, LRN, Concat, Dropout, two weighted losses): the
    analytic gradient of every parameter blob against central differences of 1.0 * loss + 0.3 * aux loss -- what the reference's
    GradientChecker does per layer (test_gradient_check_util.hpp), here through the whole graph."""
    import netoracle as no
    spec = no.mini_inception()
    shapes = no.param_shapes(spec)
    rng = np.random.default_rng(1)
    params = [(rng.standard_normal(s) * 0.3).astype(np.float32) for _, _, s in shapes]
    data = rng.standard_normal(spec[0]["shape"]).astype(np.float32)
    label = rng.integers(0, 10, spec[0]["shape"][0]).astype(np.float32)
    _, grads, _, _ = no.forward_backward(spec, params, data, label)

    def total(ps):
        l, _, vv, _ = no.forward_backward(spec, ps, data, label)
        return float(l) + 0.3 * float(vv["aux/loss1"])

    eps = 3e-3
    for i, g in enumerate(grads):
        idx = np.unravel_index(np.argmax(np.abs(g)), g.shape)
        p2 = [q.copy() for q in params]
        p2[i][idx] += eps
        up = total(p2)
        p2[i][idx] -= 2 * eps
        dn = total(p2)
        fd = (up - dn) / (2 * eps)
        assert abs(fd - g[idx]) <= 2e-3 * max(1.0, abs(g[idx])), (shapes[i], fd, g[idx])


def test_transform_oracle_matches_the_reference_loops(rng):
    """oracle.layers_oracle.transform_u8 (numpy) against a literal transcription of DataTransformer::Transform's index walk
    (data_transformer.cpp:233-283: top_index running backwards under mirror, data_index in datum coordinates)."""
    src = rng.integers(0, 256, (2, 3, 6, 7), dtype=np.uint8)
    crop, ho, wo, mir, mv, sc = (4, 5), np.array([1, 2]), np.array([0, 2]), np.array([1, 0]), [10.0, 20.0, 30.0], 0.5
    want = np.empty((2, 3, 4, 5), np.float32)
    for n in range(2):
        data, td = src[n].reshape(-1), np.empty(3 * 4 * 5, np.float32)
        for c in range(3):
            cdho, ch = c * 6 + ho[n], c * 4
            for h in range(4):
                top = (ch + h + 1) * 5 - 1 if mir[n] else (ch + h) * 5
                di = (cdho + h) * 7 + wo[n]
                for _ in range(5):
                    td[top] = (np.float32(data[di]) - np.float32(mv[c])) * np.float32(sc)
                    di += 1
                    top += -1 if mir[n] else 1
        want[n] = td.reshape(3, 4, 5)
    assert np.array_equal(lo.transform_u8(src, crop, ho, wo, mir, mv, None, sc), want)
