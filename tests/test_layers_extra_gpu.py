"""GPU parity of the AlexNet / GoogLeNet / VGG-16 extras (LRN, Dropout mask, elementwise product) against
oracle/layers_oracle.py (LRN: 1e-5 forward / 1e-4 backward relative, blob level; dropout mask and product: bit-exact)."""
import ctypes as C

import numpy as np
import pytest

from oracle import layers_oracle as lo
from cases import rel_err

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ptr(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("shape,size,alpha,beta,k", [((2, 7, 3, 3), 5, 1.0, 0.75, 1.0), ((4, 96, 27, 27), 5, 1e-4, 0.75, 1.0),
                                                   ((2, 3, 5, 5), 15, 1.0, 0.75, 2.0), ((3, 64, 8, 8), 3, 0.01, 0.5, 1.0)])
def test_lrn(rng, shape, size, alpha, beta, k):
    from caffe_mpi_b200 import capi
    L = capi.lib()
    x = rng.standard_normal(shape).astype(np.float32)
    dy = rng.standard_normal(shape).astype(np.float32)
    y_ref, s_ref = lo.lrn_forward(x, size, alpha, beta, k)
    dx_ref = lo.lrn_backward(x, y_ref, s_ref, dy, size, alpha, beta)
    N, Cc, S = shape[0], shape[1], shape[2] * shape[3]
    X, DY = dev(x), dev(dy)
    Y, SC, DX = torch.empty_like(X), torch.empty_like(X), torch.empty_like(X)
    assert L.b2c_lrn_forward(N, Cc, S, size, alpha, beta, k, ptr(X), ptr(SC), ptr(Y), None) == 0
    assert L.b2c_lrn_backward(N, Cc, S, size, alpha, beta, ptr(X), ptr(Y), ptr(SC), ptr(DY), ptr(DX), None) == 0
    torch.cuda.synchronize()
    assert rel_err(SC.cpu().numpy(), s_ref) < 1e-5
    assert rel_err(Y.cpu().numpy(), y_ref) < 1e-5
    assert rel_err(DX.cpu().numpy(), dx_ref) < 1e-4


@pytest.mark.parametrize("n,ratio,seed,offset", [(100003, 0.5, 1701, 0), (4096, 0.3, 7, 123456789), (17, 0.9, 2 ** 40 + 5, 3)])
def test_dropout_mask_is_bit_exact_and_mul(rng, n, ratio, seed, offset):
    from caffe_mpi_b200 import capi
    L = capi.lib()
    M = torch.empty(n, device="cuda")
    assert L.b2c_dropout_mask(n, ratio, seed, offset, ptr(M), None) == 0
    m = M.cpu().numpy()
    assert np.array_equal(m, lo.dropout_mask(n, ratio, seed, offset))
    x = rng.standard_normal(n).astype(np.float32)
    X, Y = dev(x), torch.empty(n, device="cuda")
    assert L.b2c_mul(n, ptr(X), ptr(M), ptr(Y), None) == 0
    assert np.array_equal(Y.cpu().numpy(), x * m)


@pytest.mark.parametrize("N,Cc,k", [(64, 1000, 1), (64, 1000, 5), (7, 10, 3), (33, 17, 17)])
def test_accuracy_top_k(rng, N, Cc, k):
    from caffe_mpi_b200 import capi
    L = capi.lib()
    z = rng.standard_normal((N, Cc)).astype(np.float32)
    z[:, : Cc // 2] = np.round(z[:, : Cc // 2] * 2) / 2          # ties: equal scores rank the higher class index first
    lab = rng.integers(0, Cc, N).astype(np.float32)
    Z, LB = dev(z), dev(lab)
    acc, scratch = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    L.b2c_accuracy.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert L.b2c_accuracy(N, Cc, k, ptr(Z), ptr(LB), ptr(acc), ptr(scratch), None) == 0
    torch.cuda.synchronize()
    assert float(acc.item()) == float(lo.accuracy(z, lab, k))


@pytest.mark.parametrize("shape,crop,mean,scale", [((5, 3, 40, 36), (32, 32), "values", 0.0078125), ((3, 1, 28, 28), (28, 28), None, 0.00390625),
                                                   ((4, 3, 19, 23), (17, 21), "image", 1.0), ((2, 3, 256, 256), (224, 224), "values", 1.0)])
def test_transform_u8_is_bit_exact(rng, shape, crop, mean, scale):
    from caffe_mpi_b200 import capi
    L = capi.lib()
    N, Cc, Hd, Wd = shape
    src = rng.integers(0, 256, shape, dtype=np.uint8)
    ho = rng.integers(0, Hd - crop[0] + 1, N).astype(np.int32)
    wo = rng.integers(0, Wd - crop[1] + 1, N).astype(np.int32)
    mir = rng.integers(0, 2, N).astype(np.uint8)
    mv = np.array([104.0, 117.0, 123.0][:Cc], np.float32) if mean == "values" else None
    mi = rng.uniform(90, 130, (Cc, Hd, Wd)).astype(np.float32) if mean == "image" else None
    want = lo.transform_u8(src, crop, ho, wo, mir, mv, mi, scale)
    S, HO, WO, MIR = dev(src), dev(ho), dev(wo), dev(mir)
    MV = dev(mv) if mv is not None else None
    MI = dev(mi) if mi is not None else None
    out = torch.empty((N, Cc) + crop, device="cuda")
    L.b2c_transform_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
    assert L.b2c_transform_u8(ptr(S), N, Cc, Hd, Wd, crop[0], crop[1], ptr(HO), ptr(WO), ptr(MIR), ptr(MV) if MV is not None else None,
                              ptr(MI) if MI is not None else None, scale, ptr(out), None) == 0
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)
