"""Generate tests/golden/conv_ref_golden.npz from oracle/_ref, i.e. from the reference's own
src/caffe/util/im2col.cpp compiled verbatim, driven through the per-image / per-group
ConvolutionLayer CPU loop with OpenBLAS cblas_sgemm/sgemv (the reference's BLAS := open).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixture travels to the GPU box; /root/reference does not."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as o  # noqa: E402
from cases import make  # noqa: E402

CASES = [
    ("simple", dict(N=2, Cin=3, H=6, W=4, O=4, k=3, s=2, p=0, d=1, G=1, bias=True)),
    ("group3", dict(N=2, Cin=6, H=6, W=4, O=3, k=3, s=2, p=0, d=1, G=3, bias=True)),
    ("dilated", dict(N=2, Cin=3, H=11, W=9, O=4, k=3, s=1, p=0, d=2, G=1, bias=True)),
    ("one_by_one", dict(N=2, Cin=8, H=6, W=4, O=4, k=1, s=1, p=0, d=1, G=1, bias=True)),
    ("rect", dict(N=2, Cin=4, H=9, W=11, O=5, k=(3, 5), s=(2, 1), p=(1, 2), d=(1, 1), G=1, bias=True)),
    ("res_3x3", dict(N=2, Cin=32, H=14, W=14, O=48, k=3, s=1, p=1, d=1, G=1, bias=False)),
    ("res_1x1_s2", dict(N=2, Cin=32, H=14, W=14, O=24, k=1, s=2, p=0, d=1, G=1, bias=False)),
    ("stem_7x7", dict(N=1, Cin=3, H=30, W=30, O=16, k=7, s=2, p=3, d=1, G=1, bias=False)),
    ("alex_g2", dict(N=2, Cin=16, H=13, W=13, O=24, k=5, s=1, p=2, d=1, G=2, bias=True)),
]


def main():
    assert o.ref() is not None and o.ref_blas_open(1), "oracle/_ref or OpenBLAS unavailable"
    rng = np.random.default_rng(1701)
    out = {}
    for name, c in CASES:
        prm = make(o, c)
        x = rng.standard_normal(prm.x_shape()).astype(np.float32)
        w = (rng.standard_normal(prm.w_shape()) * (2.0 / prm.Kd) ** 0.5).astype(np.float32)
        b = (rng.standard_normal(prm.O) * 0.1).astype(np.float32) if prm.has_bias else None
        dy = rng.standard_normal(prm.y_shape()).astype(np.float32)
        y = np.zeros(prm.y_shape(), np.float32)
        dw = np.zeros(prm.w_shape(), np.float32)
        db = np.zeros(prm.O, np.float32)
        dx = np.zeros(prm.x_shape(), np.float32)
        assert o.ref_conv_fwd_bwd(prm, x, w, b, y=y, dy=dy, dw=dw, db=db if prm.has_bias else None, dx=dx) == 0
        col0 = np.empty((prm.C * prm.kh * prm.kw, prm.Ho, prm.Wo), np.float32)
        o.ref().ref_im2col_cpu(np.ascontiguousarray(x[0]), prm.C, prm.H, prm.W, prm.kh, prm.kw, prm.ph, prm.pw,
                               prm.sh, prm.sw, prm.dh, prm.dw, col0)
        keys = [f for f, _ in o.ConvParams._fields_]
        out[name + "/keys"] = np.array(keys)
        out[name + "/vals"] = np.array([getattr(prm, k) for k in keys], np.int64)
        for k, v in (("x", x), ("w", w), ("dy", dy), ("y", y), ("dw", dw), ("dx", dx), ("col0", col0)):
            out[name + "/" + k] = v
        if prm.has_bias:
            out[name + "/b"] = b
            out[name + "/db"] = db
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conv_ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
