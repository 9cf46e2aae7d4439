"""Diagnostics for the bulk-copy-staged bf16x3 kernel (conv_tc_stg.cu): staged result vs the 3xTF32 gather kernel on the same
inputs, per shape; when they disagree, tests the hypotheses that are cheap to tell apart (K pairs swapped in the packed
tensor-memory operand, hi/lo halves exchanged)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import caffe_mpi_b200 as m
from caffe_mpi_b200 import capi

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

shapes = [(2, 32, 12, 12, 32, 1, 0), (3, 64, 12, 11, 72, 1, 0), (2, 32, 12, 12, 32, 3, 1), (3, 64, 14, 14, 128, 3, 1), (5, 32, 10, 14, 48, 5, 2), (2, 64, 56, 56, 256, 1, 0),
          (4, 128, 28, 28, 128, 3, 1), (64, 256, 14, 14, 256, 3, 1), (64, 64, 56, 56, 64, 3, 1)]
torch.manual_seed(0)
for (N, C, H, W, O, k, p) in shapes:
    prm = capi.ConvParams.make(N, C, H, W, O, k, 1, p, 1, 1, True)
    ds = m.ConvDesc(prm, capi.ENGINE_DEFAULT, math=0)
    dg = m.ConvDesc(prm, capi.ENGINE_DEFAULT, math=2)
    x = torch.randn(prm.x_shape(), device="cuda"); w = torch.randn(prm.w_shape(), device="cuda") * (2.0 / (C * k * k)) ** 0.5
    b = torch.randn(O, device="cuda") * 0.1
    dy = torch.randn(prm.y_shape(), device="cuda")
    ys, yg = torch.full(prm.y_shape(), 7.0, device="cuda"), torch.empty(prm.y_shape(), device="cuda")
    try:
        ds.forward(x, w, b, ys); torch.cuda.synchronize()
    except Exception as e:
        print("shape", (N, C, H, W, O, k, p), "staged forward FAILED:", repr(e)[:300]); break
    dg.forward(x, w, b, yg)
    dxs, dxg = torch.full(prm.x_shape(), 7.0, device="cuda"), torch.empty(prm.x_shape(), device="cuda")
    ds.backward_data(dy, w, dxs); dg.backward_data(dy, w, dxg); torch.cuda.synchronize()
    ey, edx = rel(ys, yg), rel(dxs, dxg)
    print(f"N{N} C{C} {H}x{W} O{O} k{k}: y err {ey:.3e}  dx err {edx:.3e}  nan(y)={bool(torch.isnan(ys).any())} untouched(y)={int((ys == 7.0).sum())}", flush=True)
    if ey > 1e-3:
        perm = torch.arange(C, device="cuda").view(-1, 2).flip(1).reshape(-1)
        yp = torch.empty_like(yg); dg.forward(x[:, perm].contiguous(), w, b, yp); torch.cuda.synchronize()
        print("   hypothesis K pairs swapped in TMEM packing: err vs conv(x[pair-swapped]) =", f"{rel(ys, yp):.3e}")
        d = (ys - yg).abs()
        idx = torch.nonzero(d > 1e-2 * yg.abs().max())[:8].tolist()
        print("   first bad (n,o,h,w):", idx, " frac bad:", float((d > 1e-2 * yg.abs().max()).float().mean()))
        nob = torch.empty_like(yg); dg.forward(x, w, None if False else torch.zeros_like(b), nob); torch.cuda.synchronize()
        print("   err vs no-bias result:", f"{rel(ys, nob):.3e}")
