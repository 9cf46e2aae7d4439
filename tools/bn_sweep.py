"""Time the fused BatchNorm [+ ReLU] forward / backward launches on ResNet-50's distinct BN shapes (batch 64 by default).
Algorithmic bytes: forward 8 B/element (read x, write y), backward 12 B/element (read dy, x, write dx).
Environment: B2C_BN_ONEPASS=0|1, B2C_BN_OCC=1..4 select the form being timed."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import caffe_mpi_b200 as m
from caffe_mpi_b200 import capi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
# (C, H, count in ResNet-50)
SHAPES = [(64, 112, 1), (64, 56, 6), (256, 56, 4), (128, 28, 8), (512, 28, 5), (256, 14, 12), (1024, 14, 7), (512, 7, 6), (2048, 7, 4)]
L = m.lib()
p = lambda t: C.c_void_p(t.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot_f = tot_b = 0.0
print(f"{'shape':>22} {'fwd us':>8} {'GB/s':>7} {'bwd us':>8} {'GB/s':>7}")
for (Cc, H, cnt) in SHAPES:
    S = H * H
    X = torch.randn(N, Cc, H, H, device="cuda"); DY = torch.randn_like(X); Y = torch.empty_like(X); DX = torch.empty_like(X)
    G, B = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    RM, RV, SM, SI, DG, DB = (torch.zeros(Cc, device="cuda") for _ in range(6))
    fwd = lambda: capi.check(L.b2c_bn_forward_train_fused(N, Cc, S, p(X), p(G), p(B), 1e-4, 0.9, 0, p(RM), p(RV), p(SM), p(SI), p(Y), 1, None))
    bwd = lambda: capi.check(L.b2c_bn_backward_fused(N, Cc, S, p(DY), p(X), p(SM), p(SI), p(G), p(B), p(DG), p(DB), p(DX), 1, None))
    res = []
    for fn in (fwd, bwd):
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            flush.zero_()                                  # 256 MB > L2: every timed launch starts cold
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort(); res.append(ts[len(ts) // 2])
    el = N * Cc * S
    tot_f += res[0] * cnt; tot_b += res[1] * cnt
    print(f"{cnt}x N{N} C{Cc} {H}x{H:<6} {res[0]:8.1f} {8 * el / res[0] / 1e3:7.0f} {res[1]:8.1f} {12 * el / res[1] / 1e3:7.0f}", flush=True)
print(f"ResNet-50 totals (us, weighted): fwd {tot_f:.0f} bwd {tot_b:.0f}")
