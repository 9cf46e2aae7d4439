"""First-run diagnosis of the staged weight-gradient kernel (conv_tc_wgrad_stg.cu): non-fatal mbarrier timeouts, dW against the
3xTF32 gather / TMA kernels on the same inputs, decoded stuck waits."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import caffe_mpi_b200 as m
from caffe_mpi_b200 import capi

L = m.lib()
L.b2c_debug_mbar_set_trap(0)
shapes = [(2, 64, 8, 8, 64, 1, 0), (2, 32, 12, 12, 40, 3, 1), (3, 96, 14, 14, 72, 1, 0), (4, 64, 28, 28, 64, 3, 1), (2, 64, 56, 56, 256, 1, 0),
          (64, 256, 14, 14, 256, 3, 1), (64, 64, 56, 56, 64, 3, 1), (64, 256, 56, 56, 64, 1, 0), (2, 32, 8, 80, 32, 3, 1), (5, 32, 8, 13, 40, 3, 1)]
if os.environ.get("B2C_WGRAD_STAGED_PLANE") == "1":      # 7x7 maps: plane mode
    shapes = [(9, 64, 7, 7, 72, 3, 1), (11, 96, 7, 7, 160, 1, 0), (3, 512, 7, 7, 512, 3, 1), (64, 2048, 7, 7, 512, 1, 0), (64, 512, 7, 7, 2048, 1, 0)]
if len(sys.argv) >= 8:
    shapes = [tuple(int(a) for a in sys.argv[1:8])]
torch.manual_seed(0)
for (N, Cc, H, W, O, k, p) in shapes:
    prm = capi.ConvParams.make(N, Cc, H, W, O, k, 1, p, 1, 1, False)
    ds, dg = m.ConvDesc(prm, capi.ENGINE_DEFAULT, math=0), m.ConvDesc(prm, capi.ENGINE_DEFAULT, math=2)
    x = torch.randn(prm.x_shape(), device="cuda"); dy = torch.randn(prm.y_shape(), device="cuda")
    dw0 = torch.randn(prm.w_shape(), device="cuda") * 0.1
    a, b = dw0.clone(), dw0.clone()
    dg.backward_filter(x, dy, b); torch.cuda.synchronize()
    try:
        ds.backward_filter(x, dy, a); torch.cuda.synchronize()
        err = float((a - b).abs().max() / b.abs().max())
        print(f"N{N} C{Cc} {H}x{W} O{O} k{k}: dW err vs 3xTF32 {err:.3e}  nan={bool(torch.isnan(a).any())}", flush=True)
    except Exception as e:
        print(f"N{N} C{Cc} {H}x{W} O{O} k{k}: staged wgrad raised", repr(e)[:160]); break
    buf = (C.c_uint * 512)()
    n = L.b2c_debug_mbar_timeouts(buf, 512)
    if n:
        print("  timeouts recorded:", n)
        blk = buf[384:512]
        role = lambda t: "conv%d" % (t // 32) if t < 512 else ("TMA" if t // 32 == 16 else "MMA")
        names = {0: "dy_full[0]", 8: "dy_full[1]", 16: "dy_empty[0]", 24: "dy_empty[1]", 32: "x_full[0]", 40: "x_full[1]", 48: "x_empty[0]", 56: "x_empty[1]",
                 64: "b_full[0]", 72: "b_full[1]", 80: "b_empty[0]", 88: "b_empty[1]", 96: "done"}
        for i in range(min(blk[0], 31)):
            bb, t, bar, par = blk[4 + 4 * i: 8 + 4 * i]
            print(f"    block {bb & 0xffff},{bb >> 16} thread {t} ({role(t)} lane {t % 32}) stuck on {names.get(bar & 0xff, hex(bar & 0xff))} parity {par}")
        break
