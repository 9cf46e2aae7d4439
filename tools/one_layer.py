"""Run one conv layer op a few times (for ncu captures).  args: C H O k s p N op[fwd|dgrad|wgrad] [tf32]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import caffe_mpi_b200 as m
from caffe_mpi_b200 import capi
C, H, O, k, s, p, N = (int(a) for a in sys.argv[1:8])
op = sys.argv[8]
math = capi.MATH_TF32 if len(sys.argv) > 9 and sys.argv[9] == "tf32" else capi.MATH_FP32
prm = capi.ConvParams.make(N, C, H, H, O, k, s, p, 1, 1, False)
d = m.ConvDesc(prm, capi.ENGINE_DEFAULT, math=math)
x = torch.randn(prm.x_shape(), device="cuda"); w = torch.randn(prm.w_shape(), device="cuda") * 0.05
y = torch.empty(prm.y_shape(), device="cuda"); dy = torch.randn(prm.y_shape(), device="cuda")
dx = torch.empty_like(x); dw = torch.zeros_like(w)
fn = {"fwd": lambda: d.forward(x, w, None, y), "dgrad": lambda: d.backward_data(dy, w, dx), "wgrad": lambda: d.backward_filter(x, dy, dw)}[op]
for _ in range(3):
    fn()
torch.cuda.synchronize()
print("done", op, sys.argv[1:8])
