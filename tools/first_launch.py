import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import caffe_mpi_b200 as m
from caffe_mpi_b200 import capi
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
def t(tag, fn):
    torch.cuda.synchronize(); t0 = time.time(); fn(); torch.cuda.synchronize(); print(f"{tag}: {1e3*(time.time()-t0):.1f} ms", flush=True)
for (N, C, H, O, k, p) in ((8, 256, 14, 256, 3, 1), (8, 256, 14, 64, 1, 0), (8, 64, 14, 20, 3, 1), (8, 64, 14, 100, 3, 1)):
    prm = capi.ConvParams.make(N, C, H, H, O, k, 1, p, 1, 1, False)
    d = m.ConvDesc(prm)
    x = torch.randn(prm.x_shape(), device="cuda"); w = torch.randn(prm.w_shape(), device="cuda"); y = torch.empty(prm.y_shape(), device="cuda")
    dy = torch.randn(prm.y_shape(), device="cuda"); dx = torch.empty_like(x); dw = torch.zeros_like(w)
    for rep in range(2):
        t(f"O={O} k={k} fwd   call{rep}", lambda: d.forward(x, w, None, y))
        t(f"O={O} k={k} dgrad call{rep}", lambda: d.backward_data(dy, w, dx))
        t(f"O={O} k={k} wgrad call{rep}", lambda: d.backward_filter(x, dy, dw))
