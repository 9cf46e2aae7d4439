"""Per-blob comparison of TrainNet with tests/netoracle.py on the mini ResNet (diagnostic; GPU)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import netoracle as no
from caffe_mpi_b200 import capi, host_api

SOLVER = 'base_lr: 0.05 lr_policy: "fixed" momentum: 0.9 weight_decay: 0.0005 max_iter: 100 solver_mode: GPU'
if len(sys.argv) > 1 and sys.argv[1] == "simt":
    capi.lib().b2c_set_default_algo(1)
    print("== default algo SIMT")
rng = np.random.default_rng(1701)
spec = no.mini_resnet()
shapes = no.param_shapes(spec)
t = host_api.Trainer(no.to_prototxt(spec), SOLVER, num_classes=10)
params = []
for i, (layer, kind, shp) in enumerate(shapes):
    p = (rng.standard_normal(shp) * np.sqrt(2.0 / np.prod(shp[1:]))).astype(np.float32) if kind == "w" else \
        rng.uniform(0.5, 1.5, shp).astype(np.float32) if kind == "scale" else rng.uniform(-0.2, 0.2, shp).astype(np.float32)
    t.set_param(i, p); params.append(p)
    back = t.get_param(i)
    if not np.array_equal(back, p.reshape(-1)):
        print("PARAM READBACK MISMATCH", i, layer, kind)
d0 = spec[0]["shape"]
data = rng.standard_normal(d0).astype(np.float32); label = rng.integers(0, 10, d0[0]).astype(np.float32)
t.set_blob("data", data); t.set_blob("label", label)
loss = t.forward_backward()
ref_loss, grads, v, d = no.forward_backward(spec, params, data, label)
print("loss", loss, ref_loss)
rel = lambda a, r: float(np.max(np.abs(a.reshape(-1) - r.reshape(-1))) / max(np.max(np.abs(r)), 1e-20))
print("label back", t.get_blob("label"), label)
for L in spec:
    n = L["n"]
    if L["t"] in ("relu", "loss"):
        continue
    print("fwd %-22s %.3e" % (n, rel(t.get_blob(n), v[n])))
for L in reversed(spec):
    n = L["n"]
    if L["t"] in ("relu", "loss", "data") or n not in d:
        continue
    print("bwd %-22s %.3e" % (n, rel(t.get_blob(n, diff=True), d[n])))
for i, g in enumerate(grads):
    print("grad %-3d %-22s %-6s %.3e" % (i, shapes[i][0], shapes[i][1], rel(t.get_param(i, 1), g)))
