"""A few full ResNet-50 TrainNet steps (for ncu launch lists): python tools/fullnet_step.py [batch] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_mpi_b200 import host_api, models
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
t = host_api.Trainer(models.resnet50_prototxt(N), models.RESNET50_SOLVER, batch=N)
t.step(K)
t.sync()
print("ms/step", t.timed_steps(K) / K, "loss", t.loss())
