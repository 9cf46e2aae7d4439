"""Data-parallel TrainNet over P2PSync / ReduceScheduler (C++ host layer) on N GPUs: one process per GPU,
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/trainer_multi.py [batch] [steps]
Checks that every rank holds bit-identical weights after the steps (the allreduce + fused update are deterministic and
identical on all ranks) and prints whole-job images/sec (device time, max over ranks)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from caffe_mpi_b200 import host_api, models

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.cuda.set_device(local)
dist.init_process_group("gloo")
t = host_api.Trainer(models.resnet50_prototxt(N), models.RESNET50_SOLVER.replace("base_lr: 0.001", "base_lr: 0.01"), batch=N, seed=1701 + rank)
w_before = t.get_param(0).copy()
if world > 1:
    ids = [t.new_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    t.attach_sync(world, rank, ids[0])                 # P2PSync::on_start: weights broadcast from rank 0
t.step(3)
t.sync()
first = t.loss()
dist.barrier()
ms = t.timed_steps(K)
loss = t.loss()
tm = torch.tensor([ms], dtype=torch.float64)
dist.all_reduce(tm, op=dist.ReduceOp.MAX)
# weights must be bit-identical on every rank
sums = []
for i in (0, 1, 2, 80, 159, 160):
    p = t.get_param(i)
    sums.append(float(np.sum(p.astype(np.float64))))
    sums.append(float(p.view(np.uint32).astype(np.uint64).sum()))
mine = torch.tensor(sums, dtype=torch.float64)
allv = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allv, mine)
same = all(bool(torch.equal(allv[0], v)) for v in allv)
losses = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
dist.all_gather(losses, torch.tensor([loss], dtype=torch.float64))
if rank == 0:
    print(json.dumps({"n_gpus": world, "per_gpu_batch": N, "steps": K, "ms_per_step": float(tm) / K,
                      "images_per_sec": N * world * K / (float(tm) / 1e3), "weights_identical_across_ranks": same,
                      "loss_first": first, "loss_last_per_rank": [float(l) for l in losses],
                      "weights_changed": bool(np.any(t.get_param(0) != w_before))}))
dist.barrier()
dist.destroy_process_group()
