"""Worker of tests/test_multi_gpu.py (one process per GPU under torch.distributed.run): a BatchNorm-free net (LeNet) trained
for a few iterations through the C++ TrainNet + P2PSync + ReduceScheduler on this rank's slice of a fixed global batch; dumps
the final parameters, history and losses of this rank to <outdir>/rank<r>.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_case(global_batch, seed=1701):
    import netoracle as no
    rng = np.random.default_rng(seed)
    spec = no.lenet(batch=global_batch)
    shapes = no.param_shapes(spec)
    params = []
    for (_, kind, shp) in shapes:
        if kind == "w":
            params.append(rng.standard_normal(shp).astype(np.float32) * np.float32(np.sqrt(2.0 / np.prod(shp[1:]))))
        else:
            params.append(rng.uniform(-0.2, 0.2, shp).astype(np.float32))
    data = rng.standard_normal(spec[0]["shape"]).astype(np.float32)
    label = rng.integers(0, 10, global_batch).astype(np.float32)
    return params, data, label


SOLVER = 'base_lr: 0.02 lr_policy: "fixed" momentum: 0.9 weight_decay: 0.0005 max_iter: 100 solver_mode: GPU'


def run(trainer_batch, params, data, label, steps, sync=None, iter_size=1):
    import netoracle as no
    from caffe_mpi_b200 import host_api
    spec = no.lenet(batch=trainer_batch)
    solver = SOLVER + (f" iter_size: {iter_size}" if iter_size > 1 else "")
    t = host_api.Trainer(no.to_prototxt(spec), solver, num_classes=10)
    for i, p in enumerate(params):
        t.set_param(i, p)
    if sync is not None:                  # (world, rank, torch.distributed): the NCCL id travels over the gloo group
        world, rank, dist = sync
        ids = [t.new_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        t.attach_sync(world, rank, ids[0])    # P2PSync::on_start broadcasts rank 0's weights (identical here anyway)
    t.set_blob("data", data)
    t.set_blob("label", label)
    losses = []
    for _ in range(steps):
        t.step(1)
        losses.append(t.loss())
    t.sync()
    out = {"losses": np.array(losses, np.float64)}
    for i in range(len(params)):
        out[f"p{i}"] = t.get_param(i, 0)
        out[f"h{i}"] = t.get_param(i, 2)
    return out


def main():
    import torch
    import torch.distributed as dist
    outdir, global_batch, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    params, data, label = build_case(global_batch)
    per = global_batch // world
    sl = slice(rank * per, (rank + 1) * per)
    out = run(per, params, data[sl], label[sl], steps, sync=(world, rank, dist))
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), **out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
