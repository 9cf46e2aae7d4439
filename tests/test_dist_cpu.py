"""Multi-rank logic on CPU (gloo, world_size 2): the data-parallel exchange must make N ranks x (batch/N) produce
the same update as one rank on the full batch -- the property the reference checks with real GPUs in
GradientBasedSolverTest::TestLeastSquaresUpdate (test_gradient_based_solver.cpp:471-509).  The compute legs use
the CPU oracle; what is under test is the host-side sharding / summation / 1/solver_count scaling / rank-0
broadcast protocol that bench.py and the C++ P2PSync follow on the GPU."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle as o
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prm_full = o.ConvParams.make(8, 3, 10, 10, 4, 3, 1, 1, 1, 1, True)
    rng = np.random.default_rng(1701)                      # same stream on every rank: the GLOBAL batch
    x = rng.standard_normal(prm_full.x_shape()).astype(np.float32)
    dy = rng.standard_normal(prm_full.y_shape()).astype(np.float32)
    # per-rank different initial weights; rank 0's are broadcast (P2PSync::on_start, parallel.cpp:208-227)
    wr = np.random.default_rng(100 + rank)
    w = wr.standard_normal(prm_full.w_shape()).astype(np.float32)
    b = wr.standard_normal(4).astype(np.float32)
    for t in (w, b):
        tt = torch.from_numpy(t)
        dist.broadcast(tt, src=0)
    # shard: rank r takes images [r*B, (r+1)*B), B = total / solver_count (parallel.cpp:284-293)
    B = prm_full.N // world
    prm = o.ConvParams.make(B, 3, 10, 10, 4, 3, 1, 1, 1, 1, True)
    xs, dys = x[rank * B:(rank + 1) * B], dy[rank * B:(rank + 1) * B]
    h_w, h_b = np.zeros_like(w), np.zeros_like(b)
    for it in range(3):
        dw, db, _ = o.conv_backward(prm, xs, w, dys * np.float32(1.0 / prm_full.N))   # loss normalised by global batch
        # even-padded contiguous diff arena, one in-place sum allreduce (parallel.cpp:245-253)
        arena = np.zeros(dw.size + (dw.size & 1) + db.size + (db.size & 1), np.float32)
        arena[:dw.size] = dw.reshape(-1)
        arena[dw.size + (dw.size & 1):dw.size + (dw.size & 1) + db.size] = db
        t = torch.from_numpy(arena)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        # reference: x 1/solver_count after the sum (net.cpp:910); with a loss normalised by the LOCAL batch that
        # averages the replicas.  Here the loss is already normalised by the global batch, so solver_count = 1.
        gw = arena[:dw.size].reshape(w.shape)
        gb = arena[dw.size + (dw.size & 1):dw.size + (dw.size & 1) + db.size]
        _, wf, hf = o.sgd_update(gw, w, h_w, 0.9, 0.1, 0.01)
        w, h_w = wf.reshape(w.shape), hf.reshape(w.shape)
        _, b, h_b = o.sgd_update(gb, b, h_b, 0.9, 0.1, 0.0)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), w=w, b=b)
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["w"], r1["w"]) and np.array_equal(r0["b"], r1["b"])    # replicas stay in sync
    # single-rank run on the full batch
    import oracle as o
    prm = o.ConvParams.make(8, 3, 10, 10, 4, 3, 1, 1, 1, 1, True)
    rng = np.random.default_rng(1701)
    x = rng.standard_normal(prm.x_shape()).astype(np.float32)
    dy = rng.standard_normal(prm.y_shape()).astype(np.float32)
    wr = np.random.default_rng(100)
    w = wr.standard_normal(prm.w_shape()).astype(np.float32)
    b = wr.standard_normal(4).astype(np.float32)
    h_w, h_b = np.zeros_like(w), np.zeros_like(b)
    for it in range(3):
        dw, db, _ = o.conv_backward(prm, x, w, dy * np.float32(1.0 / prm.N))
        _, wf, hf = o.sgd_update(dw, w, h_w, 0.9, 0.1, 0.01)
        w, h_w = wf.reshape(w.shape), hf.reshape(w.shape)
        _, b, h_b = o.sgd_update(db, b, h_b, 0.9, 0.1, 0.0)
    assert np.allclose(r0["w"], w, rtol=1e-5, atol=1e-6)
    assert np.allclose(r0["b"], b, rtol=1e-5, atol=1e-6)


# ---- input side: product code (C++ DataReader) under a real process group ---------------------------------------------------
def _reader_worker(rank, world, port, db_path, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from caffe_mpi_b200 import data_api
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, steps = 3, 5
    rd = data_api.DataReader(db_path, B, solver_count=dist.get_world_size(), solver_rank=dist.get_rank())
    mine = np.stack([rd.next()[2].astype(np.int64) for _ in range(steps)])            # [steps][B] record ids
    sums = np.stack([rd.next()[0].astype(np.int64).sum(axis=(1, 2, 3)) for _ in range(1)])
    rd.close()
    gathered = [torch.zeros(steps, B, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine))
    if rank == 0:
        np.save(os.path.join(out_dir, "ids.npy"), torch.stack(gathered).numpy())       # [world][steps][B]
    dist.destroy_process_group()
    assert sums.shape == (1, B)


def test_two_ranks_read_the_global_batch_one_rank_would(tmp_path):
    """Data-parallel input: rank r of S reads records [r*B, (r+1)*B) of every S*B-record cycle
    (data_reader.cpp:288-305), so the ranks' batches concatenated in rank order are exactly the batches one solver with
    batch S*B reads -- the input-side half of "N solvers x batch/N == one solver x batch"."""
    from caffe_mpi_b200 import data_api, lmdb_io
    world, B, steps = 2, 3, 5
    rng = np.random.default_rng(8)
    db = str(tmp_path / "db")
    lmdb_io.write_datum_lmdb(db, rng.integers(0, 256, (20, 1, 4, 4), dtype=np.uint8), rng.integers(0, 10, 20))
    mp.spawn(_reader_worker, args=(world, _free_port(), db, str(tmp_path)), nprocs=world, join=True)
    ids = np.load(tmp_path / "ids.npy")                                               # [world][steps][B]
    one = data_api.DataReader(db, world * B)
    for k in range(steps):
        want = one.next()[2].astype(np.int64)
        assert np.array_equal(np.concatenate([ids[r, k] for r in range(world)]), want)
    one.close()
