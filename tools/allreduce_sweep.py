"""In-place fp32 sum-allreduce bus bandwidth through b2c_comm_* for the two gradient-arena sizes of the BASELINE configs
(ResNet-50 102 MB, VGG-16 553 MB), over communicator CTA caps and plain vs. ncclMemAlloc'ed + registered buffers.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/allreduce_sweep.py"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import caffe_mpi_b200 as m
from caffe_mpi_b200 import capi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
L = m.lib()
L.b2c_comm_mem_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
L.b2c_comm_mem_free.argtypes = [C.c_void_p]
L.b2c_comm_register.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
L.b2c_comm_allreduce_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
sizes = {"resnet50_arena": 25_610_152, "vgg16_arena": 138_357_544}
out = []
st = torch.cuda.Stream()
for ctas in (None, 8, 16, 32):
    if ctas is None:
        os.environ.pop("B2C_NCCL_MAX_CTAS", None)
    else:
        os.environ["B2C_NCCL_MAX_CTAS"] = str(ctas)
    ids = [capi.Comm.get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = capi.Comm(world, rank, ids[0])
    for name, n in sizes.items():
        for registered in (False, True):
            if registered:
                p = C.c_void_p()
                if L.b2c_comm_mem_alloc(C.byref(p), n * 4) != 0:
                    continue
                if L.b2c_comm_register(comm._h, p, n * 4) != 0:
                    L.b2c_comm_mem_free(p); continue
                ptr = p
            else:
                buf = torch.zeros(n, device="cuda")
                ptr = C.c_void_p(buf.data_ptr())
            sp = C.c_void_p(st.cuda_stream)
            for _ in range(3):
                L.b2c_comm_allreduce_sum(comm._h, ptr, n, sp)
            st.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for _ in range(10):
                L.b2c_comm_allreduce_sum(comm._h, ptr, n, sp)
            b.record(st)
            st.synchronize()
            t = torch.tensor([a.elapsed_time(b) / 10], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            alg = n * 4 / (ms / 1e3) / 1e9
            out.append(dict(max_ctas=ctas, buffer=name, bytes=n * 4, registered=registered, ms=round(ms, 4), algbw_gbs=round(alg, 1),
                            busbw_gbs=round(alg * 2 * (world - 1) / world, 1)))
            if rank == 0:
                print(json.dumps(out[-1]), flush=True)
            if registered:
                pass   # freed with the process: deregistration happens in b2c_comm_destroy
    comm.destroy()
dist.barrier()
dist.destroy_process_group()
