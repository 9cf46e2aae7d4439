#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Caffe-MPI hot path.

Metric (BASELINE.json): images/sec, ResNet-50 fp32 training, batch 64 per GPU, synthetic ImageNet-shaped
input.  One "step" = one pass of the hot path over one batch: every ConvolutionLayer of
models/resnet50/train_val.prototxt forward, then backward (weight + bias gradient and bottom gradient,
no bottom gradient for conv1), the gradient allreduce over the contiguous diff arena (N > 1) and the fused
SGD-momentum update of all 25.56 M learnable parameters.  `config.workload` says exactly what is inside
the timed region.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the reference's CPU path on the host cores)

Prints ONE JSON line on rank 0 (see the key list in DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "resnet50"
PER_GPU_BATCH = 64
METRIC = "images/sec ResNet-50 fp32 train (conv fwd+bwd + grad allreduce + SGD hot path)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------- reference arm
def reference_sample_runner():
    """Returns (run(n_images) -> seconds, cores, kind, description).  The reference's CPU ConvolutionLayer
    path: verbatim im2col.cpp (oracle/_ref) + per-image/per-group loop + OpenBLAS sgemm on all host cores;
    falls back to the plain-C oracle port when oracle/_ref or OpenBLAS is unavailable."""
    import numpy as np
    import oracle as o
    from caffe_mpi_b200.shapes import MODELS
    cores = os.cpu_count() or 1
    use_ref = o.ref() is not None and o.ref_blas_open(cores)
    state = {"threads": cores}
    rng = np.random.default_rng(1701)
    layers = MODELS[MODEL]

    def run(nimg):
        t = 0.0
        first = True
        for (cnt, C, H, O, k, s, p, G, bias) in layers:
            prm = o.ConvParams.make(nimg, C, H, H, O, k, s, p, 1, G, bias)
            x = rng.standard_normal(prm.x_shape(), dtype=np.float32)
            w = rng.standard_normal(prm.w_shape(), dtype=np.float32) * np.float32((2.0 / prm.Kd) ** 0.5)
            dy = rng.standard_normal(prm.y_shape(), dtype=np.float32)
            b = np.zeros(prm.O, np.float32) if bias else None
            y = np.empty(prm.y_shape(), np.float32)
            dw = np.zeros(prm.w_shape(), np.float32)
            db = np.zeros(prm.O, np.float32) if bias else None
            dx = None if first else np.empty(prm.x_shape(), np.float32)
            t0 = time.perf_counter()
            if use_ref:
                o.ref_conv_fwd_bwd(prm, x, w, b, y=y, dy=dy, dw=dw, db=db, dx=dx)
            else:
                o.conv_forward(prm, x, w, b)
                o.conv_backward(prm, x, w, dy, want_dx=not first)
            dt = time.perf_counter() - t0
            t += dt * cnt            # identical layers are timed once and counted `cnt` times
            first = False
        # SGD update of all learnable params on the host (CPU branch of ComputeUpdateValue)
        n = 25557032
        g = np.zeros(n, np.float32); w_ = np.zeros(n, np.float32); h = np.zeros(n, np.float32)
        t0 = time.perf_counter()
        o.lib().b2o_sgd_update(n, g, w_, h, 0.9, 0.001, 1e-4, 1, 1.0, 1, 1)
        t += time.perf_counter() - t0
        return t

    if use_ref:
        # the reference's OpenBLAS uses every core by default, which is far from optimal for its per-image GEMMs on a
        # many-core host; give the CPU arm its best thread count out of {all, 64, 32, 16, 8} (1-image calibration)
        best = None
        for th in sorted({cores, 64, 32, 16, 8} & set(range(1, cores + 1)), reverse=True):
            o.ref_blas_open(th)
            run(1)
            t = run(1)
            if best is None or t < best[0]:
                best = (t, th)
        state["threads"] = best[1]
        o.ref_blas_open(best[1])
    kind = "port"   # the conv loop is a restatement; only im2col.cpp is the reference's own object code
    desc = ("ResNet-50 conv stack fwd+bwd (+ host SGD), reference im2col.cpp verbatim + OpenBLAS sgemm per image/group"
            if use_ref else "ResNet-50 conv stack fwd+bwd, plain-C oracle port (single thread)")
    if use_ref:
        desc += f" ({state['threads']} OpenBLAS threads = best of calibration, host has {cores} cores)"
    return run, (state["threads"] if use_ref else 1), kind, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, cores, kind, desc = reference_sample_runner()
    run(1)                      # cold start (BLAS threads, page faults) is not the calibration
    t1 = run(1)
    S = max(1, min(8, int(4.0 / max(t1, 1e-3))))
    for _ in range(args.warmup):
        run(S)
    t0 = time.perf_counter()
    tt = 0.0
    for _ in range(args.steps):
        tt += run(S)
    wall = time.perf_counter() - t0
    ips = S * args.steps / tt
    out = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL} conv fwd+bwd + SGD, CPU, {S} images per step (bounded sample)",
                   "per_gpu_batch": PER_GPU_BATCH},
        "cpu_baseline": {"value": ips, "unit": "images/sec", "cores": cores, "kind": kind,
                         "sample": f"{desc}; {S} images/step x {args.steps} steps; wall {wall:.1f}s"},
        "e2e": {"value": ips, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------- our arm
class ClockSampler:
    def __init__(self, idx):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import caffe_mpi_b200 as m
    from caffe_mpi_b200 import capi
    from caffe_mpi_b200.shapes import MODELS, EXTRA_PARAMS, conv_flops_per_image

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ids = [capi.Comm.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = capi.Comm(world, rank, ids[0])
    L = m.lib()
    N = PER_GPU_BATCH
    layers = MODELS[MODEL]
    g = torch.Generator(device=dev).manual_seed(1701 + rank)     # seed + rank, parallel.cpp:179-187

    # ---- parameters: one contiguous arena each for data / diff / history (net.cpp:1350-1373), even-padded slots
    convs = []
    off = 0
    segs = []
    for (cnt, C, H, O, k, s, p, G, bias) in layers:
        for _ in range(cnt):
            prm = capi.ConvParams.make(N, C, H, H, O, k, s, p, 1, G, bias)
            nW = O * (C // G) * k * k
            convs.append(dict(prm=prm, w_off=off, nW=nW, b_off=None))
            segs.append((off, nW)); off += nW + (nW & 1)
            if bias:
                convs[-1]["b_off"] = off
                segs.append((off, O)); off += O + (O & 1)
    extra = EXTRA_PARAMS[MODEL]                 # fc + BN scale/bias: updated and allreduced, no conv math
    segs.append((off, extra)); off += extra + (extra & 1)
    arena_n = off
    Wd = torch.empty(arena_n, device=dev)
    Wd.normal_(0, 0.02, generator=g)
    Gd = torch.zeros(arena_n, device=dev)
    Hd = torch.zeros(arena_n, device=dev)
    if comm is not None:
        comm.bcast(Wd, 0)                       # P2PSync::on_start, parallel.cpp:208-227
    # ---- activations (synthetic, resident): per-layer bottoms; shared top / top-diff / bottom-diff scratch
    max_y = max(c["prm"].N * c["prm"].O * c["prm"].Ho * c["prm"].Wo for c in convs)
    max_x = max(c["prm"].N * c["prm"].C * c["prm"].H * c["prm"].W for c in convs)
    Y = torch.empty(max_y, device=dev)
    DY = torch.empty(max_y, device=dev).normal_(0, 1, generator=g)
    DX = torch.empty(max_x, device=dev)
    for c in convs:
        prm = c["prm"]
        c["x"] = torch.empty(prm.x_shape(), device=dev).normal_(0, 1, generator=g)
        c["desc"] = m.ConvDesc(prm, capi.ENGINE_DEFAULT, math=capi.MATH_TF32 if args.math == "tf32" else capi.MATH_FP32)
        c["w"] = Wd[c["w_off"]:c["w_off"] + c["nW"]]
        c["dw"] = Gd[c["w_off"]:c["w_off"] + c["nW"]]
        c["b"] = Wd[c["b_off"]:c["b_off"] + prm.O] if c["b_off"] is not None else None
        c["db"] = Gd[c["b_off"]:c["b_off"] + prm.O] if c["b_off"] is not None else None
    host_in = torch.empty(convs[0]["prm"].x_shape(), pin_memory=True).normal_(0, 1)
    host_out = torch.empty(1, pin_memory=True)
    lr, momentum, wd = 0.001, 0.9, 1e-4       # models/resnet50/solver.prototxt
    offs = [s_[0] for s_ in segs]; cnts = [s_[1] for s_ in segs]
    rates = [lr] * len(segs); decays = [wd] * len(segs)
    comm_stream = torch.cuda.Stream(device=dev, priority=-1)
    ev_upd_done = torch.cuda.Event()
    timers = {"fwd": [], "wgrad": [], "dgrad": [], "sgd": []}
    # Net::ReduceAndUpdate's bucketing (net.cpp:824-862) over the arena, planned by the C++ host layer: params
    # become ready last-to-first during backward; a bucket is exchanged + updated as soon as its lowest id is done
    from caffe_mpi_b200 import host_api
    buckets = host_api.plan_buckets(cnts, 6) if world > 1 else [(0, len(segs) - 1, 0, arena_n)]
    seg_of_conv = {}            # conv index -> lowest segment id it owns
    si = 0
    for ci, c in enumerate(convs):
        seg_of_conv[ci] = si
        si += 2 if c["b_off"] is not None else 1
    bucket_events = [torch.cuda.Event() for _ in buckets]

    def flush_bucket(bi, cur):
        f, t, off_, cnt_ = buckets[bi]
        bucket_events[bi].record(cur)
        side = comm_stream if comm is not None else cur
        if comm is not None:
            comm_stream.wait_event(bucket_events[bi])
            comm.allreduce_sum(Gd[off_:off_ + cnt_], stream=comm_stream)
        capi.sgd_update_arena(offs[f:t + 1], cnts[f:t + 1], rates[f:t + 1], decays[f:t + 1], Gd, Wd, Hd, momentum, l2=True,
                              grad_scale=1.0 / world, clear_grads=True, stream=side)

    def step(e2e=False, timed=False):
        cur = torch.cuda.current_stream()
        if e2e:
            convs[0]["x"].copy_(host_in, non_blocking=True)
        cur.wait_event(ev_upd_done)            # weights of the previous iteration are final
        def rec(kind):
            if not timed:
                return None
            a = torch.cuda.Event(enable_timing=True); a.record(cur); return (kind, a)
        def end(tok):
            if tok is None:
                return
            b = torch.cuda.Event(enable_timing=True); b.record(cur); timers[tok[0]].append((tok[1], b))
        for c in convs:
            t = rec("fwd"); c["desc"].forward(c["x"], c["w"], c["b"], Y); end(t)
        nb = 0
        # the fc / BN parameters (last segment) have no conv math here: their (zero) diffs are ready at once
        while nb < len(buckets) and buckets[nb][0] >= len(segs) - 1 and world > 1:
            flush_bucket(nb, cur); nb += 1
        for i in range(len(convs) - 1, -1, -1):
            c = convs[i]
            t = rec("wgrad"); c["desc"].backward_filter(c["x"], DY, c["dw"]); end(t)
            if c["db"] is not None:
                c["desc"].backward_bias(DY, c["db"])
            if i > 0:                           # conv1's bottom is data: propagate_down = false (net.cpp:183-191)
                t = rec("dgrad"); c["desc"].backward_data(DY, c["w"], DX); end(t)
            if world > 1:
                while nb < len(buckets) and seg_of_conv[i] <= buckets[nb][0]:
                    flush_bucket(nb, cur); nb += 1
        if world > 1:
            while nb < len(buckets):
                flush_bucket(nb, cur); nb += 1
            ev_upd_done.record(comm_stream)
        else:
            t = rec("sgd"); flush_bucket(0, cur); end(t)
            ev_upd_done.record(cur)
        if e2e:
            cur.wait_event(ev_upd_done)
            host_out.copy_(Wd[:1], non_blocking=True)   # device->host read of a step result
            cur.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_run(e2e, timed_ops):
        for _ in range(args.warmup):
            step(e2e)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.b2c_launch_count()
        a.record()
        for _ in range(args.steps):
            step(e2e, timed=timed_ops)
        torch.cuda.current_stream().wait_event(ev_upd_done)
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        launches = L.b2c_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    def time_allreduce(reps=5):
        # standalone in-place sum-allreduce of the whole diff arena (nccl-tests convention for busbw)
        if comm is None:
            return None
        for _ in range(2):
            comm.allreduce_sum(Gd, stream=comm_stream)
        comm_stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(comm_stream)
        for _ in range(reps):
            comm.allreduce_sum(Gd, stream=comm_stream)
        b.record(comm_stream)
        comm_stream.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_ = float(t.item())
        Gd.zero_()
        algbw = arena_n * 4 / (ms_ / 1e3) / 1e9
        return {"bytes": arena_n * 4, "ms": ms_, "algbw_gbs": algbw, "busbw_gbs": algbw * 2 * (world - 1) / world,
                "nvlink_ref_gbs": 900.0}

    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = timed_run(False, False)        # headline: nothing but the step inside the timed region
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed_run(True, False)
    timed_run(False, True)                        # per-op breakdown for the roofline block (two events around every conv call)
    allreduce = time_allreduce()
    # whole-graph measurement: always at N = 1; at N > 1 only on request (the same data-parallel TrainNet path is
    # exercised and checked by tools/trainer_multi.py -- profiles/r01_fullnet_2gpu.json -- and kept out of the default
    # multi-rank bench so that nothing can stand between the scaling run and its headline line)
    full_net = None if (args.no_full_net or (world > 1 and not args.full_net_multi)) else full_net_measure(args, N, world, rank, dev)

    if rank == 0:
        pk = peaks()
        imgs = N * world * args.steps
        value = imgs / (ms / 1e3)
        e2e_v = imgs / (ms_e2e / 1e3)
        # dominant kernel: igemm_fwd_kernel (the persistent tcgen05 implicit-GEMM kernel; forward and dgrad launches,
        # 56% of the step's GPU time in profiles/r01_v3_launches.csv).  achieved = algorithmic FLOPs of those launches /
        # their CUDA-event time inside the timed steps (events bracket the filter-prep + main kernel of each call).
        fl_layer = [(2 * c["prm"].N * c["prm"].O * c["prm"].Kd * c["prm"].Ho * c["prm"].Wo) for c in convs]
        fl_fwd = sum(fl_layer)
        fl_dgrad = sum(fl_layer[1:])
        fl = conv_flops_per_image(layers) * N
        per = {k: sum(a.elapsed_time(b) for a, b in timers[k]) / args.steps for k in timers}
        conv_ms = per["fwd"] + per["wgrad"] + per["dgrad"]
        tf32_peak = pk["bf16_tflops"] / 2.0      # TF32 dense = half the bf16 rate on the same tensor pipe
        ach = (fl_fwd + fl_dgrad) / ((per["fwd"] + per["dgrad"]) / 1e3) / 1e12
        ach_all = fl / (conv_ms / 1e3) / 1e12
        algo = sorted(set(c["desc"].algo_used(op) for c in convs for op in (0, 1, 2)))
        out = {
            "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{MODEL}: 53 conv layers fwd+bwd (N={N}/GPU, 224x224) + "
                                   f"{'bucketed NCCL allreduce of the %.1f MB diff arena overlapped with backward + ' % (arena_n * 4 / 1e6) if world > 1 else ''}"
                                   "fused SGD update of 25.56M params; non-conv layers not in the timed region",
                       "per_gpu_batch": N, "global_batch": N * world, "parallelism": f"dp{world}",
                       "math": "fp32-equivalent (3xTF32 tcgen05 MMA)" if args.math == "fp32" else "tf32 (single-pass, ~3e-4 per-layer error; informational)", "algos_used": algo,
                       "l2_policy": "per-step working set (activations ~2.4 GB) exceeds the 126 MB L2"},
            "gpu_launches": launches,
            "e2e": {"value": e2e_v, "unit": "images/sec", "h2d_bytes_per_step": host_in.numel() * 4 * world,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "clocks": clocks,
            "allreduce": allreduce,
            "roofline": {"bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the 105 launches of one step
                         # (profiles/r01_final_launches.csv, ncu); algorithmic bytes per launch 100.5e6 -> no wasted re-reads
                         "traffic": 72.2e6 if (args.math == "fp32" and N == 64) else None,
                         "traffic_unit": "bytes per launch (mean; algorithmic 100.5e6; ncu, profiles/r01_final_launches.csv)",
                         "kernel": "igemm_fwd_kernel (tcgen05 implicit GEMM: 53 forward + 52 dgrad launches/step)",
                         "peak_source": pk["source"] + ": bf16_tflops/2 (TF32 dense); fp32-equivalent mode issues 3 TF32 MMAs per "
                                        "algorithmic MAC, so its ceiling is peak/3",
                         "frac_of_3xtf32_ceiling": ach / (tf32_peak / 3.0) if args.math == "fp32" else None,
                         "all_conv_kernels": {"achieved": ach_all, "frac": ach_all / tf32_peak},
                         "ms_per_step": {k: round(v, 3) for k, v in per.items()}},
        }
        if full_net is not None:
            out["full_net"] = full_net
        if world == 1:
            run, cores, kind, desc = reference_sample_runner()
            run(1)
            t1 = run(1)
            S = max(1, min(8, int(10.0 / max(t1, 1e-3))))
            t = run(S) if S > 1 else t1
            out["cpu_baseline"] = {"value": S / t, "unit": "images/sec", "cores": cores, "kind": kind,
                                   "sample": f"{desc}; {S} images once ({t:.1f}s)"}
        print(json.dumps(out))
    if comm is not None:
        comm.destroy()
        dist.destroy_process_group()


def full_net_measure(args, N, world=1, rank=0, dev=None):
    """Supplementary line: the WHOLE models/resnet50/train_val.prototxt graph -- conv, BatchNorm,
    ReLU, pooling, Eltwise, InnerProduct, SoftmaxWithLoss forward + backward, bucketed gradient allreduce through the
    C++ P2PSync / ReduceScheduler (N > 1) and the SGD update -- through caffe::TrainNet (host/train_net.cpp), timed with
    CUDA events on the net's own stream, max over ranks.  Not the headline while the non-conv kernels are first-cut;
    reported so the distance between the hot path and the full training step is on record."""
    try:
        from caffe_mpi_b200 import capi, host_api, models
        import torch
        import torch.distributed as dist
        torch.cuda.empty_cache()
        t = host_api.Trainer(models.resnet50_prototxt(N), models.RESNET50_SOLVER, batch=N, seed=1701 + rank,
                             math=capi.MATH_TF32 if args.math == "tf32" else capi.MATH_FP32)
        if world > 1:
            ids = [t.new_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            t.attach_sync(world, rank, ids[0])

        def mx(ms):
            if world == 1:
                return ms
            v = torch.tensor([ms], device=dev)
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            return float(v.item())

        t.step(args.warmup)
        t.sync()
        if world > 1:
            dist.barrier()
        ms = mx(t.timed_steps(args.steps))
        loss = t.loss()
        t.step(1, copy_input=True)
        t.sync()
        if world > 1:
            dist.barrier()
        ms_e2e = mx(t.timed_steps(args.steps, copy_input=True, read_loss=True))
        imgs = N * world * args.steps
        return {"metric": "images/sec ResNet-50 fp32 train, full prototxt graph (all layers fwd+bwd"
                          + (" + bucketed allreduce" if world > 1 else "") + " + SGD)",
                "value": imgs / (ms / 1e3), "unit": "images/sec", "ms_per_step": ms / args.steps, "n_gpus": world,
                "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "images/sec", "h2d_bytes_per_step": t.input_bytes() * world,
                        "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
                "loss_after_warmup": loss, "learnable_blobs": t.num_params(), "activation_floats": t.activation_floats(),
                "accuracy_layers": "skipped (no gradient, not on the training path)"}
    except Exception as e:                       # supplementary measurement: report, never mask
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default="fp32", choices=["fp32", "tf32"],
                    help="fp32 = 3xTF32 split (fp32-equivalent results, the headline); tf32 = single-pass TF32 (informational)")
    ap.add_argument("--no-full-net", action="store_true", help="skip the supplementary full-prototxt-graph measurement")
    ap.add_argument("--full-net-multi", action="store_true", help="also run the full-graph measurement when N > 1 (P2PSync / ReduceScheduler)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
