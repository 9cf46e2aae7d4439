"""CPU tests of the C++ host layer (libb2caffe.so): registry, solver schedules, bucket plan, batch division.
These mirror test_layer_factory.cpp / the schedule part of test_gradient_based_solver.cpp; no GPU needed."""
import numpy as np
import pytest

import oracle as o
from caffe_mpi_b200 import host_api as h


def test_layer_registry_knows_convolution():
    assert h.lib().b2h_registry_has(b"Convolution") == 1
    assert h.lib().b2h_registry_has(b"NoSuchLayer") == 0


POLICIES = [
    dict(lr_policy="fixed", base_lr=0.01),
    dict(lr_policy="step", base_lr=0.1, gamma=0.5, stepsize=10),
    dict(lr_policy="exp", base_lr=0.1, gamma=0.999),
    dict(lr_policy="inv", base_lr=0.01, gamma=1e-4, power=0.75),                     # examples/mnist/lenet_solver
    dict(lr_policy="poly", base_lr=0.001, power=2.0, max_iter=2400000),              # models/resnet50/solver.prototxt
    dict(lr_policy="sigmoid", base_lr=0.1, gamma=-0.01, stepsize=50),
    dict(lr_policy="fixed", base_lr=0.1, rampup_interval=20, rampup_lr=0.01),
]


@pytest.mark.parametrize("cfg", POLICIES, ids=[c["lr_policy"] + str(i) for i, c in enumerate(POLICIES)])
def test_learning_rate_matches_oracle(cfg):
    s = h.SGDSolver(**cfg)
    for it in (0, 1, 5, 19, 20, 37, 100, 1234):
        want = o.learning_rate(cfg["lr_policy"], it, cfg["base_lr"], cfg.get("gamma", 0.1), cfg.get("power", 1.0),
                               cfg.get("stepsize", 1), cfg.get("max_iter", 1), cfg.get("min_lr", 0.0), 0,
                               cfg.get("rampup_interval", 0), cfg.get("rampup_lr", 0.0))
        assert s.lr_at(it) == pytest.approx(want, rel=1e-6, abs=1e-12)


def test_multistep_policy():
    s = h.SGDSolver(lr_policy="multistep", base_lr=0.1, gamma=0.1, stepvalue=[10, 20])
    got = [s.lr_at(i) for i in (0, 9, 10, 15, 20, 25)]          # iterations visited in order, like Solver::Step
    assert got == pytest.approx([0.1, 0.1, 0.01, 0.01, 0.001, 0.001], rel=1e-5)


def test_unknown_policy_is_fatal():
    s = h.SGDSolver(lr_policy="bogus")
    with pytest.raises(h.HostError, match="Unknown learning rate policy"):
        s.lr_at(0)


def reference_bucket_walk(counts, reduce_buckets):
    """Independent restatement of Net::ReduceAndUpdate's bucketing (net.cpp:772-783,824-862) for ids arriving
    last-to-first."""
    even = lambda c: c + (c & 1)
    n = len(counts)
    space = sum(even(c) for c in counts)
    mppb = max(1, (n + 1) // reduce_buckets)
    bsc = int(np.float32(space + 1) / np.float32(n) * np.float32(mppb))
    out, lo, hi, rec = [], -1, -1, 0
    for pid in range(n - 1, -1, -1):
        if rec >= bsc or (lo != -1 and pid < lo - 1) or (hi != -1 and pid > hi + 1):
            out.append((lo, hi))
            lo = hi = pid
            rec = even(counts[pid])
        else:
            lo = pid if lo == -1 or pid < lo else lo
            hi = pid if hi == -1 or pid > hi else hi
            rec += even(counts[pid])
    if lo != -1:
        out.append((lo, hi))
    return out


@pytest.mark.parametrize("seed", range(5))
def test_bucket_plan_matches_reference_walk(seed):
    rng = np.random.default_rng(seed)
    counts = [int(c) for c in rng.integers(1, 300000, size=int(rng.integers(1, 170)))]
    for rb in (1, 2, 6, 50):
        plan = h.plan_buckets(counts, rb)
        assert [(f, t) for f, t, _, _ in plan] == reference_bucket_walk(counts, rb)
        # buckets tile the arena exactly: contiguous, even-padded, every param in exactly one bucket
        even = lambda c: c + (c & 1)
        offs = np.cumsum([0] + [even(c) for c in counts])
        seen = []
        for f, t, off, cnt in plan:
            assert off == offs[f] and cnt == offs[t + 1] - offs[f]
            seen += list(range(f, t + 1))
        assert sorted(seen) == list(range(len(counts)))


def test_resnet50_bucket_sizes():
    # 161 reduced blobs, 102.2 MB arena, reduce_buckets = 6 -> messages of roughly 1/6 of the arena
    from caffe_mpi_b200.shapes import RESNET50
    counts = []
    for (cnt, C, H, O, k, s, p, G, b) in RESNET50:
        for _ in range(cnt):
            counts += [O * (C // G) * k * k, O, O]       # conv W + BN scale + BN bias
    counts += [2048 * 1000, 1000]
    plan = h.plan_buckets(counts, 6)
    assert 4 <= len(plan) <= 12
    assert sum(c for *_, c in plan) == sum(c + (c & 1) for c in counts)


def test_divide_batch_size():
    # parallel.cpp:284-293: per-rank batch = total / solver_count rounded up
    assert h.divide_batch_size(512, 8) == 64
    assert h.divide_batch_size(100, 8) == 13
    assert h.divide_batch_size(32, 1) == 32
