"""Host-side multi-GPU logic on CPU: contiguous sample sharding and the single gradient reduce,
exercised with world_size=2 over gloo (the GPU path uses the same code over NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

from diffquantum_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 8, 8192, 8195):
        for w in (1, 2, 3, 4, 8):
            seen = []
            for r in range(w):
                lo, hi = sharding.shard_bounds(n, r, w)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [sharding.shard_bounds(n, r, w)[1] - sharding.shard_bounds(n, r, w)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def test_balanced_assignment_is_a_partition_and_balances():
    rng = np.random.RandomState(0)
    costs = rng.uniform(1000, 3000, size=64)
    for w in (1, 2, 4, 8):
        parts = sharding.balanced_assignment(costs, w)
        assert sorted(np.concatenate(parts).tolist()) == list(range(64))
        loads = [costs[p].sum() for p in parts]
        assert max(loads) - min(loads) <= costs.max()
        assert max(loads) / (costs.sum() / w) < 1.03 or w == 1


def test_single_process_estimator_is_plain_mean():
    fake = lambda coeff, s: np.stack([np.full(coeff.shape, float(x)) for x in s])
    est = sharding.ShardedEstimator(fake)
    coeff = np.zeros((3, 2))
    s = np.array([1.0, 2.0, 6.0])
    np.testing.assert_allclose(est.mean_gradient(coeff, s), np.full((3, 2), 3.0))
    np.testing.assert_allclose(est.per_sample_gradients(coeff, s)[:, 0, 0], s)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # per-sample "gradient" depends on the sample time only, so any split must give the same answer
        fake = lambda coeff, s: np.stack([np.outer(np.arange(1, 4), np.arange(2)) * x + x * x for x in s])
        est = sharding.ShardedEstimator(fake)
        coeff = np.zeros((3, 2))
        s = np.random.RandomState(0).uniform(size=7) * 2.0
        mean = est.mean_gradient(coeff, s)
        bal = sharding.ShardedEstimator(fake, cost=lambda x: 3.0 - x)      # cost-balanced shards: same mean
        np.testing.assert_allclose(bal.mean_gradient(coeff, s), mean, rtol=1e-14)
        assert len(bal.my_samples(s)) in (3, 4)
        every = est.per_sample_gradients(coeff, s)
        want = fake(coeff, s)
        np.testing.assert_allclose(mean, want.mean(axis=0), rtol=1e-14)
        np.testing.assert_array_equal(every, want)
        lo, hi = sharding.shard_bounds(7, rank, world)
        out.put((rank, lo, hi))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(2))
    assert got == [(0, 0, 4), (1, 4, 7)]
