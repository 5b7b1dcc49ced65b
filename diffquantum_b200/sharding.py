"""Sample sharding across the GPUs of one box (one process per GPU, torch.distributed).

The stochastic parameter-shift estimator (sim_plain.py:156-231) draws independent times s; every
sample is an independent set of 1 + 2*n_Hs trajectories.  So the only multi-GPU structure the path
needs is: split the sample list over ranks, run each shard on the local device with no traffic at
all, and sum the [n_Hs, n_basis] gradient once at the end (NCCL all-reduce over NVLink; gloo on CPU
for the host-logic tests).  There is no data-path collective.
"""
import numpy as np


def shard_bounds(n_items, rank, world_size):
    """Contiguous, balanced split: the first (n_items % world_size) ranks get one extra item."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(items, rank, world_size):
    lo, hi = shard_bounds(len(items), rank, world_size)
    return items[lo:hi]


def balanced_assignment(costs, world_size):
    """Cost-balanced split of independent items (longest-processing-time greedy, deterministic):
    returns one sorted index array per rank.  Sample cost is known up front — a sample at time s runs
    int(per_step (s+1)) + 2 n_Hs int(per_step (T-s+1)) trajectory-steps (sim_plain.py:123,190-215) and
    varies 3x with s — so balancing by cost instead of by count removes most of the idle tail of the
    slowest rank."""
    costs = np.asarray(costs, dtype=np.float64).reshape(-1)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world_size)
    out = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))             # first minimum: deterministic on every rank
        out[r].append(int(i))
        load[r] += costs[i]
    return [np.array(sorted(x), dtype=np.int64) for x in out]


def dist_info():
    """(rank, world_size, initialised) of the default process group; (0, 1, False) without one."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), True
    except ImportError:
        pass
    return 0, 1, False


def all_reduce_sum(array, device=None):
    """Sum a float64 numpy array over all ranks of the default group; returns a new array.
    With the nccl backend the buffer is staged on `device` (the reduce runs over NVLink);
    with gloo it stays on the host."""
    rank, world, ok = dist_info()
    a = np.ascontiguousarray(array, dtype=np.float64)
    if not ok or world == 1:
        return a.copy()
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(a.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


def all_gather_rows(array, counts, device=None):
    """Concatenate per-rank row blocks (rank r contributes counts[r] rows) in rank order."""
    rank, world, ok = dist_info()
    a = np.ascontiguousarray(array, dtype=np.float64)
    if not ok or world == 1:
        return a.copy()
    import torch
    import torch.distributed as dist
    width = int(np.prod(a.shape[1:])) if a.ndim > 1 else 1
    pad = max(counts)
    buf = np.zeros((pad, width))
    buf[:a.shape[0]] = a.reshape(a.shape[0], width)
    t = torch.from_numpy(buf)
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = [o.cpu().numpy()[:counts[r]] for r, o in enumerate(outs)]
    return np.concatenate(parts, axis=0).reshape((-1,) + a.shape[1:])


class ShardedEstimator(object):
    """K-sample gradient average over all ranks.  `local_grads(coeff, s_shard)` is any callable
    returning per-sample gradients [len(s_shard), n_Hs, n_basis] — normally
    IsingSimulator.grad_samples bound to this rank's GPU."""

    def __init__(self, local_grads, device=None, cost=None):
        self.local_grads = local_grads
        self.device = device
        self.cost = cost                # optional: callable s -> relative cost, enables cost-balanced shards

    def my_samples(self, s_list):
        rank, world, _ = dist_info()
        s_list = np.asarray(s_list, dtype=np.float64).reshape(-1)
        if self.cost is None or world == 1:
            return shard(s_list, rank, world)
        return s_list[balanced_assignment([self.cost(s) for s in s_list], world)[rank]]

    def mean_gradient(self, coeff, s_list):
        rank, world, _ = dist_info()
        s_list = np.asarray(s_list, dtype=np.float64).reshape(-1)
        mine = self.my_samples(s_list)
        shape = np.asarray(coeff).shape
        local = np.zeros(shape)
        if len(mine):
            local = np.asarray(self.local_grads(coeff, mine)).sum(axis=0)
        return all_reduce_sum(local, self.device) / max(1, len(s_list))

    def per_sample_gradients(self, coeff, s_list):
        """All per-sample gradients in sample order on every rank (parity checks: 1 vs N GPUs)."""
        rank, world, _ = dist_info()
        s_list = np.asarray(s_list, dtype=np.float64).reshape(-1)
        mine = shard(s_list, rank, world)
        shape = tuple(np.asarray(coeff).shape)
        local = np.zeros((0,) + shape)
        if len(mine):
            local = np.asarray(self.local_grads(coeff, mine)).reshape((len(mine),) + shape)
        counts = [shard_bounds(len(s_list), r, world)[1] - shard_bounds(len(s_list), r, world)[0]
                  for r in range(world)]
        return all_gather_rows(local, counts, self.device)
