"""Multi-rank CUDA parity, visible on ANY box (SURVEY 4 iv): W = 2, 4, 8 ranks are spawned here (no torchrun).

When the box has >= W GPUs every rank takes its own device and the exchange is NCCL over NVLink (the product path,
diffquantum_b200/distributed.py: CudaSliceOps).  On a box with fewer GPUs the ranks SHARE devices: the CUDA slice /
gradient kernels and the layout bookkeeping are exactly the same, only the all-to-all and the all-reduce are staged
through the host over gloo (NCCL refuses two ranks on one device).  Either way:
  (i)  DistributedState (one state split on its high qubits, one exchange per step: all-to-all, or fused into the last local
       pass's stores through peer memory) at n = 16, 18 against the oracle
       (oracle/restate.py evolve_split_structured = diffqc.cc:155-164), amplitudes and energy to 1e-10;
  (ii) ShardedEstimator.per_sample_gradients with the real IsingSimulator: the W-rank result is BIT-equal to the
       same samples computed by one rank alone.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _host_exchange_ops(device):
    from diffquantum_b200 import distributed

    class HostExchangeOps(distributed.CudaSliceOps):
        """CudaSliceOps with the two collectives staged through the host over gloo (ranks sharing a device)."""

        def all_to_all(self, recv, send):
            import torch
            import torch.distributed as dist
            self.ctx.synchronize()
            h_send = torch.view_as_real(send).cpu()
            h_recv = torch.empty_like(h_send)
            dist.all_to_all_single(h_recv, h_send)
            torch.view_as_real(recv).copy_(h_recv)
            torch.cuda.current_stream(self.device).synchronize()

        def all_reduce_scalar(self, x):
            import torch
            import torch.distributed as dist
            t = torch.tensor([x], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t.item())

    return HostExchangeOps(device)


def _init(rank, world, port):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    n_dev = torch.cuda.device_count()
    device = rank % n_dev
    torch.cuda.set_device(device)
    nccl = n_dev >= world
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return device, nccl


def _state_worker(rank, world, port, n, peer, out):
    device, nccl = _init(rank, world, port)
    import torch.distributed as dist
    try:
        from diffquantum_b200.ising import IsingProblem
        from diffquantum_b200 import distributed
        from oracle import restate as R
        edges = R.random_regular_edges(n, seed=n)
        prob = IsingProblem.maxcut(n, edges)
        ref = R.maxcut_structured(n, edges)
        coeff = np.random.RandomState(n).normal(0, 1, [len(prob.terms), 6])
        ops = None if nccl else _host_exchange_ops(device)
        st = distributed.DistributedState(prob, device=device, per_step=2, ops=ops, peer_exchange=peer)
        assert st.peer_exchange == bool(peer)                      # CUDA IPC mappings of every rank's buffers, or none
        st.fill_uniform()
        st.evolve(coeff, 0.2, 1.7)
        ns, dt, ts = R.step_grid(0.2, 1.7, 2)
        want = R.evolve_split_structured(ref, R.coef_table_plain(coeff, ref["omegas"], ref["T"], ts), dt, ref["psi0"])
        assert st.exchanges == ns                                   # one all-to-all per step
        e = st.energy()
        e_ref = R.energy_diag(ref["m_diag"], want)
        assert abs(e - e_ref) < 1e-10 * abs(e_ref), (e, e_ref)
        assert abs(st.norm2() - 1) < 1e-12
        N = 1 << (n - (world.bit_length() - 1))
        err = np.abs(st.local_slice() - want[rank * N:(rank + 1) * N]).max() / np.abs(want).max()
        assert err < 1e-10, err
        out.put((rank, "nccl" if nccl else "gloo+shared-device", float(err)))
    finally:
        dist.destroy_process_group()


def _grad_worker(rank, world, port, n, out):
    device, nccl = _init(rank, world, port)
    import torch.distributed as dist
    try:
        import diffquantum_b200 as dq
        from diffquantum_b200 import sharding
        from oracle import restate as R
        edges = R.random_regular_edges(n, seed=3)
        prob = dq.IsingProblem.maxcut(n, edges)
        coeff = np.random.RandomState(5).normal(0, 1, [len(prob.terms), 6])
        sim = dq.IsingSimulator(prob, device=device, per_step=2)
        assert sim.info("engine") == 1                              # the fused engine, as in the bench
        s = np.random.RandomState(8).uniform(size=2 * world + 1) * prob.T      # ragged: not a multiple of W
        est = sharding.ShardedEstimator(lambda c, ss: sim.grad_samples(c, ss), device=device)
        every = est.per_sample_gradients(coeff, s)
        alone = sim.grad_samples(coeff, s)                          # all samples on this rank's device alone
        np.testing.assert_array_equal(every, alone)
        mean = est.mean_gradient(coeff, s)
        assert np.abs(mean - alone.mean(axis=0)).max() < 1e-14 * max(1.0, np.abs(alone).max())
        if rank == 0:                                               # and against the oracle for one of them
            g_ref = R.grad_mc_structured(R.maxcut_structured(n, edges), coeff, float(s[1]), 2, mode="split")
            assert np.abs(every[1] - g_ref).max() / np.abs(g_ref).max() < 1e-10
        out.put((rank, "nccl" if nccl else "gloo+shared-device", 0.0))
    finally:
        dist.destroy_process_group()


def _spawn(target, world, *args):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, world, port) + args + (out,)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(world))
    assert [g[0] for g in got] == list(range(world))
    return got


@pytest.mark.parametrize("peer", [False, True])
@pytest.mark.parametrize("world,n", [(2, 16), (4, 18), (8, 18)])
def test_distributed_state_cuda_vs_oracle(world, n, peer):
    """peer=True: the exchange is fused into the stores of the last local rotation pass (peer memory through CUDA IPC, NVLink
    when every rank has its own GPU) and a barrier; peer=False: NCCL (or host-staged) all-to-all."""
    _spawn(_state_worker, world, n, peer)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_per_sample_gradients_bit_equal_to_one_rank(world):
    _spawn(_grad_worker, world, 14)
