"""Single state split over ranks on its high qubits (diffquantum_b200/distributed.py).

CPU tier: the layout bookkeeping (which qubit sits on which bit after each all-to-all, one exchange per
step) runs for real over gloo with world_size 2 and 4; the slice kernels are replaced by the NumPy
stand-in below (test infrastructure, same per-term semantics as oracle/restate.py).
GPU tier: the CUDA slice kernels on one rank, and on 2 ranks over NCCL when two GPUs are visible."""
import os
import socket
import sys

import numpy as np
import pytest

from diffquantum_b200.ising import IsingProblem
from diffquantum_b200 import distributed
from oracle import restate as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpySliceOps(object):
    """CPU stand-in for dq_slice_* (tests only)."""

    def alloc(self, n):
        import torch
        return torch.zeros(n, dtype=torch.complex128)

    def fill_uniform(self, psi, L, n):
        psi[:] = 2.0 ** (-0.5 * n)

    @staticmethod
    def _diag(L, high, pair_bits, vals, c0):
        g = (np.uint64(high) << np.uint64(L)) | np.arange(1 << L, dtype=np.uint64)
        d = np.full(1 << L, float(c0))
        for (a, b), v in zip(pair_bits, vals):
            par = ((g >> np.uint64(a)) ^ (g >> np.uint64(b))) & np.uint64(1)
            d += np.where(par == 1, -v, v)
        return d

    def phase(self, psi, L, high, n, pair_bits, angles):
        a = psi.numpy()
        a *= np.exp(-1j * self._diag(L, high, pair_bits, angles[1:], angles[0]))

    def rx(self, psi, L, bit, theta):
        a = psi.numpy().reshape(-1, 2, 1 << bit)
        c, s = np.cos(theta), np.sin(theta)
        x0, x1 = a[:, 0, :].copy(), a[:, 1, :].copy()
        a[:, 0, :] = c * x0 - 1j * s * x1
        a[:, 1, :] = c * x1 - 1j * s * x0

    def rx_many(self, psi, L, bits, thetas):
        """Stand-in of dq_slice_rx_many: the rotations commute, so any order is the fused pass."""
        assert len(set(bits)) == len(bits) and all(0 <= b < L for b in bits)
        self.rx_many_calls = getattr(self, "rx_many_calls", 0) + 1
        for b, th in zip(bits, thetas):
            self.rx(psi, L, b, th)

    def energy(self, psi, L, high, n, pair_bits, m_zz, m_const):
        a = psi.numpy()
        return float(np.sum(self._diag(L, high, pair_bits, m_zz, m_const) * (a.real ** 2 + a.imag ** 2)))

    def all_to_all(self, recv, send):
        import torch
        import torch.distributed as dist
        dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send))

    def all_reduce_scalar(self, x):
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def to_host(self, psi):
        return psi.numpy().copy()

    def from_host(self, psi, array):
        psi.numpy()[:] = array


def _check(n, rank, world, ops, device=0, steps_per=3):
    edges = R.random_regular_edges(n, seed=n)
    prob = IsingProblem.maxcut(n, edges)
    ref = R.maxcut_structured(n, edges)
    coeff = np.random.RandomState(n).normal(0, 1, [len(prob.terms), 6])
    st = distributed.DistributedState(prob, device=device, per_step=steps_per, ops=ops)
    st.fill_uniform()
    st.evolve(coeff, 0.2, 1.7)
    ns, dt, ts = R.step_grid(0.2, 1.7, steps_per)
    want = R.evolve_split_structured(ref, R.coef_table_plain(coeff, ref["omegas"], ref["T"], ts), dt, ref["psi0"])
    assert st.exchanges == (ns if world > 1 else 0)                     # one all-to-all per step, no more
    if hasattr(st.ops, "rx_many_calls"):                                # fused rotations: local set, then the swapped-in qubits
        assert st.ops.rx_many_calls == ns * (2 if world > 1 else 1)
    e = st.energy()                                                     # layout-agnostic: no restore needed
    assert abs(e - R.energy_diag(ref["m_diag"], want)) < 1e-10
    assert abs(st.norm2() - 1) < 1e-12
    mine = st.local_slice()
    N = 1 << (n - (world.bit_length() - 1))
    err = np.abs(mine - want[rank * N:(rank + 1) * N]).max() / np.abs(want).max()
    assert err < 1e-10, err
    # a second leg from a non-uniform state loaded in the reference order
    rng = np.random.RandomState(1)
    psi0 = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi0 /= np.linalg.norm(psi0)
    st.set_state(psi0)
    rows = prob.trajectory_rows(coeff, 0.0, 0.35, steps_per)
    st.evolve_rows(rows)
    ns, dt, ts = R.step_grid(0.0, 0.35, steps_per)
    want = R.evolve_split_structured(ref, R.coef_table_plain(coeff, ref["omegas"], ref["T"], ts), dt, psi0)
    err = np.abs(st.local_slice() - want[rank * N:(rank + 1) * N]).max() / np.abs(want).max()
    assert err < 1e-10, err


def _cpu_worker(rank, world, port, n):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _check(n, rank, world, NumpySliceOps())
    finally:
        dist.destroy_process_group()


def _spawn(target, world, *args):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=target, args=(r, world, port) + args) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0


@pytest.mark.parametrize("world,n", [(2, 6), (4, 8)])
def test_layout_bookkeeping_over_gloo(world, n):
    _spawn(_cpu_worker, world, n)


def test_single_rank_bookkeeping_numpy():
    _check(6, 0, 1, NumpySliceOps())


def test_bad_world_sizes_are_rejected():
    prob = IsingProblem.maxcut(4, [[0, 1], [0, 3], [1, 2], [2, 3]])
    st = distributed.DistributedState(prob, ops=NumpySliceOps())
    assert st.g == 0 and st.L == 4 and st.global_qubits() == []


@pytest.mark.gpu
@pytest.mark.parametrize("n", [4, 10, 16])
def test_cuda_slice_kernels_single_rank(n):
    _check(n, 0, 1, None, steps_per=2)


@pytest.mark.gpu
@pytest.mark.parametrize("L,bits", [(3, [0, 2]), (5, [4, 1, 0, 3, 2]), (12, list(range(12))), (14, [13, 3, 12, 0, 7]),
                                    (17, list(range(17))), (20, [19, 5, 12, 18, 13, 0, 1, 17, 14, 15, 16]), (21, [20])])
def test_fused_rotation_pass_equals_one_kernel_per_rotation(L, bits):
    """dq_slice_rx_many (up to 12 bits per read + write of the slice, any bit set) against dq_slice_rx term by term and,
    for small slices, against the NumPy stand-in."""
    import torch
    ops = distributed.CudaSliceOps(0)
    rng = np.random.RandomState(L)
    host = rng.normal(size=1 << L) + 1j * rng.normal(size=1 << L)
    host /= np.linalg.norm(host)
    thetas = rng.uniform(-1.5, 1.5, size=len(bits))
    if L in (14, 20):
        thetas[0] = np.pi / 2                   # cos = 0: that pass must take the unscaled form of the rotation
    a, b = ops.alloc(1 << L), ops.alloc(1 << L)
    ops.from_host(a, host)
    ops.from_host(b, host)
    ops.rx_many(a, L, bits, thetas)
    for bit, th in zip(bits, thetas):
        ops.rx(b, L, bit, th)
    ops.ctx.synchronize()
    got, want = ops.to_host(a), ops.to_host(b)
    assert np.abs(got - want).max() < 1e-14
    assert abs(np.linalg.norm(got) - 1.0) < 1e-13
    if L <= 14:
        ref = torch.from_numpy(host.copy())
        cpu = NumpySliceOps()
        for bit, th in zip(bits, thetas):
            cpu.rx(ref, L, bit, th)
        assert np.abs(got - ref.numpy()).max() < 1e-14
    with pytest.raises(ValueError):
        ops.rx_many(a, L, [0, 0], [0.1, 0.2])
    with pytest.raises(ValueError):
        ops.rx_many(a, L, [L], [0.1])


@pytest.mark.gpu
@pytest.mark.parametrize("L,n,high", [(12, 12, 0), (16, 16, 0), (18, 20, 3), (21, 21, 0)])
def test_phase_fused_into_the_first_rotation_pass(L, n, high):
    """dq_slice_phase_rx_many (the step's diagonal phase applied inside the contiguous 12-bit rotation pass) against
    dq_slice_phase followed by dq_slice_rx_many; high rank bits included, and a bit list without low targets (no fusion)."""
    from oracle import restate as R
    ops = distributed.CudaSliceOps(0)
    rng = np.random.RandomState(100 + L)
    edges = R.random_regular_edges(n, seed=L) if n % 2 == 0 else [(i, (i + 1) % n) for i in range(n)] + [(i, (i + 5) % n) for i in range(0, n, 3)]
    pair_bits = np.array([[n - 1 - a, n - 1 - b] for a, b in edges], dtype=np.int32)
    angles = rng.normal(size=1 + len(edges))
    host = rng.normal(size=1 << L) + 1j * rng.normal(size=1 << L)
    host /= np.linalg.norm(host)
    for bits in (list(range(L)), [b for b in range(L) if b >= 12 or b % 3 == 0], [b for b in range(L) if b >= 12]):
        if not bits:
            continue
        thetas = rng.uniform(-1.2, 1.2, size=len(bits))
        a, b = ops.alloc(1 << L), ops.alloc(1 << L)
        ops.from_host(a, host)
        ops.from_host(b, host)
        l0 = ops.ctx.launch_count
        ops.phase_rx_many(a, L, high, n, pair_bits, angles, bits, thetas)
        fused_launches = ops.ctx.launch_count - l0
        l0 = ops.ctx.launch_count
        ops.phase(b, L, high, n, pair_bits, angles)
        ops.rx_many(b, L, bits, thetas)
        split_launches = ops.ctx.launch_count - l0
        ops.ctx.synchronize()
        got, want = ops.to_host(a), ops.to_host(b)
        assert np.abs(got - want).max() < 1e-13
        assert abs(np.linalg.norm(got) - 1.0) < 1e-12
        low_target = any(bit < 12 for bit in bits)
        assert fused_launches == split_launches - (1 if low_target else 0)


def _gpu_worker(rank, world, port, n):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        _check(n, rank, world, None, device=rank, steps_per=2)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [12, 18])
def test_cuda_two_ranks_over_nccl(n):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    _spawn(_gpu_worker, 2, n)
