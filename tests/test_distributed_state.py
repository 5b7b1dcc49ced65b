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


class NumpySliceOpsStep(NumpySliceOps):
    """The stand-in with dq_slice_step: rotations owed to the previous step, phase, rotations in one call."""

    def step(self, psi, L, high, n, pair_bits, angles, pre_bits, pre_thetas, bits, thetas):
        assert len(set(pre_bits)) == len(pre_bits) and all(0 <= b < L for b in pre_bits)
        self.step_calls = getattr(self, "step_calls", 0) + 1
        self.step_pre = getattr(self, "step_pre", 0) + (1 if len(pre_bits) else 0)
        for b, th in zip(pre_bits, pre_thetas):
            self.rx(psi, L, b, th)
        self.phase(psi, L, high, n, pair_bits, angles)
        for b, th in zip(bits, thetas):
            self.rx(psi, L, b, th)


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
    if hasattr(st.ops, "step_calls"):                                   # owed rotations ride on the next step's first call
        assert st.ops.step_calls == ns and st.ops.step_pre == (ns - 1 if world > 1 else 0)
        assert getattr(st.ops, "rx_many_calls", 0) == (1 if world > 1 else 0)      # the flush after the last step
    elif hasattr(st.ops, "rx_many_calls"):                              # fused rotations: local set, then the swapped-in qubits
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
        _check(n, rank, world, NumpySliceOpsStep())
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
    _check(6, 0, 1, NumpySliceOpsStep())


def _ring_pairs(n):
    """ring + chords: degree <= 4, any n"""
    edges = [(i, (i + 1) % n) for i in range(n)] + [(i, (i + 5) % n) for i in range(0, n, 3)]
    return np.array([[n - 1 - a, n - 1 - b] for a, b in edges], dtype=np.int32)


@pytest.mark.parametrize("n,steps,sets", [(12, 4, 1), (14, 5, 2), (21, 6, 2), (22, 3, 3), (29, 6, 3), (30, 2, 3), (33, 2, 4)])
def test_chained_pass_plan(n, steps, sets):
    """The pass planner of dq_slice_evolve_steps (host logic, no device: dq_slice_plan): with the TMA tile kernel every step costs
    (tile sets - 1) passes, every bit is rotated exactly once per step -- by the step's own passes or, owed, by the first pass
    of the next step --, every step has exactly one phase and it comes before the step's rotations on the same tile."""
    plan = distributed.slice_plan(n, n, _ring_pairs(n), list(range(n)), steps)
    assert len(plan) == (steps * (sets - 1) + 1 if sets > 1 else steps)
    for k in range(steps):
        mine = [r for r in plan if r["step"] == k]
        assert sum(r["phase"] for r in mine) == 1 and mine[0]["phase"] == 1           # the step opens with its phase
        owed_to_next = sum(r["n_pre"] for r in plan if r["step"] == k + 1)
        assert sum(r["n_rot"] for r in mine) + owed_to_next == n
        assert mine[0]["n_pre"] == (0 if k == 0 else n - sum(r["n_rot"] for r in plan if r["step"] == k - 1))
        assert all(r["n_pre"] == 0 for r in mine[1:])
    for r in plan:
        assert r["T"] == 12 and r["lo"] >= 3 and bin(r["mask"]).count("1") == 12 and r["scatter"] == 0
        assert r["mask"] & ((1 << r["lo"]) - 1) == (1 << r["lo"]) - 1                 # the lo low bits are in the tile
        assert r["n_rot"] <= 12 - (0 if r["lo"] == 12 else r["lo"])
    # without the TMA kernel only the contiguous tile can carry the phase and nothing is chained
    plain = distributed.slice_plan(n, n, _ring_pairs(n), list(range(n)), steps, assume_tma=False)
    assert len(plain) == steps * sets and all(r["n_pre"] == 0 for r in plain)
    assert sum(r["phase"] for r in plain) == steps


def test_pass_plan_small_and_dense_cases():
    # below 12 bits no tile can carry the phase: phase pass + one rotation pass per step
    plan = distributed.slice_plan(10, 10, _ring_pairs(10), list(range(10)), 3)
    assert [(r["T"], r["phase"], r["n_rot"]) for r in plan] == [(0, 1, 0), (10, 0, 10)] * 3
    # a qubit with more pairs than the kernel's tables hold (17 > 16): the phase runs as its own pass, no chaining
    n = 20
    star = np.array([[0, b] for b in range(1, 18)], dtype=np.int32)
    plan = distributed.slice_plan(n, n, star, list(range(n)), 2)
    assert sum(r["phase"] for r in plan) == 2 and all(r["T"] == 0 for r in plan if r["phase"]) and all(r["n_pre"] == 0 for r in plan)
    # rotations on a subset of the bits: only the tiles that hold targets are launched
    plan = distributed.slice_plan(24, 24, _ring_pairs(24), [0, 5, 23], 2)
    assert sum(r["n_rot"] + r["n_pre"] for r in plan) == 6 and len(plan) == 3


def test_distributed_step_plan():
    """One step of the n = 32 state on 8 ranks (L = 29, the three qubits swapped in by the last exchange owed): three launches --
    the owed rotations, the phase and 8 rotations on the tile that holds the top bits; 9 rotations; the contiguous tile carrying
    the exchange.  Without owed rotations and without the TMA kernel the step degrades gracefully."""
    n, L, g = 32, 29, 3
    pairs = _ring_pairs(n)
    plan = distributed.slice_plan_step(L, n, pairs, [26, 27, 28], list(range(L)), scatter_g=g)
    assert [(r["n_pre"], r["phase"], r["n_rot"], r["scatter"]) for r in plan] == [(3, 1, 8, 0), (0, 0, 9, 0), (0, 0, 12, 1)]
    assert plan[0]["mask"] >> 21 == 0xff and plan[2]["lo"] == 12 and plan[2]["mask"] == 0xfff
    first = distributed.slice_plan_step(L, n, pairs, [], list(range(L)), scatter_g=g)           # the very first step: nothing owed
    assert len(first) == 3 and first[0]["phase"] == 1 and first[-1]["scatter"] == 1 and sum(r["n_rot"] for r in first) == L
    plain = distributed.slice_plan_step(L, n, pairs, [26, 27, 28], list(range(L)), scatter_g=g, assume_tma=False)
    # cp.async kernels only: the owed rotations as a pass of their own, the phase on the contiguous tile, the exchange last
    assert [(r["n_pre"], r["phase"], r["scatter"]) for r in plain] == [(0, 0, 0), (0, 1, 0), (0, 0, 0), (0, 0, 1)]
    assert plain[0]["n_rot"] == 3 and sum(r["n_rot"] for r in plain[1:]) == L
    # owed bits spread over two tiles (L = 13: bits 10, 11 in the contiguous tile, bit 12 alone): a pass of their own each
    spread = distributed.slice_plan_step(13, 16, _ring_pairs(16), [10, 11, 12], list(range(13)), scatter_g=3)
    assert sum(r["n_pre"] for r in spread) == 0 and sum(r["phase"] for r in spread) == 1 and spread[-1]["scatter"] == 1
    assert sum(r["n_rot"] for r in spread) == 3 + 13


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
    """dq_slice_phase_rx_many (the step's diagonal phase applied to the tile of the first rotation pass) against dq_slice_phase
    followed by dq_slice_rx_many; high rank bits included, and a bit list without low targets (the phase rides on a high-bit tile)."""
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
        assert fused_launches == split_launches - 1          # any 12-bit tile of the plan can carry the phase (TMA tile kernel)


@pytest.mark.gpu
@pytest.mark.parametrize("L,n,high,pre_bits", [(13, 16, 5, [10, 11, 12]), (15, 18, 2, [12, 13, 14]), (16, 16, 0, [15]),
                                               (20, 21, 1, [17, 18, 19]), (21, 21, 0, [3, 7])])
def test_step_with_owed_rotations(L, n, high, pre_bits):
    """dq_slice_step: [rotations owed to the previous step] [phase] [rotations] against the three pieces as separate calls.
    When one tile of the plan holds every owed bit the three ride on ONE pass (launch count checked); (13, ...) has the owed
    bits spread over two tiles and must fall back to a pass of their own."""
    from oracle import restate as R
    ops = distributed.CudaSliceOps(0)
    rng = np.random.RandomState(300 + L)
    edges = R.random_regular_edges(n, seed=L) if n % 2 == 0 else [(i, (i + 1) % n) for i in range(n)] + [(i, (i + 5) % n) for i in range(0, n, 3)]
    pair_bits = np.array([[n - 1 - a, n - 1 - b] for a, b in edges], dtype=np.int32)
    angles = rng.normal(size=1 + len(edges))
    host = rng.normal(size=1 << L) + 1j * rng.normal(size=1 << L)
    host /= np.linalg.norm(host)
    bits = list(range(L))
    for trial in range(2):
        thetas = rng.uniform(-1.2, 1.2, size=L)
        pre_thetas = rng.uniform(-1.2, 1.2, size=len(pre_bits))
        if trial == 1:
            pre_thetas[0] = np.pi / 2               # cos = 0: the unscaled form of the pass
        a, b = ops.alloc(1 << L), ops.alloc(1 << L)
        ops.from_host(a, host)
        ops.from_host(b, host)
        l0 = ops.ctx.launch_count
        ops.step(a, L, high, n, pair_bits, angles, pre_bits, pre_thetas, bits, thetas)
        fused_launches = ops.ctx.launch_count - l0
        l0 = ops.ctx.launch_count
        ops.rx_many(b, L, pre_bits, pre_thetas)
        ops.phase(b, L, high, n, pair_bits, angles)
        ops.rx_many(b, L, bits, thetas)
        split_launches = ops.ctx.launch_count - l0
        ops.ctx.synchronize()
        got, want = ops.to_host(a), ops.to_host(b)
        assert np.abs(got - want).max() < 1e-13
        assert abs(np.linalg.norm(got) - 1.0) < 1e-12
        one_tile = L != 13
        assert fused_launches == split_launches - (2 if one_tile else 1)
    with pytest.raises(ValueError):
        ops.step(a, L, high, n, pair_bits, angles, [L], [0.1], bits, thetas)


@pytest.mark.gpu
@pytest.mark.parametrize("n,steps", [(11, 3), (14, 4), (18, 5), (21, 3), (22, 4)])
def test_chained_steps_equal_step_by_step(n, steps, monkeypatch):
    """dq_slice_evolve_steps (the rotations of one tile set owed across every step boundary: tile sets - 1 passes per step)
    against one dq_slice_phase + dq_slice_rx_many per step, and against the same call with chaining switched off."""
    from oracle import restate as R
    ops = distributed.CudaSliceOps(0)
    rng = np.random.RandomState(400 + n)
    edges = R.random_regular_edges(n, seed=n) if n % 2 == 0 else [(i, (i + 1) % n) for i in range(n)] + [(i, (i + 5) % n) for i in range(0, n, 3)]
    pair_bits = np.array([[n - 1 - a, n - 1 - b] for a, b in edges], dtype=np.int32)
    angle_rows = rng.normal(size=(steps, 1 + len(edges))) * 0.4
    bits = list(rng.permutation(n))
    theta_rows = rng.uniform(-1.0, 1.0, size=(steps, n))
    host = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    host /= np.linalg.norm(host)
    a, b, c = ops.alloc(1 << n), ops.alloc(1 << n), ops.alloc(1 << n)
    for t in (a, b, c):
        ops.from_host(t, host)
    l0 = ops.ctx.launch_count
    ops.evolve_steps(a, n, 0, n, pair_bits, bits, angle_rows, theta_rows)
    chained = ops.ctx.launch_count - l0
    for k in range(steps):
        ops.phase(b, n, 0, n, pair_bits, angle_rows[k])
        ops.rx_many(b, n, bits, theta_rows[k])
    monkeypatch.setenv("DQ_SLICE_NO_CHAIN", "1")
    l0 = ops.ctx.launch_count
    ops.evolve_steps(c, n, 0, n, pair_bits, bits, angle_rows, theta_rows)
    unchained = ops.ctx.launch_count - l0
    monkeypatch.delenv("DQ_SLICE_NO_CHAIN")
    ops.ctx.synchronize()
    got, want, got2 = ops.to_host(a), ops.to_host(b), ops.to_host(c)
    assert np.abs(got - want).max() < 1e-13
    assert np.abs(got2 - want).max() < 1e-13
    assert abs(np.linalg.norm(got) - 1.0) < 1e-12
    if n < 12:                                      # no 12-bit tile: phase pass + rotation pass per step
        assert unchained == chained == 2 * steps
    else:
        sets = 1 if n == 12 else (2 if n <= 21 else 3)
        assert unchained == steps * sets
        assert chained == (steps * (sets - 1) + 1 if sets > 1 else steps)


@pytest.mark.gpu
def test_cp_async_fallback_path_agrees_with_the_tma_path(monkeypatch):
    """DQ_SLICE_NO_TMA=1 routes every pass through k_slice_rx_tile (the kernel that carries exchanges and tile shapes a tensor map
    cannot express): same results as the TMA tile kernel for rotations, a step with owed rotations and a chained sequence."""
    from oracle import restate as R
    n = 18
    ops = distributed.CudaSliceOps(0)
    rng = np.random.RandomState(77)
    edges = R.random_regular_edges(n, seed=n)
    pair_bits = np.array([[n - 1 - a, n - 1 - b] for a, b in edges], dtype=np.int32)
    angle_rows = rng.normal(size=(3, 1 + len(edges))) * 0.4
    theta_rows = rng.uniform(-1.0, 1.0, size=(3, n))
    bits = list(range(n))
    host = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    host /= np.linalg.norm(host)
    results = []
    for no_tma in (False, True):
        if no_tma:
            monkeypatch.setenv("DQ_SLICE_NO_TMA", "1")
        a = ops.alloc(1 << n)
        ops.from_host(a, host)
        l0 = ops.ctx.launch_count
        ops.rx_many(a, n, bits, theta_rows[0])
        ops.step(a, n, 0, n, pair_bits, angle_rows[0], [15, 16, 17], [0.3, -0.2, 0.9], bits, theta_rows[1])
        ops.evolve_steps(a, n, 0, n, pair_bits, bits, angle_rows, theta_rows)
        ops.ctx.synchronize()
        results.append((ops.to_host(a), ops.ctx.launch_count - l0))
    monkeypatch.delenv("DQ_SLICE_NO_TMA")
    (tma, l_tma), (plain, l_plain) = results
    assert np.abs(tma - plain).max() < 1e-13
    assert abs(np.linalg.norm(tma) - 1.0) < 1e-12
    assert l_tma == 2 + 2 + 4 and l_plain == 2 + 3 + 6           # chaining and owed rotations need the TMA tile kernel


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
