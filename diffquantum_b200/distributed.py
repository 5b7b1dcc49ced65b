"""One state vector distributed over the GPUs of a box on its high-order qubits (BASELINE
configs[4]: n = 32 is 64 GiB of complex128 — 8 GiB per B200 at 8 ranks).

Layout.  World size W = 2^g; rank r owns the 2^L amplitudes (L = n - g) whose global basis index has
top g bits equal to r.  In the reference order qubit q sits on index bit n-1-q (demo_maxcut.py:49-57),
so qubits 0..g-1 start out "global".  `pos[q]` tracks the current physical bit of every qubit.

Step (per-term product, diffqc.cc:155-164).  The diagonal phase is local in any layout (the kernel is
told where each ZZ endpoint currently sits).  X rotations commute, so a step rotates every qubit that
is local, then ONE all-to-all swaps index bits [L-g, L) with the rank bits — because those are the top
local bits, every rank's send chunks are contiguous and `all_to_all_single` needs no packing — and the
g qubits that just became local are rotated.  The next step starts from the swapped layout and swaps
back: one exchange of (W-1)/W of the slice per step, no other traffic.  Energies are reduced with one
all-reduce of a scalar.

The device work is the C ABI's dq_slice_* kernels; `ops` exists so that the CPU tests can drive the same
bookkeeping over gloo with a NumPy stand-in for the kernels (tests/test_distributed_state.py).
"""
import ctypes

import numpy as np

from . import _lib
from .sharding import dist_info


def slice_plan(L, n_total, pair_bits, bits, n_steps, assume_tma=True):
    """The launches dq_slice_evolve_steps would make (dq_slice_plan; host only, no device needed): one row per launch with the
    fields step, T, lo, n_pre, phase, n_rot, scatter, mask (the tile's physical bits as a Python int)."""
    pair_bits = np.ascontiguousarray(pair_bits, dtype=np.int32).reshape(-1, 2)
    bits = np.ascontiguousarray(bits, dtype=np.int32)
    lib = _lib.load()
    n_rows = ctypes.c_int64()
    cap = 4 * (n_steps + 1) * 8 + 16
    rows = np.zeros((cap, 9), dtype=np.int32)
    _lib.check(lib.dq_slice_plan(int(L), int(n_total), len(pair_bits), _lib.ptr(pair_bits), len(bits), _lib.ptr(bits), int(n_steps),
                                 1 if assume_tma else 0, _lib.ptr(rows), cap, ctypes.byref(n_rows)))
    assert n_rows.value <= cap
    keys = ("step", "T", "lo", "n_pre", "phase", "n_rot", "scatter")
    out = []
    for r in rows[:n_rows.value]:
        d = {k: int(v) for k, v in zip(keys, r[:7])}
        d["mask"] = (int(r[7]) & 0xffffffff) | ((int(r[8]) & 0xffffffff) << 32)
        out.append(d)
    return out


def slice_plan_step(L, n_total, pair_bits, pre_bits, bits, scatter_g=0, assume_tma=True):
    """The launches of ONE step of a distributed slice (dq_slice_plan_step; host only): rows as in slice_plan."""
    pair_bits = np.ascontiguousarray(pair_bits, dtype=np.int32).reshape(-1, 2)
    pre_bits = np.ascontiguousarray(pre_bits, dtype=np.int32)
    bits = np.ascontiguousarray(bits, dtype=np.int32)
    n_rows = ctypes.c_int64()
    rows = np.zeros((64, 9), dtype=np.int32)
    _lib.check(_lib.load().dq_slice_plan_step(int(L), int(n_total), len(pair_bits), _lib.ptr(pair_bits), len(pre_bits),
                                              _lib.ptr(pre_bits), len(bits), _lib.ptr(bits), int(scatter_g), 1 if assume_tma else 0,
                                              _lib.ptr(rows), len(rows), ctypes.byref(n_rows)))
    keys = ("step", "T", "lo", "n_pre", "phase", "n_rot", "scatter")
    out = []
    for r in rows[:n_rows.value]:
        d = {k: int(v) for k, v in zip(keys, r[:7])}
        d["mask"] = (int(r[7]) & 0xffffffff) | ((int(r[8]) & 0xffffffff) << 32)
        out.append(d)
    return out


class CudaSliceOps(object):
    """Slice kernels through the C ABI; buffers are torch CUDA tensors (device memory + NCCL only)."""

    def __init__(self, device=0):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.ctx = _lib.Context.get(device)
        self.lib = _lib.load()

    def alloc(self, n_amps):
        return self.torch.empty(n_amps, dtype=self.torch.complex128, device=self.device)

    def _p(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def fill_uniform(self, psi, L, n):
        _lib.check(self.lib.dq_slice_fill_uniform(self.ctx.handle, self._p(psi), L, n))

    def phase(self, psi, L, high, n, pair_bits, angles):
        pair_bits = np.ascontiguousarray(pair_bits, dtype=np.int32)
        angles = np.ascontiguousarray(angles, dtype=np.float64)
        _lib.check(self.lib.dq_slice_phase(self.ctx.handle, self._p(psi), L, high, n, len(pair_bits), _lib.ptr(pair_bits),
                                           _lib.ptr(angles)))

    def rx(self, psi, L, bit, theta):
        _lib.check(self.lib.dq_slice_rx(self.ctx.handle, self._p(psi), L, int(bit), float(theta)))

    def rx_many(self, psi, L, bits, thetas):
        bits = np.ascontiguousarray(bits, dtype=np.int32)
        thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        _lib.check(self.lib.dq_slice_rx_many(self.ctx.handle, self._p(psi), L, len(bits), _lib.ptr(bits), _lib.ptr(thetas)))

    def phase_rx_many(self, psi, L, high, n, pair_bits, angles, bits, thetas):
        """Phase + the step's local rotations in one call: the phase rides on the first rotation pass (dq_slice_phase_rx_many)."""
        pair_bits = np.ascontiguousarray(pair_bits, dtype=np.int32)
        angles = np.ascontiguousarray(angles, dtype=np.float64)
        bits = np.ascontiguousarray(bits, dtype=np.int32)
        thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        _lib.check(self.lib.dq_slice_phase_rx_many(self.ctx.handle, self._p(psi), L, high, n, len(pair_bits), _lib.ptr(pair_bits),
                                                   _lib.ptr(angles), len(bits), _lib.ptr(bits), _lib.ptr(thetas)))

    def step(self, psi, L, high, n, pair_bits, angles, pre_bits, pre_thetas, bits, thetas):
        """[rotations still owed to the previous step] [phase] [rotations]: when one tile of the pass plan holds the owed bits,
        all three ride on one pass over the slice (dq_slice_step)."""
        pair_bits = np.ascontiguousarray(pair_bits, dtype=np.int32)
        angles = np.ascontiguousarray(angles, dtype=np.float64)
        pre_bits = np.ascontiguousarray(pre_bits, dtype=np.int32)
        pre_thetas = np.ascontiguousarray(pre_thetas, dtype=np.float64)
        bits = np.ascontiguousarray(bits, dtype=np.int32)
        thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        _lib.check(self.lib.dq_slice_step(self.ctx.handle, self._p(psi), L, high, n, len(pair_bits), _lib.ptr(pair_bits),
                                          _lib.ptr(angles), len(pre_bits), _lib.ptr(pre_bits), _lib.ptr(pre_thetas), len(bits),
                                          _lib.ptr(bits), _lib.ptr(thetas)))

    def evolve_steps(self, psi, L, high, n, pair_bits, bits, angle_rows, theta_rows):
        """Steps of a slice that needs no exchange, chained over the step boundaries (dq_slice_evolve_steps)."""
        pair_bits = np.ascontiguousarray(pair_bits, dtype=np.int32)
        bits = np.ascontiguousarray(bits, dtype=np.int32)
        angle_rows = np.ascontiguousarray(angle_rows, dtype=np.float64)
        theta_rows = np.ascontiguousarray(theta_rows, dtype=np.float64)
        assert angle_rows.shape == (len(theta_rows), 1 + len(pair_bits)) and theta_rows.shape[1] == len(bits)
        _lib.check(self.lib.dq_slice_evolve_steps(self.ctx.handle, self._p(psi), L, high, n, len(pair_bits), _lib.ptr(pair_bits),
                                                  len(bits), _lib.ptr(bits), len(theta_rows), _lib.ptr(angle_rows),
                                                  angle_rows.shape[1], _lib.ptr(theta_rows), theta_rows.shape[1]))

    # -- exchange fused into the last local pass (peer memory) ------------------------------------------------
    def enable_peer_exchange(self, buffers, rank, world):
        """Map every rank's two slice buffers into this process (CUDA IPC; peer access over NVLink when the ranks own
        different GPUs) so that the last local rotation pass of a step can store straight into the peers' receive buffers.
        buffers: this rank's two tensors (they swap roles every step).  Returns False when the mapping is not possible
        (the all-to-all path stays in use)."""
        import torch.distributed as dist
        mine = []
        try:
            for t in buffers:
                handle = (ctypes.c_ubyte * 64)()
                off = ctypes.c_uint64()
                _lib.check(self.lib.dq_ipc_export(self.ctx.handle, self._p(t), handle, ctypes.byref(off)))
                mine.append((bytes(handle), int(off.value)))
        except Exception:
            mine = None
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        if any(e is None for e in everyone):
            return False
        self._peer_tables, self._peer_opened = [], []
        try:
            for k in range(len(buffers)):
                table = (ctypes.c_void_p * world)()
                for j in range(world):
                    if j == rank:
                        table[j] = buffers[k].data_ptr()
                    else:
                        ptr = ctypes.c_void_p()
                        h, off = everyone[j][k]
                        _lib.check(self.lib.dq_ipc_open(self.ctx.handle, ctypes.c_char_p(h), off, ctypes.byref(ptr)))
                        self._peer_opened.append((ptr.value, off))
                        table[j] = ptr.value
                self._peer_tables.append(table)
            ok = True
        except Exception:
            ok = False
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        self._peer_ptrs = {buffers[k].data_ptr(): self._peer_tables[k] for k in range(len(buffers))} if all(flags) else None
        self._rank, self._g = rank, world.bit_length() - 1
        return self._peer_ptrs is not None

    def phase_rx_many_scatter(self, psi, recv, L, high, n, pair_bits, angles, bits, thetas):
        """Phase, the local rotations and the exchange of a step: the last pass writes into every rank's `recv` buffer."""
        pair_bits = np.ascontiguousarray(pair_bits, dtype=np.int32)
        angles = np.ascontiguousarray(angles, dtype=np.float64)
        bits = np.ascontiguousarray(bits, dtype=np.int32)
        thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        table = self._peer_ptrs[recv.data_ptr()]
        _lib.check(self.lib.dq_slice_phase_rx_many_scatter(
            self.ctx.handle, self._p(psi), L, high, n, len(pair_bits), _lib.ptr(pair_bits), _lib.ptr(angles), len(bits),
            _lib.ptr(bits), _lib.ptr(thetas), self._g, self._rank, table))

    def step_scatter(self, psi, recv, L, high, n, pair_bits, angles, pre_bits, pre_thetas, bits, thetas):
        """dq_slice_step whose last pass writes into every rank's `recv` buffer."""
        pair_bits = np.ascontiguousarray(pair_bits, dtype=np.int32)
        angles = np.ascontiguousarray(angles, dtype=np.float64)
        pre_bits = np.ascontiguousarray(pre_bits, dtype=np.int32)
        pre_thetas = np.ascontiguousarray(pre_thetas, dtype=np.float64)
        bits = np.ascontiguousarray(bits, dtype=np.int32)
        thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        table = self._peer_ptrs[recv.data_ptr()]
        _lib.check(self.lib.dq_slice_step_scatter(
            self.ctx.handle, self._p(psi), L, high, n, len(pair_bits), _lib.ptr(pair_bits), _lib.ptr(angles), len(pre_bits),
            _lib.ptr(pre_bits), _lib.ptr(pre_thetas), len(bits), _lib.ptr(bits), _lib.ptr(thetas), self._g, self._rank, table))

    def barrier(self):
        """Every rank's stores into every receive buffer are complete (the pass has finished on every device)."""
        import torch.distributed as dist
        self.ctx.synchronize()
        if dist.get_backend() == "nccl":
            dist.barrier(device_ids=[self.device.index])
        else:
            dist.barrier()

    def energy(self, psi, L, high, n, pair_bits, m_zz, m_const):
        pair_bits = np.ascontiguousarray(pair_bits, dtype=np.int32)
        m_zz = np.ascontiguousarray(m_zz, dtype=np.float64)
        v = ctypes.c_double()
        _lib.check(self.lib.dq_slice_energy(self.ctx.handle, self._p(psi), L, high, n, len(pair_bits), _lib.ptr(pair_bits),
                                            _lib.ptr(m_zz), float(m_const), ctypes.byref(v)))
        return v.value

    def all_to_all(self, recv, send):
        import torch.distributed as dist
        self.ctx.synchronize()                      # kernels run on the library's stream, NCCL on torch's
        dist.all_to_all_single(self.torch.view_as_real(recv), self.torch.view_as_real(send))
        self.torch.cuda.current_stream(self.device).synchronize()

    def all_reduce_scalar(self, x):
        import torch.distributed as dist
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        dist.all_reduce(t)
        return float(t.item())

    def to_host(self, psi):
        self.ctx.synchronize()
        return psi.cpu().numpy()

    def from_host(self, psi, array):
        self.ctx.synchronize()
        psi.copy_(self.torch.from_numpy(np.ascontiguousarray(array, dtype=np.complex128)))
        self.torch.cuda.current_stream(self.device).synchronize()


class DistributedState(object):
    """The slice of one n-qubit state owned by this rank, plus the qubit -> physical-bit map."""

    def __init__(self, problem, device=0, per_step=10, basis="BSpline", ops=None, peer_exchange=None):
        self.problem = problem
        self.per_step = per_step
        self.basis = basis
        self.rank, self.world, _ = dist_info()
        g = self.world.bit_length() - 1
        if (1 << g) != self.world:
            raise ValueError("world size %d is not a power of two" % self.world)
        self.g = g
        self.n = problem.n
        self.L = self.n - g
        if self.L < max(1, g):
            raise ValueError("%d qubits cannot be split over %d ranks" % (self.n, self.world))
        self.ops = ops if ops is not None else CudaSliceOps(device)
        self.psi = self.ops.alloc(1 << self.L)
        self.recv = self.ops.alloc(1 << self.L) if g else None
        self.pos = [self.n - 1 - q for q in range(self.n)]
        self.exchanges = 0
        self.exchanged_bytes = 0
        self.fused_rx = True               # False: one kernel per rotation (dq_slice_rx), kept for cross-checks
        self.fused_phase = True            # False: the diagonal phase as its own pass (dq_slice_phase)
        # The rotations of the qubits that became local in a step's exchange are not run as a pass of their own: they stay OWED
        # and ride, with the next step's phase, on the first pass of the next step (dq_slice_step) -- one pass over the slice
        # less per step.  flush() runs what is owed (before anything reads the state).
        self.defer_post = hasattr(self.ops, "step")
        self.owed = None                   # (qubits, angles)
        # exchange fused into the stores of the last local pass (peer memory over NVLink) instead of an NCCL all-to-all
        self.peer_exchange = bool(g) and hasattr(self.ops, "enable_peer_exchange") and peer_exchange is not False and \
            self.ops.enable_peer_exchange([self.psi, self.recv], self.rank, self.world)

    # -- layout ------------------------------------------------------------------------------------
    def pair_bits(self):
        return np.array([[self.pos[a], self.pos[b]] for a, b in self.problem.zz_pairs], dtype=np.int32).reshape(-1, 2)

    def global_qubits(self):
        return [q for q in range(self.n) if self.pos[q] >= self.L]

    def flush(self):
        """Run the rotations still owed to the last step (see defer_post)."""
        if self.owed is not None:
            qubits, angles = self.owed
            self.owed = None
            self._rotate_angles(qubits, angles)

    def swap_global_local(self):
        """All-to-all that exchanges index bits [L-g, L) with the rank bits [L, n)."""
        if self.g == 0:
            return
        self.flush()
        self.ops.all_to_all(self.recv, self.psi)
        self._swapped()

    def _swapped(self):
        """Bookkeeping after the amplitudes have moved: buffers change roles, index bits [L-g, L) <-> rank bits."""
        self.psi, self.recv = self.recv, self.psi
        L, g = self.L, self.g
        for q in range(self.n):
            if L - g <= self.pos[q] < L:
                self.pos[q] += g
            elif self.pos[q] >= L:
                self.pos[q] -= g
        self.exchanges += 1
        self.exchanged_bytes += 16 * (1 << L) * (self.world - 1) // self.world

    def restore_layout(self):
        self.flush()
        if self.pos != [self.n - 1 - q for q in range(self.n)]:
            self.swap_global_local()

    # -- evolution -----------------------------------------------------------------------------------
    def fill_uniform(self):
        """The demo's start state (demo_maxcut.py:12-17)."""
        self.restore_layout()
        self.ops.fill_uniform(self.psi, self.L, self.n)

    def set_state(self, full_state):
        """Tests: load this rank's slice of a host state given in the reference order."""
        self.restore_layout()
        N = 1 << self.L
        self.ops.from_host(self.psi, np.asarray(full_state).reshape(-1)[self.rank * N:(self.rank + 1) * N])

    def step(self, row):
        """One product-formula step from an angle row [c | zz angles | x angle per qubit] (already x dt)."""
        p = self.problem
        row = np.asarray(row, dtype=np.float64)
        x = row[1 + p.n_zz:]
        was_global = self.global_qubits()
        local = [q for q in range(self.n) if self.pos[q] < self.L]
        fused = self.fused_rx and self.fused_phase and bool(local)
        if self.owed is not None and not (fused and self.defer_post):
            self.flush()
        pre_bits, pre_thetas = [], []
        if self.owed is not None:
            pre_bits, pre_thetas = [self.pos[q] for q in self.owed[0]], list(self.owed[1])
            self.owed = None
        bits, thetas = [self.pos[q] for q in local], [x[q] for q in local]
        if self.peer_exchange and was_global and local:
            # one call: (owed rotations +) phase + every local rotation, the last pass storing into the peers' receive buffers;
            # then a barrier
            if pre_bits or (self.defer_post and hasattr(self.ops, "step_scatter")):
                self.ops.step_scatter(self.psi, self.recv, self.L, self.rank, self.n, self.pair_bits(), row[:1 + p.n_zz],
                                      pre_bits, pre_thetas, bits, thetas)
            else:
                self.ops.phase_rx_many_scatter(self.psi, self.recv, self.L, self.rank, self.n, self.pair_bits(), row[:1 + p.n_zz],
                                               bits, thetas)
            self.ops.barrier()
            self._swapped()
            self._post(was_global, x)
            return
        if fused and self.defer_post:
            self.ops.step(self.psi, self.L, self.rank, self.n, self.pair_bits(), row[:1 + p.n_zz], pre_bits, pre_thetas, bits, thetas)
        elif fused and hasattr(self.ops, "phase_rx_many"):
            self.ops.phase_rx_many(self.psi, self.L, self.rank, self.n, self.pair_bits(), row[:1 + p.n_zz], bits, thetas)
        else:
            self.ops.phase(self.psi, self.L, self.rank, self.n, self.pair_bits(), row[:1 + p.n_zz])
            self._rotate(local, x)
        if was_global:
            self.swap_global_local()
            self._post(was_global, x)

    def _post(self, qubits, x):
        """The step's rotations on the qubits that have just become local: owed to the next step's first pass, or run now."""
        if self.defer_post and self.fused_rx and self.fused_phase:
            self.owed = (list(qubits), [x[q] for q in qubits])
        else:
            self._rotate(qubits, x)

    def _rotate_angles(self, qubits, angles):
        x = {q: a for q, a in zip(qubits, angles)}
        self._rotate(qubits, x)

    def _rotate(self, qubits, x):
        """X rotations of one step on local qubits: they commute, so the fused pass kernel takes them all at once."""
        if not qubits:
            return
        if self.fused_rx and hasattr(self.ops, "rx_many"):
            self.ops.rx_many(self.psi, self.L, [self.pos[q] for q in qubits], [x[q] for q in qubits])
        else:
            for q in qubits:
                self.ops.rx(self.psi, self.L, self.pos[q], x[q])

    def evolve_rows(self, rows):
        rows = np.asarray(rows, dtype=np.float64).reshape(-1, self.problem.row_len)
        if self.g == 0 and self.fused_rx and self.fused_phase and hasattr(self.ops, "evolve_steps") and len(rows):
            # nothing to exchange: the whole sequence in one call, chained over the step boundaries
            p = self.problem
            self.ops.evolve_steps(self.psi, self.L, self.rank, self.n, self.pair_bits(), [self.pos[q] for q in range(self.n)],
                                  rows[:, :1 + p.n_zz], rows[:, 1 + p.n_zz:])
            return
        for row in rows:
            self.step(row)
        self.flush()

    def evolve(self, coeff, T0, T1):
        """SimulatorPlain.trotter's step grid (sim_plain.py:123-150) with the product-formula step."""
        self.evolve_rows(self.problem.trajectory_rows(coeff, T0, T1, self.per_step, self.basis))

    def energy(self):
        """<psi|M|psi> for the problem's diagonal observable, summed over ranks."""
        p = self.problem
        self.flush()
        if p.m_diag is not None:
            raise NotImplementedError("distributed energies need the observable in ZZ form (m_zz, m_const)")
        part = self.ops.energy(self.psi, self.L, self.rank, self.n, self.pair_bits(), p.m_zz, p.m_const)
        return self.ops.all_reduce_scalar(part) if self.world > 1 else part

    def norm2(self):
        self.flush()
        zero = np.zeros(self.problem.n_zz)
        part = self.ops.energy(self.psi, self.L, self.rank, self.n, self.pair_bits(), zero, 1.0)
        return self.ops.all_reduce_scalar(part) if self.world > 1 else part

    def local_slice(self):
        """This rank's amplitudes in the reference order (restores the layout first)."""
        self.restore_layout()
        return self.ops.to_host(self.psi)
