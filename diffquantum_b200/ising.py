"""Structured (Pauli-term) problems on the B200: H(t) = H0 + sum_i u_i(t) H_i with
H0 = c0 + sum_e w_e Z_a Z_b (diagonal) and every control H_i either Z_a Z_b or X_q.

This is the large-n form of what the reference builds densely with np.kron
(demo_maxcut.py:19-85, sim_plain.py:477-482).  The host evaluates pulses (pulses.py) and hands
per-step ANGLES to the C ABI (include/diffqc_b200.h, dq_ising_*); all amplitudes stay on the device.

Step semantics are the product formula the reference keeps in its disabled variant
(diffqc.cc:155-164): exp(-i dt H0), then exp(-i dt u_h H_h) for h in list order.
"""
import ctypes

import numpy as np

from . import _lib
from . import pulses


class IsingProblem(object):
    """n qubits; `zz_pairs` lists the distinct (a, b) pairs that carry a ZZ operator anywhere
    (H0, a control, or the observable).  `terms` is the control list in the reference's order:
    ('zz', a, b) or ('x', q), one pulse u_i(t) each (demo_maxcut.py:68-79)."""

    def __init__(self, n, terms, omegas, T, h0_zz=None, h0_const=0.0, m_zz=None, m_const=0.0,
                 m_diag=None, psi0=None):
        self.n = int(n)
        self.terms = [tuple(t) for t in terms]
        self.omegas = np.asarray(omegas, dtype=np.float64)
        if len(self.omegas) != len(self.terms):
            raise ValueError("one omega per control term is required")
        self.T = float(T)
        pairs = {}

        def pair_id(a, b):
            a, b = int(a), int(b)
            if not (0 <= a < self.n and 0 <= b < self.n) or a == b:
                raise ValueError("bad ZZ pair (%d, %d) for %d qubits" % (a, b, self.n))
            key = (min(a, b), max(a, b))
            if key not in pairs:
                pairs[key] = len(pairs)
            return pairs[key]

        h0_zz = dict(h0_zz or {})
        m_zz = dict(m_zz or {})
        self.term_kind = np.zeros(len(self.terms), dtype=np.int32)     # 0 zz, 1 x
        self.term_index = np.zeros(len(self.terms), dtype=np.int32)    # pair id or qubit
        for i, t in enumerate(self.terms):
            if t[0] == 'zz':
                self.term_kind[i] = 0
                self.term_index[i] = pair_id(t[1], t[2])
            elif t[0] == 'x':
                if not 0 <= int(t[1]) < self.n:
                    raise ValueError("bad X qubit %r" % (t[1],))
                self.term_kind[i] = 1
                self.term_index[i] = int(t[1])
            else:
                raise ValueError("unsupported control %r (ZZ and X Pauli terms only)" % (t,))
        # One step is exp(-i dt H0), then exp(-i dt u_h H_h) in LIST order (diffqc.cc:155-164).  The device applies the
        # whole diagonal factor first and every X rotation after it, which is that product exactly when no X control
        # precedes a ZZ control in the list (the demo's order, demo_maxcut.py:68-79: edges, then qubits) -- diagonal
        # factors commute with each other, X rotations on different qubits commute with each other, nothing else does.
        kinds = self.term_kind.tolist()
        if 1 in kinds and 0 in kinds[kinds.index(1):]:
            raise ValueError("unsupported control order: an X control precedes a ZZ control; the per-term product "
                             "exp(-i dt u_h H_h) in list order (diffqc.cc:155-164) is only implemented for ZZ controls "
                             "first, X controls after (demo_maxcut.py:68-79)")
        for (a, b) in list(h0_zz) + list(m_zz):
            pair_id(a, b)
        self.zz_pairs = np.array(sorted(pairs, key=pairs.get), dtype=np.int32).reshape(-1, 2)
        self.n_zz = len(self.zz_pairs)
        self.h0_zz = np.zeros(self.n_zz)
        self.m_zz = np.zeros(self.n_zz)
        for (a, b), w in h0_zz.items():
            self.h0_zz[pair_id(a, b)] += w
        for (a, b), w in m_zz.items():
            self.m_zz[pair_id(a, b)] += w
        self.h0_const = float(h0_const)
        self.m_const = float(m_const)
        self.m_diag = None if m_diag is None else np.ascontiguousarray(m_diag, dtype=np.float64)
        self.psi0 = None if psi0 is None else np.ascontiguousarray(psi0, dtype=np.complex128)
        self.row_len = 1 + self.n_zz + self.n
        # shot sampling (sim.Pauli_M of the reference, demo_maxcut.py:47-65): Z-string terms [(pair id or -1 = identity, weight)]
        self.measure_terms = None

    def set_measurement(self, terms):
        """Pauli_M in structured form: [((a, b), weight), ..., (None, weight)] -- Z_a Z_b strings and the identity, in the
        reference's order (the draw order of stochastic_measure, sim_plain.py:104).  Every pair must be one of the problem's."""
        ids = {tuple(map(int, pr)): i for i, pr in enumerate(self.zz_pairs)}
        out = []
        for pair, w in terms:
            if pair is None:
                out.append((-1, float(w)))
            else:
                key = (min(pair), max(pair))
                if key not in ids:
                    raise ValueError("measurement pair %r is not a ZZ pair of this problem" % (pair,))
                out.append((ids[key], float(w)))
        self.measure_terms = out

    # -- the reference's MaxCut construction ---------------------------------------------------
    @classmethod
    def maxcut(cls, n, edges, omega0=np.pi, omega1=np.pi, n_layers=1):
        """demo_maxcut.py:19-85 in structured form: controls = ZZ per edge then X per qubit;
        H0 = sum_e (I - Z_a Z_b) (the aliasing of `H0`/`H_cost`, demo_maxcut.py:34-38,60-61);
        M = -1/2 sum_e (I - Z_a Z_b); T = pi (1/omega0 + 1/omega1) n_layers; |+...+> start."""
        edges = [(int(a), int(b)) for a, b in edges]
        terms = [('zz', a, b) for a, b in edges] + [('x', q) for q in range(n)]
        omegas = [omega0] * len(edges) + [omega1] * n
        T = np.pi * (1. / omega0 + 1. / omega1) * n_layers
        h0 = {}
        mz = {}
        for a, b in edges:
            h0[(a, b)] = h0.get((a, b), 0.0) - 1.0
            mz[(a, b)] = mz.get((a, b), 0.0) + 0.5
        prob = cls(n, terms, omegas, T, h0_zz=h0, h0_const=float(len(edges)), m_zz=mz,
                   m_const=-0.5 * len(edges))
        prob.set_measurement([((a, b), 0.5) for a, b in edges] + [(None, -0.5 * len(edges))])   # demo_maxcut.py:47-62
        return prob

    # -- host-side angle tables ------------------------------------------------------------------
    def angle_rows(self, u, dt):
        """rows[k] = [dt*c0 | dt*(w_e + sum of ZZ pulses on e) | dt*(sum of X pulses on q)]."""
        u = np.asarray(u, dtype=np.float64)
        K = u.shape[0]
        rows = np.zeros((K, self.row_len))
        rows[:, 0] = self.h0_const
        rows[:, 1:1 + self.n_zz] = self.h0_zz[None, :]
        cols = np.where(self.term_kind == 0, 1 + self.term_index, 1 + self.n_zz + self.term_index)
        if len(np.unique(cols)) == len(cols):
            rows[:, cols] += u                     # every column takes one pulse: one indexed add, the same additions
        else:
            for i in range(len(self.terms)):       # several controls on one pair / qubit: summed in list order
                rows[:, cols[i]] += u[:, i]
        return rows * dt

    def trajectory_rows(self, coeff, T0, T1, per_step, basis='BSpline'):
        n_steps, dt, ts = pulses.step_grid(T0, T1, per_step)
        if n_steps == 0:
            return np.zeros((0, self.row_len))
        return self.angle_rows(pulses.u_table(coeff, self.omegas, self.T, ts, basis), dt)


class IsingSimulator(object):
    """Device-side evolution and batched stochastic parameter-shift samples for an IsingProblem."""

    def __init__(self, problem, device=0, per_step=10, basis='BSpline', engine=None, ket_group=None, step='split'):
        self.problem = problem
        self.per_step = per_step
        self.basis = basis
        self.ctx = _lib.Context.get(device)
        lib = _lib.load()
        h = ctypes.c_void_p()
        p = problem
        pairs = np.ascontiguousarray(p.zz_pairs, dtype=np.int32)
        _lib.check(lib.dq_ising_create(self.ctx.handle, p.n, p.n_zz, _lib.ptr(pairs), _lib.ptr(p.m_zz),
                                       p.m_const, _lib.ptr(p.m_diag), ctypes.byref(h)))
        self.handle = h
        if engine is not None:
            self.set_option("engine", engine)
        if ket_group is not None:
            self.set_option("ket_group", ket_group)
        if step not in ('split', 'exact'):
            raise ValueError("step must be 'split' (per-term product, diffqc.cc:155-164) or 'exact' "
                             "(live semantics, sim_plain.py:135-150)")
        self.step = step
        if step == 'exact':
            self.set_option("step", 1)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().dq_ising_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_option(self, name, value):
        _lib.check(_lib.load().dq_ising_set_option(self.handle, name.encode(), int(value)))

    def info(self, name):
        v = ctypes.c_int64()
        _lib.check(_lib.load().dq_ising_get_info(self.handle, name.encode(), ctypes.byref(v)))
        return v.value

    def stat(self, name):
        v = ctypes.c_double()
        _lib.check(_lib.load().dq_ising_last_stat(self.handle, name.encode(), ctypes.byref(v)))
        return v.value

    # -- single / batched evolution ---------------------------------------------------------------
    def evolve_rows(self, rows, psi0=None, batch=1, want_state=True, want_energy=True):
        """Evolve through explicit angle rows.  psi0: None (uniform superposition), a host
        complex128 array [batch, 2^n], or a CUDA torch tensor of that shape (used in place as the
        input, a new tensor is returned)."""
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, self.problem.row_len)
        N = 1 << self.problem.n
        is_dev = 0
        out = None
        pin = pout = None
        keep = []
        if psi0 is not None and hasattr(psi0, "data_ptr"):      # torch tensor handoff
            import torch
            if not psi0.is_cuda or psi0.dtype != torch.complex128:
                raise TypeError("device states must be CUDA complex128 tensors")
            psi0 = psi0.contiguous()
            batch = psi0.numel() // N
            is_dev = 1
            pin = ctypes.c_void_p(psi0.data_ptr())
            if want_state:
                out = torch.empty_like(psi0)
                pout = ctypes.c_void_p(out.data_ptr())
            torch.cuda.current_stream(psi0.device).synchronize()
            keep.append(psi0)
        else:
            if psi0 is not None:
                psi0 = np.ascontiguousarray(psi0, dtype=np.complex128).reshape(-1, N)
                batch = psi0.shape[0]
                pin = _lib.ptr(psi0)
            if want_state:
                out = np.empty((batch, N), dtype=np.complex128)
                pout = _lib.ptr(out)
        en = np.empty(batch) if want_energy else None
        _lib.check(_lib.load().dq_ising_evolve(self.handle, batch, rows.shape[0], _lib.ptr(rows), pin, pout,
                                               is_dev, _lib.ptr(en)))
        return out, en

    def evolve(self, coeff, T0, T1, psi0=None, **kw):
        """SimulatorPlain.trotter semantics on the step grid of sim_plain.py:123-150."""
        rows = self.problem.trajectory_rows(coeff, T0, T1, self.per_step, self.basis)
        return self.evolve_rows(rows, psi0, **kw)

    def trotter_cc(self, channels, duration, func_type, vv, T0, T, psi0=None, **kw):
        """diffqc.trotter (diffqc.cc:173-205) for a Pauli-term Hamiltonian: the IQ-channel pulse model f_u of the native twin
        (diffqc.cc:95-135: per control a sum over channels [_, omega, w, idx] of omega (2 expit(N) - 1) / N (cos(w t) A +
        sin(w t) B), A/B = basis expansions of vv[0][idx] / vv[1][idx]) evaluated by the library's host routine on the
        native step grid (n_steps from |T - T0|, diffqc.cc:182-184), one channel list per control term of the problem."""
        p = self.problem
        if len(channels) != len(p.terms):
            raise ValueError("one channel list per control term is required (%d terms, %d lists)" % (len(p.terms), len(channels)))
        n_steps, dt, ts = pulses.step_grid(T0, T, self.per_step, use_abs=True)
        if n_steps == 0:
            raise ValueError("per_step=%r gives no steps (the reference divides by zero here)" % (self.per_step,))
        u = pulses.f_u_table_lib(channels, duration, func_type, vv, ts)
        return self.evolve_rows(p.angle_rows(u, dt), psi0, **kw)

    # -- batched estimator --------------------------------------------------------------------------
    def sample_tables(self, coeff, s_list):
        """Host tables for a batch of sampled times: step counts and packed angle rows."""
        p = self.problem
        pre_steps, suf_steps, pre_rows, suf_rows = [], [], [], []
        for s in s_list:
            a = p.trajectory_rows(coeff, 0, s, self.per_step, self.basis)
            b = p.trajectory_rows(coeff, s, p.T, self.per_step, self.basis)
            pre_steps.append(a.shape[0])
            suf_steps.append(b.shape[0])
            pre_rows.append(a)
            suf_rows.append(b)
        cat = lambda xs: np.ascontiguousarray(np.concatenate(xs, axis=0)) if xs else np.zeros((0, p.row_len))
        return (np.array(pre_steps, dtype=np.int32), cat(pre_rows),
                np.array(suf_steps, dtype=np.int32), cat(suf_rows))

    def _grad_args(self, tables, r):
        p = self.problem
        pre_steps, pre_rows, suf_steps, suf_rows = tables
        return (self.handle, len(pre_steps), _lib.ptr(pre_steps), _lib.ptr(pre_rows), _lib.ptr(suf_steps),
                _lib.ptr(suf_rows), len(p.terms), _lib.ptr(p.term_kind), _lib.ptr(p.term_index), float(r),
                _lib.ptr(p.psi0))

    def shifted_energies(self, coeff, s_list, r=0.5):
        """energies[b, i, 0|1] = <M> after the (+, -) shifted trajectories of sample b, term i
        (ps_p, ps_m of sim_plain.py:205,215)."""
        tables = self.sample_tables(coeff, s_list)
        out = np.empty((len(s_list), len(self.problem.terms), 2))
        args = self._grad_args(tables, r) + (_lib.ptr(out),)
        _lib.check(_lib.load().dq_ising_grad(*args))
        return out

    def assemble_gradients(self, coeff, s_list, energies, r=0.5, coeff_sign=1.0):
        """grad[b, i, j] = coeff_sign (1+r^2)/(2r) (ps_m - ps_p) dDdv[i, j]   (sim_plain.py:220-227)."""
        p = self.problem
        grads = np.empty((len(s_list),) + np.asarray(coeff).shape)
        for b, s in enumerate(s_list):
            ps = coeff_sign * ((1 + r ** 2) / 2 / r * (energies[b, :, 1] - energies[b, :, 0]))
            grads[b] = ps[:, None] * pulses.dudc_table(coeff, p.omegas, p.T, s, self.basis)
        return grads

    def grad_samples(self, coeff, s_list, r=0.5, coeff_sign=1.0, return_energies=False, is_noisy=False,
                     sampling_measure=False, per_Pauli=100):
        """Per-sample gradients of compute_energy_grad_MC (sim_plain.py:156-231) for explicit
        sampled times s_list (the reference draws s = np.random.uniform() * T at :167).  is_noisy: the reference's
        measurement noise on every shifted energy (sim_plain.py:207-208,217-218; pulses.add_measurement_noise).
        sampling_measure: shifted energies by shot sampling (stochastic_measure, :202-203,212-213)."""
        s_list = np.asarray(s_list, dtype=np.float64).reshape(-1)
        if sampling_measure:
            en = self.sampled_shifted_energies(coeff, s_list, r, per_Pauli, is_noisy)
        else:
            en = self.shifted_energies(coeff, s_list, r)
            if is_noisy:
                pulses.add_measurement_noise(en)
        g = self.assemble_gradients(coeff, s_list, en, r, coeff_sign)
        return (g, en) if return_energies else g

    # -- shot sampling (stochastic_measure, sim_plain.py:101-117) --------------------------------------
    def pair_expectations(self, psi):
        """<Z_a Z_b> of every ZZ pair for host states [batch, 2^n]: [batch, n_zz] (dq_ising_pair_expect)."""
        N = 1 << self.problem.n
        psi = np.ascontiguousarray(psi, dtype=np.complex128).reshape(-1, N)
        out = np.empty((psi.shape[0], self.problem.n_zz))
        _lib.check(_lib.load().dq_ising_pair_expect(self.handle, psi.shape[0], _lib.ptr(psi), 0, _lib.ptr(out)))
        return out

    def _measure_tables(self):
        mt = self.problem.measure_terms
        if mt is None:
            raise ValueError("problem.set_measurement(...) first (the Z-string terms of sim.Pauli_M)")
        return np.array([t[0] for t in mt]), [t[1] for t in mt]

    def stochastic_measure(self, psi, per_Pauli=100):
        pair, w = self._measure_tables()
        return pulses.stochastic_measure_zstrings(self.pair_expectations(psi)[0], pair, w, per_Pauli)

    def shifted_pair_expectations(self, coeff, s_list, r=0.5):
        """<Z_a Z_b> of every pair in every shifted ket: [B, n_terms, 2, n_zz] (dq_ising_grad_pairs)."""
        s_list = np.asarray(s_list, dtype=np.float64).reshape(-1)
        tables = self.sample_tables(coeff, s_list)
        out = np.empty((len(s_list), len(self.problem.terms), 2, self.problem.n_zz))
        _lib.check(_lib.load().dq_ising_grad_pairs(*(self._grad_args(tables, r) + (_lib.ptr(out),))))
        return out

    def sampled_shifted_energies(self, coeff, s_list, r=0.5, per_Pauli=100, is_noisy=False):
        """ps_p / ps_m by shot sampling in the reference's draw order (sim_plain.py:196-218)."""
        pair, w = self._measure_tables()
        zz = self.shifted_pair_expectations(coeff, s_list, r)
        en = np.empty(zz.shape[:3])
        for b in range(zz.shape[0]):
            for i in range(zz.shape[1]):
                for k in range(2):
                    v = pulses.stochastic_measure_zstrings(zz[b, i, k], pair, w, per_Pauli)
                    if is_noisy:
                        v += np.random.normal(scale=np.abs(v) / 5)
                    en[b, i, k] = v
        return en

    # -- device-resident training loop ---------------------------------------------------------------
    def train_energy_device(self, coeff0, s_all, lr=2e-2, betas=(0.9, 0.999), eps=1e-8, r=0.5, e0=None, want_state=True):
        """The whole train_energy loop (sim_plain.py:245-305) on the device: dq_ising_train.  s_all [n_epoch, K] are the
        sample times of every epoch (the caller draws them: np.random.uniform() * T each, sim_plain.py:167); e0 the lowest
        eigenvalue of the observable (:294) -- required, the dense eigendecomposition does not exist here.
        Returns (trained coefficients, losses [n_epoch] = loss_energy - e0, state of the last epoch's evolution or None)."""
        p = self.problem
        if self.basis != 'BSpline':
            raise ValueError("the device-resident loop evaluates the B-spline ansatz only")
        if e0 is None:
            raise ValueError("pass e0 (lowest eigenvalue of the observable = min of its diagonal)")
        coeff = np.array(coeff0, dtype=np.float64, order="C")
        if coeff.shape[0] != len(p.terms):
            raise ValueError("one coefficient row per control term is required")
        s_all = np.ascontiguousarray(s_all, dtype=np.float64)
        if s_all.ndim == 1:
            s_all = s_all.reshape(-1, 1)
        losses = np.empty(s_all.shape[0])
        final = np.empty(1 << p.n, dtype=np.complex128) if want_state else None
        ms = ctypes.c_double()
        _lib.check(_lib.load().dq_ising_train(
            self.handle, len(p.terms), _lib.ptr(p.term_kind), _lib.ptr(p.term_index), _lib.ptr(p.omegas), _lib.ptr(p.h0_zz),
            p.h0_const, p.T, int(self.per_step), coeff.shape[1], _lib.ptr(coeff), s_all.shape[0], s_all.shape[1],
            _lib.ptr(s_all), float(lr), float(betas[0]), float(betas[1]), float(eps), float(r), float(e0), _lib.ptr(p.psi0),
            _lib.ptr(losses), _lib.ptr(final), ctypes.byref(ms)))
        self.train_device_ms = ms.value
        return coeff, losses, final

    # staged variant for benchmarking (inputs resident in HBM before the timed region)
    def stage(self, coeff, s_list, r=0.5):
        tables = self.sample_tables(coeff, np.asarray(s_list, dtype=np.float64).reshape(-1))
        _lib.check(_lib.load().dq_ising_grad_stage(*self._grad_args(tables, r)))
        self._staged_shape = (len(s_list), len(self.problem.terms), 2)

    def run_staged(self):
        _lib.check(_lib.load().dq_ising_grad_run_staged(self.handle))

    def fetch(self):
        out = np.empty(self._staged_shape)
        _lib.check(_lib.load().dq_ising_grad_fetch(self.handle, _lib.ptr(out)))
        return out
