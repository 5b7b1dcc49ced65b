"""Dense path (dim <= 1024) behind the reference's Python-facing interfaces.

  DenseSimulator        explicit arrays: evolve / grad_samples with the live `exact` step
                        psi <- expm(-i dt (H0 + sum_i u_i(t_k) H_i)) psi  (sim_plain.py:135-150), or the
                        disabled per-term product (`split`, diffqc.cc:155-164).
  solver_for(sim)       a drop-in for SimulatorPlain.my_solver (sim_plain.py:43): same signature
                        solver(H_list, psi0, T0, T) -> Qobj, evaluates the reference's own pulse
                        closures on the reference's own step grid and runs the steps on the GPU.
  estimator_for(sim)    compute_energy_grad_MC (sim_plain.py:156-231) with the 1 + 2*n_Hs evolutions
                        batched into one device call; consumes np.random exactly like the reference.

All device work goes through the C ABI (include/diffqc_b200.h, dq_dense_*); there is no CPU path.
"""
import ctypes

import numpy as np

from . import _lib
from . import pulses

MODES = {"exact": 0, "split": 1}


def _c128(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.complex128))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _full(q):
    """ndarray of a qutip-like object (has .full()) or of anything array-like."""
    return np.asarray(q.full() if hasattr(q, "full") else q, dtype=np.complex128)


def dense_evolve(ctx, H0, Hs, u, dt, psi, mode="exact"):
    """psi: [batch, dim] (or [dim]); u: [n_steps, n_H] host-evaluated pulse values."""
    H0 = _c128(H0)
    dim = H0.shape[0]
    Hs = _c128(Hs, (-1, dim, dim))
    u = np.asarray(u, dtype=np.float64)
    n_steps = u.shape[0] if u.ndim == 2 else u.size // max(1, Hs.shape[0])
    u = np.ascontiguousarray(u.reshape(n_steps, Hs.shape[0]))
    psi = _c128(psi)
    single = psi.ndim == 1
    psi = psi.reshape(-1, dim)
    out = np.empty_like(psi)
    _lib.check(_lib.load().dq_dense_evolve(ctx.handle, dim, _lib.ptr(H0), Hs.shape[0], _lib.ptr(Hs), _lib.ptr(u),
                                           n_steps, float(dt), MODES[mode], psi.shape[0], _lib.ptr(psi),
                                           _lib.ptr(out)))
    return out[0] if single else out


class DenseSimulator(object):
    """H(t) = H0 + sum_i u_i(t) H_i on explicit dense matrices, pulses u_i as SimulatorPlain.generate_u
    builds them (sim_plain.py:73-99): u_i(t) = omega_i (2 sigma(sum_j c_ij phi_j(t/T)) - 1)."""

    def __init__(self, H0, Hs, omegas, T, M=None, psi0=None, per_step=10, basis="BSpline", device=0, mode="exact"):
        self.H0 = _c128(H0)
        self.dim = self.H0.shape[0]
        self.Hs = _c128(Hs, (-1, self.dim, self.dim))
        self.n_H = self.Hs.shape[0]
        self.omegas = np.asarray(omegas, dtype=np.float64)
        if len(self.omegas) != self.n_H:
            raise ValueError("one omega per control term is required")
        self.T = float(T)
        self.M = None if M is None else _c128(M, (self.dim, self.dim))
        self.psi0 = None if psi0 is None else _c128(psi0, (self.dim,))
        self.per_step = per_step
        self.basis = basis
        self.mode = mode
        self.ctx = _lib.Context.get(device)
        # dim <= 16 with the B-spline ansatz: the device evaluates the pulse rows itself (dq_dense_grad_times); only the
        # sample times and the coefficients cross the bus.  False = host tables (pulses.u_table) through dq_dense_grad.
        self.device_tables = True
        self.measure = None            # (weights, evals, bases) of sim.Pauli_M: shot sampling, set_measurement()

    def set_measurement(self, pauli_m):
        """Pauli_M of the reference (demo_maxcut.py:47-65): [[matrix, weight, (evals, estates)], ...] -- the measurement
        bases of stochastic_measure (sim_plain.py:101-117)."""
        self.measure = pulses.measurement_bases(pauli_m)

    def outcome_probs(self, kets):
        """distr[j] = |<e_mj|ket>|^2 for every ket and measurement basis (sim_plain.py:105-109), on the device:
        [n_kets, n_meas, dim]."""
        if self.measure is None:
            raise ValueError("set_measurement(Pauli_M) first")
        bases = self.measure[2]
        kets = _c128(kets).reshape(-1, self.dim)
        out = np.empty((kets.shape[0], bases.shape[0], self.dim))
        _lib.check(_lib.load().dq_dense_outcome_probs(self.ctx.handle, self.dim, kets.shape[0], _lib.ptr(kets), bases.shape[0],
                                                      _lib.ptr(bases), _lib.ptr(out)))
        return out

    def stochastic_measure(self, psi, per_Pauli=100):
        """SimulatorPlain.stochastic_measure (sim_plain.py:101-117): distributions on the device, draws from np.random."""
        return pulses.stochastic_measure_from_probs(self.outcome_probs(psi)[0], self.measure[0], self.measure[1], per_Pauli)

    def shifted_outcome_probs(self, coeff, s_list, r=0.5, mode=None):
        """Outcome distributions of every shifted ket: [B, n_H, 2, n_meas, dim] (dq_dense_grad_probs)."""
        if self.measure is None or self.psi0 is None:
            raise ValueError("shifted_outcome_probs needs set_measurement(Pauli_M) and psi0")
        s_list = np.asarray(s_list, dtype=np.float64).reshape(-1)
        pre_n, pre_dt, pre_ts = pulses.step_grids(0.0, s_list, self.per_step)
        suf_n, suf_dt, suf_ts = pulses.step_grids(s_list, self.T, self.per_step)
        pre_u = np.ascontiguousarray(pulses.u_table(coeff, self.omegas, self.T, pre_ts, self.basis).reshape(-1, self.n_H))
        suf_u = np.ascontiguousarray(pulses.u_table(coeff, self.omegas, self.T, suf_ts, self.basis).reshape(-1, self.n_H))
        return self._probs_call(s_list.size, pre_n, pre_dt, pre_u, suf_n, suf_dt, suf_u, r, mode)

    def _probs_call(self, B, pre_n, pre_dt, pre_u, suf_n, suf_dt, suf_u, r, mode):
        bases = self.measure[2]
        pre_dt = np.ascontiguousarray(pre_dt, dtype=np.float64); suf_dt = np.ascontiguousarray(suf_dt, dtype=np.float64)
        out = np.empty((B, self.n_H, 2, bases.shape[0], self.dim))
        _lib.check(_lib.load().dq_dense_grad_probs(
            self.ctx.handle, self.dim, _lib.ptr(self.H0), self.n_H, _lib.ptr(self.Hs), _lib.ptr(self.psi0), float(r), B,
            _lib.ptr(pre_n), _lib.ptr(pre_dt), _lib.ptr(pre_u), _lib.ptr(suf_n), _lib.ptr(suf_dt), _lib.ptr(suf_u),
            MODES[mode or self.mode], bases.shape[0], _lib.ptr(bases), _lib.ptr(out)))
        return out

    def _times_call(self, coeff, s_list, r, mode, want_u=False):
        """dq_dense_grad_times: energies [B, n_H, 2] (and the device-built pulse table when want_u)."""
        coeff = np.ascontiguousarray(coeff, dtype=np.float64)
        s_list = np.ascontiguousarray(s_list, dtype=np.float64)
        out = np.empty((len(s_list), self.n_H, 2))
        u = None
        if want_u:
            rows = int((self.per_step * (s_list + 1)).astype(np.int64).sum() + (self.per_step * ((self.T - s_list) + 1)).astype(np.int64).sum())
            u = np.empty((rows, self.n_H))
        _lib.check(_lib.load().dq_dense_grad_times(
            self.ctx.handle, self.dim, _lib.ptr(self.H0), self.n_H, _lib.ptr(self.Hs), _lib.ptr(self.M), _lib.ptr(self.psi0),
            float(r), len(s_list), _lib.ptr(s_list), self.T, int(self.per_step), _lib.ptr(coeff), _lib.ptr(self.omegas),
            coeff.shape[1], MODES[mode or self.mode], _lib.ptr(out), _lib.ptr(u)))
        return out, u

    def train_energy_device(self, coeff0, s_all, lr=2e-2, betas=(0.9, 0.999), eps=1e-8, r=0.5, e0=None, mode=None):
        """The whole train_energy loop (sim_plain.py:245-305) on the device: dq_dense_train.  s_all [n_epoch, K] are the
        sample times of every epoch (the caller draws them: np.random.uniform() * T each, sim_plain.py:167).
        Returns (trained coefficients, losses [n_epoch] = loss_energy - e0, state of the last epoch's evolution)."""
        if self.M is None or self.psi0 is None:
            raise ValueError("train_energy_device needs M and psi0")
        if self.basis != 'BSpline':
            raise ValueError("the device-resident loop evaluates the B-spline ansatz only")
        coeff = np.array(coeff0, dtype=np.float64, order="C")
        s_all = np.ascontiguousarray(s_all, dtype=np.float64)
        if s_all.ndim == 1:
            s_all = s_all.reshape(-1, 1)
        if e0 is None:
            e0 = float(np.linalg.eigvalsh(self.M)[0])                  # M.eigenenergies()[0], sim_plain.py:294
        losses = np.empty(s_all.shape[0])
        final = np.empty(self.dim, dtype=np.complex128)
        _lib.check(_lib.load().dq_dense_train(
            self.ctx.handle, self.dim, _lib.ptr(self.H0), self.n_H, _lib.ptr(self.Hs), _lib.ptr(self.M), _lib.ptr(self.psi0),
            _lib.ptr(self.omegas), self.T, int(self.per_step), coeff.shape[1], _lib.ptr(coeff), s_all.shape[0], s_all.shape[1],
            _lib.ptr(s_all), float(lr), float(betas[0]), float(betas[1]), float(eps), float(r), float(e0),
            MODES[mode or self.mode], _lib.ptr(losses), _lib.ptr(final)))
        return coeff, losses, final

    def stat(self, name):
        v = ctypes.c_double()
        _lib.check(_lib.load().dq_dense_last_stat(self.ctx.handle, name.encode(), ctypes.byref(v)))
        return v.value

    def set_option(self, name, value):
        _lib.check(_lib.load().dq_dense_set_option(self.ctx.handle, name.encode(), int(value)))
        if name == "strategy":
            self._strategy = int(value)

    def ctx_strategy_allows_resident(self):
        """The device-side tables belong to the resident engine: not when a GEMM strategy (0, 1, 2) is forced."""
        return getattr(self, "_strategy", -1) in (-1, 3)

    def evolve(self, coeff, T0, T1, psi0=None, mode=None):
        """SimulatorPlain.trotter (sim_plain.py:119-153) on the reference's step grid."""
        n_steps, dt, ts = pulses.step_grid(T0, T1, self.per_step)
        psi = self.psi0 if psi0 is None else psi0
        if n_steps == 0:
            return _c128(psi).copy()
        u = pulses.u_table(coeff, self.omegas, self.T, ts, self.basis)
        return dense_evolve(self.ctx, self.H0, self.Hs, u, dt, psi, mode or self.mode)

    def energy(self, psi):
        psi = _c128(psi, (self.dim,))
        return (psi.conj() @ self.M @ psi).real

    def shifted_energies(self, coeff, s_list, r=0.5, mode=None):
        """energies[b, i, 0|1] = <M> of the (+, -) shifted trajectories (ps_p, ps_m: sim_plain.py:205,215)."""
        if self.M is None or self.psi0 is None:
            raise ValueError("shifted_energies needs M and psi0")
        s_list = np.asarray(s_list, dtype=np.float64).reshape(-1)
        if (self.device_tables and self.basis == 'BSpline' and self.dim <= 16 and isinstance(self.per_step, (int, np.integer))
                and self.ctx_strategy_allows_resident()):
            return self._times_call(coeff, s_list, r, mode)[0]
        # all step grids and pulse tables of the batch in four vectorised calls (bit-identical to per-sample calls)
        pre_n, pre_dt, pre_ts = pulses.step_grids(0.0, s_list, self.per_step)
        suf_n, suf_dt, suf_ts = pulses.step_grids(s_list, self.T, self.per_step)
        pre_u = np.ascontiguousarray(pulses.u_table(coeff, self.omegas, self.T, pre_ts, self.basis).reshape(-1, self.n_H))
        suf_u = np.ascontiguousarray(pulses.u_table(coeff, self.omegas, self.T, suf_ts, self.basis).reshape(-1, self.n_H))
        pre_dt = np.ascontiguousarray(pre_dt); suf_dt = np.ascontiguousarray(suf_dt)
        out = np.empty((len(s_list), self.n_H, 2))
        _lib.check(_lib.load().dq_dense_grad(
            self.ctx.handle, self.dim, _lib.ptr(self.H0), self.n_H, _lib.ptr(self.Hs), _lib.ptr(self.M),
            _lib.ptr(self.psi0), float(r), len(s_list), _lib.ptr(pre_n), _lib.ptr(pre_dt), _lib.ptr(pre_u),
            _lib.ptr(suf_n), _lib.ptr(suf_dt), _lib.ptr(suf_u), MODES[mode or self.mode], _lib.ptr(out)))
        return out

    def grad_samples(self, coeff, s_list, r=0.5, coeff_sign=1.0, return_energies=False, mode=None, is_noisy=False,
                     sampling_measure=False, per_Pauli=100):
        """Per-sample gradients of compute_energy_grad_MC (sim_plain.py:156-231) at explicit times.  is_noisy: the
        reference's measurement noise on every shifted energy (pulses.add_measurement_noise).  sampling_measure: every
        shifted energy by shot sampling (stochastic_measure, :202-203,212-213) -- outcome distributions from the device,
        np.random.choice draws on the host in the reference's order."""
        s_list = np.asarray(s_list, dtype=np.float64).reshape(-1)
        if sampling_measure:
            probs = self.shifted_outcome_probs(coeff, s_list, r, mode)
            en = pulses.sampled_shifted_energies(probs, self.measure[0], self.measure[1], per_Pauli, is_noisy)
        else:
            en = self.shifted_energies(coeff, s_list, r, mode)
            if is_noisy:
                pulses.add_measurement_noise(en)
        ps = coeff_sign * ((1 + r ** 2) / 2 / r * (en[:, :, 1] - en[:, :, 0]))
        grads = ps[:, :, None] * pulses.dudc_tables(coeff, self.omegas, self.T, s_list, self.basis)
        return (grads, en) if return_energies else grads


# ---- drop-ins for SimulatorPlain --------------------------------------------------------------------

def _split_H(H_):
    """[H0, [H_1, u_1], ...] as train_energy builds it (sim_plain.py:272-274) -> arrays + closures."""
    H0 = _full(H_[0])
    Hs = np.array([_full(h[0]) for h in H_[1:]], dtype=np.complex128).reshape(-1, H0.shape[0], H0.shape[0])
    fs = [h[1] for h in H_[1:]]
    return H0, Hs, fs


def solver_for(sim, device=0, mode="exact"):
    """GPU-backed replacement for SimulatorPlain.trotter; assign it to `sim.my_solver`.
    Step grid and pulse sampling follow sim_plain.py:123,133-150 literally: n_steps by int truncation
    without abs, left-end sampling, t accumulated by `t += dt`, closures called as u(t, None)."""
    ctx = _lib.Context.get(device)

    def solver(H_, psi0_, T0, T):
        H0, Hs, fs = _split_H(H_)
        n_steps, dt, ts = pulses.step_grid(T0, T, sim.per_step)
        psi0 = _full(psi0_).reshape(-1)
        if n_steps <= 0:
            raise ZeroDivisionError("float division by zero")       # sim_plain.py:133 with n_steps == 0
        u = np.array([[f(t, None) for f in fs] for t in ts], dtype=np.float64).reshape(n_steps, len(fs))
        out = dense_evolve(ctx, H0, Hs, u, dt, psi0, mode)
        return type(psi0_)(out) if hasattr(psi0_, "full") else out

    return solver


def estimator_for(sim, device=0, mode="exact"):
    """Batched replacement for SimulatorPlain.compute_energy_grad_MC (sim_plain.py:156-231).
    Same signature and return type (torch.float64 [n_Hs, n_basis]); draws s = np.random.uniform() * T
    exactly where the reference does (:167) unless `s` is passed.  `sim.is_noisy` adds the reference's measurement noise
    (:207-208,217-218) from the same global stream; `sim.sampling_measure` replaces every shifted energy by
    stochastic_measure (:101-117,202-203,212-213): the outcome distributions of sim.Pauli_M's eigenbases come from the device
    (dq_dense_grad_probs), the np.random.choice draws happen here in the reference's order."""
    ctx = _lib.Context.get(device)

    def compute_energy_grad_MC(M, H_, psi0_, coeff=1.0, s=None):
        import torch
        H0, Hs, fs = _split_H(H_)
        if s is None:
            s = np.random.uniform() * sim.T
        c = sim.spectral_coeff.detach().cpu().numpy()
        ds = DenseSimulator(H0, Hs, sim.omegas, sim.T, M=_full(M), psi0=_full(psi0_).reshape(-1),
                            per_step=sim.per_step, basis=sim.basis, device=device, mode=mode)
        ds.ctx = ctx
        # pulse values come from the closures the caller built (they may differ from sim.spectral_coeff)
        def table(T0, T1):
            n, dt, ts = pulses.step_grid(T0, T1, sim.per_step)
            return n, dt, np.array([[f(t, None) for f in fs] for t in ts], dtype=np.float64).reshape(n, len(fs))
        pn, pdt, pu = table(0, s)
        sn, sdt, su = table(s, sim.T)
        en = np.empty((1, len(fs), 2))
        r = 1 / 2
        if getattr(sim, "sampling_measure", False):
            ds.set_measurement(sim.Pauli_M)
            probs = ds._probs_call(1, np.array([pn], dtype=np.int32), np.array([pdt]), np.ascontiguousarray(pu),
                                   np.array([sn], dtype=np.int32), np.array([sdt]), np.ascontiguousarray(su), r, mode)
            en = pulses.sampled_shifted_energies(probs, ds.measure[0], ds.measure[1], 100, getattr(sim, "is_noisy", False))
        else:
            _lib.check(_lib.load().dq_dense_grad(
                ctx.handle, ds.dim, _lib.ptr(ds.H0), ds.n_H, _lib.ptr(ds.Hs), _lib.ptr(ds.M), _lib.ptr(ds.psi0), r, 1,
                _lib.ptr(np.array([pn], dtype=np.int32)), _lib.ptr(np.array([pdt])), _lib.ptr(np.ascontiguousarray(pu)),
                _lib.ptr(np.array([sn], dtype=np.int32)), _lib.ptr(np.array([sdt])), _lib.ptr(np.ascontiguousarray(su)),
                MODES[mode], _lib.ptr(en)))
            if getattr(sim, "is_noisy", False):
                pulses.add_measurement_noise(en)
        ps = coeff * ((1 + r ** 2) / 2 / r * (en[0, :, 1] - en[0, :, 0]))
        grad = ps[:, None] * pulses.dudc_table(c, sim.omegas, sim.T, s, sim.basis)
        return torch.from_numpy(grad)

    return compute_energy_grad_MC
