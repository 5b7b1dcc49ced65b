"""The reference's comparison methods on the device -- SURVEY 8(f): finite-difference gradients and training
(compute_energy_grad_FD / train_energy_FD, sim_plain.py:308-412) and state-transfer training (train_fidelity,
sim_plain.py:414-475).

Both call qutip.mesolve for their forward runs (sim_plain.py:330,386,448): the Schroedinger equation of
H(t) = H0 + sum_i u_i(t) H_i integrated over ts = np.linspace(0, 1, n_step) -- t in [0, 1] whatever sim.T is, while the
pulses keep their t / T argument (:94).  Here that integral is a 4th-order commutator-free Magnus scheme built from the
exact-step engine (dq_dense_evolve_many): per interval h two exponentials
    exp(-i h (a1 H(t1) + a2 H(t2))) exp(-i h (a2 H(t1) + a1 H(t2))),  t_{1,2} = t + (1/2 -+ sqrt(3)/6) h,  a_{1,2} = (3 -+ 2 sqrt 3)/12
i.e. two engine steps of length h/2 with mixed pulse rows; the B-spline pulses have kinks at multiples of T / (n_basis - 2), so
the interval list is cut there and the scheme keeps its order.  All trajectories of a finite-difference gradient (two per
coefficient) run as ONE batch per segment.
"""
import math

import numpy as np

from . import _lib
from . import pulses
from .dense import MODES

_S3 = math.sqrt(3.0)
_A1, _A2 = (3 - 2 * _S3) / 12, (3 + 2 * _S3) / 12
_C1, _C2 = 0.5 - _S3 / 6, 0.5 + _S3 / 6


def kink_times(basis, n_basis, T, t0, t1):
    """Times in (t0, t1) where a pulse is not smooth: the supports of the quadratic bumps end at x = tau (b - 3) and tau b,
    x = t / T, tau = 1 / (n_basis - 2) (sim_plain.py:52-66)."""
    if basis != 'BSpline':
        return []
    step = T / (n_basis - 2)
    k0, k1 = int(math.floor(t0 / step)) - 1, int(math.ceil(t1 / step)) + 1
    return [k * step for k in range(k0, k1 + 1) if t0 + 1e-12 < k * step < t1 - 1e-12]


def segments(basis, n_basis, T, t0, t1, h):
    """[(a, b, n_sub)] covering [t0, t1], cut at the pulse kinks, n_sub Magnus intervals of length <= h each."""
    cuts = [t0] + kink_times(basis, n_basis, T, t0, t1) + [t1]
    return [(a, b, max(1, int(math.ceil((b - a) / h - 1e-9)))) for a, b in zip(cuts[:-1], cuts[1:])]


def magnus_rows(coeff, omegas, T, basis, a, b, n_sub):
    """Pulse rows [2 n_sub, n_H] and the engine step length for the segment [a, b] (see the module docstring)."""
    h = (b - a) / n_sub
    left = a + h * np.arange(n_sub)
    u1 = pulses.u_table(coeff, omegas, T, left + _C1 * h, basis)
    u2 = pulses.u_table(coeff, omegas, T, left + _C2 * h, basis)
    rows = np.empty((2 * n_sub, u1.shape[1]))
    rows[0::2] = 2 * (_A2 * u1 + _A1 * u2)            # applied first: the factor weighted towards the earlier node
    rows[1::2] = 2 * (_A1 * u1 + _A2 * u2)
    return rows, h / 2


def schrodinger(sim, coeffs, psi0s=None, t0=0.0, t1=1.0, h=2e-3, want_states=False):
    """Integrate d psi / dt = -i H(t) psi from t0 to t1 for a batch of coefficient sets coeffs [K, n_H, n_basis] (what
    qutip.mesolve(H, psi0, ts).states[-1] approximates at sim_plain.py:330-331).  Returns energies [K] (sim.M) or states."""
    coeffs = np.asarray(coeffs, dtype=np.float64)
    if coeffs.ndim == 2:
        coeffs = coeffs[None]
    K = coeffs.shape[0]
    psi = np.tile(np.asarray(sim.psi0, dtype=np.complex128), (K, 1)) if psi0s is None else \
        np.ascontiguousarray(np.asarray(psi0s, dtype=np.complex128).reshape(K, sim.dim))
    segs = segments(sim.basis, coeffs.shape[2], sim.T, t0, t1, h)
    lib = _lib.load()
    energies = None
    for si, (a, b, n_sub) in enumerate(segs):
        last = si == len(segs) - 1
        tabs = [magnus_rows(coeffs[k], sim.omegas, sim.T, sim.basis, a, b, n_sub) for k in range(K)]
        u = np.ascontiguousarray(np.concatenate([t[0] for t in tabs], axis=0))
        steps = np.full(K, 2 * n_sub, dtype=np.int32)
        dts = np.array([t[1] for t in tabs], dtype=np.float64)
        out = np.empty_like(psi)
        need_e = last and not want_states
        energies = np.empty(K) if need_e else None
        _lib.check(lib.dq_dense_evolve_many(sim.ctx.handle, sim.dim, _lib.ptr(sim.H0), sim.n_H, _lib.ptr(sim.Hs),
                                            _lib.ptr(sim.M) if need_e else None, K, _lib.ptr(psi), _lib.ptr(steps), _lib.ptr(dts),
                                            _lib.ptr(u), MODES["exact"], _lib.ptr(energies), None if need_e else _lib.ptr(out)))
        psi = out
    return psi if want_states else energies


def grad_fd(sim, coeff, delta=1e-3, is_noisy=False, t_end=1.0, h=2e-3):
    """compute_energy_grad_FD (sim_plain.py:308-353): central differences of the final energy in every coefficient, all
    2 n_H n_basis forward runs as one device batch.  is_noisy: the reference's noise on every forward energy (:339-340), drawn
    from the global stream in its order (per coefficient: E_p then E_m)."""
    coeff = np.asarray(coeff, dtype=np.float64)
    n_H, n_basis = coeff.shape
    batch = np.repeat(coeff[None], 2 * n_H * n_basis, axis=0)
    k = 0
    for i in range(n_H):
        for j in range(n_basis):
            batch[k, i, j] = coeff[i][j] + delta              # :346
            batch[k + 1, i, j] = coeff[i][j] - delta          # :349
            k += 2
    E = schrodinger(sim, batch, t0=0.0, t1=t_end, h=h)
    if is_noisy:
        for k in range(len(E)):
            E[k] += np.random.normal(scale=np.abs(E[k]) / 5)
    return ((E[0::2] - E[1::2]) / delta / 2.0).reshape(n_H, n_basis)   # :351


class FDTrainer(object):
    """train_energy_FD (sim_plain.py:355-412) on a DenseSimulator: same draws, torch Adam, finite-difference gradients."""

    def __init__(self, backend, n_basis=6, n_epoch=200, lr=2e-2, delta=1e-3, is_noisy=False, h=2e-3):
        self.backend, self.n_basis, self.n_epoch, self.lr, self.delta, self.is_noisy, self.h = backend, n_basis, n_epoch, lr, delta, is_noisy, h
        self.losses_energy, self.final_state, self.spectral_coeff = [], None, None

    def train_energy_FD(self):
        import torch
        b = self.backend
        coeff = np.random.normal(0, 1e-3, [b.n_H, self.n_basis])              # :368
        self.spectral_coeff = torch.tensor(coeff, requires_grad=True)
        optimizer = torch.optim.Adam([self.spectral_coeff], lr=self.lr)        # :376
        e0 = float(np.linalg.eigvalsh(b.M)[0])                                 # :402
        self.losses_energy = []
        for epoch in range(1, self.n_epoch + 1):
            c = self.spectral_coeff.detach().numpy().copy()
            self.final_state = schrodinger(b, c, t0=0.0, t1=1.0, h=self.h, want_states=True)[0]     # :386-387
            loss_energy = float(b.energy(self.final_state))                    # :389
            if self.is_noisy:
                loss_energy += np.random.normal(scale=np.abs(loss_energy) / 5)  # :390-391
            optimizer.zero_grad()
            self.spectral_coeff.grad = torch.from_numpy(grad_fd(b, c, self.delta, self.is_noisy, 1.0, self.h))   # :398-399
            optimizer.step()
            self.losses_energy.append(loss_energy - e0)
        return self.spectral_coeff


class FidelityTrainer(object):
    """train_fidelity (sim_plain.py:414-475) on a DenseSimulator: per epoch and per (initial, target) pair, the loss
    1 - |<target|psi(1)>|^2 from a forward run over t in [0, 1] (:448-455; logged only) and one stochastic parameter-shift
    sample of the projector observable M = |target><target| with coeff = -1 over [0, T] (:461), one Adam step each (:463-464)."""

    def __init__(self, backend, n_basis=6, n_epoch=200, lr=2e-2, is_noisy=False, h=2e-3):
        self.backend, self.n_basis, self.n_epoch, self.lr, self.is_noisy, self.h = backend, n_basis, n_epoch, lr, is_noisy, h
        self.losses_energy, self.spectral_coeff = [], None

    def train_fidelity(self, initial_states, target_states):
        import torch
        b = self.backend
        coeff = np.random.normal(0, 1, [b.n_H, self.n_basis])                  # :425
        self.spectral_coeff = torch.tensor(coeff, requires_grad=True)
        optimizer = torch.optim.Adam([self.spectral_coeff], lr=self.lr)        # :432
        inits = [np.asarray(p.full() if hasattr(p, "full") else p, dtype=np.complex128).reshape(-1) for p in initial_states]
        targets = [np.asarray(p.full() if hasattr(p, "full") else p, dtype=np.complex128).reshape(-1) for p in target_states]
        saved = (b.M, b.psi0)
        self.losses_energy = []
        try:
            for epoch in range(1, self.n_epoch + 1):
                batch_losses = []
                for psi0, psi1 in zip(inits, targets):
                    c = self.spectral_coeff.detach().numpy().copy()
                    b.psi0 = psi0
                    b.M = np.outer(psi1, psi1.conj())                          # M = psi1 * psi1.dag(), :447
                    final = schrodinger(b, c, t0=0.0, t1=1.0, h=self.h, want_states=True)[0]
                    inner = float(np.abs(np.vdot(psi1, final)) ** 2)           # M.matrix_element(final, final), :451
                    if self.is_noisy:
                        inner += np.random.normal(scale=np.abs(inner) / 5)     # :452-454
                    optimizer.zero_grad()
                    s = np.random.uniform() * b.T                              # :167
                    grads = b.grad_samples(c, [s], coeff_sign=-1.0, is_noisy=self.is_noisy)     # :461
                    self.spectral_coeff.grad = torch.from_numpy(np.asarray(grads[0]))
                    optimizer.step()
                    batch_losses.append(1 - inner)                             # :455,466
                self.losses_energy.append(np.array(batch_losses).mean())       # :468,474
        finally:
            b.M, b.psi0 = saved
        return self.spectral_coeff
