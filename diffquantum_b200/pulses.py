"""Host-side pulse tables and time grids, vectorised over steps and samples.

Mirrors, in the reference's own operation order, the scalar Python code of
  SimulatorPlain.get_func_bspline  sim_plain.py:52-70   (open-support quadratic bumps)
  SimulatorPlain.generate_u        sim_plain.py:73-99   (sequential sum over j, sigmoid, omega)
  the dDdv autograd block          sim_plain.py:169-184 (closed form of the same derivative)
  the step grid of trotter         sim_plain.py:123,133-134,150 / diffqc.cc:182-184,199
  diffqc f_u / my_expit / bspline  diffqc.cc:75-135
The device only ever sees angles; all pulse arithmetic stays here in float64.
"""
import math

import numpy as np
from scipy.special import eval_legendre


def step_grid(T0, T, per_step, use_abs=False):
    """n_steps, dt and the accumulated left-end times (t += dt, not T0 + k*dt)."""
    span = abs(T - T0) if use_abs else (T - T0)
    n_steps = int(per_step * (span + 1))
    if n_steps <= 0:
        return 0, 0.0, np.zeros(0)
    dt = (T - T0) / n_steps
    # np.add.accumulate is a sequential left fold, i.e. exactly the reference's repeated `t += dt`
    ts = np.add.accumulate(np.concatenate(([float(T0)], np.full(n_steps - 1, dt))))
    return n_steps, dt, ts


def bspline_table(n_basis, x):
    """phi[b](x) for all b: array [len(x), n_basis] (sim_plain.py:52-70 == diffqc.cc:82-93)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    tau = 1. / (n_basis - 2)
    out = np.zeros((x.size, n_basis))
    norm_factor = -(1.5 * tau) ** 2
    for b in range(n_basis):
        tau_b = tau * (b - 1.5)
        l = tau_b - 1.5 * tau
        r = tau_b + 1.5 * tau
        inside = ~((x >= r) | (x <= l))
        out[inside, b] = (x[inside] - l) * (x[inside] - r) / norm_factor
    return out


def legendre_table(n_basis, y):
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    return np.stack([eval_legendre(j, y) for j in range(n_basis)], axis=1)


def basis_table(basis, n_basis, ts, T):
    ts = np.asarray(ts, dtype=np.float64)
    if basis == 'BSpline':
        return bspline_table(n_basis, ts / T)
    if basis == 'Legendre':
        return legendre_table(n_basis, 2 * ts / T - 1)
    if basis == 'poly':
        # sim_plain.py:86-87: (t - 0.5) ** j on the raw time; Python float powers, element by element (NumPy's vectorised pow
        # may round differently)
        return np.array([[(float(t) - 0.5) ** j for j in range(n_basis)] for t in ts.reshape(-1)], dtype=np.float64).reshape(-1, n_basis)
    if basis == 'Fourier':
        # sim_plain.py:90-92: columns [cos(2 pi j t) for j < n | sin(2 pi j t) for j < n], n = int(n_basis / 2); scalar calls as
        # in the reference (the vectorised sin / cos kernels may round differently)
        n = int(n_basis / 2)
        out = np.zeros((ts.size, n_basis))
        for k, t in enumerate(ts.reshape(-1)):
            t = float(t)
            for j in range(n):
                out[k, j] = np.cos(2 * np.pi * j * t)
                out[k, j + n] = np.sin(2 * np.pi * j * t)
        return out
    raise ValueError("unsupported basis %r (sim_plain.py:84-94 knows BSpline, Legendre, poly, Fourier)" % (basis,))


def step_grids(T0s, Ts, per_step):
    """step_grid for many (T0, T) pairs at once: n[b], dt[b] and the accumulated times of all pairs packed in pair
    order (sum(n) values).  Row-wise np.add.accumulate is the same sequential `t += dt` fold as step_grid, so the
    packed times are bit-identical to per-pair calls."""
    T0s, Ts = np.broadcast_arrays(np.asarray(T0s, dtype=np.float64), np.asarray(Ts, dtype=np.float64))
    T0s, Ts = T0s.reshape(-1), Ts.reshape(-1)
    n = (per_step * ((Ts - T0s) + 1)).astype(np.int64)              # int() truncation (sim_plain.py:123)
    n = np.maximum(n, 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        dt = np.where(n > 0, (Ts - T0s) / np.maximum(n, 1), 0.0)
    width = int(n.max()) if n.size else 0
    if width == 0:
        return n.astype(np.int32), dt, np.zeros(0)
    grid = np.repeat(dt[:, None], width, axis=1)
    grid[:, 0] = T0s
    ts = np.add.accumulate(grid, axis=1)
    keep = np.arange(width)[None, :] < n[:, None]
    return n.astype(np.int32), dt, ts[keep]


def _sigmoid(a):
    return 1 / (1 + np.exp(-a))


def u_table(coeff, omegas, T, ts, basis='BSpline'):
    """u[k, i] = omega_i (2 sigma(sum_j c_ij phi_j(t_k)) - 1)   (sim_plain.py:81-98)."""
    coeff = np.asarray(coeff, dtype=np.float64)
    phi = basis_table(basis, coeff.shape[1], ts, T)             # [K, n_basis]
    a = np.zeros((phi.shape[0], coeff.shape[0]))
    if basis == 'Fourier':                                      # :90-92: u += (c_j cos + c_{j+n} sin), the pair summed first
        n = int(coeff.shape[1] / 2)
        for j in range(n):
            a = a + (coeff[None, :, j] * phi[:, j:j + 1] + coeff[None, :, j + n] * phi[:, j + n:j + n + 1])
    else:
        for j in range(coeff.shape[1]):                         # same summation order as the loop at :85
            a = a + phi[:, j:j + 1] * coeff[None, :, j]
    return (_sigmoid(a) * 2 - 1) * np.asarray(omegas, dtype=np.float64)[None, :]


def _estimator_basis(basis):
    """compute_energy_grad_MC only defines its derivative block for two bases (sim_plain.py:171-176): with 'poly' or 'Fourier'
    the reference itself stops at :178 (coeff_A is never assigned).  Same behaviour here."""
    if basis not in ('BSpline', 'Legendre'):
        raise ValueError("compute_energy_grad_MC is undefined for basis %r in the reference (sim_plain.py:171-178 assigns coeff_A "
                         "for 'Legendre' and 'BSpline' only); the evolution (generate_u) supports it" % (basis,))


def dudc_table(coeff, omegas, T, s, basis='BSpline'):
    """dDdv[i, j] = d u_i(s) / d c_ij = omega_i 2 sigma'(A_i) phi_j(s)  (sim_plain.py:169-184)."""
    _estimator_basis(basis)
    coeff = np.asarray(coeff, dtype=np.float64)
    phi = basis_table(basis, coeff.shape[1], [s], T)[0]
    a = np.zeros(coeff.shape[0])
    for j in range(coeff.shape[1]):
        a = a + coeff[:, j] * phi[j]
    sg = _sigmoid(a)
    return (np.asarray(omegas, dtype=np.float64) * 2.0 * sg * (1.0 - sg))[:, None] * phi[None, :]


def dudc_tables(coeff, omegas, T, s_list, basis='BSpline'):
    """dudc_table for many sample times: [len(s_list), n_H, n_basis], bit-identical to per-sample calls."""
    _estimator_basis(basis)
    coeff = np.asarray(coeff, dtype=np.float64)
    s_list = np.asarray(s_list, dtype=np.float64).reshape(-1)
    phi = basis_table(basis, coeff.shape[1], s_list, T)             # [B, n_basis]
    a = np.zeros((s_list.size, coeff.shape[0]))
    for j in range(coeff.shape[1]):
        a = a + coeff[None, :, j] * phi[:, j:j + 1]
    sg = _sigmoid(a)
    return (np.asarray(omegas, dtype=np.float64)[None, :] * 2.0 * sg * (1.0 - sg))[:, :, None] * phi[:, None, :]


def add_measurement_noise(en):
    """is_noisy of the reference (sim_plain.py:207-208,217-218): every shifted energy ps gets
    np.random.normal(scale=|ps|/5) added, drawn from the GLOBAL NumPy stream in the reference's order - per sample, per
    control, ps_p (column 0) before ps_m (column 1).  en: [B, n_H, 2], modified in place and returned."""
    for b in range(en.shape[0]):
        for i in range(en.shape[1]):
            for k in range(2):
                en[b, i, k] += np.random.normal(scale=np.abs(en[b, i, k]) / 5)
    return en


# ---- native twin (diffqc.cc) -------------------------------------------------------------------

def _expit_cc(x):
    out = 1 / (1 + np.exp(-np.clip(x, -700, 700)))
    out = np.where(x > 32., 1., out)           # diffqc.cc:75-80
    return np.where(x < -32., 0., out)


def f_u_table(channels, duration, func_type, vv, ts):
    """u[k, h] of the native twin (diffqc.cc:95-135).  channels: list (per term) of lists of
    [_, omega, w, idx]; vv: [2][n_param][n_basis]."""
    vv = np.asarray(vv, dtype=np.float64)
    ts = np.asarray(ts, dtype=np.float64)
    n_basis = vv.shape[2]
    if func_type == 0:
        fv = legendre_table(n_basis, 2 * ts / duration - 1)
    else:
        fv = bspline_table(n_basis, ts / duration)
    out = np.zeros((ts.size, len(channels)))
    for h, chans in enumerate(channels):
        for chan in chans:
            omega, w = float(chan[1]), float(chan[2])
            idx = int(math.floor(abs(chan[3]) + 0.5)) * (1 if chan[3] >= 0 else -1)   # C round()
            if not 0 <= idx < vv.shape[1]:
                raise ValueError("channel parameter index %d outside vv (n_param=%d)" % (idx, vv.shape[1]))
            A = np.zeros(ts.size)
            B = np.zeros(ts.size)
            for j in range(n_basis):
                A = A + vv[0, idx, j] * fv[:, j]
                B = B + vv[1, idx, j] * fv[:, j]
            N = np.sqrt(A * A + B * B)
            small = np.abs(N - 0.0) < 0.000001
            Ns = np.where(small, 1.0, N)
            term = omega * (2 * _expit_cc(Ns) - 1) / Ns * (np.cos(w * ts) * A + np.sin(w * ts) * B)
            out[:, h] += np.where(small, 0.0, term)
    return out


def f_u_table_lib(channels, duration, func_type, vv, ts):
    """The same table from the library's own host routine (dq_pulse_f_u_table, the code dq_dense_trotter evaluates its
    step grid with; diffqc.cc:95-135).  Needs the shared library, not a device."""
    from . import _lib
    counts = np.array([len(c) for c in channels], dtype=np.int32)
    flat = np.ascontiguousarray(np.array([ch for c in channels for ch in c], dtype=np.float64).reshape(-1, 4))
    v = np.ascontiguousarray(np.asarray(vv, dtype=np.float64))
    if v.ndim != 3 or v.shape[0] != 2:
        raise ValueError("vv must be [2][n_param][n_basis]")
    ts = np.ascontiguousarray(ts, dtype=np.float64).reshape(-1)
    out = np.empty((ts.size, len(channels)))
    _lib.check(_lib.load().dq_pulse_f_u_table(len(channels), _lib.ptr(counts), _lib.ptr(flat), float(duration), int(func_type),
                                              _lib.ptr(v), v.shape[1], v.shape[2], ts.size, _lib.ptr(ts), _lib.ptr(out)))
    return out


# ---- shot sampling (SimulatorPlain.stochastic_measure, sim_plain.py:101-117) ----------------------------------------------------

def measurement_bases(pauli_m):
    """sim.Pauli_M as the reference's callers build it (demo_maxcut.py:47-65): entries [matrix, weight, (evals, estates)].
    Returns (weights [n_meas], evals [n_meas, dim], bases [n_meas, dim, dim] complex128 with bases[m, j] = eigenvector j)."""
    weights, evals, bases = [], [], []
    for entry in pauli_m:
        w, (ev, es) = entry[1], entry[2]
        weights.append(w)
        evals.append(np.asarray(ev, dtype=np.float64).reshape(-1))
        bases.append(np.stack([np.asarray(e.full() if hasattr(e, "full") else e, dtype=np.complex128).reshape(-1) for e in es]))
    return weights, np.array(evals), np.ascontiguousarray(np.array(bases, dtype=np.complex128))


def stochastic_measure_from_probs(probs, weights, evals, per_Pauli=100):
    """stochastic_measure (sim_plain.py:101-117) given the outcome distributions probs[m, j] = |<e_mj|psi>|^2 (:105-109) the
    device computed: one np.random.choice(dim, per_Pauli, p=distr) per Pauli term from the GLOBAL stream (:112), then
    ans += weight * evals[j] * freq_j / per_Pauli over j in order (:113-116; outcomes never drawn add exactly 0)."""
    ans = 0
    for m in range(len(weights)):
        distr = probs[m]
        res = np.random.choice(len(distr), per_Pauli, p=distr)
        freq = np.bincount(res, minlength=len(distr))
        weight = weights[m]
        for j in np.nonzero(freq)[0]:
            ans += weight * evals[m][j] * int(freq[j]) / per_Pauli
    return ans


def sampled_shifted_energies(probs, weights, evals, per_Pauli=100, is_noisy=False):
    """ps_p / ps_m of every control of every sample by shot sampling, in the reference's draw order (sim_plain.py:196-218):
    per sample, per control, ket_p (its Pauli terms in order, then its noise draw when is_noisy), then ket_m.
    probs: [B, n_H, 2, n_meas, dim] -> energies [B, n_H, 2]."""
    B, n_H = probs.shape[:2]
    en = np.empty((B, n_H, 2))
    for b in range(B):
        for i in range(n_H):
            for k in range(2):
                v = stochastic_measure_from_probs(probs[b, i, k], weights, evals, per_Pauli)
                if is_noisy:
                    v += np.random.normal(scale=np.abs(np.real(v)) / 5)
                en[b, i, k] = np.real(v)
    return en


def stochastic_measure_zstrings(zz, term_pair, weights, per_Pauli=100):
    """stochastic_measure (sim_plain.py:101-117) for Z-string Pauli terms (demo_maxcut.py:47-65) given zz[e] = <Z_a Z_b>:
    the eigenvalues of Z_a Z_b come sorted (-1 block first, eigenstates()), so the index np.random.choice(dim, per_Pauli, p)
    draws (:112: per_Pauli uniforms, searchsorted in the cumulative distribution) has eigenvalue -1 exactly when its uniform
    falls below P(-1) = (1 - <ZZ>) / 2.  Same stream consumption as the reference; the value differs from its j-by-j sum
    only by the rounding of that sum (1e-16).  term_pair[m]: pair id, or -1 for the identity term (always +1)."""
    ans = 0.0
    for m in range(len(weights)):
        u = np.random.random_sample(per_Pauli)
        k_minus = 0 if term_pair[m] < 0 else int(np.count_nonzero(u < 0.5 * (1.0 - zz[term_pair[m]])))
        ans += weights[m] * (per_Pauli - 2 * k_minus) / per_Pauli
    return ans
