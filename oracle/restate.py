"""NumPy/SciPy restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

Every function cites the reference lines it restates (paths relative to /root/reference).
Conventions taken from the reference and kept here:
  * qubit j of an n-qubit register is bit (n-1-j) of the basis index, because operators are
    built with np.kron in qubit order (demo_maxcut.py:49-57, sim_plain.py:477-482);
  * the time grid samples the LEFT end of each step and accumulates `t += dt`
    (sim_plain.py:133-150, diffqc.cc:182-199);
  * n_steps truncates toward zero; the Python twin has no abs(), the C++ twin has
    (sim_plain.py:123 vs diffqc.cc:182).
"""
import math

import numpy as np
import scipy.linalg
from scipy.sparse.linalg import expm_multiply
import scipy.sparse as sps
from scipy.special import eval_legendre

# --------------------------------------------------------------------------------------
# time grid
# --------------------------------------------------------------------------------------


def step_grid(T0, T, per_step, use_abs=False):
    """(n_steps, dt, ts) with ts[k] the accumulated left-end time of step k.

    sim_plain.py:123,133-134,150 (use_abs=False) / diffqc.cc:182-184,199 (use_abs=True).
    """
    span = abs(T - T0) if use_abs else (T - T0)
    n_steps = int(per_step * (span + 1))
    if n_steps <= 0:
        return 0, 0.0, np.zeros(0)
    dt = (T - T0) / n_steps
    ts = np.empty(n_steps)
    t = T0
    for k in range(n_steps):
        ts[k] = t
        t += dt
    return n_steps, dt, ts


# --------------------------------------------------------------------------------------
# pulse models
# --------------------------------------------------------------------------------------


def bspline_value(b, n_basis, x):
    """Quadratic bump b of the reference's 'BSpline' basis at x = t/T.

    sim_plain.py:52-70 (open support: zero when x >= r or x <= l); same as diffqc.cc:82-93.
    """
    tau = 1. / (n_basis - 2)
    tau_b = tau * (b - 1.5)
    l = tau_b - 1.5 * tau
    r = tau_b + 1.5 * tau
    if x >= r or x <= l:
        return 0.0
    return (x - l) * (x - r) / (-(1.5 * tau) ** 2)


def sigmoid_py(x):
    """sim_plain.py:49-50 (math.exp, no cutoff)."""
    return 1 / (1 + math.exp(-x))


def u_plain(i, t, coeff, omegas, T, basis='BSpline'):
    """u_i(t) of the Python twin: sim_plain.py:73-99 (sequential accumulation from j=0)."""
    n_basis = coeff.shape[1]
    u = 0
    n = int(n_basis / 2) if basis == 'Fourier' else n_basis                # :84
    for j in range(n):
        if basis == 'BSpline':
            u += coeff[i][j] * bspline_value(j, n_basis, t / T)
        elif basis == 'Legendre':
            u += coeff[i][j] * eval_legendre(j, 2 * t / T - 1)
        elif basis == 'poly':                                              # :86-87 (raw t, not t / T)
            u += coeff[i][j] * (t - 0.5) ** j
        elif basis == 'Fourier':                                           # :90-92
            u += coeff[i][j] * np.cos(2 * np.pi * j * t) + coeff[i][j + n] * np.sin(2 * np.pi * j * t)
        else:
            raise ValueError(basis)
    return (sigmoid_py(u) * 2 - 1) * omegas[i]


def dudc_plain(i, s, coeff, omegas, T, basis='BSpline'):
    """Row i of dDdv at time s: d u_i(s) / d c_ij.

    The reference gets this from torch autograd of Ds = (sigmoid(A)*2-1)*omega_i
    (sim_plain.py:169-184); the closed form is omega_i * 2*sigma(A)(1-sigma(A)) * phi_j(s/T).
    """
    n_basis = coeff.shape[1]
    if basis == 'BSpline':
        phis = [bspline_value(j, n_basis, s / T) for j in range(n_basis)]
    else:
        phis = [float(eval_legendre(j, 2 * s / T - 1)) for j in range(n_basis)]
    A = sum(coeff[i][j] * phis[j] for j in range(n_basis))
    sg = sigmoid_py(A)
    return np.array([omegas[i] * 2.0 * sg * (1.0 - sg) * p for p in phis])


def expit_cc(x):
    """diffqc.cc:75-80 (cut off at +-32)."""
    if x > 32.:
        return 1.
    if x < -32.:
        return 0.
    return 1 / (1 + math.exp(-x))


def _round_half_away(x):
    """C `round()` as used at diffqc.cc:111."""
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def f_u_cc(h, t, vv, channels, duration, func_type):
    """IQ-modulated sigmoid-bounded envelope of the native twin: diffqc.cc:95-135.

    vv[0] / vv[1] hold the A / B coefficients, shape [2][n_param][n_basis];
    channels[h][c] = [_, omega, w, idx].
    """
    ans = 0.0
    n_basis = len(vv[0][0])
    for chan in channels[h]:
        omega = chan[1]
        w = chan[2]
        idx = _round_half_away(chan[3])
        A = 0.0
        B = 0.0
        for j in range(n_basis):
            if func_type == 0:
                fv = float(eval_legendre(j, 2 * t / duration - 1))
            else:
                fv = bspline_value(j, n_basis, t / duration)
            A += vv[0][idx][j] * fv
            B += vv[1][idx][j] * fv
        N = math.sqrt(A * A + B * B)
        if abs(N - 0.0) < 0.000001:
            ans += 0.0
        else:
            ans += omega * (2 * expit_cc(N) - 1) / N * (math.cos(w * t) * A + math.sin(w * t) * B)
    return ans


# --------------------------------------------------------------------------------------
# dense evolution (small n): live `exact` semantics and the disabled `split` product
# --------------------------------------------------------------------------------------


def coef_table_plain(coeff, omegas, T, ts, basis='BSpline'):
    """u[k, i] = u_i(ts[k]) for the Python twin."""
    n_H = coeff.shape[0]
    out = np.empty((len(ts), n_H))
    for k, t in enumerate(ts):
        for i in range(n_H):
            out[k, i] = u_plain(i, t, coeff, omegas, T, basis)
    return out


def evolve_exact_dense(H0, Hs, u, dt, psi0):
    """psi <- expm(-i dt (H0 + sum_i u[k,i] H_i)) psi for each step k.

    Live code of both twins: sim_plain.py:135-150, diffqc.cc:190-200.  The generator is summed
    in list order starting from H0, as the reference does.
    """
    psi = np.array(psi0, dtype=np.complex128).reshape(-1, 1)
    for k in range(u.shape[0]):
        dH = -1.j * dt * H0
        for i, H in enumerate(Hs):
            dH = dH + (-1.j * dt * u[k, i]) * H
        psi = scipy.linalg.expm(dH) @ psi
    return psi.reshape(-1)


def evolve_split_dense(H0, Hs, u, dt, psi0):
    """Per-term Lie-Trotter product, H0 first then H_h in list order.

    The reference's disabled variant: diffqc.cc:155-164 (and sim_plain.py:139,142).
    """
    psi = np.array(psi0, dtype=np.complex128).reshape(-1, 1)
    for k in range(u.shape[0]):
        psi = scipy.linalg.expm(-1.j * dt * H0) @ psi
        for i, H in enumerate(Hs):
            psi = scipy.linalg.expm((-1.j * dt * u[k, i]) * H) @ psi
    return psi.reshape(-1)


def trotter_plain(H0, Hs, coeff, omegas, T_total, psi0, T0, T, per_step, basis='BSpline',
                  mode='exact'):
    """SimulatorPlain.trotter (sim_plain.py:119-153) on explicit arrays."""
    n_steps, dt, ts = step_grid(T0, T, per_step, use_abs=False)
    u = coef_table_plain(coeff, omegas, T_total, ts, basis)
    f = evolve_exact_dense if mode == 'exact' else evolve_split_dense
    return f(H0, Hs, u, dt, psi0)


def trotter_cc(H0, Hs, channels, duration, func_type, psi0, T0, T, per_step, vv, mode='exact'):
    """diffqc.set_H + diffqc.trotter (diffqc.cc:43-73, 173-205)."""
    n_steps, dt, ts = step_grid(T0, T, per_step, use_abs=True)
    u = np.empty((n_steps, len(Hs)))
    for k, t in enumerate(ts):
        for h in range(len(Hs)):
            u[k, h] = f_u_cc(h, t, vv, channels, duration, func_type)
    f = evolve_exact_dense if mode == 'exact' else evolve_split_dense
    return f(np.asarray(H0, dtype=np.complex128), [np.asarray(H, dtype=np.complex128) for H in Hs],
             u, dt, psi0)


# --------------------------------------------------------------------------------------
# MaxCut problem construction
# --------------------------------------------------------------------------------------

_I2 = np.array([[1., 0.], [0., 1.]])
_X2 = np.array([[0., 1.], [1., 0.]])
_Z2 = np.array([[1., 0.], [0., -1.]])


def multi_kron(*args):
    """sim_plain.py:477-482."""
    ret = np.array([[1.0]])
    for q in args:
        ret = np.kron(ret, q)
    return ret


def z_diag(n, q):
    """Diagonal of Z on qubit q: +1 where bit (n-1-q) of the index is 0."""
    idx = np.arange(1 << n)
    return 1.0 - 2.0 * ((idx >> (n - 1 - q)) & 1)


def maxcut_structured(n, edges, omega0=np.pi, omega1=np.pi, n_layers=1):
    """The demo's problem in structured form (demo_maxcut.py:19-85).

    Controls = one ZZ per edge (graph order) then one X per qubit; H0 = sum_e (I - Z_a Z_b)
    because `H0 = OO` aliases `H_cost` and is mutated in place before `H_cost` is rebound
    (demo_maxcut.py:34-38,60-61: SURVEY F4); M = H_cost = -1/2 sum_e (I - Z_a Z_b);
    T = pi (1/omega0 + 1/omega1) n_layers (demo_maxcut.py:44); psi0 = uniform superposition.
    """
    edges = [tuple(int(v) for v in e) for e in edges]
    zz = np.zeros(1 << n)
    for a, b in edges:
        zz += z_diag(n, a) * z_diag(n, b)
    h0_diag = len(edges) - zz
    m_diag = -0.5 * (len(edges) - zz)
    terms = [('zz', a, b) for a, b in edges] + [('x', q) for q in range(n)]
    omegas = [omega0] * len(edges) + [omega1] * n
    T = np.pi * (1. / omega0 + 1. / omega1) * n_layers
    psi0 = np.full(1 << n, 1.0 / np.sqrt(2.0 ** n), dtype=np.complex128)
    return dict(n=n, edges=edges, terms=terms, omegas=np.array(omegas, dtype=float), T=float(T),
                h0_diag=h0_diag, m_diag=m_diag, psi0=psi0)


def maxcut_dense(prob):
    """Dense H0, Hs, M of a structured problem, built as the demo does (np.kron chains)."""
    n = prob['n']
    Hs = []
    for term in prob['terms']:
        if term[0] == 'zz':
            ops = [_Z2 if j in term[1:] else _I2 for j in range(n)]
        else:
            ops = [_X2 if j == term[1] else _I2 for j in range(n)]
        Hs.append(multi_kron(*ops).astype(np.complex128))
    H0 = np.diag(prob['h0_diag']).astype(np.complex128)
    M = np.diag(prob['m_diag']).astype(np.complex128)
    return H0, Hs, M


# --------------------------------------------------------------------------------------
# structured evolution (any n that fits): split and exact step semantics
# --------------------------------------------------------------------------------------


def term_diag(prob, term):
    n = prob['n']
    return z_diag(n, term[1]) * z_diag(n, term[2])


def evolve_split_structured(prob, u, dt, psi0):
    """Product-formula step on a Pauli-term problem, term order of diffqc.cc:155-164:
    exp(-i dt H0), then exp(-i dt u_h H_h) for h in list order.  ZZ terms are diagonal phases,
    X terms are 2x2 rotations [[c, -is], [-is, c]] on axis q of the (2,)*n view (qubit 0 = MSB).
    """
    n = prob['n']
    psi = np.array(psi0, dtype=np.complex128).reshape(-1)
    zz_diags = {}
    for k in range(u.shape[0]):
        psi = psi * np.exp(-1.j * dt * prob['h0_diag'])
        for i, term in enumerate(prob['terms']):
            th = dt * u[k, i]
            if term[0] == 'zz':
                if i not in zz_diags:
                    zz_diags[i] = term_diag(prob, term)
                psi = psi * np.exp(-1.j * th * zz_diags[i])
            else:
                q = term[1]
                v = psi.reshape(1 << q, 2, -1)
                c, s = math.cos(th), math.sin(th)
                a = c * v[:, 0, :] - 1.j * s * v[:, 1, :]
                b = c * v[:, 1, :] - 1.j * s * v[:, 0, :]
                psi = np.stack([a, b], axis=1).reshape(-1)
    return psi


def _sparse_terms(prob):
    n = prob['n']
    D = 1 << n
    mats = []
    idx = np.arange(D)
    for term in prob['terms']:
        if term[0] == 'zz':
            mats.append(sps.diags(term_diag(prob, term)).tocsr())
        else:
            q = term[1]
            mats.append(sps.csr_matrix((np.ones(D), (idx, idx ^ (1 << (n - 1 - q)))), shape=(D, D)))
    return sps.diags(prob['h0_diag']).tocsr(), mats


def evolve_exact_structured(prob, u, dt, psi0):
    """Exact step without the dense matrix: psi <- expm_multiply(dH, psi), the variant the
    reference leaves commented at sim_plain.py:147 (agrees with :145-146 to ~1e-14)."""
    H0, mats = _sparse_terms(prob)
    psi = np.array(psi0, dtype=np.complex128).reshape(-1)
    for k in range(u.shape[0]):
        dH = (-1.j * dt) * H0
        for i, m in enumerate(mats):
            dH = dH + (-1.j * dt * u[k, i]) * m
        psi = expm_multiply(dH.tocsc(), psi)
    return psi


def energy_diag(m_diag, psi):
    """<psi|M|psi> for diagonal M (sim_plain.py:205,215,281 with M = H_cost)."""
    return float(np.sum(m_diag * (psi.real ** 2 + psi.imag ** 2)))


def apply_shift_gate(prob, term, phi, sign, r=0.5):
    """(I + sign * i r H_i) phi / sqrt(1 + r^2): sim_plain.py:197-199."""
    n = prob['n']
    if term[0] == 'zz':
        hp = term_diag(prob, term) * phi
    else:
        q = term[1]
        hp = phi.reshape(1 << q, 2, -1)[:, ::-1, :].reshape(-1)
    return (phi + sign * r * 1.j * hp) / np.sqrt(1. + r ** 2)


def grad_mc_structured(prob, coeff, s, per_step, mode='split', basis='BSpline', coeff_sign=1.0,
                       r=0.5, return_energies=False):
    """One stochastic parameter-shift sample at time s: sim_plain.py:156-231 with the sampled
    time passed in (the reference draws s = np.random.uniform() * T at :167)."""
    T = prob['T']
    omegas = prob['omegas']
    n_H = len(prob['terms'])
    n_basis = coeff.shape[1]
    evolve = evolve_split_structured if mode == 'split' else evolve_exact_structured

    def run(psi, T0, T1):
        n_steps, dt, ts = step_grid(T0, T1, per_step)
        u = coef_table_plain(coeff, omegas, T, ts, basis)
        return evolve(prob, u, dt, psi)

    phi = run(prob['psi0'], 0, s)
    grad = np.zeros((n_H, n_basis))
    energies = np.zeros((n_H, 2))
    for i, term in enumerate(prob['terms']):
        ket_p = run(apply_shift_gate(prob, term, phi, +1, r), s, T)
        ps_p = energy_diag(prob['m_diag'], ket_p)
        ket_m = run(apply_shift_gate(prob, term, phi, -1, r), s, T)
        ps_m = energy_diag(prob['m_diag'], ket_m)
        energies[i] = (ps_p, ps_m)
        ps = coeff_sign * ((1 + r ** 2) / 2 / r * (ps_m - ps_p))
        grad[i, :] = ps * dudc_plain(i, s, coeff, omegas, T, basis)
    if return_energies:
        return grad, energies
    return grad


def stochastic_measure(psi, weights, evals, estates, per_Pauli=100):
    """sim_plain.py:101-117: per Pauli term, distr[j] = |<psi|e_j>|^2, per_Pauli draws with np.random.choice from the global
    stream, ans += weight * evals[j] * freq_j / per_Pauli over j in order."""
    ans = 0
    for i in range(len(weights)):
        distr = [abs(np.vdot(psi, estates[i][j])) ** 2 for j in range(len(evals[i]))]
        res = np.random.choice(len(evals[i]), per_Pauli, p=distr)
        for j in range(len(evals[i])):
            freq = np.count_nonzero(res == j)
            ans += weights[i] * evals[i][j] * freq / per_Pauli
    return ans


def grad_mc_dense(H0, Hs, M, psi0, coeff, omegas, T, s, per_step, mode='exact', basis='BSpline',
                  coeff_sign=1.0, r=0.5, return_energies=False, is_noisy=False, sampling=None):
    """Dense twin of grad_mc_structured: sim_plain.py:156-231 on explicit matrices.  is_noisy adds the reference's
    measurement noise (sim_plain.py:207-208,217-218): one np.random.normal(scale=|ps|/5) per shifted energy, drawn from
    the global stream in the order ps_p, ps_m per control.  sampling = (weights, evals, estates): shot sampling
    (sampling_measure=True, :202-203,212-213)."""
    n_H = len(Hs)
    d = len(psi0)

    def run(psi, T0, T1):
        return trotter_plain(H0, Hs, coeff, omegas, T, psi, T0, T1, per_step, basis, mode)

    phi = run(psi0, 0, s)
    grad = np.zeros((n_H, coeff.shape[1]))
    energies = np.zeros((n_H, 2))
    for i in range(n_H):
        gate_p = (np.eye(d) + r * 1.j * Hs[i]) / np.sqrt(1. + r ** 2)
        gate_m = (np.eye(d) - r * 1.j * Hs[i]) / np.sqrt(1. + r ** 2)
        ket_p = run(gate_p @ phi, s, T)
        ps_p = (ket_p.conj() @ M @ ket_p) if sampling is None else stochastic_measure(ket_p, *sampling) + 0j
        if is_noisy:
            ps_p += np.random.normal(scale=np.abs(ps_p.real) / 5)
        ket_m = run(gate_m @ phi, s, T)
        ps_m = (ket_m.conj() @ M @ ket_m) if sampling is None else stochastic_measure(ket_m, *sampling) + 0j
        if is_noisy:
            ps_m += np.random.normal(scale=np.abs(ps_m.real) / 5)
        energies[i] = (ps_p.real, ps_m.real)
        ps = coeff_sign * ((1 + r ** 2) / 2 / r * (ps_m - ps_p)).real
        grad[i, :] = ps * dudc_plain(i, s, coeff, omegas, T, basis)
    if return_energies:
        return grad, energies
    return grad


def random_regular_edges(n, seed, degree=3):
    """Synthetic workload graph of SURVEY 8(d): networkx.random_regular_graph, edges sorted."""
    import networkx as nx
    g = nx.random_regular_graph(degree, n, seed=seed)
    return sorted(tuple(sorted(e)) for e in g.edges())
