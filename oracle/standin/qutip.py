"""Minimal stand-in for the `qutip` names the reference touches on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  qutip is not installed in
this image; the reference (`/root/reference/sim_plain.py`, `demo_maxcut.py`)
is imported UNMODIFIED behind this module so that its own NumPy/SciPy code can
be run as the parity oracle.  Use sites in the reference:
  sim_plain.py:121,129,131 (.full), :152 (Qobj(ndarray)), :197-199 (qeye, +, -, *, /),
  :205,215,281 (matrix_element), :294 (eigenenergies), :496-499 (.data, [i]);
  demo_maxcut.py:65 (eigenstates), :81-85 (Qobj(ndarray)); sim_plain.py:330,386,448 (mesolve, comparators only).
Nothing here tidies small elements (real qutip may); the stand-in run is authoritative.
"""
import numpy as np


class Qobj(object):
    __array_priority__ = 1000

    def __init__(self, arr):
        a = np.asarray(arr.full() if isinstance(arr, Qobj) else arr)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        self._a = np.array(a, dtype=np.complex128)

    # containers -----------------------------------------------------------
    def full(self):
        return self._a.copy()

    @property
    def shape(self):
        return self._a.shape

    @property
    def data(self):
        return self._a

    def __getitem__(self, i):
        r = self._a[i]
        if isinstance(r, np.ndarray) and r.size == 1:
            return complex(r.reshape(-1)[0])
        return r

    # algebra --------------------------------------------------------------
    def dag(self):
        return Qobj(self._a.conj().T)

    def norm(self):
        if 1 in self._a.shape:
            return float(np.linalg.norm(self._a))
        return float(np.linalg.norm(self._a, 'nuc'))

    def __mul__(self, o):
        if isinstance(o, Qobj):
            return Qobj(self._a @ o._a)
        return Qobj(self._a * o)

    def __rmul__(self, o):
        return Qobj(o * self._a)

    def __truediv__(self, o):
        return Qobj(self._a / o)

    def __add__(self, o):
        return Qobj(self._a + (o._a if isinstance(o, Qobj) else o))

    __radd__ = __add__

    def __sub__(self, o):
        return Qobj(self._a - (o._a if isinstance(o, Qobj) else o))

    def __neg__(self):
        return Qobj(-self._a)

    # measurements ---------------------------------------------------------
    def matrix_element(self, bra, ket):
        b = bra._a if isinstance(bra, Qobj) else np.asarray(bra)
        k = ket._a if isinstance(ket, Qobj) else np.asarray(ket)
        if b.shape[1] == 1:          # a ket was passed as the bra (reference does this)
            b = b.conj().T
        return complex((b @ self._a @ k).reshape(-1)[0])

    def eigenenergies(self):
        return np.linalg.eigvalsh(self._a)

    def eigenstates(self):
        w, v = np.linalg.eigh(self._a)
        return w, [Qobj(v[:, i]) for i in range(v.shape[1])]


def qeye(d):
    return Qobj(np.eye(int(d)))


class _Result(object):
    def __init__(self, states):
        self.states = states


def mesolve(H, psi0, tlist, *a, **k):
    """Stand-in for qutip.mesolve on a closed system with a ket (use sites: sim_plain.py:330,386,448): the Schroedinger
    equation d psi / dt = -i (H0 + sum_i u_i(t, args) H_i) psi for H = [H0, [H_i, u_i], ...], integrated with SciPy's DOP853 at
    tolerances far below anything the comparators resolve (qutip's own default is an Adams method at rtol 1e-6).  Returns an
    object with .states = [Qobj at every t in tlist]."""
    from scipy.integrate import solve_ivp
    H0 = H[0].full() if isinstance(H[0], Qobj) else np.asarray(H[0], dtype=np.complex128)
    terms = [(h[0].full(), h[1]) for h in H[1:]]
    y0 = psi0.full().reshape(-1)

    def rhs(t, y):
        Ht = H0.copy()
        for (Hi, ui) in terms:
            Ht = Ht + ui(t, None) * Hi
        return -1j * (Ht @ y)

    tlist = np.asarray(tlist, dtype=np.float64)
    sol = solve_ivp(rhs, (float(tlist[0]), float(tlist[-1])), y0, method="DOP853", t_eval=tlist, rtol=1e-12, atol=1e-14,
                    max_step=1e-2)
    if not sol.success:
        raise RuntimeError("stand-in mesolve: %s" % sol.message)
    return _Result([Qobj(sol.y[:, i]) for i in range(sol.y.shape[1])])
