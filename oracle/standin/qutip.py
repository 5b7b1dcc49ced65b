"""Minimal stand-in for the `qutip` names the reference touches on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  qutip is not installed in
this image; the reference (`/root/reference/sim_plain.py`, `demo_maxcut.py`)
is imported UNMODIFIED behind this module so that its own NumPy/SciPy code can
be run as the parity oracle.  Use sites in the reference:
  sim_plain.py:121,129,131 (.full), :152 (Qobj(ndarray)), :197-199 (qeye, +, -, *, /),
  :205,215,281 (matrix_element), :294 (eigenenergies), :496-499 (.data, [i]);
  demo_maxcut.py:65 (eigenstates), :81-85 (Qobj(ndarray)).
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


def mesolve(*a, **k):  # pragma: no cover - off the hot path (FD / fidelity comparators)
    raise NotImplementedError("qutip.mesolve is outside the hot path; not provided by the stand-in")
