"""ctypes face of oracle/c/liboracle_c.so (plain-C, OpenMP restatement of the split step and the
shifted-ket estimator).  TEST INFRASTRUCTURE ONLY — used by tests/ as a second checker and by
bench.py as the multi-core CPU baseline; the product never imports it.

Follows the same reference lines as restate.py: split step diffqc.cc:155-164, shift gates
sim_plain.py:197-199, energies sim_plain.py:205,215, estimator loop sim_plain.py:186-230."""
import ctypes
import os

import numpy as np

from . import restate as R

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "c", "liboracle_c.so")
_lib = None


_STAMP = os.path.join(_HERE, "c", "built_for.txt")


def _host_id():
    """CPU model + ISA flags of this machine: the library is built with -march=native, so it only runs where it was built."""
    try:
        txt = open("/proc/cpuinfo").read()
        import hashlib
        model = [l for l in txt.splitlines() if l.startswith("model name")][:1]
        flags = [l for l in txt.splitlines() if l.startswith("flags")][:1]
        return hashlib.sha1(("".join(model) + "".join(flags)).encode()).hexdigest()
    except OSError:
        return "unknown"


def ensure_built():
    """(Re)build liboracle_c.so when it is missing or was built on a different CPU (the snapshot that travels to the GPU box
    carries the build container's .so).  Needs gcc and make; returns available()."""
    import subprocess
    have = os.path.isfile(LIB_PATH) and os.path.isfile(_STAMP) and open(_STAMP).read().strip() == _host_id()
    if not have:
        try:
            subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(_HERE, "c")])
            open(_STAMP, "w").write(_host_id() + "\n")
        except (OSError, subprocess.CalledProcessError):
            return False
    return os.path.isfile(LIB_PATH)


_avail = None


def available():
    global _avail
    if _avail is None:
        _avail = ensure_built()
    return _avail


def load():
    global _lib
    if _lib is None:
        available()
        lib = ctypes.CDLL(LIB_PATH)
        vp = ctypes.c_void_p
        lib.oc_num_threads.restype = ctypes.c_int
        lib.oc_set_num_threads.restype = None
        lib.oc_set_num_threads.argtypes = [ctypes.c_int]
        lib.oc_evolve_split.restype = ctypes.c_int
        lib.oc_evolve_split.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int,
                                        ctypes.c_double, vp]
        lib.oc_shift_gate.restype = None
        lib.oc_shift_gate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_double, vp, vp]
        lib.oc_energy_diag.restype = ctypes.c_double
        lib.oc_energy_diag.argtypes = [ctypes.c_int, vp, vp]
        _lib = lib
    return _lib


def num_threads():
    return load().oc_num_threads()


def use_host_cores():
    """All cores this process may run on, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    load().oc_set_num_threads(n)
    return num_threads()


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CProblem(object):
    """Term tables of a restate.maxcut_structured() problem in the layout the C code takes."""

    def __init__(self, prob):
        self.prob = prob
        self.n = prob['n']
        terms = prob['terms']
        self.kind = np.array([0 if t[0] == 'zz' else 1 for t in terms], dtype=np.int32)
        self.qa = np.array([t[1] for t in terms], dtype=np.int32)
        self.qb = np.array([t[2] if t[0] == 'zz' else 0 for t in terms], dtype=np.int32)
        self.h0 = np.ascontiguousarray(prob['h0_diag'], dtype=np.float64)
        self.m = np.ascontiguousarray(prob['m_diag'], dtype=np.float64)

    def evolve(self, u, dt, psi):
        """In-place split evolution of psi (complex128 [2^n]) through the rows of u."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        rc = load().oc_evolve_split(self.n, len(self.kind), _p(self.kind), _p(self.qa), _p(self.qb), _p(self.h0),
                                    _p(u), u.shape[0], float(dt), _p(psi))
        if rc != 0:
            raise MemoryError("oc_evolve_split")
        return psi

    def run(self, coeff, psi, T0, T1, per_step, basis='BSpline'):
        n_steps, dt, ts = R.step_grid(T0, T1, per_step)
        u = R.coef_table_plain(coeff, self.prob['omegas'], self.prob['T'], ts, basis)
        return self.evolve(u, dt, psi), n_steps

    def shift(self, i, phi, sign, r=0.5):
        out = np.empty_like(phi)
        load().oc_shift_gate(self.n, int(self.kind[i]), int(self.qa[i]), int(self.qb[i]), float(sign), float(r),
                             _p(phi), _p(out))
        return out

    def energy(self, psi):
        return load().oc_energy_diag(self.n, _p(self.m), _p(psi))


def grad_mc(cprob, coeff, s, per_step, r=0.5, basis='BSpline', coeff_sign=1.0, terms=None,
            return_energies=False):
    """One stochastic parameter-shift sample (sim_plain.py:156-231) on the C kernels.  `terms`
    restricts the shifted trajectories to a subset of controls (bench.py's bounded CPU sample);
    returns (grad, energies, trajectory_steps_executed)."""
    prob = cprob.prob
    T = prob['T']
    n_H = len(prob['terms'])
    phi, steps = cprob.run(coeff, prob['psi0'].copy(), 0, s, per_step, basis)
    grad = np.zeros((n_H, coeff.shape[1]))
    energies = np.zeros((n_H, 2))
    for i in (range(n_H) if terms is None else terms):
        ket, k = cprob.run(coeff, cprob.shift(i, phi, +1, r), s, T, per_step, basis)
        ps_p = cprob.energy(ket)
        ket, k2 = cprob.run(coeff, cprob.shift(i, phi, -1, r), s, T, per_step, basis)
        ps_m = cprob.energy(ket)
        steps += k + k2
        energies[i] = (ps_p, ps_m)
        ps = coeff_sign * ((1 + r ** 2) / 2 / r * (ps_m - ps_p))
        grad[i, :] = ps * R.dudc_plain(i, s, coeff, prob['omegas'], T, basis)
    if return_energies:
        return grad, energies, steps
    return grad
