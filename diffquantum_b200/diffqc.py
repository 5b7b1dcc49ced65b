"""`diffqc` — the reference's pybind11 module (diffqc.cc:210-228) re-hosted on the B200.

Same five functions and the same conversion rules as the pybind11 casters the reference relies on:
  * sequence arguments are any non-str Python sequence (lists, tuples, ndarrays), deep-copied
    element by element (pybind11 stl.h:129-142); a str/bytes or a ragged nest raises TypeError;
  * complex entries are anything complex() accepts (complex.h:44-56);
  * `per_step` and `func_type` are C++ ints: a Python float raises TypeError (cast.h:131-133);
  * results are fresh Python lists of Python complex / float (stl.h:155-167, complex.h:58-60);
  * the Hamiltonian, channels, duration and basis type set by set_H are module-global state
    (diffqc.cc:21-25); trotter before set_H raises RuntimeError instead of reading empty globals.
Where the reference has undefined behaviour (mismatched sizes, parameter index out of range,
zero steps) this module raises ValueError.
"""
import ctypes
import numbers

import numpy as np

from . import _lib

__version__ = "dev"                     # diffqc.cc:224-227 without VERSION_INFO

_state = {"dim": None, "n_H": 0, "device": 0}


def set_device(device):
    """Not part of the reference API: choose the GPU the module-global state lives on."""
    _state["device"] = int(device)
    _state["dim"] = None


def _seq(x, what):
    if isinstance(x, (str, bytes)) or not hasattr(x, "__len__") or not hasattr(x, "__getitem__"):
        raise TypeError("%s: expected a sequence, got %s" % (what, type(x).__name__))
    return x


def _int(x, what):
    if isinstance(x, bool) or not isinstance(x, (numbers.Integral, np.integer)):
        raise TypeError("%s: expected int, got %s" % (what, type(x).__name__))
    return int(x)


def _float(x, what):
    if isinstance(x, (str, bytes)) or not isinstance(x, (numbers.Real, np.floating, np.integer)):
        raise TypeError("%s: expected float, got %s" % (what, type(x).__name__))
    return float(x)


def _complex_array(x, what, ndim):
    _seq(x, what)
    try:
        a = np.array(x, dtype=np.complex128)
    except (TypeError, ValueError) as e:
        raise TypeError("%s: cannot convert to complex (%s)" % (what, e))
    if a.ndim != ndim:
        raise TypeError("%s: expected %d nested levels, got shape %s" % (what, ndim, a.shape))
    return np.ascontiguousarray(a)


def print_test():
    """diffqc.cc:27-29."""
    print("hello")


def complex_test(psi0):
    """Identity round trip through the complex list caster (diffqc.cc:31-34)."""
    return [complex(v) for v in _complex_array(psi0, "complex_test(psi0)", 1)]


def test_eigen(v):
    """Identity round trip through the nested float list caster (diffqc.cc:36-38)."""
    _seq(v, "test_eigen(v)")
    out = []
    for row in v:
        _seq(row, "test_eigen(v[i])")
        out.append([_float(e, "test_eigen(v[i][j])") for e in row])
    return out


def set_H(_H0, _Hs, channels, duration, func_type):
    """diffqc.cc:43-73.  channels[h][c] = [_, omega, w, idx] (diffqc.cc:108-111); func_type 0 is the
    Legendre basis, anything else the bump ("b-spline") basis (diffqc.cc:25,115-125)."""
    H0 = _complex_array(_H0, "set_H(_H0)", 2)
    dim = H0.shape[0]
    if H0.shape[1] != dim:
        raise ValueError("set_H: _H0 must be square, got %s" % (H0.shape,))
    _seq(_Hs, "set_H(_Hs)")
    if len(_Hs):
        Hs = _complex_array(_Hs, "set_H(_Hs)", 3)
        if Hs.shape[1:] != (dim, dim):
            raise ValueError("set_H: every _Hs[k] must be %dx%d, got %s" % (dim, dim, Hs.shape[1:]))
    else:
        Hs = np.zeros((0, dim, dim), dtype=np.complex128)
    _seq(channels, "set_H(channels)")
    if len(channels) < len(Hs):
        raise ValueError("set_H: %d control terms but only %d channel lists" % (len(Hs), len(channels)))
    counts, flat = [], []
    for h in range(len(Hs)):
        ch = _seq(channels[h], "set_H(channels[h])")
        counts.append(len(ch))
        for c in ch:
            c = [_float(e, "set_H(channels[h][c][k])") for e in _seq(c, "set_H(channels[h][c])")]
            if len(c) < 4:
                raise ValueError("set_H: a channel needs [_, omega, w, idx], got %d entries" % len(c))
            flat.append(c[:4])
    duration = _float(duration, "set_H(duration)")
    func_type = _int(func_type, "set_H(func_type)")
    counts = np.array(counts, dtype=np.int32)
    flat = np.ascontiguousarray(np.array(flat, dtype=np.float64).reshape(-1, 4))
    ctx = _lib.Context.get(_state["device"])
    _lib.check(_lib.load().dq_dense_set_H(ctx.handle, dim, _lib.ptr(H0), len(Hs), _lib.ptr(Hs), _lib.ptr(counts),
                                          _lib.ptr(flat), duration, func_type))
    _state["dim"] = dim
    _state["n_H"] = len(Hs)


def trotter(_psi0, T0, T, per_step, vv):
    """diffqc.cc:173-205: n_steps = (int)(per_step (|T - T0| + 1)), dt = (T - T0)/n_steps, per step
    psi <- exp(-i dt (H0 + sum_h f_u(h, t) H_h)) psi with t accumulated; returns list[complex]."""
    psi0 = _complex_array(_psi0, "trotter(_psi0)", 1)
    T0 = _float(T0, "trotter(T0)")
    T = _float(T, "trotter(T)")
    per_step = _int(per_step, "trotter(per_step)")
    _seq(vv, "trotter(vv)")
    try:
        v = np.ascontiguousarray(np.array(vv, dtype=np.float64))
    except (TypeError, ValueError) as e:
        raise TypeError("trotter(vv): cannot convert to float (%s)" % e)
    if v.ndim != 3 or v.shape[0] < 2:
        raise TypeError("trotter(vv): expected [2][n_param][n_basis], got shape %s" % (v.shape,))
    v = np.ascontiguousarray(v[:2])
    if _state["dim"] is None:
        raise RuntimeError("diffqc.trotter called before diffqc.set_H")
    if psi0.shape[0] != _state["dim"]:
        raise ValueError("trotter: psi0 has %d amplitudes, H is %dx%d" % (psi0.shape[0], _state["dim"], _state["dim"]))
    out = np.empty_like(psi0)
    ctx = _lib.Context.get(_state["device"])
    _lib.check(_lib.load().dq_dense_trotter(ctx.handle, _lib.ptr(psi0), T0, T, per_step, _lib.ptr(v), v.shape[1],
                                            v.shape[2], _lib.ptr(out), None))
    return [complex(a) for a in out]


def _pulse_table(T0, T, per_step, vv):
    """Testing aid: the f_u values the device run used, [n_steps][n_H] (not in the reference API)."""
    v = np.ascontiguousarray(np.array(vv, dtype=np.float64)[:2])
    n_steps = int(per_step * (abs(T - T0) + 1))
    u = np.empty((n_steps, _state["n_H"]))
    psi = np.zeros(_state["dim"], dtype=np.complex128)
    psi[0] = 1
    out = np.empty_like(psi)
    ctx = _lib.Context.get(_state["device"])
    _lib.check(_lib.load().dq_dense_trotter(ctx.handle, _lib.ptr(psi), float(T0), float(T), int(per_step), _lib.ptr(v),
                                            v.shape[1], v.shape[2], _lib.ptr(out), _lib.ptr(u)))
    return u
