"""ctypes face of oracle/_ref/libfu.so: the reference's OWN f_u / my_expit / bspline (diffqc.cc:75-135), compiled from
the reference source by oracle/ref_cc/Makefile.  TEST INFRASTRUCTURE ONLY.  The .so is built in the build container
(where /root/reference exists) and travels to the GPU box; fixtures made with it are committed under tests/golden/."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libfu.so")
_lib = None


def build():
    """Compile from /root/reference (no-op when the reference tree is absent)."""
    if os.path.isfile("/root/reference/diffqc.cc"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref_cc")])
    return available()


def available():
    return os.path.isfile(LIB)


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(LIB)
        lib.ref_f_u.restype = ctypes.c_double
        lib.ref_f_u.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.ref_my_expit.restype = ctypes.c_double
        lib.ref_my_expit.argtypes = [ctypes.c_double]
        lib.ref_bspline.restype = ctypes.c_double
        lib.ref_bspline.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double]
        lib.ref_set_channels.restype = None
        lib.ref_set_channels.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_int]
        _lib = lib
    return _lib


def f_u_table(channels, duration, func_type, vv, ts):
    """u[k, h] = f_u(h, ts[k], vv) evaluated by the reference's compiled code."""
    lib = load()
    counts = np.array([len(c) for c in channels], dtype=np.int32)
    flat = np.ascontiguousarray(np.array([ch for c in channels for ch in c], dtype=np.float64).reshape(-1, 4))
    lib.ref_set_channels(len(channels), counts.ctypes.data, flat.ctypes.data, float(duration), int(func_type))
    v = np.ascontiguousarray(np.asarray(vv, dtype=np.float64))
    out = np.empty((len(ts), len(channels)))
    for k, t in enumerate(ts):
        for h in range(len(channels)):
            out[k, h] = lib.ref_f_u(h, float(t), v.ctypes.data, v.shape[1], v.shape[2])
    return out
