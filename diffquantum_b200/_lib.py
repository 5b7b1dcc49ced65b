"""ctypes binding of libdiffqc_b200.so (the C ABI declared in include/diffqc_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is visible when a
context is requested, the product path raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DIFFQC_B200_LIB") or os.path.join(_HERE, "libdiffqc_b200.so")   # override: kernel experiments

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)

# name -> (restype, argtypes); must list every symbol include/diffqc_b200.h declares
_VP = ctypes.c_void_p
SIGNATURES = {
    "dq_version": (ctypes.c_char_p, []),
    "dq_last_error": (ctypes.c_char_p, []),
    "dq_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "dq_context_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_VP)]),
    "dq_context_destroy": (ctypes.c_int, [_VP]),
    "dq_context_synchronize": (ctypes.c_int, [_VP]),
    "dq_context_stream": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_uint64)]),
    "dq_context_launch_count": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_uint64)]),
    "dq_dense_set_H": (ctypes.c_int, [_VP, ctypes.c_int, _VP, ctypes.c_int, _VP, _VP, _VP,
                                      ctypes.c_double, ctypes.c_int]),
    "dq_dense_trotter": (ctypes.c_int, [_VP, _VP, ctypes.c_double, ctypes.c_double, ctypes.c_int, _VP,
                                        ctypes.c_int, ctypes.c_int, _VP, _VP]),
    "dq_pulse_f_u_table": (ctypes.c_int, [ctypes.c_int, _VP, _VP, ctypes.c_double, ctypes.c_int, _VP, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, _VP, _VP]),
    "dq_dense_evolve": (ctypes.c_int, [_VP, ctypes.c_int, _VP, ctypes.c_int, _VP, _VP, ctypes.c_int,
                                       ctypes.c_double, ctypes.c_int, ctypes.c_int, _VP, _VP]),
    "dq_dense_grad": (ctypes.c_int, [_VP, ctypes.c_int, _VP, ctypes.c_int, _VP, _VP, _VP, ctypes.c_double,
                                     ctypes.c_int, _VP, _VP, _VP, _VP, _VP, _VP, ctypes.c_int, _VP]),
    "dq_dense_grad_times": (ctypes.c_int, [_VP, ctypes.c_int, _VP, ctypes.c_int, _VP, _VP, _VP, ctypes.c_double, ctypes.c_int, _VP,
                                           ctypes.c_double, ctypes.c_int, _VP, _VP, ctypes.c_int, ctypes.c_int, _VP, _VP]),
    "dq_dense_grad_probs": (ctypes.c_int, [_VP, ctypes.c_int, _VP, ctypes.c_int, _VP, _VP, ctypes.c_double, ctypes.c_int, _VP, _VP, _VP,
                                           _VP, _VP, _VP, ctypes.c_int, ctypes.c_int, _VP, _VP]),
    "dq_dense_outcome_probs": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, ctypes.c_int, _VP, _VP]),
    "dq_dense_evolve_many": (ctypes.c_int, [_VP, ctypes.c_int, _VP, ctypes.c_int, _VP, _VP, ctypes.c_int, _VP, _VP, _VP, _VP, ctypes.c_int,
                                            _VP, _VP]),
    "dq_dense_train": (ctypes.c_int, [_VP, ctypes.c_int, _VP, ctypes.c_int, _VP, _VP, _VP, _VP, ctypes.c_double, ctypes.c_int,
                                      ctypes.c_int, _VP, ctypes.c_int, ctypes.c_int, _VP, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, _VP, _VP]),
    "dq_dense_last_stat": (ctypes.c_int, [_VP, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]),
    "dq_dense_set_option": (ctypes.c_int, [_VP, ctypes.c_char_p, ctypes.c_int64]),
    "dq_ising_create": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, _VP, ctypes.c_double, _VP,
                                       ctypes.POINTER(_VP)]),
    "dq_ising_destroy": (ctypes.c_int, [_VP]),
    "dq_ising_set_option": (ctypes.c_int, [_VP, ctypes.c_char_p, ctypes.c_int64]),
    "dq_ising_get_info": (ctypes.c_int, [_VP, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64)]),
    "dq_ising_evolve": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP, ctypes.c_int, _VP]),
    "dq_ising_grad": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _VP, _VP, _VP, ctypes.c_int, _VP, _VP,
                                     ctypes.c_double, _VP, _VP]),
    "dq_ising_grad_stage": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _VP, _VP, _VP, ctypes.c_int, _VP, _VP,
                                           ctypes.c_double, _VP]),
    "dq_ising_grad_run_staged": (ctypes.c_int, [_VP]),
    "dq_ising_grad_fetch": (ctypes.c_int, [_VP, _VP]),
    "dq_ising_last_stat": (ctypes.c_int, [_VP, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]),
    "dq_ising_grad_pairs": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _VP, _VP, _VP, ctypes.c_int, _VP, _VP, ctypes.c_double, _VP, _VP]),
    "dq_ising_pair_expect": (ctypes.c_int, [_VP, ctypes.c_int, _VP, ctypes.c_int, _VP]),
    "dq_ising_train": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _VP, _VP, _VP, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                      ctypes.c_int, _VP, ctypes.c_int, ctypes.c_int, _VP, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _VP, _VP, _VP,
                                      ctypes.POINTER(ctypes.c_double)]),
    "dq_slice_fill_uniform": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int]),
    "dq_slice_phase": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, _VP, _VP]),
    "dq_slice_rx": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_double]),
    "dq_slice_rx_many": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, _VP, _VP]),
    "dq_slice_phase_rx_many": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, _VP, _VP,
                                              ctypes.c_int, _VP, _VP]),
    "dq_slice_rx_many_scatter": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, _VP, _VP, ctypes.c_int, ctypes.c_int, _VP]),
    "dq_slice_phase_rx_many_scatter": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, _VP, _VP,
                                                      ctypes.c_int, _VP, _VP, ctypes.c_int, ctypes.c_int, _VP]),
    "dq_slice_step": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, _VP, _VP,
                                     ctypes.c_int, _VP, _VP, ctypes.c_int, _VP, _VP]),
    "dq_slice_step_scatter": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, _VP, _VP,
                                             ctypes.c_int, _VP, _VP, ctypes.c_int, _VP, _VP, ctypes.c_int, ctypes.c_int, _VP]),
    "dq_slice_evolve_steps": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, _VP,
                                             ctypes.c_int, _VP, ctypes.c_int, _VP, ctypes.c_int64, _VP, ctypes.c_int64]),
    "dq_slice_plan": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP, ctypes.c_int, _VP, ctypes.c_int, ctypes.c_int,
                                     _VP, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
    "dq_slice_plan_step": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP, ctypes.c_int, _VP, ctypes.c_int, _VP,
                                          ctypes.c_int, ctypes.c_int, _VP, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
    "dq_ipc_export": (ctypes.c_int, [_VP, _VP, _VP, ctypes.POINTER(ctypes.c_uint64)]),
    "dq_ipc_open": (ctypes.c_int, [_VP, _VP, ctypes.c_uint64, ctypes.POINTER(_VP)]),
    "dq_ipc_close": (ctypes.c_int, [_VP, _VP, ctypes.c_uint64]),
    "dq_slice_energy": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, _VP, _VP,
                                       ctypes.c_double, ctypes.POINTER(ctypes.c_double)]),
    "dq_microbench": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_double)]),
}

_lib = None


class DiffqcError(RuntimeError):
    pass


def load():
    """Load the shared library (no device needed) and attach the signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise DiffqcError("%s is missing: build it with `python -c 'import __graft_entry__ as g; "
                          "g.build()'` or `make -C diffquantum_b200/csrc` (there is no CPU fallback)"
                          % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_STATUS_EXC = {-1: ValueError, -3: RuntimeError, -4: NotImplementedError, -5: MemoryError}


def check(status):
    """Map a dq_status to the Python exception the pybind11 module would have raised
    (ValueError/TypeError for bad arguments; RuntimeError for everything CUDA)."""
    if status == 0:
        return
    msg = load().dq_last_error().decode("utf-8", "replace")
    raise _STATUS_EXC.get(status, DiffqcError)("diffqc_b200: %s (status %d)" % (msg, status))


def ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


class Context(object):
    """One CUDA context/stream on one device; shared by every problem created from it."""

    _cache = {}

    def __init__(self, device=0):
        lib = load()
        self.device = int(device)
        h = ctypes.c_void_p()
        check(lib.dq_context_create(self.device, ctypes.byref(h)))
        self.handle = h

    @classmethod
    def get(cls, device=0):
        device = int(device)
        if device not in cls._cache:
            cls._cache[device] = cls(device)
        return cls._cache[device]

    def synchronize(self):
        check(load().dq_context_synchronize(self.handle))

    @property
    def stream(self):
        v = ctypes.c_uint64()
        check(load().dq_context_stream(self.handle, ctypes.byref(v)))
        return v.value

    @property
    def launch_count(self):
        v = ctypes.c_uint64()
        check(load().dq_context_launch_count(self.handle, ctypes.byref(v)))
        return v.value

    def microbench(self, kind, nbytes=1 << 30, iters=10):
        v = ctypes.c_double()
        check(load().dq_microbench(self.handle, int(kind), int(nbytes), int(iters), ctypes.byref(v)))
        return v.value
