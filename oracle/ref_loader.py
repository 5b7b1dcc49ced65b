"""Import the UNMODIFIED reference (`/root/reference/sim_plain.py`) behind the stand-ins.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build container);
`available()` is False on the GPU box, where the committed fixtures under tests/golden/ are
used instead.
"""
import importlib
import os
import sys

REFERENCE_DIR = os.environ.get("DIFFQUANTUM_REFERENCE", "/root/reference")
_STANDIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standin")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "sim_plain.py"))


def load_sim_plain():
    """Return the reference's `sim_plain` module (class SimulatorPlain, sim_plain.py:14)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    saved = list(sys.path)
    try:
        # stand-ins first so `import qutip`, `matplotlib.pyplot`, `logger` resolve to them;
        # the reference directory after, so `sim_plain` itself is the real file.
        sys.path[:0] = [_STANDIN, REFERENCE_DIR]
        for name in ("qutip", "logger", "matplotlib", "matplotlib.pyplot"):
            mod = sys.modules.get(name)
            if mod is not None and not getattr(mod, "__file__", "").startswith(_STANDIN):
                del sys.modules[name]
        mod = importlib.import_module("sim_plain")
        assert os.path.dirname(os.path.abspath(mod.__file__)) == os.path.abspath(REFERENCE_DIR)
        return mod
    finally:
        sys.path[:] = saved


def load_qutip_standin():
    load_sim_plain() if available() else None
    if "qutip" in sys.modules and getattr(sys.modules["qutip"], "__file__", "").startswith(_STANDIN):
        return sys.modules["qutip"]
    saved = list(sys.path)
    try:
        sys.path.insert(0, _STANDIN)
        sys.modules.pop("qutip", None)
        return importlib.import_module("qutip")
    finally:
        sys.path[:] = saved


def run_demo_maxcut(seed=0, capture=True):
    """Execute the reference's demo_maxcut.py unmodified (demo_maxcut.py:1-89) and return its
    module namespace (sim, H_cost, H0, Hs, superposition, state, prob ...)."""
    import contextlib
    import io
    import runpy
    import numpy as np
    load_sim_plain()
    saved = list(sys.path)
    try:
        sys.path[:0] = [_STANDIN, REFERENCE_DIR]
        np.random.seed(seed)
        buf = io.StringIO()
        with (contextlib.redirect_stdout(buf) if capture else contextlib.nullcontext()):
            ns = runpy.run_path(os.path.join(REFERENCE_DIR, "demo_maxcut.py"), run_name="__main__")
        ns["__stdout__"] = buf.getvalue()
        return ns
    finally:
        sys.path[:] = saved
