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


# ------------------------------------------------------------------------------------------------
# The reference's DISABLED per-term product step, executed from its own lines.
#
# sim_plain.py keeps the product form as two commented statements inside `trotter`:
#   :139  # psi = expm_multiply(-1.j * dt * h[1](t,None) * h[0], psi)
#   :142  # psi = expm_multiply(-1.j * dt * h, psi)
# next to the live summed-generator code (:140, :143, :145-146, :149).  `load_sim_plain_split` reads the file,
# checks those lines are byte-for-byte what is quoted here, un-comments :139/:142 and comments the live
# statements out IN MEMORY (the file on disk is never touched), and executes the result as module
# `sim_plain_split`.  Every other line of the reference (step grid, pulse closures, estimator) runs unchanged, so
# fixtures made with it pin the split-step restatement to reference-executed output.
# ------------------------------------------------------------------------------------------------
_SPLIT_ENABLE = {
    139: "                    # psi = expm_multiply(-1.j * dt * h[1](t,None) * h[0], psi)",
    142: "                    # psi = expm_multiply(-1.j * dt * h, psi)",
}
_SPLIT_DISABLE = {
    140: "                    dH += -1.j * dt * h[1](t,None) * h[0]",
    143: "                    dH = -1.j * dt * h",
    145: "            expm = scipy.linalg.expm(dH)",
    146: "            psi = np.matmul(expm , psi)",
    149: "            dH = dH * 0",
}


def split_source():
    """Text of sim_plain.py with the per-term product enabled (see above); raises if the reference differs."""
    path = os.path.join(REFERENCE_DIR, "sim_plain.py")
    lines = open(path).read().split("\n")
    for no, want in list(_SPLIT_ENABLE.items()) + list(_SPLIT_DISABLE.items()):
        if lines[no - 1] != want:
            raise RuntimeError("sim_plain.py:%d is not the line this transform was written for: %r" % (no, lines[no - 1]))
    for no in _SPLIT_ENABLE:
        lines[no - 1] = lines[no - 1].replace("# ", "", 1)
    for no in _SPLIT_DISABLE:
        ind = len(lines[no - 1]) - len(lines[no - 1].lstrip())
        lines[no - 1] = lines[no - 1][:ind] + "pass  # " + lines[no - 1][ind:]
    return "\n".join(lines)


def load_sim_plain_split():
    """Module object of the transformed source (class SimulatorPlain with the product-form `trotter`)."""
    import types
    load_sim_plain()                           # stand-ins in sys.modules, reference importable
    saved = list(sys.path)
    try:
        sys.path[:0] = [_STANDIN, REFERENCE_DIR]
        mod = types.ModuleType("sim_plain_split")
        mod.__file__ = os.path.join(REFERENCE_DIR, "sim_plain.py") + " [split transform, in memory]"
        exec(compile(split_source(), mod.__file__, "exec"), mod.__dict__)
        return mod
    finally:
        sys.path[:] = saved
