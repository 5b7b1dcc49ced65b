"""diffquantum_b200 — B200-native (sm_100a) state-vector evolution and batched stochastic
parameter-shift gradients behind diffquantum's own interfaces (diffqc.set_H / diffqc.trotter,
SimulatorPlain.my_solver).  Host code is Python; all amplitudes live on the device and every
kernel is hand-written CUDA reached through the C ABI in include/diffqc_b200.h.
There is no CPU fallback: importing is cheap, but using any solver without the built shared
library and a B200 raises."""
from . import pulses  # noqa: F401
from ._lib import Context, DiffqcError, load  # noqa: F401
from .ising import IsingProblem, IsingSimulator  # noqa: F401
from .dense import DenseSimulator, dense_evolve, estimator_for, solver_for  # noqa: F401
from . import diffqc, sharding  # noqa: F401
from .training import EnergyTrainer  # noqa: F401
from . import comparators  # noqa: F401
from .comparators import FDTrainer, FidelityTrainer  # noqa: F401

__version__ = "dev"
