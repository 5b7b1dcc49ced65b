"""configs[1] (H2 VQE, 4 qubits, K batched parameter-shift samples) through the resident dense engine; used under ncu."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffquantum_b200 as dq  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "h2_vqe_ref.npz"), allow_pickle=False)
sim = dq.DenseSimulator(g["H0"], g["Hs"], g["omegas"], float(g["T"]), M=g["M"], psi0=g["psi0"], per_step=10)
K = int(os.environ.get("K", 4096))
np.random.seed(7)
s_list = np.random.uniform(size=K) * sim.T
sim.shifted_energies(g["coeff"], s_list[:64])
for rep in range(3):
    t = time.perf_counter(); en = sim.shifted_energies(g["coeff"], s_list); dt = time.perf_counter() - t
    print("K=%d n_H=%d: %.1f ms end to end, resident kernels %.3f ms, %.0f samples/s end to end, %.0f samples/s device"
          % (K, sim.n_H, dt * 1e3, sim.stat("kernel_ms"), K / dt, K / (sim.stat("kernel_ms") * 1e-3)))
