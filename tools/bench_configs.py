"""Timings of the other BASELINE.json configs on the GPU box (bench.py measures configs[3]):
  configs[0]  demo_maxcut.py as shipped (n=4, 202 epochs) through the dense exact path
  configs[1]  H2 VQE, 4 qubits, 4096 batched parameter-shift samples (dense path, DMMA GEMMs)
  configs[2]  random 3-regular MaxCut n=16, per_step=512 (1024 steps per full evolution), subset of the 1024 samples
Writes gpurun_out/configs.json.  CPU figures come from the oracle (test infrastructure) on one core."""
import json
import os
import sys
import time

import numpy as np
import torch  # noqa: F401  (first import takes seconds on a fresh box: keep it out of the timed demo loop)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffquantum_b200 as dq  # noqa: E402
from oracle import restate as R  # noqa: E402

out = {}
G = lambda name: np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)

# ---- configs[0] -------------------------------------------------------------------------------------
g = G("demo_training_ref")
sim = dq.DenseSimulator(g["H0"], g["Hs"], g["omegas"], float(g["T"]), M=g["H_cost"], psi0=g["psi0"], per_step=10)
tr = dq.EnergyTrainer(sim, n_basis=6, n_epoch=202, lr=2e-2)
np.random.seed(0)
t = time.perf_counter(); tr.train_energy(); dt = time.perf_counter() - t
out["config0_demo_maxcut"] = {"epochs": 202, "seconds": dt, "epochs_per_s": 202 / dt, "cut": bin(tr.find_state()[0])[2:],
                              "max_abs_loss_diff_vs_reference_run": float(np.abs(np.array(tr.losses_energy) - g["losses_energy"]).max()),
                              "reference_cpu_note": "reference demo: 24.2 s for 202 epochs in the build container (BASELINE.md)"}

# ---- configs[1] -------------------------------------------------------------------------------------
g = G("h2_vqe_ref")
sim = dq.DenseSimulator(g["H0"], g["Hs"], g["omegas"], float(g["T"]), M=g["M"], psi0=g["psi0"], per_step=10)
K = int(os.environ.get("K2", 4096))
np.random.seed(7)
s_list = np.random.uniform(size=K) * sim.T
sim.shifted_energies(g["coeff"], s_list[:64])                      # warm-up
res = {}
for strat in (-1, 0, 1, 2, 3):
    sim.set_option("strategy", strat)
    t = time.perf_counter(); en = sim.shifted_energies(g["coeff"], s_list); dt = time.perf_counter() - t
    res["strategy_%d" % strat] = {"seconds": dt, "samples_per_s": K / dt, "chosen": sim.stat("strategy"),
                                  "gemm_TFLOPs": sim.stat("gemm_flops") / dt / 1e12,
                                  "resident_kernel_ms": sim.stat("kernel_ms") if sim.stat("strategy") == 3 else None}
sim.set_option("strategy", -1)
t = time.perf_counter()
for s in s_list[:8]:
    R.grad_mc_dense(g["H0"], list(g["Hs"]), g["M"], g["psi0"], g["coeff"], g["omegas"], float(g["T"]), float(s), 10)
cpu = 8 / (time.perf_counter() - t)
e_ref = R.grad_mc_dense(g["H0"], list(g["Hs"]), g["M"], g["psi0"], g["coeff"], g["omegas"], float(g["T"]), float(s_list[0]), 10,
                        return_energies=True)[1]
res["rel_err_vs_oracle_sample0"] = float(np.abs(en[0] - e_ref).max() / np.abs(e_ref).max())
res["cpu_oracle_samples_per_s_1core"] = cpu
res["samples"] = K
out["config1_h2_vqe"] = res

# ---- configs[2] -------------------------------------------------------------------------------------
n = 16
edges = R.random_regular_edges(n, seed=0)
prob = dq.IsingProblem.maxcut(n, edges, omega0=2 * np.pi, omega1=2 * np.pi)       # T = 1.0
sim = dq.IsingSimulator(prob, per_step=512)
coeff = np.random.default_rng(0).normal(0, 1, [len(prob.terms), 6])
K = int(os.environ.get("K3", 16))
np.random.seed(0)
s_list = np.random.uniform(size=K) * prob.T
sim.stage(coeff, s_list[:2]); sim.run_staged(); sim.fetch()
res = {}
for lin in (0, 1):
    sim.set_option("linear", lin)
    sim.stage(coeff, s_list)
    t = time.perf_counter(); sim.run_staged(); en = sim.fetch(); dt = time.perf_counter() - t
    res["linear_%d" % lin] = {"samples": K, "seconds": dt, "samples_per_s": K / dt, "trajectory_steps": sim.stat("steps"),
                              "alg_GBs": sim.stat("alg_bytes") / dt / 1e9}
psi, e = sim.evolve(coeff, 0, prob.T)
res["full_evolution_steps"] = int(512 * (prob.T + 1))
res["norm_after_%d_steps" % res["full_evolution_steps"]] = float(np.linalg.norm(psi[0]))
out["config2_n16"] = res

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
