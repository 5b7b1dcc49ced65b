"""DMMA GEMM throughput of the dense path at large dimensions (GPU box): forces the per-step propagator
strategy so that every Taylor / squaring product is a Dp x Dp x Dp complex GEMM on the FP64 tensor pipe."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import diffquantum_b200 as dq

ctx = dq.Context.get(0)
out = {}
for dim in [int(x) for x in os.environ.get("DIMS", "256,512,1024").split(",")]:
    rng = np.random.RandomState(dim)
    a = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    H0 = (a + a.conj().T) / (2 * np.sqrt(dim))
    b = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    H1 = (b + b.conj().T) / (2 * np.sqrt(dim))
    psi = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    psi /= np.linalg.norm(psi)
    u = rng.uniform(-1, 1, size=(3, 1))
    sim = dq.DenseSimulator(H0, [H1], [1.0], 1.0)
    for strat in (1, 0):
        sim.set_option("strategy", strat)
        dq.dense_evolve(ctx, H0, [H1], u[:1], 0.2, psi)
        t = time.perf_counter(); o = dq.dense_evolve(ctx, H0, [H1], u, 0.2, psi); dt = time.perf_counter() - t
        out["dim%d_strategy%d" % (dim, strat)] = {"seconds": dt, "gemm_TFLOPs": sim.stat("gemm_flops") / dt / 1e12,
                                                 "squarings": sim.stat("squarings"), "degree": sim.stat("degree"),
                                                 "norm": float(np.linalg.norm(o))}
    sim.set_option("strategy", -1)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dense_probe.json", "w"), indent=1)
