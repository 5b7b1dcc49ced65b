"""Small fixed workload for ncu captures: n=20, 1 sample, fused engine (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import diffquantum_b200 as dq
from oracle import restate as R

n = int(os.environ.get("N", 20))
G = int(os.environ.get("G", 4))
edges = R.random_regular_edges(n, seed=0)
prob = dq.IsingProblem.maxcut(n, edges)
coeff = np.random.default_rng(0).normal(0, 1, [len(prob.terms), 6])
sim = dq.IsingSimulator(prob, per_step=10, engine=int(os.environ.get("ENGINE", 1)), ket_group=G)
sim.stage(coeff, [1.0])
sim.run_staged()
print(sim.fetch()[0, :2])
