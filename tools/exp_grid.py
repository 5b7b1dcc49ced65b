import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import diffquantum_b200 as dq
from oracle import restate as R
n = 20; B = 2
prob = dq.IsingProblem.maxcut(n, R.random_regular_edges(n, seed=0))
coeff = np.random.default_rng(0).normal(0, 1, [len(prob.terms), 6])
s_list = [1.0, 0.9]
for gps in (1, 2):
    for G in (2, 4):
        sim = dq.IsingSimulator(prob, per_step=10, engine=1, ket_group=G)
        sim.set_option("grid_per_sm", gps)
        sim.stage(coeff, s_list); sim.run_staged(); sim.fetch()
        t = time.time(); sim.run_staged(); sim.fetch(); dt = time.time() - t
        print(json.dumps(dict(grid_per_sm=gps, G=G, us_per_ket_step=dt / sim.stat("steps") * 1e6)))
