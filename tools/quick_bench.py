"""Quick n=20 timing of the gradient-sample driver for engine / ket_group choices (GPU box)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import diffquantum_b200 as dq
from oracle import restate as R

n = int(os.environ.get("N", 20))
B = int(os.environ.get("B", 4))
edges = R.random_regular_edges(n, seed=0)
prob = dq.IsingProblem.maxcut(n, edges)
coeff = np.random.default_rng(0).normal(0, 1, [len(prob.terms), 6])
np.random.seed(1)
s_list = np.random.uniform(size=B) * prob.T
res = []
base = None
engines = [int(x) for x in os.environ.get("ENGINES", "1").split(",")]
groups = [int(x) for x in os.environ.get("KG", "2,3,4,5,6,8").split(",")]
for engine, G in [(0, 1)] + [(e, g) for e in engines for g in groups]:
    if engine == 0 and os.environ.get("SKIP_GENERIC"):
        continue
    sim = dq.IsingSimulator(prob, per_step=10, engine=engine, ket_group=G)
    sim.stage(coeff, s_list)
    sim.run_staged(); en = sim.fetch()
    t = time.time()
    sim.run_staged(); en = sim.fetch()
    dt = time.time() - t
    steps = sim.stat("steps")
    gbs = sim.stat("alg_bytes") / dt / 1e9
    if base is None:
        base = en
    err = np.abs(en - base).max() / np.abs(base).max()
    r = dict(engine=sim.info("engine"), G=G, samples_per_s=B / dt, us_per_ket_step=dt / steps * 1e6, alg_GBs=gbs,
             launches=sim.stat("launches"), rel_vs_first=err)
    print(json.dumps(r)); res.append(r)
    del sim
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/quick_bench_n%d.json" % n, "w"), indent=1)
