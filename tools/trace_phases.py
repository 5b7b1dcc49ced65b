"""DQ_TRACE build only: per-phase cycle counts of the fused kernel's items (GPU box)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import diffquantum_b200 as dq
from diffquantum_b200 import _lib
from oracle import restate as R
n = 20
prob = dq.IsingProblem.maxcut(n, R.random_regular_edges(n, seed=0))
coeff = np.random.default_rng(0).normal(0, 1, [len(prob.terms), 6])
G = int(os.environ.get("G", 4))
sim = dq.IsingSimulator(prob, per_step=10, engine=1, ket_group=G)
if os.environ.get("GPS"): sim.set_option("grid_per_sm", int(os.environ["GPS"]))
sim.stage(coeff, [1.0]); sim.run_staged(); sim.fetch()
lib = _lib.load()
lib.dq_debug_trace.restype = ctypes.c_longlong
lib.dq_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong]
buf = np.zeros((40000, 4, 8), dtype=np.int64)
cnt = lib.dq_debug_trace(sim.handle, buf.ctypes.data_as(ctypes.c_void_p), buf.shape[0])
t = buf[:cnt]
ok = (t[:, :, 0] > 0) & (t[:, :, 7] > 0)
print("items traced", cnt, "complete", ok.all(axis=1).sum())
names = ["top->landed(wait)", "outerA", "sync2", "inner", "barrier3", "rotB0+prefetch", "rotB1-4", "energy/stores"]
for w in range(4):
    m = ok[:, w]
    d = np.diff(t[m, w, :], axis=1)   # 7 intervals: 0->1 wait, 1->2 outerA, 2->3 inner(after sync2), 3->4 barrier3, 4->5, 5->6, 6->7
    tot = (t[m, w, 7] - t[m, w, 0])
    print("warp", w, "n", m.sum(), "total/item %.0f" % tot.mean(), " ".join("%s=%.0f" % (nm, v) for nm, v in zip(
        ["wait_land", "outerA", "sync2+inner", "barrier3", "rotB0+pf", "rotB1-4", "store"], d.mean(axis=0))))
# gap between consecutive items on the same CTA is not recorded; estimate from throughput
