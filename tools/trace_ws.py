"""DQ_TRACE build only: per-phase cycle counts of the warp-specialised pass kernel's tiles (GPU box).
Slots per warp: 0 top of loop, 1 tile landed, 2 after KA, 3 before J1 (exchange 1 done), 4 after J2, 5 before KB (exchange 2
done), 6 after KB, 7 after the done-arrive."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import diffquantum_b200 as dq
from diffquantum_b200 import _lib
from oracle import restate as R
n = 20
prob = dq.IsingProblem.maxcut(n, R.random_regular_edges(n, seed=0))
coeff = np.random.default_rng(0).normal(0, 1, [len(prob.terms), 6])
G = int(os.environ.get("G", 5))
sim = dq.IsingSimulator(prob, per_step=10, engine=1, ket_group=G)
sim.stage(coeff, [1.0]); sim.run_staged(); sim.fetch()
lib = _lib.load()
lib.dq_debug_trace.restype = ctypes.c_longlong
lib.dq_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong]
buf = np.zeros((600000, 48), dtype=np.int64)
cnt = lib.dq_debug_trace(sim.handle, buf.ctypes.data_as(ctypes.c_void_p), buf.shape[0])
x = buf[:cnt, 32:]
t = buf[:cnt, :32].reshape(-1, 4, 8)
ok = ((t > 0).all(axis=2)).all(axis=1)
print("items traced", cnt, "complete", ok.sum())
names = ["wait_full", "load+KA", "exch1", "J1+phase+J2", "exch2", "KB", "store+arrive"]
for ty in (0, 1):
    m = ok & (x[:, 2] == ty)
    if not m.sum():
        continue
    d = np.diff(t[m], axis=2)            # [items, warp, 7]
    tot = t[m][:, :, 7] - t[m][:, :, 0]
    print("type %d: n %d  total/tile %.0f   " % (ty, m.sum(), tot.mean()) + "  ".join("%s=%.0f" % (nm, v) for nm, v in zip(names, d.mean(axis=(0, 1)))))
    print("         p10/p50/p90 of J1+phase+J2: %s   KB: %s   exch1: %s  exch2: %s" % tuple(
        "/".join("%.0f" % np.percentile(d[:, :, k], q) for q in (10, 50, 90)) for k in (3, 5, 2, 4)))
# per worker: throughput and overlap between the two teams of a CTA
wk = x[:, 6]
spans = []
for w in np.unique(wk[wk > 0])[:64]:
    idx = np.where((wk == w) & ok)[0]
    idx = idx[np.argsort(t[idx, 0, 0])]
    if len(idx) > 4:
        spans.append((t[idx[-1], 0, 7] - t[idx[0], 0, 0]) / len(idx))
print("cycles per tile per team (first 64 workers): mean %.0f" % np.mean(spans))
# fraction of time warp 0 of team A is in a math section while warp 0 of team B (same CTA) is too
def math_intervals(idx):
    out = []
    for i in idx:
        out += [(t[i, 0, 1] + 200, t[i, 0, 2]), (t[i, 0, 3], t[i, 0, 4]), (t[i, 0, 5], t[i, 0, 6])]
    return out
both = anyone = total = 0.0
for cta in range(8):
    a = np.where((wk == 2 * cta + 1) & ok)[0]; b = np.where((wk == 2 * cta + 2) & ok)[0]
    if len(a) < 4 or len(b) < 4:
        continue
    lo = max(t[a, 0, 0].min(), t[b, 0, 0].min()); hi = min(t[a, 0, 7].max(), t[b, 0, 7].max())
    grid = np.arange(lo, hi, 64)
    ma = np.zeros(len(grid), bool); mb = np.zeros(len(grid), bool)
    for (s, e) in math_intervals(a): ma |= (grid >= s) & (grid < e)
    for (s, e) in math_intervals(b): mb |= (grid >= s) & (grid < e)
    both += (ma & mb).sum(); anyone += (ma | mb).sum(); total += len(grid)
if total:
    print("warp 0 of the two teams: both in math %.1f%%, exactly one %.1f%%, neither %.1f%% of the time" % (
        100 * both / total, 100 * (anyone - both) / total, 100 * (total - anyone) / total))
