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
buf = np.zeros((40000, 48), dtype=np.int64)
cnt = lib.dq_debug_trace(sim.handle, buf.ctypes.data_as(ctypes.c_void_p), buf.shape[0])
x = buf[:cnt, 32:]
t = buf[:cnt, :32].reshape(-1, 4, 8)
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

# TMA load: issue (thread 0 of the loading team) -> first use (warp 0 enters process_tile), by issue site and pass type
lat = t[:, 0, 1] - x[:, 0]
for site in (1, 2, 3, 4, 5, 6, 7, 9):
    for ty in (0, 1):
        m = ok[:, 0] & (x[:, 1] == site) & (x[:, 2] == ty) & (x[:, 0] > 0)
        if m.sum():
            print("site %d type %d: n %6d  issue->use mean %.0f  min %.0f  p10 %.0f  median %.0f" % (
                site, ty, m.sum(), lat[m].mean(), lat[m].min(), np.percentile(lat[m], 10), np.median(lat[m])))
            w = t[m][:, 0, 1] - t[m][:, 0, 0]
            print("        wait at top (slot0->1): mean %.0f median %.0f" % (w.mean(), np.median(w)))

m = ok[:, 0] & (x[:, 3] > 0) & (x[:, 4] > 0)
print("thread 0 at the top: slot0->tables %.0f  mbar_wait %.0f  ->slot1 %.0f" % (
    (x[m, 3] - t[m, 0, 0]).mean(), (x[m, 4] - x[m, 3]).mean(), (t[m, 0, 1] - x[m, 4]).mean()))
m5 = m & (x[:, 5] > 0)
print("  of which bulk_wait_read+release (n %d): %.0f" % (m5.sum(), (x[m5, 5] - x[m5, 4]).mean()))

# per worker (CTA, team): time outside the traced region
wk = x[:, 6]
gaps = {1: [], 2: [], 3: [], 4: [], 5: [], 6: [], 7: [], 9: []}
busy = []; span = []
first = []; last = []
for w in np.unique(wk[wk > 0]):
    idx = np.where((wk == w) & ok[:, 0])[0]
    idx = idx[np.argsort(t[idx, 0, 0])]
    t0 = t[idx, 0, 0]; t7 = t[idx, 0, 7]
    g = t0[1:] - t7[:-1]
    for k, site in zip(g, x[idx[1:], 1]):
        gaps.setdefault(int(site), []).append(k)
    busy.append((t7 - t0).sum()); span.append(t7[-1] - t0[0]); first.append(t0[0]); last.append(t7[-1])
print("workers %d  busy/span %.3f  items/worker %.1f" % (len(busy), np.sum(busy) / np.sum(span), cnt / len(busy)))
print("launch span (first start -> last end) %.0f cycles; mean worker span %.0f; start spread %.0f; end spread %.0f" % (
    max(last) - min(first), np.mean(span), max(first) - min(first), max(last) - min(last)))
for site, g in sorted(gaps.items()):
    if len(g): print("gap before an item loaded at site %d: n %6d mean %.0f median %.0f p90 %.0f" % (site, len(g), np.mean(g), np.median(g), np.percentile(g, 90)))

# inside the gap of a warm item (thread 0): end of previous item's loop body (8) -> top (9) -> after the cold-path block (10) -> slot 0
m = ok[:, 0] & (x[:, 8] > 0) & (x[:, 9] > 0) & (x[:, 10] > 0) & (x[:, 1] > 0) & (x[:, 1] < 4)
print("warm items (sites 1-3) n %d: 8->9 %.0f  9->10 %.0f  10->slot0 %.0f" % (m.sum(), (x[m, 9] - x[m, 8]).mean(), (x[m, 10] - x[m, 9]).mean(), (t[m, 0, 0] - x[m, 10]).mean()))
m = ok[:, 0] & (x[:, 8] > 0) & (x[:, 9] > 0) & (x[:, 10] > 0) & (x[:, 1] >= 4) & (x[:, 1] < 7)
print("late items (sites 4-6) n %d: 8->9 %.0f  9->10 %.0f  10->slot0 %.0f" % (m.sum(), (x[m, 9] - x[m, 8]).mean(), (x[m, 10] - x[m, 9]).mean(), (t[m, 0, 0] - x[m, 10]).mean()))
m = ok[:, 0] & (x[:, 8] > 0) & (x[:, 9] > 0) & (x[:, 10] > 0) & (x[:, 1] == 9)
print("cold items n %d: 8->9 %.0f  9->10 %.0f  10->slot0 %.0f" % (m.sum(), (x[m, 9] - x[m, 8]).mean(), (x[m, 10] - x[m, 9]).mean(), (t[m, 0, 0] - x[m, 10]).mean()))
