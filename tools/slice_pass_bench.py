"""Slice rotation passes, TMA tile kernel vs the cp.async tile kernel (DQ_SLICE_NO_TMA=1): device time per call (CUDA events on the
library's stream, median of `reps`) and max difference between the two results.  GPU box only.
Usage: python tools/slice_pass_bench.py L [bits,bits,...;bits,...]   (default: all L bits, then each planned group)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from diffquantum_b200 import distributed

L = int(sys.argv[1])
groups = [list(range(L)), list(range(12)), list(range(12, min(L, 20))), list(range(20, L))]
if len(sys.argv) > 2:
    groups = [[int(b) for b in g.split(",")] for g in sys.argv[2].split(";")]
groups = [g for g in groups if g]
reps = int(os.environ.get("REPS", "7"))
ops = distributed.CudaSliceOps(0)
stream = torch.cuda.ExternalStream(ops.ctx.stream)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1)


N = 1 << L
a = ops.alloc(N)
b = ops.alloc(N)
a.copy_(torch.randn(N, dtype=torch.complex128, device=a.device))
a /= a.norm()
rng = np.random.RandomState(L)
out = []
for bits in groups:
    thetas = rng.uniform(-0.9, 0.9, size=len(bits))
    rec = {"L": L, "bits": "%d..%d (%d)" % (bits[0], bits[-1], len(bits))}
    res = {}
    for tag, env in (("cp_async", "1"), ("tma", None)):
        if env:
            os.environ["DQ_SLICE_NO_TMA"] = env
        else:
            os.environ.pop("DQ_SLICE_NO_TMA", None)
        times = []
        for r in range(reps + 1):
            b.copy_(a)
            ops.ctx.synchronize()
            torch.cuda.synchronize()
            l0 = ops.ctx.launch_count
            t = timed(lambda: ops.rx_many(b, L, bits, thetas))
            if r:
                times.append(t)
            rec[tag + "_launches"] = ops.ctx.launch_count - l0
        res[tag] = b.clone()
        ms = float(np.median(times))
        rec[tag + "_ms"] = ms
        rec[tag + "_GBs_per_pass"] = 32.0 * N * rec[tag + "_launches"] / (ms * 1e-3) / 1e9
    rec["max_diff"] = float((res["tma"] - res["cp_async"]).abs().max())
    rec["norm"] = float(res["tma"].norm())
    print(json.dumps(rec), flush=True)
    out.append(rec)
# the contiguous pass carrying the step's phase (3-regular graph on L qubits) vs the separate Gray-code phase pass + rotation pass
from oracle import restate as R
edges = R.random_regular_edges(L, seed=0)
pair_bits = np.array([[L - 1 - a, L - 1 - b] for a, b in edges], dtype=np.int32)
angles = rng.uniform(-0.3, 0.3, size=1 + len(edges))       # [c | one angle per pair]
bits = list(range(min(12, L)))
thetas = rng.uniform(-0.9, 0.9, size=len(bits))
rec = {"L": L, "bits": "phase + 0..11"}
res = {}
for tag, env, fused in (("separate_tma", None, False), ("fused_cp_async", "1", True), ("fused_tma", None, True)):
    if env:
        os.environ["DQ_SLICE_NO_TMA"] = env
    else:
        os.environ.pop("DQ_SLICE_NO_TMA", None)
    times = []
    for r in range(reps + 1):
        b.copy_(a)
        ops.ctx.synchronize()
        torch.cuda.synchronize()
        l0 = ops.ctx.launch_count
        if fused:
            t = timed(lambda: ops.phase_rx_many(b, L, 0, L, pair_bits, angles, bits, thetas))
        else:
            t = timed(lambda: (ops.phase(b, L, 0, L, pair_bits, angles), ops.rx_many(b, L, bits, thetas)))
        if r:
            times.append(t)
        rec[tag + "_launches"] = ops.ctx.launch_count - l0
    res[tag] = b.clone()
    rec[tag + "_ms"] = float(np.median(times))
rec["max_diff_fused_tma_vs_separate"] = float((res["fused_tma"] - res["separate_tma"]).abs().max())
rec["max_diff_fused_tma_vs_fused_cp_async"] = float((res["fused_tma"] - res["fused_cp_async"]).abs().max())
print(json.dumps(rec), flush=True)
out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/slice_pass_bench_L%d.json" % L, "w"), indent=1)
