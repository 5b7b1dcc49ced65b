"""Calibration micro-benchmarks on the B200 (copy GB/s, FP64 TFLOP/s, in-place read+write GB/s as a
function of window size -> L2 capacity seen by a streaming kernel, shared-memory read GB/s)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffquantum_b200 as dq

ctx = dq.Context.get(0)
out = {}
out["copy_GBs_1GiB"] = ctx.microbench(0, 1 << 30, 10)
out["fp64_fma_TFLOPs"] = ctx.microbench(1, 0, 4096)
out["smem_read_GBs"] = ctx.microbench(3, 0, 2000)
for mib in (8, 16, 32, 48, 64, 96, 128, 192, 256, 1024):
    out["inplace_rw_GBs_%dMiB" % mib] = ctx.microbench(2, mib << 20, 40)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
