"""TMA tile kernel vs the cp.async tile kernel (DQ_SLICE_NO_TMA=1) on one slice: max difference and time per call (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diffquantum_b200 import distributed
L = int(sys.argv[1])
bits = list(range(L)) if len(sys.argv) < 3 else [int(b) for b in sys.argv[2].split(",")]
ops = distributed.CudaSliceOps(0)
import torch
rng = np.random.RandomState(L)
a = ops.alloc(1 << L); b = ops.alloc(1 << L)
a.copy_(torch.randn(1 << L, dtype=torch.complex128, device=a.device)); a /= a.norm(); b.copy_(a)
torch.cuda.synchronize()
thetas = rng.uniform(-1.2, 1.2, size=len(bits))
for env, t in (("1", b), (None, a)):
    if env: os.environ["DQ_SLICE_NO_TMA"] = env
    else: os.environ.pop("DQ_SLICE_NO_TMA", None)
    ops.ctx.synchronize(); t0 = time.time()
    l0 = ops.ctx.launch_count
    ops.rx_many(t, L, bits, thetas)
    ops.ctx.synchronize()
    print("L=%d tma=%s launches=%d  %.3f ms" % (L, env is None, ops.ctx.launch_count - l0, 1e3 * (time.time() - t0)), flush=True)
print("max diff", float((a - b).abs().max()), "norm", float(a.norm()))
