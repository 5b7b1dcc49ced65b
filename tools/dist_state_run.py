"""BASELINE configs[4]: one MaxCut state split over the ranks on its high qubits (torchrun, one rank per GPU).
n = 32 on 8 GPUs is 8 GiB of complex128 per rank (+ an equal receive buffer).  No CPU oracle exists at this size
(SURVEY H2), so the run checks size-independent properties: unit norm after every leg, the closed form of the
mixer-only evolution of the uniform state (every amplitude exp(-i sum_q theta_q) 2^(-n/2)), the closed form of
<M> for the uniform state, and that the exchange count is one all-to-all per step."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import diffquantum_b200 as dq  # noqa: E402
from diffquantum_b200 import distributed  # noqa: E402

n = int(os.environ.get("N", 32))
steps = int(os.environ.get("STEPS", 3))
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

import networkx as nx  # noqa: E402
g = nx.random_regular_graph(3, n, seed=0)
edges = sorted(tuple(sorted(e)) for e in g.edges())
prob = dq.IsingProblem.maxcut(n, edges)
coeff = np.random.default_rng(0).normal(0, 1, [len(prob.terms), 6])
st = distributed.DistributedState(prob, device=local, per_step=10)
st.fused_rx = bool(int(os.environ.get("FUSED", "1")))       # 0: one kernel per X rotation (round-1 baseline)
out = {"n": n, "world": world, "slice_GiB": 16 * (1 << st.L) / 2 ** 30, "edges": len(edges)}

# (1) uniform state: <M> = -|E|/2 exactly, norm 1
st.fill_uniform()
out["uniform_energy"] = st.energy(); out["uniform_energy_expected"] = -0.5 * len(edges)
out["uniform_norm2"] = st.norm2()

# (2) mixer-only steps on the uniform state: closed form
rows = np.zeros((2, prob.row_len)); theta = np.linspace(0.1, 0.7, n)
rows[:, 1 + prob.n_zz:] = theta
st.evolve_rows(rows)
want = np.exp(-2j * theta.sum()) * 2.0 ** (-0.5 * n)
st.ops.ctx.synchronize()          # the kernels run on the library's stream, the copies below on torch's
mine = st.psi[:4096].cpu().numpy(); tail = st.psi[-4096:].cpu().numpy()
out["mixer_only_max_rel_err"] = float(max(np.abs(mine - want).max(), np.abs(tail - want).max()) / abs(want))
out["exchanges_after_2_steps"] = st.exchanges

# (3) timed pulse evolution (full product-formula steps incl. ZZ phases)
st.fill_uniform()
rows = prob.trajectory_rows(coeff, 0.0, prob.T, 10)[:steps]
st.ops.ctx.synchronize(); torch.cuda.synchronize()
if world > 1:
    dist.barrier(device_ids=[local])
x0, b0 = st.exchanges, st.exchanged_bytes
launches0 = st.ops.ctx.launch_count
t = time.perf_counter()
st.evolve_rows(rows)
st.ops.ctx.synchronize(); torch.cuda.synchronize()
if world > 1:
    dist.barrier(device_ids=[local])
dt = time.perf_counter() - t
out["steps"] = len(rows); out["seconds_per_step"] = dt / len(rows)
out["alg_GBs_per_gpu"] = 32.0 * (1 << st.L) * len(rows) / dt / 1e9
out["fused_rx"] = st.fused_rx
out["kernel_launches_per_step"] = (st.ops.ctx.launch_count - launches0) / len(rows)
out["exchange_GB_per_step_per_gpu"] = (st.exchanged_bytes - b0) / max(1, st.exchanges - x0) / 1e9
out["exchanges"] = st.exchanges - x0
out["norm2_after"] = st.norm2()
out["energy_after"] = st.energy()
if rank == 0:
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dist_state_n%d_w%d_fused%d.json" % (n, world, int(st.fused_rx))), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
