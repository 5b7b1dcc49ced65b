#!/usr/bin/env python
"""bench.py — gradient samples/s of the stochastic parameter-shift estimator at n=20 MaxCut
(BASELINE.json configs[3]: random 3-regular MaxCut, n=20, per_step=10 as shipped, T=2, B-spline
basis with 6 coefficients per control, 50 controls -> 1 + 100 trajectories per sample).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, sm_100a), BASELINE configs[3] (the headline)
  python bench.py --impl reference [...]                          reference arm (CPU, host cores)
  python bench.py --config {0,1,2,4} [...]                        the other BASELINE configs, same JSON contract:
      0  demo_maxcut.py as shipped (n=4, 202 epochs)              metric: training epochs/s
      1  H2 VQE, 4 qubits, 4096 batched samples (dense path)      metric: gradient samples/s
      2  n=16 MaxCut, 1024 Trotter steps, 1024 samples            metric: gradient samples/s
      4  one oversize state split on its high qubits (n=32 at 8 GPUs, 29 + log2 N below)   metric: product-formula steps/s
  python bench.py --scaling strong [...]                          configs[3] with the GLOBAL batch fixed (64 samples per step)

A "step" is one batch of --samples-per-step gradient samples PER GPU (weak scaling): the sample
times of step i are the reference's own stream, np.random.seed(i); np.random.uniform(size=B*N)*T
(sim_plain.py:167), split contiguously over the N ranks.  No rank talks to another until the
[n_Hs, n_basis] gradient sum is all-reduced once per step (NCCL).

value  : samples/s with the angle tables already staged in HBM; device time from CUDA events on the
         library's stream, summed over the K steps, max over ranks; L2 flushed between steps.
e2e    : the same K steps through the public API (IsingSimulator.grad_samples + the sharded
         reduce) with host buffers: host pulse-table evaluation, H2D of the tables, kernels, D2H of
         the shifted energies, gradient assembly and the all-reduce are all inside the timed region.
roofline: the fused pass kernel (k_fused_passes), algorithmic bytes = trajectory-steps x 2 x 16 B x 2^n
         per launch over its CUDA-event-timed launch durations in the same timed region.
cpu_baseline / --impl reference: the reference's step (diffqc.cc:155-164) and estimator
         (sim_plain.py:186-230) as the plain-C OpenMP port in oracle/c on all host cores, on a bounded
         sample (the reference's own dense code cannot represent n=20: SURVEY F3).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_JSON_OUT = None          # the process's real stdout once fd 1 has been pointed at stderr (N > 1: NCCL prints to fd 1)


def emit_json(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "gradient samples/s at n=20 MaxCut"        # configs[3], the headline; the other configs name theirs (metric_of)
UNIT = "samples/s"


def metric_of(a):
    if a.config == 0:
        return "training epochs/s, demo_maxcut.py as shipped (n=4, 202 epochs)", "epochs/s"
    if a.config == 1:
        return "gradient samples/s, H2 VQE 4 qubits (dense path)", "samples/s"
    if a.config == 2:
        return "gradient samples/s at n=16 MaxCut, 1024 Trotter steps", "samples/s"
    if a.config == 4:
        return "product-formula steps/s of one state split over the GPUs on its high qubits", "steps/s"
    return (METRIC if a.n == 20 else "gradient samples/s at n=%d MaxCut" % a.n), UNIT


FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[0, 1, 2, 3, 4], help="index into BASELINE.json configs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="configs[2,3]: weak = --samples-per-step per GPU; strong = the same GLOBAL batch at every N")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--per-step", type=int, default=None)
    ap.add_argument("--samples-per-step", type=int, default=None,
                    help="gradient samples per GPU per step (strong scaling: per step over all GPUs)")
    ap.add_argument("--cpu-terms", type=int, default=6,
                    help="controls whose +/- trajectories the bounded CPU sample runs (of n_Hs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="configs[4]: the per-step exchange as stores of the last local pass into peer memory (default) or as an "
                         "NCCL all_to_all_single")
    ap.add_argument("--ket-group", type=int, default=0)
    ap.add_argument("--item-tiles-log2", type=int, default=-1, help="fused v2: tiles per work item = 2^k (default: library default)")
    ap.add_argument("--engine", type=int, default=1, help="1 = the fused pass engine (the only product engine at n=20)")
    a = ap.parse_args()
    # per-config defaults: config 2 = omega 2 pi -> T = 1, per_step 512 -> 1024 steps per full evolution
    dflt = {3: dict(n=20, per_step=10, sps=8), 2: dict(n=16, per_step=512, sps=128), 1: dict(n=4, per_step=10, sps=4096),
            0: dict(n=4, per_step=10, sps=1), 4: dict(n=None, per_step=10, sps=1)}[a.config]
    if a.n is None:
        a.n = dflt["n"]
    if a.per_step is None:
        a.per_step = dflt["per_step"]
    if a.samples_per_step is None:
        a.samples_per_step = dflt["sps"] * (8 if (a.scaling == "strong" and a.config == 3) else 1)
    return a


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def workload(n):
    import networkx as nx
    g = nx.random_regular_graph(3, n, seed=0)
    edges = sorted(tuple(sorted(e)) for e in g.edges())
    n_H = len(edges) + n
    coeff = np.random.default_rng(0).normal(0, 1, [n_H, 6])
    return edges, coeff


def omega_of(a):
    """configs[2]: omega = 2 pi -> T = 1.0 (1024 steps at per_step 512); configs[3]: omega = pi -> T = 2.0 as shipped."""
    return 2 * np.pi if a.config == 2 else np.pi


def per_gpu_samples(a, world):
    """samples each rank sees per step: fixed per GPU (weak) or the global batch split over the ranks (strong)."""
    return a.samples_per_step if a.scaling == "weak" else max(1, a.samples_per_step // world)


def step_samples(step, per_gpu, world, T):
    np.random.seed(step)
    return np.random.uniform(size=per_gpu * world) * T


def steps_of_sample(s, T, per_step, n_H):
    return int(per_step * (s + 1)), int(per_step * (T - s + 1)), \
        int(per_step * (s + 1)) + 2 * n_H * int(per_step * (T - s + 1))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_copy_gbs", "hbm_GBs"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json %s)" % k
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


class ClockSampler(object):
    """nvidia-smi polled in the background for the whole run (its start-up enumerates every GPU and can
    stall launches for hundreds of ms, so it is started BEFORE the warm-up steps, never inside a timed
    region); `window(t0, t1)` summarises the samples whose timestamps fall inside a timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.rows = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            import select
            ready, _, _ = select.select([self.proc.stdout], [], [], 10.0)      # first sample = start-up is over
            self.first = self.proc.stdout.readline() if ready else ""
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None or self.rows is not None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        out = self.first + out
        import datetime
        self.rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                self.rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[4:8]))
            except ValueError:
                continue

    def window(self, t0, t1):
        self.stop()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        sel = [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.05] or self.rows
        reasons = set()
        for r in sel:
            for nm, v in zip(names, r[4]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median([r[1] for r in sel])), "sm_max_mhz": float(max(r[2] for r in sel)),
                "power_w_max": float(max(r[3] for r in sel)), "samples": len(sel), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: plain-C OpenMP port of the reference step + estimator (oracle/c), bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_bounded_sample(n, per_step, coeff, edges, s, n_terms, omega=np.pi):
    from oracle import c_port as C, restate as R
    if not C.available():
        raise RuntimeError("oracle/c/liboracle_c.so missing: run __graft_entry__.build()")
    C.use_host_cores()                      # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    prob = R.maxcut_structured(n, edges, omega0=omega, omega1=omega)
    cp = C.CProblem(prob)
    n_H = len(prob["terms"])
    # one ZZ control and one X control first, then alternate, so the sample sees both gate kinds
    order = []
    zz = [i for i, t in enumerate(prob["terms"]) if t[0] == "zz"]
    xx = [i for i, t in enumerate(prob["terms"]) if t[0] == "x"]
    while len(order) < n_terms and (zz or xx):
        if zz:
            order.append(zz.pop(0))
        if xx and len(order) < n_terms:
            order.append(xx.pop(0))
    t0 = time.perf_counter()
    _, energies, steps = C.grad_mc(cp, coeff, float(s), per_step, terms=order, return_energies=True)
    dt = time.perf_counter() - t0
    full = steps_of_sample(s, prob["T"], per_step, n_H)[2]
    return dict(seconds=dt, steps=steps, full_steps=full, samples_per_s=(steps / dt) / full,
                cores=C.num_threads(), n_terms=len(order), n_H=n_H, order=order, energies=energies[order])


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if a.config in (0, 1):
        return run_reference_dense(a)
    if a.config == 4:
        print(json.dumps({"impl": "reference", "unavailable": "configs[4] (n >= 29 per GPU, 64 GiB at n=32) has no CPU "
                          "reference: the state does not fit the host and the reference's dense operators stop at n~13 (SURVEY H2)"}))
        return 0
    edges, coeff = workload(a.n)
    T = np.pi * 2.0 / omega_of(a)
    vals, last = [], None
    for i in range(a.warmup + a.steps):
        # warm-up steps use a reduced sample (page-in, OpenMP pool start-up); timed ones the bounded sample
        s = step_samples(i, a.samples_per_step, 1, T)[0]
        r = cpu_bounded_sample(a.n, a.per_step, coeff, edges, s, 1 if i < a.warmup else (a.cpu_terms if a.config == 3 else 2), omega_of(a))
        if i >= a.warmup:
            vals.append(r)
            last = r
    tot_steps = sum(r["steps"] for r in vals)
    tot_sec = sum(r["seconds"] for r in vals)
    mean_full = float(np.mean([r["full_steps"] for r in vals]))
    v = (tot_steps / tot_sec) / mean_full
    sample = ("per step: 1 prefix + the +/- trajectories of %d of %d controls of one sample (%d of ~%d "
              "trajectory-steps), scaled by steps" % (last["n_terms"], last["n_H"], last["steps"], last["full_steps"]))
    line = {
        "impl": "reference", "metric": metric_of(a)[0], "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tot_sec / a.steps, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, 1),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": sample,
                         "build": "gcc -O3 -march=native -fopenmp (oracle/c/Makefile), rebuilt on this host"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference's own dense code cannot build n=20 operators (SURVEY F3); this is the plain-C OpenMP port "
                "of its per-term product step (diffqc.cc:155-164) and estimator (sim_plain.py:186-230)",
    }
    print(json.dumps(line))
    return 0


def workload_config(a, world):
    T = np.pi * 2.0 / omega_of(a)
    per_gpu = per_gpu_samples(a, world)
    full = {3: "8192 gradient samples", 2: "1024 gradient samples"}[a.config]
    return {"workload": "configs[%d]: random 3-regular MaxCut n=%d (networkx seed 0), per_step=%d, T=%.1f (%d steps per full "
                        "evolution), n_basis=6 BSpline, 1+2*n_Hs trajectories per sample" % (
                            a.config, a.n, a.per_step, T, int(a.per_step * (T + 1))),
            "samples_per_step_per_gpu": per_gpu, "global_samples_per_step": per_gpu * world,
            "samples_timed": per_gpu * world * a.steps,
            "of_the_configs_batch": "%s in BASELINE.json; samples are independent and identically distributed in cost, so the "
                                    "timed subset measures the same per-sample work" % full,
            "parallelism": "sample-sharded x%d (cost-balanced shards), one all-reduce of the gradient per step" % world,
            "l2": "L2 flushed (512 MiB write) between timed steps"}


# ---------------------------------------------------------------------------------------------
# shared plumbing for the other configs
# ---------------------------------------------------------------------------------------------
class Ranks(object):
    """torch.distributed bring-up as the driver launches it (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* in the env)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            global _JSON_OUT
            sys.stdout.flush()
            _JSON_OUT = os.fdopen(os.dup(1), "w")        # NCCL prints its banner to fd 1: keep the JSON line apart
            os.dup2(2, 1)
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(device_ids=[self.local])
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, x, op="max"):
        if self.world == 1:
            return float(x)
        import torch.distributed as dist
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)


def fp64_peak_tflops(ctx):
    """FP64 FMA rate of this GPU, measured now (dq_microbench kind 1: 8 x 256 threads per SM of dependent-free DFMA)."""
    try:
        return float(ctx.microbench(1, iters=4000)), "measured now (dq_microbench: DFMA issue rate, 2 flops per FMA)"
    except Exception as e:           # pragma: no cover
        return 36.0, "fallback 36 TFLOP/s (round-1 measurement); microbench failed: %s" % e


# ---------------------------------------------------------------------------------------------
# configs[0] and configs[1]: the dense path (dim = 16)
# ---------------------------------------------------------------------------------------------
def dense_problem(a):
    g = golden("demo_training_ref" if a.config == 0 else "h2_vqe_ref")
    M = g["H_cost"] if a.config == 0 else g["M"]
    coeff = None if a.config == 0 else g["coeff"]
    return dict(H0=g["H0"], Hs=g["Hs"], M=M, psi0=g["psi0"], omegas=g["omegas"], T=float(g["T"]), coeff=coeff, g=g)


def dense_ket_steps(s_list, T, per_step, n_H):
    pre = (per_step * (s_list + 1)).astype(np.int64)
    suf = (per_step * ((T - s_list) + 1)).astype(np.int64)
    return int(pre.sum() + 2 * n_H * suf.sum())


def _cpu_dense_samples(args):
    """Worker of the CPU arm of configs[1]: the oracle's restatement of compute_energy_grad_MC (sim_plain.py:156-231)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import restate as R
    P, s_list, per_step = args
    t0 = time.perf_counter()
    for s in s_list:
        R.grad_mc_dense(P["H0"], list(P["Hs"]), P["M"], P["psi0"], P["coeff"], P["omegas"], P["T"], float(s), per_step)
    return time.perf_counter() - t0


def cpu_dense_rate(P, per_step, n_workers, per_worker):
    """samples/s of the CPU oracle over n_workers processes (one thread each), per_worker samples per process."""
    import multiprocessing as mp
    P = {k: v for k, v in P.items() if k != "g"}
    rng = np.random.RandomState(123)
    jobs = [(P, rng.uniform(size=per_worker) * P["T"], per_step) for _ in range(n_workers)]
    t0 = time.perf_counter()
    if n_workers == 1:
        _cpu_dense_samples(jobs[0])
    else:
        with mp.get_context("spawn").Pool(n_workers) as pool:
            pool.map(_cpu_dense_samples, jobs)
    return n_workers * per_worker / (time.perf_counter() - t0)


def _cpu_demo_epochs(P, per_step, n_epoch):
    """The reference's training loop (sim_plain.py:245-305) restated on the CPU oracle; seconds per epoch."""
    import torch
    from oracle import restate as R
    np.random.seed(0)
    n_H = len(P["Hs"])
    coeff = torch.tensor(np.random.normal(0, 1e-3, [n_H, 6]), requires_grad=True)
    opt = torch.optim.Adam([coeff], lr=2e-2)
    t0 = time.perf_counter()
    for _ in range(n_epoch):
        c = coeff.detach().numpy().copy()
        R.trotter_plain(P["H0"], list(P["Hs"]), c, P["omegas"], P["T"], P["psi0"], 0, P["T"], per_step)
        s = np.random.uniform() * P["T"]
        g = R.grad_mc_dense(P["H0"], list(P["Hs"]), P["M"], P["psi0"], c, P["omegas"], P["T"], s, per_step)
        opt.zero_grad()
        coeff.grad = torch.from_numpy(g)
        opt.step()
        np.linalg.eigvalsh(P["M"])                # M.eigenenergies() every epoch, sim_plain.py:294
    return (time.perf_counter() - t0) / n_epoch


def run_reference_dense(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    P = dense_problem(a)
    metric, unit = metric_of(a)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    vals = []
    for i in range(a.warmup + a.steps):
        if a.config == 1:
            v = cpu_dense_rate(P, a.per_step, cores, 2 if i < a.warmup else 6)
            sample = "%d processes x 6 samples of the 4096 (oracle restatement of compute_energy_grad_MC, one thread each)" % cores
            used = cores
        else:
            v = 1.0 / _cpu_demo_epochs(P, a.per_step, 3 if i < a.warmup else 12)
            sample = "12 of the 202 epochs (restated train_energy loop: full evolution + one gradient sample + Adam + eigvalsh)"
            used = 1
        if i >= a.warmup:
            vals.append(v)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dense_config(a, 1), "cpu_baseline": {"value": v, "unit": unit, "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "NumPy/SciPy restatement of the reference's dense path (oracle/restate.py, pinned to the reference's own "
                    "outputs); the reference tree itself does not travel to the GPU box"}
    print(json.dumps(line))
    return 0


def dense_config(a, world):
    if a.config == 0:
        return {"workload": "configs[0]: demo_maxcut.py as shipped -- 4-qubit ring MaxCut, 8 controls, n_basis=6 BSpline, per_step=10, "
                            "202 epochs of (full evolution + energy + one stochastic gradient sample + Adam), np.random.seed(0)",
                "parallelism": "replicas x%d (the dense path does not shard: SURVEY 8e)" % world,
                "l2": "working set < 1 MB: L2 residency is the design, nothing to flush"}
    return {"workload": "configs[1]: H2 VQE, 4 qubits (dim 16), 8 controls (X_q, Y_q), n_basis=6 BSpline, per_step=10, T=1.5, "
                        "%d batched parameter-shift samples per step (1 + 16 trajectories each)" % a.samples_per_step,
            "samples_per_step_per_gpu": a.samples_per_step, "global_samples_per_step": a.samples_per_step * world,
            "parallelism": "replicas x%d (the dense path does not shard: SURVEY 8e)" % world,
            "l2": "working set (pulse tables + 17 kets per sample) < L2 by design; FP64-bound, not HBM-bound"}


def run_dense_config(a):
    import diffquantum_b200 as dq
    R_ = Ranks()
    P = dense_problem(a)
    metric, unit = metric_of(a)
    sim = dq.DenseSimulator(P["H0"], P["Hs"], P["omegas"], P["T"], M=P["M"], psi0=P["psi0"], per_step=a.per_step, device=R_.local)
    n_H = sim.n_H
    sampler = ClockSampler(R_.local)
    dev_s, e2e_s, flops, launches, h2d, d2h = [], [], 0.0, 0, 0, 0
    parity = None
    t_start = time.time()
    for i in range(a.warmup + a.steps):
        timed = i >= a.warmup
        if i == a.warmup:
            R_.barrier()
            t_start = time.time()
        l0 = sim.ctx.launch_count
        if a.config == 1:
            np.random.seed(i)
            s_list = np.random.uniform(size=a.samples_per_step) * sim.T
            R_.barrier()
            t0 = time.perf_counter()
            grads, en = sim.grad_samples(P["coeff"], s_list, return_energies=True)
            dt = time.perf_counter() - t0
            units = len(s_list)
            kms = sim.stat("kernel_ms")
            steps_i = dense_ket_steps(s_list, sim.T, a.per_step, n_H)
            fl = steps_i * sim.stat("degree") * (2.0 ** sim.stat("squarings")) * 8.0 * sim.dim * sim.dim
            bi = ((a.per_step * (s_list + 1)).astype(int).sum() + (a.per_step * (sim.T - s_list + 1)).astype(int).sum()) * n_H * 8
            bo = en.nbytes
        else:
            np.random.seed(0)
            tr = dq.EnergyTrainer(sim, n_basis=6, n_epoch=202, lr=2e-2, device_resident=True)   # dq_dense_train: the whole loop on the device
            R_.barrier()
            t0 = time.perf_counter()
            tr.train_energy()
            dt = time.perf_counter() - t0
            units = 202
            kms = getattr(tr, "device_ms", 0.0) or dt * 1e3
            fl, bi, bo = 0.0, 202 * 8 + 8 * 6 * 8, 202 * 8 + 8 * 6 * 8 + 16 * 16      # sample times + coefficients in; losses, coefficients, state out
            if timed:
                ref = P["g"]["losses_energy"]
                parity = {"max_abs_loss_diff": float(np.abs(np.array(tr.losses_energy) - ref).max()), "tol": 1e-8,
                          "what": "202-epoch loss_energy trajectory vs the reference's own demo_maxcut.py run (tests/golden/"
                                  "demo_training_ref.npz, np.random.seed(0))", "cut": bin(tr.find_state()[0])[2:]}
                parity["ok"] = bool(parity["max_abs_loss_diff"] < 1e-8)
        if timed:
            dev_s.append(kms * 1e-3)
            e2e_s.append(dt)
            flops += fl
            launches += sim.ctx.launch_count - l0
            h2d, d2h = int(bi), int(bo)
    R_.barrier()
    window = (t_start, time.time())
    t_dev = R_.reduce(sum(dev_s))
    t_e2e = R_.reduce(sum(e2e_s))
    total = units * a.steps * R_.world
    if a.config == 1:
        from oracle import restate as R
        e_ref = R.grad_mc_dense(P["H0"], list(P["Hs"]), P["M"], P["psi0"], P["coeff"], P["omegas"], P["T"], float(s_list[0]),
                                a.per_step, return_energies=True)[1]
        err = float(np.abs(en[0] - e_ref).max() / np.abs(e_ref).max())
        parity = {"max_rel_err": err, "tol": 1e-10, "ok": bool(err < 1e-10),
                  "what": "shifted energies of sample 0 of the last step: CUDA (dq_dense_grad) vs the oracle's restatement of "
                          "compute_energy_grad_MC (pinned to the reference's own run, tests/golden/h2_vqe_ref.npz)"}
    if parity and not parity["ok"]:
        raise SystemExit("bench.py: GPU result differs from the oracle: %r" % (parity,))
    peak, peak_src = fp64_peak_tflops(sim.ctx)
    ach = flops / t_dev / 1e12 if (flops and t_dev) else None
    line = {"metric": metric, "value": total / t_dev, "unit": unit, "n_gpus": R_.world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dense_config(a, R_.world),
            "e2e": {"value": total / t_e2e, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / a.steps,
                    "note": "public API with host buffers: host pulse tables + H2D + kernels + D2H + gradient assembly"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64_tensor",
                         "kernel": ("k_small_mma (resident engine: the shifted kets of a sample through the Horner recurrence on the FP64 "
                                    "tensor cores, DMMA m8n8k4) + k_small (prefix kets, DFMA)") if sim.stat("strategy") == 3 else "k_zgemm (DMMA)",
                         "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
                         "peak_source": peak_src, "traffic": None,
                         "flops_note": "degree x 2^squarings complex 16x16 mat-vecs per ket-step, 8 real flops per complex MAC; peak = the "
                                       "measured FP64 FMA rate of this device (dq_microbench), the DMMA pipe has the same nominal rate"},
            "parity": parity, "clocks": sampler.window(*window)}
    if a.config == 0:
        line["value_note"] = ("device time of the 202-epoch loop (CUDA events around the 1212 launches dq_dense_train enqueues); e2e = wall "
                              "time of EnergyTrainer.train_energy (RNG draws, uploads, the loop, downloads)")
    if R_.rank == 0 and R_.world == 1 and not a.no_cpu_baseline:
        try:
            if a.config == 1:
                v = cpu_dense_rate(P, a.per_step, 1, 8)
                line["cpu_baseline"] = {"value": v, "unit": unit, "cores": 1, "kind": "port",
                                        "sample": "8 of the %d samples, oracle restatement of compute_energy_grad_MC on one core" % a.samples_per_step}
            else:
                v = 1.0 / _cpu_demo_epochs(P, a.per_step, 12)
                line["cpu_baseline"] = {"value": v, "unit": unit, "cores": 1, "kind": "port",
                                        "sample": "12 of the 202 epochs of the restated train_energy loop on one core"}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": unit, "cores": 0, "kind": "port", "sample": "failed: %s" % e}
    if R_.rank == 0:
        emit_json(line)
    R_.close()
    return 0


# ---------------------------------------------------------------------------------------------
# configs[4]: one state split over the GPUs on its high qubits
# ---------------------------------------------------------------------------------------------
def run_distributed_state(a):
    import diffquantum_b200 as dq
    from diffquantum_b200 import distributed
    import networkx as nx
    R_ = Ranks()
    torch = R_.torch
    g_bits = R_.world.bit_length() - 1
    n = a.n if a.n else 29 + g_bits                  # 8 GiB of complex128 per GPU: n = 32 at 8 GPUs
    metric, unit = metric_of(a)

    # ---- parity first, at a size the oracle can check (same code path, same world size) ----------------
    from oracle import restate as R
    n_small = 16
    e_small = R.random_regular_edges(n_small, seed=n_small)
    p_small = dq.IsingProblem.maxcut(n_small, e_small)
    ref = R.maxcut_structured(n_small, e_small)
    c_small = np.random.RandomState(n_small).normal(0, 1, [len(p_small.terms), 6])
    st = distributed.DistributedState(p_small, device=R_.local, per_step=2)
    st.fill_uniform()
    st.evolve(c_small, 0.2, 1.7)
    ns, dt_s, ts = R.step_grid(0.2, 1.7, 2)
    want = R.evolve_split_structured(ref, R.coef_table_plain(c_small, ref["omegas"], ref["T"], ts), dt_s, ref["psi0"])
    N_loc = 1 << (n_small - g_bits)
    err = float(np.abs(st.local_slice() - want[R_.rank * N_loc:(R_.rank + 1) * N_loc]).max() / np.abs(want).max())
    err = R_.reduce(err)
    e_err = abs(st.energy() - R.energy_diag(ref["m_diag"], want))
    parity = {"max_rel_err": err, "energy_abs_err": float(e_err), "tol": 1e-10, "ok": bool(err < 1e-10 and e_err < 1e-10),
              "what": "DistributedState at n=%d over %d GPU(s), %d product-formula steps: every rank's slice and <M> vs the oracle "
                      "(oracle/restate.py evolve_split_structured = diffqc.cc:155-164)" % (n_small, R_.world, ns)}
    if not parity["ok"]:
        raise SystemExit("bench.py: distributed state differs from the oracle: %r" % (parity,))
    del st

    # ---- the timed state ---------------------------------------------------------------------------------
    if n % 2 == 0:
        gr = nx.random_regular_graph(3, n, seed=0)
        edges = sorted(tuple(sorted(e)) for e in gr.edges())
    else:                                            # no 3-regular graph on an odd number of nodes: n-1 regular + one node of degree 3
        gr = nx.random_regular_graph(3, n - 1, seed=0)
        edges = sorted(tuple(sorted(e)) for e in gr.edges()) + [(0, n - 1), (1, n - 1), (2, n - 1)]
    prob = dq.IsingProblem.maxcut(n, edges)
    coeff = np.random.default_rng(0).normal(0, 1, [len(prob.terms), 6])
    st = distributed.DistributedState(prob, device=R_.local, per_step=a.per_step,
                                      peer_exchange=(False if a.exchange == "nccl" else None))
    st.fill_uniform()
    rows = prob.trajectory_rows(coeff, 0.0, prob.T, a.per_step)
    sampler = ClockSampler(R_.local)
    k = 0
    times = []
    t_start = time.time()
    l0 = st.ops.ctx.launch_count
    x0, b0 = st.exchanges, st.exchanged_bytes
    if R_.world == 1:
        # nothing to exchange: the public call is evolve_rows, which chains the passes over the step boundaries
        # (dq_slice_evolve_steps); warm-up and the timed steps are one call each
        take = lambda k0, cnt: np.stack([rows[(k0 + j) % len(rows)] for j in range(cnt)])
        st.evolve_rows(take(0, a.warmup))
        st.ops.ctx.synchronize()
        torch.cuda.synchronize(R_.dev)
        t_start = time.time()
        l0 = st.ops.ctx.launch_count
        t0 = time.perf_counter()
        st.evolve_rows(take(a.warmup, a.steps))
        st.ops.ctx.synchronize()
        torch.cuda.synchronize(R_.dev)
        times.append(time.perf_counter() - t0)
    for i in range(a.warmup + a.steps if R_.world > 1 else 0):
        if i == a.warmup:
            t_start = time.time()
            l0 = st.ops.ctx.launch_count
            x0, b0 = st.exchanges, st.exchanged_bytes
        st.ops.ctx.synchronize()
        R_.barrier()
        t0 = time.perf_counter()
        st.step(rows[k % len(rows)])
        # the rotations a step leaves owed to the next step's first pass: every timed step but the last carries the ones of
        # its predecessor, the last one also runs its own, so the K timed steps contain exactly the work of K steps
        if i == a.warmup - 1 or i == a.warmup + a.steps - 1:
            st.flush()
        st.ops.ctx.synchronize()
        torch.cuda.synchronize(R_.dev)
        dt = time.perf_counter() - t0
        k += 1
        if i >= a.warmup:
            times.append(dt)
    R_.barrier()
    window = (t_start, time.time())
    t_tot = R_.reduce(sum(times))
    launches = st.ops.ctx.launch_count - l0
    norm2 = st.norm2()
    slice_bytes = 16.0 * (1 << st.L)
    peak, peak_src = measured_peak()
    passes = getattr(st, "last_passes", None)
    exch_bytes = (st.exchanged_bytes - b0) / max(1, a.steps)
    exch_ms = getattr(st, "exchange_ms", None)
    ach = 2 * slice_bytes * a.steps / t_tot / 1e9
    line = {"metric": metric, "value": a.steps / t_tot, "unit": unit, "n_gpus": R_.world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * t_tot / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "configs[4]: one MaxCut state of n=%d qubits (random 3-regular graph, networkx seed 0), %.1f GiB of "
                                   "complex128 per GPU, product-formula steps of the pulse trajectory (per_step=%d)" % (
                                       n, slice_bytes / 2 ** 30, a.per_step),
                       "parallelism": "state split on its %d high-order qubits over %d GPUs; one exchange of (W-1)/W of the slice per "
                                      "step: %s" % (g_bits, R_.world, "stores of the last local rotation pass into the peers' "
                                                    "buffers (peer memory over NVLink) + one barrier" if st.peer_exchange else
                                                    "NCCL all_to_all_single between two stream synchronisations"),
                       "l2": "slice (%.1f GiB) >> L2" % (slice_bytes / 2 ** 30)},
            "e2e": {"value": a.steps / t_tot, "unit": unit, "h2d_bytes_per_step": int(rows.shape[1] * 8), "d2h_bytes_per_step": 8,
                    "note": "the state lives on the devices by definition (64 GiB at n=32); per step the host sends one angle row; "
                            "`value` and `e2e` are the same wall-clock measurement through DistributedState.step (N > 1) / "
                            "DistributedState.evolve_rows (N = 1)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_slice_rx_tma (TMA tile passes: rotations, phase and the previous step's owed rotations on "
                                                    "one tile; the pass that carries the exchange: k_slice_rx_tile) -- whole step, exchange included",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "traffic": None,
                         "alg_bytes_per_step_per_gpu": 2 * slice_bytes,
                         "passes_per_step": passes, "launches_per_step": launches / max(1, a.steps),
                         "per_pass_GBs": (launches / max(1, a.steps)) * ach,
                         "per_pass_note": "every launch of a step is one read + write of the slice: "
                                          "per-pass rate = launches per step x the per-step figure; an exchange lowers it",
                         "exchange_GB_per_step_per_gpu": exch_bytes / 1e9,
                         "exchange_GBs_per_gpu": (exch_bytes / (exch_ms * 1e-3) / 1e9) if exch_ms else None,
                         "nvlink_peak_GBs": 900.0,
                         "exchange_path": "peer-memory stores fused into the last local pass" if st.peer_exchange else
                                          ("nccl all_to_all_single" if R_.world > 1 else "none"),
                         "limiter": ("the exchange (%.1f GB per GPU per step over NVLink)" % (exch_bytes / 1e9)) if R_.world > 1 else
                                    "HBM passes (no exchange on one GPU)"},
            "parity": parity, "norm2_after": norm2, "clocks": sampler.window(*window),
            "cpu_baseline": {"value": None, "unit": unit, "cores": 0, "kind": "port",
                             "sample": "not runnable: a %d-qubit state is %.0f GiB; the reference's dense operators stop at n~13" % (
                                 n, 16.0 * 2 ** n / 2 ** 30)}}
    if R_.rank == 0:
        emit_json(line)
    R_.close()
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_b200_arm(a):
    import torch
    import diffquantum_b200 as dq
    from diffquantum_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its banner ("NCCL version ...") to file descriptor 1; stdout carries the one JSON line.  Keep a private
        # copy of the real stdout for that line and point fd 1 at stderr for everything else in this process.
        global _JSON_OUT
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        import torch.distributed as dist
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        import torch.distributed as dist
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    edges, coeff = workload(a.n)
    prob = dq.IsingProblem.maxcut(a.n, edges, omega0=omega_of(a), omega1=omega_of(a))
    sim = dq.IsingSimulator(prob, device=local, per_step=a.per_step, engine=a.engine)
    per_gpu = per_gpu_samples(a, world)
    if a.ket_group:
        sim.set_option("ket_group", a.ket_group)
    if a.item_tiles_log2 >= 0:
        sim.set_option("item_tiles_log2", a.item_tiles_log2)
    if sim.info("engine") != a.engine:
        raise RuntimeError("fused engine %d not available for n=%d" % (a.engine, a.n))
    kernel_name = "k_fused_ws (warp-specialised loop-form pass kernel: every |x angle| <= 1; k_fused_passes otherwise)"
    n_H = len(prob.terms)
    stream = torch.cuda.ExternalStream(sim.ctx.stream, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    cost = lambda s: steps_of_sample(s, prob.T, a.per_step, n_H)[2]
    est = sharding.ShardedEstimator(lambda c, s: sim.grad_samples(c, s), device=dev, cost=cost)

    def my_samples(i):
        return est.my_samples(step_samples(i, per_gpu, world, prob.T))

    # ---- value: tables staged in HBM, device-timed ----------------------------------------------
    def staged_loop(linear):
        """K timed steps of run_staged (W untimed before); returns device ms per step, counters, clocks."""
        sim.set_option("linear", 1 if linear else 0)
        sim.set_option("time_launches", 1)
        r = dict(dev_ms=[], alg_bytes=0.0, traj_steps=0.0, kern_ms=0.0, kern_launches=0.0, clocks=None, launches=0, en=None)
        l0 = 0
        t_start = time.time()
        for i in range(a.warmup + a.steps):
            timed = i >= a.warmup
            if i == a.warmup:
                barrier()
                t_start = time.time()
                l0 = sim.ctx.launch_count
            sim.stage(coeff, my_samples(i))              # H2D of the angle tables: outside the timed region
            flush.zero_()
            torch.cuda.synchronize(dev)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sim.run_staged()
            e1.record(stream)
            e1.synchronize()
            r["en"] = sim.fetch()
            if timed:
                r["dev_ms"].append(e0.elapsed_time(e1))
                r["alg_bytes"] += sim.stat("alg_bytes")
                r["traj_steps"] += sim.stat("steps")
                r["kern_ms"] += sim.stat("pass_kernel_ms")
                r["kern_launches"] += sim.stat("pass_kernel_launches")
        barrier()
        r["launches"] = sim.ctx.launch_count - l0
        r["window"] = (t_start, time.time())
        sim.set_option("time_launches", 0)
        sim.set_option("linear", 0)
        return r

    sampler = ClockSampler(local)            # starts (and finishes its start-up) during the warm-up steps
    lit = staged_loop(False)
    dev_ms, alg_bytes, traj_steps = lit["dev_ms"], lit["alg_bytes"], lit["traj_steps"]
    kern_ms, kern_launches, gpu_launches, en = lit["kern_ms"], lit["kern_launches"], lit["launches"], lit["en"]
    per_rank = None
    if world > 1:
        import torch.distributed as dist
        mine = torch.tensor([sum(dev_ms), kern_ms, traj_steps, float(gpu_launches)], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"rank": r, "device_ms": float(t[0]), "pass_kernel_ms": float(t[1]), "trajectory_steps": float(t[2]),
                     "launches": int(t[3])} for r, t in enumerate(allr)]
    t_value = max_over_ranks(sum(dev_ms) * 1e-3)
    total_samples = per_gpu * world * a.steps
    value = total_samples / t_value
    linear_leg = None
    if a.engine == 1:
        lin = staged_loop(True)
        t_lin = max_over_ranks(sum(lin["dev_ms"]) * 1e-3)
        dmax = float(np.abs(lin["en"] - en).max() / np.abs(en).max())
        linear_leg = {"value": total_samples / t_lin, "unit": UNIT, "ms_per_step": 1e3 * t_lin / a.steps,
                      "trajectory_steps_per_s": sum_over_ranks(lin["traj_steps"]) / t_lin,
                      "pass_kernel_GBs": lin["alg_bytes"] / (lin["kern_ms"] * 1e-3) / 1e9 if lin["kern_ms"] else None,
                      "max_rel_diff_of_shifted_energies_vs_literal": dmax,
                      "note": "NOT the headline: same per-sample outputs from n_Hs+1 suffix trajectories instead of 2*n_Hs "
                              "(ket- = 2 U phi/sqrt(1+r^2) - ket+ by linearity of the evolution); `value` above evolves "
                              "both shifted kets of every control literally as sim_plain.py:196-215 does"}
    sim.set_option("time_launches", 0)
    last_energies = en

    # ---- e2e: public API, host buffers, reduce included ----------------------------------------------
    h2d = d2h = 0
    e2e_s = []
    for i in range(a.warmup + a.steps):
        timed = i >= a.warmup
        s_all = step_samples(i, per_gpu, world, prob.T)
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        gmean = est.mean_gradient(coeff, s_all)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if timed:
            e2e_s.append(dt)
            tabs = sim.sample_tables(coeff, est.my_samples(s_all))                  # recount the bytes copied
            h2d = sum(t.nbytes for t in tabs) + prob.term_kind.nbytes + prob.term_index.nbytes
            d2h = len(est.my_samples(s_all)) * n_H * 2 * 8
    t_e2e = max_over_ranks(sum(e2e_s))
    e2e_value = total_samples / t_e2e
    assert np.all(np.isfinite(gmean)) and np.all(np.isfinite(last_energies))

    clk = sampler.window(*lit["window"])

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
    traffic = None
    tnote = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(tp) and kern_launches:
        try:
            tj = json.load(open(tp))
            ratio = float(tj["dram_bytes_per_launch"]) / float(tj["alg_bytes_per_launch"])
            traffic = ratio * alg_bytes / kern_launches
            tnote = "ncu dram bytes / algorithmic bytes = %.3f for %s (%s), applied to this run's bytes per launch" % (
                ratio, tj.get("kernel", "k_fused_passes"), tj.get("source", "profiles/"))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "peak_source": peak_src, "traffic": traffic,
                "traffic_note": tnote,
                "alg_bytes_per_launch": alg_bytes / kern_launches if kern_launches else None,
                "avg_launch_ms": kern_ms / kern_launches if kern_launches else None,
                "launches_timed": kern_launches, "kernel_share_of_step": kern_ms / sum(dev_ms) if dev_ms else None,
                "note": "kets of a launch group stay in the 126 MB L2 between passes by design, so algorithmic GB/s can "
                        "exceed DRAM GB/s; `traffic` is the ncu DRAM figure.  The kernel is bound inside the SM (FP64 pipe 46 %, "
                        "shared-memory wavefronts 53 %, their phases do not overlap: profiles/r02_o_fused_ws_ncu_summary.txt), not by HBM: "
                        "`frac` is reported against the HBM peak because that is the contract's denominator"}
    if achieved and a.config == 3 and (a.n in (None, 20)):
        # the limits that do apply at n = 20: per 64 KiB tile (128 KiB algorithmic) 3 545 cycles of FP64 pipe
        # (52 DFMA per amplitude and step at one warp-wide DFMA per 2.24 cycles and sub-partition) and 4 570 cycles of
        # shared-memory pipe (three register<->shared exchanges, the TMA load and store, 10 % bank conflicts), 148 SMs at the
        # SM clock sampled during the run
        mhz = (clk or {}).get("sm_mhz") or 1965.0
        per_tile = 148 * 131072.0 * mhz * 1e6 / 1e9
        roofline["sm_co_limits"] = {
            "fp64_pipe_GBs": per_tile / 3545.0, "shared_memory_pipe_GBs": per_tile / 4570.0,
            "frac_of_fp64_pipe": achieved / (per_tile / 3545.0), "frac_of_shared_memory_pipe": achieved / (per_tile / 4570.0),
            "source": "profiles/r02_o_fused_ws_ncu_summary.txt (instruction and wavefront counts per tile), "
                      "profiles/r02_al_fp64_lsu_overlap_microbench.txt (the two pipes overlap across warps: 88 % of the slower one)"}

    line = {
        "metric": metric_of(a)[0], "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t_value / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(a, world),
        "trajectory_steps_per_s": sum_over_ranks(traj_steps) / t_value,
        "alg_GBs_whole_job": sum_over_ranks(alg_bytes) / t_value / 1e9,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * t_e2e / a.steps,
                "note": "host pulse tables + H2D (cudaMemcpyAsync from the caller's numpy buffers) + kernels + D2H "
                        "energies + gradient assembly + all-reduce"},
        "gpu_launches": int(gpu_launches), "roofline": roofline, "clocks": clk,
    }
    if per_rank:
        line["per_rank"] = per_rank
    if world > 1:
        # untimed: per-sample gradients sharded over the N GPUs must be BIT-equal to the same samples on one GPU (rank 0)
        s_par = step_samples(12345, 2, world, prob.T)
        plain = sharding.ShardedEstimator(lambda c, s_: sim.grad_samples(c, s_), device=dev)
        every = plain.per_sample_gradients(coeff, s_par)
        if rank == 0:
            alone = sim.grad_samples(coeff, s_par)
            diff = float(np.abs(every - alone).max())
            line["parity_multi_gpu"] = {"max_abs_diff": diff, "bit_equal": bool(diff == 0.0), "samples": int(len(s_par)),
                                        "what": "ShardedEstimator.per_sample_gradients over %d GPUs vs the same samples on "
                                                "GPU 0 alone (n=%d, per_step=%d)" % (world, a.n, a.per_step)}
            if diff != 0.0:
                raise SystemExit("bench.py: %d-GPU per-sample gradients differ from the 1-GPU result: %r" % (world, diff))
    if linear_leg:
        line["linear_estimator"] = linear_leg
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            s0 = step_samples(a.warmup, per_gpu, 1, prob.T)[0]
            n_cpu = a.cpu_terms if a.config == 3 else 1                           # configs[2]: ~1300 steps per trajectory
            cpu_bounded_sample(a.n, a.per_step, coeff, edges, s0, 1, omega_of(a))  # warm the OpenMP pool / page in
            r = cpu_bounded_sample(a.n, a.per_step, coeff, edges, s0, n_cpu, omega_of(a))
            # parity at the headline size (untimed): the CPU port's shifted energies of this very sample against the
            # GPU's, through the public API -- same graph, coefficients, sampled time, per_step as the timed steps
            gpu_en = sim.shifted_energies(coeff, [s0])[0][r["order"]]
            err = float(np.abs(gpu_en - r["energies"]).max() / np.abs(r["energies"]).max())
            line["parity"] = {"max_rel_err": err, "tol": 1e-10, "ok": bool(err < 1e-10),
                              "what": "shifted energies <ket+-|M|ket+-> of %d controls (%d values) of the sample at s=%.4f, "
                                      "n=%d per_step=%d: CUDA (dq_ising_grad) vs the plain-C port of diffqc.cc:155-164 / "
                                      "sim_plain.py:186-230" % (r["n_terms"], 2 * r["n_terms"], s0, a.n, a.per_step)}
            if not line["parity"]["ok"]:
                raise SystemExit("bench.py: GPU energies differ from the CPU oracle: %r" % (line["parity"],))
            line["cpu_baseline"] = {
                "value": r["samples_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                "build": "gcc -O3 -march=native -fopenmp (oracle/c/Makefile), rebuilt on this host",
                "sample": "1 prefix + the +/- trajectories of %d of %d controls of one sample at s=%.4f (%d of %d "
                          "trajectory-steps, %.1f s), scaled by steps" % (r["n_terms"], r["n_H"], s0, r["steps"],
                                                                         r["full_steps"], r["seconds"])}
        except Exception as e:            # the baseline is a report, never a reason to lose the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %s" % e}
    if rank == 0:
        emit_json(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)
    if a.config in (0, 1):
        return run_dense_config(a)
    if a.config == 4:
        return run_distributed_state(a)
    return run_b200_arm(a)


if __name__ == "__main__":
    sys.exit(main())
