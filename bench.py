#!/usr/bin/env python
"""bench.py — gradient samples/s of the stochastic parameter-shift estimator at n=20 MaxCut
(BASELINE.json configs[3]: random 3-regular MaxCut, n=20, per_step=10 as shipped, T=2, B-spline
basis with 6 coefficients per control, 50 controls -> 1 + 100 trajectories per sample).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                          reference arm (CPU, host cores)

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


METRIC = "gradient samples/s at n=20 MaxCut"
UNIT = "samples/s"
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=20)
    ap.add_argument("--per-step", type=int, default=10)
    ap.add_argument("--samples-per-step", type=int, default=8, help="gradient samples per GPU per step")
    ap.add_argument("--cpu-terms", type=int, default=6,
                    help="controls whose +/- trajectories the bounded CPU sample runs (of n_Hs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ket-group", type=int, default=0)
    ap.add_argument("--item-tiles-log2", type=int, default=-1, help="fused v2: tiles per work item = 2^k (default: library default)")
    ap.add_argument("--engine", type=int, default=1, help="1 = the fused pass engine (the only product engine at n=20)")
    return ap.parse_args()


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
def cpu_bounded_sample(n, per_step, coeff, edges, s, n_terms):
    from oracle import c_port as C, restate as R
    if not C.available():
        raise RuntimeError("oracle/c/liboracle_c.so missing: run __graft_entry__.build()")
    C.use_host_cores()                      # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    prob = R.maxcut_structured(n, edges)
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
    edges, coeff = workload(a.n)
    T = 2.0
    vals, last = [], None
    for i in range(a.warmup + a.steps):
        # warm-up steps use a reduced sample (page-in, OpenMP pool start-up); timed ones the bounded sample
        s = step_samples(i, a.samples_per_step, 1, T)[0]
        r = cpu_bounded_sample(a.n, a.per_step, coeff, edges, s, 1 if i < a.warmup else a.cpu_terms)
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
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tot_sec / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, 1),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference's own dense code cannot build n=20 operators (SURVEY F3); this is the plain-C OpenMP port "
                "of its per-term product step (diffqc.cc:155-164) and estimator (sim_plain.py:186-230)",
    }
    print(json.dumps(line))
    return 0


def workload_config(a, world):
    return {"workload": "configs[3]: random 3-regular MaxCut n=%d (networkx seed 0), per_step=%d, T=2.0, "
                        "n_basis=6 BSpline, 1+2*n_Hs trajectories per sample" % (a.n, a.per_step),
            "samples_per_step_per_gpu": a.samples_per_step, "global_samples_per_step": a.samples_per_step * world,
            "parallelism": "sample-sharded x%d (cost-balanced shards), one all-reduce of the gradient per step" % world,
            "l2": "L2 flushed (512 MiB write) between timed steps"}


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
    prob = dq.IsingProblem.maxcut(a.n, edges)
    sim = dq.IsingSimulator(prob, device=local, per_step=a.per_step, engine=a.engine)
    if a.ket_group:
        sim.set_option("ket_group", a.ket_group)
    if a.item_tiles_log2 >= 0:
        sim.set_option("item_tiles_log2", a.item_tiles_log2)
    if sim.info("engine") != a.engine:
        raise RuntimeError("fused engine %d not available for n=%d" % (a.engine, a.n))
    kernel_name = "k_fused_passes"
    n_H = len(prob.terms)
    stream = torch.cuda.ExternalStream(sim.ctx.stream, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    cost = lambda s: steps_of_sample(s, prob.T, a.per_step, n_H)[2]
    est = sharding.ShardedEstimator(lambda c, s: sim.grad_samples(c, s), device=dev, cost=cost)

    def my_samples(i):
        return est.my_samples(step_samples(i, a.samples_per_step, world, prob.T))

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
    total_samples = a.samples_per_step * world * a.steps
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
        s_all = step_samples(i, a.samples_per_step, world, prob.T)
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
                        "exceed DRAM GB/s; `traffic` is the ncu DRAM figure"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t_value / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
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
            s0 = step_samples(a.warmup, a.samples_per_step, 1, prob.T)[0]
            cpu_bounded_sample(a.n, a.per_step, coeff, edges, s0, 1)            # warm the OpenMP pool / page in
            r = cpu_bounded_sample(a.n, a.per_step, coeff, edges, s0, a.cpu_terms)
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
    return run_b200_arm(a)


if __name__ == "__main__":
    sys.exit(main())
