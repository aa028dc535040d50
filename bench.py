#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native soft-robot-control hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Default workload (BASELINE.json configs[2], the one the target is quoted on): Trunk-SSM batched iLQR, 4096
independent solves per GPU, horizon 100, Gauss-Newton tracking of randomised figure-8 targets.  One "step" = one
batched solve of the whole batch (ONE launch of ilqr_solve_kernel).  Metric: iLQR solves/s (whole job).

  value      : device-resident inputs, CUDA events on the launching stream, L2 flushed between steps (untimed).
  e2e        : the public host API path -- pinned host buffers, H2D of x0 / targets, solve, D2H of x, u, K, cost.
  roofline   : algorithmic FP64 flops of the solve kernel / its event time against the measured cuBLAS DGEMM rate.
  cpu_baseline / --impl reference : the CPU port of the reference algorithm (oracle/, pinned bitwise to the
               reference classes) on the box's host cores, bounded sample.
Other workloads (--workload): ilqr_tpwl, tpwl_rollout_nn, tpwl_rollout_weighting, ssm_rollout, ssm_eval, pod_gram, mpc.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FP64_PEAK_TFLOPS = 35.4   # cuBLAS DGEMM 8192^3 measured on this pool's B200 (profiles/fp64_peaks_r01.json); the
#                           driver's MEASURED_PEAKS.json has no FP64 entry.  HBM peak comes from MEASURED_PEAKS.json.


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    hbm, src = 6650.0, "fallback"
    if os.path.exists(p):
        try:
            hbm, src = float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    f = os.path.join(REPO, "profiles", "fp64_peaks_r01.json")
    fp64 = FP64_PEAK_TFLOPS
    if os.path.exists(f):
        try:
            fp64 = float(json.load(open(f))["cublas_dgemm_tflops_sustained"])
        except Exception:
            pass
    return hbm, src, fp64


# ---------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(mhz)}


# ---------------------------------------------------------------------------------------------------------------
# algorithmic work (DESIGN.md "Algorithmic work per unit")
# ---------------------------------------------------------------------------------------------------------------
def ilqr_flops(n, m, nz, nfeat, N, fwd_passes, bwd_passes):
    """Algorithmic FP64 flops of the iLQR solve: dense counts of the matmuls the reference performs."""
    ssm_eval = 2 * nfeat * (n + n * n + nz + nz * n)              # f, A = r dphi, z, H = w dphi (dense contraction)
    be = 2 * (2 * n ** 3) + 2 * n ** 3 + 2 * n * n * m + 2 * n * n  # two inverses (~2n^3 each), sep, B_d, d_d
    fwd_step = ssm_eval + be + 2 * m * n + 2 * (n * n + n * m) + 2 * (nz * nz + m * m) + 2 * n * m
    bwd_step = (2 * (n * nz * nz + n * n * nz) + 2 * n * nz + 2 * m * m          # c_xx, c_x, c_u
                + 2 * (n * n + n * m)                                             # Q_x, Q_u
                + 2 * (n ** 3) * 2 + 2 * (m * n * n) * 2 + 2 * (m * m * n)        # A'P, (A'P)A, B'P, (B'P)A, (B'P)B
                + 2 * (m * n * n) * 2 + 2 * (m * m * n)                           # regularised B'(P+rho I), its two products
                + m ** 3 / 3 + 2 * m ** 3 + 2 * m * m * n + 2 * m * m             # Cholesky, inverse, K, k
                + 2 * (n * m * m) + 2 * 3 * n * m + 2 * 3 * n * n * m)            # K'Quu, p, P
    return fwd_passes * N * fwd_step + bwd_passes * N * bwd_step


# ---------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------
def build_ilqr(batch, N, seed):
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.SSM.ssm import SSMDynamics
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost
    w = synth.trunk_ilqr_batch(batch, N=N, seed=seed, m=8)
    s = w['ssm']
    model = SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    solver = iLQR(w['dt'], model, QuadraticCost(Q, R, Qf), N)
    return w, solver


def cpu_ilqr_worker(args):
    """One CPU worker: solves a slice of the same batch with the reference algorithm's CPU port."""
    os.environ["OMP_NUM_THREADS"] = "1"
    idx, N, seed, batch = args
    import sofacontrol_b200.synth as synth
    from oracle.ssm_np import SSMDynamicsNP, GaussNewtonSSM
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost
    w = synth.trunk_ilqr_batch(batch, N=N, seed=seed, m=8)
    s = w['ssm']
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    t0 = time.perf_counter()
    its = 0
    for b in idx:
        o = ILQRNP(w['dt'], GaussNewtonSSM(SSMDynamicsNP(s['z_ref'], discrete=False, discr_method='be', model=s['model'],
                                                         params=s['params'])), QuadraticCost(Q, R, Qf), N)
        o.set_target(w['z_target'][b])
        o.ilqr_computation(w['x0'][b])
        its += o.iterations
    return len(idx), its, time.perf_counter() - t0


def cpu_ilqr_baseline(batch, N, seed, per_core=2):
    """Reference CPU path on the host cores: `per_core` solves per core taken from the same batch, one process per
    core, OMP_NUM_THREADS=1.  Returns (solves/s aggregate, cores, sample description)."""
    import multiprocessing as mp
    cores = max(1, len(os.sched_getaffinity(0)))
    cores = min(cores, 64)
    jobs = [(list(range(c * per_core, (c + 1) * per_core)), N, seed, max(batch, cores * per_core)) for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(cpu_ilqr_worker, jobs)
    wall = time.perf_counter() - t0
    solved = sum(r[0] for r in res)
    busy = max(r[2] for r in res)
    return solved / busy, cores, "%d solves (%d per core, first problems of the same seeded batch), horizon %d; " \
                                 "wall %.1fs incl. process start, slowest worker %.1fs" % (solved, per_core, N, wall, busy)


def run_ilqr(args, rank, world, dev_index):
    import torch
    import torch.distributed as dist
    from sofacontrol_b200 import _lib as L
    batch, N = args.batch, args.horizon
    w, solver = build_ilqr(batch, N, seed=3 + rank)
    x0 = L.to_dev(w['x0'])
    zt = L.to_dev(w['z_target'])
    flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    for _ in range(args.warmup):
        out = solver.solve_device(x0, zt)
    torch.cuda.synchronize()
    iters = out['iterations'].cpu().numpy()
    trials = out['trials'].cpu().numpy()
    status = out['status'].cpu().numpy()

    # ---- device-resident timing (value)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    with ClockSampler(dev_index) as clk:
        for s, e in ev:
            flush.fill_(1.0)                       # L2 flush, untimed
            s.record()
            out = solver.solve_device(x0, zt)
            e.record()
        barrier()
    t_dev = sum(s.elapsed_time(e) for s, e in ev) * 1e-3
    clocks = clk.summary()

    # ---- end-to-end through the host API with pinned buffers (e2e)
    x0_h = torch.from_numpy(w['x0']).pin_memory()
    zt_h = torch.from_numpy(w['z_target']).pin_memory()
    outs_h = None
    for _ in range(2):
        outs_h = solver.solve_pinned(x0_h, zt_h, outs_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        outs_h = solver.solve_pinned(x0_h, zt_h, outs_h)
    barrier()
    t_e2e = time.perf_counter() - t0
    h2d = x0_h.numel() * 8 + zt_h.numel() * 8
    d2h = sum(v.numel() * v.element_size() for v in outs_h.values())

    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    total = batch * world * args.steps
    flops = ilqr_flops(6, 8, 6, 83, N, float((trials + 1).sum()), float(iters.sum()))
    hbm, hsrc, fp64 = measured_peaks()
    per_launch = t_dev / args.steps
    ach = flops / per_launch / 1e12
    res = {
        "metric": "ilqr_solves_per_sec", "value": total / t_dev, "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Trunk-SSM batched iLQR (BASELINE configs[2]): %d independent solves per GPU, horizon %d, "
                               "n=6 m=8 order-3 SSM (83 monomials), be discretisation, dt=0.02, Gauss-Newton figure-8 "
                               "tracking, randomised amplitude/phase/x0 (seed 3+rank)" % (batch, N),
                   "batch_per_gpu": batch, "horizon": N, "parallelism": "dp%d (problems sharded, no collective)" % world,
                   "l2": "256 MB buffer written between timed steps (untimed)",
                   "converged_frac": float((status & 1).mean()), "mean_iterations": float(iters.mean()),
                   "mean_forward_passes": float((trials + 1).mean())},
        "e2e": {"value": total / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": 2 * args.steps,         # ilqr_queue_init_kernel + ilqr_ssm_fast_kernel<8> per step
        "clocks": clocks,
        "roofline": {"kernel": "ilqr_ssm_fast_kernel<8>", "bound": "tensor", "achieved": ach, "peak": fp64,
                     "unit": "TFLOP/s", "frac": ach / fp64,
                     "traffic": 36.12e9 * batch / 4096.0 if (batch == 4096 and N == 100) else None,
                     "note": "FP64 pipe (DMMA + DFMA): algorithmic flops of the executed passes (dense counts, bench.py:"
                             "ilqr_flops, DESIGN.md) / event time of the single launch; peak = cuBLAS DGEMM 8192^3 "
                             "measured on this pool (profiles/fp64_peaks_r01.json), of measured; the kernel is "
                             "dependent-issue-latency bound (n = 6), see profiles/ncu_ilqr_v7_r01.txt; traffic = "
                             "dram read+write bytes of one launch from that ncu capture (trajectory records)"},
    }
    return res


def run_tpwl_rollout(args, rank, world, dev_index, method):
    import torch
    import torch.distributed as dist
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200 import _lib as L
    from sofacontrol_b200.tpwl.tpwl import TPWLATV
    batch, N = args.batch, args.horizon
    data, Hf = synth.tpwl_bank()
    params = {'tpwl_method': method, 'dist_weights': {'q': 1.0, 'v': 0.0}, 'beta_weighting': 25.0}
    g = TPWLATV(data, params=params, Hf=Hf, discr_method='fe' if method == 'weighting' else 'zoh')
    if method == 'nn':
        g.pre_discretize(0.01)
    x0h, uh = synth.tpwl_rollout_batch(batch, N=N, seed=2 + rank)
    x0, u = L.to_dev(x0h), L.to_dev(uh)
    flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)
    for _ in range(args.warmup):
        g.rollout_device(x0, u, 0.01)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(dev_index) as clk:
        for s, e in ev:
            flush.fill_(1.0)
            s.record()
            x, z = g.rollout_device(x0, u, 0.01)
            e.record()
        torch.cuda.synchronize()
    t_dev = sum(s.elapsed_time(e) for s, e in ev) * 1e-3
    # e2e: host API
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.rollout(x0h, uh, 0.01)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    steps_total = batch * N * world * args.steps
    hbm, hsrc, fp64 = measured_peaks()
    n, m, P, r = 72, 4, 1000, 36
    if method == 'nn':
        byt = batch * N * ((n * n + n * m + n) * 8 + (2 * n + m) * 8) + N * P * r * 8
        ops = batch * N * 2.0 * P * r
        roof = {"kernel": "tpwl_rollout_nn_screen_kernel<36,4>", "bound": "hbm", "achieved": byt / (t_dev / args.steps) / 1e9,
                "peak": hbm, "unit": "GB/s", "traffic": None,
                "fp32_screen_tops": ops / (t_dev / args.steps) / 1e12,
                "note": "algorithmic bytes per trajectory-step = gathered bank entry 44352 B + state I/O, distance bank "
                        "288000 B once per time step for the whole batch (SURVEY 8d), of " + hsrc + "; the 44 MB bank is "
                        "L2 resident, so these bytes move L2 -> SM, not HBM -> L2 (ncu: profiles/ncu_tpwl_screen_r01.txt). "
                        "The kernel alternates an FP32-issue-bound exact two-stage nearest search (fp32_screen_tops = "
                        "2 P r FP32 instructions per trajectory-step, T lane-ops/s) with the L2-bandwidth-bound gather"}
    else:
        fl = batch * N * 2.0 * P * (n * n + n * m + n)
        roof = {"kernel": "dgemm_kernel (bank blend)", "bound": "tensor", "achieved": fl / (t_dev / args.steps) / 1e12,
                "peak": fp64, "unit": "TFLOP/s", "traffic": None,
                "note": "2 P (n^2+nm+n) flop per trajectory-step; whole step time (weights + blend + discretise + step)"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    return {"metric": "tpwl_%s_rollout_steps_per_sec" % method, "value": steps_total / t_dev, "unit": "steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Diamond TPWL batched rollout (BASELINE configs[1]): %d trajectories x %d steps per GPU, "
                                   "n=72 m=4 P=1000 r=36, method %s (%s)" % (batch, N, method, "pre-discretised zoh bank" if method == "nn" else "fe per step"), "batch_per_gpu": batch,
                       "horizon": N, "l2": "256 MB buffer written between timed steps (untimed)"},
            "e2e": {"value": steps_total / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": int((x0h.size + uh.size) * 8),
                    "d2h_bytes_per_step": int(batch * (N + 1) * (n + 6) * 8)},
            "gpu_launches": args.steps * (1 if method == 'nn' else N * 7) + args.steps, "clocks": clk.summary(),
            "roofline": roof}


def run_ssm_rollout(args, rank, world, dev_index):
    import torch
    import torch.distributed as dist
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200 import _lib as L
    from sofacontrol_b200.SSM.ssm import SSMDynamics
    batch, N = args.batch, args.horizon
    s = synth.trunk_ssm(8)
    g = SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    rng = np.random.default_rng(1 + rank)
    x0h = np.zeros((batch, 6)); x0h[:, :3] = rng.uniform(-0.5, 0.5, size=(batch, 3))
    uh = rng.uniform(0, 800, size=(batch, N, 8))
    x0, u = L.to_dev(x0h), L.to_dev(uh)
    flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)
    for _ in range(args.warmup):
        g.rollout_device(x0, u, 0.02)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(dev_index) as clk:
        for s_, e_ in ev:
            flush.fill_(1.0)
            s_.record()
            g.rollout_device(x0, u, 0.02)
            e_.record()
        torch.cuda.synchronize()
    t_dev = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.rollout(x0h, uh, 0.02)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    total = batch * N * world * args.steps
    hbm, hsrc, fp64 = measured_peaks()
    fl = batch * N * (2.0 * 83 * (6 + 36 + 6) + 2 * (2 * 216) + 2 * 216 + 2 * 36 * 8 + 2 * (36 + 48) * 2)
    ach = fl / (t_dev / args.steps) / 1e12
    return {"metric": "ssm_rollout_steps_per_sec", "value": total / t_dev, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Trunk SSM batched open-loop rollout (BASELINE configs[0] batched): %d trajectories x %d steps per "
                                   "GPU, n=6 m=8 order 3, be discretisation, dt=0.02, u~U(0,800)" % (batch, N),
                       "batch_per_gpu": batch, "horizon": N, "l2": "256 MB buffer written between timed steps (untimed)"},
            "e2e": {"value": total / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": int((x0h.size + uh.size) * 8),
                    "d2h_bytes_per_step": int(batch * (N + 1) * 12 * 8)},
            "gpu_launches": args.steps, "clocks": clk.summary(),
            "roofline": {"kernel": "ssm_rollout_fast_kernel<8>", "bound": "tensor", "achieved": ach, "peak": fp64,
                         "unit": "TFLOP/s", "frac": ach / fp64, "traffic": None,
                         "note": "dense algorithmic flops (model contraction + discretisation + step) / event time, of measured cuBLAS DGEMM"}}


def run_ssm_eval(args, rank, world, dev_index):
    """Kernel (b): batched evaluation + linearisation, continuous Jacobians (A, d, H, c, z) of `count` states."""
    import torch
    import torch.distributed as dist
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200 import _lib as L
    from sofacontrol_b200.SSM.ssm import SSMDynamics
    count = args.batch * 1024
    s = synth.trunk_ssm(8)
    g = SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    gen = torch.Generator(device="cuda").manual_seed(1 + rank)
    x = torch.randn((count, 6), device="cuda", dtype=torch.float64, generator=gen)
    u = torch.rand((count, 8), device="cuda", dtype=torch.float64, generator=gen) * 800
    want = ('A', 'd', 'H', 'c', 'z')
    for _ in range(args.warmup):
        out = g._eval_device(x, u, -1.0, 'cont_raw', want)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(dev_index) as clk:
        for s_, e_ in ev:
            s_.record()
            out = g._eval_device(x, u, -1.0, 'cont_raw', want)
            e_.record()
        torch.cuda.synchronize()
    t_dev = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    if world > 1:
        tt = torch.tensor([t_dev], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt[0])
    hbm, hsrc, fp64 = measured_peaks()
    ach = count * 13944.0 / (t_dev / args.steps) / 1e12
    byts = count * (14 + 36 + 6 + 36 + 6 + 6) * 8
    return {"metric": "ssm_eval_linearize_states_per_sec", "value": count * world * args.steps / t_dev, "unit": "states/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Trunk SSM batched evaluation + linearisation: %d states per GPU, outputs A_c, d_c, H, c, z "
                                   "(inputs %.0f MB + outputs %.0f MB per launch: larger than L2)" % (count, count * 14 * 8 / 1e6, byts / 1e6 - count * 14 * 8 / 1e6)},
            "e2e": None, "gpu_launches": args.steps, "clocks": clk.summary(),
            "roofline": {"kernel": "ssm_eval_dmma_kernel<8>", "bound": "tensor", "achieved": ach, "peak": fp64, "unit": "TFLOP/s",
                         "frac": ach / fp64, "traffic": None, "hbm_gbs": byts / (t_dev / args.steps) / 1e9,
                         "note": "13944 algorithmic flop per state (dense count, SURVEY 8d); 42 DMMA m8n8k4 per state execute "
                                 "21504 flop (padding 16x84x8); peak = measured cuBLAS DGEMM"}}


def run_pod_gram(args, rank, world, dev_index):
    """Kernel (d): POD Gram X^T X with the rows (DOFs) of X sharded across ranks + ONE NCCL all-reduce of G."""
    import torch
    import torch.distributed as dist
    from sofacontrol_b200.mor import pod
    from sofacontrol_b200 import parallel
    nf_local, ns = 131072, 8192
    gen = torch.Generator(device="cuda").manual_seed(5 + rank)
    X = torch.randn((nf_local, ns), device="cuda", dtype=torch.float64, generator=gen)
    G = torch.empty((ns, ns), device="cuda", dtype=torch.float64)
    for _ in range(max(1, args.warmup - 1)):
        pod.gram_device(X, G)
        parallel.allreduce_sum_(G)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(dev_index) as clk:
        for s_, m_, e_ in ev:
            s_.record()
            pod.gram_device(X, G)
            m_.record()
            parallel.allreduce_sum_(G)
            e_.record()
        torch.cuda.synchronize()
    t_gram = sum(a.elapsed_time(b) for a, b, _ in ev) * 1e-3
    t_all = sum(a.elapsed_time(c) for a, _, c in ev) * 1e-3
    if world > 1:
        tt = torch.tensor([t_gram, t_all], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_gram, t_all = float(tt[0]), float(tt[1])
    hbm, hsrc, fp64 = measured_peaks()
    fl = 2.0 * nf_local * ns * ns
    ach = fl / (t_gram / args.steps) / 1e12
    return {"metric": "pod_gram_tflops", "value": world * fl * args.steps / t_all / 1e12, "unit": "TFLOP/s (algorithmic, incl. all-reduce)",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "POD snapshot Gram (BASELINE configs[4] per-GPU slice, reduced): X_g %d x %d FP64 per GPU (%.1f GB), "
                                   "G = sum_g X_g^T X_g, one all-reduce of %d x %d" % (nf_local, ns, nf_local * ns * 8 / 1e9, ns, ns),
                       "allreduce_ms": 1e3 * (t_all - t_gram) / args.steps},
            "e2e": None, "gpu_launches": args.steps, "clocks": clk.summary(),
            "roofline": {"kernel": "dgemm_kernel<true,true,true> (SYRK)", "bound": "tensor", "achieved": ach, "peak": fp64,
                         "unit": "TFLOP/s", "frac": ach / fp64, "traffic": None,
                         "note": "algorithmic 2 nf ns^2 flop; the SYRK kernel executes only the upper tiles (half), so frac can exceed 1; "
                                 "peak = measured cuBLAS DGEMM"}}


def run_mpc(args, rank, world, dev_index):
    """BASELINE configs[3]: closed-loop receding-horizon Monte Carlo -- every control step re-solves all problems
    (horizon 20, warm start = shifted previous plan, u_last) and steps the plant with process noise."""
    import torch
    import torch.distributed as dist
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200 import _lib as L
    from sofacontrol_b200.SSM.ssm import SSMDynamics
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.mpc import RecedingHorizonILQR
    from sofacontrol_b200.utils import QuadraticCost
    batch = args.batch if args.batch != 4096 else 16384
    N, steps = 20, 20
    s = synth.trunk_ssm(8)
    model = SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    solver = iLQR(0.02, model, QuadraticCost(Q, R, Qf), N)
    rng = np.random.default_rng(4 + rank)
    amp, ph = rng.uniform(2, 15, size=batch), rng.uniform(0, 2 * np.pi, size=batch)
    T = steps + N
    th = np.linspace(0, 2 * np.pi * T / 100.0, T + 1)[None, :] + ph[:, None]
    zref = np.tile(s['z_ref'], (batch, T + 1, 1))
    zref[:, :, 0] += -amp[:, None] * np.sin(th); zref[:, :, 1] += amp[:, None] * np.sin(2 * th)
    x0 = np.zeros((batch, 6)); x0[:, :3] = rng.uniform(-0.5, 0.5, size=(batch, 3))
    mpc = RecedingHorizonILQR(solver, process_noise_std=1e-3, seed=4 + rank)
    x0d, zd = L.to_dev(x0), L.to_dev(zref)
    mpc.run_device(x0d, zd, 2)                                   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    with ClockSampler(dev_index) as clk:
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        out = mpc.run_device(x0d, zd, steps)
        e_.record()
        torch.cuda.synchronize()
    t_dev = s_.elapsed_time(e_) * 1e-3
    if world > 1:
        tt = torch.tensor([t_dev], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt[0])
    its = out['iterations'].float().mean().item()
    return {"metric": "mpc_problem_steps_per_sec", "value": batch * steps * world / t_dev, "unit": "receding-horizon solves/s",
            "n_gpus": world, "steps": steps, "warmup": 2, "ms_per_step": 1e3 * t_dev / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "closed-loop receding-horizon iLQR Monte Carlo (BASELINE configs[3], SSM plant): %d problems per "
                                   "GPU x %d control steps, horizon %d, warm start + u_last, process noise 1e-3" % (batch, steps, N),
                       "mean_iterations_per_solve": its, "converged_frac": float((out['status'] & 1).float().mean().item())},
            "e2e": None, "gpu_launches": steps * 2, "clocks": clk.summary(), "roofline": None}


def run_ilqr_tpwl(args, rank, world, dev_index):
    """Diamond TPWL iLQR (the reference's TPWL controller, tpwl/controllers.py + lqr/ilqr.py): n = 72, m = 4, P = 1000,
    nearest-neighbour linearisation on the zoh pre-discretised bank, horizon 100, figure-8 targets of random
    amplitude.  One CTA per problem (generic kernel, Diamond instantiation)."""
    import torch
    import torch.distributed as dist
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200 import _lib as L
    from sofacontrol_b200.tpwl.tpwl import TPWLATV
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost
    batch = args.batch if args.batch != 4096 else 1184
    N = args.horizon
    data, Hf = synth.tpwl_bank()
    g = TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='zoh')
    g.pre_discretize(0.01)
    Q = np.zeros((6, 6)); Q[3, 3] = Q[4, 4] = 100.0
    R = 1e-5 * np.eye(4)
    th = np.linspace(0, 2 * np.pi, N + 1)
    rng = np.random.default_rng(rank)
    x0, _ = synth.tpwl_rollout_batch(batch, N=1, seed=21 + rank)
    amp = rng.uniform(0.3, 1.5, size=batch)
    zt = np.tile(g.z_ref, (batch, N + 1, 1))
    zt[:, :, 3] += amp[:, None] * np.sin(th)[None]; zt[:, :, 4] += amp[:, None] * np.sin(2 * th)[None]
    solver = iLQR(0.01, g, QuadraticCost(Q, R, np.zeros((6, 6))), N)
    x0d, ztd = L.to_dev(x0), L.to_dev(zt)
    flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)
    out = None
    for _ in range(max(1, min(args.warmup, 3))):
        out = solver.solve_device(x0d, ztd)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(dev_index) as clk:
        for s_, e_ in ev:
            flush.fill_(1.0)
            s_.record()
            out = solver.solve_device(x0d, ztd)
            e_.record()
        torch.cuda.synchronize()
    t_dev = sum(s_.elapsed_time(e_) for s_, e_ in ev) * 1e-3
    if world > 1:
        tt = torch.tensor([t_dev], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt[0])
    it = out['iterations'].cpu().numpy(); tr = out['trials'].cpu().numpy(); st = out['status'].cpu().numpy()
    # DMMA work of the backward sweeps: (76x73x72 + 8x76x72 + 72x73x84) fused multiply-adds per step
    flops = 2.0 * (76 * 73 * 72 + 8 * 76 * 72 + 72 * 73 * 84) * N * float(it.sum())
    hbm, hsrc, fp64 = measured_peaks()
    ach = flops * args.steps / t_dev / 1e12
    return {"metric": "ilqr_solves_per_sec", "value": batch * world * args.steps / t_dev, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Diamond TPWL batched iLQR: %d independent solves per GPU, horizon %d, n=72 m=4 P=1000, nearest-"
                                   "neighbour linearisation on the zoh bank (dt 0.01), figure-8 tracking of random amplitude"
                                   % (batch, N),
                       "batch_per_gpu": batch, "horizon": N, "l2": "256 MB buffer written between timed steps (untimed)",
                       "converged_frac": float((st & 1).mean()), "mean_iterations": float(it.mean()),
                       "mean_forward_passes": float((tr + 1).mean())},
            "e2e": None, "gpu_launches": args.steps, "clocks": clk.summary(),
            "roofline": {"kernel": "ilqr_solve_kernel<TpwlPolicyT<72,4,6>>", "bound": "tensor", "achieved": ach, "peak": fp64,
                         "unit": "TFLOP/s", "frac": ach / fp64, "traffic": None,
                         "note": "flops of the three DMMA products of every backward step only (forward passes and the "
                                 "nearest-point search not counted) / whole-solve time; peak = measured cuBLAS DGEMM"}}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU port on the host cores, same workload/metric (bounded sample)."""
    t_all = []
    val = cores = sample = None
    for i in range(max(1, min(args.steps, 2))):
        val, cores, sample = cpu_ilqr_baseline(args.batch, args.horizon, seed=3, per_core=args.cpu_per_core)
        t_all.append(val)
    val = max(t_all)
    return {"impl": "reference", "metric": "ilqr_solves_per_sec", "value": val, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Trunk-SSM batched iLQR (BASELINE configs[2]), horizon %d -- reference algorithm on CPU" % args.horizon},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ilqr_trunk_ssm",
                    choices=["ilqr_trunk_ssm", "tpwl_rollout_nn", "tpwl_rollout_weighting", "ssm_rollout", "ssm_eval", "pod_gram", "mpc", "ilqr_tpwl"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--cpu-per-core", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.warmup = max(args.warmup, 3)
    if args.workload == "ilqr_trunk_ssm":
        res = run_ilqr(args, rank, world, local)
    elif args.workload == "ssm_rollout":
        res = run_ssm_rollout(args, rank, world, local)
    elif args.workload == "ssm_eval":
        res = run_ssm_eval(args, rank, world, local)
    elif args.workload == "pod_gram":
        res = run_pod_gram(args, rank, world, local)
    elif args.workload == "ilqr_tpwl":
        res = run_ilqr_tpwl(args, rank, world, local)
    elif args.workload == "mpc":
        res = run_mpc(args, rank, world, local)
    else:
        res = run_tpwl_rollout(args, rank, world, local, "nn" if args.workload.endswith("nn") else "weighting")
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "ilqr_trunk_ssm":
        v, cores, sample = cpu_ilqr_baseline(args.batch, args.horizon, seed=3, per_core=args.cpu_per_core)
        res["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample}
    elif rank == 0:
        res.setdefault("cpu_baseline", None)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
