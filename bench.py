#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native soft-robot-control hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Default workload (BASELINE.json configs[2], the one the target is quoted on): Trunk-SSM batched iLQR, 4096
independent solves per GPU, horizon 100, Gauss-Newton tracking of randomised figure-8 targets.  One "step" = one
batched solve of the whole batch (ONE launch of ilqr_solve_kernel).  Metric: iLQR solves/s (whole job).

  value      : device-resident inputs, CUDA events on the launching stream, L2 flushed between steps (untimed).
  e2e        : the public host API path -- pinned host buffers, H2D of x0 / targets, solve, D2H of x, u, K, cost, every
               step; double-buffered (the D2H of step i overlaps the solve of step i + 1, iLQR.solve_pinned_stream).
  roofline   : algorithmic FP64 flops of the solve kernel / its event time against the measured cuBLAS DGEMM rate.
  strong     : (N > 1) the SAME 4096 problems split across the N ranks (BASELINE configs[2] "sharded across 1/2/4/8
               GPUs"), measured in the same run next to the weak-scaling `value` (4096 per GPU, same seed per rank).
  secondary  : the other kernels of the path (TPWL nn rollout, SSM evaluation, POD Gram [+ all-reduce for N > 1],
               closed-loop MPC) measured outside the headline's timed region, each with roofline / e2e / cpu_baseline.
  cpu_baseline / --impl reference : the reference's own iLQR class (imported from /root/reference when that tree is
               present: kind "reference") or else its CPU port (oracle/, pinned bitwise to the reference classes:
               kind "port") on the box's host cores, bounded sample.
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


def _reference_available():
    try:
        from oracle import refimport
        return refimport.available()
    except Exception:
        return False


def cpu_ilqr_worker(args):
    """One CPU worker: solves a slice of the same batch with the reference's own iLQR class (when /root/reference is
    present; it drives the SSM through the Gauss-Newton H-property adapter of SURVEY App. C.2) or its CPU port."""
    os.environ["OMP_NUM_THREADS"] = "1"
    idx, N, seed, batch, use_ref = args
    import contextlib
    import io
    import warnings
    import sofacontrol_b200.synth as synth
    from oracle.ssm_np import SSMDynamicsNP, GaussNewtonSSM
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost
    w = synth.trunk_ilqr_batch(batch, N=N, seed=seed, m=8)
    s = w['ssm']
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    cls, qc = ILQRNP, QuadraticCost
    if use_ref:
        warnings.simplefilter("ignore")
        from oracle import refimport
        ref = refimport.load()
        cls, qc = ref.ilqr.iLQR, ref.utils.QuadraticCost
    t0 = time.perf_counter()
    its = 0
    for b in idx:
        o = cls(w['dt'], GaussNewtonSSM(SSMDynamicsNP(s['z_ref'], discrete=False, discr_method='be', model=s['model'],
                                                      params=s['params'])), qc(Q, R, Qf), N)
        o.set_target(w['z_target'][b])
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):     # the reference prints in its hot loop
            o.ilqr_computation(w['x0'][b])
        its += getattr(o, 'iterations', 0)
    return len(idx), its, time.perf_counter() - t0


def cpu_ilqr_baseline(batch, N, seed, per_core=2):
    """Reference CPU path on the host cores: `per_core` solves per core taken from the same batch, one process per
    core, OMP_NUM_THREADS=1.  Returns (solves/s aggregate, cores, sample description, kind,
    seconds of the slowest worker)."""
    import multiprocessing as mp
    cores = max(1, len(os.sched_getaffinity(0)))
    cores = min(cores, 64)
    use_ref = _reference_available()
    jobs = [(list(range(c * per_core, (c + 1) * per_core)), N, seed, max(batch, cores * per_core), use_ref) for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(cpu_ilqr_worker, jobs)
    wall = time.perf_counter() - t0
    solved = sum(r[0] for r in res)
    busy = max(r[2] for r in res)
    kind = "reference" if use_ref else "port"
    what = ("the reference's iLQR class (sofacontrol/lqr/ilqr.py, unmodified) on the pinned numpy SSM" if use_ref else
            "the CPU port of the reference iLQR (oracle/ilqr_np.py, pinned bitwise to the reference class)")
    return solved / busy, cores, "%s: %d solves (%d per core, first problems of the same seeded batch), horizon %d; " \
                                 "wall %.1fs incl. process start, slowest worker %.1fs" % (what, solved, per_core, N, wall, busy), kind, busy


def ilqr_record_traffic(n, m, nz, N, fwd_passes, bwd_passes, batch):
    """Bytes the solve kernel moves through L2/HBM by design: every forward step writes one
    trajectory record (x, u, e, H_t, A_t, B_t) and reads the nominal step (x, u, K_t, k_t, target); every backward
    step reads the record (+ u_{t-1}) and writes K_t, k_t and two line-search scalars; plus the result copy."""
    rec = n + m + nz + nz * n + n * n + n * m
    fwd = (rec + (n + m + m * n + m + nz)) * 8.0
    bwd = (rec + m + (m * n + m + 2)) * 8.0
    out = batch * ((N + 1) * n + N * m) * 8.0 * 2
    return fwd_passes * N * fwd + bwd_passes * N * bwd + out


def run_ilqr(args, rank, world, dev_index):
    import torch
    import torch.distributed as dist
    from sofacontrol_b200 import _lib as L
    from sofacontrol_b200.parallel import shard_slice
    batch, N = args.batch, args.horizon
    # the same seeded batch on every rank: weak scaling measures the system, not a different draw of problems
    w, solver = build_ilqr(batch, N, seed=3)
    x0 = L.to_dev(w['x0'])
    zt = L.to_dev(w['z_target'])
    flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(x0_, zt_, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        o = None
        for s_, e_ in ev:
            flush.fill_(1.0)                       # L2 flush, untimed
            s_.record()
            o = solver.solve_device(x0_, zt_)
            e_.record()
        barrier()
        return sum(s_.elapsed_time(e_) for s_, e_ in ev) * 1e-3, o

    out = None
    for _ in range(args.warmup):
        out = solver.solve_device(x0, zt)
    torch.cuda.synchronize()
    iters = out['iterations'].cpu().numpy()
    trials = out['trials'].cpu().numpy()
    status = out['status'].cpu().numpy()

    # ---- device-resident timing (value): weak scaling, `batch` problems per GPU
    with ClockSampler(dev_index) as clk:
        t_dev, out = timed(x0, zt, args.steps)
    clocks = clk.summary()

    # ---- strong scaling: the SAME `batch` problems split across the ranks (BASELINE configs[2])
    strong = None
    if world > 1:
        sl = shard_slice(batch, rank, world)
        xs, zs = x0[sl].contiguous(), zt[sl].contiguous()
        for _ in range(2):
            solver.solve_device(xs, zs)
        t_strong, _ = timed(xs, zs, args.steps)

    # ---- end-to-end through the host API with pinned buffers (e2e)
    x0_h = torch.from_numpy(w['x0']).pin_memory()
    zt_h = torch.from_numpy(w['z_target']).pin_memory()
    outs_h = None
    for _ in range(2):
        outs_h = solver.solve_pinned(x0_h, zt_h, outs_h)
    for outs_h in solver.solve_pinned_stream([(x0_h, zt_h)] * 3):
        pass
    barrier()
    t0 = time.perf_counter()
    # every step: H2D of its inputs from pinned memory, the solve, D2H of its results into pinned memory; the D2H of step i
    # overlaps the solve of step i + 1 (iLQR.solve_pinned_stream), the last one is drained inside the timed region
    n_out = 0
    for outs_h in solver.solve_pinned_stream([(x0_h, zt_h)] * args.steps):
        n_out += 1
    assert n_out == args.steps
    barrier()
    t_e2e = time.perf_counter() - t0
    h2d = x0_h.numel() * 8 + zt_h.numel() * 8
    d2h = sum(v.numel() * v.element_size() for v in outs_h.values())

    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, t_strong], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_strong = float(tt[0]), float(tt[1]), float(tt[2])
        strong = {"value": batch * args.steps / t_strong, "unit": "solves/s", "ms_per_step": 1e3 * t_strong / args.steps,
                  "batch_total": batch, "batch_per_gpu": batch // world,
                  "note": "the same %d seed-3 problems split contiguously across the %d ranks, no collective; "
                          "%d problems per GPU is below the %d resident warps of one B200, so this is the latency of "
                          "the longest solves, not throughput" % (batch, world, batch // world, 148 * 16)}
    total = batch * world * args.steps
    fwd_passes, bwd_passes = float((trials + 1).sum()), float(iters.sum())
    flops = ilqr_flops(6, 8, 6, 83, N, fwd_passes, bwd_passes)
    hbm, hsrc, fp64 = measured_peaks()
    per_launch = t_dev / args.steps
    ach = flops / per_launch / 1e12
    traffic_model = ilqr_record_traffic(6, 8, 6, N, fwd_passes, bwd_passes, batch)
    res = {
        "metric": "ilqr_solves_per_sec", "value": total / t_dev, "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Trunk-SSM batched iLQR (BASELINE configs[2]): %d independent solves per GPU, horizon %d, "
                               "n=6 m=8 order-3 SSM (83 monomials), be discretisation, dt=0.02, Gauss-Newton figure-8 "
                               "tracking, randomised amplitude/phase/x0 (seed 3, the same batch on every rank)" % (batch, N),
                   "batch_per_gpu": batch, "horizon": N, "parallelism": "dp%d (problems sharded, no collective)" % world,
                   "l2": "256 MB buffer written between timed steps (untimed)",
                   "converged_frac": float((status & 1).mean()), "mean_iterations": float(iters.mean()),
                   "mean_forward_passes": float((trials + 1).mean()),
                   "total_iterations_per_rank": int(iters.sum())},
        "e2e": {"value": total / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": 3 * args.steps,         # queue_init_kernel + ilqr_ssm_fast_kernel<8,0> (until 1776 problems are left) + <8,2> (tail) per step
        "clocks": clocks,
        "roofline": {"kernel": "ilqr_ssm_fast_kernel<8,0> + <8,2> (tail hand-over)", "bound": "tensor", "achieved": ach, "peak": fp64,
                     "unit": "TFLOP/s", "frac": ach / fp64,
                     "traffic": traffic_model,
                     "traffic_source": "modelled from the executed passes (bench.py:ilqr_record_traffic: record + gain "
                                       "bytes per pass-step); ncu dram read+write of the two launches of one step of this workload: "
                                       "31.6e9 + 4.0e9 (profiles/ncu_ilqr_r2_final.txt)",
                     "algorithmic_io_bytes": float(batch * (55e3)),
                     "note": "FP64 pipe (DMMA + DFMA): algorithmic flops of the executed passes (dense counts, bench.py:"
                             "ilqr_flops, DESIGN.md) / event time of the single launch; peak = cuBLAS DGEMM 8192^3 "
                             "measured on this pool (profiles/fp64_peaks_r01.json), of measured; the kernel is bound by the "
                             "dependent-issue latency of one warp (n = 6: issue slots 32 %, L1/shared 58-72 %, FP64 pipe "
                             "16 %) and, at 4096 problems, by the serial chain of the longest solves (22 % of the warp-time "
                             "idle in the task queue): profiles/ncu_ilqr_r2_l2.txt, ncu_ilqr_r2_l2_lines.txt, launches_bench_r2.csv, DESIGN.md 4.1"},
    }
    if strong is not None:
        res["strong"] = strong
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
    # e2e: pinned host tensors in (x0, u), pinned host tensors out (x, z), copies inside the timed region
    x0p, up_ = torch.from_numpy(x0h).pin_memory(), torch.from_numpy(uh).pin_memory()
    g.rollout_pinned(x0p, up_, 0.01)                     # allocates the pinned outputs (untimed)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.rollout_pinned(x0p, up_, 0.01)
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
        ops = batch * N * 2.0 * P * r              # one FMA per point-coordinate (dot-product screen)
        dram = 0.25e9 * batch / 4096.0          # ncu dram read+write of one launch (profiles/ncu_tpwl_screen_r01.txt)
        l2_peak = 17978.0                       # GB/s, tools/l2_bw_microbench.cu gather pattern (profiles/l2_bw_microbench_r2.txt)
        roof = {"kernel": "tpwl_rollout_nn_screen_kernel<36,4>", "bound": "l2/lsu (on-chip: the 44 MB bank is L2 resident)",
                "achieved": byt / (t_dev / args.steps) / 1e9, "peak": l2_peak, "unit": "GB/s (L2 -> SM, algorithmic)",
                "traffic": dram, "hbm_frac": dram / (t_dev / args.steps) / 1e9 / hbm,
                "l1tex_pct_ncu": 74.2, "lts_pct_ncu": 24.7, "issue_active_pct_ncu": 43.8,
                "fp32_screen_tops": ops / (t_dev / args.steps) / 1e12,
                "note": "algorithmic bytes per trajectory-step = gathered bank entry 44352 B + state I/O, distance bank "
                        "288000 B once per time step for the whole batch (SURVEY 8d).  These bytes move L2 -> SM, not HBM -> "
                        "L2: DRAM traffic is 0.25 GB per launch (hbm_frac, of " + hsrc + "); ncu (profiles/ncu_tpwl_screen_r2.txt): "
                        "L1 74 %, issue slots 43.8 % at 4 warps per scheduler -- latency-bound at this occupancy "
                        "(profiles/phase_clocks_tpwl_nn_r2.txt).  peak = L2 -> SM bandwidth measured on this pool for the same "
                        "access pattern (random 44 KB entries, 16-byte loads, all SMs: 18.0 TB/s; coalesced sweep 20.5 TB/s; "
                        "tools/l2_bw_microbench.cu), of measured; frac = achieved / peak."}
        roof["frac"] = roof["achieved"] / l2_peak
    else:
        fl = batch * N * 2.0 * P * (n * n + n * m + n)
        roof = {"kernel": "dgemm_kernel (bank blend)", "bound": "tensor", "achieved": fl / (t_dev / args.steps) / 1e12,
                "peak": fp64, "unit": "TFLOP/s", "traffic": None,
                "note": "2 P (n^2+nm+n) flop per trajectory-step over the WHOLE step time; two launches per time step (profiles/launches_tpwl_weighting_r2.csv): blend GEMM over the concatenated [A|B|d] bank 1.555 ms = 29.2 TFLOP/s (82 % of measured DGEMM peak), fused TMA-fed discretise + step + next-weights kernel 0.287 ms (profiles/launches_tpwl_weighting_r2_mbar.csv); the 44.4 MB bank is L2 resident (bank stream 25 GB/s: not a bound at this batch)"}
        roof["frac"] = roof["achieved"] / roof["peak"]
    return {"metric": "tpwl_%s_rollout_steps_per_sec" % method, "value": steps_total / t_dev, "unit": "steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Diamond TPWL batched rollout (BASELINE configs[1]): %d trajectories x %d steps per GPU, "
                                   "n=72 m=4 P=1000 r=36, method %s (%s)" % (batch, N, method, "pre-discretised zoh bank" if method == "nn" else "fe per step"), "batch_per_gpu": batch,
                       "horizon": N, "l2": "256 MB buffer written between timed steps (untimed)"},
            "e2e": {"value": steps_total / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": int((x0h.size + uh.size) * 8),
                    "d2h_bytes_per_step": int(batch * (N + 1) * (n + 6) * 8)},
            "gpu_launches": args.steps * (1 if method == 'nn' else 2 * N + 2), "clocks": clk.summary(),
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
    # e2e: pinned host tensors in (x0, u), pinned host tensors out (x, z), copies inside the timed region
    x0p, up_ = torch.from_numpy(x0h).pin_memory(), torch.from_numpy(uh).pin_memory()
    g.rollout_pinned(x0p, up_, 0.02)                     # allocates the pinned outputs (untimed)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.rollout_pinned(x0p, up_, 0.02)
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
    # e2e: pinned host x / u in, pinned A, d, H, c, z out (a quarter of the states: 0.9 GB of results per step cross PCIe)
    cnt_e = count // 4
    xh, uh = x[:cnt_e].cpu().pin_memory(), u[:cnt_e].cpu().pin_memory()
    outh = None
    t_e2e = None
    for rep in range(3):
        if rep == 1:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        o = g._eval_device(xh.cuda(non_blocking=True), uh.cuda(non_blocking=True), -1.0, 'cont_raw', want)
        if outh is None:
            outh = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in o.items()}
        for k, v in outh.items():
            v.copy_(o[k], non_blocking=True)
        torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / 2
    hbm, hsrc, fp64 = measured_peaks()
    ach = count * 13944.0 / (t_dev / args.steps) / 1e12
    byts = count * (14 + 36 + 6 + 36 + 6 + 6) * 8
    return {"metric": "ssm_eval_linearize_states_per_sec", "value": count * world * args.steps / t_dev, "unit": "states/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Trunk SSM batched evaluation + linearisation: %d states per GPU, outputs A_c, d_c, H, c, z "
                                   "(inputs %.0f MB + outputs %.0f MB per launch: larger than L2)" % (count, count * 14 * 8 / 1e6, byts / 1e6 - count * 14 * 8 / 1e6)},
            "e2e": {"value": cnt_e * world / t_e2e, "unit": "states/s", "h2d_bytes_per_step": int(cnt_e * 14 * 8),
                    "d2h_bytes_per_step": int(cnt_e * 90 * 8),
                    "note": "PCIe-bound: 720 B of results per state leave the device; consumers on the device (SCP adapters, "
                            "iLQR) never take this path"},
            "gpu_launches": args.steps, "clocks": clk.summary(),
            "roofline": {"kernel": "ssm_eval_sparse_dmma_kernel<8>", "bound": "tensor", "achieved": ach, "peak": fp64, "unit": "TFLOP/s",
                         "frac": ach / fp64, "traffic": None, "hbm_gbs": byts / (t_dev / args.steps) / 1e9,
                         "hbm_frac": byts / (t_dev / args.steps) / 1e9 / hbm,
                         "executed_tflops": 12 * 512.0 * count / (t_dev / args.steps) / 1e12,
                         "executed_frac": 12 * 512.0 * count / (t_dev / args.steps) / 1e12 / fp64,
                         "note": "achieved = 13944 ALGORITHMIC flop per state (the dense count of SURVEY 8d: f, A = df/dx, C(x), H) / time; "
                                 "the kernel contracts only the structural non-zeros of d phi / d x (12 DMMA m8n8k4 = 6144 executed flop "
                                 "per state, executed_tflops; the dense formulation needs 42 DMMA, of which 35 % padding: "
                                 "SRCB200_SSM_EVAL_DENSE=1, 0.67 G states/s), so the algorithmic fraction can approach 1 while the tensor "
                                 "pipe is ~45 % busy; the other bound is the 832 B of I/O per state (hbm_frac of the measured copy "
                                 "bandwidth); peak = measured cuBLAS DGEMM"}}


def pod_sharded_matrix(nf_local, ns, rank, world, r=192, seed=5):
    """Row shard of a synthetic snapshot matrix with a PRESCRIBED spectrum: X = A diag(s) B^T
    with s = synth.pod_spectrum(r) (the Diamond-like decay of the small-scale twin), A (sum_g nf_g x r) orthonormal
    ACROSS the shards (Cholesky-QR with one r x r all-reduce), B (ns x r) orthonormal and identical on every rank.
    Generated on the device (torch, untimed: data generation is not the path).  Returns (X_g, s)."""
    import torch
    import torch.distributed as dist
    import sofacontrol_b200.synth as synth
    g = torch.Generator(device="cuda").manual_seed(seed * 1000 + rank)
    A = torch.randn((nf_local, r), device="cuda", dtype=torch.float64, generator=g)
    C_ = A.t() @ A
    if world > 1:
        dist.all_reduce(C_)
    A = A @ torch.linalg.inv(torch.linalg.cholesky(C_).t())          # A R^-1 with C = R^T R: orthonormal across shards
    gb = torch.Generator(device="cuda").manual_seed(seed)
    B, _ = torch.linalg.qr(torch.randn((ns, r), device="cuda", dtype=torch.float64, generator=gb))
    sv = torch.from_numpy(synth.pod_spectrum(r)).cuda()
    X = (A * sv) @ B.t()
    return X.contiguous(), sv


def run_pod_gram(args, rank, world, dev_index):
    """Kernel (d) + its callers: POD of a row-sharded snapshot matrix -- Gram X^T X per rank (DMMA SYRK; block rows with
    the all-reduce of finished pieces overlapped when world > 1), leading eigenpairs + energy rule (replicated), sharded
    back-projection U_g = X_g V S^-1.  --pod-full: the per-GPU slice of BASELINE configs[4] (250 000 x 20 000)."""
    import torch
    import torch.distributed as dist
    from sofacontrol_b200.mor import pod, eig
    from sofacontrol_b200 import parallel
    import sofacontrol_b200.synth as synth
    nf_local, ns = (250000, 20000) if args.pod_full else (131072, 8192)
    tol = 5e-5
    X, sv = pod_sharded_matrix(nf_local, ns, rank, world)
    lam_true = (sv ** 2)
    modes_expected = eig.energy_mode_count(lam_true, lam_true.sum(), tol)
    nblk = max(1, args.pod_overlap) if world > 1 else 1

    def gram():
        if world > 1 and nblk > 1:
            return parallel.overlapped_gram_allreduce(X, nblk)
        G_ = pod.gram_device(X)
        return parallel.allreduce_sum_(G_)

    for _ in range(max(1, args.warmup - 1)):
        G = gram()
    torch.cuda.synchronize()
    E = lambda: torch.cuda.Event(enable_timing=True)
    ev = [(E(), E(), E(), E()) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(dev_index) as clk:
        for e0, e1, e2, e3 in ev:
            e0.record()
            G = gram()
            e1.record()
            lam, V, nb, info = eig.leading_eigenpairs(G, tol)
            e2.record()
            S = lam[:nb].sqrt()
            U = pod.dgemm_device(X, (V[:, :nb] / S).contiguous())
            e3.record()
        torch.cuda.synchronize()
    t_gram = sum(a.elapsed_time(b) for a, b, _, _ in ev) * 1e-3
    t_eig = sum(b.elapsed_time(c) for _, b, c, _ in ev) * 1e-3
    t_all = sum(a.elapsed_time(d) for a, _, _, d in ev) * 1e-3
    fl = 2.0 * nf_local * ns * ns
    # untimed checks: the modes are the prescribed ones, U is orthonormal across the shards
    UtU = U.t() @ U
    if world > 1:
        dist.all_reduce(UtU)
        tt = torch.tensor([t_gram, t_eig, t_all], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_gram, t_eig, t_all = float(tt[0]), float(tt[1]), float(tt[2])
    orth = float((UtU - torch.eye(nb, device="cuda", dtype=torch.float64)).abs().max())
    sig_err = float(((S - sv[:nb]).abs().max() / sv[0]))
    e2e = None
    if world == 1 and not args.pod_full:
        # e2e: the reference-facing call compute_POD (pod.py:181-200) on a HOST snapshot matrix, modes back on the host
        Xh = X.cpu().numpy()
        del U, G
        t0 = time.perf_counter()
        U_full, U_h, nb_h, S_h = pod.compute_POD(Xh, tol)
        t_e2e = time.perf_counter() - t0
        e2e = {"value": fl / t_e2e / 1e12, "unit": "TFLOP/s (algorithmic Gram flops / whole compute_POD call)",
               "h2d_bytes_per_step": int(Xh.nbytes), "d2h_bytes_per_step": int(U_full.nbytes + S_h.nbytes),
               "seconds": t_e2e, "modes": int(nb_h),
               "note": "host numpy matrix in (pageable, 8.6 GB over PCIe), U / Sigma out: the copy dominates"}
        del Xh
    hbm, hsrc, fp64 = measured_peaks()
    ach = fl / (t_gram / args.steps) / 1e12
    return {"metric": "pod_gram_tflops", "value": world * fl * args.steps / t_gram / 1e12,
            "unit": "TFLOP/s (algorithmic Gram flops, all-reduce inside the timed region)",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "POD of a row-sharded snapshot matrix (BASELINE configs[4]%s): X_g %d x %d FP64 per GPU (%.1f GB), "
                                   "prescribed spectrum; G = sum_g X_g^T X_g (%s), leading eigenpairs by block subspace iteration, "
                                   "U_g = X_g V S^-1" % ("" if args.pod_full else ", reduced per-GPU slice", nf_local, ns,
                                                         nf_local * ns * 8 / 1e9,
                                                         "%d block rows, all-reduce of finished rows overlapped" % nblk if nblk > 1
                                                         else "one SYRK launch" + (" + one all-reduce" if world > 1 else "")),
                       "gram_ms": 1e3 * t_gram / args.steps, "eig_ms": 1e3 * t_eig / args.steps,
                       "project_ms": 1e3 * (t_all - t_gram - t_eig) / args.steps,
                       "modes": int(nb), "modes_expected": int(modes_expected), "sigma_relerr": sig_err,
                       "U_orthonormality": orth, "eig_iterations": info.get("iterations"), "eig_block": info.get("block")},
            "e2e": e2e, "gpu_launches": args.steps * nblk + args.steps * 12, "clocks": clk.summary(),
            "roofline": {"kernel": "dgemm_kernel<true,true,true> (SYRK)" if nblk == 1 else "dgemm_kernel<true,false,true> (block rows)",
                         "bound": "tensor", "achieved": ach, "peak": fp64,
                         "unit": "TFLOP/s", "frac": ach / fp64, "traffic": 86.4e9 if (world == 1 and not args.pod_full) else None,
                         "note": "algorithmic 2 nf ns^2 flop / Gram time (incl. the overlapped all-reduce when world > 1); the kernel "
                                 "executes only the upper tiles (half), so frac can exceed 1; peak = measured cuBLAS DGEMM; traffic = ncu "
                                 "dram read+write of one SYRK launch at this size (profiles/ncu_gram_r2_mbar_pace.txt: 83-168 GB between captures, 8.6 GB algorithmic; tensor pipe 88-89 % active, DRAM at 5-9 % of its peak: the re-reads of free-running CTAs are not the bound, the pace keeper that removes them costs 5 %)"}}


def run_mpc(args, rank, world, dev_index):
    """BASELINE configs[3] "Diamond SSM + TPWL closed-loop MPC Monte Carlo: 16k receding-horizon problems, re-linearised
    at every step": plant = Diamond-shaped TPWL bank (n = 72, nearest-neighbour on the zoh bank), controller =
    receding-horizon iLQR (N = 20) on the Diamond SSM fed by the SSM observer, warm start = shifted previous plan,
    u_last, process noise; every control step re-solves ALL problems (each forward step re-linearises the SSM)."""
    import torch
    import torch.distributed as dist
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200 import _lib as L
    from sofacontrol_b200.mpc import RecedingHorizonILQR, SSMOutputBelief
    batch = args.batch if args.batch != 4096 else 16384
    N, steps = 20, args.mpc_steps
    w = synth.mpc_ssm_tpwl_workload(batch, steps=steps, N=N, seed=4)
    mpc = RecedingHorizonILQR(w['solver'], plant=w['plant'], observer=SSMOutputBelief(w['ssm']), process_noise_std=1e-3,
                              seed=4 + rank)
    xb, xp, zd = L.to_dev(w['x0_belief']), L.to_dev(w['x0_plant']), L.to_dev(w['z_ref'])
    mpc.run_device(xb, zd, 3, xp)                                # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    with ClockSampler(dev_index) as clk:
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        out = mpc.run_device(xb, zd, steps, xp)
        e_.record()
        torch.cuda.synchronize()
    t_dev = s_.elapsed_time(e_) * 1e-3
    # e2e: host arrays in, closed-loop record out
    t0 = time.perf_counter()
    outh = mpc.run(w['x0_belief'], w['z_ref'], steps, x0_plant=w['x0_plant'])
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    its = out['iterations'].float().mean().item()
    h2d = (w['x0_belief'].size + w['x0_plant'].size + w['z_ref'].size) * 8
    d2h = sum(v.size * v.itemsize for v in outh.values())
    return {"metric": "mpc_problem_steps_per_sec", "value": batch * steps * world / t_dev, "unit": "receding-horizon solves/s",
            "n_gpus": world, "steps": steps, "warmup": 3, "ms_per_step": 1e3 * t_dev / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "closed-loop receding-horizon iLQR Monte Carlo (BASELINE configs[3]): %d problems per GPU x %d "
                                   "control steps; plant = Diamond TPWL bank (n=72, P=1000, nn, zoh), controller = Diamond-SSM "
                                   "iLQR horizon %d with SSM observer, warm start + u_last, process noise 1e-3"
                                   % (batch, steps, N),
                       "mean_iterations_per_solve": its, "converged_frac": float((out['status'] & 1).float().mean().item()),
                       "launches_per_control_step": "2 solver + 1 plant step + 1 observer map + 1 glue (mpc_shift) + 1 index_select"},
            "e2e": {"value": batch * steps * world / t_e2e, "unit": "receding-horizon solves/s", "h2d_bytes_per_step": int(h2d // steps),
                    "d2h_bytes_per_step": int(d2h // steps)},
            "gpu_launches": steps * 5, "clocks": clk.summary(),
            "roofline": {"kernel": "ilqr_ssm_fast_kernel<4> (per control step)", "bound": "tensor", "achieved": None, "peak": None,
                         "unit": "TFLOP/s", "frac": None, "traffic": None,
                         "note": "launch- and latency-bound loop of short solves (horizon 20, warm-started: a few iterations); "
                                 "the per-step kernels are the headline iLQR kernel and the nn rollout kernel, whose rooflines "
                                 "are reported by their own workloads"}}


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
    # e2e: the reference-facing call, host arrays in (x0, targets), host arrays out (x, u, K)
    solver.set_target(zt)
    t0 = time.perf_counter()
    xh, uh_, Kh = solver.ilqr_computation(x0)
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt[0])
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
            "e2e": {"value": batch * world / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": int((x0.size + zt.size) * 8),
                    "d2h_bytes_per_step": int((xh.size + uh_.size + Kh.size) * 8),
                    "note": "one iLQR.ilqr_computation call on host arrays (pageable), results x, u, K back on the host"},
            "gpu_launches": args.steps, "clocks": clk.summary(),
            "roofline": {"kernel": "ilqr_solve_kernel<TpwlPolicyT<72,4,6>>", "bound": "tensor", "achieved": ach, "peak": fp64,
                         "unit": "TFLOP/s", "frac": ach / fp64, "traffic": None,
                         "note": "flops of the three DMMA products of every backward step only (forward passes and the "
                                 "nearest-point search not counted) / whole-solve time; peak = measured cuBLAS DGEMM"}}


def run_reference(args):
    """--impl reference: the reference's own iLQR class on the host cores (imported from /root/reference when present,
    else its CPU port), same workload/metric, bounded sample per step."""
    steps = max(1, min(args.steps, 3))
    vals, busys = [], []
    cores = sample = kind = None
    for i in range(steps):
        val, cores, sample, kind, busy = cpu_ilqr_baseline(args.batch, args.horizon, seed=3, per_core=args.cpu_per_core)
        vals.append(val); busys.append(busy)
    val = max(vals)
    return {"impl": "reference", "metric": "ilqr_solves_per_sec", "value": val, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 0, "ms_per_step": 1e3 * min(busys), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Trunk-SSM batched iLQR (BASELINE configs[2]): horizon %d, n=6 m=8 order-3 SSM, be "
                                   "discretisation, dt=0.02, seed 3 -- the reference algorithm on the host cores; one step = "
                                   "%d solves per core of the same batch" % (args.horizon, args.cpu_per_core),
                       "batch_per_gpu": args.batch, "horizon": args.horizon},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


# ---------------------------------------------------------------------------------------------------------------
# CPU baselines of the secondary workloads (rank 0, N = 1): the oracle ports on the host cores, bounded samples
# ---------------------------------------------------------------------------------------------------------------
def _cpu_tpwl_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    idx, N = args
    import sofacontrol_b200.synth as synth
    from oracle.tpwl_np import TPWLATVNP
    data, Hf = synth.tpwl_bank()
    o = TPWLATVNP(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='fe')
    o.pre_discretize(0.01)          # fe bank: the per-step work (nearest point + affine step) does not depend on the method
    x0, u = synth.tpwl_rollout_batch(max(idx) + 1, N=N, seed=2)
    t0 = time.perf_counter()
    for b in idx:
        o.rollout(x0[b], u[b], 0.01)
    return len(idx) * N, time.perf_counter() - t0


def _cpu_tpwl_weighting_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    idx, N = args
    import sofacontrol_b200.synth as synth
    from oracle.tpwl_np import TPWLATVNP
    data, Hf = synth.tpwl_bank()
    o = TPWLATVNP(data, params={'tpwl_method': 'weighting', 'dist_weights': {'q': 1.0, 'v': 0.0}, 'beta_weighting': 25.0},
                  Hf=Hf, discr_method='fe')
    x0, u = synth.tpwl_rollout_batch(max(idx) + 1, N=N, seed=2)
    t0 = time.perf_counter()
    for b in idx:
        o.rollout(x0[b], u[b], 0.01)
    return len(idx) * N, time.perf_counter() - t0


def _cpu_ssm_rollout_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    count, N, seed = args
    import sofacontrol_b200.synth as synth
    from oracle.ssm_np import SSMDynamicsNP
    s = synth.trunk_ssm(8)
    o = SSMDynamicsNP(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    for _ in range(count):
        o.rollout(0.05 * rng.normal(size=6), rng.uniform(0, 800, size=(N, 8)), 0.02)
    return count * N, time.perf_counter() - t0


def _cpu_ssm_eval_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    count, seed = args
    import sofacontrol_b200.synth as synth
    from oracle.ssm_np import SSMDynamicsNP
    s = synth.trunk_ssm(8)
    o = SSMDynamicsNP(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    rng = np.random.default_rng(seed)
    X, U = rng.normal(size=(count, 6)), rng.uniform(0, 800, size=(count, 8))
    t0 = time.perf_counter()
    for x, u in zip(X, U):
        o.get_continuous_jacobians(x, u)
        o.get_observer_jacobians(x)
        o.x_to_zfyf(x)
    return count, time.perf_counter() - t0


def _cpu_pool(worker, jobs):
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(len(jobs)) as pool:
        res = pool.map(worker, jobs)
    return sum(r[0] for r in res) / max(r[1] for r in res)


def secondary_cpu_baseline(name):
    cores = min(64, max(1, len(os.sched_getaffinity(0))))
    if name == "tpwl_rollout_nn":
        v = _cpu_pool(_cpu_tpwl_worker, [(list(range(c * 2, c * 2 + 2)), 100) for c in range(cores)])
        return {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                "sample": "2 trajectories x 100 steps per core of the same seeded batch (oracle/tpwl_np.py, numpy)"}
    if name == "tpwl_rollout_weighting":
        v = _cpu_pool(_cpu_tpwl_weighting_worker, [([c], 40) for c in range(cores)])
        return {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                "sample": "1 trajectory x 40 steps per core, method weighting, fe per step (oracle/tpwl_np.py, numpy)"}
    if name == "ssm_rollout":
        v = _cpu_pool(_cpu_ssm_rollout_worker, [(2, 100, c) for c in range(cores)])
        return {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                "sample": "2 trajectories x 100 steps per core, be discretisation (oracle/ssm_np.py, numpy)"}
    if name == "ssm_eval":
        v = _cpu_pool(_cpu_ssm_eval_worker, [(400, c) for c in range(cores)])
        return {"value": v, "unit": "states/s", "cores": cores, "kind": "port",
                "sample": "400 states per core: continuous Jacobians + observer Jacobians + output (oracle/ssm_np.py)"}
    if name == "pod_gram":
        nf, ns = 16384, 2048
        X = np.random.default_rng(5).normal(size=(nf, ns))
        t0 = time.perf_counter()
        X.T @ X
        dt = time.perf_counter() - t0
        return {"value": 2.0 * nf * ns * ns / dt / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                "sample": "numpy X^T X (multi-threaded BLAS) on a %d x %d slice" % (nf, ns)}
    return None


def _trim(res):
    keep = ("metric", "value", "unit", "ms_per_step", "steps", "config", "e2e", "roofline", "gpu_launches", "cpu_baseline")
    return {k: res[k] for k in keep if k in res}


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
    ap.add_argument("--mpc-steps", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--pod-overlap", type=int, default=1, help="pod_gram on > 1 GPU: block rows of the Gram matrix whose all-reduce overlaps the next rows (1 = one SYRK + one all-reduce, the faster choice: the reduction is < 1 %% of the contraction)")
    ap.add_argument("--pod-full", action="store_true", help="pod_gram at the per-GPU slice of config 5 (250000 x 20000, 40 GB)")
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
    runners = {"ilqr_trunk_ssm": run_ilqr, "ssm_rollout": run_ssm_rollout, "ssm_eval": run_ssm_eval, "pod_gram": run_pod_gram,
               "ilqr_tpwl": run_ilqr_tpwl, "mpc": run_mpc,
               "tpwl_rollout_nn": lambda a, r, w_, l: run_tpwl_rollout(a, r, w_, l, "nn"),
               "tpwl_rollout_weighting": lambda a, r, w_, l: run_tpwl_rollout(a, r, w_, l, "weighting")}
    res = runners[args.workload](args, rank, world, local)
    headline = (args.workload == "ilqr_trunk_ssm")
    if headline and not args.no_secondary:
        # the other kernels of the path, outside the headline's timed region (their own events, warm-up and L2 policy)
        import copy
        sec = {}
        for name, steps in (("tpwl_rollout_nn", 5), ("tpwl_rollout_weighting", 2), ("ssm_eval", 5), ("ssm_rollout", 5),
                            ("ilqr_tpwl", 2), ("pod_gram", 2), ("mpc", 1)):
            a2 = copy.copy(args)
            a2.batch, a2.steps, a2.warmup, a2.horizon = 4096, steps, 3, 100
            if name == "ilqr_tpwl":
                a2.batch, a2.warmup = 1184, 1                # 8 problems per SM: the Diamond TPWL solver runs one CTA per problem
            a2.mpc_steps = 25
            try:
                r2 = runners[name](a2, rank, world, local)
                if rank == 0 and world == 1 and not args.no_cpu_baseline:
                    cb = secondary_cpu_baseline(name)
                    if cb is not None:
                        r2["cpu_baseline"] = cb
                sec[name] = _trim(r2)
            except Exception as e:                      # a secondary workload must never take the headline down
                sec[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        res["secondary"] = sec
    if rank == 0 and world == 1 and not args.no_cpu_baseline and headline:
        v, cores, sample, kind, _ = cpu_ilqr_baseline(args.batch, args.horizon, seed=3, per_core=args.cpu_per_core)
        res["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": cores, "kind": kind, "sample": sample}
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
