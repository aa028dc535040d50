"""Trunk-SSM iLQR: hand-over threshold of the two-launch schedule (SRCB200_ILQR_HANDOVER), ms per batch, best of 3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sofacontrol_b200.synth as synth
from sofacontrol_b200 import _lib as L
from sofacontrol_b200.SSM.ssm import SSMDynamics
from sofacontrol_b200.lqr.ilqr import iLQR
from sofacontrol_b200.utils import QuadraticCost

N = 100
for batch in (3072, 4096, 8192):
    w = synth.trunk_ilqr_batch(batch, N=N, seed=3, m=8)
    s = w['ssm']
    model = SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    solver = iLQR(w['dt'], model, QuadraticCost(Q, R, Qf), N)
    x0, zt = L.to_dev(w['x0']), L.to_dev(w['z_target'])
    ref = None
    for ho in ("0", "1184", "1776", "2368", "2960", "3552"):
        os.environ["SRCB200_ILQR_HANDOVER"] = ho
        out = solver.solve_device(x0, zt)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = solver.solve_device(x0, zt); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        sig = (out['x'].clone(), out['iterations'].clone())
        same = True if ref is None else (torch.equal(sig[0], ref[0]) and torch.equal(sig[1], ref[1]))
        if ref is None: ref = sig
        print("batch %5d handover %5s: %7.2f ms  (%.1f k solves/s)  results identical to handover 0: %s" % (batch, ho, best, batch / best, same))
