import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import sofacontrol_b200.synth as synth
from sofacontrol_b200 import _lib as L
from sofacontrol_b200.SSM.ssm import SSMDynamics
from sofacontrol_b200.lqr.ilqr import iLQR
from sofacontrol_b200.utils import QuadraticCost
N = 100
for batch in (4096, 8192):
    w = synth.trunk_ilqr_batch(batch, N=N, seed=3, m=8)
    s = w['ssm']
    model = SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    solver = iLQR(w['dt'], model, QuadraticCost(Q, R, Qf), N)
    x0, zt = L.to_dev(w['x0']), L.to_dev(w['z_target'])
    for ch in ("0:1776,2:0", "0:3552,1:1776,2:0", "0:2960,1:1776,2:0", "0:2960,1:1184,2:0", "1:1776,2:0", "0:2368,1:1184,2:0", "0:1776,2:592,0:0", "0:1480,2:0"):
        os.environ["SRCB200_ILQR_CHAIN"] = ch
        solver.solve_device(x0, zt); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = solver.solve_device(x0, zt); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print("batch %5d chain %-24s %7.2f ms (%.1f k/s)" % (batch, ch, best, batch / best))
