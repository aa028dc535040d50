"""Launch shapes of the Trunk-SSM iLQR kernel (SRCB200_ILQR_SHAPE = 0 / 1 / 2) over batch sizes: ms per batch, best of 3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sofacontrol_b200.synth as synth
from sofacontrol_b200 import _lib as L
from sofacontrol_b200.SSM.ssm import SSMDynamics
from sofacontrol_b200.lqr.ilqr import iLQR
from sofacontrol_b200.utils import QuadraticCost

N = 100
for batch in (256, 512, 1024, 1184, 1536, 1776, 2048, 2368, 3072, 3584, 4096, 6144):
    w = synth.trunk_ilqr_batch(batch, N=N, seed=3, m=8)
    s = w['ssm']
    model = SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    solver = iLQR(w['dt'], model, QuadraticCost(Q, R, Qf), N)
    x0, zt = L.to_dev(w['x0']), L.to_dev(w['z_target'])
    row = []
    for shape in ("0", "1", "2"):
        os.environ["SRCB200_ILQR_SHAPE"] = shape
        solver.solve_device(x0, zt)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = solver.solve_device(x0, zt); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        row.append(best)
    print("batch %5d: shape 0 %7.2f ms, shape 1 %7.2f ms, shape 2 %7.2f ms  (iterations %d)" % (batch, row[0], row[1], row[2], int(out['iterations'].sum())))
