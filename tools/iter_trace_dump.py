"""Dump the per-iteration traces of the bench workload (cost, alpha, rho, restarts per iteration) for offline study of
how early an expensive solve can be told from a cheap one (scheduling only; results do not depend on it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sofacontrol_b200 import _lib as L
from sofacontrol_b200.lqr.ilqr import iLQR
from sofacontrol_b200.utils import QuadraticCost
w, solver0 = bench.build_ilqr(4096, 100, 3)
solver = iLQR(w["dt"], solver0.model, solver0.cost_params, 100, trace=True)
out = solver.solve_device(L.to_dev(w['x0']), L.to_dev(w['z_target']))
os.makedirs('gpurun_out', exist_ok=True)
np.savez_compressed('gpurun_out/iter_trace.npz', it=out['iterations'].cpu().numpy(), trials=out['trials'].cpu().numpy(),
                    cost0=out['cost0'].cpu().numpy(), trace=out['trace'].cpu().numpy())
print('saved', out['trace'].shape)
