import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sofacontrol_b200 import _lib as L
w, solver = bench.build_ilqr(4096, 100, 3)
out = solver.solve_device(L.to_dev(w['x0']), L.to_dev(w['z_target']))
it = out['iterations'].cpu().numpy().astype(float); c0 = out['cost0'].cpu().numpy(); tr = out['trials'].cpu().numpy().astype(float)
work = it + tr
print('corr(cost0, iterations)', np.corrcoef(c0, it)[0,1], 'corr(log cost0, work)', np.corrcoef(np.log(c0), work)[0,1])
print('spearman-ish: top-10% cost0 mean work', work[np.argsort(-c0)[:410]].mean(), 'overall mean', work.mean(), 'max', work.max())
print('hist iterations', np.bincount(it.astype(int))[:60])
amp = None
