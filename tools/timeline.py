"""Timeline of the 4096-problem benchmark batch (build with -DSRCB_PHASE_TIMING: the per-iteration trace then carries
%globaltimer instead of the PD flag): active problems over time and the duration of an iteration early / late in the run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sofacontrol_b200 import _lib as L
from sofacontrol_b200.lqr.ilqr import iLQR
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w, solver0 = bench.build_ilqr(batch, 100, 3)
solver = iLQR(w["dt"], solver0.model, solver0.cost_params, 100, trace=True)
x0, zt = L.to_dev(w['x0']), L.to_dev(w['z_target'])
solver.solve_device(x0, zt)
out = solver.solve_device(x0, zt)
it = out['iterations'].cpu().numpy().astype(int); tr = out['trace'].cpu().numpy(); trials = out['trials'].cpu().numpy().astype(int)
T = tr[:, :, 3]
t0 = min(T[b, 0] for b in range(batch))
end = np.array([T[b, it[b] - 1] for b in range(batch)]) - t0
print("batch %d: makespan %.2f ms" % (batch, end.max() * 1e-6))
for frac in (0.25, 0.5, 0.75, 0.9, 0.95, 0.99, 1.0):
    print("  %5.1f %% of the problems finished by %.2f ms" % (100 * frac, np.quantile(end, frac) * 1e-6))
grid = np.linspace(0, end.max(), 11)
for a, b_ in zip(grid[:-1], grid[1:]):
    act = ((end > a)).sum()
    durs = []
    for b in range(batch):
        tt = T[b, :it[b]] - t0
        d = np.diff(tt)
        m = (tt[1:] > a) & (tt[1:] <= b_)
        durs.extend(d[m])
    print("  window %5.1f-%5.1f ms: %4d problems still active at its start, %6d iterations ended, median iteration %.0f us"
          % (a * 1e-6, b_ * 1e-6, act, len(durs), np.median(durs) * 1e-3 if durs else 0))
lw = np.argmax(it + trials)
tt = T[lw, :it[lw]] - t0
print("longest solve: %d iterations, %d passes, finished at %.2f ms; its iteration durations (us):" % (it[lw], it[lw] + trials[lw], tt[-1] * 1e-6))
print("  ", np.round(np.diff(tt) * 1e-3).astype(int).tolist())
