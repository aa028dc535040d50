"""One forward pass and one backward pass of the Diamond TPWL iLQR on a batch (unit entry points), for per-pass timing
under `ncu --metrics gpu__time_duration.sum`.

Phase clocks of the backward sweep: compile csrc/ilqr_tpwl_diamond.cu with -DSRCB_PHASE_TIMING (the kernel then writes
its per-phase clock64 totals into the Q_u output), relink, and run with SRCB_PHASE_TIMING=1 in the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sofacontrol_b200.synth as synth
from sofacontrol_b200.tpwl.tpwl import TPWLATV
from sofacontrol_b200.lqr.ilqr import iLQR
from sofacontrol_b200.utils import QuadraticCost

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 148
N = 100
data, Hf = synth.tpwl_bank()
g = TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='zoh')
g.pre_discretize(0.01)
Q = np.zeros((6, 6)); Q[3, 3] = Q[4, 4] = 100.0
R = 1e-5 * np.eye(4)
th = np.linspace(0, 2 * np.pi, N + 1)
x0, u = synth.tpwl_rollout_batch(batch, N=N, seed=21)
zt = np.tile(g.z_ref, (batch, N + 1, 1))
zt[:, :, 3] += np.sin(th)[None]; zt[:, :, 4] += np.sin(2 * th)[None]
s = iLQR(0.01, g, QuadraticCost(Q, R, np.zeros((6, 6))), N)
s.set_target(zt)
xp = np.zeros((batch, N + 1, 72)); xp[:, 0] = x0
x, uu, cost, A, B, d = s.forward_pass(xp, 0.1 * u)
torch.cuda.synchronize()
s.rho, s.drho = 0.0, 0.0
K, k, Qu, Quu = s.dlqr_recursion(x, uu, A, B, d)
torch.cuda.synchronize()
x2, u2, cost2, *_ = s.forward_pass(x, uu, 1.0, K, k)
torch.cuda.synchronize()
print('ok', float(np.mean(cost)), float(np.mean(cost2)))
if os.environ.get('SRCB_PHASE_TIMING'):
    q = np.asarray(Qu).reshape(batch, N, 4)
    ph = q[:, :2, :].reshape(batch, 8)
    names = ['top wait+sync', 'P1', 'sync', 'P2+sync', 'chol/inv+sync', 'gains+sync', '-', 'P3 (+prologue)']
    print('cycles per step (mean over CTAs):', {n: round(float(v) / N) for n, v in zip(names, ph.mean(0))})
