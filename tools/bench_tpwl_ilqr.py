"""Times the generic iLQR kernel on the Diamond TPWL shape (n=72, m=4, P=1000, nn on the zoh bank)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sofacontrol_b200.synth as synth
from sofacontrol_b200 import _lib as L
from sofacontrol_b200.tpwl.tpwl import TPWLATV
from sofacontrol_b200.lqr.ilqr import iLQR
from sofacontrol_b200.utils import QuadraticCost

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
data, Hf = synth.tpwl_bank()
g = TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='zoh')
g.pre_discretize(0.01)
Q = np.zeros((6, 6)); Q[3, 3] = Q[4, 4] = 100.0
R = 1e-5 * np.eye(4)
th = np.linspace(0, 2 * np.pi, N + 1)
rng = np.random.default_rng(0)
x0, _ = synth.tpwl_rollout_batch(batch, N=1, seed=21)
amp = rng.uniform(0.3, 1.5, size=batch)
zt = np.tile(g.z_ref, (batch, N + 1, 1))
zt[:, :, 3] += amp[:, None] * np.sin(th)[None]; zt[:, :, 4] += amp[:, None] * np.sin(2 * th)[None]
s = iLQR(0.01, g, QuadraticCost(Q, R, np.zeros((6, 6))), N)
x0d, ztd = L.to_dev(x0), L.to_dev(zt)
out = s.solve_device(x0d, ztd); torch.cuda.synchronize()
t0 = time.perf_counter(); out = s.solve_device(x0d, ztd); torch.cuda.synchronize(); dt = time.perf_counter() - t0
it = out['iterations'].cpu().numpy(); tr = out['trials'].cpu().numpy(); st = out['status'].cpu().numpy()
print(json.dumps({"batch": batch, "N": N, "seconds": dt, "solves_per_s": batch / dt, "mean_iterations": float(it.mean()),
                  "mean_fwd": float(tr.mean() + 1), "converged": float((st & 1).mean()), "status_hist": np.bincount(st).tolist()}))
