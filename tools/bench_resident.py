"""tpwl_resident.cu (SRCB200_TPWL_RESIDENT=1) against tpwl_screen.cu (default) on the config-2 rollout, at the
benchmark's dt = 0.01 and at smaller steps (slower movement against the spacing of the stored points)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sofacontrol_b200.synth as synth
from sofacontrol_b200 import _lib as L
from sofacontrol_b200.tpwl.tpwl import TPWLATV

data, Hf = synth.tpwl_bank()
x0h, uh = synth.tpwl_rollout_batch(4096, N=100, seed=2)
x0, u = L.to_dev(x0h), L.to_dev(uh)
for dt in (0.01, 0.003, 0.001):
    g = TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='zoh')
    g.pre_discretize(dt)
    res = {}
    for tag, env in (("screen", "0"), ("resident", "1")):
        os.environ["SRCB200_TPWL_RESIDENT"] = env
        for _ in range(2):
            x, z, idx = g.rollout_device(x0, u, dt, want_idx=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            g.rollout_device(x0, u, dt, want_z=False)
        b.record()
        torch.cuda.synchronize()
        res[tag] = (a.elapsed_time(b) / 3, idx.cpu().numpy(), x.cpu().numpy())
    os.environ["SRCB200_TPWL_RESIDENT"] = "0"
    idx = res["screen"][1]
    same = float((idx[:, 1:] == idx[:, :-1]).mean())
    print("dt %.3f: screen %.2f ms, resident %.2f ms; index unchanged on %.1f %% of the steps; traces equal %s, states equal %s"
          % (dt, res["screen"][0], res["resident"][0], 100 * same, np.array_equal(idx, res["resident"][1]),
             np.array_equal(res["screen"][2], res["resident"][2])))
