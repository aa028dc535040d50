import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import sofacontrol_b200.synth as synth
from sofacontrol_b200 import _lib as L
from sofacontrol_b200.tpwl.tpwl import TPWLATV
data, Hf = synth.tpwl_bank()
g = TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='zoh')
g.pre_discretize(0.01)
x0h, uh = synth.tpwl_rollout_batch(4096, N=100, seed=2)
x0, u = L.to_dev(x0h), L.to_dev(uh)
flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)
for mode in ("flush", "noflush", "flush_want_z_false"):
    ts = []
    for i in range(14):
        if mode != "noflush": flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.rollout_device(x0, u, 0.01, want_z=(mode != "flush_want_z_false"))
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(mode, " ".join("%.2f" % t for t in ts))
