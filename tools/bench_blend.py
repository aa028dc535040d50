"""Times the TPWL weighted bank blend (kernel a3) at the Diamond size for small and large batches."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sofacontrol_b200.synth as synth
from sofacontrol_b200 import _lib as L
from sofacontrol_b200.tpwl.tpwl import TPWLATV

data, Hf = synth.tpwl_bank()
g = TPWLATV(data, params={'tpwl_method': 'weighting', 'dist_weights': {'q': 1.0, 'v': 0.0}, 'beta_weighting': 25.0}, Hf=Hf, discr_method='fe')
out = {}
flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)
for cnt in (1, 8, 32, 64, 256, 4096):
    x0, _ = synth.tpwl_rollout_batch(cnt, N=1, seed=3)
    xd = L.to_dev(x0)
    for _ in range(3):
        g.linearize_device(xd, None)
    torch.cuda.synchronize()
    ts = []
    for cold in (True, False):
        best = 1e9
        for _ in range(5):
            if cold:
                flush.fill_(1.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); g.linearize_device(xd, None); e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) * 1e-3)
        ts.append(best)
    bank_bytes = 1000 * (72 * 72 + 72 * 4 + 72) * 8
    out[cnt] = {"cold_l2_ms": ts[0] * 1e3, "warm_l2_ms": ts[1] * 1e3, "bank_gbs_cold": bank_bytes * max(1, -(-cnt // 8) if cnt <= 32 else 1) / ts[0] / 1e9,
                "tflops": 2.0 * cnt * bank_bytes / 8 / ts[1] / 1e12}
print(json.dumps(out))
