"""Per-phase clock64 totals of the Trunk-SSM iLQR kernel (build csrc/ilqr_fast.cu with -DSRCB_PHASE_TIMING first:
SRCB_NVCC_EXTRA=-DSRCB_PHASE_TIMING python -m sofacontrol_b200._build --force).  Prints cycles per pass-step for a lone
warp (batch 1) and under load (batch 4096)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sofacontrol_b200 import _lib as L

lib = L.lib()
buf = (C.c_ulonglong * 32)()
names = {0: 'fwd u_t', 1: 'fwd prefetch issue', 2: 'fwd model eval', 3: 'fwd vector products + d_c', 4: 'fwd 2 x inverse (GJ)',
         5: 'fwd sep + A_d record', 6: 'fwd B_d, x_next', 8: 'bwd stage step', 9: 'bwd level 1', 10: 'bwd level 2',
         11: 'bwd gain solve', 12: 'bwd level 4', 13: 'bwd level 5'}
for batch in (1, 4096):
    w, solver = bench.build_ilqr(batch, 100, 3)
    x0, zt = L.to_dev(w['x0']), L.to_dev(w['z_target'])
    solver.solve_device(x0, zt)
    lib.srcb200_debug_phase(buf, 1)
    out = solver.solve_device(x0, zt)
    lib.srcb200_debug_phase(buf, 1)
    it = out['iterations'].cpu().numpy().astype(np.int64); tr = out['trials'].cpu().numpy().astype(np.int64)
    fwd_steps = float((tr.sum() + batch) * 101); bwd_steps = float(it.sum() * 100)
    print("batch %d: %d iterations, %d forward passes" % (batch, it.sum(), tr.sum() + batch))
    tf = tb = 0.0
    for i in range(16):
        if buf[i]:
            per = buf[i] / (fwd_steps if i < 8 else bwd_steps)
            print("  %-28s %8.1f cycles per step" % (names.get(i, str(i)), per))
            if i < 8: tf += per
            else: tb += per
    print("  forward step %.0f cycles, backward step %.0f cycles" % (tf, tb))
