"""Golden vectors AT THE BENCHMARK SHAPES, produced by the UNMODIFIED reference classes (imported from
/root/reference; osqp stubbed) -- TEST INFRASTRUCTURE ONLY, runs only in the build container.

    python -m oracle.make_golden_bench [ilqr] [tpwl]

  ilqr_bench_seed3.npz   BASELINE config 3 (the headline): 32 members of the seed-3 Trunk-SSM batch (N = 100, m = 8,
                         4096 problems) solved one by one by the reference iLQR class (ilqr.py) driving the SSM
                         through the Gauss-Newton H-property adapter.  Members are chosen from the iteration
                         histogram of the batch to cover the shortest solves (5 iterations), the only 51-iteration
                         (max_iter) member, abandoned line searches and the middle of both modes.
                         Stored: member indices, x, u, K, iterations, final rho, how the loop ended.
  tpwl_bench.npz         BASELINE config 2 shapes (P = 1000, n = 72, m = 4): reference TPWLATV `weighting` mode
                         (beta = 25) get_jacobians (fe / be / bil / zoh) at 6 states and a 20-step rollout;
                         nn rollout of 100 steps on the pre-discretised zoh bank for 6 trajectories.
"""
import contextlib
import io
import os
import sys
import warnings
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
GOLD = os.path.join(REPO, "tests", "golden")

# chosen from the iteration histogram of the 4096-problem seed-3 batch (tools/iter_trace_dump.py):
# shortest (5 it), the 51-iteration member, members with abandoned / failed line searches, both modes of the histogram
BENCH_MEMBERS = [10, 225, 256, 3048, 6, 33, 113, 125, 229, 266, 0, 1, 2, 3, 4, 5, 7, 8, 9, 11, 12, 100, 500, 1000,
                 1500, 2000, 2500, 3000, 3500, 4000, 4094, 4095]


def _solve_member(b):
    warnings.simplefilter("ignore")
    from oracle import refimport, ssm_np
    import sofacontrol_b200.synth as synth
    ref = refimport.load()
    w = synth.trunk_ilqr_batch(4096, N=100, seed=3, m=8)
    s = w['ssm']
    mdl = ssm_np.GaussNewtonSSM(ssm_np.SSMDynamicsNP(s['z_ref'], discrete=False, discr_method='be', model=s['model'],
                                                     params=s['params']))
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    sol = ref.ilqr.iLQR(w['dt'], mdl, ref.utils.QuadraticCost(Q, R, Qf), 100)
    sol.set_target(w['z_target'][b])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), np.errstate(all='ignore'):
        x, u, K = sol.ilqr_computation(w['x0'][b])
    log = buf.getvalue()
    it = log.count('Iteration')
    ended = 4 if 'abandoning search' in log else (1 if log.rstrip().endswith('Cost converged') else 2)   # status bit
    return b, x, u, K, it, float(sol.rho), ended, log.count('No improved cost found')


def ilqr_bench():
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        res = list(ex.map(_solve_member, BENCH_MEMBERS))
    out = dict(members=np.array([r[0] for r in res]), x=np.array([r[1] for r in res]), u=np.array([r[2] for r in res]),
               K=np.array([r[3] for r in res]), iterations=np.array([r[4] for r in res]),
               rho=np.array([r[5] for r in res]), status=np.array([r[6] for r in res]),
               failed_linesearches=np.array([r[7] for r in res]))
    np.savez_compressed(os.path.join(GOLD, "ilqr_bench_seed3.npz"), **out)
    for r in res:
        print("member %4d: %2d iterations, rho %.6g, status %d, failed line searches %d" % (r[0], r[4], r[5], r[6], r[7]))


def tpwl_bench():
    from oracle import refimport
    import sofacontrol_b200.synth as synth
    ref = refimport.load()
    data, Hf = synth.tpwl_bank()
    x0, u = synth.tpwl_rollout_batch(6, N=100, seed=2)
    out = dict(x0=x0, u=u)
    prm = {'tpwl_method': 'weighting', 'dist_weights': {'q': 1.0, 'v': 0.0}, 'beta_weighting': 25.0}
    for meth in ('fe', 'be', 'bil', 'zoh'):
        mw = ref.tpwl.TPWLATV(data, params=prm, Hf=Hf, discr_method=meth)
        J = [mw.get_jacobians(x, dt=0.01) for x in x0]
        for i, nm in enumerate('ABd'):
            out['w_%s_%s' % (nm, meth)] = np.array([j[i] for j in J])
    mw = ref.tpwl.TPWLATV(data, params=prm, Hf=Hf, discr_method='be')
    out['weights'] = np.array([mw.calc_weighting_factors(x) for x in x0])
    R = [mw.rollout(x0[b], u[b, :20], 0.01) for b in range(3)]
    out['w_roll_x'] = np.array([r[0] for r in R]); out['w_roll_z'] = np.array([r[1] for r in R])
    mn = ref.tpwl.TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='zoh')
    with contextlib.redirect_stdout(io.StringIO()):
        mn.pre_discretize(0.01)
    idx = []
    R = []
    for b in range(6):
        xs = np.zeros((101, 72)); xs[0] = x0[b]
        ii = []
        for t in range(100):
            xs[t + 1] = mn.update_state(xs[t], u[b, t], 0.01)
            ii.append(mn.ref_point)
        R.append(xs); idx.append(ii)
    out['nn_zoh_x'] = np.array(R); out['nn_zoh_idx'] = np.array(idx)
    out['nn_zoh_z'] = np.array([mn.x_to_zfyf(x, zf=True) for x in R])
    # three bank entries of the reference's scipy-expm pre-discretisation (the device uses its own Pade-13 kernel)
    out['zoh_bank_idx'] = np.array([0, 499, 999])
    out['zoh_A_d'] = np.array([mn.A_d[i] for i in (0, 499, 999)])
    out['zoh_B_d'] = np.array([mn.B_d[i] for i in (0, 499, 999)])
    out['zoh_d_d'] = np.array([mn.d_d[i] for i in (0, 499, 999)])
    np.savez_compressed(os.path.join(GOLD, "tpwl_bench.npz"), **out)
    print("tpwl bench golden: nn distinct indices", len(set(np.array(idx).ravel().tolist())))


if __name__ == "__main__":
    what = sys.argv[1:] or ['ilqr', 'tpwl']
    os.makedirs(GOLD, exist_ok=True)
    if 'ilqr' in what:
        ilqr_bench()
    if 'tpwl' in what:
        tpwl_bench()
