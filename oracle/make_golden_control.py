"""Golden vectors for the callers either side of the hot path, produced by the UNMODIFIED
reference modules (observer.py, lqr.py, traj_tracking_lqr.py, utils.extract_AB, tpwl_utils add_continuous_TPWL
arithmetic, gusto.compute_accuracy arithmetic on the reference TPWLGuSTO-style model).  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_control        ->  tests/golden/control_small.npz, control_diamond.npz
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
GOLD = os.path.join(REPO, "tests", "golden")


def meas_matrix(rows, nf):
    Cf = np.zeros((len(rows), nf))
    for i, j in enumerate(rows):
        Cf[i, j] = 1.0
    return Cf


class _T:
    pass


def small(ref):
    import sofacontrol_b200.synth as synth
    out = {}
    rng = np.random.default_rng(0)
    # ---- infinite-horizon gains (lqr.py:6-31)
    n, m = 10, 3
    A = np.eye(n) + 0.05 * rng.normal(size=(n, n)); B = rng.normal(size=(n, m)); Q = np.eye(n); R = 0.1 * np.eye(m)
    A2 = np.stack([A, 0.9 * A, A.T]); B2 = np.stack([B, 2.0 * B, B])
    Ls, Ps, Kd, Pd = [], [], [], []
    for a, b in zip(A2, B2):
        L_, P_ = ref.lqr.solve_riccati(a, b, Q, R); Ls.append(L_); Ps.append(P_)
        K_, P2 = ref.lqr.dare(a, b, Q, R); Kd.append(K_); Pd.append(P2)
    out.update(lqr_A=A2, lqr_B=B2, lqr_Q=Q, lqr_R=R, lqr_L=np.array(Ls), lqr_P=np.array(Ps), dare_K=np.array(Kd),
               dare_P=np.array(Pd))
    # ---- EKF on the small TPWL bank (observer.py:94-126), 15 steps, two filters
    data, Hf = synth.tpwl_bank(seed=11, r=5, m=3, P=40, num_nodes=20, tip_node=7, spread=1.0)
    Cf = meas_matrix((3, 17, 64, 90), 120)
    prm = {'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}
    mr = ref.tpwl.TPWLATV(data, params=prm, Hf=Hf, Cf=Cf, discr_method='be')
    W, V, S0 = 0.5 * np.eye(10), 0.01 * np.eye(4), 2.0 * np.eye(10)
    us = rng.uniform(0, 1000, size=(2, 15, 3)); ys = mr.y_ref + rng.normal(size=(2, 15, 4))
    xs, Ss, zs = [], [], []
    for b in range(2):
        e = ref.observer.DiscreteEKFObserver(mr, W=W, V=V, Sigma0=S0)
        xb, Sb, zb = [], [], []
        for k in range(15):
            e.update(us[b, k], ys[b, k], 0.01)
            xb.append(e.x.copy()); Sb.append(e.Sigma.copy()); zb.append(np.asarray(e.z).copy())
        xs.append(xb); Ss.append(Sb); zs.append(zb)
    out.update(ekf_Cf=Cf, ekf_W=W, ekf_V=V, ekf_S0=S0, ekf_u=us, ekf_y=ys, ekf_x=np.array(xs), ekf_Sigma=np.array(Ss),
               ekf_z=np.array(zs))
    # ---- TrajTrackingLQR (traj_tracking_lqr.py:18-48)
    tg = _T(); tg.t = np.linspace(0, 0.3, 31); tg.x = rng.normal(size=(31, 10)); tg.u = rng.uniform(0, 100, size=(31, 3))
    qc = ref.utils.QuadraticCost(Q=np.eye(10), R=0.01 * np.eye(3))
    tv = ref.traj_tracking_lqr.TrajTrackingLQR(0.01, mr, qc)
    K, P = tv.perform_dlqr_recursion(tg)
    out.update(tv_t=tg.t, tv_x=tg.x, tv_u=tg.u, tv_K=K, tv_P=P, tv_xbar=tv.x_bar, tv_ubar=tv.u_bar)
    # ---- bank construction (utils.py:251-286, tpwl_utils.py:263-276)
    r = 6
    Ks, Ds, Ms, Hs, fs, qs, As, Bs, ds = [], [], [], [], [], [], [], [], []
    for i in range(5):
        Kk = rng.normal(size=(r, r)); Kk = Kk @ Kk.T + r * np.eye(r)
        D = 0.1 * Kk + np.eye(r)
        M = np.eye(r) + 0.1 * rng.normal(size=(r, r)); M = M @ M.T
        H = rng.normal(size=(r, 2)); f = rng.normal(size=r); q = rng.normal(size=r)
        A_, B_ = ref.utils.extract_AB(Kk, D, M, H)
        d_ = np.hstack((np.linalg.solve(M, f + Kk @ q), np.zeros(r)))       # tpwl_utils.py:269-272
        for lst, v in zip((Ks, Ds, Ms, Hs, fs, qs, As, Bs, ds), (Kk, D, M, H, f, q, A_, B_, d_)):
            lst.append(v)
    out.update(bank_K=np.array(Ks), bank_D=np.array(Ds), bank_M=np.array(Ms), bank_H=np.array(Hs), bank_f=np.array(fs),
               bank_q=np.array(qs), bank_A=np.array(As), bank_B=np.array(Bs), bank_d=np.array(ds))
    # ---- GuSTO accuracy ratio (gusto.py:203-223) with the reference TPWL model's continuous dynamics
    class G:                                              # scp/models/tpwl.py:32-50, verbatim arithmetic
        def get_continuous_dynamics(self, x, u):
            A, B, d = mr.get_jacobians(x)
            return A @ x + B @ u + d, A, B
    g = G()
    N = 12
    xk = rng.normal(size=(N + 1, 10)); uk = rng.uniform(0, 500, size=(N, 3))
    x = xk + 1.5 * rng.normal(size=xk.shape); u = uk + 5.0 * rng.normal(size=uk.shape)     # far enough to change the nearest point
    fscale = rng.uniform(0.5, 2.0, size=10)
    err = 0; approx = 0; dt = 0.05; J = 3.7
    for i in range(N):
        fk, Ak, Bk = g.get_continuous_dynamics(xk[i], uk[i])
        f, _, _ = g.get_continuous_dynamics(x[i], u[i])
        fa = fk + Ak @ (x[i] - xk[i]) + Bk @ (u[i] - uk[i])
        err += dt * np.linalg.norm(np.multiply(fscale, f - fa), 2)
        approx += dt * np.linalg.norm(np.multiply(fscale, fa), 2)
    out.update(acc_xk=xk, acc_uk=uk, acc_x=x, acc_u=u, acc_fscale=fscale, acc_dt=dt, acc_J=J, acc_rho=err / (J + approx))
    np.savez_compressed(os.path.join(GOLD, "control_small.npz"), **out)
    print("control_small: riccati passes per system ->", [int(np.isfinite(l).all()) for l in Ls], "rho", out['acc_rho'])


def diamond(ref):
    """EKF at the Diamond size (n = 72, 6 measured DOFs): 8 steps of the reference class, one filter."""
    import sofacontrol_b200.synth as synth
    data, Hf = synth.tpwl_bank()
    rng = np.random.default_rng(3)
    nf = 2 * data['rom_info']['U'].shape[0]
    Cf = meas_matrix(tuple(int(v) for v in rng.choice(nf, size=6, replace=False)), nf)
    mr = ref.tpwl.TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, Cf=Cf,
                          discr_method='be')
    W, V = 1e-2 * np.eye(72), 1e-3 * np.eye(6)
    e = ref.observer.DiscreteEKFObserver(mr, W=W, V=V)
    us = rng.uniform(0, 1500, size=(8, 4)); ys = mr.y_ref + 0.5 * rng.normal(size=(8, 6))
    xs, Ss = [], []
    for k in range(8):
        e.update(us[k], ys[k], 0.01)
        xs.append(e.x.copy()); Ss.append(e.Sigma.copy())
    np.savez_compressed(os.path.join(GOLD, "control_diamond.npz"), Cf_rows=np.nonzero(Cf)[1], W=W, V=V, u=us, y=ys,
                        x=np.array(xs), Sigma=np.array(Ss))
    print("control_diamond: |x| max", np.abs(xs[-1]).max())


if __name__ == "__main__":
    warnings.simplefilter("ignore")
    from oracle import refimport
    ref = refimport.load()
    small(ref)
    diamond(ref)
