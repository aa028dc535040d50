"""GPU parity of the kernels for the callers either side of the hot path (csrc/control.cu):
batched EKF, infinite-horizon and time-varying LQR gains, TPWL bank construction, the GuSTO linearisation adapters
and model-accuracy ratio, the receding-horizon glue.  Golden vectors come from the UNMODIFIED reference modules
(oracle/make_golden_control.py); tolerance relative 1e-9."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _small_model(Cf=None, method='nn', discr='be', beta=None):
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.tpwl.tpwl import TPWLATV
    data, Hf = synth.tpwl_bank(seed=11, r=5, m=3, P=40, num_nodes=20, tip_node=7, spread=1.0)
    prm = {'tpwl_method': method, 'dist_weights': {'q': 1.0, 'v': 0.0}, 'beta_weighting': beta}
    return TPWLATV(data, params=prm, Hf=Hf, Cf=Cf, discr_method=discr)


def test_ekf_small_vs_reference_golden(golden):
    """DiscreteEKFObserver: 15 predict/update steps, single filter (reference shapes) and both filters as a batch."""
    from sofacontrol_b200.tpwl.observer import DiscreteEKFObserver
    g = golden("control_small.npz")
    m = _small_model(Cf=g['ekf_Cf'])
    kw = dict(W=g['ekf_W'], V=g['ekf_V'], Sigma0=g['ekf_S0'])
    e = DiscreteEKFObserver(m, **kw)
    assert e.x.shape == (10,) and e.Sigma.shape == (10, 10)
    for k in range(15):
        e.update(g['ekf_u'][0, k], g['ekf_y'][0, k], 0.01)
        assert relerr(e.x, g['ekf_x'][0, k]) < TOL and relerr(e.Sigma, g['ekf_Sigma'][0, k]) < TOL
    assert relerr(e.z, g['ekf_z'][0, 14]) < TOL
    eb = DiscreteEKFObserver(m, **kw)
    eb.initialize_reduced(np.tile(eb.x, (2, 1)))
    for k in range(15):
        eb.update(g['ekf_u'][:, k], g['ekf_y'][:, k], 0.01)
    assert relerr(eb.x, g['ekf_x'][:, 14]) < TOL and relerr(eb.Sigma, g['ekf_Sigma'][:, 14]) < TOL
    with pytest.raises(RuntimeError):
        DiscreteEKFObserver(_small_model())                # no measurement model (observer.py:52-53)


def test_ekf_diamond_size_vs_reference_golden(golden):
    """n = 72, 6 measured DOFs, P = 1000: the covariance products run on the FP64 tensor pipe."""
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.tpwl.tpwl import TPWLATV
    from sofacontrol_b200.tpwl.observer import DiscreteEKFObserver
    g = golden("control_diamond.npz")
    data, Hf = synth.tpwl_bank()
    nf = 2 * data['rom_info']['U'].shape[0]
    Cf = np.zeros((6, nf)); Cf[np.arange(6), g['Cf_rows']] = 1.0
    m = TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, Cf=Cf, discr_method='be')
    e = DiscreteEKFObserver(m, W=g['W'], V=g['V'])
    for k in range(8):
        e.update(g['u'][k], g['y'][k], 0.01)
        assert relerr(e.x, g['x'][k]) < TOL and relerr(e.Sigma, g['Sigma'][k]) < TOL


def test_full_state_and_ssm_observers():
    from sofacontrol_b200.tpwl.observer import FullStateObserver
    from sofacontrol_b200.SSM.observer import SSMObserver
    rng = np.random.default_rng(0)
    H = rng.normal(size=(3, 6)); x = rng.normal(size=6)
    o = FullStateObserver(6, H=H)
    o.update(None, None, 0.01, x=x)
    assert np.array_equal(o.z, H @ x) and o.get_meas_dim() == 6
    s = SSMObserver(None)
    y = rng.normal(size=6)
    s.update(None, y, 0.01)
    assert np.array_equal(s.z, np.hstack((y[3:], y[:3])))            # vq2qv (SSM/controllers.py:302-309)


def test_riccati_gains_vs_reference_golden(golden):
    from sofacontrol_b200.lqr.lqr import solve_riccati, solve_riccati_info, dare, DLQR
    from oracle import lqr_np
    g = golden("control_small.npz")
    L_, P_, it = solve_riccati_info(g['lqr_A'], g['lqr_B'], g['lqr_Q'], g['lqr_R'])
    for i in range(3):
        lqr_np.solve_riccati(g['lqr_A'][i], g['lqr_B'][i], g['lqr_Q'], g['lqr_R'])
        assert it[i] == lqr_np.solve_riccati.last_iterations             # same stopping pass as the reference loop
    assert relerr(L_, g['lqr_L']) < TOL and relerr(P_, g['lqr_P']) < TOL
    L1, P1 = solve_riccati(g['lqr_A'][0], g['lqr_B'][0], g['lqr_Q'], g['lqr_R'])
    assert L1.shape == (3, 10) and relerr(L1, g['lqr_L'][0]) < TOL
    K, P = dare(g['lqr_A'], g['lqr_B'], g['lqr_Q'], g['lqr_R'])          # scipy solve_discrete_are in the reference
    assert relerr(K, g['dare_K']) < TOL and relerr(P, g['dare_P']) < TOL
    # DLQR.compute_gain_matrix (lqr.py:52-55): discretise the target linearisation, then solve_riccati
    m = _small_model(discr='be')
    A_c, B_c = np.asarray(m.tpwl_dict['A_c'][3]), np.asarray(m.tpwl_dict['B_c'][3])
    from oracle.tpwl_np import TPWLATVNP
    o = TPWLATVNP(m.tpwl_dict, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, discr_method='be')
    Ad, Bd, _ = o.discretize_dynamics(A_c, B_c, np.zeros(10), 0.01)
    from sofacontrol_b200.utils import QuadraticCost
    Kg = DLQR(0.01, m, QuadraticCost(Q=np.eye(10), R=0.1 * np.eye(3))).compute_gain_matrix(A_c, B_c, np.eye(10), 0.1 * np.eye(3))
    assert relerr(Kg, lqr_np.solve_riccati(Ad, Bd, np.eye(10), 0.1 * np.eye(3))[0]) < TOL


def test_traj_tracking_lqr_vs_reference_golden(golden):
    from sofacontrol_b200.lqr.traj_tracking_lqr import TrajTrackingLQR
    from sofacontrol_b200.utils import QuadraticCost
    g = golden("control_small.npz")
    m = _small_model(discr='be')

    class T:
        pass
    tg = T(); tg.t, tg.x, tg.u = g['tv_t'], g['tv_x'], g['tv_u']
    tv = TrajTrackingLQR(0.01, m, QuadraticCost(Q=np.eye(10), R=0.01 * np.eye(3)))
    xb, ub, K = tv.compute_policy(tg)
    K2, P = tv.perform_dlqr_recursion(tg)
    assert K.shape == g['tv_K'].shape and P.shape == g['tv_P'].shape
    assert relerr(K, g['tv_K']) < TOL and relerr(P, g['tv_P']) < TOL
    assert relerr(xb, g['tv_xbar']) < 1e-14 and relerr(ub, g['tv_ubar']) < 1e-14
    tg.x, tg.u = np.stack([g['tv_x'], g['tv_x'][::-1]]), np.stack([g['tv_u'], g['tv_u']])     # two trajectories at once
    Kb, Pb = tv.perform_dlqr_recursion(tg)
    assert Kb.shape == (2,) + g['tv_K'].shape and relerr(Kb[0], g['tv_K']) < TOL


def test_bank_construction_vs_reference_golden(golden):
    """extract_AB + add_continuous_TPWL arithmetic, batched; TPWLSnapshotData.add_point through the POD projection."""
    from sofacontrol_b200.utils import extract_AB
    from sofacontrol_b200.tpwl.tpwl_utils import continuous_tpwl_points, TPWLSnapshotData, SnapshotPoint
    from sofacontrol_b200.mor.pod import POD
    from oracle import lqr_np
    g = golden("control_small.npz")
    A, B = extract_AB(g['bank_K'], g['bank_D'], g['bank_M'], g['bank_H'])
    assert relerr(A, g['bank_A']) < TOL and relerr(B, g['bank_B']) < TOL
    A1, B1 = extract_AB(g['bank_K'][2], g['bank_D'][2], g['bank_M'][2], g['bank_H'][2])
    assert A1.shape == (12, 12) and np.array_equal(A1, A[2])
    A, B, d = continuous_tpwl_points(g['bank_K'], g['bank_D'], g['bank_M'], g['bank_H'], g['bank_f'], g['bank_q'])
    assert relerr(A, g['bank_A']) < TOL and relerr(d, g['bank_d']) < TOL
    # add_point: full-order snapshot (nf = 30) -> POD projection (r = 6) -> continuous TPWL entry, vs the oracle chain
    rng = np.random.default_rng(8)
    nf, r = 30, 6
    U, _ = np.linalg.qr(rng.normal(size=(nf, r)))
    rom = POD({'U': U, 'q_ref': rng.normal(size=nf), 'v_ref': np.zeros(nf), 'type': 'POD'})
    Kf = rng.normal(size=(nf, nf)); Kf = Kf @ Kf.T + nf * np.eye(nf)
    Mf = np.diag(rng.uniform(1, 2, size=nf)); Df = 0.1 * Kf + 2.0 * Mf
    Hfm, ff = rng.normal(size=(nf, 2)), rng.normal(size=nf)
    pt = SnapshotPoint(t=0.1, q=rng.normal(size=nf), v=rng.normal(size=nf), u=np.array([1.0, 2.0]), K=Kf, D=Df, M=Mf,
                       H=Hfm, f=ff, dt=0.01)
    snap = TPWLSnapshotData(rom)
    snap.add_point(pt)
    qr = U.T @ (pt.q - rom.q_ref)
    Ao, Bo, do = lqr_np.continuous_tpwl_point(U.T @ Kf @ U, U.T @ Df @ U, U.T @ Mf @ U, U.T @ Hfm, U.T @ ff, qr)
    assert relerr(snap.dict['A_c'][0], Ao) < TOL and relerr(snap.dict['B_c'][0], Bo) < TOL and relerr(snap.dict['d_c'][0], do) < TOL
    assert relerr(snap.dict['q'][0], qr) < TOL and snap.dict['dt'] == 0.01


def test_gusto_adapters_and_accuracy_vs_reference_golden(golden):
    """scp/models/{tpwl,ssm}.py adapters + gusto.py:203-281 (get_traj_dynamics, get_observer_linearizations,
    compute_accuracy) as batched launches."""
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.scp.models import TPWLGuSTO, SSMGuSTO
    from sofacontrol_b200.scp.linearize import TrajectoryLinearizer
    from sofacontrol_b200.SSM.ssm import SSMDynamics
    from oracle.ssm_np import SSMDynamicsNP
    from oracle.tpwl_np import TPWLATVNP
    g = golden("control_small.npz")
    m = _small_model(discr='be')
    gm = TPWLGuSTO(m)
    assert (gm.n_x, gm.n_u, gm.n_z) == (10, 3, 6) and not gm.nonlinear_observer
    lin = TrajectoryLinearizer(gm, float(g['acc_dt']), f_scale=g['acc_fscale'])
    lin.set_iterate(g['acc_xk'], g['acc_uk'])
    rho = lin.compute_accuracy(g['acc_x'], g['acc_u'], float(g['acc_J']))
    assert abs(rho - float(g['acc_rho'])) < TOL * float(g['acc_rho'])
    rb = lin.__class__(gm, float(g['acc_dt']), f_scale=g['acc_fscale'])
    rb.set_iterate(np.stack([g['acc_xk']] * 3), np.stack([g['acc_uk']] * 3))
    rr = rb.compute_accuracy(np.stack([g['acc_x'], g['acc_xk'], g['acc_x']]), np.stack([g['acc_u'], g['acc_uk'], g['acc_u']]),
                             float(g['acc_J']))
    assert rr.shape == (3,) and abs(rr[0] - rho) < 1e-15 and rr[1] == 0.0
    # get_traj_dynamics vs the oracle model point by point
    o = TPWLATVNP(m.tpwl_dict, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=None, discr_method='be')
    A, B, d = lin.get_traj_dynamics(g['acc_x'], g['acc_u'])
    assert A.shape == (12, 10, 10) and d.shape == (12, 10)
    for i in range(12):
        Ao, Bo, do = o.get_jacobians(g['acc_x'][i], dt=float(g['acc_dt']))
        assert relerr(A[i], Ao) < TOL and relerr(B[i], Bo) < TOL and relerr(d[i], do) < TOL
    xc, fc = gm.get_characteristic_vals()
    fo = np.array([o.get_jacobians(x)[0] @ x + o.get_jacobians(x)[1] @ u + o.get_jacobians(x)[2]
                   for x, u in zip(np.concatenate((m.tpwl_dict['v'], m.tpwl_dict['q']), axis=1), m.tpwl_dict['u'])])
    assert relerr(fc, np.abs(fo).max(axis=0)) < TOL
    # SSM adapter: observer linearisations along a trajectory
    s = synth.trunk_ssm(4)
    kw = dict(discrete=False, discr_method='be', model=s['model'], params=s['params'])
    sg = SSMGuSTO(SSMDynamics(s['z_ref'], **kw))
    so = SSMDynamicsNP(s['z_ref'], **kw)
    rng = np.random.default_rng(1)
    xs, us = rng.normal(size=(9, 6)), rng.uniform(0, 800, size=(8, 4))
    ls = TrajectoryLinearizer(sg, 0.02)
    H, c = ls.get_observer_linearizations(xs)
    A, B, d = ls.get_traj_dynamics(xs, us)
    f, Ac, Bc = sg.get_continuous_dynamics(xs[:-1], us)
    for i in range(8):
        Ho, co = so.get_observer_jacobians(xs[i])
        Ao, Bo, do = so.get_jacobians(xs[i], us[i], 0.02)
        assert relerr(H[i], Ho) < TOL and relerr(c[i], co) < TOL and relerr(A[i], Ao) < TOL and relerr(B[i], Bo) < TOL
        assert relerr(f[i], so.reduced_dynamics(xs[i], us[i])) < TOL
    assert H.shape == (9, 6, 6) and sg.nonlinear_observer


def test_mpc_shift_kernel():
    import torch
    from sofacontrol_b200 import _lib as L
    rng = np.random.default_rng(0)
    Bt, N, m, nz, T, k = 5, 7, 3, 4, 11, 6
    up, zr = rng.normal(size=(Bt, N, m)), rng.normal(size=(Bt, T + N + 1, nz))
    upd, zrd = L.to_dev(up), L.to_dev(zr)
    uw, ua, zw, ul = L.empty((Bt, N, m)), L.empty((Bt, m)), L.empty((Bt, N + 1, nz)), L.zeros((Bt, T, m))
    L.check(L.lib().srcb200_mpc_shift_batch(Bt, N, m, nz, T, k, L.ptr(upd), L.ptr(zrd), L.ptr(uw), L.ptr(ua), L.ptr(zw),
                                            L.ptr(ul), L.stream_ptr()))
    assert np.array_equal(L.to_host(uw), np.concatenate((up[:, 1:], up[:, -1:]), axis=1))
    assert np.array_equal(L.to_host(ua), up[:, 0]) and np.array_equal(L.to_host(ul)[:, k], up[:, 0])
    assert np.array_equal(L.to_host(zw), zr[:, k + 1:k + 2 + N])


def test_closed_loop_tpwl_plant_with_ekf_matches_oracle_loop(golden):
    """Config-4 loop as TemplateController.evaluate runs it (tpwl/controllers.py:85-117): TPWL plant, EKF belief from
    y = C x + y_ref, receding-horizon iLQR (N = 8) with shifted warm start and u_last -- vs the same loop written
    with the oracle classes (each pinned bitwise to the reference), noise-free, 2 problems x 5 control steps."""
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost
    from sofacontrol_b200.mpc import RecedingHorizonILQR, EKFBelief
    from sofacontrol_b200.tpwl.observer import DiscreteEKFObserver
    from oracle.tpwl_np import TPWLATVNP
    from oracle.ilqr_np import ILQRNP
    from oracle.observer_np import DiscreteEKFObserverNP
    from oracle.utils_np import QuadraticCost as QCo
    g = golden("control_small.npz")
    Cf = g['ekf_Cf']
    m = _small_model(Cf=Cf, discr='be')
    N, steps, Bt, dt = 8, 5, 2, 0.01
    Q = np.zeros((6, 6)); Q[3, 3] = Q[4, 4] = 100.0; Q[5, 5] = 10.0
    R = 1e-5 * np.eye(3)
    rng = np.random.default_rng(21)
    x0 = 0.2 * rng.normal(size=(Bt, 10))
    th = np.linspace(0, 1.5, steps + N + 1)
    zref = np.tile(m.z_ref, (Bt, steps + N + 1, 1))
    zref[:, :, 3] += np.array([0.05, 0.08])[:, None] * np.sin(th)[None]
    kw = dict(W=g['ekf_W'], V=g['ekf_V'], Sigma0=g['ekf_S0'])
    sol = iLQR(dt, m, QuadraticCost(Q, R, np.zeros((6, 6))), N)
    sol.set_target(zref[:, :N + 1])
    out = RecedingHorizonILQR(sol, observer=EKFBelief(DiscreteEKFObserver(m, **kw))).run(x0, zref, steps)
    om = TPWLATVNP(m.tpwl_dict, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=None, Cf=Cf,
                   discr_method='be')
    om.H, om.z_ref = m.H, m.z_ref
    for b in range(Bt):
        ekf = DiscreteEKFObserverNP(om, **kw)
        ekf.x = x0[b].copy()
        xp, belief, u_plan, u_last = x0[b].copy(), x0[b].copy(), None, np.zeros(3)
        for k in range(steps):
            o = ILQRNP(dt, om, QCo(Q, R, np.zeros((6, 6))), N)
            o.set_target(zref[b, k:k + N + 1]); o.set_u_last(u_last)
            ws = None if u_plan is None else np.vstack((u_plan[1:], u_plan[-1:]))
            _, u_plan, _ = o.ilqr_computation(belief, ws)
            assert out['iterations'][b, k] == o.iterations
            assert relerr(out['u'][b, k], u_plan[0]) < TOL
            xp = om.update_state(xp, u_plan[0], dt)
            ekf.update(u_plan[0], om.C @ xp + om.y_ref, dt)
            belief, u_last = ekf.x.copy(), u_plan[0]
            assert relerr(out['x'][b, k + 1], xp) < TOL


def test_closed_loop_tpwl_plant_ssm_controller_runs():
    """Config 4 'Diamond SSM + TPWL': TPWL plant (Diamond size), Diamond-SSM iLQR controller fed by the SSM observer
    (tip output [v; q] -> [q; v] -> W_map).  Smoke of the plumbing: shapes, finite results, statuses set."""
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.tpwl.tpwl import TPWLATV
    from sofacontrol_b200.SSM.ssm import SSMDynamics
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost
    from sofacontrol_b200.mpc import RecedingHorizonILQR, SSMOutputBelief
    w = synth.mpc_ssm_tpwl_workload(16, steps=4, N=10, seed=4)
    out = RecedingHorizonILQR(w['solver'], plant=w['plant'], observer=SSMOutputBelief(w['ssm']), process_noise_std=1e-4) \
        .run(w['x0_belief'], w['z_ref'], 4, x0_plant=w['x0_plant'])
    assert out['x'].shape == (16, 5, 72) and out['u'].shape == (16, 4, 4)
    assert np.all(np.isfinite(out['u'])) and np.all(out['status'] != 0) and np.all(out['iterations'] >= 1)
