"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference, osqp stubbed)
on seeded inputs.  TEST INFRASTRUCTURE ONLY; runs only in the build container (the reference does not travel to
the GPU box -- the vectors do).

    python -m oracle.make_golden

Fixtures written
  ssm_diamond_model.npz   Diamond SSM coefficients of examples/hardware/SSMmodels/SSM_model.mat + the equilibrium
                          output z_eq = linearModel([1354],1628).evaluate(x_eq, qv=True) from rest_qv.pkl
                          (examples/hardware/diamond_SSM.py:88-102).  Model INPUT data for every SSM test/bench.
  ssm_module_test.npz     the reference's manual module_test (diamond_SSM.py:83-140): recorded inputs u_big.csv,
                          recorded outputs z_big.csv, rollouts and MSEs of the reference SSMDynamics class
                          (ssm.py UNMODIFIED, run on oracle/jax_shim.py) for be / fe / bil / discrete.
  ssm_units.npz           reference SSMDynamics on 64 seeded states: f, C, W, (A, B, d) continuous / fe / be / bil /
                          discrete, (H, c), update_state -- Diamond (m=4) and Trunk (m=8) models.
  ssm_ilqr.npz            reference iLQR class (ilqr.py, unmodified) driving the reference SSMDynamics class through
                          the Gauss-Newton H-property adapter: Diamond (m=4) and Trunk (m=8) figure-8 solves.
  ilqr_nonpd.npz          reference iLQR class on indefinite stage costs: the non-PD branch of dlqr_recursion
                          (ilqr.py:276-299, no restart) -- full solves + one backward pass.
  tpwl_small.npz          reference TPWLATV (tpwl.py, unmodified) on a small seeded bank: nearest indices, weights,
                          Jacobians (fe/be/bil/zoh), rollouts (nn + weighting), and a reference iLQR solve.
  tpwl_diamond_nn.npz     reference calc_nearest_point on the Diamond-shaped bank (P=1000, r=36): 512 states.
  pod_known.npz           Sigma of examples/diamond/pod_model.pkl + tolerance + the known answer (36 modes), and a
                          small reference compute_POD (np.linalg.svd) case.
"""
import contextlib
import io
import os
import pickle
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
GOLD = os.path.join(REPO, "tests", "golden")

from oracle import refimport, ssm_np  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def small_tpwl_bank(seed=11, r=5, m=3, P=40):
    """A small stable bank with the reference's dict schema (same construction as synth.tpwl_bank)."""
    import sofacontrol_b200.synth as synth
    return synth.tpwl_bank(seed=seed, r=r, m=m, P=P, num_nodes=20, tip_node=7, spread=1.0)


NONPD_CASES = [  # tag, m, N, amplitude, Q[2,2], max_iter
    ("d4_first_step", 4, 15, 3.0, -5000.0, 50), ("d4_five", 4, 20, 3.0, -30.0, 4), ("t8_five", 8, 40, 6.0, -30.0, 4),
    ("t8_hard", 8, 40, 3.0, -500.0, 4), ("t8_twelve", 8, 40, 6.0, -500.0, 11),
    # long solves: the indefinite cost makes these closed loops unstable (|x| ~ 1e5..1e6), a ONE-ulp change of the target
    # moves the reference's own result by 1e-6..1e-2 (stored as *_ulp_sensitivity) -- they pin the branch sequence
    ("t8_mid", 8, 40, 6.0, -50.0, 50), ("d4_mid", 4, 30, 6.0, -200.0, 50)]


def nonpd_golden(ref):
    """Drives the UNMODIFIED reference iLQR class through `Q_uu not PD` (ilqr.py:282-299).  The failing step of each
    backward sweep is recovered by wrapping (not editing) dlqr_recursion: the first all-zero K row from the top."""
    import sofacontrol_b200.synth as synth
    out = {}
    for tag, m, N, amp, q22, max_iter in NONPD_CASES:
        s = synth.trunk_ssm(m)
        mdl = ssm_np.GaussNewtonSSM(ssm_np.SSMDynamicsNP(s['z_ref'], discrete=False, discr_method='be', model=s['model'],
                                                         params=s['params']))
        Q, R, Qf = synth.trunk_ilqr_costs(6, m)
        Q = Q.copy(); Q[2, 2] = q22
        sol = ref.ilqr.iLQR(0.02, mdl, ref.utils.QuadraticCost(Q, R, Qf), N)
        sol.params.max_iter = max_iter
        sol.set_target(synth.figure8_targets(s['z_ref'], N, amp)[0])
        fails, calls = [], []
        inner = sol.dlqr_recursion

        def wrapped(x, u, A, B, d, inner=inner):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                K, k, Qu, Quu = inner(x, u, A, B, d)
            tf = -1
            if 'not PD' in buf.getvalue():
                # the sweep stopped at t_fail: Q_uu[t_fail] was assigned (non-zero, it contains R), Q_uu below is zero
                tf = min(t for t in range(N) if Quu[t].any())
            fails.append(tf)
            calls.append((x.copy(), u.copy(), A.copy(), B.copy(), d.copy(), K.copy(), k.copy(), Qu.copy(), Quu.copy()))
            return K, k, Qu, Quu
        sol.dlqr_recursion = wrapped
        x, u, K = quiet(sol.ilqr_computation, np.zeros(6))
        out.update({tag + '_x': x, tag + '_u': u, tag + '_K': K, tag + '_rho': sol.rho, tag + '_iterations': len(fails),
                    tag + '_pd_fail_t': np.array(fails)})
        # conditioning of the case: the same reference solve with the target moved by ONE ulp.  An indefinite stage
        # cost makes some of these closed loops unstable (|x| grows to 1e5..1e6 along the horizon), and rounding-level
        # input changes are then amplified over the iterations; the GPU test allows 10 x this measured sensitivity.
        sol2 = ref.ilqr.iLQR(0.02, mdl, ref.utils.QuadraticCost(Q, R, Qf), N)
        sol2.params.max_iter = max_iter
        sol2.set_target(synth.figure8_targets(s['z_ref'], N, amp)[0] * (1.0 + 2.0 ** -52))
        with np.errstate(all='ignore'):
            x2, u2, K2 = quiet(sol2.ilqr_computation, np.zeros(6))
        rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
        out[tag + '_ulp_sensitivity'] = np.array([rel(x2, x), rel(u2, u), rel(K2, K)])
        print("nonpd", tag, "1-ulp sensitivity of the reference solve (x, u, K):", out[tag + '_ulp_sensitivity'])
        print("nonpd", tag, "iterations", len(fails), "pd_fail_t", fails[:8], "rho", sol.rho)
        if tag == "t8_twelve":
            # unit backward pass: replay the first interrupted sweep from rho = drho = 0
            i = next(j for j, f in enumerate(fails) if f >= 0)
            xx, uu, AA, BB, dd = calls[i][:5]
            sol.dlqr_recursion = inner
            sol.rho, sol.drho = 0.0, 0.0
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                Kb, kb, Qub, Quub = sol.dlqr_recursion(xx, uu, AA, BB, dd)
            assert 'not PD' in buf.getvalue()
            out.update(unit_x=xx, unit_u=uu, unit_A=AA, unit_B=BB, unit_d=dd, unit_K=Kb, unit_k=kb, unit_Qu=Qub,
                       unit_Quu=Quub, unit_rho=sol.rho, unit_drho=sol.drho,
                       unit_pd_fail_t=min(t for t in range(N) if Quub[t].any()))
    return out


def main():
    ref = refimport.load()
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == 'nonpd':
        np.savez_compressed(os.path.join(GOLD, "ilqr_nonpd.npz"), **nonpd_golden(ref))
        return
    only_ssm = len(sys.argv) > 1 and sys.argv[1] == 'ssm'
    from scipy.io import loadmat
    from scipy.interpolate import interp1d
    import sofacontrol_b200.synth as synth

    # ---------------------------------------------------------------- SSM model + equilibrium
    hw = os.path.join(ref.root, "examples", "hardware")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rest = ref.utils.load_data(os.path.join(hw, "rest_qv.pkl"))
    qv = np.array(rest['rest'])
    x_eq = ref.utils.qv2x(q=qv[0], v=qv[1])
    z_eq = ref.measurement_models.linearModel([1354], 1628).evaluate(x_eq, qv=True)
    mat = loadmat(os.path.join(hw, "SSMmodels", "SSM_model.mat"))['py_data'][0, 0]
    mdl, prm = mat['model'], mat['params']
    g = lambda k: np.asarray(mdl[k][0, 0], dtype=np.float64)
    np.savez(os.path.join(GOLD, "ssm_diamond_model.npz"), r_coeff=g('r_coeff'), w_coeff=g('w_coeff'),
             v_coeff=g('v_coeff'), rd_coeff=g('rd_coeff'), B=g('B'), Bd=g('Bd'), Ts=float(mdl['Ts'][0, 0][0, 0]),
             z_eq=np.asarray(z_eq, dtype=np.float64),
             dims=np.array([int(prm[k][0, 0][0, 0]) for k in ('state_dim', 'input_dim', 'output_dim', 'SSM_order', 'ROM_order')]))

    # ---------------------------------------------------------------- module_test (diamond_SSM.py:83-140)
    z_true = np.genfromtxt(os.path.join(hw, "checkModel", "z_big.csv"), delimiter=',')
    u_true = np.genfromtxt(os.path.join(hw, "checkModel", "u_big.csv"), delimiter=',')
    zq, zv = ref.utils.x2qv(z_true)
    dt = 0.01
    T = 10.01
    N = int(T / dt)
    t_orig = np.linspace(0, T, int(T / 0.01) + 1)
    t_int = np.linspace(0, T, N + 1)
    u_int = interp1d(t_orig, u_true, axis=0)(t_int)
    z_qv = interp1d(t_orig, np.hstack((zq, zv)), axis=0)(t_int)
    out = dict(u=u_int, z_true_qv=z_qv, dt=dt)
    rssm = refimport.load_ssm()        # sofacontrol/SSM/ssm.py UNMODIFIED, on the jax stand-in of oracle/jax_shim.py
    for name, kw in (('be', dict(discrete=False, discr_method='be')), ('fe', dict(discrete=False, discr_method='fe')),
                     ('bil', dict(discrete=False, discr_method='bil')), ('disc', dict(discrete=True, discr_method='be'))):
        m = rssm.SSMDynamics(z_eq, model=mdl, params=prm, **kw)
        x, z = m.rollout(np.zeros(6), u_int, dt)
        err = z_qv - z[:-1]
        out['x_' + name] = x
        out['z_' + name] = z
        out['mse_' + name] = np.linalg.norm(np.linalg.norm(err, axis=1)) ** 2 / err.shape[0]
        print("module_test", name, out['mse_' + name])
    np.savez_compressed(os.path.join(GOLD, "ssm_module_test.npz"), **out)

    # ---------------------------------------------------------------- unit evaluations of the reference SSM class
    units = {}
    for tag, mm in (('diamond', 4), ('trunk', 8)):
        sm = synth.trunk_ssm(mm)
        rng = np.random.default_rng(100 + mm)
        xs = rng.normal(size=(64, 6)) * np.array([3.0, 3.0, 3.0, 30.0, 30.0, 30.0])
        us = rng.uniform(0, 800, size=(64, mm))
        zs = sm['z_ref'] + rng.normal(size=(64, 6)) * np.array([3.0, 3.0, 3.0, 20.0, 20.0, 20.0])
        units[tag + '_x'], units[tag + '_u'], units[tag + '_z'] = xs, us, zs
        c = rssm.SSMDynamics(sm['z_ref'], discrete=False, discr_method='fe', model=sm['model'], params=sm['params'])
        units[tag + '_f'] = np.array([c.reduced_dynamics(x, u) for x, u in zip(xs, us)])
        units[tag + '_C'] = np.array([c.C_map(x) for x in xs])
        units[tag + '_zf'] = np.asarray(c.x_to_zfyf(xs))
        units[tag + '_W'] = np.array([c.compute_RO_state(z) for z in zs])
        J = [c.get_continuous_jacobians(x, u) for x, u in zip(xs, us)]
        for i, nm in enumerate('ABd'):
            units[tag + '_c' + nm] = np.array([np.asarray(j[i]) for j in J])
        O = [c.get_observer_jacobians(x) for x in xs]
        units[tag + '_H'] = np.array([np.asarray(o[0]) for o in O])
        units[tag + '_c'] = np.array([np.asarray(o[1]) for o in O])
        for meth, kw in (('fe', dict(discrete=False, discr_method='fe')), ('be', dict(discrete=False, discr_method='be')),
                         ('bil', dict(discrete=False, discr_method='bil')), ('disc', dict(discrete=True, discr_method='be'))):
            mdl_ = rssm.SSMDynamics(sm['z_ref'], model=sm['model'], params=sm['params'], **kw)
            J = [mdl_.get_jacobians(x, u, 0.02) for x, u in zip(xs, us)]
            for i, nm in enumerate('ABd'):
                units['%s_%s_%s' % (tag, meth, nm)] = np.array([np.asarray(j[i]) for j in J])
            units['%s_%s_next' % (tag, meth)] = np.array([mdl_.update_state(x, u, 0.02) for x, u in zip(xs, us)])
    np.savez_compressed(os.path.join(GOLD, "ssm_units.npz"), **units)

    # ---------------------------------------------------------------- SSM iLQR through the unmodified reference class
    res = {}
    for tag, mm in (('diamond', 4), ('trunk', 8)):
        s = synth.trunk_ssm(mm)
        m = rssm.SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
        Nh = 100
        zt = synth.figure8_targets(s['z_ref'], Nh, 5.0)[0]
        Q, R, Qf = synth.trunk_ilqr_costs(6, mm)
        solver = ref.ilqr.iLQR(0.02, ssm_np.GaussNewtonSSM(m), ref.utils.QuadraticCost(Q, R, Qf), Nh)
        solver.set_target(zt)
        x, u, K = quiet(solver.ilqr_computation, np.zeros(6))
        res.update({tag + '_x': x, tag + '_u': u, tag + '_K': K, tag + '_zt': zt, tag + '_rho': solver.rho})
        # one forward pass + one backward pass of the reference on the converged trajectory (unit-level parity)
        xf, uf, cf, Af, Bf, df = solver.forward_pass(x, u)
        solver.rho, solver.drho = 0.0, 0.0
        Kb, kb, Qub, Quub = quiet(solver.dlqr_recursion, xf, uf, Af, Bf, df)
        res.update({tag + '_fp_cost': cf, tag + '_fp_A': Af, tag + '_fp_B': Bf, tag + '_fp_d': df, tag + '_bp_K': Kb,
                    tag + '_bp_k': kb, tag + '_bp_Qu': Qub, tag + '_bp_Quu': Quub, tag + '_bp_rho': solver.rho})
        # and on the initial zero-input rollout, far from the optimum: k, Q_u are O(1) there (no cancellation)
        x00 = np.zeros((Nh + 1, 6))
        x0f, u0f, c0f, A0f, B0f, d0f = solver.forward_pass(x00, np.zeros((Nh, mm)))
        solver.rho, solver.drho = 0.0, 0.0
        K0, k0, Qu0, Quu0 = quiet(solver.dlqr_recursion, x0f, u0f, A0f, B0f, d0f)
        res.update({tag + '_init_x': x0f, tag + '_init_u': u0f, tag + '_init_cost': c0f, tag + '_init_A': A0f,
                    tag + '_init_B': B0f, tag + '_init_d': d0f, tag + '_init_K': K0, tag + '_init_k': k0,
                    tag + '_init_Qu': Qu0, tag + '_init_Quu': Quu0})
        print("ssm ilqr", tag, "cost", cf)
    np.savez_compressed(os.path.join(GOLD, "ssm_ilqr.npz"), **res)

    if only_ssm:
        return
    # ---------------------------------------------------------------- non-PD branch of the reference class
    np.savez_compressed(os.path.join(GOLD, "ilqr_nonpd.npz"), **nonpd_golden(ref))

    # ---------------------------------------------------------------- small TPWL bank through the reference class
    data, Hf = small_tpwl_bank()
    rng = np.random.default_rng(12)
    r, mI, P = 5, 3, 40
    n = 2 * r
    xs = np.concatenate((rng.normal(0, 1, size=(64, r)), rng.normal(0, 1.0, size=(64, r))), axis=1)
    xs[3] = ref.utils.qv2x(data['q'][17], data['v'][17])        # exact hit: min distance 0 (one-hot weights)
    x0 = xs[0]
    Nr = 60
    useq = rng.uniform(0, 1500, size=(Nr, mI))
    out = dict(xs=xs, useq=useq, dt=0.01)
    dw = {'q': 1.0, 'v': 0.0}
    dw2 = {'q': 0.7, 'v': 0.05}
    for wname, w in (('w10', dw), ('w7', dw2)):
        mdl_nn = ref.tpwl.TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': w}, Hf=Hf, discr_method='fe')
        out['idx_' + wname] = np.array([mdl_nn.calc_nearest_point(x) for x in xs])
    mdl_w = ref.tpwl.TPWLATV(data, params={'tpwl_method': 'weighting', 'dist_weights': dw2, 'beta_weighting': 25.0},
                             Hf=Hf, discr_method='fe')
    out['weights'] = np.array([mdl_w.calc_weighting_factors(x) for x in xs])
    for meth in ('fe', 'be', 'bil', 'zoh'):
        mn = ref.tpwl.TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': dw}, Hf=Hf, discr_method=meth)
        J = [mn.get_jacobians(x, dt=0.01) for x in xs[:8]]
        out['nn_A_' + meth] = np.array([j[0] for j in J])
        out['nn_B_' + meth] = np.array([j[1] for j in J])
        out['nn_d_' + meth] = np.array([j[2] for j in J])
        xr, zr = mn.rollout(x0, useq, 0.01)
        out['nn_x_' + meth], out['nn_z_' + meth] = xr, zr
        mw = ref.tpwl.TPWLATV(data, params={'tpwl_method': 'weighting', 'dist_weights': dw2, 'beta_weighting': 25.0},
                              Hf=Hf, discr_method=meth)
        J = [mw.get_jacobians(x, dt=0.01) for x in xs[:8]]
        out['w_A_' + meth] = np.array([j[0] for j in J])
        out['w_B_' + meth] = np.array([j[1] for j in J])
        out['w_d_' + meth] = np.array([j[2] for j in J])
        xr, zr = mw.rollout(x0, useq, 0.01)
        out['w_x_' + meth], out['w_z_' + meth] = xr, zr
    # pre-discretised zoh bank (the reference default configuration) for the nn rollout
    mz = ref.tpwl.TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': dw}, Hf=Hf, discr_method='zoh')
    quiet(mz.pre_discretize, 0.01)
    out['zoh_A_d'], out['zoh_B_d'], out['zoh_d_d'] = np.array(mz.A_d), np.array(mz.B_d), np.array(mz.d_d)
    # reference iLQR on the small TPWL model (constant H)
    Nh = 40
    Q = np.zeros((6, 6)); Q[3, 3] = Q[4, 4] = 100.0; Q[5, 5] = 10.0
    R = 1e-5 * np.eye(mI)
    zt = np.tile(mz.z_ref, (Nh + 1, 1))
    th = np.linspace(0, 2 * np.pi, Nh + 1)
    zt[:, 3] += 0.05 * np.sin(th); zt[:, 4] += 0.05 * np.sin(2 * th)
    solver = ref.ilqr.iLQR(0.01, mz, ref.utils.QuadraticCost(Q, R, np.zeros((6, 6))), Nh)
    solver.set_target(zt)
    x, u, K = quiet(solver.ilqr_computation, x0)
    out.update(ilqr_x=x, ilqr_u=u, ilqr_K=K, ilqr_zt=zt, ilqr_Q=Q, ilqr_R=R, ilqr_rho=solver.rho)
    np.savez_compressed(os.path.join(GOLD, "tpwl_small.npz"), **out)

    # ---------------------------------------------------------------- Diamond-shaped nearest point (P=1000, r=36)
    data, Hf = synth.tpwl_bank()
    x0b, _ = synth.tpwl_rollout_batch(512, N=1)
    mdl = ref.tpwl.TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, discr_method='fe')
    idx = np.array([mdl.calc_nearest_point(x) for x in x0b])
    mdl2 = ref.tpwl.TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.3}}, discr_method='fe')
    idx2 = np.array([mdl2.calc_nearest_point(x) for x in x0b])
    xr, _ = mdl.rollout(x0b[0], synth.tpwl_rollout_batch(1, N=100)[1][0], 0.01)
    np.savez_compressed(os.path.join(GOLD, "tpwl_diamond_nn.npz"), idx_q=idx, idx_qv=idx2, rollout_x=xr)
    print("diamond nn: distinct indices", len(set(idx.tolist())))

    # ---------------------------------------------------------------- POD known answers
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pm = pickle.load(open(os.path.join(ref.root, "examples", "diamond", "pod_model.pkl"), "rb"))
    Sigma = np.asarray(pm['Sigma'])
    tol = pm['config']['pod_tolerance']
    s2 = Sigma ** 2
    i = 0
    while (np.sum(s2[i:]) / np.sum(s2)) > tol or i == 0:     # pod.py:193-196
        i += 1
    X, _, _ = synth.pod_snapshots(600, 150, seed=5)
    Uf, U, nb, S = ref.pod.compute_POD(X, 5e-5)
    np.savez_compressed(os.path.join(GOLD, "pod_known.npz"), Sigma=Sigma, tol=tol, modes=i,
                        U_orth_err=np.abs(pm['POD_info']['U'].T @ pm['POD_info']['U'] - np.eye(36)).max(),
                        small_U=U, small_S=S, small_modes=nb)
    print("pod: fixture modes", i, "small case modes", nb)


if __name__ == "__main__":
    main()
