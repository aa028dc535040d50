"""Pins the numpy oracle (oracle/*.py) bit-for-bit against the UNMODIFIED reference modules imported from
/root/reference (available only in the build container; skipped elsewhere -- the committed golden vectors that the
same reference produced are checked in tests/test_oracle_golden.py on every box)."""
import contextlib
import io

import numpy as np
import pytest

from oracle import refimport

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference tree not present on this box")


@pytest.fixture(scope="module")
def ref():
    return refimport.load()


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _small_bank():
    import sofacontrol_b200.synth as synth
    return synth.tpwl_bank(seed=11, r=5, m=3, P=40, num_nodes=20, tip_node=7, spread=1.0)


@pytest.mark.parametrize("method", ["nn", "weighting"])
@pytest.mark.parametrize("discr", ["fe", "be", "bil", "zoh"])
def test_tpwl_restatement_bitwise(ref, method, discr):
    from oracle.tpwl_np import TPWLATVNP
    data, Hf = _small_bank()
    params = {'tpwl_method': method, 'dist_weights': {'q': 0.7, 'v': 0.05}, 'beta_weighting': 25.0}
    r = ref.tpwl.TPWLATV(data, params=params, Hf=Hf, discr_method=discr)
    o = TPWLATVNP(data, params=params, Hf=Hf, discr_method=discr)
    rng = np.random.default_rng(0)
    for _ in range(5):
        x = rng.normal(size=10)
        assert r.calc_nearest_point(x) == o.calc_nearest_point(x)
        assert np.array_equal(r.calc_weighting_factors(x), o.calc_weighting_factors(x))
        for a, b in zip(r.get_jacobians(x, dt=0.01), o.get_jacobians(x, dt=0.01)):
            assert np.array_equal(a, b)
        for a, b in zip(r.get_jacobians(x), o.get_jacobians(x)):
            assert np.array_equal(a, b)
    u = rng.uniform(0, 1500, size=(20, 3))
    xr, zr = r.rollout(rng.normal(size=10), u, 0.01)
    xo, zo = o.rollout(np.asarray(xr[0]), u, 0.01)
    assert np.array_equal(xr, xo) and np.array_equal(zr, zo)
    if method == 'nn':
        _quiet(r.pre_discretize, 0.01)
        o.pre_discretize(0.01)
        assert np.array_equal(np.array(r.A_d), np.array(o.A_d)) and np.array_equal(np.array(r.d_d), np.array(o.d_d))
        assert np.array_equal(r.get_characteristic_dx(0.01), o.get_characteristic_dx(0.01))


def test_ilqr_restatement_bitwise_on_tpwl(ref):
    from oracle.tpwl_np import TPWLATVNP
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost
    data, Hf = _small_bank()
    params = {'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}
    N = 25
    Q = np.zeros((6, 6)); Q[3, 3] = Q[4, 4] = 100.0
    R = 1e-5 * np.eye(3)
    Qf = 5 * Q
    rng = np.random.default_rng(1)
    x0 = rng.normal(size=10)
    results = []
    for cls, mcls, qc in ((ref.ilqr.iLQR, ref.tpwl.TPWLATV, ref.utils.QuadraticCost), (ILQRNP, TPWLATVNP, QuadraticCost)):
        m = mcls(data, params=params, Hf=Hf, discr_method='be')
        zt = np.tile(m.z_ref, (N + 1, 1)); zt[:, 3] += 0.05 * np.sin(np.linspace(0, 6, N + 1))
        s = cls(0.01, m, qc(Q, R, Qf), N)
        s.set_target(zt)
        s.set_u_last(np.array([10.0, 20.0, 30.0]))
        results.append(_quiet(s.ilqr_computation, x0, rng.uniform(0, 50, size=(N, 3)) * 0 + 25.0) + (s.rho,))
    for a, b in zip(*results):
        assert np.array_equal(a, b)


def test_ilqr_restatement_bitwise_on_ssm_adapter(ref):
    """The Gauss-Newton adapter drives the unmodified reference class and the restatement to identical bits."""
    import sofacontrol_b200.synth as synth
    from oracle.ssm_np import SSMDynamicsNP, GaussNewtonSSM
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost
    s = synth.trunk_ssm(8)
    N = 30
    zt = synth.figure8_targets(s['z_ref'], N, 7.0, 0.4)[0]
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    out = []
    for cls, qc in ((ref.ilqr.iLQR, ref.utils.QuadraticCost), (ILQRNP, QuadraticCost)):
        m = GaussNewtonSSM(SSMDynamicsNP(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params']))
        sol = cls(0.02, m, qc(Q, R, Qf), N)
        sol.set_target(zt)
        out.append(_quiet(sol.ilqr_computation, np.zeros(6)))
    for a, b in zip(*out):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("m,N,amp,q22,max_iter", [(4, 15, 3.0, -5000.0, 2), (8, 40, 6.0, -50.0, 50), (8, 40, 6.0, -500.0, 12),
                                                  (4, 30, 6.0, -200.0, 50)])
def test_ilqr_restatement_bitwise_on_the_non_pd_branch(ref, m, N, amp, q22, max_iter):
    """ilqr.py:276-299 on an indefinite stage cost: the UNMODIFIED reference class prints 'Q_uu not PD', raises rho,
    leaves the sweep and falls through to the decrease -- no restart, zero gains at and below the failing step.
    The restatement must follow it bit for bit (x, u, K, rho) and report the same number of failed PD tests."""
    import sofacontrol_b200.synth as synth
    from oracle.ssm_np import SSMDynamicsNP, GaussNewtonSSM
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost
    s = synth.trunk_ssm(m)
    zt = synth.figure8_targets(s['z_ref'], N, amp)[0]
    Q, R, Qf = synth.trunk_ilqr_costs(6, m)
    Q = Q.copy(); Q[2, 2] = q22
    out, counts = [], []
    for cls, qc in ((ref.ilqr.iLQR, ref.utils.QuadraticCost), (ILQRNP, QuadraticCost)):
        mdl = GaussNewtonSSM(SSMDynamicsNP(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params']))
        sol = cls(0.02, mdl, qc(Q, R, Qf), N)
        sol.params.max_iter = max_iter
        sol.set_target(zt)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), np.errstate(all='ignore'):
            out.append(sol.ilqr_computation(np.zeros(6)) + (sol.rho,))
        if cls is ILQRNP:
            counts.append((sum(1 for ev in sol.trace if ev['pd_fail_t'] >= 0), sol.iterations))
        else:
            counts.append((buf.getvalue().count('not PD'), buf.getvalue().count('Iteration')))
    assert counts[0] == counts[1] and counts[0][0] > 0
    for a, b in zip(*out):
        assert np.array_equal(a, b)


@pytest.fixture(scope="module")
def ref_ssm():
    return refimport.load_ssm()


@pytest.mark.parametrize("m", [4, 8])
@pytest.mark.parametrize("kw", [dict(discrete=False, discr_method='fe'), dict(discrete=False, discr_method='be'),
                                dict(discrete=False, discr_method='bil'), dict(discrete=True, discr_method='be')])
def test_ssm_restatement_vs_unmodified_reference_class(ref_ssm, m, kw):
    """sofacontrol/SSM/ssm.py imported UNMODIFIED (jax replaced by oracle/jax_shim.py: numpy float64 + exact
    forward-mode dual-number jacobian) vs oracle/ssm_np.py.  Maps, Jacobians and A_d agree to 1e-14 of their scale (bitwise on most
    inputs: the only difference in arithmetic is x**3 by libm pow in the lambdified basis vs (x*x)*x in the
    restatement, which is also how XLA lowers integer_pow); the cancellation residue d (ssm.py:203) within a few
    ulp of its largest term."""
    import sofacontrol_b200.synth as synth
    from oracle.ssm_np import SSMDynamicsNP
    s = synth.trunk_ssm(m)
    r = ref_ssm.SSMDynamics(s['z_ref'], model=s['model'], params=s['params'], **kw)
    o = SSMDynamicsNP(s['z_ref'], model=s['model'], params=s['params'], **kw)
    for attr in ('state_dim', 'input_dim', 'output_dim', 'SSM_order', 'ROM_order', 'Ts', 'nonlinear_observer'):
        assert getattr(r, attr) == getattr(o, attr)
    assert np.array_equal(r.H, o.H) and not r.H.any()
    close = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max()) <= 1e-14 * max(float(np.abs(np.asarray(b)).max()), 1e-300)
    rng = np.random.default_rng(m)
    eps = np.finfo(np.float64).eps
    for _ in range(12):
        x = rng.normal(size=6) * np.array([3, 3, 3, 30, 30, 30.0])
        u = rng.uniform(0, 800, size=m)
        z = s['z_ref'] + rng.normal(size=6)
        assert close(np.asarray(r.rom_phi(*x)), __import__('oracle.ssm_np', fromlist=['x']).poly_features(x, o.rom_table))
        assert close(r.reduced_dynamics(x, u), o.reduced_dynamics(x, u))
        assert close(r.reduced_dynamics_discrete(x, u), o.reduced_dynamics_discrete(x, u))
        assert close(r.C_map(x), o.C_map(x)) and close(r.W_map(x), o.W_map(x))
        assert close(r.compute_RO_state(z), o.compute_RO_state(z))
        assert close(r.x_to_zy(x), o.x_to_zy(x))
        A, B, d = r.get_continuous_jacobians(x, u)
        Ao, Bo, do = o.get_continuous_jacobians(x, u)
        scale = (np.abs(r.reduced_dynamics(x, u)) + np.abs(A) @ np.abs(x) + np.abs(B) @ np.abs(u)).max()
        assert close(A, Ao) and close(B, Bo) and np.abs(d - do).max() <= 8 * eps * scale
        if kw['discrete']:
            A, B, d = r.get_discrete_jacobians(x, u)
            Ao, Bo, do = o.get_discrete_jacobians(x, u)
            assert close(A, Ao) and close(B, Bo) and np.abs(d - do).max() <= 8 * eps * scale
        for a, b in zip(r.get_observer_jacobians(x), o.get_observer_jacobians(x)):
            assert close(a, b)
        assert close(r.update_observer_state(x), o.update_observer_state(x))
        Ad, Bd, dd = r.get_jacobians(x, u, 0.02)
        Ado, Bdo, ddo = o.get_jacobians(x, u, 0.02)
        assert close(Ad, Ado) and np.abs(Bd - Bdo).max() <= 4 * eps * np.abs(Bdo).max()
        xn, xno = r.update_state(x, u, 0.02), o.update_state(x, u, 0.02)
        assert np.abs(xn - xno).max() <= 1e-14 * np.abs(xno).max()
    xs = rng.normal(size=(9, 6))
    assert close(r.x_to_zfyf(xs), o.x_to_zfyf(xs))
    assert np.array_equal(r.zfyf_to_zy(zf=xs), o.zfyf_to_zy(zf=xs)) and np.array_equal(r.zy_to_zfyf(z=xs), o.zy_to_zfyf(z=xs))
    uu = rng.uniform(0, 800, size=(40, m))
    (xr, zr), (xo, zo) = r.rollout(np.zeros(6), uu, 0.02), o.rollout(np.zeros(6), uu, 0.02)
    assert np.abs(xr - xo).max() <= 1e-13 * np.abs(xo).max() and np.abs(zr - zo).max() <= 1e-13 * np.abs(zo).max()


def test_ssm_zoh_raises_like_the_reference(ref_ssm):
    import sofacontrol_b200.synth as synth
    from oracle.ssm_np import SSMDynamicsNP
    s = synth.trunk_ssm(4)
    for cls in (ref_ssm.SSMDynamics, SSMDynamicsNP):
        mdl = cls(s['z_ref'], discrete=False, discr_method='zoh', model=s['model'], params=s['params'])
        with pytest.raises(RuntimeError):
            mdl.get_jacobians(np.zeros(6), np.zeros(4), 0.01)


def test_gauss_newton_ilqr_reference_classes_vs_restatements(ref, ref_ssm):
    """Reference iLQR class driving the reference SSM class (through the H-property adapter of SURVEY App. C.2)
    vs ILQRNP driving SSMDynamicsNP: same iteration count and rho, x / u / K to 1e-12."""
    import sofacontrol_b200.synth as synth
    from oracle.ssm_np import SSMDynamicsNP, GaussNewtonSSM
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost
    s = synth.trunk_ssm(8)
    N = 30
    zt = synth.figure8_targets(s['z_ref'], N, 9.0, 1.1)[0]
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    kw = dict(discrete=False, discr_method='be', model=s['model'], params=s['params'])
    rs = ref.ilqr.iLQR(0.02, GaussNewtonSSM(ref_ssm.SSMDynamics(s['z_ref'], **kw)), ref.utils.QuadraticCost(Q, R, Qf), N)
    os_ = ILQRNP(0.02, GaussNewtonSSM(SSMDynamicsNP(s['z_ref'], **kw)), QuadraticCost(Q, R, Qf), N)
    res = []
    for sol in (rs, os_):
        sol.set_target(zt)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res.append(sol.ilqr_computation(0.05 * np.ones(6)))
        if sol is rs:
            its = buf.getvalue().count('Iteration')
    assert its == os_.iterations and rs.rho == os_.rho
    for a, b in zip(*res):
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()


def test_pod_restatement(ref):
    import sofacontrol_b200.synth as synth
    from oracle import pod_np
    X, _, _ = synth.pod_snapshots(300, 80, seed=5)
    Uf, U, nb, S = ref.pod.compute_POD(X, 1e-4)
    Uf2, U2, nb2, S2 = pod_np.compute_POD(X, 1e-4)
    assert nb == nb2 and np.array_equal(S, S2) and np.array_equal(U, U2)
    info = {'U': U, 'q_ref': X[:, 0], 'v_ref': X[:, 1], 'type': 'POD'}
    r, o = ref.pod.POD(info), pod_np.PODNP(info)
    v = np.random.default_rng(0).normal(size=300)
    assert np.array_equal(r.compute_RO_state(qf=v), o.compute_RO_state(qf=v))
    assert np.array_equal(r.compute_FO_state(q=v[:nb]), o.compute_FO_state(q=v[:nb]))
    assert np.array_equal(r.V, o.V)


def test_utils_restatement(ref):
    from oracle import utils_np
    rng = np.random.default_rng(0)
    A, B, d = rng.normal(size=(6, 6)), rng.normal(size=(6, 2)), rng.normal(size=6)
    for a, b in zip(ref.utils.zoh_affine(A, B, d, 0.05), utils_np.zoh_affine(A, B, d, 0.05)):
        assert np.array_equal(a, b)
    x = rng.normal(size=(4, 10))
    assert all(np.array_equal(a, b) for a, b in zip(ref.utils.x2qv(x), utils_np.x2qv(x)))
    assert np.array_equal(ref.utils.qv2x(x[:, :5], x[:, 5:]), utils_np.qv2x(x[:, :5], x[:, 5:]))


def test_monomial_order_matches_reference_sympy_construction():
    """ssm.py:158-164 builds the basis with sympy (itermonomials sorted by grevlex on reversed variables, constant
    dropped).  The same sympy construction is repeated here and compared with the combinatorial table used by the
    oracle and by the CUDA path; the analytic Jacobian is checked against sympy differentiation."""
    import sympy as sp
    from sympy.polys.monomials import itermonomials
    from sympy.polys.orderings import monomial_key
    from oracle.ssm_np import monomial_index_table, poly_features, poly_features_jac
    for dim, order in ((3, 2), (3, 3), (6, 3), (4, 4)):
        zeta = sp.Matrix(sp.symbols('x1:{}'.format(dim + 1)))
        polys = sorted(itermonomials(list(zeta), order), key=monomial_key('grevlex', list(reversed(zeta))))[1:]
        table = monomial_index_table(dim, order)
        assert len(polys) == table.shape[0]
        for p, row in zip(polys, table):
            expect = sp.Integer(1)
            for j in row:
                if j >= 0:
                    expect *= zeta[int(j)]
            assert sp.simplify(p - expect) == 0
        x = np.random.default_rng(dim).normal(size=dim)
        f = sp.lambdify(zeta, polys, 'numpy')
        assert np.allclose(poly_features(x, table), np.array(f(*x), dtype=float), rtol=1e-14)
        J = sp.lambdify(zeta, sp.Matrix(polys).jacobian(zeta), 'numpy')
        assert np.allclose(poly_features_jac(x, table), np.array(J(*x), dtype=float), rtol=1e-13, atol=1e-15)


def test_golden_vectors_are_what_the_reference_produces_today(ref, golden):
    """Re-runs a slice of oracle/make_golden.py and compares with the committed fixtures."""
    gt = golden("tpwl_small.npz")
    data, Hf = _small_bank()
    m = ref.tpwl.TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='be')
    assert np.array_equal(np.array([m.calc_nearest_point(x) for x in gt['xs']]), gt['idx_w10'])
    x, z = m.rollout(gt['xs'][0], gt['useq'], 0.01)
    assert np.array_equal(x, gt['nn_x_be']) and np.array_equal(z, gt['nn_z_be'])


# ------------------------------------------------------------------------------------------------------------
# Callers either side of the hot path: observers, infinite-horizon / time-varying gains,
# bank construction -- the reference modules are importable (python-control and np.infty are stubbed in
# oracle/refimport.py), so the restatements are pinned bit for bit.
# ------------------------------------------------------------------------------------------------------------
def _meas(rows, nf):
    Cf = np.zeros((len(rows), nf))
    for i, j in enumerate(rows):
        Cf[i, j] = 1.0
    return Cf


def test_ekf_restatement_bitwise(ref):
    from oracle.tpwl_np import TPWLATVNP
    from oracle.observer_np import DiscreteEKFObserverNP, FullStateObserverNP
    data, Hf = _small_bank()
    Cf = _meas((3, 17, 64, 90), 120)
    prm = {'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}
    mr = ref.tpwl.TPWLATV(data, params=prm, Hf=Hf, Cf=Cf, discr_method='be')
    mo = TPWLATVNP(data, params=prm, Hf=Hf, Cf=Cf, discr_method='be')
    kw = dict(W=0.5 * np.eye(10), V=0.01 * np.eye(4), Sigma0=2.0 * np.eye(10))
    er, eo = ref.observer.DiscreteEKFObserver(mr, **kw), DiscreteEKFObserverNP(mo, **kw)
    assert np.array_equal(er.x, eo.x) and np.array_equal(er.z, eo.z)
    rng = np.random.default_rng(0)
    for _ in range(12):
        u, y = rng.uniform(0, 1000, size=3), mr.y_ref + rng.normal(size=4)
        er.update(u, y, 0.01); eo.update(u, y, 0.01)
        assert np.array_equal(er.x, eo.x) and np.array_equal(er.Sigma, eo.Sigma) and np.array_equal(er.z, eo.z)
    fr, fo = ref.observer.FullStateObserver(10, H=mr.H), FullStateObserverNP(10, H=mo.H)
    x = rng.normal(size=10)
    fr.update(None, None, 0.01, x=x); fo.update(None, None, 0.01, x=x)
    assert np.array_equal(fr.z, fo.z)


def test_lqr_gain_restatements_bitwise(ref):
    from oracle import lqr_np
    rng = np.random.default_rng(0)
    n, m = 10, 3
    A = np.eye(n) + 0.05 * rng.normal(size=(n, n)); B = rng.normal(size=(n, m)); Q = np.eye(n); R = 0.1 * np.eye(m)
    for a, b in zip(ref.lqr.solve_riccati(A, B, Q, R), lqr_np.solve_riccati(A, B, Q, R)):
        assert np.array_equal(a, b)
    for a, b in zip(ref.lqr.dare(A, B, Q, R), lqr_np.dare(A, B, Q, R)):
        assert np.array_equal(a, b)


def test_traj_tracking_lqr_restatement_bitwise(ref):
    from oracle import lqr_np
    from oracle.tpwl_np import TPWLATVNP
    data, Hf = _small_bank()
    prm = {'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}
    mr = ref.tpwl.TPWLATV(data, params=prm, Hf=Hf, discr_method='be')
    mo = TPWLATVNP(data, params=prm, Hf=Hf, discr_method='be')
    rng = np.random.default_rng(2)

    class T:
        pass
    tg = T(); tg.t = np.linspace(0, 0.3, 31); tg.x = rng.normal(size=(31, 10)); tg.u = rng.uniform(0, 100, size=(31, 3))
    qc = ref.utils.QuadraticCost(Q=np.eye(10), R=0.01 * np.eye(3))
    a, b = ref.traj_tracking_lqr.TrajTrackingLQR(0.01, mr, qc), lqr_np.TrajTrackingLQRNP(0.01, mo, qc)
    (Ka, Pa), (Kb, Pb) = a.perform_dlqr_recursion(tg), b.perform_dlqr_recursion(tg)
    assert np.array_equal(Ka, Kb) and np.array_equal(Pa, Pb) and np.array_equal(a.x_bar, b.x_bar)
    assert np.array_equal(a.u_bar, b.u_bar)


def test_extract_AB_restatement_bitwise(ref):
    from oracle import lqr_np
    rng = np.random.default_rng(4)
    r = 6
    K = rng.normal(size=(r, r)); K = K @ K.T + r * np.eye(r)
    M = np.eye(r) + 0.1 * rng.normal(size=(r, r)); M = M @ M.T
    D, H = 0.1 * K + np.eye(r), rng.normal(size=(r, 2))
    for a, b in zip(ref.utils.extract_AB(K, D, M, H), lqr_np.extract_AB(K, D, M, H)):
        assert np.array_equal(a, b)
