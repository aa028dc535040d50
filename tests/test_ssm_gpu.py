"""GPU parity: SSM kernels (csrc/ssm.cu, csrc/ilqr_fast.cu) vs golden vectors produced by the reference's own
SSMDynamics class (ssm.py imported unmodified on oracle/jax_shim.py: ssm_units.npz, ssm_module_test.npz) and vs the
FP64 numpy oracle (oracle/ssm_np.py, pinned to that class).  Tolerance: relative 1e-9 (BASELINE.json north_star);
the affine residue d = f - A x - B u is a difference of large terms and is held to 1e-9 of |d| wherever |d| is not
itself below the rounding of its terms, and always to 64 ulp of the largest term (d_bound)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _models(m=4, **kw):
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.SSM.ssm import SSMDynamics
    from oracle.ssm_np import SSMDynamicsNP
    s = synth.trunk_ssm(m)
    g = SSMDynamics(s['z_ref'], model=s['model'], params=s['params'], **kw)
    o = SSMDynamicsNP(s['z_ref'], model=s['model'], params=s['params'], **kw)
    return g, o


def d_ok(d, d_ref, o, x, u, dt=0.02):
    """d = f - A x - B u (ssm.py:203): relative 1e-9 of max|d|, unless the reference's own float64 evaluation of d is
    less accurate than that -- its rounding error is a few ulp of (|f| + |A||x| + |B||u|), so two correct float64
    evaluations (the oracle itself differs from the reference class by this much, tests/test_oracle_golden.py) can only
    be asked to agree to that.  64 ulp covers the 6 / 8-term dot products and the FMA contraction of the kernel."""
    Ac, Bc, dc = o.get_continuous_jacobians(x, u) if not o.discrete else o.get_discrete_jacobians(x, u)
    f = o.reduced_dynamics(x, u) if not o.discrete else o.reduced_dynamics_discrete(x, u)
    floor = 64 * np.finfo(np.float64).eps * (np.abs(f) + np.abs(Ac) @ np.abs(x) + np.abs(Bc) @ np.abs(u)).max()
    if not o.discrete and o.discr_method != 'fe':
        floor *= max(1.0, np.abs(np.linalg.inv(Ac)).max() * 2.0)      # d_d = A_c^-1 (A_d - I) d_c
    elif not o.discrete:
        floor *= dt
    err = np.abs(np.asarray(d) - d_ref).max()
    return err <= max(TOL * np.abs(d_ref).max(), floor)


@pytest.mark.parametrize("tag,m", [("diamond", 4), ("trunk", 8)])
def test_units_vs_reference_class_golden(golden, tag, m):
    """Every SSM entry point on 64 seeded states vs the outputs of the REFERENCE class (golden ssm_units.npz)."""
    gu = golden("ssm_units.npz")
    X, U, Z = gu[tag + '_x'], gu[tag + '_u'], gu[tag + '_z']
    g, o = _models(m, discrete=False, discr_method='fe')
    assert relerr(g.reduced_dynamics(X, U), gu[tag + '_f']) < TOL
    assert relerr(g.x_to_zfyf(X), gu[tag + '_zf']) < TOL
    assert relerr(g.C_map(X.T).T, gu[tag + '_C']) < TOL
    assert relerr(g.compute_RO_state(Z.T).T, gu[tag + '_W']) < TOL            # (n_z, N) column convention of W_map
    A, B, d = g.get_continuous_jacobians(X, U)
    assert relerr(A, gu[tag + '_cA']) < TOL and relerr(B, gu[tag + '_cB']) < TOL
    H, c = g.get_observer_jacobians(X)
    assert relerr(H, gu[tag + '_H']) < TOL and relerr(c, gu[tag + '_c']) < TOL
    for meth, kw in (('fe', dict(discrete=False, discr_method='fe')), ('be', dict(discrete=False, discr_method='be')),
                     ('bil', dict(discrete=False, discr_method='bil')), ('disc', dict(discrete=True, discr_method='be'))):
        gm, om = _models(m, **kw)
        A, B, d = gm.get_jacobians(X, U, 0.02)
        assert relerr(A, gu['%s_%s_A' % (tag, meth)]) < TOL and relerr(B, gu['%s_%s_B' % (tag, meth)]) < TOL
        for i in range(X.shape[0]):
            assert d_ok(d[i], gu['%s_%s_d' % (tag, meth)][i], om, X[i], U[i])
        assert relerr(gm.update_state(X, U, 0.02), gu['%s_%s_next' % (tag, meth)]) < TOL


@pytest.mark.parametrize("m", [4, 8])
@pytest.mark.parametrize("method", ["fe", "be", "bil"])
def test_jacobians_single_and_batch(m, method):
    g, o = _models(m, discrete=False, discr_method=method)
    rng = np.random.default_rng(0)
    X = rng.normal(0, 1.5, size=(33, 6))
    U = rng.uniform(0, 800, size=(33, m))
    A, B, d = g.get_jacobians(X, U, 0.02)
    for i in range(X.shape[0]):
        Ao, Bo, do = o.get_jacobians(X[i], U[i], 0.02)
        assert relerr(A[i], Ao) < TOL and relerr(B[i], Bo) < TOL
        assert d_ok(d[i], do, o, X[i], U[i])
    A1, B1, d1 = g.get_jacobians(X[5], U[5], 0.02)          # 1-D call keeps the reference's shapes
    assert A1.shape == (6, 6) and B1.shape == (6, m) and d1.shape == (6,)
    assert np.array_equal(A1, A[5]) and np.array_equal(d1, d[5])


def test_continuous_discrete_and_observer_jacobians():
    g, o = _models(4, discrete=False, discr_method='be')
    gd, od = _models(4, discrete=True, discr_method='be')
    rng = np.random.default_rng(1)
    for _ in range(5):
        x = rng.normal(0, 1.0, size=6)
        u = rng.uniform(0, 800, size=4)
        for a, b in zip(g.get_continuous_jacobians(x, u), o.get_continuous_jacobians(x, u)):
            assert relerr(a, b) < TOL
        for a, b in zip(gd.get_discrete_jacobians(x, u), od.get_discrete_jacobians(x, u)):
            assert relerr(a, b) < TOL
        for a, b in zip(gd.get_jacobians(x, u, 0.01), od.get_jacobians(x, u, 0.01)):
            assert relerr(a, b) < TOL
        H, c = g.get_observer_jacobians(x)
        Ho, co = o.get_observer_jacobians(x)
        assert relerr(H, Ho) < TOL and relerr(c, co) < TOL
        assert relerr(g.update_observer_state(x), o.update_observer_state(x)) < TOL
        assert relerr(g.update_state(x, u, 0.02), o.update_state(x, u, 0.02)) < TOL
        assert relerr(g.reduced_dynamics(x, u), o.reduced_dynamics(x, u)) < TOL
        assert relerr(gd.reduced_dynamics_discrete(x, u), od.reduced_dynamics_discrete(x, u)) < TOL


def test_maps_and_shapes():
    g, o = _models(4)
    rng = np.random.default_rng(2)
    X = rng.normal(0, 1.0, size=(17, 6))
    assert relerr(g.x_to_zfyf(X), o.x_to_zfyf(X)) < TOL
    assert relerr(g.x_to_zfyf(X[0]), o.x_to_zfyf(X[0])) < TOL
    assert relerr(g.C_map(X.T), o.C_map(X.T)) < TOL                   # (n, N) column convention
    assert relerr(g.x_to_zy(X[3]), o.x_to_zy(X[3])) < TOL
    Z = o.x_to_zfyf(X)
    assert relerr(g.compute_RO_state(Z[4]), o.compute_RO_state(Z[4])) < TOL
    assert relerr(g.W_map(X.T), o.W_map(X.T)) < TOL
    assert np.array_equal(g.zfyf_to_zy(Z), o.zfyf_to_zy(Z)) and np.array_equal(g.zy_to_zfyf(Z), o.zy_to_zfyf(Z))
    assert g.get_state_dim() == 6 and g.get_input_dim() == 4 and g.get_output_dim() == 6
    assert g.H.shape == (6, 6) and not g.H.any() and g.nonlinear_observer


def test_bad_discretisation_raises_like_reference():
    g, _ = _models(4, discrete=False, discr_method='zoh')
    with pytest.raises(RuntimeError):
        g.get_jacobians(np.zeros(6), np.zeros(4), 0.01)


@pytest.mark.parametrize("name,kw", [("be", dict(discrete=False, discr_method='be')),
                                     ("fe", dict(discrete=False, discr_method='fe')),
                                     ("bil", dict(discrete=False, discr_method='bil')),
                                     ("disc", dict(discrete=True, discr_method='be'))])
def test_module_test_rollout_golden(golden, name, kw):
    """The reference's module_test (examples/hardware/diamond_SSM.py:83-140): 1001-step open-loop rollout on the
    recorded inputs; states/outputs vs the rollouts of the REFERENCE class (golden) and the MSE vs the recorded SOFA outputs."""
    gm = golden("ssm_module_test.npz")
    g, _ = _models(4, **kw)
    x, z = g.rollout(np.zeros(6), gm['u'], float(gm['dt']))
    assert x.shape == gm['x_' + name].shape and z.shape == gm['z_' + name].shape
    assert relerr(x, gm['x_' + name]) < TOL
    assert relerr(z, gm['z_' + name]) < TOL
    err = gm['z_true_qv'] - z[:-1]
    mse = np.linalg.norm(np.linalg.norm(err, axis=1)) ** 2 / err.shape[0]
    assert abs(mse - float(gm['mse_' + name])) < 1e-9 * float(gm['mse_' + name])


def test_batched_rollout_matches_oracle_and_is_batch_invariant():
    g, o = _models(8, discrete=False, discr_method='be')
    rng = np.random.default_rng(3)
    Bt, N = 37, 50
    x0 = rng.normal(0, 0.3, size=(Bt, 6))
    u = rng.uniform(0, 800, size=(Bt, N, 8))
    x, z = g.rollout(x0, u, 0.02)
    assert x.shape == (Bt, N + 1, 6) and z.shape == (Bt, N + 1, 6)
    for b in (0, 7, 36):
        xo, zo = o.rollout(x0[b], u[b], 0.02)
        assert relerr(x[b], xo) < TOL and relerr(z[b], zo) < TOL
    x1, z1 = g.rollout(x0[7], u[7], 0.02)                              # same bits alone or inside a batch
    assert np.array_equal(x1, x[7]) and np.array_equal(z1, z[7])


def test_empty_batch_and_zero_horizon():
    g, _ = _models(4)
    A, B, d = g.get_jacobians(np.zeros((0, 6)), np.zeros((0, 4)), 0.01)
    assert A.shape == (0, 6, 6) and B.shape == (0, 6, 4) and d.shape == (0, 6)
    x, z = g.rollout(np.ones((3, 6)) * 0.1, np.zeros((3, 0, 4)), 0.01)
    assert x.shape == (3, 1, 6) and z.shape == (3, 1, 6)


def test_c_abi_error_codes_on_device():
    """Argument / workspace validation of the C ABI with a live device: negative codes, no launch, message set."""
    import ctypes
    import torch
    from sofacontrol_b200 import _lib as L
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.lqr.ilqr import iLQR, C_addr
    from sofacontrol_b200.utils import QuadraticCost
    g, _ = _models(4, discrete=False, discr_method='be')
    lib = L.lib()
    h = g.device_model()
    x = torch.zeros((2, 6), device="cuda", dtype=torch.float64)
    assert lib.srcb200_ssm_eval_linearize_batch(h, 2, L.ptr(x), None, 0.01, L.ptr(x), None, None, None, None, None, None) == L.E_NULL
    assert b"u is NULL" in lib.srcb200_last_error_string()
    assert lib.srcb200_ssm_map_batch(h, 5, 0, 2, L.ptr(x), None, L.ptr(x), None) == L.E_DIM
    s_ = synth.trunk_ssm(4)
    zt = synth.figure8_targets(s_['z_ref'], 10, 5.0)[0]
    Q, R, Qf = synth.trunk_ilqr_costs(6, 4)
    sol = iLQR(0.02, g, QuadraticCost(Q, R, Qf), 10)
    sol.set_target(zt)
    pr = sol._problem(2, x, None, L.to_dev(zt), None)
    res = L.IlqrResult()
    assert lib.srcb200_ilqr_solve_batch(L.ILQR_MODEL_SSM, C_addr(h), sol._cfg(), pr, res, None, 0, None) == L.E_NULL
    out = {k: L.empty(s) for k, s in (('x', (2, 11, 6)), ('u', (2, 10, 4)), ('K', (2, 10, 4, 6)), ('cost', (2,)))}
    it = L.empty((2,), torch.int32); st = L.empty((2,), torch.int32)
    res = L.IlqrResult(x=L.ptr(out['x']), u=L.ptr(out['u']), K=L.ptr(out['K']), cost=L.ptr(out['cost']),
                       iterations=L.ptr(it), status=L.ptr(st))
    small = L.empty((16,))
    assert lib.srcb200_ilqr_solve_batch(L.ILQR_MODEL_SSM, C_addr(h), sol._cfg(), pr, res, L.ptr(small), 128, None) == L.E_WORKSPACE
    assert lib.srcb200_ilqr_solve_batch(7, C_addr(h), sol._cfg(), pr, res, L.ptr(small), 128, None) == L.E_DIM
    with pytest.raises(L.Srcb200Error):
        L.check(L.E_WORKSPACE)


def test_device_copies_follow_modified_model_data():
    """The device handle is keyed on a fingerprint of the coefficient arrays: editing the model after first use
    (here r_coeff and z_ref) must change the results like it does in the reference class."""
    g, o = _models(4, discrete=False, discr_method='fe')
    rng = np.random.default_rng(5)
    x, u = 0.3 * rng.normal(size=6), rng.uniform(0, 800, size=4)
    A0, B0, d0 = g.get_jacobians(x, u=u, dt=0.02)
    g.r_coeff = g.r_coeff * 1.5
    o.r_coeff = o.r_coeff * 1.5
    g.z_ref = g.z_ref + 1.0
    o.z_ref = o.z_ref + 1.0
    A1, B1, d1 = g.get_jacobians(x, u=u, dt=0.02)
    Ao, Bo, do_ = o.get_jacobians(x, u=u, dt=0.02)
    assert relerr(A1, Ao) < TOL and not np.allclose(A1, A0)
    assert relerr(g.x_to_zfyf(x[None, :], zf=True), o.x_to_zfyf(x[None, :], zf=True)) < TOL


@pytest.mark.parametrize("dense", ["0", "1"])
def test_sparse_and_dense_evaluation_kernels_agree_with_oracle(dense, monkeypatch):
    """Kernel (b) in both formulations (eight-states-per-warp sparse contraction / dense 42-DMMA contraction) on a batch
    that is not a multiple of 8: A_c, d_c, H, c, z vs the oracle."""
    monkeypatch.setenv("SRCB200_SSM_EVAL_DENSE", dense)
    from sofacontrol_b200 import _lib as L
    g, o = _models(8, discrete=False, discr_method='be')
    rng = np.random.default_rng(11)
    X = 0.4 * rng.normal(size=(45, 6))
    U = rng.uniform(0, 800, size=(45, 8))
    out = g._eval_device(L.to_dev(X), L.to_dev(U), -1.0, 'cont_raw', ('A', 'd', 'H', 'c', 'z'))
    for i in (0, 7, 8, 44):
        Ac, Bc, dc = o.get_continuous_jacobians(X[i], U[i])
        H, c = o.get_observer_jacobians(X[i])
        assert relerr(out['A'][i].cpu().numpy(), Ac) < TOL and relerr(out['H'][i].cpu().numpy(), H) < TOL
        assert d_ok(out["d"][i].cpu().numpy(), dc, o, X[i], U[i], dt=1.0)
        assert np.abs(out['c'][i].cpu().numpy() - c).max() <= 1e-9 * max(1.0, np.abs(o.C_map(X[i])).max())
        assert relerr(out['z'][i].cpu().numpy(), o.x_to_zfyf(X[i][None, :], zf=True)[0]) < TOL
