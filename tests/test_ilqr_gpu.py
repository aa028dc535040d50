"""GPU parity: batched iLQR kernel (csrc/ilqr.cu) vs golden vectors produced by the UNMODIFIED reference iLQR class
and vs the numpy oracle (oracle/ilqr_np.py) including its branch trace.  Tolerance: relative 1e-9 on states, inputs,
gains and costs."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _ssm(m, **kw):
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.SSM.ssm import SSMDynamics
    s = synth.trunk_ssm(m)
    return s, SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])


def _oracle_ssm(m):
    import sofacontrol_b200.synth as synth
    from oracle.ssm_np import SSMDynamicsNP, GaussNewtonSSM
    s = synth.trunk_ssm(m)
    return GaussNewtonSSM(SSMDynamicsNP(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params']))


def _solver(model, m, zt, **kw):
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost
    Q, R, Qf = synth.trunk_ilqr_costs(6, m)
    s = iLQR(0.02, model, QuadraticCost(Q, R, Qf), zt.shape[-2] - 1, **kw)
    s.set_target(zt)
    return s


def _floor_tol(m, zt, x, u, rho=0.0, factor=8.0):
    """1e-9, or `factor` x the measured float64 noise floor of the reference's own Riccati recursion on this
    trajectory (oracle.ilqr_np.riccati_noise_floor: float64 vs 80-bit evaluation of ilqr.py:258-295) when that is
    larger.  Returns (tol, floor_K, floor_k)."""
    import sofacontrol_b200.synth as synth
    from oracle.ilqr_np import ILQRNP, riccati_noise_floor
    from oracle.utils_np import QuadraticCost
    Q, R, Qf = synth.trunk_ilqr_costs(6, m)
    om = _oracle_ssm(m)
    o = ILQRNP(0.02, om, QuadraticCost(Q, R, Qf), u.shape[0])
    o.set_target(zt)
    xf, uf, _, Af, Bf, _ = o.forward_pass(x, u)
    fK, fk = riccati_noise_floor(om, QuadraticCost(Q, R, Qf), zt, xf, uf, Af, Bf, rho=rho)
    return max(TOL, factor * max(fK, fk)), fK, fk


@pytest.mark.parametrize("tag,m", [("diamond", 4), ("trunk", 8)])
def test_forward_and_backward_pass_units(golden, tag, m, capsys):
    """One forward pass and one backward pass of the reference class (golden), no branch sensitivity:
    (a) on the INITIAL zero-input rollout, where k and Q_u are O(1): x, u, cost, A, B, d, K, k, Q_u, Q_uu at 1e-9;
    (b) on the CONVERGED trajectory: forward quantities and K at 1e-9; k and Q_u are pure cancellation residues
        there (|k| ~ 1e-6 |u|, |Q_u| ~ 1e-9 of its terms) and Q_uu carries the rounding the un-symmetrised value
        recursion amplifies, so they are held to 8 x the measured float64 noise floor of the reference's own
        recursion relative to the scale of the terms that cancel (printed)."""
    gi = golden("ssm_ilqr.npz")
    _, model = _ssm(m)
    s = _solver(model, m, gi[tag + '_zt'])
    # (a) initial rollout
    N = gi[tag + '_u'].shape[0]
    x, u, cost, A, B, d = s.forward_pass(np.zeros((N + 1, 6)), np.zeros((N, m)))
    assert relerr(x, gi[tag + '_init_x']) < TOL and relerr(A, gi[tag + '_init_A']) < TOL
    assert relerr(B, gi[tag + '_init_B']) < TOL and relerr(d, gi[tag + '_init_d']) < TOL
    assert abs(cost - float(gi[tag + '_init_cost'])) < TOL * abs(float(gi[tag + '_init_cost']))
    s.rho, s.drho = 0.0, 0.0
    K, k, Qu, Quu = s.dlqr_recursion(gi[tag + '_init_x'], gi[tag + '_init_u'], gi[tag + '_init_A'], gi[tag + '_init_B'],
                                     gi[tag + '_init_d'])
    tol0, fK0, fk0 = _floor_tol(m, gi[tag + '_zt'], gi[tag + '_init_x'], gi[tag + '_init_u'])
    e0 = [relerr(K, gi[tag + '_init_K']), relerr(k, gi[tag + '_init_k']), relerr(Qu, gi[tag + '_init_Qu']),
          relerr(Quu, gi[tag + '_init_Quu'])]
    assert max(e0) < tol0, (e0, tol0)
    # (b) converged trajectory
    x, u, cost, A, B, d = s.forward_pass(gi[tag + '_x'], gi[tag + '_u'])
    assert relerr(x, gi[tag + '_x']) < TOL and np.array_equal(u, gi[tag + '_u'])
    assert abs(cost - float(gi[tag + '_fp_cost'])) < TOL * abs(float(gi[tag + '_fp_cost']))
    assert relerr(A, gi[tag + '_fp_A']) < TOL and relerr(B, gi[tag + '_fp_B']) < TOL and relerr(d, gi[tag + '_fp_d']) < TOL
    s.rho, s.drho = 0.0, 0.0
    K, k, Qu, Quu = s.dlqr_recursion(gi[tag + '_x'], gi[tag + '_u'], gi[tag + '_fp_A'], gi[tag + '_fp_B'], gi[tag + '_fp_d'])
    tol, fK, fk = _floor_tol(m, gi[tag + '_zt'], gi[tag + '_x'], gi[tag + '_u'])
    eK, eQuu = relerr(K, gi[tag + '_bp_K']), relerr(Quu, gi[tag + '_bp_Quu'])
    # scale of the terms that cancel in Q_u = R du + B^T p and k = -Quu~^-1 Q_u: the same quantities far from the optimum
    su, sk = np.abs(gi[tag + '_init_Qu']).max(), np.abs(gi[tag + '_init_k']).max()
    eQu = np.abs(Qu - gi[tag + '_bp_Qu']).max() / su
    ek = np.abs(k - gi[tag + '_bp_k']).max() / sk
    with capsys.disabled():
        print("\n[unit passes %s] initial: K %.1e k %.1e Qu %.1e Quu %.1e (tol %.1e) | converged: K %.1e Quu %.1e, "
              "k %.1e Qu %.1e of their un-cancelled scale; float64 Riccati floor K %.1e k %.1e -> tol %.1e"
              % (tag, e0[0], e0[1], e0[2], e0[3], tol0, eK, eQuu, ek, eQu, fK, fk, tol))
    assert eK < tol and eQuu < tol and eQu < tol and ek < tol
    assert abs(float(s.rho) - float(gi[tag + '_bp_rho'])) < 1e-15


@pytest.mark.parametrize("tag,m", [("diamond", 4), ("trunk", 8)])
def test_solve_matches_reference_golden(golden, tag, m, capsys):
    """Full solve from x0 = 0 on the figure-8 target: x, u, K vs the unmodified reference iLQR class driving the
    reference SSM class (golden).  1e-9, or 8 x the measured float64 noise floor of the reference's own Riccati
    recursion on this trajectory if that is larger (printed)."""
    gi = golden("ssm_ilqr.npz")
    _, model = _ssm(m)
    s = _solver(model, m, gi[tag + '_zt'], trace=True)
    x, u, K = s.ilqr_computation(np.zeros(6))
    assert x.shape == (101, 6) and u.shape == (100, m) and K.shape == (100, m, 6)
    e = [relerr(x, gi[tag + '_x']), relerr(u, gi[tag + '_u']), relerr(K, gi[tag + '_K'])]
    tol, fK, fk = (TOL, 0.0, 0.0) if max(e) < TOL else _floor_tol(m, gi[tag + '_zt'], gi[tag + '_x'], gi[tag + '_u'])
    with capsys.disabled():
        print("\n[solve %s N=100] x %.1e u %.1e K %.1e ; tol %.1e (Riccati floor K %.1e k %.1e)" % (tag, *e, tol, fK, fk))
    assert max(e) < tol
    assert abs(float(s.info['rho']) - float(gi[tag + '_rho'])) <= 1e-12 * max(1.0, abs(float(gi[tag + '_rho'])))
    assert s.info['status'] & 1                                            # converged


def test_branch_trace_matches_oracle():
    """Iteration by iteration: accepted step size, cost, rho after the backward pass and PD restarts vs the oracle."""
    import sofacontrol_b200.synth as synth
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost
    m = 8
    s_, model = _ssm(m)
    rng = np.random.default_rng(7)
    zt = synth.figure8_targets(s_['z_ref'], 60, [4.0, 11.0, 14.5], [0.3, 2.0, 4.4])
    x0 = rng.uniform(-0.5, 0.5, size=(3, 6)) * np.array([1, 1, 1, 0, 0, 0])
    s = _solver(model, m, zt, trace=True)
    x, u, K = s.ilqr_computation(x0)
    Q, R, Qf = synth.trunk_ilqr_costs(6, m)
    for b in range(3):
        o = ILQRNP(0.02, _oracle_ssm(m), QuadraticCost(Q, R, Qf), 60)
        o.set_target(zt[b])
        xo, uo, Ko = o.ilqr_computation(x0[b])
        assert s.info['iterations'][b] == o.iterations
        tr = s.info['trace'][b]
        for ev in o.trace:
            i = ev['it']
            assert tr[i, 3] == ev['pd_fail_t']
            assert abs(tr[i, 2] - ev['rho_after_bwd']) <= 1e-12 * max(1.0, ev['rho_after_bwd'])
            assert (tr[i, 1] > 0) == ev['accepted']
            if ev['accepted']:
                assert tr[i, 1] == 0.5 ** (len(ev['trials']) - 1)
            assert abs(tr[i, 0] - ev['cost']) <= 1e-9 * abs(ev['cost'])
        assert relerr(x[b], xo) < TOL and relerr(u[b], uo) < TOL and relerr(K[b], Ko) < TOL
        assert abs(s.info['cost'][b] - o.final_cost) < TOL * abs(o.final_cost)
        assert abs(s.info['cost0'][b] - o.initial_cost) < TOL * abs(o.initial_cost)


def test_literal_constant_H_mode_is_degenerate_like_reference():
    """With the literal SSM.H = zeros (ssm.py:72-73) the reference returns u = 0 after one iteration."""
    import sofacontrol_b200.synth as synth
    s_, model = _ssm(4)
    zt = synth.figure8_targets(s_['z_ref'], 30, 5.0)[0]
    s = _solver(model, 4, zt, gauss_newton=False)
    x, u, K = s.ilqr_computation(np.zeros(6))
    assert not u.any() and s.info['iterations'] == 1 and not K.any()


def test_warm_start_u_last_and_config_switches():
    import sofacontrol_b200.synth as synth
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost
    m = 4
    s_, model = _ssm(m)
    rng = np.random.default_rng(9)
    N = 25
    zt = synth.figure8_targets(s_['z_ref'], N, 6.0, 1.0)[0]
    uw = rng.uniform(0, 300, size=(N, m))
    ul = rng.uniform(0, 300, size=m)
    Q, R, Qf = synth.trunk_ilqr_costs(6, m)
    Qf = 10.0 * Q
    for variant in range(3):
        from sofacontrol_b200.lqr.ilqr import iLQR
        from sofacontrol_b200.utils import QuadraticCost as QC
        s = iLQR(0.02, model, QC(Q, R, Qf), N)
        o = ILQRNP(0.02, _oracle_ssm(m), QuadraticCost(Q, R, Qf), N)
        for obj in (s, o):
            obj.set_target(zt)
            obj.set_u_last(ul)
            if variant == 1:
                obj.params.state_regularization = False
                obj.params.rho0 = 0.5
                obj.params.drho0 = 1.0
            if variant == 2:
                obj.params.include_input_var_constraint = False
                obj.params.max_iter = 3
        x, u, K = s.ilqr_computation(0.1 * np.ones(6), u_warmstart=uw)
        xo, uo, Ko = o.ilqr_computation(0.1 * np.ones(6), u_warmstart=uw)
        assert s.info['iterations'] == o.iterations
        assert relerr(x, xo) < TOL and relerr(u, uo) < TOL and relerr(K, Ko) < TOL


def test_tpwl_solve_matches_reference_golden(golden):
    """TPWL (constant H, nn on the pre-discretised zoh bank of the reference) vs the unmodified reference class."""
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.tpwl.tpwl import TPWLATV
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost
    gt = golden("tpwl_small.npz")
    data, Hf = synth.tpwl_bank(seed=11, r=5, m=3, P=40, num_nodes=20, tip_node=7, spread=1.0)
    g = TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='zoh')
    g.set_pre_discretized(gt['zoh_A_d'], gt['zoh_B_d'], gt['zoh_d_d'], 0.01)
    s = iLQR(0.01, g, QuadraticCost(gt['ilqr_Q'], gt['ilqr_R'], np.zeros((6, 6))), 40)
    s.set_target(gt['ilqr_zt'])
    x, u, K = s.ilqr_computation(gt['xs'][0])
    assert relerr(x, gt['ilqr_x']) < TOL and relerr(u, gt['ilqr_u']) < TOL and relerr(K, gt['ilqr_K']) < TOL


def test_batch_is_consistent_and_statuses_reported():
    """4096-problem shape at a reduced batch (257): every member equals its own single solve bit for bit."""
    import sofacontrol_b200.synth as synth
    w = synth.trunk_ilqr_batch(257, N=40, seed=3, m=8)
    _, model = _ssm(8)
    s = _solver(model, 8, w['z_target'])
    x, u, K = s.ilqr_computation(w['x0'])
    assert x.shape == (257, 41, 6) and np.all(np.isfinite(x)) and np.all(np.isfinite(K))
    assert np.all(s.info['iterations'] >= 1) and np.all(s.info['cost'] <= s.info['cost0'])
    for b in (0, 100, 256):
        s1 = _solver(model, 8, w['z_target'][b])
        x1, u1, K1 = s1.ilqr_computation(w['x0'][b])
        assert np.array_equal(x1, x[b]) and np.array_equal(u1, u[b]) and np.array_equal(K1, K[b])


def test_specialised_kernels_agree_with_generic_kernels():
    """ilqr_fast.cu (warp per problem, DMMA tiles, shuffle sweeps) vs the generic cooperative kernels of ilqr.cu /
    ssm.cu on the same inputs (SRCB200_ILQR_GENERIC=1 forces the generic path): same iterations, 1e-9 on x, u, K."""
    import os
    import sofacontrol_b200.synth as synth
    w = synth.trunk_ilqr_batch(24, N=50, seed=5, m=8)
    _, model = _ssm(8)
    res = {}
    for tag, env in (("fast", "0"), ("generic", "1")):
        os.environ["SRCB200_ILQR_GENERIC"] = env
        s = _solver(model, 8, w['z_target'])
        res[tag] = s.ilqr_computation(w['x0']) + (s.info['iterations'], s.info['cost'])
        rng = np.random.default_rng(0)
        res[tag + "_roll"] = model.rollout(w['x0'], rng.uniform(0, 800, size=(24, 50, 8)), 0.02)
    os.environ["SRCB200_ILQR_GENERIC"] = "0"
    assert np.array_equal(res["fast"][3], res["generic"][3])
    for a, b in zip(res["fast"][:3], res["generic"][:3]):
        assert relerr(a, b) < TOL
    assert relerr(res["fast"][4], res["generic"][4]) < TOL
    for a, b in zip(res["fast_roll"], res["generic_roll"]):
        assert relerr(a, b) < 1e-11


@pytest.mark.parametrize("m", [4, 8])
def test_launch_shapes_of_the_fast_kernel_agree(m, monkeypatch):
    """SRCB200_ILQR_SHAPE 0 / 1 / 2 (16 / 12 / 8 warps per SM; Jacobian-table rows in shared memory / one / two in
    registers) run the same arithmetic: same iterations and statuses, x, u, K to 1e-12."""
    import sofacontrol_b200.synth as synth
    w = synth.trunk_ilqr_batch(40, N=30, seed=9, m=m)
    _, model = _ssm(m)
    res = {}
    for shape in ("0", "1", "2"):
        monkeypatch.setenv("SRCB200_ILQR_SHAPE", shape)
        s = _solver(model, m, w['z_target'])
        res[shape] = s.ilqr_computation(w['x0']) + (s.info['iterations'], s.info['status'])
    for shape in ("1", "2"):
        assert np.array_equal(res[shape][3], res["0"][3]) and np.array_equal(res[shape][4], res["0"][4])
        for a, b in zip(res[shape][:3], res["0"][:3]):
            assert relerr(a, b) < 1e-12


def test_tail_handover_between_launch_shapes_is_bit_identical(monkeypatch):
    """Large batches: the throughput launch shape stops taking tasks when 1776 problems are left and a second launch in
    the 8-warp shape finishes them from the same task rings (csrc/ilqr_fast.cu).  Suspended problems carry their whole
    state in global memory, so the schedule must not change a bit: x, u, K, costs, iterations, statuses with the
    hand-over (default), without it, and at an odd threshold -- for the shape-1 and the shape-0 range of batch sizes."""
    import torch
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200 import _lib as L
    _, model = _ssm(8)
    for batch in (2700, 4000):
        w = synth.trunk_ilqr_batch(batch, N=24, seed=31, m=8)
        s = _solver(model, 8, w['z_target'])
        x0, zt = L.to_dev(w['x0']), L.to_dev(w['z_target'])
        res = {}
        for tag, ho in (("default", None), ("off", "0"), ("odd", "777")):
            if ho is None:
                monkeypatch.delenv("SRCB200_ILQR_HANDOVER", raising=False)
            else:
                monkeypatch.setenv("SRCB200_ILQR_HANDOVER", ho)
            out = s.solve_device(x0, zt)
            res[tag] = {k: out[k].clone() for k in ('x', 'u', 'K', 'cost', 'iterations', 'status')}
        for tag in ("off", "odd"):
            for k, v in res["default"].items():
                assert torch.equal(v, res[tag][k]), (batch, tag, k)
        assert int(res["default"]['iterations'].max()) > int(res["default"]['iterations'].min())


def test_pinned_stream_api_equals_single_calls():
    """iLQR.solve_pinned_stream (double-buffered D2H) returns, batch by batch, exactly what solve_pinned returns."""
    import torch
    import sofacontrol_b200.synth as synth
    _, model = _ssm(8)
    ws = [synth.trunk_ilqr_batch(16, N=20, seed=s_, m=8) for s_ in (1, 2, 3, 4, 5)]
    s = _solver(model, 8, ws[0]['z_target'])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    batches = [(pin(w['x0']), pin(w['z_target'])) for w in ws]
    single = []
    for x0h, zth in batches:
        o = s.solve_pinned(x0h, zth)
        single.append({k: v.clone() for k, v in o.items()})
    n = 0
    for o, ref in zip(s.solve_pinned_stream(batches), single):
        for k in ref:
            assert torch.equal(o[k], ref[k]), k
        n += 1
    assert n == len(batches)


def test_receding_horizon_closed_loop_matches_oracle_loop():
    """Config-4 driver (receding-horizon iLQR with shifted warm start + u_last) vs the same loop written with the
    numpy oracle solver and model, noise-free, small case."""
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.mpc import RecedingHorizonILQR
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost
    m, N, steps, Bt = 4, 10, 4, 3
    s_, model = _ssm(m)
    rng = np.random.default_rng(11)
    zref = synth.figure8_targets(s_['z_ref'], steps + N, [3.0, 6.0, 9.0], [0.0, 1.0, 2.0])      # (Bt, steps+N+1, 6)
    x0 = rng.uniform(-0.3, 0.3, size=(Bt, 6)) * np.array([1, 1, 1, 0, 0, 0])
    sol = _solver(model, m, zref[:, :N + 1])
    out = RecedingHorizonILQR(sol).run(x0, zref, steps)
    Q, R, Qf = synth.trunk_ilqr_costs(6, m)
    for b in range(Bt):
        om = _oracle_ssm(m)
        x = x0[b].copy()
        u_plan, u_last = None, np.zeros(m)
        for k in range(steps):
            o = ILQRNP(0.02, om, QuadraticCost(Q, R, Qf), N)
            o.set_target(zref[b, k:k + N + 1])
            o.set_u_last(u_last)
            ws = None if u_plan is None else np.vstack((u_plan[1:], u_plan[-1:]))
            _, u_plan, _ = o.ilqr_computation(x, ws)
            assert out['iterations'][b, k] == o.iterations
            assert relerr(out['u'][b, k], u_plan[0]) < TOL
            x = om.ssm.update_state(x, u_plan[0], 0.02)
            u_last = u_plan[0]
            assert relerr(out['x'][b, k + 1], x) < TOL


def test_tpwl_diamond_size_solve_vs_oracle():
    """Diamond TPWL shape (n = 72, m = 4, P = 1000, zoh pre-discretised on the device), horizon 30: generic
    one-CTA-per-problem kernel vs the numpy oracle (which is pinned bitwise to the reference class)."""
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.tpwl.tpwl import TPWLATV
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost
    from oracle.tpwl_np import TPWLATVNP
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost as QCo
    data, Hf = synth.tpwl_bank()
    params = {'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}
    g = TPWLATV(data, params=params, Hf=Hf, discr_method='zoh')
    g.pre_discretize(0.01)
    o_model = TPWLATVNP(data, params=params, Hf=Hf, discr_method='zoh')
    # the oracle gets the DEVICE-discretised bank so that both sides linearise with identical matrices
    o_model.A_d, o_model.B_d, o_model.d_d, o_model.pre_discretized_dt = list(g.A_d), list(g.B_d), list(g.d_d), 0.01
    N = 30
    Q = np.zeros((6, 6)); Q[3, 3] = Q[4, 4] = 100.0
    R = 1e-5 * np.eye(4)
    Qf = np.zeros((6, 6))
    th = np.linspace(0, 2 * np.pi, N + 1)
    x0s, _ = synth.tpwl_rollout_batch(2, N=1, seed=21)
    zts = []
    for b in range(2):
        zt = np.tile(g.z_ref, (N + 1, 1))
        zt[:, 3] += (0.5 + b) * np.sin(th); zt[:, 4] += (0.5 + b) * np.sin(2 * th)
        zts.append(zt)
    zts = np.array(zts)
    s = iLQR(0.01, g, QuadraticCost(Q, R, Qf), N)
    s.set_target(zts)
    x, u, K = s.ilqr_computation(x0s)
    for b in range(2):
        o = ILQRNP(0.01, o_model, QCo(Q, R, Qf), N)
        o.set_target(zts[b])
        xo, uo, Ko = o.ilqr_computation(x0s[b])
        assert s.info['iterations'][b] == o.iterations
        assert relerr(x[b], xo) < TOL and relerr(u[b], uo) < TOL and relerr(K[b], Ko) < TOL


def _pair(m, N, tweak, zt, Qscale=1.0, x0=None, Qneg=False):
    """Solve the same problem with the CUDA kernel and the oracle after applying `tweak(params)` to both configs."""
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost as QC
    from oracle.ilqr_np import ILQRNP
    from oracle.utils_np import QuadraticCost
    _, model = _ssm(m)
    Q, R, Qf = synth.trunk_ilqr_costs(6, m)
    Q = Q * Qscale
    if Qneg:
        Q = Q.copy(); Q[2, 2] = -5000.0          # indefinite stage cost: Q_uu~ loses positive definiteness
    s = iLQR(0.02, model, QC(Q, R, Qf), N, trace=True)
    o = ILQRNP(0.02, _oracle_ssm(m), QuadraticCost(Q, R, Qf), N)
    for obj in (s, o):
        tweak(obj.params)
        obj.set_target(zt)
    x0 = np.zeros(6) if x0 is None else x0
    return s, o, s.ilqr_computation(x0), x0


def test_branch_paths_line_search_failure_and_abandon():
    """improv_lb = 0.99999 rejects every step: five trials per iteration, rho bumped (scaled + 10) each time,
    abandoned after counter_limit failures (ilqr.py:76-103)."""
    import sofacontrol_b200.synth as synth
    s_, _ = _ssm(4)
    zt = synth.figure8_targets(s_['z_ref'], 20, 5.0)[0]

    def tweak(p):
        p.improv_lb = 0.99999
    s, o, (x, u, K), x0 = _pair(4, 20, tweak, zt)
    xo, uo, Ko = o.ilqr_computation(x0)
    assert s.info['iterations'] == o.iterations
    assert s.info['status'] & 4                                            # abandoned after counter_limit failures
    assert sum(1 for ev in o.trace if not ev['accepted']) == 5
    assert s.info['trials'] == sum(len(ev['trials']) for ev in o.trace)
    assert abs(float(s.info['rho']) - float(o.rho)) <= 1e-12 * float(o.rho)
    assert relerr(x, xo) < TOL and relerr(K, Ko) < TOL and relerr(u, uo) < TOL
    for ev in o.trace:
        assert (s.info['trace'][ev['it'], 1] > 0) == ev['accepted']
        assert abs(s.info['trace'][ev['it'], 2] - ev['rho_after_bwd']) <= 1e-12 * max(1.0, ev['rho_after_bwd'])


NONPD_CASES = [  # tag, m, N, amplitude, Q[2,2], max_iter
    ("d4_first_step", 4, 15, 3.0, -5000.0, 50), ("d4_five", 4, 20, 3.0, -30.0, 4), ("t8_five", 8, 40, 6.0, -30.0, 4),
    ("t8_hard", 8, 40, 3.0, -500.0, 4), ("t8_twelve", 8, 40, 6.0, -500.0, 11),
    # long solves: the indefinite cost makes these closed loops unstable (|x| ~ 1e5..1e6), a ONE-ulp change of the target
    # moves the reference's own result by 1e-6..1e-2 (stored as *_ulp_sensitivity) -- they pin the branch sequence
    ("t8_mid", 8, 40, 6.0, -50.0, 50), ("d4_mid", 4, 30, 6.0, -200.0, 50)]


@pytest.mark.parametrize("tag,m,N,amp,q22,max_iter", NONPD_CASES)
def test_non_pd_branch_matches_reference_golden(golden, tag, m, N, amp, q22, max_iter, capsys):
    """Indefinite stage cost -> Q_uu~ fails the Cholesky test.  The reference (ilqr.py:282-299) raises rho, LEAVES the
    backward sweep (K_t = k_t = 0 at and below the failing step), lowers rho once and line-searches with those
    gains; it never restarts.  Golden = the unmodified reference class (oracle/make_golden.py: ilqr_nonpd.npz);
    the oracle restatement is pinned bitwise to it (tests/test_oracle_vs_reference.py).
    Exact on every case: iteration count, final rho, the failing step of every backward sweep, the zero pattern of K.
    x, u, K: 1e-9 on the well-conditioned cases (<= 12 iterations).  The long cases are unstable closed loops: the
    reference's OWN result moves by `ulp_sensitivity` (1e-6 .. 1e-2, measured in make_golden.py) when its target is
    changed by one ulp, so there the values are held to 10 x that measured sensitivity (printed)."""
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost as QC
    g = golden("ilqr_nonpd.npz")
    s_, model = _ssm(m)
    Q, R, Qf = synth.trunk_ilqr_costs(6, m)
    Q = Q.copy(); Q[2, 2] = q22
    zt = synth.figure8_targets(s_['z_ref'], N, amp)[0]
    s = iLQR(0.02, model, QC(Q, R, Qf), N, trace=True)
    s.params.max_iter = max_iter
    s.set_target(zt)
    x, u, K = s.ilqr_computation(np.zeros(6))
    it = int(g[tag + '_iterations'])
    assert int(s.info['iterations']) == it
    assert np.array_equal(s.info['trace'][:it, 3].astype(int), g[tag + '_pd_fail_t'])
    assert (g[tag + '_pd_fail_t'] >= 0).any() and (s.info['status'] & 8)      # the path is really exercised
    assert abs(float(s.info['rho']) - float(g[tag + '_rho'])) <= 1e-12 * max(1.0, float(g[tag + '_rho']))
    assert np.array_equal(K == 0.0, g[tag + '_K'] == 0.0)
    e = [relerr(x, g[tag + '_x']), relerr(u, g[tag + '_u']), relerr(K, g[tag + '_K'])]
    sens = g[tag + '_ulp_sensitivity']
    tol = [max(TOL, 10.0 * float(v)) for v in sens]
    with capsys.disabled():
        print("\n[non-PD %s] %d iterations: x %.1e u %.1e K %.1e | 1-ulp sensitivity of the reference %s -> tol %s"
              % (tag, it, *e, ['%.1e' % v for v in sens], ["%.1e" % t for t in tol]))
    if it <= 12:
        assert max(tol) == TOL                                                 # the short cases ARE held to 1e-9
    assert e[0] < tol[0] and e[1] < tol[1] and e[2] < tol[2]


def test_non_pd_backward_pass_unit_matches_reference_golden(golden):
    """dlqr_recursion alone on the non-PD case: K, k, Q_u, Q_uu with the reference's zero pattern (gains zero at and
    below the failing step, Q_u / Q_uu zero strictly below it), rho raised once then lowered once."""
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost as QC
    g = golden("ilqr_nonpd.npz")
    s_, model = _ssm(8)
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    Q = Q.copy(); Q[2, 2] = -500.0
    s = iLQR(0.02, model, QC(Q, R, Qf), 40)
    s.set_target(synth.figure8_targets(s_['z_ref'], 40, 6.0)[0])
    s.rho, s.drho = 0.0, 0.0
    K, k, Qu, Quu = s.dlqr_recursion(g['unit_x'], g['unit_u'], g['unit_A'], g['unit_B'], g['unit_d'])
    tf = int(g['unit_pd_fail_t'])
    assert int(np.ravel(s.info['pd_fail_step'])[0]) == tf and tf >= 0
    assert not K[:tf + 1].any() and not k[:tf + 1].any() and not Qu[:tf].any() and not Quu[:tf].any()
    assert relerr(K, g['unit_K']) < TOL and relerr(k, g['unit_k']) < TOL
    assert relerr(Qu, g['unit_Qu']) < TOL and relerr(Quu, g['unit_Quu']) < TOL
    assert float(s.rho) == float(g['unit_rho']) and float(s.drho) == float(g['unit_drho'])


def test_branch_paths_max_iter_and_no_linesearch():
    import sofacontrol_b200.synth as synth
    s_, _ = _ssm(8)
    zt = synth.figure8_targets(s_['z_ref'], 30, 12.0, 0.7)[0]

    def tweak(p):
        p.max_iter = 1
        p.epsilon = 1e-12
    s, o, (x, u, K), x0 = _pair(8, 30, tweak, zt)
    xo, uo, Ko = o.ilqr_computation(x0)
    assert s.info['iterations'] == o.iterations == 2 and (s.info['status'] & 2)       # nbr_iter <= max_iter quirk: 2 passes
    assert relerr(x, xo) < TOL and relerr(u, uo) < TOL and relerr(K, Ko) < TOL

    def tweak2(p):
        p.do_linesearch = False
        p.max_iter = 3
        p.regularize = False
    s, o, (x, u, K), x0 = _pair(8, 30, tweak2, zt)
    xo, uo, Ko = o.ilqr_computation(x0)
    assert s.info['iterations'] == o.iterations
    assert relerr(x, xo) < TOL and relerr(u, uo) < TOL and relerr(K, Ko) < TOL


def test_tpwl_forward_pass_two_stage_search_is_exact_on_ties():
    """The iLQR forward pass for TPWL-nn picks its linearisation with an FP32-screened search that must return
    EXACTLY the index of the full FP64 numpy-order search (tpwl.py:160-168), also when stored points coincide
    (first occurrence wins) or differ by less than FP32 can resolve.  Adversarial bank: duplicated and
    1e-13-perturbed points, trajectories started on top of them; the indices implied by the pass's A_t are compared
    with calc_nearest_point (bit-exact vs numpy, tests/test_tpwl_gpu.py) evaluated on the visited states."""
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.tpwl.tpwl import TPWLATV
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost
    data, Hf = synth.tpwl_bank()
    q = data['q']
    q[500] = q[100]                                  # exact tie: index 100 must win
    q[700] = q[200] * (1.0 + 1e-13)                  # below FP32 resolution: only the FP64 rescoring can tell
    q[701] = q[200] * (1.0 - 1e-13)
    q[3] = q[900]                                    # exact tie where the LOWER index is the later-made copy
    for i in range(1000):                            # make the linearisations distinguishable by A[0, 0]
        data['A_c'][i][0, 0] = -1.0 - i
    params = {'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}
    g = TPWLATV(data, params=params, Hf=Hf, discr_method='fe')
    g.pre_discretize(0.001)                          # the step's linearisation is then an index into the bank
    N, Bt = 12, 8
    rng = np.random.default_rng(5)
    x0 = np.zeros((Bt, 72))
    starts = [100, 500, 200, 700, 701, 900, 3, 42]
    for b, p in enumerate(starts):
        x0[b, 36:] = q[p] + (0.0 if b % 2 == 0 else 1e-9 * rng.normal(size=36))
        x0[b, :36] = 0.1 * rng.normal(size=36)
    Q = np.zeros((6, 6)); Q[3, 3] = Q[4, 4] = 100.0
    s = iLQR(0.001, g, QuadraticCost(Q, 1e-5 * np.eye(4), np.zeros((6, 6))), N)
    s.set_target(np.tile(g.z_ref, (Bt, N + 1, 1)))
    xp = np.zeros((Bt, N + 1, 72)); xp[:, 0] = x0
    x, u, cost, A, B, d = s.forward_pass(xp, np.zeros((Bt, N, 4)))
    A = np.asarray(A).reshape(Bt, N, 72, 72)
    x = np.asarray(x).reshape(Bt, N + 1, 72)
    idx_pass = np.rint((A[:, :, 0, 0] - 1.0) / 0.001 * -1.0 - 1.0).astype(int)     # fe: A_d[0,0] = 1 + dt * (-1 - i)
    idx_full = np.asarray(g.calc_nearest_point(x[:, :-1].reshape(-1, 72))).reshape(Bt, N)
    assert np.array_equal(idx_pass, idx_full)
    assert idx_full[0, 0] == 100 and idx_full[1, 0] == 100 and idx_full[5, 0] == 3 and idx_full[6, 0] == 3
    # and against numpy itself on the first states
    for b in range(Bt):
        dist = 1.0 * np.linalg.norm(q - x0[b, 36:], axis=1) + 0.0 * np.linalg.norm(data['v'] - x0[b, :36], axis=1)
        assert idx_pass[b, 0] == int(np.argmin(dist))


def test_tpwl_diamond_instantiation_agrees_with_runtime_dimension_kernels():
    """Compile-time-dimension Diamond instantiation (DMMA views, single-thread factorisation) vs the run-time-dimension
    instantiation of the same templates (SRCB200_ILQR_GENERIC=1): same iteration counts, 1e-9 on x, u, K."""
    import os
    import sofacontrol_b200.synth as synth
    from sofacontrol_b200.tpwl.tpwl import TPWLATV
    from sofacontrol_b200.lqr.ilqr import iLQR
    from sofacontrol_b200.utils import QuadraticCost
    data, Hf = synth.tpwl_bank()
    g = TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='zoh')
    g.pre_discretize(0.01)
    N, Bt = 40, 6
    Q = np.zeros((6, 6)); Q[3, 3] = Q[4, 4] = 100.0
    th = np.linspace(0, 2 * np.pi, N + 1)
    x0, _ = synth.tpwl_rollout_batch(Bt, N=1, seed=33)
    zt = np.tile(g.z_ref, (Bt, N + 1, 1))
    amp = np.linspace(0.3, 1.5, Bt)
    zt[:, :, 3] += amp[:, None] * np.sin(th)[None]; zt[:, :, 4] += amp[:, None] * np.sin(2 * th)[None]
    res = {}
    for tag, env in (("fixed", "0"), ("runtime", "1")):
        os.environ["SRCB200_ILQR_GENERIC"] = env
        s = iLQR(0.01, g, QuadraticCost(Q, 1e-5 * np.eye(4), np.zeros((6, 6))), N)
        s.set_target(zt)
        res[tag] = s.ilqr_computation(x0) + (s.info['iterations'],)
    os.environ["SRCB200_ILQR_GENERIC"] = "0"
    assert np.array_equal(res["fixed"][3], res["runtime"][3])
    for a, b in zip(res["fixed"][:3], res["runtime"][:3]):
        assert relerr(a, b) < 1e-9


def test_headline_batch_members_match_reference_golden(golden, capsys):
    """The BENCH WORKLOAD at its own shape: the 4096-problem seed-3 Trunk-SSM batch (N = 100, m = 8) in ONE launch of
    the specialised kernel, 32 members compared with solves of the UNMODIFIED reference iLQR class
    (oracle/make_golden_bench.py: ilqr_bench_seed3.npz) -- the 5-iteration members, the 51-iteration member, six
    abandoned line searches (status 4, rho 1192.65) and both modes of the iteration histogram.
    Exact: iteration count, how the loop ended, final rho.  x, u, K: relative 1e-9, or -- for a member whose OWN
    float64 Riccati recursion is less accurate than that -- 8 x its measured noise floor (riccati_noise_floor:
    reference recursion in float64 vs 80-bit on the member's final trajectory), printed per member."""
    import sofacontrol_b200.synth as synth
    from oracle.ilqr_np import ILQRNP, riccati_noise_floor
    from oracle.utils_np import QuadraticCost
    g = golden("ilqr_bench_seed3.npz")
    w = synth.trunk_ilqr_batch(4096, N=100, seed=3, m=8)
    _, model = _ssm(8)
    s = _solver(model, 8, w['z_target'])
    x, u, K = s.ilqr_computation(w['x0'])
    mem = g['members']
    assert np.array_equal(s.info['iterations'][mem], g['iterations'])
    assert np.array_equal(s.info['status'][mem] & 7, g['status'])
    assert np.all(np.abs(s.info['rho'][mem] - g['rho']) <= 1e-12 * np.maximum(1.0, g['rho']))
    Q, R, Qf = synth.trunk_ilqr_costs(6, 8)
    om = _oracle_ssm(8)
    rows, worst = [], 0.0
    for j, b in enumerate(mem):
        ex, eu, eK = relerr(x[b], g['x'][j]), relerr(u[b], g['u'][j]), relerr(K[b], g['K'][j])
        tol = TOL
        if max(ex, eu, eK) >= TOL:
            o = ILQRNP(0.02, om, QuadraticCost(Q, R, Qf), 100)
            o.set_target(w['z_target'][b])
            xf, uf, _, Af, Bf, _ = o.forward_pass(g['x'][j], g['u'][j])
            fK, fk = riccati_noise_floor(om, QuadraticCost(Q, R, Qf), w['z_target'][b], xf, uf, Af, Bf)
            tol = max(TOL, 8.0 * max(fK, fk))
            rows.append("member %4d (%2d it): x %.1e u %.1e K %.1e | float64 Riccati floor K %.1e k %.1e -> tol %.1e"
                        % (b, g['iterations'][j], ex, eu, eK, fK, fk, tol))
        worst = max(worst, ex, eu, eK)
        assert ex < tol and eu < tol and eK < tol, rows[-1] if rows else (b, ex, eu, eK)
    with capsys.disabled():
        print("\n[headline parity] 32 members, worst rel err %.2e; members above 1e-9: %d" % (worst, len(rows)))
        for r in rows:
            print("   ", r)
    # every problem of the batch ended in one of the reference's three ways and returned finite results
    st = s.info['status']
    assert np.all((st & 7) != 0) and np.all(np.isfinite(x)) and np.all(np.isfinite(K))
