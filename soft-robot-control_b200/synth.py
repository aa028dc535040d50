"""Seeded synthetic models and workloads of the shapes BASELINE.json names.

Used by bench.py, the tests and __graft_entry__.smoke() to build identical inputs for the CUDA path and for the
CPU oracle / reference.  Pure numpy data generation -- no hot-path numerics.
"""
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIAMOND_SSM_FIXTURE = os.path.join(REPO, "tests", "golden", "ssm_diamond_model.npz")


# ------------------------------------------------------------------------------------------------------------
# MATLAB-struct plumbing: the SSM classes take `model` / `params` the way loadmat(...)['py_data'][0,0] yields them
# (examples/hardware/diamond_SSM.py:99-102)
# ------------------------------------------------------------------------------------------------------------
def ssm_structs(n, m, nz, order, r_coeff, w_coeff, v_coeff, B, Ts=0.01, rd_coeff=None, Bd=None):
    model = np.empty((1, 1), dtype=[(k, 'O') for k in ('w_coeff', 'v_coeff', 'r_coeff', 'B', 'Ts', 'rd_coeff', 'Bd')])
    model['w_coeff'][0, 0] = np.asarray(w_coeff, dtype=np.float64)
    model['v_coeff'][0, 0] = np.asarray(v_coeff, dtype=np.float64)
    model['r_coeff'][0, 0] = np.asarray(r_coeff, dtype=np.float64)
    model['B'][0, 0] = np.asarray(B, dtype=np.float64)
    model['Ts'][0, 0] = np.array([[Ts]])
    model['rd_coeff'][0, 0] = np.asarray(r_coeff if rd_coeff is None else rd_coeff, dtype=np.float64)
    model['Bd'][0, 0] = np.asarray(B if Bd is None else Bd, dtype=np.float64)
    params = np.empty((1, 1), dtype=[(k, 'O') for k in ('SSM_order', 'ROM_order', 'state_dim', 'input_dim', 'output_dim')])
    for k, v in (('SSM_order', order), ('ROM_order', order), ('state_dim', n), ('input_dim', m), ('output_dim', nz)):
        params[k][0, 0] = np.array([[v]], dtype=np.uint8)
    return model, params


def num_monomials(dim, order):
    from math import comb
    return sum(comb(dim + d - 1, d) for d in range(1, order + 1))


def diamond_ssm_fixture():
    """The real Diamond SSM coefficients (examples/hardware/SSMmodels/SSM_model.mat) + equilibrium output, as
    exported to tests/golden/ssm_diamond_model.npz by oracle/make_golden.py.  n = n_z = 6, m = 4, order 3."""
    d = np.load(DIAMOND_SSM_FIXTURE)
    return {k: d[k] for k in d.files}


def trunk_ssm(m=8):
    """BASELINE config 1/3 'Trunk SSM' (n = n_z = 6, 8 cable inputs, order 3): no Trunk model ships with the
    reference, so the Diamond fixture coefficients are widened deterministically to 8 inputs,
    B8 = [B, 0.7 * B[:, ::-1]].  Returns dict(model, params, z_ref, dt, ...)."""
    f = diamond_ssm_fixture()
    B, Bd = f['B'], f['Bd']
    if m == 8:
        B = np.hstack((B, 0.7 * B[:, ::-1]))
        Bd = np.hstack((Bd, 0.7 * Bd[:, ::-1]))
    elif m != 4:
        raise ValueError("m must be 4 (Diamond) or 8 (Trunk)")
    model, params = ssm_structs(6, m, 6, 3, f['r_coeff'], f['w_coeff'], f['v_coeff'], B, Ts=float(f['Ts']),
                                rd_coeff=f['rd_coeff'], Bd=Bd)
    return dict(model=model, params=params, z_ref=f['z_eq'].copy(), n=6, m=m, nz=6)


def synthetic_ssm(seed=0, n=6, m=8, order=3):
    """Fully synthetic polynomial SSM (secondary workload): linear part = damped oscillator pairs in modal form,
    small quadratic/cubic terms, near-identity observation map."""
    rng = np.random.default_rng(seed)
    nf = num_monomials(n, order)
    h = n // 2
    r = np.zeros((n, nf))
    om = 2 * np.pi * np.array([1.0, 2.5, 4.0, 5.5])[:h]
    zeta = 0.05
    # state = [positions (h); velocities (h)]:  qdot = v, vdot = -om^2 q - 2 zeta om v
    for i in range(h):
        r[i, h + i] = 1.0
        r[h + i, i] = -om[i] ** 2
        r[h + i, h + i] = -2 * zeta * om[i]
    nq = num_monomials(n, 2) - n
    r[:, n:n + nq] += rng.normal(0, 1e-3, size=(n, nq))
    r[:, n + nq:] += rng.normal(0, 1e-5, size=(n, nf - n - nq))
    w = np.zeros((n, nf)); w[:, :n] = np.eye(n)
    w[:, n:n + nq] += rng.normal(0, 1e-3, size=(n, nq))
    w[:, n + nq:] += rng.normal(0, 1e-5, size=(n, nf - n - nq))
    v = np.zeros((n, nf)); v[:, :n] = np.eye(n)
    v[:, n:n + nq] -= w[:, n:n + nq]
    B = np.zeros((n, m)); B[h:, :] = rng.normal(0, 0.05, size=(n - h, m))
    model, params = ssm_structs(n, m, n, order, r, w, v, B)
    z_ref = np.concatenate((rng.normal(0, 10, size=h), np.zeros(n - h)))
    return dict(model=model, params=params, z_ref=z_ref, n=n, m=m, nz=n)


def figure8_targets(z_ref, N, amp, phase=0.0):
    """z* = z_ref + [-a sin(th), a sin(2 th), 0, ...], th in [phase, phase + 2 pi]  (the figure-8 of
    examples/hardware/diamond_SSM.py).  amp / phase may be arrays (batch) -> (Bt, N+1, n_z)."""
    amp = np.atleast_1d(np.asarray(amp, dtype=np.float64))
    phase = np.broadcast_to(np.atleast_1d(np.asarray(phase, dtype=np.float64)), amp.shape)
    th = np.linspace(0, 2 * np.pi, N + 1)[None, :] + phase[:, None]
    zt = np.tile(np.asarray(z_ref, dtype=np.float64), (amp.shape[0], N + 1, 1))
    zt[:, :, 0] += -amp[:, None] * np.sin(th)
    zt[:, :, 1] += amp[:, None] * np.sin(2 * th)
    return zt


def trunk_ilqr_costs(nz=6, m=8):
    """Q = diag(100,100,0,...), R = 0.003 I, Qf = 0 (examples/hardware/diamond_SSM.py:199-204)."""
    Q = np.zeros((nz, nz)); Q[0, 0] = 100.0; Q[1, 1] = 100.0
    return Q, 0.003 * np.eye(m), np.zeros((nz, nz))


def trunk_ilqr_batch(batch, N=100, seed=3, m=8):
    """BASELINE config 3: randomised figure-8 amplitude/phase and initial conditions for `batch` problems."""
    rng = np.random.default_rng(seed)
    s = trunk_ssm(m)
    amp = rng.uniform(2.0, 15.0, size=batch)
    phase = rng.uniform(0.0, 2 * np.pi, size=batch)
    zt = figure8_targets(s['z_ref'], N, amp, phase)
    x0 = np.zeros((batch, 6))
    x0[:, :3] = rng.uniform(-0.5, 0.5, size=(batch, 3))     # small reduced-coordinate offsets
    return dict(ssm=s, z_target=zt, x0=x0, dt=0.02, N=N)


# ------------------------------------------------------------------------------------------------------------
# Diamond-shaped TPWL bank (config 2): r = 36, n = 72, m = 4, P = 1000
# ------------------------------------------------------------------------------------------------------------
def tip_output_matrix(node, num_nodes):
    """Dense Hf of measurement_models.linearModel([node], num_nodes).C (velocity rows then position rows of one
    node; measurement_models.py:29-37, 87-103): (6, 6 num_nodes)."""
    Hf = np.zeros((6, 6 * num_nodes))
    for k in range(3):
        Hf[k, 3 * node + k] = 1.0
        Hf[3 + k, 3 * num_nodes + 3 * node + k] = 1.0
    return Hf


def tpwl_bank(seed=0, r=36, m=4, P=1000, num_nodes=1628, tip_node=1354, spread=5.0):
    """Synthetic TPWL dict with the reference's schema: keys q, v, u, A_c, B_c, d_c, rom_info.
    Mass-normalised stiffness K_i (eigenvalues 50..5000 perturbed 5 % per point), Rayleigh damping
    D_i = 2.5 I + 0.01 K_i (examples/hardware/model.py:14-15), A_i = [[-D_i, -K_i], [I, 0]], B_i = [H_i; 0],
    d_i = [0.01 K_i q_i; 0]; state x = [v; q]."""
    rng = np.random.default_rng(seed)
    Qm, _ = np.linalg.qr(rng.normal(size=(r, r)))
    K0 = Qm @ np.diag(np.linspace(50.0, 5000.0, r)) @ Qm.T
    H0 = rng.normal(0, 1e-2, size=(r, m))
    q = rng.normal(0, spread, size=(P, r))
    v = rng.normal(0, 4 * spread, size=(P, r))
    u = rng.uniform(0, 1500, size=(P, m))
    A = np.zeros((P, 2 * r, 2 * r)); B = np.zeros((P, 2 * r, m)); d = np.zeros((P, 2 * r))
    I = np.eye(r)
    for i in range(P):
        S = rng.normal(size=(r, r)); S = 0.5 * (S + S.T)
        Ki = K0 * (1.0 + 0.05 * S)
        Di = 2.5 * I + 0.01 * Ki
        A[i, :r, :r] = -Di
        A[i, :r, r:] = -Ki
        A[i, r:, :r] = I
        B[i, :r, :] = H0 * (1.0 + 0.05 * rng.normal(size=(r, m)))
        d[i, :r] = 0.01 * Ki @ q[i]
    nf = 3 * num_nodes
    U, _ = np.linalg.qr(rng.normal(size=(nf, r)))
    rom_info = {'U': U, 'q_ref': rng.normal(0, 50, size=nf), 'v_ref': np.zeros(nf), 'type': 'POD'}
    data = {'q': q, 'v': v, 'u': u, 'A_c': A, 'B_c': B, 'd_c': d, 'rom_info': rom_info}
    return data, tip_output_matrix(tip_node, num_nodes)


def tpwl_rollout_batch(batch, N=100, seed=2, r=36, m=4, spread=5.0):
    """x0 and inputs for config 2: positions spread like the stored points so trajectories cross several regions."""
    rng = np.random.default_rng(seed)
    x0 = np.concatenate((rng.normal(0, 1.0, size=(batch, r)), rng.normal(0, spread, size=(batch, r))), axis=1)
    u = rng.uniform(0, 1500, size=(batch, N, m))
    return x0, u


# ------------------------------------------------------------------------------------------------------------
# POD snapshots with a prescribed spectrum (config 5)
# ------------------------------------------------------------------------------------------------------------
def pod_spectrum(ns):
    """A decaying singular spectrum shaped like the Diamond fixture's Sigma (1.4e4 ... 1e-11 over 3061 values)."""
    i = np.arange(ns, dtype=np.float64)
    return 1.4e4 * np.exp(-0.16 * np.minimum(i, 60.0)) * np.exp(-0.012 * np.maximum(i - 60.0, 0.0)) + 1e-9


def pod_snapshots(nf, ns, seed=5, rank=None):
    """X = Uo diag(s) Vo^T with orthonormal factors (host, small scale): known singular values/vectors."""
    rng = np.random.default_rng(seed)
    k = min(nf, ns) if rank is None else rank
    Uo, _ = np.linalg.qr(rng.normal(size=(nf, k)))
    Vo, _ = np.linalg.qr(rng.normal(size=(ns, k)))
    s = pod_spectrum(k)
    return (Uo * s) @ Vo.T, Uo, s


# ------------------------------------------------------------------------------------------------------------
# Closed-loop Monte Carlo (config 4): "Diamond SSM + TPWL"
# ------------------------------------------------------------------------------------------------------------
def mpc_ssm_tpwl_workload(batch, steps=100, N=20, seed=4, dt=0.02):
    """Plant = the Diamond-shaped TPWL bank (config 2; nearest-neighbour, zoh pre-discretised at the control period),
    controller = receding-horizon iLQR on the Diamond SSM (m = 4) fed by the SSM observer (tip output [v; q] ->
    [q; v] -> W_map).  The plant's reference configuration is placed so that its tip output at rest equals the SSM's
    equilibrium output (the two synthetic models then describe the same operating point).  Returns the model
    objects, the solver and seeded initial conditions / figure-8 references."""
    from .SSM.ssm import SSMDynamics
    from .tpwl.tpwl import TPWLATV
    from .lqr.ilqr import iLQR
    from .utils import QuadraticCost
    rng = np.random.default_rng(seed)
    s = trunk_ssm(4)
    ssm = SSMDynamics(s['z_ref'], discrete=False, discr_method='be', model=s['model'], params=s['params'])
    data, Hf = tpwl_bank()
    tip = 1354
    data['rom_info']['q_ref'] = data['rom_info']['q_ref'].copy()
    data['rom_info']['q_ref'][3 * tip:3 * tip + 3] = s['z_ref'][:3]
    plant = TPWLATV(data, params={'tpwl_method': 'nn', 'dist_weights': {'q': 1.0, 'v': 0.0}}, Hf=Hf, discr_method='zoh')
    plant.pre_discretize(dt)
    Q, R, Qf = trunk_ilqr_costs(6, 4)
    solver = iLQR(dt, ssm, QuadraticCost(Q, R, Qf), N)
    T = steps + N
    amp, ph = rng.uniform(1.0, 4.0, size=batch), rng.uniform(0, 2 * np.pi, size=batch)
    th = np.linspace(0, 2 * np.pi * T / 100.0, T + 1)[None, :] + ph[:, None]
    z_ref = np.tile(s['z_ref'], (batch, T + 1, 1))
    z_ref[:, :, 0] += -amp[:, None] * np.sin(th)
    z_ref[:, :, 1] += amp[:, None] * np.sin(2 * th)
    x0_plant = np.concatenate((rng.normal(0, 0.05, size=(batch, 36)), rng.normal(0, 1.0, size=(batch, 36))), axis=1)
    zp = plant_output_host(plant, x0_plant)
    x0_belief = ssm_belief_host(s, zp)
    solver.set_target(z_ref[:, :N + 1])
    return dict(ssm=ssm, plant=plant, solver=solver, z_ref=z_ref, x0_plant=x0_plant, x0_belief=x0_belief, dt=dt)


def plant_output_host(plant, x):
    """z = H x + z_ref of a TPWL model on host arrays (workload construction only)."""
    return x @ np.asarray(plant.H).T + np.asarray(plant.z_ref)


def ssm_belief_host(s, z_vq):
    """Initial belief of the SSM controller from a tip output in [v; q] order: W_map(vq2qv(z) - z_ref) on the host
    (workload construction only; in the loop this is the batched map kernel)."""
    f = diamond_ssm_fixture()
    h = z_vq.shape[-1] // 2
    dz = np.concatenate((z_vq[:, h:], z_vq[:, :h]), axis=1) - s['z_ref']
    from itertools import combinations_with_replacement
    feats = []
    for d in (1, 2, 3):
        for c in combinations_with_replacement(range(6), d):
            feats.append(np.prod(dz[:, list(c)], axis=1))
    return np.stack(feats, axis=1) @ f['v_coeff'].T
