"""FP64 numpy restatement of sofacontrol/SSM/ssm.py.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
PINNED against the reference file itself, imported unmodified on oracle/jax_shim.py
(tests/test_oracle_vs_reference.py::test_ssm_restatement_vs_unmodified_reference_class, golden ssm_units.npz /
ssm_module_test.npz / ssm_ilqr.npz produced by the reference class).

Monomial order (ssm.py:158-164): sympy `itermonomials(vars, order)` sorted by grevlex on reversed variables with the
constant dropped == for d = 1..order: itertools.combinations_with_replacement(range(dim), d)
(verified against sympy and against the reference's lambdified basis in tests/test_oracle_vs_reference.py).  Jacobians (ssm.py:198-235, jax.jacobian in the reference)
are the analytic  coeff @ dphi/dx.
"""
import itertools
import numpy as np


def monomial_index_table(dim, order):
    """Rows of variable indices (padded with -1) for every monomial of total degree 1..order."""
    rows = []
    for d in range(1, order + 1):
        for c in itertools.combinations_with_replacement(range(dim), d):
            rows.append(tuple(c) + (-1,) * (order - d))
    return np.array(rows, dtype=np.int64)


def poly_features(x, table):
    """phi(x): x is (dim,) or (dim, N) like the lambdified basis of ssm.py:164.  Products are taken left to right
    over the sorted variable indices."""
    x = np.asarray(x, dtype=np.float64)
    out = []
    for row in table:
        v = x[row[0]]
        for j in row[1:]:
            if j >= 0:
                v = v * x[j]
        out.append(v)
    return np.array(out)


def poly_features_jac(x, table):
    """dphi/dx at a single point x (dim,): (nfeat, dim).  d/dx_j of a monomial = multiplicity * product of the
    remaining factors (left to right), the multiplicity applied last."""
    x = np.asarray(x, dtype=np.float64)
    dim = x.shape[0]
    J = np.zeros((table.shape[0], dim))
    for k, row in enumerate(table):
        idx = [j for j in row if j >= 0]
        for j in set(idx):
            mult = idx.count(j)
            rest = list(idx)
            rest.remove(j)
            v = 1.0
            for r in rest:
                v = v * x[r]
            J[k, j] = mult * v
    return J


class SSMDynamicsNP:
    """Same public surface as ssm.SSMDynamics (ssm.py:18-344) on plain numpy."""

    def __init__(self, eq_point, discrete=False, discr_method='fe', **kwargs):
        # ssm.py:24-74
        self.maps = {}
        self.discrete = discrete
        self.discr_method = discr_method
        self.model = kwargs.pop('model', None)
        self.params = kwargs.pop('params', None)
        g = lambda s, k: s[k][0, 0][0, 0]
        self.state_dim = int(g(self.params, 'state_dim'))
        self.input_dim = int(g(self.params, 'input_dim'))
        self.output_dim = int(g(self.params, 'output_dim'))
        self.SSM_order = int(g(self.params, 'SSM_order'))
        self.ROM_order = int(g(self.params, 'ROM_order'))
        self.Ts = g(self.model, 'Ts')
        self.rom_table = monomial_index_table(self.state_dim, self.ROM_order)
        self.ssm_table = monomial_index_table(self.output_dim, self.SSM_order)
        m = lambda k: np.asarray(self.model[k][0, 0], dtype=np.float64)
        self.w_coeff, self.v_coeff, self.r_coeff, self.B_r = m('w_coeff'), m('v_coeff'), m('r_coeff'), m('B')
        self.rd_coeff, self.Bd_r = m('rd_coeff'), m('Bd')
        self.C_map = self.reduced_to_observed
        self.W_map = self.observed_to_reduced
        self.maps['f_nl'] = self.reduced_dynamics
        if self.discrete:
            self.maps['f_nl_d'] = self.reduced_dynamics_discrete
        self.z_ref = eq_point
        self.A_d = self.B_d = self.d_d = None
        self.H = np.zeros((self.output_dim, self.state_dim))
        self.nonlinear_observer = True

    # ssm.py:83-101
    def zfyf_to_zy(self, zf=None):
        if zf is not None and self.z_ref is not None:
            return zf - self.z_ref
        raise RuntimeError('Need to specify equilibrium point')

    def zy_to_zfyf(self, z=None):
        if z is not None and self.z_ref is not None:
            return z + self.z_ref
        raise RuntimeError('Need to specify equilibrium point')

    # ssm.py:105-119
    def x_to_zfyf(self, x, zf=True):
        return self.C_map(x.T).T + self.z_ref

    def x_to_zy(self, x):
        return self.C_map(x)

    def get_state_dim(self):
        return self.state_dim

    def get_input_dim(self):
        return self.input_dim

    def get_output_dim(self):
        return self.output_dim

    def get_ref_point(self):
        return self.z_ref

    # ssm.py:167-178
    def reduced_dynamics(self, x, u):
        return np.dot(self.r_coeff, poly_features(x, self.rom_table)) + np.dot(self.B_r, u)

    def reduced_to_observed(self, x):
        return np.dot(self.w_coeff, poly_features(x, self.ssm_table))

    def observed_to_reduced(self, z):
        return np.dot(self.v_coeff, poly_features(z, self.ssm_table))

    def reduced_dynamics_discrete(self, x, u):
        return np.dot(self.rd_coeff, poly_features(x, self.rom_table)) + np.dot(self.Bd_r, u)

    # ssm.py:198-212
    def get_continuous_jacobians(self, x, u):
        A = np.dot(self.r_coeff, poly_features_jac(x, self.rom_table))
        B = self.B_r
        d = self.reduced_dynamics(x, u) - np.dot(A, x) - np.dot(B, u)
        return A, B, d

    def get_discrete_jacobians(self, x, u):
        A = np.dot(self.rd_coeff, poly_features_jac(x, self.rom_table))
        B = self.Bd_r
        d = self.reduced_dynamics_discrete(x, u) - np.dot(A, x) - np.dot(B, u)
        return A, B, d

    # ssm.py:215-225
    def get_jacobians(self, x, u, dt):
        x = np.asarray(x, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        if not self.discrete:
            Ac, Bc, dc = self.get_continuous_jacobians(x, u)
            return self.discretize_dynamics(Ac, Bc, dc, dt)
        return self.get_discrete_jacobians(x, u)

    # ssm.py:228-235
    def get_observer_jacobians(self, x):
        H = np.dot(self.w_coeff, poly_features_jac(x, self.ssm_table))
        c_res = self.C_map(x) - np.dot(H, x)
        return H, c_res

    # ssm.py:271-277
    def update_observer_state(self, x, dt=None, u=None):
        H, c = self.get_observer_jacobians(x)
        return np.squeeze(np.dot(H, x)) + np.squeeze(c)

    # ssm.py:279-301
    def discretize_dynamics(self, A_c, B_c, d_c, dt):
        I = np.eye(A_c.shape[0])
        if self.discr_method == 'fe':
            return I + dt * A_c, dt * B_c, dt * d_c
        if self.discr_method == 'be':
            A_d = np.linalg.inv(I - dt * A_c)
        elif self.discr_method == 'bil':
            A_d = np.dot(I + 0.5 * dt * A_c, np.linalg.inv(I - 0.5 * dt * A_c))
        else:
            raise RuntimeError('self.discr_method must be in [fe, be, bil, zoh]')
        sep = np.dot(np.linalg.inv(A_c), A_d - I)
        return A_d, np.dot(sep, B_c), np.dot(sep, d_c)

    # ssm.py:187-195, 330-333
    def update_state(self, x, u, dt):
        A_d, B_d, d_d = self.get_jacobians(x, dt=dt, u=u)
        return self.update_dynamics(x, u, A_d, B_d, d_d)

    @staticmethod
    def update_dynamics(x, u, A_d, B_d, d_d):
        return np.squeeze(A_d @ x) + np.squeeze(B_d @ u) + np.squeeze(d_d)

    # ssm.py:134-156
    def rollout(self, x0, u, dt):
        N = u.shape[0]
        x = np.zeros((N + 1, self.state_dim))
        x[0, :] = x0
        for i in range(N):
            x[i + 1, :] = self.update_state(x[i, :], u[i, :], dt)
        return x, self.x_to_zfyf(x)

    # ssm.py:338-344
    def compute_RO_state(self, z):
        return self.W_map(z - self.z_ref)


class GaussNewtonSSM:
    """Adapter of SURVEY.md Appendix C.2: lets the UNMODIFIED reference iLQR (ilqr.py) do Gauss-Newton tracking on an
    SSM.  `H` is a property returning dC/dx at the x of the most recent x_to_zfyf call -- the reference always calls
    x_to_zfyf immediately before reading model.H (ilqr.py:164-190)."""

    def __init__(self, ssm):
        self.ssm = ssm
        self._H = np.zeros((ssm.output_dim, ssm.state_dim))

    def get_state_dim(self):
        return self.ssm.state_dim

    def get_input_dim(self):
        return self.ssm.input_dim

    def get_jacobians(self, x, u=None, dt=None):
        return self.ssm.get_jacobians(x, u, dt)

    def update_dynamics(self, x, u, A, B, d):
        return self.ssm.update_dynamics(x, u, A, B, d)

    def x_to_zfyf(self, x, zf=True):
        if x.ndim == 1:
            self._H = self.ssm.get_observer_jacobians(x)[0]
        return self.ssm.x_to_zfyf(x)

    @property
    def H(self):
        return self._H


def mat_structs(n, m, nz, order, r_coeff, w_coeff, v_coeff, B, Ts=0.01, rd_coeff=None, Bd=None):
    """Builds the (1,1) MATLAB-struct-array layout that `loadmat(...)['py_data'][0,0]` yields
    (examples/hardware/diamond_SSM.py:99-102) from plain arrays."""
    def box(a):
        o = np.empty((1, 1), dtype=object)
        o[0, 0] = a
        return o
    model = np.empty((1, 1), dtype=[(k, 'O') for k in
                                    ('w_coeff', 'v_coeff', 'r_coeff', 'B', 'Ts', 'rd_coeff', 'Bd')])
    model['w_coeff'][0, 0] = w_coeff
    model['v_coeff'][0, 0] = v_coeff
    model['r_coeff'][0, 0] = r_coeff
    model['B'][0, 0] = B
    model['Ts'][0, 0] = np.array([[Ts]])
    model['rd_coeff'][0, 0] = r_coeff if rd_coeff is None else rd_coeff
    model['Bd'][0, 0] = B if Bd is None else Bd
    params = np.empty((1, 1), dtype=[(k, 'O') for k in
                                     ('SSM_order', 'ROM_order', 'state_dim', 'input_dim', 'output_dim')])
    for k, v in (('SSM_order', order), ('ROM_order', order), ('state_dim', n), ('input_dim', m),
                 ('output_dim', nz)):
        params[k][0, 0] = np.array([[v]], dtype=np.uint8)
    return model, params
