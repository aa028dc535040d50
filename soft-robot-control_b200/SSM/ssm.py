"""SSM polynomial reduced-order model on B200 -- drop-in for sofacontrol/SSM/ssm.py (SSM, SSMDynamics).

Same constructor, attributes and method names/argument meaning as the reference classes (ssm.py:18-344); every
numerical method runs the CUDA kernels of libsrcb200 (csrc/ssm.cu) and returns host numpy FP64 arrays with the
reference's shapes.  Extension: every method also accepts a leading batch axis ((Bt, n) states, (Bt, N, m)
inputs) and `*_device` variants keep inputs/outputs as CUDA torch tensors.

Differences, on purpose:
  * FP64 throughout (the reference runs JAX at its float32 default; BASELINE.json asks for FP64 parity with the
    numpy restatement in oracle/ssm_np.py).
  * Jacobians are the analytic coeff @ dphi/dx instead of jax.jacobian.
"""
import itertools

import numpy as np

from .. import _lib as L

###  DEFAULT VALUES (ssm.py:11-15)
DISCR_METHOD = 'zoh'
TPWL_METHOD = 'nn'
DISCR_DICT = {'fe': 'forward Euler', 'be': 'implicit Euler', 'bil': 'bilinear transform', 'zoh': 'zero-order hold'}


def monomial_table(dim, order):
    """Variable-index rows of every monomial of degree 1..order in the order of SSM.get_poly_basis (ssm.py:158-164:
    sympy itermonomials sorted grevlex on reversed variables, constant dropped), padded with 0xFF to 4 columns."""
    rows = []
    for d in range(1, order + 1):
        for c in itertools.combinations_with_replacement(range(dim), d):
            rows.append(list(c) + [0xFF] * (L.SSM_MAX_ORDER - d))
    return np.array(rows, dtype=np.uint8)


def _field(struct, key):
    """MATLAB (1,1) struct-array access used by the reference: params['state_dim'][0, 0][0, 0] (ssm.py:33-51)."""
    return struct[key][0, 0]


class SSM:
    """ssm.py:18-178."""

    def __init__(self, eq_point, discrete=False, discr_method='fe', **kwargs):
        self.maps = {}
        self.discrete = discrete
        self.discr_method = discr_method
        self.model = kwargs.pop('model', None)
        self.params = kwargs.pop('params', None)

        self.state_dim = int(_field(self.params, 'state_dim')[0, 0])
        self.input_dim = int(_field(self.params, 'input_dim')[0, 0])
        self.output_dim = int(_field(self.params, 'output_dim')[0, 0])
        self.SSM_order = int(_field(self.params, 'SSM_order')[0, 0])
        self.ROM_order = int(_field(self.params, 'ROM_order')[0, 0])
        self.Ts = _field(self.model, 'Ts')[0, 0]
        if self.SSM_order != self.ROM_order or self.output_dim != self.state_dim:
            # the reference applies the output basis (output_dim vars, SSM_order) to x (ssm.py:39-40, 170-174)
            raise NotImplementedError("sofacontrol_b200 SSM needs output_dim == state_dim and SSM_order == ROM_order")
        self.mono_table = monomial_table(self.state_dim, self.ROM_order)

        f64 = lambda k: np.ascontiguousarray(_field(self.model, k), dtype=np.float64)
        self.w_coeff = f64('w_coeff')   # reduced to observed
        self.v_coeff = f64('v_coeff')   # observed to reduced
        self.r_coeff = f64('r_coeff')   # reduced coefficients
        self.B_r = f64('B')             # reduced control matrix
        self.rd_coeff = f64('rd_coeff')
        self.Bd_r = f64('Bd')

        self.C_map = self.reduced_to_observed
        self.W_map = self.observed_to_reduced
        self.maps['f_nl'] = self.reduced_dynamics
        if self.discrete:
            self.maps['f_nl_d'] = self.reduced_dynamics_discrete

        self.z_ref = eq_point
        self.A_d = None
        self.B_d = None
        self.d_d = None
        self.H = np.zeros((self.output_dim, self.state_dim))
        self.nonlinear_observer = True
        self._dev = {}

    # ---- device model handles ---------------------------------------------------------------------------------
    def _handle(self, kind):
        """kind: 'cont' (r_coeff, B_r, discretised by discr_method), 'cont_raw' (no discretisation),
        'disc' (rd_coeff, Bd_r)."""
        # the device copies capture the coefficient arrays: key them on a content fingerprint (a few KB, microseconds)
        # so a model whose coefficients / z_ref were modified after first use is uploaded again
        fp = hash(b"".join(np.ascontiguousarray(a, dtype=np.float64).tobytes() for a in
                           (self.rd_coeff if kind == 'disc' else self.r_coeff, self.Bd_r if kind == 'disc' else self.B_r,
                            self.w_coeff, self.v_coeff, np.asarray(self.z_ref, dtype=np.float64).reshape(-1))))
        key = (kind, self.discr_method)
        if key in self._dev and self._dev[key][2] == fp:
            return self._dev[key][0]
        L.require_gpu()
        torch = L.torch_mod()
        if kind == 'disc':
            r, B, method = self.rd_coeff, self.Bd_r, 'none'
        else:
            r, B = self.r_coeff, self.B_r
            method = self.discr_method if kind == 'cont' else 'fe'
            if method not in ('fe', 'be', 'bil'):
                raise RuntimeError('self.discr_method must be in [fe, be, bil, zoh]')   # ssm.py:299
        bufs = dict(r=L.to_dev(r), w=L.to_dev(self.w_coeff), v=L.to_dev(self.v_coeff), B=L.to_dev(B),
                    z=L.to_dev(np.asarray(self.z_ref, dtype=np.float64).reshape(-1)),
                    mono=L.to_dev(self.mono_table, torch.uint8))
        h = L.SsmModel(n=self.state_dim, m=self.input_dim, nz=self.output_dim, order=self.ROM_order,
                       nfeat=self.mono_table.shape[0], discr_method=L.DISCR[method],
                       r_coeff=L.ptr(bufs['r']), w_coeff=L.ptr(bufs['w']), v_coeff=L.ptr(bufs['v']),
                       B_r=L.ptr(bufs['B']), z_ref=L.ptr(bufs['z']), mono=L.ptr(bufs['mono']))
        self._dev[key] = (h, bufs, fp)
        return h

    def device_model(self):
        """The srcb200_ssm_model the batched solvers consume (discrete or continuous+discr_method)."""
        return self._handle('disc' if self.discrete else 'cont')

    def update_state(self, x, u, dt):
        raise NotImplementedError("update_state must be overriden by a child class")

    def get_jacobians(self, x, u, dt):
        raise NotImplementedError("get_jacobians must be overriden by a child class")

    # ---- shifts (ssm.py:83-101)
    def zfyf_to_zy(self, zf=None):
        if zf is not None and self.z_ref is not None:
            return zf - self.z_ref
        raise RuntimeError('Need to specify equilibrium point')

    def zy_to_zfyf(self, z=None):
        if z is not None and self.z_ref is not None:
            return z + self.z_ref
        raise RuntimeError('Need to specify equilibrium point')

    # ---- polynomial maps ------------------------------------------------------------------------------------
    def _map_device(self, which, add_ref, pts, u=None, kind=None):
        """pts: CUDA tensor (count, n) -> (count, n)."""
        out = L.empty(pts.shape)
        kind = kind or ('disc' if self.discrete else 'cont_raw')
        L.check(L.lib().srcb200_ssm_map_batch(self._handle(kind), which, int(add_ref), pts.shape[0], L.ptr(pts),
                                              L.ptr(u), L.ptr(out), L.stream_ptr()))
        return out

    def _dyn_map(self, x, u, kind):
        x = np.asarray(x, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        out = self._map_device(2, False, L.to_dev(x.reshape(-1, self.state_dim)),
                               L.to_dev(u.reshape(-1, self.input_dim)), kind)
        return L.to_host(out).reshape(x.shape)

    def _map_cols(self, which, x):
        """Reference convention of C_map / W_map: x is (n,) or (n, N) with points in columns (ssm.py:103-104)."""
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            return L.to_host(self._map_device(which, False, L.to_dev(x[None, :])))[0]
        return L.to_host(self._map_device(which, False, L.to_dev(x.T))).T

    def x_to_zfyf(self, x, zf=True):
        """(N, n_x) or (n_x,) -> C_map(x) + z_ref  (ssm.py:105-111)."""
        x = np.asarray(x, dtype=np.float64)
        pts = L.to_dev(x.reshape(-1, self.state_dim))
        return L.to_host(self._map_device(0, True, pts)).reshape(x.shape[:-1] + (self.output_dim,))

    def x_to_zy(self, x):
        """ssm.py:113-119 -- C_map(x) with the reference's column convention."""
        return self.C_map(x)

    def get_sim_params(self):
        # ssm.py:121-123 reads attributes the SSM class never sets; kept for surface compatibility
        return {'beta_weighting': getattr(self, 'beta_weighting', None), 'discr_method': self.discr_method,
                'dist_weights': getattr(self, 'dist_weights', None)}

    def get_state_dim(self):
        return self.state_dim

    def get_input_dim(self):
        return self.input_dim

    def get_output_dim(self):
        return self.output_dim

    def rollout(self, x0, u, dt):
        """ssm.py:134-156.  x0 (n,) & u (N, m) -> x (N+1, n), z (N+1, n_z); batched: x0 (Bt, n) & u (Bt, N, m)."""
        x0 = np.asarray(x0, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        single = (x0.ndim == 1)
        x0b = x0.reshape(-1, self.state_dim)
        xd, zd = self.rollout_device(L.to_dev(x0b), L.to_dev(u.reshape((x0b.shape[0],) + u.shape[-2:])), dt)
        x, z = L.to_host(xd), L.to_host(zd)
        return (x[0], z[0]) if single else (x, z)

    def rollout_pinned(self, x0_h, u_h, dt, out_h=None):
        """Extension (end-to-end batched rollout on PINNED host torch tensors): async H2D of x0 / u, one rollout, async
        D2H of x, z into pinned outputs -- allocated on first use, cached on the model and REUSED by the next call of
        the same shape --, one synchronisation."""
        torch = L.torch_mod()
        xd, zd = self.rollout_device(x0_h.cuda(non_blocking=True), u_h.cuda(non_blocking=True), dt)
        if out_h is None:
            cache = self.__dict__.setdefault('_pin_out', {})
            out_h = cache.get(tuple(xd.shape))
            if out_h is None:
                out_h = cache[tuple(xd.shape)] = {'x': torch.empty(xd.shape, dtype=xd.dtype, pin_memory=True),
                                                  'z': torch.empty(zd.shape, dtype=zd.dtype, pin_memory=True)}
        out_h['x'].copy_(xd, non_blocking=True)
        out_h['z'].copy_(zd, non_blocking=True)
        torch.cuda.synchronize()
        return out_h

    def rollout_device(self, x0, u, dt, want_z=True):
        """CUDA tensors x0 (Bt, n), u (Bt, N, m) -> CUDA tensors x (Bt, N+1, n), z (Bt, N+1, n_z)."""
        Bt, N = u.shape[0], u.shape[1]
        x = L.empty((Bt, N + 1, self.state_dim))
        z = L.empty((Bt, N + 1, self.output_dim)) if want_z else None
        L.check(L.lib().srcb200_ssm_rollout_batch(self.device_model(), Bt, N, L.ptr(x0), L.ptr(u), float(dt),
                                                  L.ptr(x), L.ptr(z), L.stream_ptr()))
        return x, z

    def get_poly_basis(self, dim, order):
        """ssm.py:158-164 returns a lambdified sympy basis; here a host callable with the same monomial order."""
        table = monomial_table(dim, order)

        def basis(*xs):
            xs = [np.asarray(v, dtype=np.float64) for v in xs]
            out = []
            for row in table:
                v = xs[row[0]]
                for j in row[1:]:
                    if j != 0xFF:
                        v = v * xs[j]
                out.append(v)
            return out
        return basis

    # Continuous maps (ssm.py:167-174) / discrete map (ssm.py:177-178)
    def reduced_dynamics(self, x, u):
        return self._dyn_map(x, u, 'cont_raw')

    def reduced_to_observed(self, x):
        return self._map_cols(0, x)

    def observed_to_reduced(self, z):
        return self._map_cols(1, z)

    def reduced_dynamics_discrete(self, x, u):
        return self._dyn_map(x, u, 'disc')

    def _eval_device(self, x, u, dt, kind, want):
        """want: subset of ('A','B','d','H','c','z') -> dict of CUDA tensors."""
        cnt, n, m, nz = x.shape[0], self.state_dim, self.input_dim, self.output_dim
        shapes = {'A': (cnt, n, n), 'B': (cnt, n, m), 'd': (cnt, n), 'H': (cnt, nz, n), 'c': (cnt, nz), 'z': (cnt, nz)}
        out = {k: L.empty(shapes[k]) for k in want}
        g = lambda k: L.ptr(out.get(k))
        L.check(L.lib().srcb200_ssm_eval_linearize_batch(self._handle(kind), cnt, L.ptr(x), L.ptr(u), float(dt),
                                                         g('A'), g('B'), g('d'), g('H'), g('c'), g('z'),
                                                         L.stream_ptr()))
        return out


class SSMDynamics(SSM):
    """ssm.py:181-344."""

    def __init__(self, eq_point, discrete=False, discr_method='fe', **kwargs):
        super(SSMDynamics, self).__init__(eq_point, discrete=discrete, discr_method=discr_method, **kwargs)

    def _lin(self, x, u, dt, kind, want):
        x = np.asarray(x, dtype=np.float64)
        single = (x.ndim == 1)
        xd = L.to_dev(x.reshape(-1, self.state_dim))
        ud = None if u is None else L.to_dev(np.asarray(u, dtype=np.float64).reshape(-1, self.input_dim))
        out = self._eval_device(xd, ud, dt, kind, want)
        res = tuple(L.to_host(out[k]) for k in want)
        return tuple(r[0] for r in res) if single else res

    def update_state(self, x, u, dt):
        """x+ for a step dt (ssm.py:187-195)."""
        A_d, B_d, d_d = self.get_jacobians(x, dt=dt, u=u)
        if np.asarray(x).ndim == 1:
            return self.update_dynamics(x, u, A_d, B_d, d_d)
        return np.einsum('bij,bj->bi', A_d, x) + np.einsum('bij,bj->bi', B_d, u) + d_d

    def get_continuous_jacobians(self, x, u):
        """A = df/dx, B = df/du, d = f - A x - B u of the continuous model (ssm.py:198-204)."""
        return self._lin(x, u, -1.0, 'cont_raw', ('A', 'B', 'd'))

    def get_discrete_jacobians(self, x, u):
        """Same for the identified discrete map rd_coeff/Bd (ssm.py:206-212)."""
        return self._lin(x, u, -1.0, 'disc', ('A', 'B', 'd'))

    def get_jacobians(self, x, u, dt):
        """ssm.py:215-225."""
        if not self.discrete:
            return self._lin(x, u, dt, 'cont', ('A', 'B', 'd'))
        return self._lin(x, u, -1.0, 'disc', ('A', 'B', 'd'))

    def get_observer_jacobians(self, x):
        """H = dC/dx, c_res = C(x) - H x (ssm.py:228-235)."""
        return self._lin(x, None, -1.0, 'disc' if self.discrete else 'cont_raw', ('H', 'c'))

    def update_observer_state(self, x, dt=None, u=None):
        """ssm.py:271-277."""
        H, c = self.get_observer_jacobians(x)
        if np.asarray(x).ndim == 1:
            return np.squeeze(np.dot(H, x)) + np.squeeze(c)
        return np.einsum('bij,bj->bi', H, x) + c

    def discretize_dynamics(self, A_c, B_c, d_c, dt):
        """ssm.py:279-301 (fe / be / bil; zoh raises like the reference)."""
        if self.discr_method not in ('fe', 'be', 'bil'):
            raise RuntimeError('self.discr_method must be in [fe, be, bil, zoh]')
        L.require_gpu()
        A_c = np.asarray(A_c, dtype=np.float64)
        single = (A_c.ndim == 2)
        n, m = A_c.shape[-1], np.asarray(B_c).shape[-1]
        A = L.to_dev(A_c.reshape(-1, n, n))
        B = L.to_dev(np.asarray(B_c, dtype=np.float64).reshape(-1, n, m))
        d = L.to_dev(np.asarray(d_c, dtype=np.float64).reshape(-1, n))
        L.check(L.lib().srcb200_discretize_batch(n, m, L.DISCR[self.discr_method], A.shape[0], float(dt), L.ptr(A),
                                                 L.ptr(B), L.ptr(d), L.ptr(A), L.ptr(B), L.ptr(d), L.stream_ptr()))
        res = (L.to_host(A), L.to_host(B), L.to_host(d))
        return tuple(r[0] for r in res) if single else res

    @staticmethod
    def update_dynamics(x, u, A_d, B_d, d_d):
        """ssm.py:330-333 -- the caller-side affine step on host arrays (three tiny products)."""
        return np.squeeze(A_d @ x) + np.squeeze(B_d @ u) + np.squeeze(d_d)

    def get_ref_point(self):
        return self.z_ref

    def compute_RO_state(self, z):
        """W_map(z - z_ref) (ssm.py:338-344); z is (n_z,) or (n_z, N) like the reference's W_map input."""
        z = np.asarray(z, dtype=np.float64)
        if z.ndim == 1:
            return L.to_host(self._map_device(1, True, L.to_dev(z[None, :])))[0]
        return L.to_host(self._map_device(1, True, L.to_dev(z.T))).T
