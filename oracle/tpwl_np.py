"""numpy restatement of sofacontrol/tpwl/tpwl.py (TPWL / TPWLATV).  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).
PINNED bit-for-bit against the imported reference in tests/test_oracle_vs_reference.py."""
import numpy as np
from . import utils_np as U


def pairwise_sumsq_row(d):
    """Bit-level model of `np.add.reduce(d*d)` for one contiguous row (what np.linalg.norm(axis=1) does before the
    sqrt; tpwl.py:166-167): squares rounded first, then numpy's pairwise summation --
    8 accumulators with stride 8, combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), sequential tail; rows longer
    than 128 are split recursively at n/2 rounded down to a multiple of 8.  This is the order the CUDA distance
    kernel reproduces; tests check it equals numpy's bits."""
    s = d * d

    def rec(a):
        n = a.shape[0]
        if n < 8:
            r = np.float64(0.0) if n == 0 else a[0]
            # numpy starts from a[0] (no leading 0.0 add) for the short path
            for i in range(1, n):
                r = r + a[i]
            return r
        if n <= 128:
            r = [a[i] for i in range(8)]
            i = 8
            while i < n - (n % 8):
                for j in range(8):
                    r[j] = r[j] + a[i + j]
                i += 8
            res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
            while i < n:
                res = res + a[i]
                i += 1
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return rec(a[:n2]) + rec(a[n2:])

    return rec(s)


class TPWLATVNP:
    """tpwl.py:14-342 restated.  `data` is the TPWL dict (keys q, v, u, A_c, B_c, d_c, rom_info)."""

    def __init__(self, data, params=None, Cf=None, Hf=None, discr_method='fe'):
        # tpwl.py:20-73
        from .pod_np import PODNP
        self.tpwl_dict = data
        self.num_points = len(data['q'])
        self.discr_method = discr_method
        if data['rom_info']['type'] != 'POD':
            raise NotImplementedError("Unknown ROM type")
        self.rom = PODNP(data['rom_info'])
        self.state_dim = np.asarray(data['q']).shape[-1] * 2
        self.input_dim = np.asarray(data['u']).shape[-1]
        params = params or {}
        self.tpwl_method = params.get('tpwl_method', 'nn')
        self.beta_weighting = params.get('beta_weighting', None)
        self.dist_weights = params.get('dist_weights')
        self.C = self.y_ref = self.meas_dim = None
        self.H = self.z_ref = self.output_dim = None
        if Cf is not None:
            self.C = Cf @ self.rom.V
            self.y_ref = Cf @ self.rom.x_ref
            self.meas_dim = self.C.shape[0]
        if Hf is not None:
            self.H = Hf @ self.rom.V
            self.z_ref = Hf @ self.rom.x_ref
            self.output_dim = self.H.shape[0]
        self.nonlinear_observer = False
        self.pre_discretized_dt = None
        self.A_d = self.B_d = self.d_d = None
        self.ref_point = None

    def get_state_dim(self):
        return self.state_dim

    def get_input_dim(self):
        return self.input_dim

    # tpwl.py:91-113
    def zfyf_to_zy(self, zf=None, yf=None):
        if zf is not None and self.z_ref is not None:
            return zf - self.z_ref
        elif yf is not None and self.y_ref is not None:
            return yf - self.y_ref
        raise RuntimeError('Need to set output or meas. model')

    def zy_to_zfyf(self, z=None, y=None):
        if z is not None and self.z_ref is not None:
            return z + self.z_ref
        elif y is not None and self.y_ref is not None:
            return y + self.y_ref
        raise RuntimeError('Need to set output or meas. model')

    # tpwl.py:115-126
    def x_to_zfyf(self, x, zf=False, yf=False):
        if zf and self.H is not None:
            return np.transpose(self.H @ x.T) + self.z_ref
        if yf and self.C is not None:
            return np.transpose(self.C @ x.T) + self.y_ref
        raise RuntimeError('Need to set output or meas. model')

    def _distances(self, x):
        # tpwl.py:165-167 / 174-177
        q, v = U.x2qv(x)
        qd = self.dist_weights['q'] * np.linalg.norm(self.tpwl_dict['q'] - q, axis=1)
        vd = self.dist_weights['v'] * np.linalg.norm(self.tpwl_dict['v'] - v, axis=1)
        return qd + vd

    def calc_nearest_point(self, x):
        # tpwl.py:160-168
        return np.argmin(self._distances(x))

    def calc_weighting_factors(self, x):
        # tpwl.py:170-191
        dist = self._distances(x)
        i = np.argmin(dist)
        m = dist[i]
        if m == 0:
            w = np.zeros(np.shape(dist))
            w[i] = 1
            return w
        w = np.exp(-self.beta_weighting * dist / m)
        return w / np.sum(w)

    def get_jacobians(self, x, dt=None, u=None):
        # tpwl.py:236-270
        D = self.tpwl_dict
        if self.tpwl_method == 'weighting':
            w = self.calc_weighting_factors(x)
            A = np.einsum("i, ijk -> jk", w, D['A_c'])
            B = np.einsum("i, ijk -> jk", w, D['B_c'])
            d = np.einsum("i, ij -> j", w, D['d_c'])
            if dt is not None:
                A, B, d = self.discretize_dynamics(A, B, d, dt)
        elif self.tpwl_method == 'nn':
            self.ref_point = self.calc_nearest_point(x)
            i = self.ref_point
            if self.pre_discretized_dt is not None and dt == self.pre_discretized_dt:
                A, B, d = self.A_d[i], self.B_d[i], self.d_d[i]
            else:
                A, B, d = D['A_c'][i], D['B_c'][i], D['d_c'][i]
                if dt is not None:
                    A, B, d = self.discretize_dynamics(A, B, d, dt)
        else:
            raise RuntimeError('tpwl method should be nn or weighting')
        return A, B, d

    def discretize_dynamics(self, A_c, B_c, d_c, dt):
        # tpwl.py:272-297
        I = np.eye(A_c.shape[0])
        if self.discr_method == 'fe':
            return I + dt * A_c, dt * B_c, dt * d_c
        if self.discr_method == 'zoh':
            return U.zoh_affine(A_c, B_c, d_c, dt)
        if self.discr_method == 'be':
            A_d = np.linalg.inv(I - dt * A_c)
        elif self.discr_method == 'bil':
            A_d = (I + 0.5 * dt * A_c) @ np.linalg.inv(I - 0.5 * dt * A_c)
        else:
            raise RuntimeError('self.discr_method must be in [fe, be, bil, zoh]')
        sep = np.linalg.inv(A_c) @ (A_d - I)
        return A_d, sep @ B_c, sep @ d_c

    def pre_discretize(self, dt):
        # tpwl.py:299-322
        if self.tpwl_method != 'nn':
            raise RuntimeError('tpwl method should be nn to pre-discretize')
        D = self.tpwl_dict
        out = [self.discretize_dynamics(D['A_c'][i], D['B_c'][i], D['d_c'][i], dt)
               for i in range(self.num_points)]
        self.A_d = [o[0] for o in out]
        self.B_d = [o[1] for o in out]
        self.d_d = [o[2] for o in out]
        self.pre_discretized_dt = dt

    def update_state(self, x, u, dt):
        # tpwl.py:226-234
        A_d, B_d, d_d = self.get_jacobians(x, dt)
        return self.update_dynamics(x, u, A_d, B_d, d_d)

    @staticmethod
    def update_dynamics(x, u, A_d, B_d, d_d):
        # tpwl.py:336-339
        return A_d @ x + np.squeeze(B_d @ u) + d_d

    def get_characteristic_dx(self, dt):
        # tpwl.py:324-334
        x = U.qv2x(self.tpwl_dict['q'], self.tpwl_dict['v'])
        dx = np.zeros(x.shape)
        for i in range(x.shape[0]):
            dx[i, :] = self.update_state(x[i, :], self.tpwl_dict['u'][i, :], dt) - x[i, :]
        return np.abs(dx).max(axis=0)

    def rollout(self, x0, u, dt):
        # tpwl.py:193-216
        N = u.shape[0]
        x = np.zeros((N + 1, self.state_dim))
        x[0, :] = x0
        for i in range(N):
            x[i + 1, :] = self.update_state(x[i, :], u[i, :], dt)
        z = self.x_to_zfyf(x, zf=True) if self.H is not None else None
        return x, z
