"""TPWL reduced-order model on B200 -- drop-in for sofacontrol/tpwl/tpwl.py (TPWL, TPWLATV).

Same constructor, attributes and method names/argument meaning as the reference (tpwl.py:14-342).  The bank of
stored linearisations lives in HBM (positions/velocities transposed so that consecutive points are contiguous);
every numerical method runs the CUDA kernels of libsrcb200 (csrc/tpwl.cu, csrc/gemm.cu) and returns host numpy
FP64 arrays with the reference's shapes.  Extension: a leading batch axis is accepted everywhere, and `*_device`
variants keep inputs/outputs as CUDA torch tensors.

The nearest-point index is bit-exact with the reference's np.argmin (numpy's pairwise summation order is
reproduced in the kernel).
"""
import numpy as np

from .. import _lib as L
from .. import utils as scutils
from ..mor import pod

###  DEFAULT VALUES (tpwl.py:7-11)
DISCR_METHOD = 'zoh'
TPWL_METHOD = 'nn'
DISCR_DICT = {'fe': 'forward Euler', 'be': 'implicit Euler', 'bil': 'bilinear transform', 'zoh': 'zero-order hold'}


class TPWL:
    """tpwl.py:14-216."""

    def __init__(self, data, params=None, Cf=None, Hf=None, **kwargs):
        if isinstance(data, dict):
            self.tpwl_dict = data
        else:
            self.tpwl_dict = scutils.load_data(data)
        self.num_points = len(self.tpwl_dict['q'])
        self.discr_method = kwargs.get('discr_method', 'fe')

        if self.tpwl_dict['rom_info']['type'] == 'POD':
            self.rom = pod.POD(self.tpwl_dict['rom_info'])
        else:
            raise NotImplementedError("Unknown ROM type")

        self.state_dim = int(np.asarray(self.tpwl_dict['q'][0]).shape[-1] * 2)
        self.input_dim = int(np.asarray(self.tpwl_dict['u'][0]).shape[-1])

        if params is None:
            params = dict()
        self.tpwl_method = params.get('tpwl_method', TPWL_METHOD)
        self.beta_weighting = params.get('beta_weighting', None)
        self.dist_weights = params.get('dist_weights')

        if Cf is not None:
            self.set_measurement_model(Cf)
        else:
            self.C = None
            self.y_ref = None
            self.meas_dim = None

        if Hf is not None:
            self.set_output_model(Hf)
        else:
            self.H = None
            self.z_ref = None
            self.output_dim = None

        self.nonlinear_observer = False
        self.pre_discretized_dt = None
        self.A_d = None
        self.B_d = None
        self.d_d = None
        self._dev = None      # device copies of the bank
        self._dev_d = None    # device copies of the pre-discretised bank

    # ---- device bank --------------------------------------------------------------------------------------------
    def invalidate_device_cache(self):
        """Drops the device copies of the bank, the output model and the discretised banks.  They are captured at
        first use (44 MB at the Diamond size: too large to fingerprint per call); call this after modifying
        tpwl_dict / H / z_ref in place."""
        self._dev = None
        self._zoh_cache = None

    def _bank(self):
        if self._dev is None:
            L.require_gpu()
            D = self.tpwl_dict
            f64 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float64)
            self._dev = dict(qT=L.to_dev(f64(D['q']).T.copy()), vT=L.to_dev(f64(D['v']).T.copy()),
                             A=L.to_dev(f64(D['A_c'])), B=L.to_dev(f64(D['B_c'])), d=L.to_dev(f64(D['d_c'])))
        return self._dev

    def _out(self):
        if self.H is None:
            return None, None
        if '_H' not in self._bank():
            self._dev['_H'] = L.to_dev(np.asarray(self.H, dtype=np.float64))
            self._dev['_z'] = L.to_dev(np.asarray(self.z_ref, dtype=np.float64))
        return self._dev['_H'], self._dev['_z']

    def device_model(self, dt=None, continuous=False):
        """srcb200_tpwl_model for evaluations with step dt.  Mirrors tpwl.py:254-263: the pre-discretised bank is
        used iff tpwl_method == 'nn' and dt == pre_discretized_dt; otherwise the continuous bank is discretised per
        evaluation with self.discr_method (or left continuous when dt is None)."""
        bank = self._bank()
        if self.tpwl_method not in L.TPWL_METHOD:
            raise RuntimeError('tpwl method should be nn or weighting')   # tpwl.py:268
        A, B, d = bank['A'], bank['B'], bank['d']
        method = 'none'
        if not continuous and dt is not None:
            if self.tpwl_method == 'nn' and self.pre_discretized_dt is not None and dt == self.pre_discretized_dt:
                A, B, d = self._dev_d['A'], self._dev_d['B'], self._dev_d['d']
            else:
                method = self.discr_method
                if method not in ('fe', 'be', 'bil', 'zoh'):
                    raise RuntimeError('self.discr_method must be in [fe, be, bil, zoh]')   # tpwl.py:295
        H, z = self._out()
        # the reference subscripts dist_weights / multiplies by beta_weighting unconditionally (tpwl.py:166-167, 176):
        # None raises TypeError there, so it does here instead of silently selecting point 0 / uniform weights
        if self.dist_weights is None:
            raise TypeError("'NoneType' object is not subscriptable (params['dist_weights'] is required, tpwl.py:166)")
        if self.tpwl_method == 'weighting' and self.beta_weighting is None:
            raise TypeError("bad operand type for unary -: 'NoneType' (params['beta_weighting'] is required, tpwl.py:176)")
        dw = self.dist_weights
        h = L.TpwlModel(n=self.state_dim, m=self.input_dim, nz=(0 if self.H is None else int(self.H.shape[0])),
                        P=self.num_points, method=L.TPWL_METHOD[self.tpwl_method], discr_method=L.DISCR[method],
                        wq=float(dw.get('q', 0.0)), wv=float(dw.get('v', 0.0)),
                        beta=float(self.beta_weighting if self.beta_weighting is not None else 0.0),
                        qT=L.ptr(bank['qT']), vT=L.ptr(bank['vT']), A=L.ptr(A), B=L.ptr(B), d=L.ptr(d),
                        H=L.ptr(H), z_ref=L.ptr(z))
        h._keep = (A, B, d, H, z, bank)
        return h

    def update_state(self, x, u, dt):
        raise NotImplementedError("update_state must be overriden by a child class")

    def get_jacobians(self, x, dt=None):
        raise NotImplementedError("get_jacobians must be overriden by a child class")

    def set_measurement_model(self, Cf):
        """tpwl.py:81-84 (one-off host projection of the measurement matrix onto the POD basis)."""
        self.C = Cf @ self.rom.V
        self.y_ref = Cf @ self.rom.x_ref
        self.meas_dim = self.C.shape[0]

    def set_output_model(self, Hf):
        """tpwl.py:86-89."""
        self.H = np.asarray(Hf @ self.rom.V)
        self.z_ref = np.asarray(Hf @ self.rom.x_ref)
        self.output_dim = self.H.shape[0]
        if getattr(self, '_dev', None) is not None:
            self._dev.pop('_H', None)
            self._dev.pop('_z', None)

    # ---- shifts / linear output maps (tpwl.py:91-137): affine bookkeeping on host arrays
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

    def x_to_zfyf(self, x, zf=False, yf=False):
        """tpwl.py:115-126; the performance output z = H x + z_ref runs on the device."""
        if zf and self.H is not None:
            x = np.asarray(x, dtype=np.float64)
            xd = L.to_dev(x.reshape(-1, self.state_dim))
            z = L.empty((xd.shape[0], self.H.shape[0]))
            L.check(L.lib().srcb200_tpwl_output_batch(self.device_model(continuous=True), xd.shape[0], L.ptr(xd),
                                                      L.ptr(z), L.stream_ptr()))
            return L.to_host(z).reshape(x.shape[:-1] + (self.H.shape[0],))
        elif yf and self.C is not None:
            return np.transpose(self.C @ x.T) + self.y_ref
        raise RuntimeError('Need to set output or meas. model')

    def x_to_zy(self, x, z=False, y=False):
        if z and self.H is not None:
            return np.transpose(self.H @ x.T)
        elif y and self.C is not None:
            return np.transpose(self.C @ x.T)
        raise RuntimeError('Need to set output or meas. model')

    def get_state_dim(self):
        return self.state_dim

    def get_input_dim(self):
        return self.input_dim

    def get_output_dim(self):
        return self.output_dim

    def get_meas_dim(self):
        return self.meas_dim

    def get_rom_info(self):
        return self.tpwl_dict['rom_info']

    def get_sim_params(self):
        return {'beta_weighting': self.beta_weighting, 'discr_method': self.discr_method,
                'tpwl_method': self.tpwl_method, 'dist_weights': self.dist_weights}

    # ---- point selection ----------------------------------------------------------------------------------------
    def _states(self, x):
        x = np.asarray(x, dtype=np.float64)
        return x.ndim == 1, L.to_dev(x.reshape(-1, self.state_dim))

    def nearest_device(self, x, want_dist=False):
        """CUDA tensor x (count, n) -> int32 CUDA tensor idx (count) [, dist (count)]."""
        torch = L.torch_mod()
        idx = L.empty((x.shape[0],), torch.int32)
        dist = L.empty((x.shape[0],)) if want_dist else None
        h = self.device_model(continuous=True)
        L.check(L.lib().srcb200_tpwl_nearest_batch(h, x.shape[0], L.ptr(x), L.ptr(idx), L.ptr(dist), L.stream_ptr()))
        return (idx, dist) if want_dist else idx

    def calc_nearest_point(self, x):
        """Index of the stored point minimising w_q |q_i - q| + w_v |v_i - v| (tpwl.py:160-168); bit-exact."""
        single, xd = self._states(x)
        idx = L.to_host(self.nearest_device(xd)).astype(np.int64)
        return idx[0] if single else idx

    def weights_device(self, x):
        w = L.empty((x.shape[0], self.num_points))
        h = self.device_model(continuous=True)
        L.check(L.lib().srcb200_tpwl_weights_batch(h, x.shape[0], L.ptr(x), L.ptr(w), L.stream_ptr()))
        return w

    def calc_weighting_factors(self, x):
        """Normalised exponential weights of every stored point (tpwl.py:170-191)."""
        single, xd = self._states(x)
        w = L.to_host(self.weights_device(xd))
        return w[0] if single else w

    def rollout_device(self, x0, u, dt, want_z=True, want_idx=False):
        """CUDA tensors x0 (Bt, n), u (Bt, N, m) -> x (Bt, N+1, n), z (Bt, N+1, n_z) or None [, idx (Bt, N)]."""
        torch = L.torch_mod()
        Bt, N = u.shape[0], u.shape[1]
        if (self.tpwl_method == 'nn' and self.discr_method == 'zoh' and dt is not None and
                not (self.pre_discretized_dt is not None and dt == self.pre_discretized_dt)):
            # the reference runs expm on the selected entry at every step (tpwl.py:261-263); discretising the whole
            # bank once for this dt gives the same matrices
            h = self._zoh_bank_model(dt)
        else:
            h = self.device_model(dt)
        x = L.empty((Bt, N + 1, self.state_dim))
        z = L.empty((Bt, N + 1, h.nz)) if (want_z and h.nz > 0) else None
        idx = L.empty((Bt, N), torch.int32) if (want_idx and self.tpwl_method == 'nn') else None
        wsb = L.lib().srcb200_tpwl_rollout_workspace(h, Bt)
        ws = L.empty((max(int(wsb), 8) // 8,))
        L.check(L.lib().srcb200_tpwl_rollout_batch(h, Bt, N, L.ptr(x0), L.ptr(u), float(dt), L.ptr(x), L.ptr(z),
                                                   L.ptr(idx), L.ptr(ws), ws.numel() * 8, L.stream_ptr()))
        return (x, z, idx) if want_idx else (x, z)

    def _discretize_bank_device(self, dt):
        """(A_d, B_d, d_d) CUDA tensors of the whole bank for step dt with self.discr_method."""
        bank = self._bank()
        n, m = self.state_dim, self.input_dim
        A, B, d = L.empty(bank['A'].shape), L.empty(bank['B'].shape), L.empty(bank['d'].shape)
        if self.discr_method == 'zoh':
            wsb = int(L.lib().srcb200_zoh_workspace(n, m, self.num_points))
            ws = L.empty((wsb // 8 + 1,))
            L.check(L.lib().srcb200_zoh_batch(n, m, self.num_points, float(dt), L.ptr(bank['A']), L.ptr(bank['B']),
                                              L.ptr(bank['d']), L.ptr(A), L.ptr(B), L.ptr(d), L.ptr(ws),
                                              ws.numel() * 8, L.stream_ptr()))
        else:
            L.check(L.lib().srcb200_discretize_batch(n, m, L.DISCR[self.discr_method], self.num_points, float(dt),
                                                     L.ptr(bank['A']), L.ptr(bank['B']), L.ptr(bank['d']), L.ptr(A),
                                                     L.ptr(B), L.ptr(d), L.stream_ptr()))
        return A, B, d

    def _zoh_bank_model(self, dt):
        cache = getattr(self, '_zoh_cache', None)
        key = (dt, self.discr_method, id(self._bank()))        # a new device bank (invalidate_device_cache) re-discretises
        if cache is None or cache[0] != key:
            cache = (key, self._discretize_bank_device(dt))
            self._zoh_cache = cache
        h = self.device_model(continuous=True)
        A, B, d = cache[1]
        h.A, h.B, h.d = L.ptr(A), L.ptr(B), L.ptr(d)
        h._keep = h._keep + (A, B, d)
        return h

    def rollout(self, x0, u, dt):
        """tpwl.py:193-216.  x0 (n,) & u (N, m) -> x (N+1, n), z (N+1, n_z) or None; batched with a leading axis."""
        x0 = np.asarray(x0, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        single = (x0.ndim == 1)
        x0b = x0.reshape(-1, self.state_dim)
        xd, zd = self.rollout_device(L.to_dev(x0b), L.to_dev(u.reshape((x0b.shape[0],) + u.shape[-2:])), dt)
        x = L.to_host(xd)
        z = None if zd is None else L.to_host(zd)
        if single:
            return x[0], (None if z is None else z[0])
        return x, z


    def rollout_pinned(self, x0_h, u_h, dt, out_h=None):
        """Extension (end-to-end batched rollout on PINNED host torch tensors): async H2D of x0 (Bt, n) / u (Bt, N, m),
        one rollout, async D2H of x (and z) into pinned outputs -- allocated on first use, cached on the model and REUSED
        by the next call of the same shape (copy them out if they must survive it) --, one synchronisation."""
        torch = L.torch_mod()
        xd, zd = self.rollout_device(x0_h.cuda(non_blocking=True), u_h.cuda(non_blocking=True), dt)
        if out_h is None:
            key = (tuple(xd.shape), None if zd is None else tuple(zd.shape))
            cache = self.__dict__.setdefault('_pin_out', {})
            out_h = cache.get(key)
            if out_h is None:
                out_h = cache[key] = {'x': torch.empty(xd.shape, dtype=xd.dtype, pin_memory=True),
                                      'z': None if zd is None else torch.empty(zd.shape, dtype=zd.dtype, pin_memory=True)}
        out_h['x'].copy_(xd, non_blocking=True)
        if zd is not None and out_h.get('z') is not None:
            out_h['z'].copy_(zd, non_blocking=True)
        torch.cuda.synchronize()
        return out_h


class TPWLATV(TPWL):
    """tpwl.py:219-342."""

    def __init__(self, data, params=None, Cf=None, Hf=None, **kwargs):
        super(TPWLATV, self).__init__(data, params, Cf=Cf, Hf=Hf, **kwargs)
        self.ref_point = None

    def update_state(self, x, u, dt):
        """x+ for a step dt (tpwl.py:226-234)."""
        A_d, B_d, d_d = self.get_jacobians(x, dt)
        if np.asarray(x).ndim == 1:
            return self.update_dynamics(x, u, A_d, B_d, d_d)
        return np.einsum('bij,bj->bi', A_d, x) + np.einsum('bij,bj->bi', B_d, u) + d_d

    def linearize_device(self, x, dt=None):
        """CUDA tensor x (count, n) -> CUDA tensors A (count, n, n), B (count, n, m), d (count, n), idx or None."""
        torch = L.torch_mod()
        cnt, n, m = x.shape[0], self.state_dim, self.input_dim
        h = self.device_model(dt)
        A, B, d = L.empty((cnt, n, n)), L.empty((cnt, n, m)), L.empty((cnt, n))
        idx = L.empty((cnt,), torch.int32) if self.tpwl_method == 'nn' else None
        wsb = L.lib().srcb200_tpwl_linearize_workspace(h, cnt)
        ws = L.empty((max(int(wsb), 8) // 8,))
        L.check(L.lib().srcb200_tpwl_linearize_batch(h, cnt, L.ptr(x), -1.0 if dt is None else float(dt), L.ptr(A),
                                                     L.ptr(B), L.ptr(d), L.ptr(idx), L.ptr(ws), ws.numel() * 8,
                                                     L.stream_ptr()))
        return A, B, d, idx

    def get_jacobians(self, x, dt=None, u=None):
        """(A, B, d) at the state x (tpwl.py:236-270): continuous if dt is None, else discretised (or taken from
        the pre-discretised bank when dt == pre_discretized_dt).  Sets self.ref_point in nn mode.  Unlike the
        reference the nn result is a fresh array, not a view into the bank."""
        single, xd = self._states(x)
        A, B, d, idx = self.linearize_device(xd, dt)
        if idx is not None:
            ih = L.to_host(idx).astype(np.int64)
            self.ref_point = ih[0] if single else ih
        A, B, d = L.to_host(A), L.to_host(B), L.to_host(d)
        return (A[0], B[0], d[0]) if single else (A, B, d)

    def discretize_dynamics(self, A_c, B_c, d_c, dt):
        """tpwl.py:272-297 for one or a stack of (A_c, B_c, d_c); fe/be/bil run csrc/tpwl.cu discretize_kernel."""
        if self.discr_method not in ('fe', 'be', 'bil', 'zoh'):
            raise RuntimeError('self.discr_method must be in [fe, be, bil, zoh]')
        L.require_gpu()
        A_c = np.asarray(A_c, dtype=np.float64)
        single = (A_c.ndim == 2)
        n, m = A_c.shape[-1], np.asarray(B_c).shape[-1]
        A = L.to_dev(A_c.reshape(-1, n, n))
        B = L.to_dev(np.asarray(B_c, dtype=np.float64).reshape(-1, n, m))
        d = L.to_dev(np.asarray(d_c, dtype=np.float64).reshape(-1, n))
        if self.discr_method == 'zoh':
            wsb = int(L.lib().srcb200_zoh_workspace(n, m, A.shape[0]))
            ws = L.empty((wsb // 8 + 1,))
            L.check(L.lib().srcb200_zoh_batch(n, m, A.shape[0], float(dt), L.ptr(A), L.ptr(B), L.ptr(d), L.ptr(A),
                                              L.ptr(B), L.ptr(d), L.ptr(ws), ws.numel() * 8, L.stream_ptr()))
        else:
            L.check(L.lib().srcb200_discretize_batch(n, m, L.DISCR[self.discr_method], A.shape[0], float(dt), L.ptr(A),
                                                     L.ptr(B), L.ptr(d), L.ptr(A), L.ptr(B), L.ptr(d), L.stream_ptr()))
        res = (L.to_host(A), L.to_host(B), L.to_host(d))
        return tuple(r[0] for r in res) if single else res

    def pre_discretize(self, dt):
        """Discretises the whole bank once on the device and keeps it in HBM (tpwl.py:299-322)."""
        if self.tpwl_method != 'nn':
            raise RuntimeError('tpwl method should be nn to pre-discretize')
        if self.discr_method not in ('fe', 'be', 'bil', 'zoh'):
            raise RuntimeError('self.discr_method must be in [fe, be, bil, zoh]')
        A, B, d = self._discretize_bank_device(dt)
        self._dev_d = dict(A=A, B=B, d=d)
        self.A_d, self.B_d, self.d_d = L.to_host(A), L.to_host(B), L.to_host(d)
        self.pre_discretized_dt = dt

    def set_pre_discretized(self, A_d, B_d, d_d, dt):
        """Extension: install an externally discretised bank (e.g. the reference's zoh output) as the
        pre-discretised model for step dt."""
        f64 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float64)
        self.A_d, self.B_d, self.d_d = f64(A_d), f64(B_d), f64(d_d)
        L.require_gpu()
        self._dev_d = dict(A=L.to_dev(self.A_d), B=L.to_dev(self.B_d), d=L.to_dev(self.d_d))
        self.pre_discretized_dt = dt

    def get_characteristic_dx(self, dt):
        """max_i |x+_i - x_i| over the stored points (tpwl.py:324-334), evaluated as one batch."""
        x = scutils.qv2x(np.asarray(self.tpwl_dict['q']), np.asarray(self.tpwl_dict['v']))
        dx = self.update_state(x, np.asarray(self.tpwl_dict['u'], dtype=np.float64), dt) - x
        return np.abs(dx).max(axis=0)

    @staticmethod
    def update_dynamics(x, u, A_d, B_d, d_d):
        """tpwl.py:336-339 -- the caller-side affine step on host arrays."""
        return A_d @ x + np.squeeze(B_d @ u) + d_d

    def get_ref_point(self):
        return self.ref_point
