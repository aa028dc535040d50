"""Observers of the closed loop on B200 -- drop-in for sofacontrol/tpwl/observer.py (FullStateObserver,
DiscreteEKFObserver), with a leading batch axis for Monte-Carlo closed loops (BASELINE config 4).

`DiscreteEKFObserver(dyn_sys, Sigma0=, W=, V=)`, `initialize(xf)`, `update(u, y, dt)`, `predict_state(u, dt)`,
`update_state(y)`, attributes `x`, `z`, `Sigma`, `C`, `W`, `V` keep the reference's names, argument meaning and shapes
(observer.py:33-126).  A single filter works on 1-D host arrays exactly like the reference; `x (Bt, n)` /
`y (Bt, p)` / `u (Bt, m)` run Bt independent filters in one launch each of
    srcb200_tpwl_linearize_batch (A_d, B_d, d_d at the belief states -- nearest-point search per filter),
    srcb200_ekf_predict_batch    (x <- A x + B u + d, Sigma <- A Sigma A^T + W, csrc/control.cu),
    srcb200_ekf_update_batch     (gain through the p x p innovation covariance, Joseph-free update like the reference).
The device state (`x_dev`, `Sigma_dev`) stays in HBM between calls; `x`, `Sigma`, `z` are host views fetched on access.
"""
import numpy as np

from .. import _lib as L


class FullStateObserver:
    """observer.py:3-30: full-state perfect measurement; the output map z = H x is a host matrix-vector product on
    the measured state (pure bookkeeping: there is nothing to estimate)."""

    def __init__(self, n_x, H=None):
        self.x = None
        self.z = None
        self.meas_dim = n_x
        self.state_dim = n_x
        self.H = H

    def get_meas_dim(self):
        return self.meas_dim

    def get_observer_params(self):
        return {'meas_dim': self.meas_dim, 'state_dim': self.state_dim}

    def update(self, u, y, dt, x=None):
        self.x = x
        if self.H is not None:
            self.z = np.asarray(x) @ np.asarray(self.H).T if np.asarray(x).ndim == 2 else self.H @ x
        else:
            self.z = x


class DiscreteEKFObserver:
    """observer.py:33-126 on a sofacontrol_b200 TPWLATV model."""

    def __init__(self, dyn_sys, **kwargs):
        self.dyn_sys = dyn_sys
        if self.dyn_sys.C is None:
            raise RuntimeError('Need to set meas. model in dyn_sys')
        self.C = self.dyn_sys.C
        self.state_dim = self.dyn_sys.get_state_dim()
        self.meas_dim = self.C.shape[0]
        self._Sigma0 = np.asarray(kwargs.get('Sigma0', np.eye(self.state_dim)), dtype=np.float64)
        self.W = kwargs.get('W', 100 * np.eye(self.state_dim))
        self.V = kwargs.get('V', np.eye(self.meas_dim))
        self._single = True
        self._x_dev = self._S_dev = None
        self._const = None
        self.initialize(self.dyn_sys.rom.x_ref)

    # ---- device state ----------------------------------------------------------------------------------------
    def _constants(self):
        if self._const is None:
            f64 = lambda a: L.to_dev(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))
            self._const = dict(C=f64(self.C), W=f64(self.W), V=f64(self.V), yref=f64(self.dyn_sys.y_ref))
        return self._const

    def _set_state(self, x):
        """x (n,) or (Bt, n) host -> device belief + covariance Sigma0 per filter."""
        x = np.asarray(x, dtype=np.float64)
        self._single = (x.ndim == 1)
        xb = x.reshape(-1, self.state_dim)
        self._x_dev = L.to_dev(xb)
        self._S_dev = L.to_dev(np.broadcast_to(self._Sigma0, (xb.shape[0],) + self._Sigma0.shape[-2:]))

    @property
    def x_dev(self):
        return self._x_dev

    @property
    def Sigma_dev(self):
        return self._S_dev

    @property
    def x(self):
        h = L.to_host(self._x_dev)
        return h[0] if self._single else h

    @x.setter
    def x(self, value):
        v = np.asarray(value, dtype=np.float64)
        self._single = (v.ndim == 1)
        self._x_dev = L.to_dev(v.reshape(-1, self.state_dim))

    @property
    def Sigma(self):
        h = L.to_host(self._S_dev)
        return h[0] if self._single else h

    @Sigma.setter
    def Sigma(self, value):
        v = np.asarray(value, dtype=np.float64)
        self._S_dev = L.to_dev(v.reshape((-1,) + v.shape[-2:]))

    @property
    def z(self):
        """observer.py:78-81 / 121-124: the performance (or measurement) output of the belief state."""
        if self.dyn_sys.H is not None:
            return self.dyn_sys.x_to_zfyf(self.x, zf=True)
        return self.dyn_sys.x_to_zfyf(self.x, yf=True)

    # ---- reference API ---------------------------------------------------------------------------------------
    def get_meas_dim(self):
        return self.meas_dim

    def get_observer_params(self):
        return {'W': self.W, 'V': self.V, 'meas_dim': self.meas_dim, 'state_dim': self.state_dim,
                'C': self.C, 'H': self.dyn_sys.H}

    def initialize(self, xf):
        """observer.py:71-81: belief = reduced-order projection of the full-order state(s) xf ((nf,) or (Bt, nf))."""
        xf = np.asarray(xf, dtype=np.float64)
        if xf.ndim == 1:
            self._set_state(self.dyn_sys.rom.compute_RO_state(xf=xf))
        else:
            self._set_state(np.stack([self.dyn_sys.rom.compute_RO_state(xf=v) for v in xf]))

    def initialize_reduced(self, x):
        """Extension: start from reduced-order belief state(s) x (n,) or (Bt, n)."""
        self._set_state(x)

    def update(self, u, y, dt, **kwargs):
        """observer.py:83-92: full EKF step with the input of step k and the measurement of step k+1."""
        self.predict_state(u, dt)
        self.update_state(y)

    def predict_state(self, u, dt):
        """observer.py:94-104."""
        ud = L.to_dev(np.asarray(u, dtype=np.float64).reshape(-1, self.dyn_sys.get_input_dim()))
        self.predict_device(ud, dt)

    def predict_device(self, u, dt):
        """CUDA tensor u (Bt, m)."""
        n, m, Bt = self.state_dim, self.dyn_sys.get_input_dim(), self._x_dev.shape[0]
        A, B, d, _ = self.dyn_sys.linearize_device(self._x_dev, dt)
        c = self._constants()
        L.check(L.lib().srcb200_ekf_predict_batch(n, m, Bt, L.ptr(A), L.ptr(B), L.ptr(d), L.ptr(u), L.ptr(c['W']),
                                                  L.ptr(self._x_dev), L.ptr(self._S_dev), L.stream_ptr()))

    def update_state(self, y):
        """observer.py:106-126; y is the FULL-order measurement (y_ref is subtracted inside the kernel)."""
        yd = L.to_dev(np.asarray(y, dtype=np.float64).reshape(-1, self.meas_dim))
        self.update_device(yd)
        return self.x

    def update_device(self, y):
        """CUDA tensor y (Bt, p)."""
        c = self._constants()
        L.check(L.lib().srcb200_ekf_update_batch(self.state_dim, self.meas_dim, self._x_dev.shape[0], L.ptr(c['C']),
                                                 L.ptr(c['V']), L.ptr(c['yref']), L.ptr(y), L.ptr(self._x_dev),
                                                 L.ptr(self._S_dev), L.stream_ptr()))
