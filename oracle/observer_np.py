"""numpy restatement of sofacontrol/tpwl/observer.py (FullStateObserver, DiscreteEKFObserver) and of
SSM/controllers.py:302-309 (SSMObserver).  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).
PINNED bit-for-bit against the unmodified reference module (importable: it only needs numpy) in
tests/test_oracle_vs_reference.py::test_ekf_restatement_bitwise; golden vectors: tests/golden/ekf_small.npz.

Every product keeps the reference's association order (Python's left-to-right `@` chain).
"""
import numpy as np

from .utils_np import vq2qv


class FullStateObserverNP:
    """observer.py:3-30."""

    def __init__(self, n_x, H=None):
        self.x = None
        self.z = None
        self.meas_dim = n_x
        self.state_dim = n_x
        self.H = H

    def update(self, u, y, dt, x=None):
        self.x = x
        self.z = self.H @ x if self.H is not None else x


class DiscreteEKFObserverNP:
    """observer.py:33-126 on a TPWL model object (oracle or reference class: duck-typed)."""

    def __init__(self, dyn_sys, **kwargs):
        self.dyn_sys = dyn_sys
        if self.dyn_sys.C is None:
            raise RuntimeError('Need to set meas. model in dyn_sys')
        self.C = self.dyn_sys.C
        self.state_dim = self.dyn_sys.get_state_dim()
        self.meas_dim = self.C.shape[0]
        self.Sigma = kwargs.get('Sigma0', np.eye(self.state_dim))
        self.W = kwargs.get('W', 100 * np.eye(self.state_dim))
        self.V = kwargs.get('V', np.eye(self.meas_dim))
        self.initialize(self.dyn_sys.rom.x_ref)

    def _z(self):
        if self.dyn_sys.H is not None:
            return self.dyn_sys.x_to_zfyf(self.x, zf=True)
        return self.dyn_sys.x_to_zfyf(self.x, yf=True)

    def initialize(self, xf):
        # observer.py:71-81
        self.x = self.dyn_sys.rom.compute_RO_state(xf=xf)
        self.z = self._z()

    def update(self, u, y, dt, **kwargs):
        # observer.py:83-92
        self.predict_state(u, dt)
        self.update_state(y)

    def predict_state(self, u, dt):
        # observer.py:94-104
        A_d, B_d, d_d = self.dyn_sys.get_jacobians(self.x, dt)
        self.x = self.dyn_sys.update_dynamics(self.x, u, A_d, B_d, d_d)
        self.Sigma = A_d @ self.Sigma @ A_d.T + self.W

    def update_state(self, y):
        # observer.py:106-126
        y = self.dyn_sys.zfyf_to_zy(yf=y)
        S = self.C @ self.Sigma @ self.C.T + self.V
        K = self.Sigma @ self.C.T @ np.linalg.inv(S)
        self.x = self.x + K @ (y - self.C @ self.x)
        self.Sigma = (np.eye(self.state_dim) - K @ self.C) @ self.Sigma
        self.z = self._z()
        return self.x


class SSMObserverNP:
    """SSM/controllers.py:302-309: the measurement [v; q] reordered to the SSM output layout [q; v]."""

    def __init__(self, dyn_sys):
        self.z = None
        self.x = None
        self.dyn_sys = dyn_sys

    def update(self, u, y, dt, x=None):
        self.z = vq2qv(y)
