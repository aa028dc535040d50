"""Time-varying LQR tracking on B200 -- drop-in for sofacontrol/lqr/traj_tracking_lqr.py (TrajTrackingLQR).

`perform_dlqr_recursion(target)` linearises the model at every sample of the nominal trajectory in ONE batched launch
(the reference calls model.get_jacobians per step, traj_tracking_lqr.py:36) and runs the backward Riccati recursion
(traj_tracking_lqr.py:38-41) in csrc/control.cu: tvlqr_kernel.  `target.x` may carry a leading batch axis
(Bt, T, n) with `target.u` (Bt, T, m): one CTA per trajectory.
"""
import numpy as np

from .. import _lib as L


def _interp_rows(t_src, vals, t_new):
    """scipy.interpolate.interp1d(t_src, vals, axis=-2)(t_new), linear (the reference's default kind)."""
    t_src = np.asarray(t_src, dtype=np.float64)
    vals = np.asarray(vals, dtype=np.float64)
    hi = np.clip(np.searchsorted(t_src, t_new, side='left'), 1, len(t_src) - 1)
    lo = hi - 1
    slope = (vals[..., hi, :] - vals[..., lo, :]) / (t_src[hi] - t_src[lo])[:, None]
    return slope * (t_new - t_src[lo])[:, None] + vals[..., lo, :]


class TrajTrackingLQR:
    def __init__(self, dt, model, cost_params):
        self.dt = dt
        self.model = model
        self.cost_params = cost_params
        self.x_bar = None
        self.u_bar = None

    def compute_policy(self, target):
        K, _ = self.perform_dlqr_recursion(target)
        return self.x_bar, self.u_bar, K

    def perform_dlqr_recursion(self, target):
        """traj_tracking_lqr.py:18-48 -> (K (steps, m, n), P (steps + 1, n, n)) in time order."""
        L.require_gpu()
        final_time = target.t[-1]
        nbr_steps = int(final_time / self.dt)
        t_steps = np.arange(nbr_steps) * self.dt
        x_nom = _interp_rows(target.t, target.x, t_steps)
        u_nom = _interp_rows(target.t, target.u, t_steps)
        single = (x_nom.ndim == 2)
        n, m = x_nom.shape[-1], u_nom.shape[-1]
        xb = x_nom.reshape(-1, nbr_steps, n)
        bt = xb.shape[0]
        A, B, _, _ = self.model.linearize_device(L.to_dev(xb.reshape(-1, n)), self.dt) \
            if hasattr(self.model, 'linearize_device') else self._linearize_ssm(xb, u_nom.reshape(-1, nbr_steps, m))
        Q = L.to_dev(np.asarray(self.cost_params.Q, dtype=np.float64))
        R = L.to_dev(np.asarray(self.cost_params.R, dtype=np.float64))
        K, P = L.empty((bt, nbr_steps, m, n)), L.empty((bt, nbr_steps + 1, n, n))
        L.check(L.lib().srcb200_tvlqr_batch(n, m, nbr_steps, bt, L.ptr(A), L.ptr(B), L.ptr(Q), L.ptr(R), L.ptr(K),
                                            L.ptr(P), L.stream_ptr()))
        self.x_bar, self.u_bar = x_nom, u_nom
        Kh, Ph = L.to_host(K), L.to_host(P)
        return (Kh[0], Ph[0]) if single else (Kh, Ph)

    def _linearize_ssm(self, xb, ub):
        """SSM models take the input as well (ssm.py:215-225)."""
        bt, T, n = xb.shape
        out = self.model._eval_device(L.to_dev(xb.reshape(-1, n)), L.to_dev(ub.reshape(bt * T, -1)), float(self.dt),
                                      'disc' if self.model.discrete else 'cont', ('A', 'B'))
        return out['A'], out['B'], None, None
