"""The model-side work of one GuSTO / SCP iteration on B200 -- the three methods of sofacontrol/scp/gusto.py that call
the reduced-order model once per trajectory point (gusto.py:203-281), as ONE batched launch each over all
(trajectories x points): `get_traj_dynamics`, `get_observer_linearizations`, `compute_accuracy`.  The convex
sub-problem (scp/locp.py, cvxpy + OSQP/Gurobi) is a different algorithm family and stays with the caller.

`TrajectoryLinearizer(model, dt)` holds what GuSTO holds for these calls (`model`, `dt`, `x_k`, `u_k`, `f_scale`);
x is (N + 1, n_x) and u (N, n_u) like in GuSTO, or (Bt, N + 1, n_x) / (Bt, N, n_u) for Bt problems at once.
"""
import numpy as np

from .. import _lib as L


class TrajectoryLinearizer:
    def __init__(self, model, dt, f_scale=None):
        self.model = model
        self.dt = dt
        self.x_k = None
        self.u_k = None
        self.f_scale = np.ones(model.n_x) if f_scale is None else np.asarray(f_scale, dtype=np.float64)

    def set_iterate(self, x_k, u_k):
        self.x_k = np.asarray(x_k, dtype=np.float64)
        self.u_k = np.asarray(u_k, dtype=np.float64)

    def get_traj_dynamics(self, x, u):
        """gusto.py:225-238: (A_d, B_d, d_d) at the first N points of the trajectory.  Returns arrays with a leading
        (N,) or (Bt, N) axis (the reference returns lists of N arrays)."""
        x = np.asarray(x, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        n, m = self.model.n_x, self.model.n_u
        pts = x[..., :-1, :]
        A, B, d = self.model.get_discrete_dynamics(pts.reshape(-1, n), u.reshape(-1, m), self.dt)
        lead = pts.shape[:-1]
        return A.reshape(lead + (n, n)), B.reshape(lead + (n, m)), d.reshape(lead + (n,))

    def get_observer_linearizations(self, x, u=None):
        """gusto.py:240-251: (H_d, c_d) at all N + 1 points (SSM models)."""
        x = np.asarray(x, dtype=np.float64)
        n = self.model.n_x
        H, c = self.model.get_observer_jacobians(x.reshape(-1, n), None, self.dt)
        return H.reshape(x.shape[:-1] + H.shape[-2:]), c.reshape(x.shape[:-1] + c.shape[-1:])

    def compute_accuracy(self, x, u, J):
        """gusto.py:203-223: rho_k = model error / (J + model approximation) about the iterate (x_k, u_k)."""
        L.require_gpu()
        x = np.asarray(x, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        single = (x.ndim == 2)
        n, m = self.model.n_x, self.model.n_u
        xb, ub = x.reshape((-1,) + x.shape[-2:]), u.reshape((-1,) + u.shape[-2:])
        xk, uk = self.x_k.reshape(xb.shape), self.u_k.reshape(ub.shape)
        Bt, N = ub.shape[0], ub.shape[1]
        fk, Ak, Bk = self.model.get_continuous_dynamics(xk[:, :-1].reshape(-1, n), uk.reshape(-1, m))
        f, _, _ = self.model.get_continuous_dynamics(xb[:, :-1].reshape(-1, n), ub.reshape(-1, m))
        dev = [L.to_dev(np.ascontiguousarray(a)) for a in (fk, Ak, Bk, f, xb, xk, ub, uk, self.f_scale,
                                                            np.broadcast_to(np.asarray(J, dtype=np.float64), (Bt,)))]
        rho = L.empty((Bt,))
        L.check(L.lib().srcb200_gusto_accuracy_batch(n, m, N, Bt, float(self.dt), *[L.ptr(a) for a in dev], L.ptr(rho),
                                                     None, None, L.stream_ptr()))
        r = L.to_host(rho)
        return float(r[0]) if single else r
