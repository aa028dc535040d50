"""Infinite-horizon LQR gains on B200 -- drop-in for sofacontrol/lqr/lqr.py (solve_riccati, dare, DLQR).

`solve_riccati(A, B, Q, R) -> (L, P)` keeps the reference's algorithm literally (lqr.py:6-21: value iteration from
P = 0 until the Frobenius norm of the gain change is <= 1e-4), so its result -- a not fully converged P -- matches
the reference's; `dare(Ad, Bd, Q, R) -> (K, P)` returns the stabilising solution to working precision (the reference
calls scipy.linalg.solve_discrete_are; here a structure-preserving doubling iteration, csrc/control.cu).  Both accept
a leading batch axis on A and B: one CTA per system, e.g. one gain per stored TPWL point
(tpwl/controllers.py:238-246).  `CLQR` needs python-control + slycot in the reference and is out of scope.
"""
import numpy as np

from .. import _lib as L


def _riccati(A, B, Q, R, mode, tol, max_iter):
    L.require_gpu()
    torch = L.torch_mod()
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    single = (A.ndim == 2)
    n, m = A.shape[-1], B.shape[-1]
    Ad, Bd = L.to_dev(A.reshape(-1, n, n)), L.to_dev(B.reshape(-1, n, m))
    Qd, Rd = L.to_dev(np.asarray(Q, dtype=np.float64).reshape(n, n)), L.to_dev(np.asarray(R, dtype=np.float64).reshape(m, m))
    bt = Ad.shape[0]
    K, P, it = L.empty((bt, m, n)), L.empty((bt, n, n)), L.empty((bt,), torch.int32)
    L.check(L.lib().srcb200_dlqr_riccati_batch(n, m, bt, L.ptr(Ad), L.ptr(Bd), L.ptr(Qd), L.ptr(Rd), 1, float(tol),
                                               int(max_iter), int(mode), L.ptr(K), L.ptr(P), L.ptr(it), L.stream_ptr()))
    Kh, Ph, ith = L.to_host(K), L.to_host(P), L.to_host(it)
    return (Kh[0], Ph[0], int(ith[0])) if single else (Kh, Ph, ith)


def solve_riccati(A, B, Q, R, max_iter=1000000):
    """lqr.py:6-21 -> (L, P) with u = +L x."""
    K, P, _ = _riccati(A, B, Q, R, 0, 1e-4, max_iter)
    return K, P


def solve_riccati_info(A, B, Q, R, max_iter=1000000):
    """Same, also returning the number of value-iteration passes (per system)."""
    return _riccati(A, B, Q, R, 0, 1e-4, max_iter)


def dare(Ad, Bd, Q, R):
    """lqr.py:24-31 -> (K, P), K = -inv(B^T P B + R) (B^T P A)."""
    K, P, _ = _riccati(Ad, Bd, Q, R, 1, 1e-15, 200)
    return K, P


class DLQR:
    """lqr.py:34-55: infinite-horizon discrete LQR about a target linearisation."""

    def __init__(self, dt, model, cost_params):
        self.dt = dt
        self.model = model
        self.cost_params = cost_params

    def compute_policy(self, target):
        u_nom = np.atleast_1d(target.u)
        x_nom = target.x
        K = self.compute_gain_matrix(target.A, target.B, self.cost_params.Q, self.cost_params.R)
        return x_nom, u_nom, K

    def compute_gain_matrix(self, A, B, Q, R):
        A = np.asarray(A, dtype=np.float64)
        d_c = np.zeros(A.shape[:-1])
        Ad, Bd, _ = self.model.discretize_dynamics(A_c=A, B_c=B, d_c=d_c, dt=self.dt)
        K, _ = solve_riccati(Ad, Bd, Q, R)
        return K
