"""numpy restatement of sofacontrol/lqr/lqr.py (solve_riccati, dare) and lqr/traj_tracking_lqr.py
(TrajTrackingLQR.perform_dlqr_recursion).  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).
PINNED bit-for-bit against the unmodified reference modules in tests/test_oracle_vs_reference.py (lqr.py imports the
absent `control` package at module top for CLQR only; an empty stub module is injected like `osqp`,
oracle/refimport.py); golden vectors: tests/golden/lqr_gains.npz.
"""
import numpy as np
import scipy.linalg
from scipy.interpolate import interp1d


def solve_riccati(A, B, Q, R):
    """lqr.py:6-21 -- value iteration until the gain moves by <= 1e-4 (Frobenius); returns (L, P), u = +L x."""
    n = A.shape[0]
    m = B.shape[1]
    P = np.zeros((n, n))
    L = np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
    Lold = np.inf * np.ones((m, n))
    iters = 0
    while (np.linalg.norm(L - Lold)) > 1e-4:
        Lold = L
        P = A.T @ P @ A - A.T @ P @ B @ np.linalg.inv(R + B.T @ P @ B) @ (B.T @ P @ A) + Q
        L = -np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
        iters += 1
    solve_riccati.last_iterations = iters
    return L, P


def dare(Ad, Bd, Q, R):
    """lqr.py:24-31."""
    P = scipy.linalg.solve_discrete_are(Ad, Bd, Q, R)
    K = -scipy.linalg.inv(Bd.T @ P @ Bd + R) @ (Bd.T @ P @ Ad)
    return K, P


class TrajTrackingLQRNP:
    """traj_tracking_lqr.py:5-48."""

    def __init__(self, dt, model, cost_params):
        self.dt, self.model, self.cost_params = dt, model, cost_params
        self.x_bar = self.u_bar = None

    def compute_policy(self, target):
        K, _ = self.perform_dlqr_recursion(target)
        return self.x_bar, self.u_bar, K

    def perform_dlqr_recursion(self, target):
        P = [self.cost_params.Q]
        K, x_nom, u_nom = [], [], []
        x_nom_interp = interp1d(target.t, target.x, axis=0)
        u_nom_interp = interp1d(target.t, target.u, axis=0)
        final_time = target.t[-1]
        nbr_steps = int(final_time / self.dt)
        for i in reversed(range(nbr_steps)):
            t_step = i * self.dt
            x_nom_i = x_nom_interp(t_step)
            u_nom_i = u_nom_interp(t_step)
            A, B, d = self.model.get_jacobians(x_nom_i, dt=self.dt)
            x_nom.append(x_nom_i)
            u_nom.append(u_nom_i)
            K.append(-1. * np.linalg.solve(self.cost_params.R + B.T @ P[-1] @ B, B.T @ P[-1] @ A))
            P.append(self.cost_params.Q + K[-1].T @ self.cost_params.R @ K[-1] +
                     (A + B @ K[-1]).T @ P[-1] @ (A + B @ K[-1]))
        K = np.flip(np.asarray(K), axis=0)
        P = np.flip(np.asarray(P), axis=0)
        self.x_bar = np.flip(np.asarray(x_nom), axis=0)
        self.u_bar = np.flip(np.asarray(u_nom), axis=0)
        return K, P


def extract_AB(K, D, M, H):
    """utils.py:251-286, dense branch."""
    Minv = np.linalg.inv(M)
    K_tilde = Minv @ K
    D_tilde = Minv @ D
    H_tilde = Minv @ H
    A11 = -D_tilde
    A12 = -K_tilde
    A21 = np.eye(np.shape(A11)[0])
    A22 = np.zeros(np.shape(A12))
    A = np.block([[A11, A12], [A21, A22]])
    B = np.block([[H_tilde], [np.zeros(np.shape(H_tilde))]])
    return A, B


def continuous_tpwl_point(K, D, M, H, f, q):
    """tpwl_utils.py:263-276 add_continuous_TPWL -> (A_c, B_c, d_c)."""
    A, B = extract_AB(K, D, M, H)
    b_normalized = np.linalg.solve(M, f + K @ q)
    d = np.hstack((b_normalized, np.zeros(np.shape(b_normalized))))
    return A, B, d


def gusto_accuracy(model, x_k, u_k, x, u, J, dt, f_scale):
    """scp/gusto.py:203-223 compute_accuracy with a duck-typed model (get_continuous_dynamics -> f, A, B)."""
    error = 0
    approx = 0
    for i in range(x.shape[0] - 1):
        fk, Ak, Bk = model.get_continuous_dynamics(x_k[i, :], u_k[i, :])
        f, _, _ = model.get_continuous_dynamics(x[i, :], u[i, :])
        f_approx = fk + Ak @ (x[i, :] - x_k[i, :]) + Bk @ (u[i, :] - u_k[i, :])
        error += dt * np.linalg.norm(np.multiply(f_scale, f - f_approx), 2)
        approx += dt * np.linalg.norm(np.multiply(f_scale, f_approx), 2)
    return error / (J + approx)
