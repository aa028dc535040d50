"""numpy restatement of sofacontrol/mor/pod.py.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).
PINNED against the imported reference and the pod_model.pkl known answer (Sigma + tol 5e-5 -> 36 modes)."""
import numpy as np
from . import utils_np as U


class PODNP:
    """pod.py:9-78."""

    def __init__(self, info):
        self.q_ref, self.v_ref = info['q_ref'], info['v_ref']
        self.x_ref = U.qv2x(self.q_ref, self.v_ref)
        self.U = info['U']
        self.V = np.kron(np.eye(2), self.U)
        self.rom_dim = self.U.shape[1]

    def compute_FO_state(self, q=None, v=None, x=None):
        if q is not None:
            return self.U @ q + self.q_ref
        if v is not None:
            return self.U @ v + self.v_ref
        if x is not None:
            return self.V @ x + self.x_ref
        raise RuntimeError('Must specify vector type')

    def compute_RO_state(self, qf=None, vf=None, xf=None):
        if qf is not None:
            return self.U.T @ (qf - self.q_ref)
        if vf is not None:
            return self.U.T @ (vf - self.v_ref)
        if xf is not None:
            return self.V.T @ (xf - self.x_ref)
        raise RuntimeError('Must specify vector type')

    def compute_RO_matrix(self, matrix, left=False, right=False):
        if left == right:
            return self.U.T @ matrix @ self.U
        return self.U.T @ matrix if left else matrix @ self.U


def energy_mode_count(S, tol):
    """pod.py:193-199 -- smallest i >= 1 with sum(S[i:]^2)/sum(S^2) <= tol."""
    s2 = S ** 2
    i = 0
    while (np.sum(s2[i:]) / np.sum(s2)) > tol or i == 0:
        i += 1
    return i


def compute_POD(snapshots, tol, rom_dim=None):
    """pod.py:181-200 -- thin SVD + energy truncation; rom_dim is ignored by the reference."""
    U_full, S, _ = np.linalg.svd(snapshots, full_matrices=False)
    nb = energy_mode_count(S, tol)
    return U_full, U_full[:, :nb], nb, S


def subspace_angle(U1, U2):
    """Largest principal angle between the column spaces of two orthonormal bases (the POD parity metric)."""
    s = np.linalg.svd(U1.T @ U2, compute_uv=False)
    # sin(theta_max) from the projector residual is better conditioned than acos near 0
    R = U2 - U1 @ (U1.T @ U2)
    return float(np.arcsin(min(1.0, np.linalg.norm(R, 2)))), float(s.min())
