"""numpy restatement of the state-layout helpers and ZOH discretisation of sofacontrol/utils.py.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import numpy as np
from scipy.sparse.linalg import expm as _ref_expm   # the routine the reference calls (utils.py:4,313)


class QuadraticCost:
    """utils.py:8-16 -- plain container (Q, R, Qf)."""

    def __init__(self, Q=None, R=None, Qf=None):
        self.Q, self.R, self.Qf = Q, R, Qf


def qv2x(q, v):
    """utils.py:129-130 -- reduced/full state is x = [v; q]."""
    return np.concatenate((v, q), axis=-1)


def x2qv(x):
    """utils.py:133-142 -- returns (q, v) from x = [v; q]."""
    h = x.shape[-1] // 2
    if x.ndim == 1:
        return x[h:], x[:h]
    if x.ndim == 2:
        return x[:, h:], x[:, :h]
    raise IndexError('Unable to process x.ndim > 2')


def vq2qv(x):
    """utils.py:144-146."""
    q, v = x2qv(x)
    return np.hstack((q, v))


def zoh_affine(A, B, d, dt):
    """utils.py:302-335 -- exact ZOH of xdot = A x + B u + d via expm of the (n+m+1) augmented matrix.
    The reference calls scipy.sparse.linalg.expm on the dense augmented matrix (Al-Mohy & Higham scaling and
    squaring); the same routine is called here so the restatement is bit-identical."""
    n, m = A.shape[0], B.shape[1]
    M = np.zeros((n + m + 1, n + m + 1))
    M[:n, :n] = A
    M[:n, n:n + m] = B
    M[:n, n + m] = d
    E = _ref_expm(M * dt)
    return E[:n, :n], E[:n, n:n + m], E[:n, n + m]
