"""State-layout helpers and containers of sofacontrol/utils.py that the hot path uses (utils.py:8-16, 129-159):
pure data plumbing; plus the zero-order-hold discretisation `zoh_affine` / `zoh_linear` (utils.py:302-335), which runs
the batched Pade-13 expm kernel of csrc/expm.cu, and `extract_AB` (utils.py:251-286), the batched second-order ->
first-order conversion of csrc/bank.cu used when a TPWL bank is built."""
import os
import pickle

import numpy as np


class QuadraticCost:
    """utils.py:8-16."""

    def __init__(self, Q=None, R=None, Qf=None):
        self.Qf = Qf
        self.Q = Q
        self.R = R


def qv2x(q, v):
    """Reduced/full state is x = [v; q]  (utils.py:129-130); extends to stacked points."""
    return np.concatenate((v, q), axis=-1)


def x2qv(x):
    """Returns (q, v) from x = [v; q]  (utils.py:133-142)."""
    half = x.shape[-1] // 2
    if x.ndim == 1:
        return x[half:], x[:half]
    if x.ndim == 2:
        return x[:, half:], x[:, :half]
    raise IndexError('Unable to process x.ndim > 2')


def vq2qv(x):
    """utils.py:144-146."""
    q, v = x2qv(x)
    return np.hstack((q, v))


def save_data(filename, data):
    """utils.py:148-153."""
    folder = os.path.split(filename)[0]
    if folder and not os.path.isdir(folder):
        os.mkdir(folder)
    with open(filename, 'wb') as file:
        pickle.dump(data, file, protocol=pickle.HIGHEST_PROTOCOL)


def load_data(filename):
    """utils.py:156-159."""
    with open(filename, 'rb') as file:
        return pickle.load(file)


def _zoh_device(A, B, d, dt):
    from . import _lib as L
    L.require_gpu()
    A = np.asarray(A, dtype=np.float64)
    single = (A.ndim == 2)
    n, m = A.shape[-1], np.asarray(B).shape[-1]
    Ad = L.to_dev(A.reshape(-1, n, n))
    Bd = L.to_dev(np.asarray(B, dtype=np.float64).reshape(-1, n, m))
    dd = L.to_dev(np.asarray(d, dtype=np.float64).reshape(-1, n))
    wsb = int(L.lib().srcb200_zoh_workspace(n, m, Ad.shape[0]))
    ws = L.empty((wsb // 8 + 1,))
    L.check(L.lib().srcb200_zoh_batch(n, m, Ad.shape[0], float(dt), L.ptr(Ad), L.ptr(Bd), L.ptr(dd), L.ptr(Ad), L.ptr(Bd),
                                      L.ptr(dd), L.ptr(ws), ws.numel() * 8, L.stream_ptr()))
    res = (L.to_host(Ad), L.to_host(Bd), L.to_host(dd))
    return tuple(r[0] for r in res) if single else res


def zoh_affine(A, B, d, dt):
    """utils.py:322-335: exact discretisation of x' = A x + B u + d under zero-order hold -> (A_d, B_d, d_d).
    One system or a stack with a leading axis; runs srcb200_zoh_batch (scaling-and-squaring Pade-13 of the
    (n+m+1)^2 augmented matrix, one CTA per system)."""
    return _zoh_device(A, B, d, dt)


def zoh_linear(A, B, dt):
    """utils.py:302-319 -> (A_d, B_d).  Same kernel with a zero affine column (the extra zero column of the
    augmented matrix does not touch the A_d / B_d blocks of its exponential)."""
    A = np.asarray(A, dtype=np.float64)
    Ad, Bd, _ = _zoh_device(A, B, np.zeros(A.shape[:-1]), dt)
    return Ad, Bd


def extract_AB(K, D, M, H):
    """utils.py:251-286 (dense branch): first-order (A, B) of  M q'' + D q' + K q = H u  with x = [v; q]:
    A = [[-inv(M) D, -inv(M) K], [I, 0]], B = [[inv(M) H], [0]].  One system or a stack (leading axis); runs
    csrc/control.cu: bank_point_kernel."""
    from . import _lib as L
    L.require_gpu()
    K = np.asarray(K, dtype=np.float64)
    single = (K.ndim == 2)
    r, m = K.shape[-1], np.asarray(H).shape[-1]
    dev = [L.to_dev(np.asarray(a, dtype=np.float64).reshape((-1,) + s)) for a, s in
           ((K, (r, r)), (D, (r, r)), (M, (r, r)), (H, (r, m)))]
    cnt = dev[0].shape[0]
    A, B = L.empty((cnt, 2 * r, 2 * r)), L.empty((cnt, 2 * r, m))
    L.check(L.lib().srcb200_tpwl_bank_point_batch(r, m, cnt, *[L.ptr(a) for a in dev], None, None, L.ptr(A), L.ptr(B),
                                                  None, L.stream_ptr()))
    Ah, Bh = L.to_host(A), L.to_host(B)
    return (Ah[0], Bh[0]) if single else (Ah, Bh)
