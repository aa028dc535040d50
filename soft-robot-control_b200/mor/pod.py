"""POD on B200 -- drop-in for sofacontrol/mor/pod.py (POD, compute_POD, run_POD, load_POD, pod_config).

compute_POD replaces np.linalg.svd of the (nf x ns) snapshot matrix (pod.py:191) by the Gram route:
    G = X^T X  (ns x ns)            FP64 tensor-core (DMMA) SYRK kernel, csrc/gemm.cu  [+ NCCL all-reduce when the
                                     rows (DOFs) of X are sharded across GPUs, see compute_POD_sharded]
    leading eigenpairs of G         block subspace iteration + Rayleigh-Ritz on this repo's DMMA GEMM and one-CTA
                                     Jacobi kernel (mor/eig.py, csrc/eig.cu) -- the energy rule only needs the kept
                                     modes and sum(S^2) = trace(G), not the full spectrum
    U = X V S^-1                    FP64 DMMA GEMM kernel
followed by the reference's energy truncation rule (pod.py:193-199).  Singular values below ~sqrt(eps)*S_max are
not resolved by the Gram route (the condition number is squared); the leading modes that the energy rule keeps
are, to a subspace angle < 1e-8 (tests/test_pod_gpu.py).

Return-value note: the reference returns every left singular vector / singular value (U_full, Sigma of length
min(nf, ns)); here U_full / Sigma hold the leading block the solver resolved (>= nbModes + 4 modes, 64 by default).
`compute_POD(..., full_spectrum=True)` returns all of them through torch.linalg.eigh (cuSOLVER) -- an explicit
opt-in library call that nothing else in this package uses.
"""
import os

import numpy as np

from .. import _lib as L
from .. import utils as scutils


def gram_device(X, G=None, accumulate=False):
    """X: CUDA tensor (nf, ns) row-major -> G (ns, ns) = X^T X (both triangles)."""
    nf, ns = X.shape
    if G is None:
        G = L.empty((ns, ns))
        accumulate = False
    L.check(L.lib().srcb200_pod_gram(nf, ns, L.ptr(X), X.stride(0), L.ptr(G), G.stride(0), int(accumulate),
                                     L.stream_ptr()))
    return G


def dgemm_device(A, B, transA=False, alpha=1.0):
    """C = alpha * op(A) @ B on CUDA tensors (row-major, last dim contiguous)."""
    K, N = B.shape
    M = A.shape[1] if transA else A.shape[0]
    assert (A.shape[0] if transA else A.shape[1]) == K
    C_ = L.empty((M, N))
    L.check(L.lib().srcb200_dgemm(int(transA), M, N, K, float(alpha), L.ptr(A), A.stride(0), L.ptr(B), B.stride(0),
                                  L.ptr(C_), C_.stride(0), L.stream_ptr()))
    return C_


def energy_mode_count_device(s2, tol):
    """pod.py:193-199 on a CUDA tensor of squared singular values (descending): smallest i >= 1 with
    sum(s2[i:]) / sum(s2) <= tol.  Returns a python int (one tiny device->host read)."""
    torch = L.torch_mod()
    total = s2.sum()
    tail = total - torch.cumsum(s2, 0)          # tail[i-1] = sum(s2[i:])
    ok = (tail / total) <= tol
    idx = torch.nonzero(ok)
    return int(idx[0].item()) + 1 if idx.numel() else int(s2.numel())


def _finish_pod(Xd, G, tol, full_U, gemm=None, ops=None):
    """Leading eigenpairs of the (replicated) Gram matrix -> U = X V S^-1 for the kept modes (full_U: for every mode
    of the resolved block).  Returns (U, nbModes, S)."""
    from . import eig
    lam, V, nb, _ = eig.leading_eigenpairs(G, tol, ops=ops)
    lam = lam.clamp_min(0.0)
    S = lam.sqrt()
    keep = V.shape[1] if full_U else nb
    # unresolved directions (S below the Gram route's sqrt(eps) floor) are returned as zero columns, never inf / nan
    ok = S[:keep] > 1e-7 * S[0]
    Vs = (V[:, :keep] * (ok / S[:keep].clamp_min(np.finfo(np.float64).tiny))).contiguous()
    U = (gemm or dgemm_device)(Xd, Vs)           # (nf, keep)
    return U, nb, S


def compute_POD_device(Xd, tol, full_U=False):
    """CUDA tensor X (nf, ns) -> (U (nf, nb or block), nbModes, S (block)) as CUDA tensors."""
    G = gram_device(Xd)
    return _finish_pod(Xd, G, tol, full_U)


def compute_POD_sharded(X_local, tol, group=None, gram=None, gemm=None, ops=None, chunks=1):
    """Row-sharded POD: every rank holds a block of DOF rows X_g (nf_g x ns).  G = sum_g X_g^T X_g through an
    all-reduce (NCCL over NVLink on GPUs), the small leading-eigenpair solve is replicated (deterministic start:
    every rank gets the same modes), U_g = X_g V S^-1 stays row-sharded.  Returns (U_local, nbModes, S).
    chunks > 1: the local rows are contracted in `chunks` row slabs and each slab's partial Gram is reduced on a side
    stream while the next slab computes (parallel.overlapped_gram_allreduce).  `gram` / `gemm` / `ops` default to
    the DMMA kernels; the gloo CPU tests of the multi-process logic inject plain torch stand-ins."""
    from ..parallel import allreduce_sum_, overlapped_gram_allreduce
    if ops is None and (gram is not None or gemm is not None):
        from . import eig
        ops = eig.TorchOps()                     # injected stand-ins (CPU tests): the small solves follow suit
    if chunks > 1 and gram is None:
        G = overlapped_gram_allreduce(X_local, chunks, group)
    else:
        G = (gram or gram_device)(X_local)
        allreduce_sum_(G, group)
    return _finish_pod(X_local, G, tol, False, gemm, ops)


def _full_spectrum(G):
    """Opt-in only: every eigenpair of the Gram matrix through torch.linalg.eigh (cuSOLVER)."""
    torch = L.torch_mod()
    lam, V = torch.linalg.eigh(G)                # ascending
    return torch.flip(lam, (0,)).clamp_min(0.0), torch.flip(V, (1,)).contiguous()


def compute_POD(snapshots, tol, rom_dim=None, full_spectrum=False):
    """pod.py:181-200.  snapshots: (nf x num_snapshots) host array.  Returns (U_full, U, nbModes, Sigma) like the
    reference (rom_dim is ignored there too).  U_full / Sigma cover the leading block of modes the solver resolved
    (all min(nf, ns) of them with full_spectrum=True, see the module docstring); U = U_full[:, :nbModes]."""
    L.require_gpu()
    from . import eig
    Xd = L.to_dev(np.asarray(snapshots, dtype=np.float64))
    nf, ns = Xd.shape
    if nf < ns:
        # thin SVD has min(nf, ns) modes: work on X X^T instead (same kernel on the transposed matrix); its
        # eigenvectors ARE the left singular vectors
        G = gram_device(Xd.t().contiguous())     # (nf x nf) = X X^T
        lam, Uf = _full_spectrum(G) if full_spectrum else eig.leading_eigenpairs(G, tol)[:2]
        lam = lam.clamp_min(0.0)
        nb = eig.energy_mode_count(lam, L.torch_mod().diagonal(G).sum(), tol) or int(lam.numel())
        U_full = L.to_host(Uf)
        return U_full, U_full[:, 0:nb], nb, L.to_host(lam.sqrt())
    if full_spectrum:
        G = gram_device(Xd)
        lam, V = _full_spectrum(G)
        S = lam.sqrt()
        nb = energy_mode_count_device(lam, tol)
        ok = S > 1e-7 * S[0]
        U = dgemm_device(Xd, (V * (ok / S.clamp_min(np.finfo(np.float64).tiny))).contiguous())
    else:
        U, nb, S = compute_POD_device(Xd, tol, full_U=True)
    U_full = L.to_host(U)
    return U_full, U_full[:, 0:nb], nb, L.to_host(S)


class POD:
    """pod.py:9-78.  The projections run the DMMA GEMM on the device for stacked inputs; the object keeps the
    same attributes (q_ref, v_ref, x_ref, U, V, rom_dim)."""

    def __init__(self, POD_info):
        self.q_ref = POD_info['q_ref']
        self.v_ref = POD_info['v_ref']
        self.x_ref = scutils.qv2x(self.q_ref, self.v_ref)
        self.U = POD_info['U']
        self.rom_dim = self.U.shape[1]
        self._V = None
        self._Ud = None

    @property
    def V(self):
        """kron(I_2, U) (pod.py:18), built lazily: 2 nf x 2 r."""
        if self._V is None:
            self._V = np.kron(np.eye(2), self.U)
        return self._V

    def _U_dev(self):
        if self._Ud is None:
            L.require_gpu()
            self._Ud = L.to_dev(np.asarray(self.U, dtype=np.float64))
        return self._Ud

    def _lift(self, r, ref):
        """U @ r + ref for r (rom_dim,) or (rom_dim, N)."""
        r = np.asarray(r, dtype=np.float64)
        cols = r.reshape(self.rom_dim, -1)
        out = L.to_host(dgemm_device(self._U_dev(), L.to_dev(cols)))
        return out.reshape((self.U.shape[0],) + r.shape[1:]) + (ref if r.ndim == 1 else np.asarray(ref)[:, None])

    def _project(self, f, ref):
        """U^T @ (f - ref) for f (nf,) or (nf, N)."""
        f = np.asarray(f, dtype=np.float64)
        diff = f - (ref if f.ndim == 1 else np.asarray(ref)[:, None])
        cols = L.to_dev(diff.reshape(self.U.shape[0], -1))
        out = L.to_host(dgemm_device(self._U_dev(), cols, transA=True))
        return out.reshape((self.rom_dim,) + f.shape[1:])

    def compute_FO_state(self, q=None, v=None, x=None):
        """pod.py:22-37."""
        if q is not None:
            return self._lift(q, self.q_ref)
        elif v is not None:
            return self._lift(v, self.v_ref)
        elif x is not None:
            x = np.asarray(x, dtype=np.float64)
            r = self.rom_dim
            return scutils.qv2x(self._lift(x[r:], self.q_ref), self._lift(x[:r], self.v_ref)) if x.ndim == 1 else \
                np.concatenate((self._lift(x[:r], self.v_ref), self._lift(x[r:], self.q_ref)), axis=0)
        raise RuntimeError('Must specify vector type')

    def compute_RO_state(self, qf=None, vf=None, xf=None):
        """pod.py:39-54."""
        if qf is not None:
            return self._project(qf, self.q_ref)
        elif vf is not None:
            return self._project(vf, self.v_ref)
        elif xf is not None:
            xf = np.asarray(xf, dtype=np.float64)
            nf = self.U.shape[0]
            return np.concatenate((self._project(xf[:nf], self.v_ref), self._project(xf[nf:], self.q_ref)), axis=0)
        raise RuntimeError('Must specify vector type')

    def compute_RO_matrix(self, matrix, left=False, right=False):
        """pod.py:56-72: U^T M U, U^T M or M U for a dense (or scipy coo) matrix."""
        if hasattr(matrix, 'toarray'):
            matrix = matrix.toarray()
        if not isinstance(matrix, np.ndarray):
            raise RuntimeError('Matrix is not numpy ndarray or sparse coo_matrix')
        if matrix.ndim == 1:                      # U^T f for a force vector (tpwl_utils.py:95-96): one column
            return self.compute_RO_matrix(matrix[:, None], left=left, right=right)[:, 0]
        Md = L.to_dev(matrix)
        Ud = self._U_dev()
        if (left and right) or (not left and not right):
            return L.to_host(dgemm_device(Ud, dgemm_device(Md, Ud), transA=True))
        if left:
            return L.to_host(dgemm_device(Ud, Md, transA=True))
        return L.to_host(dgemm_device(Md, Ud))

    def get_info(self):
        """pod.py:74-78."""
        return {'q_ref': self.q_ref, 'v_ref': self.v_ref, 'U': self.U, 'type': 'POD'}


class pod_config():
    """pod.py:81-90."""

    def __init__(self):
        self.pod_type = 'v'  # current 'v' or 'q'
        self.pod_tolerance = 0.0001
        self.preprocess = []  # string names of preprocess options to run on data
        self.preprocess_args = {'nbr_clusters': 0}


def load_POD(POD_file):
    """pod.py:93-107."""
    if not os.path.isfile(POD_file):
        raise RuntimeError('POD file specified is not a valid file')
    POD_data = scutils.load_data(POD_file)
    return POD(POD_data['POD_info'])


def get_snapshots(data, pod_type):
    """pod.py:144-154."""
    if pod_type == 'q':
        return np.asarray(data['q']) - data['q'][0]
    elif pod_type == 'v':
        return np.asarray(data['v'])
    elif pod_type == 'a':
        return np.asarray(data['v+']) - np.asarray(data['v'])
    raise RuntimeError('unknown pod_type')


def process_snapshots(snapshots, preprocess, args):
    """pod.py:157-178 ('clustering' needs sklearn KMeans and is outside the hot path: not supported here)."""
    if 'normalize' in preprocess:
        snapshots = (snapshots - snapshots.min(axis=0)) / (snapshots.max(axis=0) + 1e-15 - snapshots.min(axis=0))
    if 'substract_mean' in preprocess:
        snapshots = snapshots - snapshots.mean(axis=0, keepdims=True)
    if 'clustering' in preprocess and args.get('nbr_clusters', 0) > 0:
        raise NotImplementedError("k-means snapshot clustering is optional preprocessing outside the hot path")
    return snapshots


def run_POD(snapshots_file, POD_file, config, rom_dim=None):
    """pod.py:110-141."""
    data = scutils.load_data(snapshots_file)
    snapshots = get_snapshots(data, config.pod_type)
    snapshots = process_snapshots(snapshots, config.preprocess, config.preprocess_args)
    U_full, U, rom_dim, Sigma = compute_POD(snapshots.T, config.pod_tolerance)
    POD_info = {'U': U, 'q_ref': data['q'][0], 'v_ref': np.zeros(data['v'][0].shape)}
    results = {'POD_info': POD_info, 'config': vars(config), 'Sigma': Sigma}
    scutils.save_data(POD_file, results)
    return results
