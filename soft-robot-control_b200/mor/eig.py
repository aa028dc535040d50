"""Leading eigenpairs of the snapshot Gram matrix -- what compute_POD (sofacontrol/mor/pod.py:181-200) actually needs.

The reference takes the FULL SVD of the (nf x ns) snapshot matrix and then keeps the first i modes with
sum_{j>=i} S_j^2 / sum_j S_j^2 <= tol (pod.py:193-199).  Only those i modes (36 for the Diamond fixture) and the total
energy sum_j S_j^2 = trace(X^T X) enter the result, so on the Gram route the ns x ns eigenproblem shrinks to a
LEADING-eigenpair problem:

    block subspace iteration with Rayleigh-Ritz on G = X^T X (ns x ns, symmetric PSD)
        Z = G Q                      DMMA GEMM (csrc/gemm.cu), 2 ns^2 b flop -- the only O(ns^2) work
        T = Q^T Z   (b x b)          DMMA GEMM
        T = W diag(theta) W^T        one-CTA Jacobi kernel (csrc/eig.cu)
        Y = Q W, R = Z W - Y theta   Ritz vectors / residuals
        Q <- orth((Z W) theta^-1)    SVQB: C = P^T P (DMMA SYRK), C = V L V^T (Jacobi kernel), Q = P V L^-1/2, twice

There is no cuSOLVER / LAPACK call on this path.  Convergence is geometric in lambda_{b+1} / lambda_i, i.e. a handful
of iterations for POD spectra; the block grows (64 -> 128 -> 160) if the energy rule needs more modes than it holds.
The linear-algebra back end is injectable (`ops`) so the host logic is tested on CPU with numpy stand-ins
(tests/test_pod_host.py) and on two gloo ranks (tests/test_multi_gpu_cpu.py).
"""
import numpy as np

from .. import _lib as L

MAX_BLOCK = 160          # csrc/eig.cu keeps the b x b matrix in one CTA's shared memory


class DeviceOps:
    """CUDA tensors + libsrcb200 kernels."""

    def __init__(self):
        self.torch = L.torch_mod()

    def matmul(self, A, B, transA=False):
        K, N = B.shape
        M = A.shape[1] if transA else A.shape[0]
        C_ = L.empty((M, N))
        L.check(L.lib().srcb200_dgemm(int(transA), M, N, K, 1.0, L.ptr(A), A.stride(0), L.ptr(B), B.stride(0),
                                      L.ptr(C_), C_.stride(0), L.stream_ptr()))
        return C_

    def gram(self, Z):
        nf, ns = Z.shape
        G = L.empty((ns, ns))
        L.check(L.lib().srcb200_pod_gram(nf, ns, L.ptr(Z), Z.stride(0), L.ptr(G), G.stride(0), 0, L.stream_ptr()))
        return G

    def eig_psd(self, T):
        n = T.shape[0]
        ev, V = L.empty((n,)), L.empty((n, n))
        L.check(L.lib().srcb200_sym_eig_psd(n, L.ptr(T), T.stride(0), L.ptr(ev), L.ptr(V), V.stride(0), None,
                                            L.stream_ptr()))
        return ev, V

    def randn(self, rows, cols, seed):
        g = self.torch.Generator(device='cuda')
        g.manual_seed(seed)
        return self.torch.randn((rows, cols), dtype=self.torch.float64, device='cuda', generator=g)


class TorchOps:
    """Plain torch stand-ins (CPU tensors): used by the CPU tests of the host logic only."""

    def __init__(self):
        import torch
        self.torch = torch

    def matmul(self, A, B, transA=False):
        return (A.t() if transA else A) @ B

    def gram(self, Z):
        return Z.t() @ Z

    def eig_psd(self, T):
        lam, V = self.torch.linalg.eigh(0.5 * (T + T.t()))
        return self.torch.flip(lam, (0,)).clamp_min(0.0), self.torch.flip(V, (1,)).contiguous()

    def randn(self, rows, cols, seed):
        g = self.torch.Generator()
        g.manual_seed(seed)
        return self.torch.randn((rows, cols), dtype=self.torch.float64, generator=g)


def energy_mode_count(theta, total, tol):
    """pod.py:193-199 from the leading eigenvalues `theta` (descending, 1-D tensor) and the total energy
    `total` = trace(G): smallest i >= 1 with (total - sum(theta[:i])) / total <= tol, or None if the block does not
    reach the tolerance."""
    tail = (total - theta.cumsum(0)) / total
    ok = (tail <= tol).nonzero()
    return (int(ok[0].item()) + 1) if ok.numel() else None


def _orth(P, ops, passes=2):
    """SVQB orthonormalisation of the columns of P through the small Gram matrix (no Cholesky, no QR)."""
    torch = ops.torch
    for _ in range(passes):
        lam, V = ops.eig_psd(ops.gram(P))
        lam = lam.clamp_min(lam[0] * 1e-30 + 1e-300)
        P = ops.matmul(P, (V / lam.sqrt()).contiguous())
    return P


def leading_eigenpairs(G, tol, block=64, guard=4, res_tol=1e-13, max_iter=80, ops=None, seed=20260501):
    """Leading eigenpairs of the symmetric PSD matrix G (n x n tensor) -- enough of them for the energy rule at
    tolerance `tol`.  Returns (theta (k,), Y (n, k) orthonormal, nb, info) with nb = mode count of pod.py:193-199,
    k >= nb eigenpairs converged to a residual ||G y - theta y|| <= res_tol * theta_1."""
    ops = ops or DeviceOps()
    torch = ops.torch
    n = G.shape[0]
    total = torch.diagonal(G).sum()
    if n <= MAX_BLOCK:
        theta, Y = ops.eig_psd(G)                       # small problem: one Jacobi call, no iteration
        nb = energy_mode_count(theta, total, tol) or n
        return theta, Y, nb, {'iterations': 0, 'block': n, 'direct': True}
    b = min(block, MAX_BLOCK)
    Q = _orth(ops.randn(n, b, seed), ops)
    info = {'iterations': 0, 'block': b, 'direct': False}
    prev = None
    for it in range(max_iter):
        info['iterations'] = it + 1
        Z = ops.matmul(G, Q)
        T = ops.matmul(Q, Z, transA=True)
        theta, W = ops.eig_psd(T)
        ZW = ops.matmul(Z, W)
        Y = ops.matmul(Q, W)
        res = (ZW - Y * theta).norm(dim=0) / theta[0]
        nb = energy_mode_count(theta[:b - guard], total, tol)      # the trailing `guard` Ritz values are not trusted
        if nb is not None:
            worst = float(res[:nb + guard].max())
            # converged, or stagnated at the rounding floor of the residual (~ eps ||G|| sqrt(n))
            if worst <= res_tol or (prev is not None and worst <= 1e-11 and worst > 0.5 * prev):
                info['residual'] = worst
                return theta, Y, nb, info
            prev = worst
        elif bool((res[:max(b - 2 * guard, 1)] <= 1e-6).all()) or it >= 8:
            # the block does not hold the energy tail: grow it around the current Ritz vectors
            if b >= MAX_BLOCK:
                raise RuntimeError("compute_POD: tolerance %g needs more than %d modes; call with "
                                   "full_spectrum=True" % (tol, b - guard))
            b_new = min(MAX_BLOCK, 2 * b)
            Q = _orth(torch.cat((Y, ops.randn(n, b_new - b, seed + it + 1)), dim=1).contiguous(), ops)
            b = b_new
            info['block'] = b
            prev = None
            continue
        # power step on the Ritz vectors, columns rescaled so that the small Gram matrix stays well conditioned
        Q = _orth((ZW / theta.clamp_min(theta[0] * 1e-14)).contiguous(), ops)
    raise RuntimeError("leading_eigenpairs: no convergence in %d iterations" % max_iter)
