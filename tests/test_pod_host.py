"""CPU-only: host logic of the leading-eigenpair POD solver (mor/eig.py) with torch stand-ins for the device kernels:
the block subspace iteration finds the same modes / mode count as the reference's full SVD + energy rule
(sofacontrol/mor/pod.py:181-200), grows its block when the tolerance needs more modes, and fails loudly past the
one-CTA Jacobi kernel's size."""
import numpy as np
import pytest
import torch

import sofacontrol_b200.synth as synth
from sofacontrol_b200.mor import eig
from oracle import pod_np


def _solve(X, tol, **kw):
    Xt = torch.from_numpy(X)
    G = Xt.t() @ Xt
    theta, Y, nb, info = eig.leading_eigenpairs(G, tol, ops=eig.TorchOps(), **kw)
    U = (Xt @ (Y[:, :nb] / theta[:nb].sqrt())).numpy()
    return theta.sqrt().numpy(), U, nb, info


def test_subspace_iteration_matches_reference_svd_route():
    X, _, _ = synth.pod_snapshots(1500, 400, seed=5)
    _, Uo, nbo, So = pod_np.compute_POD(X, 5e-5)
    S, U, nb, info = _solve(X, 5e-5)
    assert not info['direct'] and info['block'] == 64 and info['iterations'] <= 12
    assert nb == nbo
    assert np.abs(S[:nb] - So[:nb]).max() <= 1e-12 * So[0]
    assert pod_np.subspace_angle(Uo, U)[0] < 1e-8


def test_energy_rule_equals_reference_loop_on_fixture_sigma(golden):
    """Known answer of the reference fixture (pod_model.pkl Sigma, tol 5e-5 -> 36 modes) through trace + cumsum."""
    gk = golden("pod_known.npz")
    lam = torch.from_numpy(gk['Sigma'] ** 2)
    assert eig.energy_mode_count(lam, lam.sum(), float(gk['tol'])) == int(gk['modes']) == 36
    assert eig.energy_mode_count(lam[:20], lam.sum(), float(gk['tol'])) is None      # block too small: caller grows it


def test_block_grows_when_the_tolerance_needs_more_modes():
    rng = np.random.default_rng(0)
    ns = 300
    Q1, _ = np.linalg.qr(rng.normal(size=(900, ns)))
    Q2, _ = np.linalg.qr(rng.normal(size=(ns, ns)))
    sv = 0.93 ** np.arange(ns)                       # slow decay: tol 1e-5 keeps ~ 80 modes
    X = (Q1 * sv) @ Q2.T
    _, Uo, nbo, So = pod_np.compute_POD(X, 1e-5)
    S, U, nb, info = _solve(X, 1e-5)
    assert nbo > 60 and nb == nbo and info['block'] in (128, 160)
    assert pod_np.subspace_angle(Uo, U)[0] < 1e-8


def test_fails_loudly_past_the_kernel_size():
    rng = np.random.default_rng(1)
    X = rng.normal(size=(600, 400))                  # flat spectrum: tol 1e-6 needs ~ all 400 modes
    with pytest.raises(RuntimeError, match="full_spectrum"):
        _solve(X, 1e-6)


def test_small_problems_go_straight_to_the_jacobi_call():
    X, _, _ = synth.pod_snapshots(300, 60, seed=7)
    _, Uo, nbo, So = pod_np.compute_POD(X, 1e-4)
    S, U, nb, info = _solve(X, 1e-4)
    assert info['direct'] and nb == nbo and pod_np.subspace_angle(Uo, U)[0] < 1e-8
