"""GPU parity: POD Gram / DMMA GEMM kernels (csrc/gemm.cu) and compute_POD vs the reference's np.linalg.svd route.
Parity metric: same mode count, singular values of the kept modes to 1e-9 relative, subspace angle < 1e-8."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.mark.parametrize("shape", [(1, 1, 1), (7, 5, 3), (130, 129, 17), (257, 300, 64), (512, 256, 1000), (200, 131, 0)])
@pytest.mark.parametrize("transA", [False, True])
def test_dgemm_vs_numpy(shape, transA):
    from sofacontrol_b200.mor import pod
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A = rng.normal(size=(K, M) if transA else (M, K))
    B = rng.normal(size=(K, N))
    C = pod.dgemm_device(_dev(A), _dev(B), transA=transA, alpha=0.5).cpu().numpy()
    ref = 0.5 * ((A.T if transA else A) @ B)
    assert C.shape == (M, N)
    assert np.max(np.abs(C - ref)) <= 1e-12 * max(1.0, np.abs(ref).max()) * max(K, 1)


def test_dgemm_mbarrier_pipeline_is_deterministic_under_load():
    """The K loop hands its shared-memory stages over through full / empty mbarriers (no CTA barrier).  A stage read
    before its copies landed or refilled while a slow warp still reads it would show up as run-to-run differences:
    a many-tile GEMM (persistent CTAs walk ten tiles each, 129 K chunks per tile) and a long-K SYRK are repeated and
    must reproduce their own bits every time, next to the float64 reference; odd leading dimensions take the 8-byte
    copy path."""
    import torch
    from sofacontrol_b200.mor import pod
    g = torch.Generator(device="cuda").manual_seed(7)
    for (M, N, K) in ((2048, 1536, 2056), (1111, 777, 1001)):
        A = torch.randn((M, K), device="cuda", dtype=torch.float64, generator=g)
        B = torch.randn((K, N), device="cuda", dtype=torch.float64, generator=g)
        ref = A @ B
        first = pod.dgemm_device(A, B)
        assert float((first - ref).abs().max()) <= 1e-12 * K
        for _ in range(12):
            assert torch.equal(pod.dgemm_device(A, B), first)
    X = torch.randn((40000, 700), device="cuda", dtype=torch.float64, generator=g)
    G0 = pod.gram_device(X)
    assert float((G0 - X.t() @ X).abs().max()) <= 1e-12 * 40000
    for _ in range(8):
        assert torch.equal(pod.gram_device(X), G0)


@pytest.mark.parametrize("nf,ns", [(100, 3), (333, 129), (4884, 306), (1000, 257)])
def test_gram_vs_numpy(nf, ns):
    from sofacontrol_b200.mor import pod
    rng = np.random.default_rng(nf)
    X = rng.normal(size=(nf, ns))
    G = pod.gram_device(_dev(X)).cpu().numpy()
    ref = X.T @ X
    assert np.max(np.abs(G - ref)) <= 1e-12 * np.abs(ref).max() * nf ** 0.5
    assert np.array_equal(G, G.T)                                   # both triangles written from the same tile
    G2 = pod.gram_device(_dev(X), G=_dev(ref), accumulate=True).cpu().numpy()
    assert relerr(G2, 2 * ref) < 1e-12


@pytest.mark.parametrize("nf,ns", [(700, 1920), (520, 2500)])
def test_gram_and_gemm_super_block_tile_walk(nf, ns):
    """Tile lists longer than one super-block (the SYRK / GEMM kernels enumerate sb x sb blocks of tiles so that the
    tiles in flight share column panels of X): every tile is still visited exactly once."""
    from sofacontrol_b200.mor import pod
    rng = np.random.default_rng(ns)
    X = rng.normal(size=(nf, ns))
    ref = X.T @ X
    G = pod.gram_device(_dev(X)).cpu().numpy()
    assert np.max(np.abs(G - ref)) <= 1e-12 * np.abs(ref).max() * nf ** 0.5 and np.array_equal(G, G.T)
    Xd = _dev(X)
    C = pod.dgemm_device(Xd[:, 128:1024], Xd[:, 128:], transA=True).cpu().numpy()      # strided views, rectangular grid
    assert np.max(np.abs(C - ref[128:1024, 128:])) <= 1e-12 * np.abs(ref).max() * nf ** 0.5


@pytest.mark.parametrize("n", [1, 2, 7, 33, 64, 129, 160])
def test_sym_eig_psd_kernel_vs_numpy(n):
    """One-CTA Jacobi kernel (csrc/eig.cu): eigenvalues to 1e-13 of the largest, V orthonormal, A V = V diag."""
    import torch
    from sofacontrol_b200.mor import eig
    rng = np.random.default_rng(n)
    B = rng.normal(size=(n + 3, n)) * 10.0 ** (-4.0 * np.arange(n) / max(n - 1, 1))   # eigenvalues over 8 decades, like a Rayleigh-Ritz block
    A = B.T @ B
    ev, V = eig.DeviceOps().eig_psd(_dev(A))
    ev, V = ev.cpu().numpy(), V.cpu().numpy()
    ref = np.linalg.eigvalsh(A)[::-1]
    assert np.all(np.diff(ev) <= 0)
    assert np.abs(ev - ref).max() <= 1e-13 * ref[0]
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-9
    assert np.abs(A @ V - V * ev).max() <= 1e-12 * ref[0]


def test_overlapped_gram_single_process_equals_syrk():
    from sofacontrol_b200.mor import pod
    from sofacontrol_b200 import parallel
    rng = np.random.default_rng(3)
    X = _dev(rng.normal(size=(900, 1500)))
    G0 = pod.gram_device(X).cpu().numpy()
    for nb in (1, 3, 5):
        G = parallel.overlapped_gram_allreduce(X, nb).cpu().numpy()
        assert np.abs(G - G0).max() <= 1e-12 * np.abs(G0).max() and np.array_equal(G, G.T)


def test_compute_pod_full_spectrum_opt_in():
    from sofacontrol_b200.mor import pod
    import sofacontrol_b200.synth as synth
    from oracle import pod_np
    X, _, _ = synth.pod_snapshots(800, 300, seed=9)
    Uf, U, nb, S = pod.compute_POD(X, 5e-5, full_spectrum=True)
    _, Uo, nbo, So = pod_np.compute_POD(X, 5e-5)
    assert Uf.shape == (800, 300) and S.shape == (300,) and nb == nbo and np.all(np.isfinite(Uf))
    assert relerr(S[:nb], So[:nb]) < 1e-9 and pod_np.subspace_angle(Uo, U)[0] < 1e-8
    Uf2, U2, nb2, S2 = pod.compute_POD(X, 5e-5)                       # default: leading block only
    assert nb2 == nbo and Uf2.shape[1] == S2.shape[0] >= nb + 4 and pod_np.subspace_angle(Uo, U2)[0] < 1e-8


def test_compute_pod_small_golden(golden):
    from sofacontrol_b200.mor import pod
    import sofacontrol_b200.synth as synth
    from oracle.pod_np import subspace_angle
    gk = golden("pod_known.npz")
    X, _, _ = synth.pod_snapshots(600, 150, seed=5)
    U_full, U, nb, S = pod.compute_POD(X, 5e-5)
    assert nb == int(gk['small_modes']) and U.shape == gk['small_U'].shape
    assert relerr(S[:nb], gk['small_S'][:nb]) < 1e-9
    ang, _ = subspace_angle(gk['small_U'], U)
    assert ang < 1e-8
    assert np.abs(U.T @ U - np.eye(nb)).max() < 1e-9


def test_energy_rule_on_fixture_sigma(golden):
    """Known answer: pod_model.pkl's Sigma with pod_tolerance 5e-5 keeps exactly 36 modes (pod.py:193-199)."""
    from sofacontrol_b200.mor import pod
    gk = golden("pod_known.npz")
    assert pod.energy_mode_count_device(_dev(gk['Sigma'] ** 2), float(gk['tol'])) == int(gk['modes']) == 36


def test_compute_pod_fixture_scale_vs_svd():
    """Fixture-scale twin (4884 x 612, Diamond-like spectrum): same modes and subspace as np.linalg.svd."""
    from sofacontrol_b200.mor import pod
    import sofacontrol_b200.synth as synth
    from oracle import pod_np
    X, _, _ = synth.pod_snapshots(4884, 612, seed=5)
    _, Uo, nbo, So = pod_np.compute_POD(X, 5e-5)
    _, U, nb, S = pod.compute_POD(X, 5e-5)
    assert nb == nbo
    assert relerr(S[:nb], So[:nb]) < 1e-9
    assert pod_np.subspace_angle(Uo, U)[0] < 1e-8


def test_pod_projections_vs_oracle():
    from sofacontrol_b200.mor import pod
    from oracle.pod_np import PODNP
    rng = np.random.default_rng(0)
    nf, r = 300, 12
    U, _ = np.linalg.qr(rng.normal(size=(nf, r)))
    info = {'U': U, 'q_ref': rng.normal(size=nf), 'v_ref': rng.normal(size=nf), 'type': 'POD'}
    g, o = pod.POD(info), PODNP(info)
    qf, q = rng.normal(size=nf), rng.normal(size=r)
    xf, x = rng.normal(size=2 * nf), rng.normal(size=2 * r)
    M = rng.normal(size=(nf, nf))
    assert relerr(g.compute_RO_state(qf=qf), o.compute_RO_state(qf=qf)) < 1e-12
    assert relerr(g.compute_RO_state(vf=qf), o.compute_RO_state(vf=qf)) < 1e-12
    assert relerr(g.compute_RO_state(xf=xf), o.compute_RO_state(xf=xf)) < 1e-12
    assert relerr(g.compute_FO_state(q=q), o.compute_FO_state(q=q)) < 1e-12
    assert relerr(g.compute_FO_state(x=x), o.compute_FO_state(x=x)) < 1e-12
    assert relerr(g.compute_RO_matrix(M), o.compute_RO_matrix(M)) < 1e-12
    assert relerr(g.compute_RO_matrix(M, left=True), o.compute_RO_matrix(M, left=True)) < 1e-12
    assert relerr(g.compute_RO_matrix(M, right=True), o.compute_RO_matrix(M, right=True)) < 1e-12
    assert np.array_equal(g.V, o.V) and g.rom_dim == r


def test_row_sharded_pod_two_gpus_nccl():
    """Row-sharded POD over 2 GPUs: per-rank DMMA Gram + ONE NCCL all-reduce + sharded back-projection vs numpy SVD
    (tools/pod_sharded_check.py under torchrun).  Skipped on a single-GPU box."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(repo, "tools", "pod_sharded_check.py")],
                       capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"] and out["world"] == 2 and out["subspace_angle"] < 1e-8


def test_run_pod_and_load_pod_roundtrip(tmp_path):
    """run_POD / load_POD (pod.py:93-141) on a snapshot pickle with the reference's schema."""
    from sofacontrol_b200.mor import pod
    from sofacontrol_b200 import utils as scutils
    import sofacontrol_b200.synth as synth
    from oracle import pod_np
    X, _, _ = synth.pod_snapshots(300, 60, seed=7)          # (nf, ns)
    snaps = {'q': list(X.T + 1.0), 'v': list(X.T), 'v+': list(2 * X.T)}
    f_in, f_out = str(tmp_path / "snap.pkl"), str(tmp_path / "out" / "pod.pkl")
    scutils.save_data(f_in, snaps)
    cfg = pod.pod_config()
    cfg.pod_tolerance = 1e-4
    res = pod.run_POD(f_in, f_out, cfg)
    _, Uo, nbo, So = pod_np.compute_POD(X, 1e-4)
    assert res['POD_info']['U'].shape == Uo.shape and res['config']['pod_type'] == 'v'
    assert pod_np.subspace_angle(Uo, res['POD_info']['U'])[0] < 1e-8
    rom = pod.load_POD(f_out)
    assert rom.rom_dim == nbo and np.array_equal(rom.q_ref, snaps['q'][0])
    with pytest.raises(RuntimeError):
        pod.load_POD(str(tmp_path / "missing.pkl"))
