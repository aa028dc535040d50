"""The FP32 screening bounds of the two-stage nearest-point search -- the rollout kernel's centred dot-product form
(csrc/tpwl_screen.cu) and the iLQR forward pass's difference form (csrc/ilqr_fwd_tpwl.cuh) --, emulated
in numpy float32: for every state the set {p : a^_p <= thr2} must contain the float64 argmin of the weighted distance
(tpwl.py:160-168) -- also with huge / tiny coordinate scales, states sitting on stored points, duplicated points and
points that differ by less than float32 can resolve.  CPU-only property test of the bound the CUDA kernels implement."""
import numpy as np
import pytest

U24 = 2.0 ** -24


def candidates(bank64, x64, w):
    """Mirror of the kernel arithmetic (csrc/tpwl_screen.cu): bank and state centred on the bank mean in float64, then
    float32; dot products by a float32 FMA chain; a^ = (|b|^2 + |c|^2) - 2 b.c; threshold min a^ + 2 E2 in float64,
    rounded up to float32.  (w > 0 scales every distance alike and does not enter the bound.)"""
    r = bank64.shape[1]
    mu = bank64.sum(0) / bank64.shape[0]
    bc, xc = bank64 - mu[None, :], x64 - mu
    bank32, x32 = bc.astype(np.float32), xc.astype(np.float32)
    nb = (bc ** 2).sum(1).astype(np.float32)
    nc = np.float32((xc ** 2).sum())
    dot = np.zeros(bank64.shape[0], dtype=np.float32)
    for j in range(r):                                     # fused multiply-add: exact product, one rounding
        dot = (dot.astype(np.float64) + bank32[:, j].astype(np.float64) * np.float64(x32[j])).astype(np.float32)
    s = (nb + nc).astype(np.float32)
    a = (s.astype(np.float64) - 2.0 * dot.astype(np.float64)).astype(np.float32)
    bank_norm = np.sqrt((bc ** 2).sum(1)).max() * (1.0 + 1e-6)
    xnorm = np.sqrt((xc ** 2).sum()) * (1.0 + 1e-9)
    c2 = (0.5 * r + 8.0) * U24 * 1.001 + 1e-12
    sn = bank_norm + xnorm
    E2 = c2 * sn * sn + 1e-36
    am = float(a.min())
    T2 = am + 2.0 * E2 + 1e-6 * (abs(am) + 2.0 * E2)
    if not (sn < 1e17 and T2 < 3.0e38):
        return np.arange(bank64.shape[0])                  # the kernel falls back to the full float64 search
    thr2 = np.nextafter(np.float32(T2), np.float32(np.inf)) if np.float32(T2) < T2 else np.float32(T2)
    return np.nonzero(a <= thr2)[0]


def candidates_diff(bank64, x64, w):
    """Mirror of the iLQR forward pass's screen (csrc/ilqr_fwd_tpwl.cuh), the difference form: float32 bank / state, squared distances accumulated in float32, threshold in
    float64 rounded up to float32."""
    bank32 = bank64.astype(np.float32)
    x32 = x64.astype(np.float32)
    d = bank32 - x32[None, :]
    a = np.zeros(bank64.shape[0], dtype=np.float32)
    for j in range(bank64.shape[1]):                       # sequential float32 accumulation (the kernel uses FMA;
        a = (a + d[:, j] * d[:, j]).astype(np.float32)     # un-fused rounds more often: covered by the same bound)
    bank_norm = w * np.sqrt((bank64 ** 2).sum(1)).max() * (1.0 + 1e-6)
    xnorm = 1.001 * float(np.sqrt(np.float32((x32.astype(np.float32) ** 2).sum(dtype=np.float32))))
    c1 = (0.5 * bank64.shape[1] + 16.0) * U24 + 1e-14      # r/2 + 16 ulp: 34 at r = 36, grows with the sum length
    c0 = U24 * 2.0 * (bank_norm + w * xnorm)
    U = (1.0 + c1) * w * np.sqrt(float(a.min())) + c0
    T = (U + c0) / (w * (1.0 - c1))
    T2 = T * T * (1.0 + 1e-6)
    thr2 = np.nextafter(np.float32(T2), np.float32(np.inf)) if np.float32(T2) < T2 else np.float32(T2)
    return np.nonzero(a <= thr2)[0]


FORMS = {"dot": candidates, "diff": candidates_diff}


@pytest.mark.parametrize("form", ["dot", "diff"])
@pytest.mark.parametrize("r", [36, 5, 64, 128])
@pytest.mark.parametrize("scale", [1e-6, 1e-2, 1.0, 37.0, 1e4, 1e8])
@pytest.mark.parametrize("w", [1.0, 0.3, 250.0])
def test_argmin_is_always_a_candidate(scale, w, r, form):
    candidates = FORMS[form]
    rng = np.random.default_rng(int(scale * 7) % 1000 + int(w * 10) + r)
    P = 300
    for trial in range(40):
        bank = rng.normal(0, scale, size=(P, r)) + rng.normal(0, 3 * scale, size=(1, r)) * (trial % 3)
        kind = trial % 5
        x = rng.normal(0, scale, size=r)
        if kind == 1:
            x = bank[rng.integers(P)].copy()                               # exactly on a stored point
        elif kind == 2:
            x = bank[rng.integers(P)] * (1.0 + 1e-9 * rng.normal(size=r))  # next to one, far inside float32 resolution
        elif kind == 3:
            i, j = rng.integers(P, size=2)
            bank[j] = bank[i]                                              # duplicated point
            x = bank[i] + 1e-3 * scale * rng.normal(size=r)
        elif kind == 4:
            i, j = rng.integers(P, size=2)
            bank[j] = bank[i] * (1.0 + 1e-13)                              # differ below float32 resolution
            x = 0.5 * (bank[i] + bank[j])
        d64 = w * np.linalg.norm(bank - x, axis=1)
        best = int(np.argmin(d64))
        cand = candidates(bank, x, w)
        assert best in cand, (scale, w, trial, kind)
        # every exact tie of the minimum must be a candidate too (first occurrence is decided by the FP64 rescoring)
        assert set(np.nonzero(d64 == d64[best])[0]).issubset(set(cand))


@pytest.mark.parametrize("form", ["dot", "diff"])
def test_screen_is_selective_on_a_diamond_like_bank(form):
    candidates = FORMS[form]
    """On a bank like the bench's (points spread 5, 36 dims) the screen leaves one or two candidates."""
    rng = np.random.default_rng(0)
    bank = rng.normal(0, 5.0, size=(1000, 36))
    counts = [len(candidates(bank, rng.normal(0, 5.0, size=36), 1.0)) for _ in range(50)]
    assert max(counts) <= 3 and np.mean(counts) < 1.5


@pytest.mark.parametrize("form", ["dot", "diff"])
@pytest.mark.parametrize("offset", [1e2, 1e4, 1e6])
def test_large_common_offset(offset, form):
    candidates = FORMS[form]
    """Coordinates dominated by a common offset: float32 rounding of the operands is then comparable to the gaps
    between the distances -- the norm term of the bound has to absorb it (more candidates, never a lost minimum)."""
    rng = np.random.default_rng(int(offset) % 97)
    for trial in range(60):
        bank = offset + rng.normal(0, 1.0, size=(200, 36))
        x = offset + rng.normal(0, 1.0, size=36)
        d64 = np.linalg.norm(bank - x, axis=1)
        assert int(np.argmin(d64)) in candidates(bank, x, 1.0)
