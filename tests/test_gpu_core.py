"""GPU parity: DMMA GEMM family, recursive Cholesky+inverse, covariance build, llh+grad
-- all through the C-ABI (gp_emu_uqsa_b200._lib), checked against the oracle / goldens."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import gp_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    from gp_emu_uqsa_b200 import _lib
    d = _lib.Device(0)
    yield d
    d.close()


def _tril_mask(n, device):
    return torch.tril(torch.ones(n, n, dtype=torch.float64, device=device))


@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 128, 3), (256, 384, 256, 2), (1024, 1024, 512, 2), (256, 32, 256, 2)])
@pytest.mark.parametrize("layout", [0, 1, 2])
def test_gemm_layouts(dev, M, N, K, batch, layout):
    from gp_emu_uqsa_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(M + N + K + layout)
    A = torch.randn(batch, M, K, dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn(batch, N, K, dtype=torch.float64, device="cuda", generator=g)
    C0 = torch.randn(batch, M, N, dtype=torch.float64, device="cuda", generator=g)
    ref = torch.matmul(A, B.transpose(1, 2))
    # storage per layout: A_KC -> [M,K], else [K,M];  B_KC -> [N,K], else [K,N]
    As = A.contiguous() if layout in (0, 1) else A.transpose(1, 2).contiguous()
    Bs = B.contiguous() if layout == 0 else B.transpose(1, 2).contiguous()
    lda = K if layout in (0, 1) else M
    ldb = K if layout == 0 else N
    Cm = C0.clone()
    dev.dbg_gemm(As, Bs, Cm, M, N, K, lda, ldb, N, sA=M * K, sB=N * K, sC=M * N, alpha=-0.5, accumulate=1, batch=batch, layout=layout)
    torch.cuda.synchronize()
    want = C0 - 0.5 * ref
    assert torch.allclose(Cm, want, rtol=1e-12, atol=1e-11)
    Cm2 = torch.zeros_like(C0)
    dev.dbg_gemm(As, Bs, Cm2, M, N, K, lda, ldb, N, sA=M * K, sB=N * K, sC=M * N, alpha=1.0, accumulate=0, batch=batch, layout=layout)
    torch.cuda.synchronize()
    assert torch.allclose(Cm2, ref, rtol=1e-12, atol=1e-11)


@pytest.mark.parametrize("n,batch", [(512, 2), (1024, 8), (1536, 4)])   # 64x64 tiles / warp-specialised 128x128 tiles
def test_gemm_triangular_kmodes(dev, n, batch):
    from gp_emu_uqsa_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(7)
    T = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g) * _tril_mask(n, "cuda")
    F = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g)
    sz = n * n
    # NT, B lower [n][k]: C = F T^T, k <= j
    Cm = torch.empty_like(F)
    dev.dbg_gemm(F, T, Cm, n, n, n, n, n, n, sz, sz, sz, kmode=_lib.KM_LE_J, batch=batch, layout=0)
    assert torch.allclose(Cm, F @ T.transpose(1, 2), rtol=1e-12, atol=1e-11)
    # NN, B lower stored [k][n]: C = F T, k >= j
    dev.dbg_gemm(F, T, Cm, n, n, n, n, n, n, sz, sz, sz, kmode=_lib.KM_GE_J, batch=batch, layout=1)
    assert torch.allclose(Cm, F @ T, rtol=1e-12, atol=1e-11)
    # NN, A lower: C = T F, k <= i
    dev.dbg_gemm(T, F, Cm, n, n, n, n, n, n, sz, sz, sz, kmode=_lib.KM_LE_I, batch=batch, layout=1)
    assert torch.allclose(Cm, T @ F, rtol=1e-12, atol=1e-11)
    # TN, lower-only output: C = T^T T (LAUUM), k >= i
    Cm.zero_()
    dev.dbg_gemm(T, T, Cm, n, n, n, n, n, n, sz, sz, sz, kmode=_lib.KM_GE_I, lower=1, batch=batch, layout=2)
    want = T.transpose(1, 2) @ T
    m = _tril_mask(n, "cuda")
    assert torch.allclose(Cm * m, want * m, rtol=1e-12, atol=1e-11)
    # TN, A lower stored [k][m] against a dense panel: C = T^T F, k >= i  (back substitution panels)
    dev.dbg_gemm(T, F, Cm, n, n, n, n, n, n, sz, sz, sz, kmode=_lib.KM_GE_I, batch=batch, layout=2)
    assert torch.allclose(Cm, T.transpose(1, 2) @ F, rtol=1e-12, atol=1e-11)
    # NT, lower-only accumulate: C -= F F^T (SYRK); entries above the diagonal of C are left alone
    C0 = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g)
    Cm = C0.clone()
    torch.cuda.synchronize()        # the handle launches on its own (non-blocking) stream
    dev.dbg_gemm(F, F, Cm, n, n, n, n, n, n, sz, sz, sz, alpha=-1.0, accumulate=1, kmode=_lib.KM_FULL, lower=1, batch=batch, layout=0)
    want = C0 - F @ F.transpose(1, 2)
    assert torch.allclose(Cm * m, want * m, rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("n,batch", [(1, 1), (7, 3), (60, 2), (128, 3), (129, 2), (200, 2), (640, 2), (1000, 1), (2000, 2)])
def test_potrf_inv(dev, n, batch):
    rng = np.random.default_rng(n)
    As = []
    for b in range(batch):
        X = rng.random((n, 4))
        A = O.make_A(X, np.full(4, 0.4 + 0.2 * b), 1e-4, 0)
        As.append(A)
    As = np.array(As)
    Li, logdet, st = dev.dbg_potrf_inv(As)
    assert (st == 0).all()
    for b in range(batch):
        L = np.linalg.cholesky(As[b])
        Lref = np.linalg.inv(L)
        assert np.allclose(np.triu(Li[b], 1), 0.0)
        err = np.abs(Li[b] @ L - np.eye(n)).max()
        assert err < 1e-9, err
        assert np.allclose(Li[b], Lref, rtol=1e-7, atol=1e-7 * np.abs(Lref).max())
        assert abs(logdet[b] - 2 * np.log(np.diag(L)).sum()) < 1e-9 * max(1.0, abs(logdet[b]))


def test_potrf_reports_non_pd(dev):
    n = 200
    rng = np.random.default_rng(0)
    M = rng.random((n, n))
    A = M @ M.T + n * np.eye(n)
    A[150, 150] = -1.0
    Li, logdet, st = dev.dbg_potrf_inv(A)
    assert st[0] == 151          # LAPACK-style 1-based index of the first non-positive pivot
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.cholesky(A)


@pytest.mark.parametrize("kind,predict", [(0, True), (0, False), (1, True), (1, False)])
def test_cov_build(dev, kind, predict):
    rng = np.random.default_rng(3)
    n, d = 300, 5
    X = rng.random((n, d))
    y = rng.random(n)
    H = O.make_H_linear(X)
    r = 0.01 + 0.02 * rng.random(n)
    dev.set_training(X, y, H, r)
    delta = 0.2 + rng.random(d)
    A = dev.cov_build(delta, 0.01, kind=kind, predict=predict, s2=0.7)
    ref = O.make_A(X, delta, 0.01, kind, r, 0.7, predict)
    assert np.allclose(A, ref, rtol=1e-14, atol=1e-15)
    assert np.array_equal(A, A.T)


MODES = [("mucm_k_fixT", 1), ("mucm_k_fixF", 1 | 4), ("gp4ml_k_fixT", 0), ("gp4ml_k_fixF", 4),
         ("gp4ml_alt_fixT", 2), ("gp4ml_alt_fixF", 2 | 4)]


@pytest.mark.parametrize("fname", ["llh_n60_d2.npz", "llh_n200_d4.npz", "llh_n500_d8.npz"])
def test_llh_grad_vs_reference_golden(dev, golden_dir, fname):
    """Tolerances: rel 1e-10 on llh (north_star), 1e-9 rel on the gradient."""
    G = np.load(os.path.join(golden_dir, fname))
    X, y, H = G["X"], G["y"], G["H"]
    for tag, mode in MODES:
        r = G[tag + "_r"] if tag + "_r" in G.files else None
        dev.set_training(X, y, H, r)
        theta = G[tag + "_theta"]
        llh, grad, sig, st = dev.llh_grad_batch(theta, mode, fixed_nugget=float(G["nugget_belief"]))
        assert (st == 0).all(), tag
        assert np.allclose(llh, G[tag + "_llh"], rtol=1e-10, atol=0), (tag, llh, G[tag + "_llh"])
        gscale = np.abs(G[tag + "_grad"]).max(axis=1, keepdims=True)
        assert np.all(np.abs(grad - G[tag + "_grad"]) <= 1e-9 * gscale + 1e-12), (tag, grad, G[tag + "_grad"])
        assert np.allclose(sig, G[tag + "_sigma"], rtol=1e-10)


def test_llh_grad_vs_oracle_n1000(dev):
    """Config-2 shape (n=1000, d=8, q=9), a 6-guess batch, against the oracle restatement."""
    rng = np.random.default_rng(0)
    n, d = 1000, 8
    X = rng.random((n, d)); w = rng.normal(size=d)
    y = np.sin(X @ w) + 0.1 * (X ** 2).sum(1)
    H = O.make_H_linear(X)
    dev.set_training(X, y, H)
    B = 6
    hp = np.column_stack([0.2 + 0.8 * rng.random((B, d)), 0.5 + rng.random(B)])
    theta = O.transform(hp)
    llh, grad, sig, st = dev.llh_grad_batch(theta, 0, fixed_nugget=1e-4)
    assert (st == 0).all()
    for b in range(B):
        ref = O.loglikelihood_gp4ml(theta[b], X, y, H, 0, 1e-4)
        assert abs(llh[b] - ref[0]) <= 1e-10 * abs(ref[0])
        assert np.all(np.abs(grad[b] - ref[1]) <= 1e-9 * np.abs(ref[1]).max())
    llh2, grad2, sig2, st2 = dev.llh_grad_batch(theta[:, :d], 1, fixed_nugget=1e-4)
    for b in range(0, B, 3):
        ref = O.loglikelihood_mucm(theta[b, :d], X, y, H, 0, 1e-4)
        assert abs(llh2[b] - ref[0]) <= 1e-10 * abs(ref[0])
        assert np.all(np.abs(grad2[b] - ref[1]) <= 1e-9 * np.abs(ref[1]).max())
        assert abs(sig2[b] - ref[2]) <= 1e-10 * ref[2]


def test_llh_non_pd_item_is_flagged_not_fatal(dev):
    rng = np.random.default_rng(1)
    n, d = 150, 2
    X = rng.random((n, d)); X[1] = X[0]          # duplicate point + zero nugget => singular
    y = rng.random(n)
    H = O.make_H_linear(X)
    dev.set_training(X, y, H)
    # nugget is a free parameter: item 0 has nugget 1e-17 (1 - nugget rounds to 1 => the duplicate
    # pair gives an exactly zero pivot), item 1 a healthy 1e-3
    theta = O.transform(np.array([[0.5, 0.5, 1e-17, 1.0], [0.3, 0.3, 1e-3, 1.0]]))
    llh, grad, sig, st = dev.llh_grad_batch(theta, 4, fixed_nugget=0.0)
    assert st[0] != 0 and st[1] == 0
    assert O.loglikelihood_gp4ml(theta[0], X, y, H, 0, 0.0) is None
    ref = O.loglikelihood_gp4ml(theta[1], X, y, H, 0, 0.0)
    assert abs(llh[1] - ref[0]) <= 1e-10 * abs(ref[0])


def test_llh_batch_larger_than_resident_capacity(dev, golden_dir):
    """More guesses than the resident workspace holds (64): the call walks sub-batches; every item must equal
    the value it gets in a small batch (ragged last sub-batch included)."""
    G = np.load(os.path.join(golden_dir, "llh_n60_d2.npz"))
    dev.set_training(G["X"], G["y"], G["H"])
    tag, mode = "gp4ml_k_fixF", 4
    base = G[tag + "_theta"]
    rng = np.random.default_rng(5)
    theta = base[rng.integers(0, len(base), 150)] + 0.05 * rng.normal(size=(150, base.shape[1]))
    nug = float(G["nugget_belief"])
    llh, grad, sig, st = dev.llh_grad_batch(theta, mode, fixed_nugget=nug)
    for lo in range(0, 150, 37):
        l2, g2, s2, t2 = dev.llh_grad_batch(theta[lo:lo + 37], mode, fixed_nugget=nug)
        assert np.array_equal(st[lo:lo + 37], t2)
        ok = t2 == 0
        assert np.array_equal(llh[lo:lo + 37][ok], l2[ok]) and np.array_equal(grad[lo:lo + 37][ok], g2[ok])
        assert np.array_equal(sig[lo:lo + 37][ok], s2[ok])
