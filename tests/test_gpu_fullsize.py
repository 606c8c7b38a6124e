"""BASELINE.json's full sizes (n = 4096, d = 16 likelihood batch; n = 2000, d = 8 posterior) checked
through size-independent properties -- the oracle needs a minute per evaluation there, so these are
the checks that scale: factor * inverse = I, log-det consistency, gradient vs finite differences of
the value, permutation invariance, the MUCM output-scale law, interpolation at the training points,
grid-index vs explicit-point prediction, implausibility counts vs a direct evaluation."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _synth(n, d, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    w = rng.normal(size=d)
    return X, np.sin(X @ w) + 0.1 * (X ** 2).sum(1)


@pytest.fixture(scope="module")
def dev():
    from gp_emu_uqsa_b200 import _lib
    d = _lib.Device(0)
    yield d
    d.close()


def test_factor_times_inverse_is_identity_n4096(dev):
    from gp_emu_uqsa_b200 import _lib
    n, d = 4096, 16
    X, y = _synth(n, d)
    dev.set_training(X, y, np.column_stack([np.ones(n), X]))
    A = dev.cov_build(np.full(d, 0.6), 1e-4, 0, True, 1.0)
    Lf, Li = np.empty_like(A), np.empty_like(A)
    ld, st = np.empty(1), np.zeros(1, dtype=np.int32)
    dev._ck(dev.L.gpe_potrf(dev.h, _lib._ptr(A), n, 1, _lib._ptr(Lf), _lib._ptr(Li), _lib._ptr(ld), _lib._ptr(st)))
    assert st[0] == 0
    assert np.allclose(np.triu(Lf, 1), 0) and np.allclose(np.triu(Li, 1), 0)
    R = Li @ Lf - np.eye(n)
    assert np.abs(R).max() < 1e-9, np.abs(R).max()
    assert abs(ld[0] - 2.0 * np.log(np.diag(Lf)).sum()) <= 1e-11 * abs(ld[0])
    resid = Lf @ Lf.T - A
    assert np.abs(resid).max() <= 1e-13 * n


def test_llh_gradient_permutation_and_scale_laws_n4096(dev):
    n, d = 4096, 16
    X, y = _synth(n, d)
    H = np.column_stack([np.ones(n), X])
    rng = np.random.default_rng(5)
    hp = np.column_stack([0.4 + 0.5 * rng.random((2, d)), 0.8 + 0.4 * rng.random(2)])
    theta = 2 * np.log(hp)
    eps = 1e-4
    dirs = rng.normal(size=(2, d + 1))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    batch = [theta[0], theta[1]]
    for v in dirs:                                          # central differences around theta[0]
        batch += [theta[0] + eps * v, theta[0] - eps * v]
    dev.set_training(X, y, H)
    llh, grad, sig, st = dev.llh_grad_batch(np.array(batch), 0, fixed_nugget=1e-4)
    assert (st == 0).all()
    for k, v in enumerate(dirs):
        fd = (llh[2 + 2 * k] - llh[3 + 2 * k]) / (2 * eps)
        assert abs(fd - grad[0] @ v) <= 2e-6 * max(1.0, np.abs(grad[0]).max()), (fd, grad[0] @ v)
    # same data in another order: the likelihood is a function of the set of points
    perm = rng.permutation(n)
    dev.set_training(X[perm], y[perm], H[perm])
    llh_p, grad_p, _, st = dev.llh_grad_batch(theta, 0, fixed_nugget=1e-4)
    assert (st == 0).all()
    assert np.allclose(llh_p, llh[:2], rtol=1e-10, atol=0)
    assert np.allclose(grad_p, grad[:2], rtol=0, atol=1e-8 * np.abs(grad[:2]).max())
    # MUCM: y -> c y  =>  sigma_hat -> c sigma_hat, LLH -> LLH + (n - q) ln c
    th_m = theta[:, :d]
    dev.set_training(X, y, H)
    l1, g1, s1, st1 = dev.llh_grad_batch(th_m, 1, fixed_nugget=1e-4)
    dev.set_training(X, 3.0 * y, H)
    l3, g3, s3, st3 = dev.llh_grad_batch(th_m, 1, fixed_nugget=1e-4)
    assert (st1 == 0).all() and (st3 == 0).all()
    assert np.allclose(s3, 3.0 * s1, rtol=1e-10)
    assert np.allclose(l3, l1 + (n - (d + 1)) * np.log(3.0), rtol=1e-10)
    assert np.allclose(g3, 9.0 * g1, rtol=1e-7, atol=1e-9 * np.abs(g3).max())      # reference quirk: grad carries sigma_hat^2


def test_posterior_properties_n2000_and_grid_consistency(dev):
    n, d = 2000, 8
    X, y = _synth(n, d, seed=2)
    dev.set_training(X, y, np.column_stack([np.ones(n), X]))
    dev.set_basis(list(range(d)), [1] * d)
    beta, _, st = dev.fit_state(np.full(d, 0.5), 1e-4, 1.0, 0)
    assert st == 0
    mean, var = dev.predict(X[:700])                        # at training points: interpolation up to the nugget
    assert np.abs(mean - y[:700]).max() < 2e-3 and var.min() > -1e-12 and var.max() < 5e-3
    # flat-index grid == the same points given explicitly (ragged count, crossing chunk boundaries)
    levels = np.full(d, 10, dtype=np.int32)
    start, count = 123457, 70001
    gm, gv = dev.predict_grid(levels, np.zeros(d), np.ones(d), start, count)
    idx = np.arange(start, start + count)
    P = np.empty((count, d))
    for k in range(d - 1, -1, -1):
        P[:, k] = 0.0 + (idx % 10 + 0.5) * ((1.0 - 0.0) / 10.0)      # the device's formula: lo + (digit + 0.5) * step
        idx = idx // 10
    em, ev = dev.predict(P)
    # two evaluation orders of the same kernel entries: the grid path multiplies tabulated one-dimensional factors
    # (prod_k exp(-Delta_k^2)), the explicit path takes exp(-sum_k Delta_k^2); they differ in the last ulps of C, which
    # the variance (a difference of O(1) terms that is ~1e-3 here) magnifies.  Both are held to 1e-8 against the oracle
    # below and in the config-4 test.
    assert np.allclose(gm, em, rtol=1e-11, atol=1e-13) and np.allclose(gv, ev, rtol=1e-9, atol=1e-14)
    assert ev.min() > 0
    # implausibility reductions vs direct evaluation of history_match.py:121-136 on the same arrays
    z, ve, cm = float(np.median(y)), 1e-2, 3.0
    Imax, keep, cnt, cmin, ccnt = dev.implausibility(gm[None, :70000], gv[None, :70000], [z], [ve], cm, maxno=1, ncell=7)
    I = np.sqrt((gm[:70000] - z) ** 2 / (gv[:70000] + ve))
    assert np.array_equal(Imax[:, 0], I) or np.allclose(Imax[:, 0], I, rtol=1e-15)
    assert np.array_equal(keep.astype(bool), I < cm) and cnt[0] == (I < cm).sum()
    assert np.allclose(cmin[:, 0], I.reshape(7, -1).min(1), rtol=1e-15)
    assert np.array_equal(ccnt[:, 0], (I.reshape(7, -1) < cm).sum(1))


def test_config4_subsample_matches_oracle_and_index_sets_are_identical(dev):
    """Config 4: n = 2000, d = 8 emulators, points of the 10^8 grid.  A random subsample of
    grid indices is predicted on the GPU and by the oracle (chunked g.posterior arithmetic); mean / variance
    agree to 1e-8, the non-implausible index sets of two emulators are identical, and no point sits within the
    guard band |I - cm| < 1e-9 where a last-bit difference could flip the decision."""
    from gp_emu_uqsa_b200 import _lib
    from oracle import gp_oracle as O
    n, d, m = 2000, 8, 3000
    X, y = _synth(n, d)
    y2 = np.cos(X @ np.random.default_rng(1).normal(size=d))
    H = np.column_stack([np.ones(n), X])
    delta = np.full(d, 0.5)
    rng = np.random.default_rng(4)
    idx = np.sort(rng.choice(10 ** 8, size=m, replace=False))
    P = np.empty((m, d))
    t = idx.copy()
    for k in range(d - 1, -1, -1):
        P[:, k] = 0.0 + (t % 10 + 0.5) * ((1.0 - 0.0) / 10.0)
        t = t // 10
    Hs = np.column_stack([np.ones(m), P])
    means, variances = [], []
    A = O.make_A(X, delta, 1e-4, 0)
    for yy in (y, y2):
        dev.set_training(X, yy, H)
        dev.set_basis(list(range(d)), [1] * d)
        beta, _, st = dev.fit_state(delta, 1e-4, 1.0, 0)
        assert st == 0
        gm, gv = dev.predict(P)
        om, ov = O.posterior_diag_chunked(P, Hs, X, yy, H, A, beta, 1.0, delta, 1e-4, 0, chunk=1000)
        assert np.allclose(beta, O.optimalbeta(A, H, yy), rtol=1e-8, atol=1e-10)
        assert np.allclose(gm, om, rtol=1e-8, atol=1e-10)
        assert np.allclose(gv, ov, rtol=1e-8, atol=1e-9 * ov.max())
        # a single flat-index prediction agrees with the explicit-point one
        g1, v1 = dev.predict_grid(np.full(d, 10, dtype=np.int32), np.zeros(d), np.ones(d), int(idx[7]), 1)
        assert abs(g1[0] - gm[7]) <= 1e-11 * abs(gm[7]) and abs(v1[0] - gv[7]) <= 1e-9 * abs(gv[7])
        # the grid path itself (tabulated separable factors) against the oracle: a run of consecutive flat indices that
        # crosses a boundary of the slow digits (... 999 -> ... 000) and is not a multiple of the tile
        g0 = 37 * 1000 - 333
        Pg = np.empty((777, d))
        t2 = np.arange(g0, g0 + 777)
        for k in range(d - 1, -1, -1):
            Pg[:, k] = 0.0 + (t2 % 10 + 0.5) * ((1.0 - 0.0) / 10.0)
            t2 = t2 // 10
        gg, gvv = dev.predict_grid(np.full(d, 10, dtype=np.int32), np.zeros(d), np.ones(d), g0, 777)
        omg, ovg = O.posterior_diag_chunked(Pg, np.column_stack([np.ones(777), Pg]), X, yy, H, A, beta, 1.0, delta, 1e-4, 0, chunk=1000)
        assert np.allclose(gg, omg, rtol=1e-8, atol=1e-10)
        assert np.allclose(gvv, ovg, rtol=1e-8, atol=1e-9 * ovg.max())
        means.append(gm); variances.append(gv)
    zs, ve, cm = [float(np.median(y)), float(np.median(y2))], [1e-2, 1e-2], 3.0
    Imax, keep, cnt, _, _ = dev.implausibility(np.array(means), np.array(variances), zs, ve, cm, maxno=1)
    om1, ov1 = O.posterior_diag_chunked(P, Hs, X, y, H, A, O.optimalbeta(A, H, y), 1.0, delta, 1e-4, 0, chunk=1000)
    om2, ov2 = O.posterior_diag_chunked(P, Hs, X, y2, H, A, O.optimalbeta(A, H, y2), 1.0, delta, 1e-4, 0, chunk=1000)
    Iref, kref, _ = O.implausibility(np.array([om1, om2]), np.array([ov1, ov2]), zs, ve, cm, 1)
    assert np.abs(np.asarray(Iref)[:, -1] - cm).min() > 1e-9          # guard band is empty on this sample
    assert np.array_equal(keep.astype(bool), np.asarray(kref)) and int(cnt[0]) == int(np.asarray(kref).sum())
    assert 0 < int(cnt[0]) < m


def test_large_ragged_n10000_gradient_matches_finite_differences(dev):
    """Beyond BASELINE's sizes: n = 10000 (pads to 10112 = 79 tiles, odd tile count at several recursion levels,
    2.4 GB of workspace per item).  The gradient must agree with central differences of the value, and the
    value with the factor's own log-determinant route on a second mode (free nugget)."""
    n, d = 10000, 6
    X, y = _synth(n, d, seed=3)
    H = np.column_stack([np.ones(n), X])
    rng = np.random.default_rng(11)
    hp = np.r_[0.5 + 0.4 * rng.random(d), 1e-3, 0.9]            # delta, nugget, sigma
    theta = np.r_[2 * np.log(hp[:d]), 2 * np.log(hp[d]), 2 * np.log(hp[d + 1])]
    eps = 1e-4
    v = rng.normal(size=theta.size)
    v /= np.linalg.norm(v)
    dev.set_training(X, y, H)
    llh, grad, sig, st = dev.llh_grad_batch(np.array([theta, theta + eps * v, theta - eps * v]), 4, fixed_nugget=0.0)
    assert (st == 0).all()
    fd = (llh[1] - llh[2]) / (2 * eps)
    assert abs(fd - grad[0] @ v) <= 5e-6 * max(1.0, np.abs(grad[0]).max()), (fd, grad[0] @ v)
    assert np.isfinite(grad).all() and np.isfinite(sig).all()


def test_potrf_leaves_training_set_and_fit_state_alone(dev):
    """gpe_fit_state -> gpe_potrf (host and device buffers, another size) -> gpe_predict on ONE handle returns the
    pre-potrf prediction bit for bit: the dense Cholesky works in its own temporaries (noise_fit.py:130-150 and
    Posterior.mahalanobis_distance call it between predictions)."""
    from gp_emu_uqsa_b200 import _lib
    n, d = 700, 5
    X, y = _synth(n, d, seed=9)
    H = np.column_stack([np.ones(n), X])
    dev.set_training(X, y, H)
    dev.set_basis(list(range(d)), [1] * d)
    beta, _, st = dev.fit_state(np.full(d, 0.4), 1e-4, 1.1, 0)
    assert st == 0
    P = np.random.default_rng(3).random((300, d))
    m0, v0 = dev.predict(P)
    l0, g0, _, st0 = dev.llh_grad_batch(2 * np.log(np.r_[np.full(d, 0.4), 1.1])[None], 0, fixed_nugget=1e-4)
    # a 300 x 300 posterior covariance, host buffers
    _, V = dev.predict_fullcov(P)
    Lf = dev.cholesky(V)
    assert np.allclose(Lf @ Lf.T, V, rtol=0, atol=1e-12 * np.abs(V).max())
    assert np.allclose(Lf, np.linalg.cholesky(V), rtol=1e-6, atol=1e-9 * np.sqrt(np.abs(V).max()))
    # device buffers, a batch of two, size not a multiple of the tile
    Ad = torch.tensor(np.stack([V[:200, :200], 2.0 * V[:200, :200]]), device="cuda")
    Ld = torch.empty_like(Ad); Lid = torch.empty_like(Ad)
    ld = torch.empty(2, dtype=torch.float64, device="cuda"); sd = torch.zeros(2, dtype=torch.int32, device="cuda")
    dev._ck(dev.L.gpe_potrf(dev.h, Ad.data_ptr(), 200, 2, Ld.data_ptr(), Lid.data_ptr(), ld.data_ptr(), sd.data_ptr()))
    torch.cuda.synchronize()
    assert int(sd.abs().sum()) == 0
    assert torch.allclose(Ld @ Ld.transpose(1, 2), Ad, rtol=0, atol=1e-12 * float(Ad.abs().max()))
    assert torch.allclose(Lid @ Ld, torch.eye(200, dtype=torch.float64, device="cuda").expand(2, 200, 200), rtol=0, atol=1e-8)
    assert torch.allclose(ld, 2.0 * torch.log(torch.diagonal(Ld, dim1=1, dim2=2)).sum(1), rtol=1e-11)
    # the handle still serves the same training set and fit
    assert dev.n == n
    m1, v1 = dev.predict(P)
    assert np.array_equal(m0, m1) and np.array_equal(v0, v1)
    l1, g1, _, st1 = dev.llh_grad_batch(2 * np.log(np.r_[np.full(d, 0.4), 1.1])[None], 0, fixed_nugget=1e-4)
    assert np.array_equal(l0, l1) and np.array_equal(g0, g1)


def test_implausibility_cells_of_a_shard_that_starts_and_ends_inside_a_cell(dev):
    """Point-sharded history matching: cell statistics of arbitrary index ranges combine (min / sum) to the
    statistics of the whole set, cells straddling shard boundaries included."""
    rng = np.random.default_rng(12)
    m, cell_pts, cm = 10000, 700, 1.2
    mean = rng.normal(size=(2, m)); var = 0.1 + rng.random((2, m))
    zs, ve = [0.1, -0.2], [0.05, 0.02]
    ncell = (m + cell_pts - 1) // cell_pts
    I = np.sqrt((mean - np.array(zs)[:, None]) ** 2 / (var + np.array(ve)[:, None])).max(0)
    ref_min = np.array([I[c * cell_pts:(c + 1) * cell_pts].min() for c in range(ncell)])
    ref_cnt = np.array([(I[c * cell_pts:(c + 1) * cell_pts] < cm).sum() for c in range(ncell)])
    tot_min, tot_cnt, kept = np.full(ncell, np.inf), np.zeros(ncell, dtype=np.uint64), 0
    for lo, hi in ((0, 1234), (1234, 1234), (1234, 6001), (6001, m)):          # one empty shard
        _, keep, cnt, cmin, ccnt = dev.implausibility(mean[:, lo:hi], var[:, lo:hi], zs, ve, cm, maxno=1, cell_pts=cell_pts,
                                                      first_index=lo, want_imax=False)
        kept += int(cnt[0])
        if hi > lo:
            c0 = lo // cell_pts
            tot_min[c0:c0 + cmin.shape[0]] = np.minimum(tot_min[c0:c0 + cmin.shape[0]], cmin[:, 0])
            tot_cnt[c0:c0 + ccnt.shape[0]] += ccnt[:, 0]
            assert np.array_equal(keep.astype(bool), I[lo:hi] < cm)
    assert np.array_equal(tot_min, ref_min) and np.array_equal(tot_cnt, ref_cnt.astype(np.uint64)) and kept == int((I < cm).sum())
