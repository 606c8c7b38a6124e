"""GPU parity: posterior mean / variance (K4, K4f), cross-covariance (K1x) and implausibility (K5)
through the C-ABI, against the reference's golden outputs and the oracle."""
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


def _setup(dev, G, tag, kind, d):
    X, y = G["X"], G["y"]
    H = O.make_H_linear(X)
    r = G[tag + "_r"] if tag + "_r" in G.files else None
    dev.set_training(X, y, H, r)
    dev.set_basis(list(range(d)), [1] * d)
    delta, sigma, nugget = G[tag + "_delta"], float(G[tag + "_sigma"]), float(G[tag + "_nugget"])
    beta, sig_mucm, st = dev.fit_state(delta, nugget, sigma, kind)
    assert st == 0
    return X, y, H, r, delta, sigma, nugget, beta


@pytest.mark.parametrize("fname,d", [("post_n200_d4.npz", 4), ("post_n500_d8.npz", 8)])
@pytest.mark.parametrize("tag,kind", [("mucm_k_fixT", 0), ("gp4ml_k_fixT", 0), ("gp4ml_alt_fixT", 1)])
def test_posterior_vs_reference_golden(dev, golden_dir, fname, d, tag, kind):
    """north_star tolerance: rel 1e-8 on posterior mean and variance."""
    G = np.load(os.path.join(golden_dir, fname))
    X, y, H, r, delta, sigma, nugget, beta = _setup(dev, G, tag, kind, d)
    assert np.allclose(beta, G[tag + "_beta"], rtol=1e-9, atol=1e-12)     # optimalbeta
    Xs = G["Xs"]
    mean, var = dev.predict(Xs)                                            # device basis
    vref = np.diag(G[tag + "_var"])
    assert np.allclose(mean, G[tag + "_mean"], rtol=1e-8, atol=1e-10)
    assert np.allclose(var, vref, rtol=1e-8, atol=1e-8 * np.abs(vref).max())
    mean2, var2 = dev.predict(Xs, Hs=O.make_H_linear(Xs))                  # explicit H*
    assert np.allclose(mean2, mean, rtol=1e-13) and np.allclose(var2, var, rtol=1e-12, atol=1e-16)
    mean3, _ = dev.predict(Xs, want_var=False)
    assert np.array_equal(mean3, mean)
    # full covariance (posterior_sample / mahalanobis / noisefit consumers)
    mu, V = dev.predict_fullcov(Xs)
    assert np.allclose(mu, G[tag + "_mean"], rtol=1e-8, atol=1e-10)
    assert np.allclose(V, G[tag + "_var"], rtol=1e-8, atol=1e-8 * np.abs(vref).max())
    # cross covariance
    Cm = dev.cross_cov(delta, nugget, kind, Xs)
    assert np.allclose(Cm, O.cov_covar(X, Xs, delta, nugget, kind), rtol=1e-14, atol=1e-300)


def test_mucm_sigma_and_user_beta(dev, golden_dir):
    G = np.load(os.path.join(golden_dir, "post_n200_d4.npz"))
    tag = "mucm_k_fixT"
    X, y = G["X"], G["y"]
    H = O.make_H_linear(X)
    dev.set_training(X, y, H)
    dev.set_basis([0, 1, 2, 3], [1, 1, 1, 1])
    delta, nugget = G[tag + "_delta"], float(G[tag + "_nugget"])
    A = O.make_A(X, delta, nugget, 0)
    beta, sig, st = dev.fit_state(delta, nugget, 1.0, 0)
    assert abs(sig - O.sigma_analytic_mucm(A, H, y)) < 1e-11 * sig
    user_beta = np.array([0.3, -0.2, 0.1, 0.4, 1.5])
    dev.fit_state(delta, nugget, 0.9, 0, beta=user_beta)
    Xs = G["Xs"][:40]
    mean, var = dev.predict(Xs)
    mref, Vref = O.posterior(Xs, O.make_H_linear(Xs), X, y, H, A, user_beta, 0.9, delta, nugget, 0)
    assert np.allclose(mean, mref, rtol=1e-8, atol=1e-10)
    assert np.allclose(var, np.diag(Vref), rtol=1e-8, atol=1e-8 * np.diag(Vref).max())


def test_grid_prediction_matches_explicit_points_and_ragged_sizes(dev, golden_dir):
    G = np.load(os.path.join(golden_dir, "post_n200_d4.npz"))
    _setup(dev, G, "gp4ml_k_fixT", 0, 4)
    levels = np.array([5, 4, 3, 7])
    lo, hi = np.zeros(4), np.array([1.0, 1.0, 0.5, 2.0])
    total = int(np.prod(levels))
    idx = np.arange(total)
    digs = np.stack(np.unravel_index(idx, levels), axis=1)
    pts = lo + (digs + 0.5) * (hi - lo) / levels
    m_all, v_all = dev.predict(pts)
    for start, count in [(0, total), (17, 131), (400, 20), (1, 1)]:
        mg, vg = dev.predict_grid(levels, lo, hi, start, count)
        # grid path: products of tabulated one-dimensional factors; explicit path: exp of the summed exponent -- equal up
        # to the last ulps of the cross-covariance, which the variance (a small difference of O(1) terms) magnifies
        assert np.allclose(mg, m_all[start:start + count], rtol=1e-11, atol=1e-12)
        assert np.allclose(vg, v_all[start:start + count], rtol=1e-9, atol=1e-13)
    # device-resident in/out, chunked (chunk < m) and not a multiple of 128
    os.environ["GPE_PRED_CHUNK"] = "256"
    try:
        from gp_emu_uqsa_b200 import _lib
        d2 = _lib.Device(0)
        X, y = G["X"], G["y"]
        d2.set_training(X, y, O.make_H_linear(X)); d2.set_basis([0, 1, 2, 3], [1] * 4)
        d2.fit_state(G["gp4ml_k_fixT_delta"], float(G["gp4ml_k_fixT_nugget"]), float(G["gp4ml_k_fixT_sigma"]), 0)
        P = torch.tensor(pts[:389], device="cuda")
        mo = torch.empty(389, dtype=torch.float64, device="cuda"); vo = torch.empty_like(mo)
        d2.predict(P, out=(mo, vo))
        assert np.allclose(mo.cpu().numpy(), m_all[:389], rtol=1e-12, atol=1e-13)
        assert np.allclose(vo.cpu().numpy(), v_all[:389], rtol=1e-10, atol=1e-14)
        d2.close()
    finally:
        del os.environ["GPE_PRED_CHUNK"]


def test_implausibility_vs_reference_golden(dev, golden_dir):
    """Identical non-implausible index sets (north_star) vs history_match.nonimp_data outputs."""
    G = np.load(os.path.join(golden_dir, "hm_n100_d3.npz"))
    pts = G["pts_scaled"]
    means, variances = [], []
    for o in range(2):
        X, y = G["Xtrain%d" % o], G["ytrain%d" % o]
        dev.set_training(X, y, O.make_H_linear(X)); dev.set_basis([0, 1, 2], [1, 1, 1])
        dev.fit_state(G["delta%d" % o], float(G["nugget%d" % o]), float(G["sigma%d" % o]), 0, beta=G["beta%d" % o])
        mu, v = dev.predict(pts)
        assert np.allclose(mu, G["means"][o], rtol=1e-8, atol=1e-10)
        assert np.allclose(v, G["vars"][o], rtol=1e-7, atol=1e-8 * G["vars"][o].max())
        means.append(mu); variances.append(v)
    means, variances = np.array(means), np.array(variances)
    for maxno in (1, 2):
        Imax, keep, count, cmin, ccnt = dev.implausibility(means, variances, G["zs"], G["var_extra"], float(G["cm"]), maxno, ncell=4)
        Iref, kref, cref = O.implausibility(means, variances, G["zs"], G["var_extra"], float(G["cm"]), maxno)
        assert np.allclose(Imax, Iref, rtol=1e-14, atol=0)
        guard = np.abs(Iref[:, 0] - float(G["cm"])) < 1e-9          # guard band for index-set flips
        assert not guard.any()
        assert np.array_equal(keep.astype(bool), kref)
        assert int(keep.sum()) == int(G["count_maxno%d" % maxno])
        assert np.allclose(pts[keep.astype(bool)], G["kept_maxno%d" % maxno], atol=1e-15)
        assert np.array_equal(count, cref.astype(np.uint64))
        cells = Iref.reshape(4, -1, maxno)
        imp_ref, odp_ref = O.implausibility_cells([c for c in cells], float(G["cm"]))
        assert np.allclose(cmin, imp_ref, rtol=1e-14)
        assert np.allclose(ccnt / cells.shape[1], odp_ref)


def test_empty_and_single_unit_calls(dev, golden_dir):
    """Zero units is a successful no-op (an empty rank of a partitioned job); one unit equals the first of many."""
    L = np.load(os.path.join(golden_dir, "llh_n200_d4.npz"))
    tag, mode = "gp4ml_k_fixT", 0
    dev.set_training(L["X"], L["y"], L["H"])
    theta = L[tag + "_theta"]
    p, nug = theta.shape[1], float(L["nugget_belief"])
    llh0, grad0, sig0, st0 = dev.llh_grad_batch(np.empty((0, p)), mode, fixed_nugget=nug)
    assert llh0.shape == (0,) and grad0.shape == (0, p) and st0.shape == (0,)
    lB, gB, _, sB = dev.llh_grad_batch(theta, mode, fixed_nugget=nug)
    l1, g1, _, s1 = dev.llh_grad_batch(theta[:1], mode, fixed_nugget=nug)
    assert (sB == 0).all() and s1.tolist() == [0]
    assert l1[0] == lB[0] and np.array_equal(g1[0], gB[0])      # an item's value does not depend on its batch

    G = np.load(os.path.join(golden_dir, "post_n200_d4.npz"))
    d = 4
    _setup(dev, G, "gp4ml_k_fixT", 0, d)
    Xs = G["Xs"]
    mean0, var0 = dev.predict(np.empty((0, d)))
    assert mean0.shape == (0,) and var0.shape == (0,)
    mean1, var1 = dev.predict(Xs[:1])
    meanm, varm = dev.predict(Xs)
    np.testing.assert_allclose(mean1[0], meanm[0], rtol=1e-12)
    np.testing.assert_allclose(var1[0], varm[0], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(mean1[0], G["gp4ml_k_fixT_mean"][0], rtol=1e-8)
    mg, vg = dev.predict_grid([3] * d, np.zeros(d), np.ones(d), 5, 0)
    assert mg.shape == (0,) and vg.shape == (0,)
    mf, Vf = dev.predict_fullcov(np.empty((0, d)))
    assert mf.shape == (0,) and Vf.shape == (0, 0)
    # implausibility of no points: zero counts, nothing kept
    Imax, keep, count, cmin, ccnt = dev.implausibility(np.empty((2, 0)), np.empty((2, 0)), [0.1, 0.2], [1e-2, 1e-2], 3.0, maxno=1)
    assert Imax.shape == (0, 1) and keep.shape == (0,) and int(count.sum()) == 0


def test_many_inputs_d48_needs_large_shared_memory_tiles(dev):
    """d = 48 inputs: the k-major X tiles of the covariance, gradient and cross-covariance kernels exceed the default
    48 KB of dynamic shared memory (opt-in per kernel and device).  Likelihood + gradient and prediction against the
    oracle at the usual tolerances."""
    rng = np.random.default_rng(48)
    n, d, m = 150, 48, 200
    X = rng.random((n, d))
    w = rng.normal(size=d) / np.sqrt(d)
    y = np.sin(X @ w) + 0.05 * rng.normal(size=n)
    H = np.ones((n, 1))
    delta = 2.0 + rng.random(d)
    sigma, nugget = 0.8, 1e-3
    dev.set_training(X, y, H)
    theta = np.r_[O.transform(delta), O.transform(np.array([sigma]))][None, :]
    llh, grad, _, st = dev.llh_grad_batch(theta, 0, fixed_nugget=nugget)
    want = O.loglikelihood_gp4ml(theta[0], X, y, H, kind=0, nugget_fixed=nugget)
    assert st[0] == 0 and want is not None
    assert abs(llh[0] - want[0]) <= 1e-10 * abs(want[0])
    assert np.all(np.abs(grad[0] - want[1]) <= 1e-9 * np.abs(want[1]).max() + 1e-12)
    beta, _, st = dev.fit_state(delta, nugget, sigma, kind=0)
    assert st == 0
    Xs = rng.random((m, d))
    mean, var = dev.predict(Xs, Hs=np.ones((m, 1)))
    A = O.make_A(X, delta, nugget, 0)
    mo, vo = O.posterior_diag_chunked(Xs, np.ones((m, 1)), X, y, H, A, O.optimalbeta(A, H, y), sigma, delta, nugget, 0)
    np.testing.assert_allclose(mean, mo, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, vo, rtol=1e-8, atol=1e-12)


def test_async_mode_two_handles_overlap_and_match_synchronous_results(dev, golden_dir):
    """gpe_set_async: with device-resident points and outputs a call returns once enqueued, so one host thread can keep
    two emulators' handles busy (history_match.py:96-118); results equal the synchronous ones bit for bit, and a call
    with a host pointer stays synchronous."""
    from gp_emu_uqsa_b200 import _lib
    G = np.load(os.path.join(golden_dir, "post_n200_d4.npz"))
    X, y = G["X"], G["y"]
    rng = np.random.default_rng(8)
    P = rng.random((5000, 4))
    devs, ref = [], []
    for k in range(2):
        dv = _lib.Device(0)
        dv.set_training(X, y * (1.0 + k), O.make_H_linear(X)); dv.set_basis([0, 1, 2, 3], [1] * 4)
        dv.fit_state(G["gp4ml_k_fixT_delta"], float(G["gp4ml_k_fixT_nugget"]), float(G["gp4ml_k_fixT_sigma"]), 0)
        devs.append(dv)
        ref.append(dv.predict(P))
    Pd = torch.tensor(P, device="cuda")
    outs = [(torch.empty(5000, dtype=torch.float64, device="cuda"), torch.empty(5000, dtype=torch.float64, device="cuda")) for _ in devs]
    torch.cuda.synchronize()
    for dv, o in zip(devs, outs):
        dv.set_async(True)
        dv.predict(Pd, None, out=o)            # returns once enqueued
    for dv in devs:
        dv.synchronize()
        dv.set_async(False)
    for (m, v), (mr, vr) in zip(outs, ref):
        assert np.array_equal(m.cpu().numpy(), mr) and np.array_equal(v.cpu().numpy(), vr)
    devs[0].set_async(True)
    mh, vh = devs[0].predict(P)                    # host buffers: complete on return even in asynchronous mode
    assert np.array_equal(mh, ref[0][0]) and np.array_equal(vh, ref[0][1])
    for dv in devs:
        dv.close()


def test_fused_implausibility_equals_the_two_step_route(dev, golden_dir):
    """gpe_predict_implaus (prediction with the implausibility folded in, one emulator per call) against gpe_predict +
    gpe_implausibility on the same points: identical lists, masks, counts and cell statistics -- explicit points and the
    flat-index grid, maxno 1 and 2, a shard that starts inside a cell, and an emulator entered as I = 0."""
    from gp_emu_uqsa_b200 import _lib
    G = np.load(os.path.join(golden_dir, "post_n200_d4.npz"))
    X, y = G["X"], G["y"]
    devs = []
    for k in range(3):
        dv = _lib.Device(0)
        dv.set_training(X, np.cos((k + 1) * y) + 0.1 * k, O.make_H_linear(X)); dv.set_basis([0, 1, 2, 3], [1] * 4)
        dv.fit_state(G["gp4ml_k_fixT_delta"], float(G["gp4ml_k_fixT_nugget"]), float(G["gp4ml_k_fixT_sigma"]), 0)
        devs.append(dv)
    zs, ve, cm = [0.3, 0.1, -0.2], [0.02, 0.05, 0.01], 1.5
    rng = np.random.default_rng(17)
    levels, lo, hi = np.array([6, 5, 7, 9]), np.zeros(4), np.ones(4)
    for mode in ("points", "grid"):
        m, first_index, cell_pts = 1337, 211, 100
        if mode == "points":
            P = rng.random((m, 4))
            mv = [dv.predict(P) for dv in devs]
        else:
            mv = [dv.predict_grid(levels, lo, hi, first_index, m) for dv in devs]
        means, variances = np.array([a for a, _ in mv]), np.array([b for _, b in mv])
        for maxno in (1, 2):
            Iref, kref, cref, minref, cntref = devs[0].implausibility(means, variances, zs, ve, cm, maxno=maxno, cell_pts=cell_pts,
                                                                    first_index=first_index)
            Itop = torch.full((m, maxno), -1.0, dtype=torch.float64, device="cuda")
            keep = torch.empty(m, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            for o, dv in enumerate(devs):
                kw = dict(points=P) if mode == "points" else dict(grid=(levels, lo, hi, first_index, m))
                res = dv.predict_implaus(zs[o], ve[o], Itop, first=(o == 0), last=(o == 2), maxno=maxno, cm=cm, cell_pts=cell_pts,
                                         first_index=first_index, keep=keep if o == 2 else None, **kw)
            cnt, cmin, ccnt = res
            assert np.array_equal(Itop.cpu().numpy(), Iref)
            assert np.array_equal(keep.cpu().numpy(), kref) and np.array_equal(cnt, cref)
            assert np.array_equal(cmin, minref) and np.array_equal(ccnt, cntref)
        # an emulator that contributes I = 0 is a zero in the initial list: same as mean = z, var = 1 in the two-step route
        means0, vars0 = means.copy(), variances.copy()
        means0[1], vars0[1] = zs[1], 1.0
        Iref, kref, cref, _, _ = devs[0].implausibility(means0, vars0, zs, ve, cm, maxno=2)
        Itop = torch.full((m, 2), -1.0, dtype=torch.float64, device="cuda")
        Itop[:, 1] = 0.0
        keep = np.empty(m, dtype=np.uint8)                          # host mask
        torch.cuda.synchronize()
        for o in (0, 2):
            kw = dict(points=P) if mode == "points" else dict(grid=(levels, lo, hi, first_index, m))
            res = devs[o].predict_implaus(zs[o], ve[o], Itop, first=False, last=(o == 2), maxno=2, cm=cm, keep=keep if o == 2 else None, **kw)
        assert np.array_equal(Itop.cpu().numpy(), Iref) and np.array_equal(keep, kref) and np.array_equal(res[0], cref)
    for dv in devs:
        dv.close()


@pytest.mark.parametrize("n,d,m,kind", [(1500, 6, 2300, 0), (2000, 8, 4096, 0), (1024, 3, 1024, 0), (1100, 5, 2048, 0),
                                        (1500, 6, 2300, 1), (4096, 16, 1024, 0)])
def test_int8_route_of_the_prediction_product_matches_dmma_and_the_oracle(n, d, m, kind, monkeypatch):
    """Chunks of 1024 points and more (a multiple of 256 after padding) over 1024 and more padded training points send
    Z = L^-1 C down the INT8 tensor-core route (gpe_ozaki.cuh: one scale for the bounded slab, the residue planes of L^-1 kept per
    fit, the product with its roles swapped and row norms taken inside the CRT pass when the padded size is a multiple of 256 --
    n = 1100 keeps L^-1 on the row side).  Held against GPE_OZAKI=0 (FP64 DMMA) and, for kernel 0, against the oracle's route
    (north_star: 1e-8), at a length-scale mix that makes L^-1 badly scaled; the last 16 points lie far outside the training
    inputs (tiny slab columns under the common scale); kind 1 = the alternative nugget with a per-point r."""
    from gp_emu_uqsa_b200 import _lib
    rng = np.random.default_rng(11 + n)
    X = rng.random((n, d))
    y = np.sin(X @ rng.normal(size=d)) + 0.1 * (X ** 2).sum(1)
    H = O.make_H_linear(X)
    Xs = rng.random((m, d))
    Xs[-16:] += 3.0
    delta = np.linspace(0.3, 1.5, d)
    nugget, sigma = (1e-5, 1.3) if kind == 0 else (3e-3, 1.3)
    r = None if kind == 0 else 1e-4 * (1.0 + rng.random(n))
    res = {}
    for route in ("dmma", "int8"):
        if route == "dmma":
            monkeypatch.setenv("GPE_OZAKI", "0")
        else:
            monkeypatch.delenv("GPE_OZAKI", raising=False)
        dv = _lib.Device(0)
        try:
            dv.set_training(X, y, H, r)
            dv.set_basis(list(range(d)), [1] * d)
            _, _, st = dv.fit_state(delta, nugget, sigma, kind)
            assert st == 0
            before = dv.int8_products
            mean, var = dv.predict(Xs)
            took = dv.int8_products - before
            # a second call re-uses the residue planes of L^-1 (same fit generation); a new fit must not
            mean_b, var_b = dv.predict(Xs)
            assert np.array_equal(mean_b, mean) and np.array_equal(var_b, var)
            dv.fit_state(delta * 1.1, nugget, sigma, kind)
            dv.predict(Xs[:1024])
            dv.fit_state(delta, nugget, sigma, kind)
            mean_c, var_c = dv.predict(Xs)
            assert np.array_equal(mean_c, mean) and np.array_equal(var_c, var)
            res[route] = (mean, var, took)
        finally:
            dv.close()
    assert res["dmma"][2] == 0
    assert res["int8"][2] >= 1, "the INT8 route was not taken"
    # (the mean does not go through Z, but from npad = 2048 on the fit's own large products take the INT8 route too)
    assert np.abs(res["int8"][0] - res["dmma"][0]).max() <= 1e-10 * np.abs(res["dmma"][0]).max()
    vd, vi = res["dmma"][1], res["int8"][1]
    scale = sigma ** 2                                                         # the prior variance the subtraction starts from
    assert np.abs(vi - vd).max() <= 1e-10 * scale, np.abs(vi - vd).max() / scale
    assert np.all(vi[-16:] > 0.99 * scale)                                     # far points: the prior variance (and more)
    if kind != 0:
        return
    A = O.make_A(X, delta, nugget, 0)
    beta = O.optimalbeta(A, H, y)
    sel = np.r_[0:240, m - 16:m]
    mref, Vref = O.posterior(Xs[sel], O.make_H_linear(Xs[sel]), X, y, H, A, beta, sigma, delta, nugget, 0)
    vref = np.diag(Vref)
    assert np.allclose(res["int8"][0][sel], mref, rtol=1e-8, atol=1e-10)
    assert np.allclose(vi[sel], vref, rtol=1e-8, atol=1e-8 * np.abs(vref).max())
