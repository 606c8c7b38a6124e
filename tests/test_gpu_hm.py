"""History matching and noise fit through the reference-facing functions on the B200, against
goldens produced by the real reference's imp_plot / nonimp_data / new_wave_design
(tests/golden/make_golden.py hmapi)."""
import contextlib
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


@contextlib.contextmanager
def _quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def _rebuild_emulators(g, gold, tmp):
    emuls = []
    with _cwd(tmp), _quiet():
        for o in (0, 1):
            for fn in ("hm%d_config_r" % o, "hm%d_beliefs-0f" % o, "hm%d_input-o0-0f" % o, "hm%d_output-o0-0f" % o):
                with open(fn, "wb") as f:
                    f.write(bytes(gold["file_" + fn]))
            emuls.append(g.setup("hm%d_config_r" % o, datashuffle=False, scaleinputs=True))
    return emuls


def test_history_match_functions_match_reference(golden_dir, tmp_path):
    import gp_emu_uqsa_b200 as g
    import gp_emu_uqsa_b200.history_match as h
    gold = np.load(os.path.join(golden_dir, "hmapi_n100_d3.npz"))
    emuls = _rebuild_emulators(g, gold, tmp_path)
    zs, ve, cm = list(gold["zs"]), list(gold["var_extra"]), float(gold["cm"])
    with _cwd(tmp_path), _quiet():
        np.random.seed(77)
        h.imp_plot(emuls, zs, cm, ve, maxno=2, olhcmult=30, grid=4, plot=False, fileStr="g")
        for s_ in ([0, 1], [0, 2], [1, 2]):
            assert np.array_equal(np.loadtxt("imp_input_%d_%d" % tuple(s_)), gold["lhc_%d_%d" % tuple(s_)])
            for m in (1, 2):
                imp = np.loadtxt("g_%d_IMP_%d_%d" % (m, s_[0], s_[1]))
                odp = np.loadtxt("g_%d_ODP_%d_%d" % (m, s_[0], s_[1]))
                assert np.allclose(imp, gold["g_%d_IMP_%d_%d" % (m, s_[0], s_[1])], rtol=1e-7, atol=1e-9)
                assert np.array_equal(odp, gold["g_%d_ODP_%d_%d" % (m, s_[0], s_[1])])
                # byte level: the optical-depth files hold count ratios -> the reference's bytes exactly;
                # the implausibility files hold minima that agree to rounding -> same layout, every number parsed equal to 1e-7
                assert open("g_%d_ODP_%d_%d" % (m, s_[0], s_[1]), "rb").read() == bytes(gold["bytes_g_%d_ODP_%d_%d" % (m, s_[0], s_[1])])
                mine_txt = open("g_%d_IMP_%d_%d" % (m, s_[0], s_[1])).read().split("\n")
                ref_txt = bytes(gold["bytes_g_%d_IMP_%d_%d" % (m, s_[0], s_[1])]).decode().split("\n")
                assert len(mine_txt) == len(ref_txt) and [len(a) for a in mine_txt] == [len(b) for b in ref_txt]
        rec = h.imp_plot_recon(cm, maxno=1, act=[0, 1, 2], fileStr="g")
        assert set(rec) == {(0, 1), (0, 2), (1, 2)}
        np.savetxt("sim_in", gold["sim_in"], fmt="%.17g")
        np.savetxt("sim_out", np.column_stack([gold["sim_in"][:, 0], gold["sim_in"][:, 1]]), fmt="%.17g")
        cnt = h.nonimp_data(emuls, zs, cm, ve, ["sim_in", "sim_out"], maxno=1)
        assert cnt == int(gold["nonimp_count"])
        assert np.allclose(np.atleast_2d(np.loadtxt("nonimp_sim_in")), gold["nonimp_in"], rtol=0, atol=1e-15)
        assert np.allclose(np.atleast_2d(np.loadtxt("noninp_sim_out")), gold["nonimp_out"], rtol=0, atol=1e-15)
        for fn in ("nonimp_sim_in", "noninp_sim_out"):          # identical keep set -> identical bytes
            assert open(fn, "rb").read() == bytes(gold["bytes_" + fn]), fn
        np.random.seed(78)
        cnt2 = h.new_wave_design(emuls, zs, cm, ve, ["nonimp_sim_in", "noninp_sim_out"], maxno=1, olhcmult=40, fileStr="w2")
        assert np.array_equal(np.loadtxt("olhc_des"), gold["olhc_des"])
        assert cnt2 == int(gold["wave_count"])
        assert np.allclose(np.atleast_2d(np.loadtxt("w2_nonimp_sim_in")), gold["wave_in"], rtol=0, atol=1e-15)
        for fn in ("olhc_des", "w2_nonimp_sim_in"):
            assert open(fn, "rb").read() == bytes(gold["bytes_" + fn]), fn


def _beliefs(text):
    return {ln.split(" ", 1)[0]: ln.split(" ", 1)[1].strip() for ln in text.splitlines() if " " in ln}


def test_noisefit_matches_reference_run(tmp_path, golden_dir):
    """noise_fit.noisefit on the reduced noisefit2D case against the real reference's run of the same case
    (tests/golden/make_golden.py noisefit): two alternations of data GP (alt nugget with an r vector, free nugget)
    and noise GP, 50 posterior samples each.  The global NumPy RNG is consumed in the reference's order
    (shuffles, multistart guesses, posterior draws, Latin hypercubes), so the design and the result points are
    bit-identical; the fitted values agree to the optimiser's tolerance."""
    import importlib.util
    import gp_emu_uqsa_b200.design_inputs as d
    import gp_emu_uqsa_b200.noise_fit as gn
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    G = np.load(os.path.join(golden_dir, "noisefit_n150.npz"))
    with _cwd(tmp_path), _quiet():
        mg.write_noisefit_case(d)
        x, y = np.loadtxt("INPUTS"), np.loadtxt("OUTPUTS")
        assert np.array_equal(x, G["INPUTS"]) and np.array_equal(y, G["OUTPUTS"])
        np.random.seed(11)
        gn.noisefit("config-data", "config-noise", stopat=2, olhcmult=10, samples=50)
        xin, out, zp = np.loadtxt("noise-inputs"), np.loadtxt("noise-outputs"), np.loadtxt("zp-outputs")
        mine = {fn: _beliefs(open(fn).read()) for fn in ("beliefs-data-0f", "beliefs-noise-0f")}
    assert np.array_equal(xin, G["noise_inputs"])
    assert out.shape == (20, 3) and np.all(out[:, 1] <= out[:, 0]) and np.all(out[:, 0] <= out[:, 2])
    np.testing.assert_allclose(zp, G["zp_outputs"], rtol=0, atol=2e-3 * np.abs(G["zp_outputs"]).max())
    np.testing.assert_allclose(out, G["noise_outputs"], rtol=1e-2)
    for fn, got in mine.items():
        want = _beliefs(bytes(G["file_" + fn]).decode())
        assert got.keys() == want.keys()
        for key in ("beta", "delta", "sigma", "nugget"):
            np.testing.assert_allclose(np.array(got[key].split(), dtype=float), np.array(want[key].split(), dtype=float),
                                       rtol=2e-2, err_msg=fn + " " + key)
        for key in ("active_index", "active", "output", "basis_str", "fix_nugget", "alt_nugget", "mucm", "input_minmax"):
            assert got[key] == want[key], (fn, key)


@pytest.mark.parametrize("n,dim,ne", [(30, 1, 0), (100, 3, 0), (257, 8, 0), (120, 3, 202), (64, 10, 1000)])
def test_device_maximin_criterion_is_bit_identical_to_scipy(n, dim, ne):
    """gpe_pdist_argmin == np.argmin(scipy pdist 'sqeuclidean') for every candidate design, including
    near-ties (duplicated points make exact ties: the first index must win)."""
    import gp_emu_uqsa_b200.design_inputs.design_inputs as d
    rng = np.random.default_rng(n + dim + ne)
    designs = rng.random((7, n, dim))
    designs[3, n // 2] = designs[3, 1]            # exact tie at distance 0 with another pair below
    designs[3, n - 1] = designs[3, 0]
    extra = rng.random((ne, dim)) if ne else None
    want = d._host_criterion(designs, extra)
    got = d.device_criterion(designs, extra)
    assert np.array_equal(got, want)


def test_config4_shape_matches_real_reference(golden_dir, tmp_path):
    """Config 4's shape against the REAL reference (tests/golden/make_golden.py config4): two n = 2000, d = 8 emulators
    rebuilt from the reference's own checkpoint files; g.posterior's mean and diagonal variance on 3000 points of the
    10-level grid to 1e-8 -- through the explicit-point path and, for the run of consecutive flat indices, through the
    flat-index grid path (tabulated factors) -- and history_match.nonimp_data's kept rows identical."""
    import gp_emu_uqsa_b200 as g
    import gp_emu_uqsa_b200.history_match as h
    G = np.load(os.path.join(golden_dir, "config4_n2000_d8.npz"))
    emuls = []
    with _cwd(tmp_path), _quiet():
        for o in (0, 1):
            for fn in ("c4_%d_config_r" % o, "c4_%d_beliefs-0f" % o, "c4_%d_input-o0-0f" % o, "c4_%d_output-o0-0f" % o):
                with open(fn, "wb") as f:
                    f.write(bytes(G["file_" + fn]))
            emuls.append(g.setup("c4_%d_config_r" % o, datashuffle=False, scaleinputs=True))
        P, idx = G["P_scaled"], G["grid_index"]
        run = np.nonzero((idx >= 36999700) & (idx < 36999700 + 600))[0]
        assert run.size == 600 and np.array_equal(idx[run], np.arange(36999700, 36999700 + 600))
        for o, E in enumerate(emuls):
            assert np.allclose(E.par.beta, G["beta%d" % o], rtol=1e-12, atol=0)       # read back from the checkpoint
            mean, var = g.posterior_diag(E, P)
            assert np.allclose(mean, G["mean%d" % o], rtol=1e-8, atol=1e-10)
            assert np.allclose(var, G["var%d" % o], rtol=1e-8, atol=1e-9 * G["var%d" % o].max())
            dev, _, _, st = E.training.fit(beta=E.par.beta, r_div=E.training._A_args[0])
            assert st == 0
            gm, gv = dev.predict_grid(np.full(8, 10, dtype=np.int32), np.zeros(8), np.ones(8), 36999700, 600)
            assert np.allclose(gm, G["mean%d" % o][run], rtol=1e-8, atol=1e-10)
            assert np.allclose(gv, G["var%d" % o][run], rtol=1e-8, atol=1e-9 * G["var%d" % o].max())
        np.savetxt("sim_in", G["sim_in"], fmt="%.17g")
        np.savetxt("sim_out", np.column_stack([G["sim_in"][:, 0], G["sim_in"][:, 1]]), fmt="%.17g")
        cnt = h.nonimp_data(emuls, list(G["zs"]), float(G["cm"]), list(G["var_extra"]), ["sim_in", "sim_out"], maxno=1)
        kept = np.atleast_2d(np.loadtxt("nonimp_sim_in"))
    assert cnt == int(G["nonimp_count"])
    assert np.allclose(kept, G["nonimp_in"], rtol=0, atol=1e-15)
