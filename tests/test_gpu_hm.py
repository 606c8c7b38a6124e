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
        rec = h.imp_plot_recon(cm, maxno=1, act=[0, 1, 2], fileStr="g")
        assert set(rec) == {(0, 1), (0, 2), (1, 2)}
        np.savetxt("sim_in", gold["sim_in"], fmt="%.17g")
        np.savetxt("sim_out", np.column_stack([gold["sim_in"][:, 0], gold["sim_in"][:, 1]]), fmt="%.17g")
        cnt = h.nonimp_data(emuls, zs, cm, ve, ["sim_in", "sim_out"], maxno=1)
        assert cnt == int(gold["nonimp_count"])
        assert np.allclose(np.atleast_2d(np.loadtxt("nonimp_sim_in")), gold["nonimp_in"], rtol=0, atol=1e-15)
        assert np.allclose(np.atleast_2d(np.loadtxt("noninp_sim_out")), gold["nonimp_out"], rtol=0, atol=1e-15)
        np.random.seed(78)
        cnt2 = h.new_wave_design(emuls, zs, cm, ve, ["nonimp_sim_in", "noninp_sim_out"], maxno=1, olhcmult=40, fileStr="w2")
        assert np.array_equal(np.loadtxt("olhc_des"), gold["olhc_des"])
        assert cnt2 == int(gold["wave_count"])
        assert np.allclose(np.atleast_2d(np.loadtxt("w2_nonimp_sim_in")), gold["wave_in"], rtol=0, atol=1e-15)


def test_noisefit_runs_on_device(tmp_path):
    """noisefit2D-style driver at reduced size: data GP (alt nugget, r vector) + noise GP alternate once;
    result files have the reference's shapes and the fitted noise level is of the right magnitude."""
    import gp_emu_uqsa_b200.design_inputs as d
    import gp_emu_uqsa_b200.noise_fit as gn
    with _cwd(tmp_path), _quiet():
        np.random.seed(3)
        d.optLatinHyperCube(2, 150, 20, [[0.0, 1.0], [0.0, 1.0]], "INPUTS")
        x = np.loadtxt("INPUTS")
        mean = 3.0 * x[:, 0] ** 3 + np.exp(np.cos(10.0 * x[:, 1]) * np.cos(5.0 * x[:, 0]) ** 2)
        noise = 0.5 * (x[:, 1] * (np.cos(6 * x[:, 0]) ** 2 + 0.1))
        np.savetxt("OUTPUTS", mean + noise * np.random.randn(x.shape[0]))
        for name, outputs, extra in (("data", "OUTPUTS", ("alt_nugget T", "constraints none", "[[0.05,10.0],[0.05,10.00]]", "[[0.1,3.0]]", "[[0.001,1.05]]")),
                                     ("noise", "zp-outputs", ("alt_nugget F", "constraints bounds", "[[0.05,1.0],[0.05,10.00]]", "[[0.001,10.0]]", "[[0.0001,1.0]]"))):
            with open("config-" + name, "w") as f:
                f.write("beliefs beliefs-%s\ninputs INPUTS\noutputs %s\ntv_config 10 0 0\ndelta_bounds %s\nsigma_bounds %s\n"
                        "nugget_bounds %s\ntries 2\n%s\n" % (name, outputs, extra[2], extra[3], extra[4], extra[1]))
            with open("beliefs-" + name, "w") as f:
                f.write("active all\noutput 0\nbasis_str 1.0\nbasis_inf NA\nbeta 1.0\ndelta 1.0 1.0\nsigma 1.0\nnugget 0.00001\n"
                        "fix_nugget F\n%s\nmucm F\n" % extra[0])
        gn.noisefit("config-data", "config-noise", stopat=1, olhcmult=10, samples=20)
        xin, out = np.loadtxt("noise-inputs"), np.loadtxt("noise-outputs")
    assert xin.shape == (20, 2) and out.shape == (20, 3)
    assert np.all(np.isfinite(out)) and np.all(out > 0) and np.all(out[:, 1] <= out[:, 0]) and np.all(out[:, 0] <= out[:, 2])
    assert 0.01 < np.median(out[:, 0]) < 2.0


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
