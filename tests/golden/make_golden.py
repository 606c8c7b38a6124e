"""Generate golden vectors from the REAL reference (/root/reference) -- build container only.

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Every file stores the inputs (X, y, theta, ...) and what the
unmodified reference returned for them, so the GPU box (which has no /root/reference) can
check both the NumPy oracle and the CUDA path against the reference's own outputs.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_loader as RL  # noqa: E402


def synth(n, d, seed=0):
    """BASELINE.md section 3 generator."""
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    w = rng.normal(size=d)
    y = np.sin(X @ w) + 0.1 * (X ** 2).sum(1)
    return X, y, w


MODES = [  # (tag, mucm, alt_nugget, fix_nugget)
    ("mucm_k_fixT", "T", "F", "T"),
    ("mucm_k_fixF", "T", "F", "F"),
    ("gp4ml_k_fixT", "F", "F", "T"),
    ("gp4ml_k_fixF", "F", "F", "F"),
    ("gp4ml_alt_fixT", "F", "T", "T"),
    ("gp4ml_alt_fixF", "F", "T", "F"),
]


def build(g, tmp, X, y, mucm, alt, fix, nugget, name, tv_config="10 0 0"):
    with RL.cwd(tmp), RL.quiet():
        cfg = RL.write_emulator_files(tmp, X, y, mucm=mucm, fix_nugget=fix, alt_nugget=alt,
                                      nugget=nugget, name=name, tv_config=tv_config)
        E = g.setup(cfg, datashuffle=False, scaleinputs=True)
    return E


def gen_llh(g, n, d, seed, n_theta, nugget=1e-4, with_r=False):
    X, y, _ = synth(n, d, seed)
    out = {"X_raw": X, "y": y, "nugget_belief": nugget}
    rng = np.random.default_rng(100 + seed)
    with tempfile.TemporaryDirectory() as tmp:
        for tag, mucm, alt, fix in MODES:
            E = build(g, tmp, X, y, mucm, alt, fix, nugget, "m_" + tag)
            out["X"] = E.training.inputs.copy()          # scaled inputs as the reference holds them
            out["H"] = E.training.H.copy()
            r = None
            if with_r and alt == "T":
                r = 0.01 + 0.02 * rng.random(n)
                with RL.quiet():
                    E.training.set_r(r)
                out[tag + "_r"] = r
            p = d + (1 if fix == "F" else 0) + (1 if mucm == "F" else 0)
            thetas, llhs, grads, sig = [], [], [], []
            for t in range(n_theta):
                delta = 0.15 + 0.85 * rng.random(d)
                hp = list(delta)
                if fix == "F":
                    hp.append(10 ** rng.uniform(-4, -2))
                if mucm == "F":
                    hp.append(0.3 + 1.2 * rng.random())
                theta = E.K.transform(np.array(hp))
                with RL.quiet():
                    res = (E.opt_T.loglikelihood_mucm if mucm == "T" else E.opt_T.loglikelihood_gp4ml)(theta.copy())
                assert res is not None
                thetas.append(theta); llhs.append(res[0]); grads.append(res[1].copy())
                sig.append(float(E.par.sigma))
            out[tag + "_theta"] = np.array(thetas)
            out[tag + "_llh"] = np.array(llhs)
            out[tag + "_grad"] = np.array(grads)
            out[tag + "_sigma"] = np.array(sig)
            assert out[tag + "_theta"].shape[1] == p
    return out


def gen_posterior(g, n, d, seed, m, nugget=1e-4):
    X, y, _ = synth(n, d, seed)
    out = {"X_raw": X, "y": y}
    rng = np.random.default_rng(200 + seed)
    Xs = rng.random((m, d))
    out["Xs"] = Xs
    with tempfile.TemporaryDirectory() as tmp:
        for tag, mucm, alt, fix in (MODES[0], MODES[2], MODES[4]):
            E = build(g, tmp, X, y, mucm, alt, fix, nugget, "p_" + tag)
            out["X"] = E.training.inputs.copy()
            delta = 0.3 + 0.5 * rng.random(d)
            sigma = 0.8 + 0.4 * rng.random()
            E.par.delta = delta.copy(); E.K.d = E.par.delta; E.K.n = E.par.nugget
            E.par.sigma = sigma
            with RL.quiet():
                if alt == "T":
                    r = 0.01 + 0.02 * rng.random(n)
                    E.training.set_r(r)
                    out[tag + "_r"] = r
                E.training.remake()
                E.opt_T.optimalbeta()
                mean, var = g.posterior(E, Xs.copy())
            out[tag + "_delta"] = delta
            out[tag + "_sigma"] = sigma
            out[tag + "_nugget"] = float(E.par.nugget)
            out[tag + "_beta"] = np.array(E.par.beta)
            out[tag + "_mean"] = mean
            out[tag + "_var"] = var
    return out


def gen_toysim(g):
    """Config 1: examples/toy-sim as shipped, np.random.seed(0)."""
    import shutil
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("toy-sim_config", "toy-sim_beliefs", "toy-sim_input", "toy-sim_output"):
            shutil.copy(os.path.join(RL.REF_ROOT, "examples", "toy-sim", f), tmp)
        with RL.cwd(tmp), RL.quiet():
            np.random.seed(0)
            E = g.setup("toy-sim_config")
            g.train(E)
            xs = np.array([[0.1, 0.2], [0.5, 0.5], [0.9, 0.3], [0.25, 0.75]])
            mean, var = g.posterior(E, xs)
        out.update(delta=np.array(E.par.delta), sigma=float(E.par.sigma), nugget=float(E.par.nugget),
                   beta=np.array(E.par.beta), X=E.training.inputs, y=E.training.outputs,
                   xs=xs, mean=mean, var=var)
    return out


def gen_toysim_log(g):
    """Everything the reference prints during g.setup + g.train of examples/toy-sim (seed 0): the per-guess
    optimiser lines, the validation diagnostics (Mahalanobis distances, individual standard errors), the
    include-V-into-T steps and the checkpoint file names -- the train/validation orchestration as text."""
    import contextlib
    import io
    import shutil
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("toy-sim_config", "toy-sim_beliefs", "toy-sim_input", "toy-sim_output"):
            shutil.copy(os.path.join(RL.REF_ROOT, "examples", "toy-sim", f), tmp)
        buf = io.StringIO()
        with RL.cwd(tmp), contextlib.redirect_stdout(buf):
            np.random.seed(0)
            E = g.setup("toy-sim_config")
            g.train(E)
    return {"log": np.frombuffer(buf.getvalue().encode(), dtype=np.uint8)}


def gen_hm(g, h, n, seed):
    """History matching: two emulators on 3 inputs; nonimp_data-style flat implausibility.
    Stores means/variances from the reference Posterior and the reference's own
    keep decision computed with history_match.py:237-250 arithmetic via nonimp_data."""
    X, y, w = synth(n, 3, seed)
    rng = np.random.default_rng(300 + seed)
    y2 = np.cos(X @ rng.normal(size=3))
    out = {"X_raw": X, "y0": y, "y1": y2}
    with tempfile.TemporaryDirectory() as tmp, RL.cwd(tmp), RL.quiet():
        emuls = []
        for o, yy in enumerate((y, y2)):
            cfg = RL.write_emulator_files(tmp, X, yy, mucm="F", fix_nugget="T", alt_nugget="F",
                                          nugget=1e-4, name="hm%d" % o, tries=3)
            np.random.seed(5 + o)
            E = g.setup(cfg, datashuffle=False, scaleinputs=True)
            g.train(E)
            # rebuild from the updated beliefs (active_index / input_minmax needed by HM)
            with open("hm%d_config_r" % o, "w") as f:
                f.write("beliefs hm%d_beliefs-0f\ninputs hm%d_input-o0-0f\noutputs hm%d_output-o0-0f\n" % (o, o, o))
                f.write("tv_config 10 0 0\ndelta_bounds [ ]\nsigma_bounds [ ]\nnugget_bounds [ ]\ntries 1\nconstraints bounds\n")
            E2 = g.setup("hm%d_config_r" % o, datashuffle=False, scaleinputs=True)
            emuls.append(E2)
            out["delta%d" % o] = np.array(E2.par.delta); out["sigma%d" % o] = float(E2.par.sigma)
            out["beta%d" % o] = np.array(E2.par.beta); out["nugget%d" % o] = float(E2.par.nugget)
            out["Xtrain%d" % o] = E2.training.inputs.copy(); out["ytrain%d" % o] = E2.training.outputs.copy()
        m = 400
        pts = rng.random((m, 3))
        np.savetxt("sim_in", pts, fmt="%.17g")
        np.savetxt("sim_out", np.column_stack([pts[:, 0], pts[:, 1]]), fmt="%.17g")
        zs = [float(np.median(y)), float(np.median(y2))]
        ve = [1e-2, 1e-2]
        cm = 3.0
        for maxno in (1, 2):
            cnt = h.nonimp_data(emuls, zs, cm, ve, ["sim_in", "sim_out"], maxno=maxno)
            kept = np.atleast_2d(np.loadtxt("nonimp_sim_in")) if cnt else np.zeros((0, 3))
            out["kept_maxno%d" % maxno] = kept
            out["count_maxno%d" % maxno] = cnt
        out.update(pts=pts, zs=np.array(zs), var_extra=np.array(ve), cm=cm)
        # nonimp_data scales the file's inputs with the emulators' input_minmax
        # (_hmutilfunctions.py:126-139) before predicting; the kept rows are the scaled ones
        mm = np.array(emuls[0].beliefs.input_minmax)
        pts_scaled = (pts - mm[:, 0]) / (mm[:, 1] - mm[:, 0])
        out["pts_scaled"] = pts_scaled
        means, variances = [], []
        for E2 in emuls:
            mu, V = g.posterior(E2, pts_scaled.copy())
            means.append(mu); variances.append(np.diag(V))
        out["means"] = np.array(means); out["vars"] = np.array(variances)
    return out


def gen_sens(g, s, n, d, seed, points=25):
    """Sensitivity (MUCM case 2): uncertainty / sensitivity / main_effect / totaleffectvariance of the
    real reference for a fixed-hyper-parameter emulator."""
    X, y, _ = synth(n, d, seed)
    rng = np.random.default_rng(400 + seed)
    out = {"X_raw": X, "y": y}
    with tempfile.TemporaryDirectory() as tmp:
        E = build(g, tmp, X, y, "F", "F", "T", 1e-4, "sens")
        delta = 0.3 + 0.5 * rng.random(d)
        sigma = 0.8 + 0.4 * rng.random()
        E.par.delta = delta.copy(); E.K.d = E.par.delta; E.K.n = E.par.nugget
        E.par.sigma = sigma
        with RL.quiet():
            E.training.remake()
            E.opt_T.optimalbeta()
            m = 0.4 + 0.2 * rng.random(d)
            v = 0.01 + 0.03 * rng.random(d)
            S = s.setup(E, list(m), list(v))
            S.uncertainty()
            S.sensitivity()
            S.main_effect(plot=False, points=points)
            S.totaleffectvariance()
        out.update(X=E.training.inputs.copy(), H=E.training.H.copy(), A=E.training.A.copy(), delta=delta, sigma=sigma,
                   nugget=float(E.par.nugget), beta=np.array(E.par.beta), m=m, v=v,
                   input_range=np.array(E.all_data.input_range), uE=S.uE, uV=S.uV, uEV=S.uEV,
                   senseindex=S.senseindex, effect=S.effect, mean_effect=S.mean_effect, EVTw=S.EVTw,
                   e=S.e, G=S.G, W=S.W)
    return out


def gen_hm_api(g, h, n, seed):
    """History matching through the reference's public functions: imp_plot (IMP/ODP matrices per input
    pair), nonimp_data and new_wave_design (kept rows), for two trained 3-input emulators rebuilt from
    their updated beliefs files.  The text of those checkpoint files is stored so the test can rebuild
    byte-identical emulators without the reference."""
    X, y, w = synth(n, 3, seed)
    rng = np.random.default_rng(300 + seed)
    y2 = np.cos(X @ rng.normal(size=3))
    out = {}
    with tempfile.TemporaryDirectory() as tmp, RL.cwd(tmp), RL.quiet():
        emuls = []
        for o, yy in enumerate((y, y2)):
            cfg = RL.write_emulator_files(tmp, X, yy, mucm="F", fix_nugget="T", alt_nugget="F",
                                          nugget=1e-4, name="hm%d" % o, tries=3)
            np.random.seed(5 + o)
            E = g.setup(cfg, datashuffle=False, scaleinputs=True)
            g.train(E)
            with open("hm%d_config_r" % o, "w") as f:
                f.write("beliefs hm%d_beliefs-0f\ninputs hm%d_input-o0-0f\noutputs hm%d_output-o0-0f\n" % (o, o, o))
                f.write("tv_config 10 0 0\ndelta_bounds [ ]\nsigma_bounds [ ]\nnugget_bounds [ ]\ntries 1\nconstraints bounds\n")
            for fn in ("hm%d_config_r" % o, "hm%d_beliefs-0f" % o, "hm%d_input-o0-0f" % o, "hm%d_output-o0-0f" % o):
                out["file_" + fn] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
            emuls.append(g.setup("hm%d_config_r" % o, datashuffle=False, scaleinputs=True))
        zs = [float(np.median(y)), float(np.median(y2))]
        ve = [1e-2, 1e-2]
        cm = 3.0
        out.update(zs=np.array(zs), var_extra=np.array(ve), cm=cm)
        np.random.seed(77)
        h.imp_plot(emuls, zs, cm, ve, maxno=2, olhcmult=30, grid=4, plot=False, fileStr="g")
        for s_ in ([0, 1], [0, 2], [1, 2]):
            for m in (1, 2):
                for kind in ("IMP", "ODP"):
                    name = "g_%d_%s_%d_%d" % (m, kind, s_[0], s_[1])
                    out[name] = np.loadtxt(name)
                    out["bytes_" + name] = np.frombuffer(open(name, "rb").read(), dtype=np.uint8)
            out["lhc_%d_%d" % tuple(s_)] = np.loadtxt("imp_input_%d_%d" % tuple(s_))
        pts = rng.random((300, 3))
        np.savetxt("sim_in", pts, fmt="%.17g")
        np.savetxt("sim_out", np.column_stack([pts[:, 0], pts[:, 1]]), fmt="%.17g")
        out["sim_in"] = pts
        cnt = h.nonimp_data(emuls, zs, cm, ve, ["sim_in", "sim_out"], maxno=1)
        out["nonimp_count"] = cnt
        out["nonimp_in"] = np.atleast_2d(np.loadtxt("nonimp_sim_in"))
        out["nonimp_out"] = np.atleast_2d(np.loadtxt("noninp_sim_out"))
        for fn in ("nonimp_sim_in", "noninp_sim_out"):
            out["bytes_" + fn] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
        np.random.seed(78)
        cnt2 = h.new_wave_design(emuls, zs, cm, ve, ["nonimp_sim_in", "noninp_sim_out"], maxno=1, olhcmult=40, fileStr="w2")
        out["wave_count"] = cnt2
        out["wave_in"] = np.atleast_2d(np.loadtxt("w2_nonimp_sim_in"))
        out["olhc_des"] = np.loadtxt("olhc_des")
        for fn in ("w2_nonimp_sim_in", "olhc_des"):
            out["bytes_" + fn] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
    return out


def gen_toysim_recon(g):
    """examples/toy-sim/reconstruct: an emulator rebuilt from shipped (old-format) beliefs + data files, no
    training; posterior mean / covariance of the real reference at fixed points."""
    import shutil
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(RL.REF_ROOT, "examples", "toy-sim", "reconstruct")
        for f in os.listdir(src):
            if f.startswith("toy-sim_"):
                shutil.copy(os.path.join(src, f), tmp)
        with RL.cwd(tmp), RL.quiet():
            E = g.setup("toy-sim_config_recon", datashuffle=False, scaleinputs=True)
            xs = np.random.default_rng(9).random((64, 2))
            mean, var = g.posterior(E, xs.copy())
        out.update(xs=xs, mean=mean, var=var, X=E.training.inputs.copy(), y=E.training.outputs.copy(),
                   nT=E.training.inputs.shape[0], nV=E.validation.inputs.shape[0])
    return out


NOISEFIT_CFG = (("data", "OUTPUTS", ("alt_nugget T", "constraints none", "[[0.05,10.0],[0.05,10.00]]", "[[0.1,3.0]]", "[[0.001,1.05]]")),
                ("noise", "zp-outputs", ("alt_nugget F", "constraints bounds", "[[0.05,1.0],[0.05,10.00]]", "[[0.001,10.0]]", "[[0.0001,1.0]]")))


def write_noisefit_case(design_module, n=150, seed=3, tries=3):
    """The noisefit2D example (examples/noisefit2D/emulator.py:12-32) at reduced size, written into the cwd.
    Shared by the generator below and by tests/test_gpu_hm.py (which passes its own design module)."""
    np.random.seed(seed)
    design_module.optLatinHyperCube(2, n, 20, [[0.0, 1.0], [0.0, 1.0]], "INPUTS")
    x = np.loadtxt("INPUTS")
    mean = 3.0 * x[:, 0] ** 3 + np.exp(np.cos(10.0 * x[:, 1]) * np.cos(5.0 * x[:, 0]) ** 2)
    noise = 0.5 * (x[:, 1] * (np.cos(6 * x[:, 0]) ** 2 + 0.1))
    np.savetxt("OUTPUTS", mean + noise * np.random.randn(x.shape[0]))
    for name, outputs, extra in NOISEFIT_CFG:
        with open("config-" + name, "w") as f:
            f.write("beliefs beliefs-%s\ninputs INPUTS\noutputs %s\ntv_config 10 0 0\ndelta_bounds %s\nsigma_bounds %s\n"
                    "nugget_bounds %s\ntries %d\n%s\n" % (name, outputs, extra[2], extra[3], extra[4], tries, extra[1]))
        with open("beliefs-" + name, "w") as f:
            f.write("active all\noutput 0\nbasis_str 1.0\nbasis_inf NA\nbeta 1.0\ndelta 1.0 1.0\nsigma 1.0\nnugget 0.00001\n"
                    "fix_nugget F\n%s\nmucm F\n" % extra[0])


def gen_noisefit(gn):
    """noise_fit.noisefit of the real reference on the reduced noisefit2D case: two alternations, 50 posterior
    samples; stores the design, the simulated outputs, the last zp-outputs and the result files."""
    import gp_emu_uqsa.design_inputs as d
    out = {}
    with tempfile.TemporaryDirectory() as tmp, RL.cwd(tmp), RL.quiet():
        write_noisefit_case(d)
        np.random.seed(11)
        gn.noisefit("config-data", "config-noise", stopat=2, olhcmult=10, samples=50)
        for fn in ("INPUTS", "OUTPUTS", "zp-outputs", "noise-inputs", "noise-outputs"):
            out[fn.replace("-", "_")] = np.loadtxt(fn)
        for fn in sorted(os.listdir(".")):
            if fn.startswith("beliefs-") and fn.endswith("f"):
                out["file_" + fn] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
    return out


def bench_thetas(B_total, d, y, seed=0):
    """bench.py:draw_thetas restated (the multistart draw of _emulatoroptimise.py:206-211 from the auto bounds), so the
    full-size goldens sit exactly where the benchmark evaluates."""
    np.random.seed(seed)
    bounds = [[0.001, 1.0]] * d + [[0.001, float(np.sqrt(np.amax(y) - np.amin(y)))]]
    tb = 2.0 * np.log(np.array(bounds))
    grid = np.zeros((d + 1, B_total))
    for R in range(d + 1):
        grid[R, :] = tb[R, 0] + (tb[R, 1] - tb[R, 0]) * np.random.random_sample(B_total)
    return np.ascontiguousarray(grid.T)


def sha16(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


FULL_MODES = [MODES[2], MODES[3], MODES[0], MODES[5]]   # gp4ml fixed / free nugget, mucm fixed, gp4ml alt-nugget free + r


def gen_llh_fullsize(g, n=4096, d=16, seed=0, nugget=1e-4, which="corner", save=None):
    """Headline shape (config 3): the REAL reference's loglikelihood_* at n=4096, d=16 (about a minute of CPU
    per evaluation).  Stores theta, llh, grad, sigma only; X, y come from synth(n, d, seed) and the reference's
    scaling (x-min)/(max-min), whose SHA-256 is stored so the test can prove it rebuilt the same inputs.
    which="corner": thetas [0] guess 0 of the benchmark's draw (bench.py:draw_thetas; many tiny delta's: A is close
    to the identity), [1] the well-correlated corner of the auto bounds (delta in [0.7, 1.0]: cond(A) ~ 1e7, the
    ill-conditioned end), [2] guess 1 of the benchmark's draw.
    which="mid": two thetas with delta in [0.15, 0.6] (strong but not extreme correlation) for the two gp4ml modes
    and MUCM with a free nugget.
    which="exact": tv_config "8 0 0", so that the training set is all 4096 points (with "10 0 0" the reference's
    split keeps the first 10 * (n // 10) = 4090 points when noV = 0 -- _emulatorclasses.py:507-520 -- which is what the
    "corner" and "mid" files hold: n_train = 4090, padded to 4096 on the device); gp4ml, MUCM and the alt-nugget kernel
    with an r vector at the ill-conditioned corner and one mid-range theta.
    `save(out)` is called after every mode so that an interrupted run keeps what it has."""
    import time
    X, y, _ = synth(n, d, seed)
    out = {"n": n, "d": d, "seed": seed, "nugget_belief": nugget}
    bt = bench_thetas(256, d, y)
    rng = np.random.default_rng(500 + seed)
    ill = 0.7 + 0.3 * rng.random(d)
    mids = [0.15 + 0.45 * rng.random(d) for _ in range(2)]
    if which == "corner":
        modes = FULL_MODES
        points = [(np.exp(bt[0][:d] / 2.0), float(np.exp(bt[0][d] / 2.0)), 3e-3), (ill, 0.9, 2e-4),
                  (np.exp(bt[1][:d] / 2.0), float(np.exp(bt[1][d] / 2.0)), 8e-3)]
    elif which == "mid":
        modes = [MODES[2], MODES[1], MODES[5]]
        points = [(mids[0], 0.7, 1e-3), (mids[1], 1.3, 5e-3)]
    else:
        modes = [MODES[2], MODES[0], MODES[5]]
        points = [(ill, 0.9, 2e-4), (mids[0], 0.7, 1e-3)]
    with tempfile.TemporaryDirectory() as tmp:
        for tag, mucm, alt, fix in modes:
            E = build(g, tmp, X, y, mucm, alt, fix, nugget, "f_" + tag, tv_config="8 0 0" if which == "exact" else "10 0 0")
            out["n_train"] = E.training.inputs.shape[0]
            out["X_sha16"] = sha16(E.training.inputs)
            out["H_sha16"] = sha16(E.training.H)
            if alt == "T":
                r = 0.01 + 0.02 * np.random.default_rng(501 + seed).random(E.training.inputs.shape[0])
                with RL.quiet():
                    E.training.set_r(r)
                out[tag + "_r"] = r
            thetas, llhs, grads, sig = [], [], [], []
            for t, (hp_d, hp_s, hp_n) in enumerate(points):
                hp = list(hp_d)
                if fix == "F":
                    hp.append(hp_n)
                if mucm == "F":
                    hp.append(hp_s)
                theta = E.K.transform(np.array(hp))
                t0 = time.time()
                with RL.quiet():
                    res = (E.opt_T.loglikelihood_mucm if mucm == "T" else E.opt_T.loglikelihood_gp4ml)(theta.copy())
                print(tag, t, "%.1f s" % (time.time() - t0), None if res is None else res[0], flush=True)
                assert res is not None
                thetas.append(theta); llhs.append(res[0]); grads.append(res[1].copy()); sig.append(float(E.par.sigma))
            out[tag + "_theta"] = np.array(thetas)
            out[tag + "_llh"] = np.array(llhs)
            out[tag + "_grad"] = np.array(grads)
            out[tag + "_sigma"] = np.array(sig)
            del E
            if save is not None:
                save(out)
    return out


def gen_sens_fullsize(g, s, n=2000, d=8, seed=0, points=100):
    """Config 5's size: uncertainty / sensitivity / main_effect(points=100) / totaleffectvariance of the REAL
    reference at n=2000, d=8 (m=0.5, v=0.02 as SURVEY 8(d) config 5; delta 0.5, sigma 1, nugget 1e-4).  Run once
    (tens of minutes of Python loops).  Stores the scalars/curves only; X, y from synth(n, d, seed)."""
    import time
    X, y, _ = synth(n, d, seed)
    out = {"n": n, "d": d, "seed": seed}
    with tempfile.TemporaryDirectory() as tmp:
        E = build(g, tmp, X, y, "F", "F", "T", 1e-4, "sensfull")
        delta = np.full(d, 0.5)
        sigma = 1.0
        E.par.delta = delta.copy(); E.K.d = E.par.delta; E.K.n = E.par.nugget
        E.par.sigma = sigma
        with RL.quiet():
            E.training.remake()
            E.opt_T.optimalbeta()
            m = [0.5] * d
            v = [0.02] * d
            S = s.setup(E, list(m), list(v))
        for name, fn in (("uncertainty", S.uncertainty), ("sensitivity", S.sensitivity),
                         ("main_effect", lambda: S.main_effect(plot=False, points=points)),
                         ("totaleffectvariance", S.totaleffectvariance)):
            t0 = time.time()
            with RL.quiet():
                fn()
            out["seconds_" + name] = time.time() - t0
            print(name, "%.1f s" % out["seconds_" + name], flush=True)
        out.update(X_sha16=sha16(E.training.inputs), delta=delta, sigma=sigma, nugget=float(E.par.nugget),
                   beta=np.array(E.par.beta), m=np.array(m), v=np.array(v),
                   input_range=np.array(E.all_data.input_range), uE=S.uE, uV=S.uV, uEV=S.uEV,
                   senseindex=S.senseindex, effect=S.effect, mean_effect=S.mean_effect, EVTw=S.EVTw)
    return out


class _Swallow:
    """Drawing stub: any attribute, call, index or arithmetic on it yields the stub (or 1.0), so the reference's
    matplotlib calls after the arithmetic of interaction_effect (:385-401) run through without a display."""
    def __getattr__(self, name): return self
    def __call__(self, *a, **k): return self
    def __getitem__(self, k): return self
    def __sub__(self, o): return 1.0
    __rsub__ = __truediv__ = __rtruediv__ = __mul__ = __rmul__ = __sub__
    def __abs__(self): return 1.0


def gen_writers(g, s):
    """Byte-level goldens of the on-disk formats and interaction_effect (f4) for a FIXED-hyper-parameter
    emulator (no optimiser in the way): the reference's own bytes of beliefs-N[f], inputs/outputs-oK-N[f], sense_file,
    and the interaction matrix, together with every number that went into them."""
    import sys as _sys
    n, d = 40, 3
    X, y, _ = synth(n, d, 12)
    out = {"X_raw": X, "y": y}
    with tempfile.TemporaryDirectory() as tmp, RL.cwd(tmp):
        with RL.quiet():
            cfg = RL.write_emulator_files(tmp, X, y, mucm="F", fix_nugget="T", alt_nugget="F", nugget=1e-3, name="wr",
                                          tv_config="10 0 2")
            E = g.setup(cfg, datashuffle=False, scaleinputs=True)
            E.par.delta = np.array([0.41, 0.73, 0.58]); E.K.d = E.par.delta; E.K.n = E.par.nugget
            E.par.sigma = 0.9375
            E.training.remake(); E.validation.remake()
            E.opt_T.optimalbeta()
            E.post.remake()
            for final in (False, True):
                E.beliefs.final_beliefs(E, final)
                E.post.final_design_points(E, final)
        for fn in sorted(os.listdir(".")):
            if fn.startswith("wr_") and "-" in fn:
                out["file_" + fn] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
        out.update(delta=np.array(E.par.delta), sigma=float(E.par.sigma), nugget=float(E.par.nugget), beta=np.array(E.par.beta),
                   no_of_trains=int(E.tv_conf.no_of_trains))
        with RL.quiet():
            m, v = [0.5, 0.45, 0.55], [0.02, 0.03, 0.025]
            S = s.setup(E, m, v)
            S.uncertainty(); S.sensitivity(); S.main_effect(plot=False, points=11); S.totaleffectvariance()
            S.to_file("sense_file")
            out["file_sense_file"] = np.frombuffer(open("sense_file", "rb").read(), dtype=np.uint8)
            out.update(m=np.array(m), v=np.array(v), uE=S.uE, uV=S.uV, uEV=S.uEV, senseindex=np.array(S.senseindex),
                       EVTw=np.array(S.EVTw), effect=np.array(S.effect), mean_effect=np.array(S.mean_effect))
            # interaction_effect with the drawing swallowed
            mod = _sys.modules[type(S).__module__]
            old_plt = getattr(mod, "plt", None)
            mod.plt = _Swallow()
            try:
                S.interaction_effect(0, 2, points=7)
            finally:
                mod.plt = old_plt
            out["interaction_0_2"] = np.array(S.interaction)
            out["interaction_mean_effect"] = np.array(S.mean_effect)
    return out


def gen_config4(g, h, n=2000, d=8, m=3000):
    """Config 4's shape with the REAL reference: two n = 2000, d = 8 emulators (outputs sin(Xw)+0.1 sum x^2 and cos(X w2) as in
    bench.py) with fixed hyper-parameters (delta 0.5, sigma 1, nugget 1e-4, beta = optimalbeta), checkpointed with the
    reference's own writers and rebuilt from those files; g.posterior (mean, diag of the full covariance, m = 1000 per call)
    and history_match.nonimp_data on m points of the 10-level tensor grid.  Stores the checkpoint files (bytes), the points,
    mean / variance per emulator and the kept rows."""
    X, y, _ = synth(n, d, 0)
    y2 = np.cos(X @ np.random.default_rng(1).normal(size=d))
    rng = np.random.default_rng(44)
    out = {}
    with tempfile.TemporaryDirectory() as tmp, RL.cwd(tmp), RL.quiet():
        emuls = []
        for o, yy in enumerate((y, y2)):
            cfg = RL.write_emulator_files(tmp, X, yy, mucm="F", fix_nugget="T", alt_nugget="F", nugget=1e-4, name="c4_%d" % o)
            E = g.setup(cfg, datashuffle=False, scaleinputs=True)
            E.par.delta = np.full(d, 0.5); E.K.d = E.par.delta; E.K.n = E.par.nugget
            E.par.sigma = 1.0
            E.training.remake()
            E.opt_T.optimalbeta()
            E.beliefs.final_beliefs(E, True)
            E.post.final_design_points(E, True)
            with open("c4_%d_config_r" % o, "w") as f:
                f.write("beliefs c4_%d_beliefs-0f\ninputs c4_%d_input-o0-0f\noutputs c4_%d_output-o0-0f\n" % (o, o, o))
                f.write("tv_config 10 0 0\ndelta_bounds [ ]\nsigma_bounds [ ]\nnugget_bounds [ ]\ntries 1\nconstraints bounds\n")
            for fn in ("c4_%d_config_r" % o, "c4_%d_beliefs-0f" % o, "c4_%d_input-o0-0f" % o, "c4_%d_output-o0-0f" % o):
                out["file_" + fn] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
            emuls.append(g.setup("c4_%d_config_r" % o, datashuffle=False, scaleinputs=True))
        # points of the 10-level grid in scaled units: random flat indices + a run of consecutive ones
        idx = np.sort(np.concatenate([rng.choice(10 ** 8, size=m - 600, replace=False), np.arange(36999700, 36999700 + 600)]))
        P = np.empty((m, d))
        t = idx.copy()
        for k in range(d - 1, -1, -1):
            P[:, k] = (t % 10 + 0.5) / 10.0
            t = t // 10
        out["grid_index"] = idx
        out["P_scaled"] = P
        for o, E2 in enumerate(emuls):
            mu, vd = np.empty(m), np.empty(m)
            for c in range(0, m, 1000):
                mean, var = g.posterior(E2, P[c:c + 1000].copy())
                mu[c:c + 1000], vd[c:c + 1000] = mean, np.diag(var)
            out["mean%d" % o], out["var%d" % o] = mu, vd
            out["beta%d" % o] = np.array(E2.par.beta)
        # the same points in the data files' units for nonimp_data (it scales them with input_minmax, _hmutilfunctions.py:126-139)
        mm = np.array(emuls[0].beliefs.input_minmax)
        pts = P * (mm[:, 1] - mm[:, 0]) + mm[:, 0]
        np.savetxt("sim_in", pts, fmt="%.17g")
        np.savetxt("sim_out", np.column_stack([pts[:, 0], pts[:, 1]]), fmt="%.17g")
        out["sim_in"] = pts
        zs = [float(np.median(y)), float(np.median(y2))]
        ve = [1e-2, 1e-2]
        out.update(zs=np.array(zs), var_extra=np.array(ve), cm=3.0)
        cnt = h.nonimp_data(emuls, zs, 3.0, ve, ["sim_in", "sim_out"], maxno=1)
        out["nonimp_count"] = cnt
        out["nonimp_in"] = np.atleast_2d(np.loadtxt("nonimp_sim_in"))
    return out


def main():
    assert RL.available(), "reference not present"
    g, h, s, gn = RL.load()
    save = lambda name, d: (np.savez_compressed(os.path.join(HERE, name), **d), print("wrote", name))
    if len(sys.argv) > 1 and sys.argv[1] == "fullsize_llh":
        gen_llh_fullsize(g, which="corner", save=lambda o: save("llh_n4096_d16.npz", o))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "fullsize_llh_exact":
        gen_llh_fullsize(g, which="exact", save=lambda o: save("llh_n4096_d16_exact.npz", o))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "fullsize_llh_mid":
        gen_llh_fullsize(g, which="mid", save=lambda o: save("llh_n4096_d16_mid.npz", o))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "fullsize_sens":
        save("sens_n2000_d8.npz", gen_sens_fullsize(g, s))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "config4":
        save("config4_n2000_d8.npz", gen_config4(g, h))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "writers":
        save("writers_n40_d3.npz", gen_writers(g, s))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sens":
        save("sens_n60_d3.npz", gen_sens(g, s, 60, 3, 7))
        save("sens_n150_d4.npz", gen_sens(g, s, 150, 4, 8))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "recon":
        save("toysim_recon.npz", gen_toysim_recon(g))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "toysimlog":
        save("toysim_log.npz", gen_toysim_log(g))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "noisefit":
        save("noisefit_n150.npz", gen_noisefit(gn))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "hmapi":
        save("hmapi_n100_d3.npz", gen_hm_api(g, h, 100, 6))
        return
    save("llh_n60_d2.npz", gen_llh(g, 60, 2, 1, 3, nugget=1e-2))
    save("llh_n200_d4.npz", gen_llh(g, 200, 4, 2, 3, with_r=True))
    save("llh_n500_d8.npz", gen_llh(g, 500, 8, 3, 2))
    save("post_n200_d4.npz", gen_posterior(g, 200, 4, 4, 150))
    save("post_n500_d8.npz", gen_posterior(g, 500, 8, 5, 100))
    save("toysim.npz", gen_toysim(g))
    save("hm_n100_d3.npz", gen_hm(g, h, 100, 6))


if __name__ == "__main__":
    main()
