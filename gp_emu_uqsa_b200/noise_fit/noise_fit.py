"""Heteroscedastic noise fit (most-likely heteroscedastic GP, Kersting et al.) with the reference's
call surface (gp_emu_uqsa/noise_fit/noise_fit.py:38).  A driver loop only: every numerical step it
takes -- g.train on the data and noise emulators, the full posterior covariance at the training
points, its Cholesky factor for the posterior samples, the noise-GP predictions -- is one of the
device paths of this package (gpe_llh_grad_batch, gpe_predict_fullcov, gpe_potrf, gpe_predict)."""
import numpy as np

from .. import _emulatorclasses as _emuc
from .. import _lib
from .. import design_inputs as _gd
from .. import emulatorfunctions as g

__all__ = ["noisefit"]


def _read_file(ifile):
    print("*** Reading file:", ifile, "***")
    table = {}
    try:
        with open(ifile, 'r') as f:
            for line in f:
                key, val = line.split(' ', 1)
                table[key] = val.strip()
    except OSError:
        print("ERROR: Problem reading file.")
        raise SystemExit(1)
    return table


def _configs_consistent(data, noise):
    """The reference's pre-flight checks (:53-70); returns False (after the same warning) when one fails."""
    datac, noisec = _read_file(data), _read_file(noise)
    datab, noiseb = _read_file(datac["beliefs"]), _read_file(noisec["beliefs"])
    checks = (
        (datac["inputs"] != noisec["inputs"], "\nWARNING: different inputs files in config files. Exiting."),
        (datab["alt_nugget"] == 'F', "\nWARNING: data beliefs must have alt_nugget T. Exiting."),
        (datab["fix_nugget"] == 'T' or noiseb["fix_nugget"] == 'T', "\nWARNING: data and noise beliefs need fix_nugget F. Exiting."),
        (datac["tv_config"] != noisec["tv_config"], "\nWARNING: different tv_config in config files. Exiting."),
        (noisec["outputs"] != "zp-outputs", "\nWARNING: config outputs file must be 'zp-outputs'. Exiting."),
    )
    for bad, msg in checks:
        if bad:
            print(msg)
            return False
    return True


def _log_noise_estimate(GD, pts, targets, samples):
    """log of the mean over `samples` posterior draws of 0.5 (t - t_j)^2 (reference :128-138): the
    draws are mean + L u with L = chol(V) from the device and u from the global NumPy RNG, consumed
    sample by sample in the reference's order (randn(n) per sample == rows of randn(samples, n))."""
    if targets.size == 0:
        return np.zeros(0)
    post = _emuc.Posterior(pts, GD.training, GD.par, GD.beliefs, GD.K)
    Lf = _lib.scratch_device().cholesky(post.var)
    u = np.random.randn(samples, targets.size)
    draws = post.mean[None, :] + u.dot(Lf.T)
    return np.log((0.5 * (targets[None, :] - draws) ** 2).sum(axis=0) / float(samples))


def _points(E, x, r=None):
    d = _emuc.Data(x, None, E.basis, E.par, E.beliefs, E.K)
    if r is not None:
        d.set_r(r)
        d.make_A(s2=E.par.sigma ** 2, predict=True)
    return d


def _noise_mean(GN, x):
    """exp of the noise-GP posterior mean at x (reference :171-180): the variances r for the data GP."""
    if x.shape[0] == 0:
        return np.zeros(0)
    p = _emuc.Posterior(_points(GN, x), GN.training, GN.par, GN.beliefs, GN.K, diag_only=True)
    return np.exp(p.mean)


def _retrain(E, valsets):
    E.tv_conf.no_of_trains = 0       # same training set again, against the same validation set (:166-168, :186-189)
    E.tv_conf.retrain = 'y'
    g.train(E, no_retrain=valsets)


def noisefit(data, noise, stopat=20, olhcmult=100, samples=200, fileStr=""):
    """Fit one emulator to the data and another to the (log) noise level, alternating `stopat` times;
    write 'noise-inputs' / 'noise-outputs' (noise sigma and its 95% band on an optimised LHC).
    Returns None."""
    if not _configs_consistent(data, noise):
        return None
    GD = g.setup(data, datashuffle=True, scaleinputs=False)
    np.savetxt("zp-outputs", np.zeros(GD.training.outputs.size + GD.validation.outputs.size * GD.tv_conf.noV).T)
    GN = g.setup(noise, datashuffle=True, scaleinputs=False)
    GN.training.inputs = GD.training.inputs          # both emulators must see the same (shuffled) inputs
    GN.validation.inputs = GD.validation.inputs
    GN.training.remake()
    GN.validation.remake()
    if GD.all_data.tv.noV > 1:
        print("\nWARNING: should have 0 or 1 validation sets for noise fitting. Exiting.")
        raise SystemExit(1)
    valsets = GD.all_data.tv.noV != 0

    print("\n****************"
          "\nTRAIN GP ON DATA"
          "\n****************")
    x, t = GD.training.inputs, GD.training.outputs
    xv, tv = GD.validation.inputs, GD.validation.outputs
    g.train(GD, no_retrain=valsets)
    r, rv = None, None
    for count in range(1, int(stopat) + 1):
        print("\n***********************"
              "\nESTIMATING NOISE LEVELS " + str(count) +
              "\n***********************")
        z_prime = _log_noise_estimate(GD, _points(GD, x, r), t, samples)
        np.savetxt('zp-outputs', z_prime)
        z_prime_V = _log_noise_estimate(GD, _points(GD, xv, rv), tv, samples)

        print("\n*****************"
              "\nTRAIN GP ON NOISE " + str(count) +
              "\n*****************")
        GN.training.outputs = z_prime.T
        GN.training.remake()
        GN.validation.outputs = z_prime_V.T
        GN.validation.remake()
        _retrain(GN, valsets)

        print("\n***********************************"
              "\nTRAIN GP ON DATA WITH NOISE FROM GP " + str(count) +
              "\n***********************************")
        r = _noise_mean(GN, x)
        GD.training.set_r(r)
        rv = _noise_mean(GN, xv)
        if rv.size:
            GD.validation.set_r(rv)
        _retrain(GD, valsets)

    print("\nCompleted", count, "fits, stopping here.")
    print("\nGenerating input points to predict noise values at...")
    ndim = x[0].size
    n = ndim * int(olhcmult)
    olhc_range = [[np.amin(col), np.amax(col)] for col in x.T]
    _gd.optLatinHyperCube(ndim, n, int(n), olhc_range, "x_range_input", _criterion=_gd.device_criterion)
    x_range = np.loadtxt("x_range_input").reshape(n, ndim)
    p_plot = _emuc.Posterior(_points(GN, x_range), GN.training, GN.par, GN.beliefs, GN.K, diag_only=True)
    p_plot.interval()
    print("\nSaving results to file...")
    nfileStr = fileStr + "_" if fileStr != "" else fileStr
    np.savetxt(nfileStr + 'noise-inputs', x_range)
    np.savetxt(nfileStr + 'noise-outputs',
               np.transpose([np.sqrt(np.exp(p_plot.mean)), np.sqrt(np.exp(p_plot.LI)), np.sqrt(np.exp(p_plot.UI))]))
    return None
