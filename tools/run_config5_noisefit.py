"""Config 5, second half: the noisefit2D example's generator (examples/noisefit2D/emulator.py:12-32) scaled to
n = 2000, noisefit(stopat=2, samples=200).  Prints the wall time and the fitted-vs-true noise level.
    python tools/run_config5_noisefit.py [n] [stopat]"""
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import gp_emu_uqsa_b200.noise_fit as gn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
stopat = int(sys.argv[2]) if len(sys.argv) > 2 else 2


def mfunc(x):
    return 3.0 * x[:, 0] ** 3 + np.exp(np.cos(10.0 * x[:, 1]) * np.cos(5.0 * x[:, 0]) ** 2)


def nfunc(x):
    return np.abs(0.500 * (x[:, 1] * (np.cos(6 * x[:, 0]) ** 2 + 0.1)))


with tempfile.TemporaryDirectory() as tmp:
    os.chdir(tmp)
    np.random.seed(1)
    x = np.random.rand(n, 2)
    np.savetxt("INPUTS", x)
    np.savetxt("OUTPUTS", mfunc(x) + nfunc(x) * np.random.randn(n))
    for name, outputs, alt, cons, db, sb, nb in (("data", "OUTPUTS", "T", "none", "[[0.05,10.0],[0.05,10.00]]", "[[0.1,3.0]]", "[[0.001,1.05]]"),
                                                 ("noise", "zp-outputs", "F", "bounds", "[[0.05,1.0],[0.05,10.00]]", "[[0.001,10.0]]", "[[0.0001,1.0]]")):
        with open("config-" + name, "w") as f:
            f.write("beliefs beliefs-%s\ninputs INPUTS\noutputs %s\ntv_config 10 0 0\ndelta_bounds %s\nsigma_bounds %s\n"
                    "nugget_bounds %s\ntries 3\nconstraints %s\n" % (name, outputs, db, sb, nb, cons))
        with open("beliefs-" + name, "w") as f:
            f.write("active all\noutput 0\nbasis_str 1.0\nbasis_inf NA\nbeta 1.0\ndelta 1.0 1.0\nsigma 1.0\nnugget 0.00001\n"
                    "fix_nugget F\nalt_nugget %s\nmucm F\n" % alt)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        gn.noisefit("config-data", "config-noise", stopat=stopat, olhcmult=100, samples=200)
    dt = time.perf_counter() - t0
    xin, out = np.loadtxt("noise-inputs"), np.loadtxt("noise-outputs")
    true = nfunc(xin)
    print("noisefit n=%d stopat=%d samples=200: %.2f s; median |fit - true| / median true = %.3f; corr(fit, true) = %.3f"
          % (n, stopat, dt, np.median(np.abs(out[:, 0] - true)) / np.median(true), np.corrcoef(out[:, 0], true)[0, 1]))
