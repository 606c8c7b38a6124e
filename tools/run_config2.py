"""Config 2: synthetic 8-input emulator, n = 1000, 64 multistart llh+gradient optimisation
through the reference-facing API (g.setup + Optimize.llh_optimize), all four {mucm, gp4ml} x {fix_nugget T, F}.
Prints wall time, evaluation rounds, evaluations and evals/s per mode.
    python tools/run_config2.py [n] [d] [tries]"""
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import gp_emu_uqsa_b200 as g
from oracle import ref_loader as RL          # only its text-file writer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
tries = int(sys.argv[3]) if len(sys.argv) > 3 else 64
rng = np.random.default_rng(0)
X = rng.random((n, d)); w = rng.normal(size=d)
y = np.sin(X @ w) + 0.1 * (X ** 2).sum(1)
with tempfile.TemporaryDirectory() as tmp:
    os.chdir(tmp)
    for mucm, fix in (("F", "T"), ("F", "F"), ("T", "T"), ("T", "F")):
        with contextlib.redirect_stdout(io.StringIO()):
            cfg = RL.write_emulator_files(tmp, X, y, mucm=mucm, fix_nugget=fix, alt_nugget="F", nugget=1e-4,
                                          name="c2_%s%s" % (mucm, fix), tries=tries, constraints="bounds")
            E = g.setup(cfg, datashuffle=False, scaleinputs=True)
            E.opt_T.llh_optimize()                      # warm-up: workspace allocation, graph capture
            np.random.seed(0)
            t0 = time.perf_counter()
            E.opt_T.llh_optimize()
            dt = time.perf_counter() - t0
        o = E.opt_T
        ok = int((o.last_table[:, 0] == 1.0).sum())
        print("mucm=%s fix_nugget=%s: %d starts (%d ok) in %.3f s, %d rounds, %d evaluations -> %.0f evals/s; best llh %.6f (guess %d), "
              "delta[:3]=%s sigma=%.5f" % (mucm, fix, tries, ok, dt, o.last_rounds, o.last_evals, o.last_evals / dt, -o.best_llh, o.best_guess,
                                           np.round(E.par.delta[:3], 4), E.par.sigma), flush=True)
