"""Config 5 timing (development aid): Sensitivity.uncertainty / sensitivity / main_effect at n = 2000, d = 8
through the reference-facing API.   python tools/perf_sens.py [n] [d]"""
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
import gp_emu_uqsa_b200.sensitivity as s
from oracle import ref_loader as RL          # only its text-file writer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = np.random.default_rng(0)
X = rng.random((n, d)); w = rng.normal(size=d)
y = np.sin(X @ w) + 0.1 * (X ** 2).sum(1)
with tempfile.TemporaryDirectory() as tmp:
    os.chdir(tmp)
    with contextlib.redirect_stdout(io.StringIO()):
        cfg = RL.write_emulator_files(tmp, X, y, mucm="F", fix_nugget="T", alt_nugget="F", nugget=1e-4, name="c5", delta=[0.5] * d)
        E = g.setup(cfg, datashuffle=False, scaleinputs=True)
        E.training.remake(); E.opt_T.optimalbeta()
    for rep in range(2):
        t = {}
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter(); S = s.setup(E, [0.5] * d, [0.02] * d); t["setup"] = time.perf_counter() - t0
            t0 = time.perf_counter(); S.uncertainty(); t["uncertainty"] = time.perf_counter() - t0
            t0 = time.perf_counter(); S.sensitivity(); t["sensitivity"] = time.perf_counter() - t0
            t0 = time.perf_counter(); S.main_effect(points=100); t["main_effect"] = time.perf_counter() - t0
        print("n=%d d=%d:" % (n, d), {k: round(v, 4) for k, v in t.items()}, "uE=%.6f uEV=%.6f sum(S)/EV=%.4f" % (S.uE, S.uEV, S.senseindex.sum() / S.uEV), flush=True)
