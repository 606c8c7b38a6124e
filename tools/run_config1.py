"""Config 1: examples/toy-sim as shipped (setup + train + the two shipped plot grids), timed end to end.
    python tools/run_config1.py"""
import contextlib
import io
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import gp_emu_uqsa_b200 as g

src = os.path.join(ROOT, "tests", "golden", "toy-sim")
with tempfile.TemporaryDirectory() as tmp:
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), tmp)
    os.chdir(tmp)
    for rep in range(3):
        with contextlib.redirect_stdout(io.StringIO()):
            np.random.seed(0)
            t0 = time.perf_counter()
            E = g.setup("toy-sim_config")
            t1 = time.perf_counter()
            g.train(E)
            t2 = time.perf_counter()
            g.plot(E, [0], [1], [0.3], "mean")
            g.plot(E, [0, 1], [2], [0.3], "mean")
            t3 = time.perf_counter()
        print("run %d: setup %.3f s, train %.3f s (3 rounds x 10 guesses), two plot grids %.3f s; delta=%s sigma=%.6f"
              % (rep, t1 - t0, t2 - t1, t3 - t2, np.round(E.par.delta, 5), E.par.sigma), flush=True)
