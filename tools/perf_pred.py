"""Quick look at gpe_predict_grid throughput (development aid): python tools/perf_pred.py [n] [d] [points]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gp_emu_uqsa_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m = int(float(sys.argv[3])) if len(sys.argv) > 3 else 1 << 22
rng = np.random.default_rng(0)
X = rng.random((n, d)); w = rng.normal(size=d)
y = np.sin(X @ w) + 0.1 * (X ** 2).sum(1)
dev = _lib.Device(0)
dev.set_training(X, y, np.column_stack([np.ones(n), X]))
dev.set_basis(list(range(d)), [1] * d)
dev.fit_state(np.full(d, 0.5), 1e-4, 1.0, 0)
levels = np.full(d, 10, dtype=np.int32)
mean = torch.empty(m, dtype=torch.float64, device="cuda"); var = torch.empty(m, dtype=torch.float64, device="cuda")
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev.predict_grid(levels, np.zeros(d), np.ones(d), it * m, m, out=(mean, var))
    dt = time.perf_counter() - t0
    print("chunk=%s: %.1f ms, %.3f Mpreds/s, %.2f TFLOP/s algorithmic" % (os.environ.get("GPE_PRED_CHUNK", "default"), dt * 1e3, m / dt * 1e-6,
                                                                          m / dt * (n * n + 2 * n * (2 * d + 4)) * 1e-12), flush=True)
