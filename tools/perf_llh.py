"""Quick wall-clock look at gpe_llh_grad_batch (development aid, not the benchmark)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gp_emu_uqsa_b200 import _lib

n, d, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rng = np.random.default_rng(0)
X = rng.random((n, d)); w = rng.normal(size=d)
y = np.sin(X @ w) + 0.1 * (X ** 2).sum(1)
H = np.column_stack([np.ones(n), X])
dev = _lib.Device(0)
dev.set_training(X, y, H)
hp = np.column_stack([0.3 + 0.7 * rng.random((B, d)), 0.5 + rng.random(B)])
theta = 2 * np.log(hp)
for it in range(reps + 1):
    t0 = time.perf_counter()
    llh, grad, sig, st = dev.llh_grad_batch(theta + 1e-3 * it, 0, fixed_nugget=1e-4)
    dt = time.perf_counter() - t0
    fl = B * (float(n) ** 3 + n * n * (3 * d + (d + 1) + 2 * (d + 1) + 4))
    print("iter %d: %.2f ms  -> %.1f evals/s, %.2f TFLOP/s (algorithmic), status_ok=%s" % (it, dt * 1e3, B / dt, fl / dt * 1e-12, (st == 0).all()), flush=True)
print("launches", dev.launches)
