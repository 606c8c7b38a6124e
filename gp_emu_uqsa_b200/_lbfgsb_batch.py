"""Lock-step batched multistart L-BFGS-B.

The reference runs ``scipy.optimize.minimize(..., method='L-BFGS-B', jac=True)`` once per initial
guess, sequentially (``_emulatoroptimise.py:227-247``).  Here every start still runs SciPy's own
L-BFGS-B (same optimiser arithmetic, same options), each in its own thread, but a start that
needs ``(f, g)`` parks its request; when all live starts are parked the coordinator evaluates the
whole round with ONE batched device call and releases them.  Starts finish at different rounds,
so the batch shrinks (ragged batch); results do not depend on thread timing because an
item's value does not depend on which other items share its batch.

A start whose evaluation reports a numerical failure (non-PD covariance) is abandoned, the
counterpart of the reference's ``return None`` -> ``TypeError`` -> "Trying next guess..." path.
"""
import threading

import numpy as np
from scipy.optimize import minimize


class _EvalFailed(Exception):
    pass


def minimize_batch(eval_batch, x0s, bounds=None, shard=None):
    """eval_batch(X [b,p]) -> (f [b], g [b,p], ok [b] bool).  x0s: [B,p].  Returns a list of
    scipy OptimizeResult (or None for abandoned starts), in start order, plus the number of
    evaluation rounds and of evaluations."""
    x0s = np.asarray(x0s, dtype=float)
    B = x0s.shape[0]
    cond = threading.Condition()
    pending, ready = {}, {}
    state = {"live": B}
    results = [None] * B

    def worker(i):
        def fun(x):
            with cond:
                if state.get("abort"):
                    raise _EvalFailed()
                pending[i] = np.array(x, dtype=float, copy=True)
                cond.notify_all()
                while i not in ready:
                    cond.wait()
                f, g, ok = ready.pop(i)
            if not ok:
                raise _EvalFailed()
            return f, g

        try:
            kw = {} if bounds is None else {"bounds": bounds}
            results[i] = minimize(fun, list(x0s[i]), method="L-BFGS-B", jac=True, **kw)
        except _EvalFailed:
            results[i] = None
        finally:
            with cond:
                state["live"] -= 1
                cond.notify_all()

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(B)]
    for t in threads:
        t.start()
    rounds = evals = 0
    while True:
        with cond:
            while state["live"] > 0 and len(pending) < state["live"]:
                cond.wait()
            if state["live"] == 0:
                break
            idx = sorted(pending)
            X = np.stack([pending.pop(i) for i in idx])
        try:
            f, g, ok = eval_batch(X)
        except BaseException:
            # a device/library error: release every parked start as "failed" so no thread is left waiting,
            # then let the error reach the caller
            with cond:
                for i in idx:
                    ready[i] = (float("nan"), np.zeros(x0s.shape[1]), False)
                state["abort"] = True
                cond.notify_all()
            for t in threads:
                t.join()
            raise
        rounds += 1
        evals += len(idx)
        with cond:
            for k, i in enumerate(idx):
                ready[i] = (float(f[k]), np.array(g[k], dtype=float, copy=True), bool(ok[k]))
            cond.notify_all()
    for t in threads:
        t.join()
    return results, rounds, evals
