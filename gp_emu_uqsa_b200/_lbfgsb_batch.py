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

Two drivers with identical results (tests/test_host_cpu.py compares both with ``scipy.optimize.minimize``
start by start, bit for bit):

* ``_minimize_batch_rc`` -- reverse communication: SciPy's compiled L-BFGS-B step routine
  (``scipy.optimize._lbfgsb.setulb``, the very call ``minimize`` loops over) is advanced start by start
  until it asks for ``(f, g)``; no threads, none of ``minimize``'s per-call Python wrapping (about 0.25 ms
  per evaluation, which dominates a 64-start round at n = 1000).  Uses a private SciPy interface, so it is
  only selected when that interface has the signature this code was written against (SciPy 1.15+).
* ``_minimize_batch_threads`` -- public API only: one thread per start inside ``minimize``, parked on an
  event while the coordinator evaluates the round.
"""
import threading

import numpy as np
from scipy.optimize import OptimizeResult, minimize


class _EvalFailed(Exception):
    pass


class _Slot:
    __slots__ = ("x", "out", "go")

    def __init__(self):
        self.x = None                      # parked request of this start (None: not waiting)
        self.out = None                    # (f, g, ok) of the round it took part in
        self.go = threading.Event()


def _setulb_interface():
    """SciPy's step routine and integer dtype if the private interface looks like the one of SciPy 1.15-1.18
    (17 positional arguments ending in maxls, ln_task); None otherwise."""
    try:
        from scipy.optimize import _lbfgsb
        from scipy.optimize._lbfgsb_py import HAS_ILP64
        from scipy.optimize._constraints import old_bound_to_new
        doc = (_lbfgsb.setulb.__doc__ or "").splitlines()[0]
        if "setulb(m, x, l, u, nbd, f, g, factr, pgtol, wa, iwa, task, lsave, isave, dsave, maxls, ln_task)" not in doc:
            return None
        return _lbfgsb.setulb, (np.int64 if HAS_ILP64 else np.int32), old_bound_to_new
    except Exception:
        return None


_STATUS = {0: "START", 1: "NEW_X", 2: "RESTART", 3: "FG", 4: "CONVERGENCE", 5: "STOP", 6: "WARNING", 7: "ERROR", 8: "ABNORMAL"}


class _RcStart:
    """State of one start for the reverse-communication driver: exactly the arrays and counters of
    scipy.optimize._lbfgsb_py._minimize_lbfgsb (defaults of minimize(method='L-BFGS-B'))."""
    M, FTOL, GTOL, MAXFUN, MAXITER, MAXLS = 10, 2.2204460492503131e-09, 1e-5, 15000, 15000, 20

    def __init__(self, x0, bounds, iface):
        self.setulb, idt, old_bound_to_new = iface
        x0 = np.asarray(x0, dtype=float).ravel()
        n = x0.size
        self.nbd = np.zeros(n, dtype=idt)
        self.lo, self.hi = np.zeros(n), np.zeros(n)
        if bounds is not None:
            if len(bounds) != n:
                raise ValueError('length of x0 != length of bounds')
            nb = np.array(old_bound_to_new(bounds))
            if (nb[0] > nb[1]).any():
                raise ValueError("LBFGSB - one of the lower bounds is greater than an upper bound.")
            x0 = np.clip(x0, nb[0], nb[1])
            code = {(False, False): 0, (True, False): 1, (True, True): 2, (False, True): 3}
            for i in range(n):
                has_l, has_u = not np.isinf(nb[0, i]), not np.isinf(nb[1, i])
                if has_l:
                    self.lo[i] = nb[0, i]
                if has_u:
                    self.hi[i] = nb[1, i]
                self.nbd[i] = code[has_l, has_u]
        m = self.M
        self.x = np.array(x0, dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n)
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m)
        self.iwa = np.zeros(3 * n, dtype=idt)
        self.task = np.zeros(2, dtype=idt)
        self.ln_task = np.zeros(2, dtype=idt)
        self.lsave = np.zeros(4, dtype=idt)
        self.isave = np.zeros(44, dtype=idt)
        self.dsave = np.zeros(29)
        self.factr = self.FTOL / np.finfo(float).eps
        self.nit = self.nfev = 0
        self.last_x = None                 # SciPy's ScalarFunction re-uses the value at an unchanged x

    def advance(self):
        """Run the optimiser until it wants (f, g) at self.x (-> True) or terminates (-> False)."""
        while True:
            self.setulb(self.M, self.x, self.lo, self.hi, self.nbd, self.f, self.g, self.factr, self.GTOL, self.wa, self.iwa,
                        self.task, self.lsave, self.isave, self.dsave, self.MAXLS, self.ln_task)
            t = self.task[0]
            if t == 3:
                if self.last_x is not None and np.array_equal(self.x, self.last_x):
                    continue                # cached value: f and g are already those of this x
                return True
            if t == 1:
                self.nit += 1
                if self.nit >= self.MAXITER:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > self.MAXFUN:
                    self.task[0], self.task[1] = 5, 502
            else:
                return False

    def give(self, f, g):
        self.f = float(f)
        self.g = np.array(g, dtype=np.float64, copy=True)
        self.last_x = self.x.copy()
        self.nfev += 1

    def result(self):
        if self.task[0] == 4:
            warnflag = 0
        elif self.nfev > self.MAXFUN or self.nit >= self.MAXITER:
            warnflag = 1
        else:
            warnflag = 2
        return OptimizeResult(fun=self.f, jac=self.g, nfev=self.nfev, njev=self.nfev, nit=self.nit, status=warnflag,
                              message=_STATUS.get(int(self.task[0]), "?"), x=self.x, success=(warnflag == 0))


def _minimize_batch_rc(eval_batch, x0s, bounds, iface):
    starts = [_RcStart(x0, bounds, iface) for x0 in x0s]
    results = [None] * len(starts)
    live = list(range(len(starts)))
    rounds = evals = 0
    while live:
        need = []
        for i in live:
            if starts[i].advance():
                need.append(i)
            else:
                results[i] = starts[i].result()
        if not need:
            break
        X = np.stack([starts[i].x for i in need])
        f, g, ok = eval_batch(X)
        rounds += 1
        evals += len(need)
        live = []
        for k, i in enumerate(need):
            if ok[k]:
                starts[i].give(f[k], g[k])
                live.append(i)
            else:
                results[i] = None          # non-PD: the reference's "Trying next guess..." path
    return results, rounds, evals


def minimize_batch(eval_batch, x0s, bounds=None, driver=None):
    """eval_batch(X [b,p]) -> (f [b], g [b,p], ok [b] bool).  x0s: [B,p].  Returns a list of
    scipy OptimizeResult (or None for abandoned starts), in start order, plus the number of
    evaluation rounds and of evaluations.  driver: None (reverse communication when SciPy's step routine
    has the expected interface, else threads), "rc" or "threads"."""
    x0s = np.asarray(x0s, dtype=float)
    iface = _setulb_interface() if driver in (None, "rc") else None
    if driver == "rc" and iface is None:
        raise RuntimeError("scipy.optimize._lbfgsb.setulb does not have the expected interface")
    if iface is not None:
        return _minimize_batch_rc(eval_batch, x0s, bounds, iface)
    return _minimize_batch_threads(eval_batch, x0s, bounds)


def _minimize_batch_threads(eval_batch, x0s, bounds=None):
    B = x0s.shape[0]
    slots = [_Slot() for _ in range(B)]
    lock = threading.Lock()
    round_ready = threading.Event()        # set when every live start is parked (or none is left)
    state = {"live": B, "parked": 0, "abort": False}
    results = [None] * B

    def _maybe_wake():                     # call with the lock held
        if state["parked"] >= state["live"]:
            round_ready.set()

    def worker(i):
        slot = slots[i]

        def fun(x):
            if state["abort"]:
                raise _EvalFailed()
            slot.x = np.array(x, dtype=float, copy=True)
            with lock:
                state["parked"] += 1
                _maybe_wake()
            slot.go.wait()
            slot.go.clear()
            f, g, ok = slot.out
            if not ok:
                raise _EvalFailed()
            return f, g

        try:
            kw = {} if bounds is None else {"bounds": bounds}
            results[i] = minimize(fun, list(x0s[i]), method="L-BFGS-B", jac=True, **kw)
        except _EvalFailed:
            results[i] = None
        finally:
            with lock:
                state["live"] -= 1
                _maybe_wake()

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(B)]
    for t in threads:
        t.start()
    rounds = evals = 0
    while True:
        round_ready.wait()
        with lock:
            round_ready.clear()
            if state["live"] == 0:
                break
            if state["parked"] < state["live"]:        # a start finished and another is still computing its step
                continue
            idx = [i for i in range(B) if slots[i].x is not None]
            state["parked"] = 0
        X = np.stack([slots[i].x for i in idx])
        for i in idx:
            slots[i].x = None
        try:
            f, g, ok = eval_batch(X)
        except BaseException:
            # a device/library error: release every parked start as "failed" so no thread is left waiting,
            # then let the error reach the caller
            state["abort"] = True
            for i in idx:
                slots[i].out = (float("nan"), np.zeros(x0s.shape[1]), False)
                slots[i].go.set()
            for t in threads:
                t.join()
            raise
        rounds += 1
        evals += len(idx)
        for k, i in enumerate(idx):
            slots[i].out = (float(f[k]), np.array(g[k], dtype=float, copy=True), bool(ok[k]))
            slots[i].go.set()
    for t in threads:
        t.join()
    return results, rounds, evals
