"""Lock-step batching of the L-BFGS-B restarts that refine the best sampled candidates.

The reference runs `n_restarts_optimizer` independent `fmin_l_bfgs_b` searches, each calling the one-point
`gaussian_acquisition_1D` about 22 times, sequentially or in joblib processes (skopt/optimizer/optimizer.py:570-586);
77 % of a Hartmann-6 run is spent there.  Here every search still uses SciPy's L-BFGS-B
driver unchanged (same maxiter=20, bounds, exact gradients), but the searches advance in lock step: each runs in its
own thread, blocks when it needs f(x) and grad f(x), and the coordinator evaluates ALL pending points (all restarts of
all hedge acquisitions) with one batched GPU call.  A search only ever sees values of its own points, and the GPU
returns bit-identical results for a row regardless of the batch it travels in, so the outcome does not depend on
thread scheduling and equals running the searches one after another.
"""
import threading

import numpy as np
from scipy.optimize import fmin_l_bfgs_b


class _Coordinator(object):
    def __init__(self, n_tasks):
        self.cv = threading.Condition()
        self.pending = {}        # task -> x awaiting evaluation
        self.answers = {}        # task -> (f, g)
        self.running = n_tasks   # searches that are neither finished nor blocked on a request
        self.failed = None

    def request(self, task, x):
        with self.cv:
            self.pending[task] = np.array(x, dtype=np.float64, copy=True)
            self.running -= 1
            self.cv.notify_all()
            while task not in self.answers and self.failed is None:
                self.cv.wait()
            if self.failed is not None:
                raise RuntimeError("batched acquisition evaluation failed") from self.failed
            return self.answers.pop(task)

    def finished(self):
        with self.cv:
            self.running -= 1
            self.cv.notify_all()


def _reverse_communication_driver():
    """SciPy's compiled L-BFGS-B stepper (`scipy.optimize._lbfgsb.setulb`, the routine fmin_l_bfgs_b itself loops
    over) if it is present with the argument list this module knows (SciPy >= 1.15), else None."""
    try:
        from scipy.optimize import _lbfgsb
        n, m = 1, 10
        itype = _lbfgsb.setulb.__doc__ and np.int32
        x = np.array([0.5]); g = np.zeros(1)
        wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m)
        iwa = np.zeros(3 * n, dtype=itype)
        task, ln_task = np.zeros(2, dtype=itype), np.zeros(2, dtype=itype)
        lsave, isave, dsave = np.zeros(4, dtype=itype), np.zeros(44, dtype=itype), np.zeros(29)
        _lbfgsb.setulb(m, x, np.zeros(1), np.ones(1), np.full(1, 2, dtype=itype), 0.0, g, 1e7, 1e-5, wa, iwa, task,
                       lsave, isave, dsave, 20, ln_task)
        return _lbfgsb.setulb if task[0] == 3 else None          # first call must ask for f and g
    except Exception:    # noqa: BLE001 - any mismatch means "use the public API instead"
        return None


_SETULB = _reverse_communication_driver()


def _stepper_agrees_with_public_api():
    """The stepper drives a PRIVATE SciPy routine with a hand-written argument list; a SciPy whose workspace layout or task
    codes changed could pass the one-call probe above and still produce other iterates.  So, once per process: a small
    bounded problem (Rosenbrock in 3 variables, active bounds) through the stepper and through fmin_l_bfgs_b must give
    identical x, f, iteration and evaluation counts - otherwise the threaded public-API driver is used."""
    def fg(x):
        f = float(np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1.0 - x[:-1]) ** 2))
        g = np.zeros_like(x)
        g[:-1] = -400.0 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2.0 * (1.0 - x[:-1])
        g[1:] += 200.0 * (x[1:] - x[:-1] ** 2)
        return f, g

    bounds = [(-1.5, 0.8), (-0.5, 2.0), (0.1, 3.0)]
    starts = [np.array([-1.2, 1.0, 0.5]), np.array([0.5, 0.5, 2.5])]
    try:
        def evaluate(tasks, X):
            vals = [fg(x) for x in X]
            return np.array([v[0] for v in vals]), np.vstack([v[1] for v in vals])
        got = _minimize_batch_stepper(starts, evaluate, bounds, maxiter=25)
        for x0, (x, f, info) in zip(starts, got):
            xr, fr, ir = fmin_l_bfgs_b(fg, x0, bounds=bounds, approx_grad=False, maxiter=25)
            if not (np.array_equal(x, xr) and f == fr and info["nit"] == ir["nit"] and info["funcalls"] == ir["funcalls"]):
                return False
        return True
    except Exception:    # noqa: BLE001
        return False


def _minimize_batch_stepper(x0s, evaluate, bounds, maxiter, m=10, factr=1e7, pgtol=1e-5, maxls=20, maxfun=15000):
    """Same searches as fmin_l_bfgs_b (identical arguments to the same compiled stepper, same bookkeeping as
    scipy/optimize/_lbfgsb_py.py:_minimize_lbfgsb), but all of them are stepped from ONE Python loop: every search
    runs until it asks for f and g, the requests are evaluated in one batch, and so on.  No threads, and SciPy's
    per-evaluation Python overhead (~100 us per search and step) disappears."""
    itype = np.int32
    lo = np.array([b[0] for b in bounds], dtype=np.float64)
    hi = np.array([b[1] for b in bounds], dtype=np.float64)
    n = lo.shape[0]
    nbd = np.zeros(n, dtype=itype)
    low_bnd, upper_bnd = np.zeros(n), np.zeros(n)
    for i in range(n):
        has_lo, has_hi = np.isfinite(lo[i]), np.isfinite(hi[i])
        if has_lo:
            low_bnd[i] = lo[i]
        if has_hi:
            upper_bnd[i] = hi[i]
        nbd[i] = {(False, False): 0, (True, False): 1, (True, True): 2, (False, True): 3}[(bool(has_lo), bool(has_hi))]

    class Search(object):
        __slots__ = ("x", "f", "g", "wa", "iwa", "task", "ln_task", "lsave", "isave", "dsave", "nit", "nfev", "done")

    searches = []
    for x0 in x0s:
        s = Search()
        s.x = np.array(np.clip(x0, lo, hi), dtype=np.float64)
        s.f = 0.0
        s.g = np.zeros(n)
        s.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m)
        s.iwa = np.zeros(3 * n, dtype=itype)
        s.task, s.ln_task = np.zeros(2, dtype=itype), np.zeros(2, dtype=itype)
        s.lsave, s.isave, s.dsave = np.zeros(4, dtype=itype), np.zeros(44, dtype=itype), np.zeros(29)
        s.nit = s.nfev = 0
        s.done = False
        searches.append(s)

    while True:
        pending = []
        for t, s in enumerate(searches):
            if s.done:
                continue
            while True:
                _SETULB(m, s.x, low_bnd, upper_bnd, nbd, s.f, s.g, factr, pgtol, s.wa, s.iwa, s.task, s.lsave, s.isave,
                        s.dsave, maxls, s.ln_task)
                if s.task[0] == 3:                  # wants f and g at the current x
                    if s.nfev >= maxfun:            # SciPy's evaluation budget (_lbfgsb_py.py: task 5 / 502, warnflag 1)
                        s.task[0], s.task[1] = 5, 502
                        continue
                    pending.append(t)
                    break
                if s.task[0] == 1:                  # a new iterate has been accepted
                    s.nit += 1
                    if s.nit >= maxiter:
                        s.task[0], s.task[1] = 5, 504
                    continue
                s.done = True
                break
        if not pending:
            break
        X = np.vstack([searches[t].x for t in pending])
        f, G = evaluate(pending, X)
        for i, t in enumerate(pending):
            s = searches[t]
            s.f = float(f[i])
            s.g = np.array(G[i], dtype=np.float64, copy=True)
            s.nfev += 1
    return [(s.x, s.f, {"grad": s.g, "funcalls": s.nfev, "nit": s.nit,
                        "warnflag": 0 if s.task[0] == 4 else (1 if (s.nit >= maxiter or s.nfev >= maxfun) else 2)})
            for s in searches]


def minimize_batch(x0s, evaluate, bounds, maxiter=20, driver="auto"):
    """Run one L-BFGS-B search per row of `x0s`; `evaluate(tasks, X)` -> (f, G) scores the rows of X, where tasks[i]
    is the index of the search that asked for row i.  Returns the list of (x, f, info) tuples of fmin_l_bfgs_b.

    driver: "stepper" steps SciPy's compiled routine directly (fast path), "threads" runs the public
    fmin_l_bfgs_b in one thread per search; "auto" prefers the stepper when this SciPy exposes it.  Both give the
    iterates of a plain sequential fmin_l_bfgs_b call."""
    x0s = [np.asarray(x, dtype=np.float64) for x in x0s]
    n = len(x0s)
    if n == 0:
        return []
    if driver == "stepper" and _SETULB is None:
        raise RuntimeError("this SciPy does not expose the reverse-communication L-BFGS-B stepper")
    if driver == "stepper" or (driver == "auto" and _SETULB is not None):
        return _minimize_batch_stepper(x0s, evaluate, bounds, maxiter)
    co = _Coordinator(n)
    results = [None] * n
    errors = [None] * n

    def worker(t):
        try:
            def fun(x):
                f, g = co.request(t, x)
                return float(f), np.array(g, dtype=np.float64, copy=True)
            results[t] = fmin_l_bfgs_b(fun, x0s[t], bounds=bounds, approx_grad=False, maxiter=maxiter)
        except BaseException as e:    # noqa: BLE001 - reported to the caller below
            errors[t] = e
        finally:
            co.finished()

    threads = [threading.Thread(target=worker, args=(t,), daemon=True) for t in range(n)]
    for th in threads:
        th.start()
    alive = n
    while alive > 0:
        with co.cv:
            while co.running > 0:
                co.cv.wait()
            batch = sorted(co.pending.items())
            co.pending = {}
            alive = len(batch)
            if alive == 0:
                break
        tasks = [t for t, _ in batch]
        X = np.vstack([x for _, x in batch])
        try:
            f, G = evaluate(tasks, X)
        except BaseException as e:    # noqa: BLE001
            with co.cv:
                co.failed = e
                co.cv.notify_all()
            for th in threads:
                th.join()
            raise
        with co.cv:
            for i, t in enumerate(tasks):
                co.answers[t] = (f[i], G[i])
            co.running += len(tasks)
            co.cv.notify_all()
    for th in threads:
        th.join()
    for e in errors:
        if e is not None:
            raise e
    return results


if _SETULB is not None and not _stepper_agrees_with_public_api():
    _SETULB = None                                     # fall back to fmin_l_bfgs_b in threads (same iterates, slower)
