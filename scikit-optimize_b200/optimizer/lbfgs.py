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


def minimize_batch(x0s, evaluate, bounds, maxiter=20):
    """Run one L-BFGS-B search per row of `x0s`; `evaluate(tasks, X)` -> (f, G) scores the rows of X, where tasks[i]
    is the index of the search that asked for row i.  Returns the list of (x, f, info) tuples of fmin_l_bfgs_b."""
    x0s = [np.asarray(x, dtype=np.float64) for x in x0s]
    n = len(x0s)
    if n == 0:
        return []
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
