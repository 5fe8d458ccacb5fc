"""First-call and steady-state cost of the fit objective at small n (what a BO loop at n <= 200 sees: one skb_fit per tell):
create, first batch of 3, next 30 batches, destroy.  Run with SKB_FIT_GRAPH=0 / 1."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skopt_b200.engine import DeviceFit  # noqa: E402

for n, d in ((40, 2), (200, 6), (600, 10)):
    rng = np.random.RandomState(0)
    X = rng.rand(n, d)
    y = np.sin(X.dot(rng.randn(d)))
    y = (y - y.mean()) / y.std()
    theta = np.concatenate([[0.0], np.zeros(d), [np.log(1e-2)]])
    thetas = np.stack([theta, theta + 0.1, theta - 0.1])
    DeviceFit(X, y, 3, 1e-10).close()                      # context, attributes
    rows = []
    for rep in range(5):
        t0 = time.perf_counter()
        fit = DeviceFit(X, y, 3, 1e-10)
        t1 = time.perf_counter()
        fit.lml_batch(thetas)
        t2 = time.perf_counter()
        for _ in range(30):
            fit.lml_batch(thetas)
        t3 = time.perf_counter()
        fit.close()
        t4 = time.perf_counter()
        rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) / 30 * 1e3, (t4 - t3) * 1e3))
    r = np.median(np.asarray(rows), axis=0)
    print("graph=%s n=%4d d=%2d: create %.2f ms, first batch of 3 %.2f ms, later batches %.3f ms each, close %.2f ms" % (
        os.environ.get("SKB_FIT_GRAPH", "1"), n, d, r[0], r[1], r[2], r[3]))
