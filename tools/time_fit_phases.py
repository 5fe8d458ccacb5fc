"""Cost of one fit-objective evaluation on the GPU by phase: value only (K build + potrf + potrs + lml reduction) against
value + gradient (adds potri + the gradient contraction), single and 3 concurrent evaluations (what a lock-step round of
the default 2-restart search issues).  Usage: python tools/time_fit_phases.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skopt_b200.engine import DeviceFit  # noqa: E402

for n, d in ((512, 10), (1024, 10), (2048, 20), (4096, 50)):
    rng = np.random.RandomState(0)
    X = rng.rand(n, d)
    y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n)
    y = (y - y.mean()) / y.std()
    fit = DeviceFit(X, y, 3, 1e-10)
    theta = np.concatenate([[0.0], np.zeros(d), [np.log(1e-2)]])
    thetas = np.stack([theta, theta + 0.1, theta - 0.1])
    for _ in range(3):
        fit.lml(theta, eval_gradient=True)
        fit.lml_batch(thetas)

    def timed(fn, reps=10):
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    t_val = timed(lambda: fit.lml(theta, eval_gradient=False))
    t_grad = timed(lambda: fit.lml(theta, eval_gradient=True))
    t_b3 = timed(lambda: fit.lml_batch(thetas))
    print("n=%5d d=%3d: value %.2f ms, value+grad %.2f ms (potri + grad kernel %.2f ms), 3 concurrent %.2f ms" % (
        n, d, t_val, t_grad, t_grad - t_val, t_b3), flush=True)
    fit.close()
