"""A few GPU LML evaluations at (n, d) for launch-list profiling.  Usage: python tools/lml_once.py n d [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skopt_b200.engine import DeviceFit  # noqa: E402

n, d = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.RandomState(0)
X = rng.rand(n, d)
y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n)
y = (y - y.mean()) / y.std()
fit = DeviceFit(X, y, 3, 1e-10)
theta = np.concatenate([[0.0], np.zeros(d), [np.log(0.01)]])
fit.lml(theta)
t = time.perf_counter()
for _ in range(reps):
    v, g = fit.lml(theta)
print("n=%d d=%d: %.3f ms per evaluation, lml=%.6f" % (n, d, (time.perf_counter() - t) / reps * 1e3, v))
