"""cProfile of one GPU fit (cooked Matern-5/2 ARD + White, 2 restarts) at n = 1024, d = 10: GPU objective vs host overhead.
Usage: python tools/profile_fit.py [n] [d]"""
import cProfile
import os
import pstats
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skopt_b200  # noqa: E402
from skopt_b200.utils import cook_estimator  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rng = np.random.RandomState(0)
X = rng.rand(n, d)
y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n)
skopt_b200.set_fit_device("gpu")
warnings.simplefilter("ignore")
est = cook_estimator("GP", space=[(0.0, 1.0)] * d, random_state=1)
est.fit(X, y)
t0 = time.perf_counter()
for _ in range(3):
    cook_estimator("GP", space=[(0.0, 1.0)] * d, random_state=1).fit(X, y)
print("fit: %.1f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
pr = cProfile.Profile()
pr.enable()
cook_estimator("GP", space=[(0.0, 1.0)] * d, random_state=1).fit(X, y)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
