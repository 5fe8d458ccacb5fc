"""Where a fit() goes with the GPU objective: search (LML evaluations), scikit-learn's final factorisation,
L_inv / K_inv post-processing, device state creation.  Usage: python tools/profile_fit_tell.py [n d]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skopt_b200 import engine  # noqa: E402
from skopt_b200.utils import cook_estimator  # noqa: E402

n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 10)
rng = np.random.RandomState(0)
X = rng.rand(n, d)
y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n)

acc = {"lml_s": 0.0, "evals": 0}
orig = engine.DeviceFit.lml


def timed(self, theta_full, eval_gradient=True):
    t = time.perf_counter()
    out = orig(self, theta_full, eval_gradient)
    acc["lml_s"] += time.perf_counter() - t
    acc["evals"] += 1
    return out


engine.DeviceFit.lml = timed
for rep in range(2):
    acc.update(lml_s=0.0, evals=0)
    est = cook_estimator("GP", space=[(0.0, 1.0)] * d, random_state=1)
    est._skb_fit_device = "gpu"
    t0 = time.perf_counter()
    est.fit(X, y)
    t_fit = time.perf_counter() - t0
    t0 = time.perf_counter()
    dev = est.device_state(0)
    dev.predict_mean(X[:8])
    t_state = time.perf_counter() - t0
    t0 = time.perf_counter()
    from scipy.linalg import solve_triangular
    Li = solve_triangular(est.L_, np.eye(n), lower=True)
    t_linv = time.perf_counter() - t0
    t0 = time.perf_counter()
    Li.T.dot(Li)
    t_kinv = time.perf_counter() - t0
    print("n=%d d=%d fit %.3fs: %d LML evals %.3fs (%.2f ms each), rest of fit %.3fs [of which L_inv %.3fs, K_inv %.3fs]; "
          "state upload + first predict %.3fs" % (n, d, t_fit, acc["evals"], acc["lml_s"], 1e3 * acc["lml_s"] / max(1, acc["evals"]),
                                                   t_fit - acc["lml_s"], t_linv, t_kinv, t_state), flush=True)
