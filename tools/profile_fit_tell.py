"""Where a fit() goes with the GPU objective: search (LML evaluations), final factorisation + state build on the device
(skb_fit_finalize_host / skb_state_create_from_fit), against the same fit finished on the host (scikit-learn's final
Cholesky, solve_triangular, L_inv^T L_inv, upload).  Usage: python tools/profile_fit_tell.py [n d]"""
import json
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

acc = {}


def wrap(cls, name):
    orig = getattr(cls, name)

    def timed(self, *a, **k):
        t = time.perf_counter()
        out = orig(self, *a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
        acc[name + "_calls"] = acc.get(name + "_calls", 0) + 1
        return out

    setattr(cls, name, timed)


for meth in ("lml", "lml_batch", "finalize", "make_state"):
    wrap(engine.DeviceFit, meth)

for rep in range(3):
    acc.clear()
    est = cook_estimator("GP", space=[(0.0, 1.0)] * d, random_state=1)
    est._skb_fit_device = "gpu"
    t0 = time.perf_counter()
    est.fit(X, y)
    t_fit = time.perf_counter() - t0
    t0 = time.perf_counter()
    est.predict(X[:8], return_std=True)
    t_first = time.perf_counter() - t0
    # the host tail this replaces, timed on the same factor
    from scipy.linalg import cho_solve, cholesky, solve_triangular
    t0 = time.perf_counter()
    K = est.kernel_(est.X_train_)
    K[np.diag_indices_from(K)] += est.noise_ if est.noise_ else 0.0
    L = cholesky(K + 1e-10 * np.eye(n), lower=True)
    cho_solve((L, True), est.y_train_)
    t_host_chol = time.perf_counter() - t0
    t0 = time.perf_counter()
    Li = solve_triangular(L, np.eye(n), lower=True)
    Ki = Li.T.dot(Li)
    t_host_inv = time.perf_counter() - t0
    t0 = time.perf_counter()
    hs = engine.HostState(est.X_train_, est.alpha_, Li, np.ones(d), 3, 1.0, 0.0, 0.0, 0.0, 1.0, K_inv=Ki)
    ds = engine.DeviceState(hs, 0)
    ds._attach_kinv()
    ds.close()
    t_host_upload = time.perf_counter() - t0
    rec = {"n": n, "d": d, "rep": rep, "fit_s": t_fit, "search_s": acc.get("lml", 0.0) + acc.get("lml_batch", 0.0),
           "lml_calls": acc.get("lml_calls", 0), "lml_batch_calls": acc.get("lml_batch_calls", 0),
           "finalize_s": acc.get("finalize", 0.0), "make_state_s": acc.get("make_state", 0.0),
           "first_predict_s": t_first, "host_tail_chol_s": t_host_chol, "host_tail_inverse_s": t_host_inv,
           "host_tail_upload_pack_s": t_host_upload}
    print(json.dumps(rec), flush=True)
