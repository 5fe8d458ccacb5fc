"""Time of one log-marginal-likelihood evaluation (value + gradient) and of a whole fit, host (scikit-learn) vs GPU
(skb_fit_lml_host).  Usage: python tools/bench_fit.py [--full-fit]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skopt_b200.learning import GaussianProcessRegressor  # noqa: E402
from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel  # noqa: E402


def problem(n, d, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n)
    kernel = ConstantKernel(1.0, (0.01, 1000.0)) * Matern(length_scale=np.ones(d), length_scale_bounds=[(0.01, 100)] * d,
                                                            nu=2.5) + WhiteKernel(0.01, (1e-6, 10.0))
    return X, y, kernel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-fit", action="store_true")
    ap.add_argument("--sizes", default="512x10,1024x10,2048x20,4096x50")
    args = ap.parse_args()
    out = []
    for spec in args.sizes.split(","):
        n, d = map(int, spec.split("x"))
        X, y, kernel = problem(n, d)
        gpr = GaussianProcessRegressor(kernel=kernel, normalize_y=True, optimizer=None).fit(X, y)
        gpr.kernel_ = kernel.clone_with_theta(kernel.theta)
        theta = kernel.theta
        rec = {"n": n, "d": d}
        gpr._skb_fit_device = "gpu"
        gpr.log_marginal_likelihood(theta, eval_gradient=True)            # create the device handle, warm up
        t = time.perf_counter()
        reps = 10
        for _ in range(reps):
            v_g, g_g = gpr.log_marginal_likelihood(theta, eval_gradient=True)
        rec["gpu_ms"] = (time.perf_counter() - t) / reps * 1e3
        if n <= 2048:
            gpr._skb_fit_device = "host"
            t = time.perf_counter()
            v_h, g_h = gpr.log_marginal_likelihood(theta, eval_gradient=True)
            rec["host_ms"] = (time.perf_counter() - t) * 1e3
            rec["lml_rel_diff"] = abs(v_g - v_h) / abs(v_h)
            rec["grad_rel_diff"] = float(np.max(np.abs(g_g - g_h)) / np.max(np.abs(g_h)))
        gpr._close_device_fit()
        if args.full_fit and n <= 1024:
            for where in ("gpu", "host"):
                est = GaussianProcessRegressor(kernel=kernel, normalize_y=True, n_restarts_optimizer=2, random_state=0)
                est._skb_fit_device = where
                t = time.perf_counter()
                est.fit(X, y)
                rec["fit_%s_s" % where] = time.perf_counter() - t
                rec["fit_%s_lml" % where] = float(est.log_marginal_likelihood_value_)
        out.append(rec)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
