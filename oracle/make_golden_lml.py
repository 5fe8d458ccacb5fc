"""Fixtures for the fit objective: log_marginal_likelihood(theta, eval_gradient=True) of the UNMODIFIED reference
estimator (skopt.learning.GaussianProcessRegressor -> scikit-learn _gpr.py:541-657) on seeded training sets, for every
kernel kind of the grammar, anisotropic and isotropic length scales, free and fixed hyper-parameters.
Build container only (needs /root/reference).  TEST INFRASTRUCTURE ONLY."""
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
if not hasattr(np, "int"):
    np.int = int
import sklearn  # noqa: E402
from skopt.learning import GaussianProcessRegressor  # noqa: E402
from skopt.learning.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

CASES = [  # name, n, d, nu (inf = RBF), isotropic, fixed amplitude, noise kind
    ("m52_ard", 96, 5, 2.5, False, False, "free"),
    ("m32_ard", 70, 3, 1.5, False, False, "free"),
    ("m12_ard", 50, 4, 0.5, False, True, "free"),
    ("rbf_ard", 64, 6, np.inf, False, False, "fixed"),
    ("m52_iso", 80, 4, 2.5, True, False, "free"),
    ("m52_nonoise", 40, 2, 2.5, False, False, "none"),
    ("m52_big", 300, 20, 2.5, False, False, "free"),
]


def build(nu, d, iso, fixed_amp, noise_kind, rng):
    ls = 0.5 + rng.rand() if iso else 0.3 + 1.5 * rng.rand(d)
    base = RBF(length_scale=ls, length_scale_bounds=(0.01, 100)) if nu == np.inf else \
        Matern(length_scale=ls, length_scale_bounds=(0.01, 100), nu=nu)
    k = ConstantKernel(1.7, "fixed" if fixed_amp else (0.01, 1000.0)) * base
    if noise_kind == "free":
        k = k + WhiteKernel(0.05, (1e-6, 10.0))
    elif noise_kind == "fixed":
        k = k + WhiteKernel(0.02, "fixed")
    return k


def main():
    warnings.simplefilter("ignore")
    rec = {"sklearn_version": np.array(sklearn.__version__)}
    for name, n, d, nu, iso, fixed_amp, noise_kind in CASES:
        rng = np.random.RandomState(sum(map(ord, name)))
        X = rng.rand(n, d)
        y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n)
        kernel = build(nu, d, iso, fixed_amp, noise_kind, rng)
        gpr = GaussianProcessRegressor(kernel=kernel, normalize_y=True, optimizer=None, alpha=1e-10).fit(X, y)
        # after skopt's fit the White term of kernel_ is silenced: evaluate on a clone of the ORIGINAL kernel
        gpr.kernel_ = kernel.clone_with_theta(kernel.theta)
        thetas = [kernel.theta.copy()]
        for _ in range(3):
            thetas.append(kernel.theta + 0.6 * rng.randn(kernel.theta.size))
        vals, grads = [], []
        for th in thetas:
            v, g = gpr.log_marginal_likelihood(th, eval_gradient=True, clone_kernel=True)
            vals.append(v)
            grads.append(g)
        rec[name + "_X"] = X
        rec[name + "_y"] = gpr.y_train_
        rec[name + "_nu"] = np.array(nu)
        rec[name + "_iso"] = np.array(iso)
        rec[name + "_fixed_amp"] = np.array(fixed_amp)
        rec[name + "_noise_kind"] = np.array(noise_kind)
        rec[name + "_theta0"] = kernel.theta
        rec[name + "_length_scale"] = np.atleast_1d(np.asarray(kernel.get_params()["k1__k2__length_scale" if noise_kind != "none" else "k2__length_scale"], dtype=float))
        rec[name + "_thetas"] = np.asarray(thetas)
        rec[name + "_lml"] = np.asarray(vals)
        rec[name + "_grad"] = np.asarray(grads)
    np.savez_compressed(os.path.join(OUT, "fit_lml.npz"), **rec)
    print("wrote fit_lml.npz:", {k: v.shape for k, v in rec.items() if k.endswith("_lml")})


if __name__ == "__main__":
    main()
