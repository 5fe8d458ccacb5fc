"""Fixtures for the all-categorical GP (HammingKernel, skopt/learning/gaussian_process/kernels.py:317-407) from the
UNMODIFIED reference: two fitted states in the format of oracle/make_golden.py (no gradients: the kernel has no
gradient_x, which the fixture records as `grad_error`) and gp_minimize trajectories on all-categorical spaces, where
cook_estimator picks the HammingKernel (utils.py:377-378) and acq_optimizer="auto" resolves to "sampling"
(utils.py:320-330, optimizer.py:236-247).
Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_hamming.py
TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, emit  # noqa: E402  (puts /root/reference on sys.path, np.int shim)

from skopt import gp_minimize  # noqa: E402
from skopt.learning import GaussianProcessRegressor  # noqa: E402
from skopt.learning.gaussian_process.kernels import ConstantKernel, HammingKernel  # noqa: E402
from skopt.space import Space  # noqa: E402
from skopt.utils import cook_estimator, normalize_dimensions  # noqa: E402


def label_grid(rng, m, n_cats):
    """Rows of a normalised all-categorical space: label / (n_categories - 1) per dimension (space.py:643-650)."""
    return np.column_stack([rng.randint(0, k, size=m) / (k - 1.0) for k in n_cats])


def objective(x):
    score = {"a": 0.0, "b": 1.0, "c": -0.5, "d": 0.3}[x[0]] + {"x": 0.2, "y": -0.4, "z": 0.9}[x[1]]
    return score * (1.5 if x[2] == "hi" else 0.7) + 0.1 * len(x[3])


DIMS = [["a", "b", "c", "d"], ["x", "y", "z"], ["lo", "hi"], ["p", "qq", "rrr", "ssss", "ttttt"]]


def main():
    warnings.simplefilter("ignore")
    rng = np.random.RandomState(4321)

    # 1. fixed hyper-parameters, anisotropic weights, scaled by a constant; candidates contain training rows and ties
    n_cats = [4, 3, 2, 5, 3]
    X = label_grid(rng, 45, n_cats)
    w = np.array([0.3, 1.1, 0.7, 2.0, 0.05])
    y = np.sin(3.0 * X @ w) + 0.05 * rng.randn(45)
    Xc = label_grid(rng, 300, n_cats)
    Xc[5:9] = X[:4]
    Xc[12] = Xc[11]
    emit("hamming_aniso_const", GaussianProcessRegressor(
        ConstantKernel(1.7, "fixed") * HammingKernel(w, "fixed"), normalize_y=True, noise=1e-3, optimizer=None),
        X, y, Xc, 4, ctor=dict(kernel="c*hamming", amplitude=1.7, length_scale=w.tolist(), normalize_y=True, noise=1e-3))

    # 2. the cooked estimator of an all-categorical space after a real marginal-likelihood fit
    space = Space(normalize_dimensions(DIMS))
    Xo = space.rvs(30, random_state=8)
    y = np.array([objective(x) for x in Xo])
    Xt = space.transform(Xo)
    est = cook_estimator("GP", space=space, random_state=5, noise="gaussian")
    Xc = space.transform(space.rvs(400, random_state=9))
    emit("hamming_cooked_categorical", est, Xt, y, Xc, 4,
         ctor=dict(kernel="cooked", noise="gaussian", random_state=5, dims=4))

    # 3. trajectories
    rec = {}
    runs = {
        "cat_auto_ei": dict(n_calls=16, n_initial_points=6, acq_func="EI", n_points=1000, random_state=3,
                            noise="gaussian"),
        "cat_sampling_hedge": dict(n_calls=15, n_initial_points=6, acq_func="gp_hedge", acq_optimizer="sampling",
                                   n_points=800, random_state=4, noise="gaussian"),
    }
    for name, kw in runs.items():
        res = gp_minimize(objective, DIMS, **kw)
        rec[name + "_x_iters"] = np.array(json.dumps(res.x_iters))
        rec[name + "_func_vals"] = np.asarray(res.func_vals, dtype=float)
        rec[name + "_kernel"] = np.array(repr(res.models[-1].kernel_))
        print(name, "fun=%.6f" % res.fun, res.models[-1].kernel_)
    rec["runs"] = np.array(json.dumps(runs))
    rec["dims"] = np.array(json.dumps(DIMS))

    # 4. the kernel class itself: K(X), dK/dtheta (what scikit-learn's marginal-likelihood fit evaluates), K(X, Y)
    Xk, Yk = label_grid(rng, 20, [4, 3, 5]), label_grid(rng, 7, [4, 3, 5])
    for tag, ls in (("aniso", np.array([0.4, 1.3, 2.2])), ("iso", 0.8), ("iso1", [1.7])):
        K, G = HammingKernel(ls)(Xk, eval_gradient=True)
        rec["hk_%s_K" % tag], rec["hk_%s_grad" % tag] = K, G
        rec["hk_%s_cross" % tag] = HammingKernel(ls)(Xk, Yk)
    rec["hk_X"], rec["hk_Y"] = Xk, Yk
    np.savez_compressed(os.path.join(OUT, "traj_categorical.npz"), **rec)


if __name__ == "__main__":
    main()
