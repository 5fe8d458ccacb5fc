"""Trajectory fixtures on WIDE search spaces (more than 64 transformed dimensions) from the UNMODIFIED reference:
a 70-dimensional Real space, and a mixed Real / Integer / Categorical space of 66 dimensions (the GP Optimizer gives every
dimension, categoricals included, ONE normalised model dimension: utils.normalize_dimensions).
Sampling optimiser, so every suggested point must be reproduced exactly.  Run in the build container only.
TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
if not hasattr(np, "int"):
    np.int = int
from skopt import gp_minimize  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CATS = list("abcdefghi")


def sphere70(x):
    x = np.asarray(x, dtype=float)
    return float(np.sum((x - 0.3) ** 2) + np.sin(5.0 * x[0]) * x[3])


def mixed66(x):
    real = sum((v - 0.1 * (i % 7)) ** 2 for i, v in enumerate(x[:40]))
    ints = sum(0.02 * (v - 3) ** 2 for v in x[40:56])
    cat = sum((CATS.index(c) - 4) ** 2 * (0.05 + 0.01 * i) for i, c in enumerate(x[56:]))
    return float(real + ints + cat)


def main():
    warnings.simplefilter("ignore")
    runs = {
        "real70_sampling_ei": dict(func="sphere70", dims=[(0.0, 1.0)] * 70, n_calls=12, n_initial_points=8,
                                   acq_func="EI", acq_optimizer="sampling", n_points=1500, random_state=21,
                                   noise="gaussian"),
        "mixed66_sampling_hedge": dict(func="mixed66", dims=[(-1.0, 1.0)] * 40 + [(0, 9)] * 16 + [CATS] * 10, n_calls=11,
                                       n_initial_points=7, acq_func="gp_hedge", acq_optimizer="sampling",
                                       n_points=1200, random_state=22, noise="gaussian"),
    }
    funcs = {"sphere70": sphere70, "mixed66": mixed66}
    rec = {}
    for name, cfg in runs.items():
        kw = {k: v for k, v in cfg.items() if k not in ("func", "dims")}
        res = gp_minimize(funcs[cfg["func"]], cfg["dims"], **kw)
        model_dims = res.models[-1].X_train_.shape[1]
        assert model_dims > 64, model_dims
        rec[name + "_x_iters"] = np.array(json.dumps(res.x_iters, default=lambda o: o.item()))
        rec[name + "_func_vals"] = np.asarray(res.func_vals, dtype=float)
        print(name, "model dims", model_dims, "fun=%.6f" % res.fun)
    rec["runs"] = np.array(json.dumps(runs))
    np.savez_compressed(os.path.join(OUT, "traj_wide.npz"), **rec)


if __name__ == "__main__":
    main()
