"""Fixture for partial dependence (skopt/plots.py:896-1055) from the UNMODIFIED reference: a GP fitted by the
reference's Optimizer on a mixed Real / Integer / Categorical space, random transformed samples, and the reference's
partial_dependence_1D / partial_dependence_2D results.  matplotlib is not installed in the build container and
skopt/plots.py imports it at module level, so inert stand-in modules are registered first (none of the functions called
here touches them).  Build container only.  TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys
import types
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
if not hasattr(np, "int"):
    np.int = int
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker"):
    mod = types.ModuleType(name)
    mod.__dict__.update(cm=None, pyplot=None, LogLocator=None, MaxNLocator=None, FuncFormatter=None)
    sys.modules[name] = mod
from skopt import Optimizer  # noqa: E402
from skopt.benchmarks import branin  # noqa: E402
from skopt.plots import partial_dependence_1D, partial_dependence_2D  # noqa: E402
from make_golden import OUT, VERSIONS, kernel_params  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    dims = [(-5.0, 10.0), (0, 15), ["a", "b", "c"]]
    opt = Optimizer(dims, "GP", n_initial_points=6, acq_func="EI", acq_optimizer="sampling",
                    acq_optimizer_kwargs=dict(n_points=500), random_state=11)
    for _ in range(16):
        x = opt.ask()
        opt.tell(x, float(branin(x[:2]) + {"a": 0.0, "b": 3.0, "c": -2.0}[x[2]]))
    est, space = opt.models[-1], opt.space
    samples = space.transform(space.rvs(n_samples=50, random_state=np.random.RandomState(5)))
    rec = dict(X_train=est.X_train_, alpha=est.alpha_, K_inv=est.K_inv_, L=est.L_,
               y_train_mean=np.float64(np.ravel(est.y_train_mean_)[0]), y_train_std=np.float64(np.ravel(est.y_train_std_)[0]),
               samples=samples)
    for i in range(3):
        xi, yi = partial_dependence_1D(space, est, i, samples, n_points=9)
        rec["pd1_x_%d" % i], rec["pd1_y_%d" % i] = np.asarray(xi, dtype=float), np.asarray(yi, dtype=float)
    for i, j in ((0, 1), (2, 0), (1, 2)):
        xi, yi, zi = partial_dependence_2D(space, est, i, j, samples, n_points=6)
        rec["pd2_x_%d%d" % (i, j)], rec["pd2_y_%d%d" % (i, j)] = np.asarray(xi, dtype=float), np.asarray(yi, dtype=float)
        rec["pd2_z_%d%d" % (i, j)] = zi
    meta = dict(name="traj_partial_dependence", versions=VERSIONS, kernel_repr=repr(est.kernel_),
                kernel_params=kernel_params(est.kernel_), kernel_class=type(est.kernel_).__name__)
    rec["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, "traj_partial_dependence.npz"), **rec)
    print("partial dependence fixture: n=%d, transformed dims=%d, kernel_=%s" % (est.X_train_.shape[0], samples.shape[1], est.kernel_))


if __name__ == "__main__":
    main()
