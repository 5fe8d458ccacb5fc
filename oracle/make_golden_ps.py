"""Fixtures for the per-second acquisitions (EIps / PIps) from the UNMODIFIED reference: Optimizer trajectories and
_gaussian_acquisition values / gradients on the fitted (objective, log-time) model pair.  Build container only.
TEST INFRASTRUCTURE ONLY."""
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
if not hasattr(np, "int"):
    np.int = int
from skopt import Optimizer  # noqa: E402
from skopt.acquisition import _gaussian_acquisition, gaussian_acquisition_1D  # noqa: E402
from skopt.benchmarks import branin  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def objective(x):
    return branin(x), 1.0 + 0.1 * (x[0] + 5.0) + 0.05 * x[1]


def main():
    warnings.simplefilter("ignore")
    rec = {}
    for acq, optm in (("EIps", "sampling"), ("PIps", "sampling"), ("EIps", "lbfgs")):
        opt = Optimizer([(-5.0, 10.0), (0.0, 15.0)], "GP", n_initial_points=5, acq_func=acq, acq_optimizer=optm,
                        acq_optimizer_kwargs=dict(n_points=1500, n_restarts_optimizer=4), random_state=3)
        for _ in range(10):
            x = opt.ask()
            opt.tell(x, objective(x))
        key = "%s_%s" % (acq, optm)
        rec[key + "_Xi"] = np.asarray(opt.Xi, dtype=float)
        rec[key + "_yi"] = np.asarray(opt.yi, dtype=float)
        est = opt.models[-1]
        Xc = opt.space.transform(opt.space.rvs(64, random_state=np.random.RandomState(9)))
        y_opt = np.min(opt.yi)
        rec[key + "_Xc"] = Xc
        rec[key + "_vals"] = _gaussian_acquisition(Xc, est, y_opt, acq, acq_func_kwargs=None)
        g = np.empty((8, 2)); v = np.empty(8)
        for i in range(8):
            vi, gi = gaussian_acquisition_1D(Xc[i], est, y_opt, acq, None)
            v[i], g[i] = vi[0], gi
        rec[key + "_v1d"], rec[key + "_g1d"] = v, g
        print(key, opt.Xi[-1])
    np.savez_compressed(os.path.join(OUT, "traj_per_second.npz"), **rec)


if __name__ == "__main__":
    main()
