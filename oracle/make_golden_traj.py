"""Trajectory fixtures from the UNMODIFIED reference: gp_minimize / Optimizer runs whose suggested points the GPU
build must reproduce (sampling optimiser: identical points; lbfgs: within tolerance).
Run in the build container only.  TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
if not hasattr(np, "int"):
    np.int = int
from skopt import Optimizer, gp_minimize  # noqa: E402
from skopt.benchmarks import branin, hart6  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    warnings.simplefilter("ignore")
    rec = {}
    runs = {
        "branin_sampling_ei": dict(func="branin", dims=[(-5.0, 10.0), (0.0, 15.0)], n_calls=14, n_initial_points=5,
                                   acq_func="EI", acq_optimizer="sampling", n_points=2000, random_state=1,
                                   noise="gaussian"),
        "branin_sampling_hedge": dict(func="branin", dims=[(-5.0, 10.0), (0.0, 15.0)], n_calls=14, n_initial_points=5,
                                      acq_func="gp_hedge", acq_optimizer="sampling", n_points=2000, random_state=2,
                                      noise="gaussian"),
        "branin_sampling_lcb_noise1e-10": dict(func="branin", dims=[(-5.0, 10.0), (0.0, 15.0)], n_calls=12,
                                               n_initial_points=5, acq_func="LCB", acq_optimizer="sampling",
                                               n_points=2000, random_state=3, noise=1e-10),
        "hart6_lbfgs_ei": dict(func="hart6", dims=[(0.0, 1.0)] * 6, n_calls=14, n_initial_points=8, acq_func="EI",
                               acq_optimizer="lbfgs", n_points=2000, n_restarts_optimizer=5, random_state=4,
                               noise="gaussian"),
        "hart6_lbfgs_hedge": dict(func="hart6", dims=[(0.0, 1.0)] * 6, n_calls=13, n_initial_points=8,
                                  acq_func="gp_hedge", acq_optimizer="lbfgs", n_points=2000, n_restarts_optimizer=4,
                                  random_state=5, noise="gaussian"),
        "mixed_sampling_pi": dict(func="mixed", dims=[(-2.0, 2.0), (1, 6), ["a", "b", "c"]], n_calls=13,
                                  n_initial_points=6, acq_func="PI", acq_optimizer="sampling", n_points=1500,
                                  random_state=6, noise="gaussian"),
    }
    funcs = {"branin": branin, "hart6": hart6,
             "mixed": lambda x: (x[0] - 0.5) ** 2 + 0.3 * (x[1] - 3) ** 2 + {"a": 0.0, "b": 1.0, "c": -0.5}[x[2]]}
    for name, cfg in runs.items():
        kw = {k: v for k, v in cfg.items() if k not in ("func", "dims")}
        res = gp_minimize(funcs[cfg["func"]], cfg["dims"], **kw)
        rec[name + "_x_iters"] = np.array(json.dumps(res.x_iters, default=lambda o: o.item()))
        rec[name + "_func_vals"] = np.asarray(res.func_vals, dtype=float)
        print(name, "fun=%.6f" % res.fun)
    rec["runs"] = np.array(json.dumps(runs))

    # constant liar: ask(n_points=4) on an optimizer that has seen 7 points (optimizer.py:366-421)
    for strategy in ("cl_min", "cl_mean", "cl_max"):
        opt = Optimizer([(-5.0, 10.0), (0.0, 15.0)], "GP", n_initial_points=5, acq_func="EI", acq_optimizer="sampling",
                        acq_optimizer_kwargs=dict(n_points=1500), random_state=11)
        for _ in range(7):
            x = opt.ask()
            opt.tell(x, branin(x))
        pts = opt.ask(n_points=4, strategy=strategy)
        rec["cl_" + strategy] = np.asarray(pts, dtype=float)
        rec["cl_" + strategy + "_next_after"] = np.asarray(opt.ask(), dtype=float)
        print("constant liar", strategy, pts)
    np.savez_compressed(os.path.join(OUT, "traj_gp_minimize.npz"), **rec)


if __name__ == "__main__":
    main()
