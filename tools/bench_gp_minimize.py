"""BASELINE.json configs[0] and configs[1] as whole runs (the reference's benchmarks/bench_branin.py:10-35 and
bench_hart6.py with the GP minimiser only):

  branin : gp_minimize on Branin, n_calls=50, acq_optimizer='sampling', n_points=10000, Matern(nu=2.5)+WhiteKernel
  hart6  : gp_minimize on Hartmann-6, acq_func='gp_hedge', acq_optimizer='lbfgs', n_restarts_optimizer=20

Prints one JSON line per run: wall time, its split into model fits and scoring blocks (Optimizer.scoring_times_, the part
that runs on the GPU), the minimum found and the call that found it (the columns of the reference's benchmarks/README).

    python tools/bench_gp_minimize.py [branin|hart6] [--n-calls 50] [--runs 3] [--fit-device host|gpu] [--abi-model]

--abi-model swaps the CUDA library for the CPU software model (tests/_abi_model.py, TEST INFRASTRUCTURE) so that the script
itself can be exercised on a machine without a GPU; timings taken that way say nothing about the product.
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skopt_b200  # noqa: E402
from skopt_b200 import Optimizer  # noqa: E402
from skopt_b200.utils import cook_estimator  # noqa: E402


def branin(x, a=1, b=5.1 / (4 * np.pi ** 2), c=5. / np.pi, r=6, s=10, t=1. / (8 * np.pi)):
    return a * (x[1] - b * x[0] ** 2 + c * x[0] - r) ** 2 + s * (1 - t) * np.cos(x[0]) + s


def hart6(x):
    alpha = np.asarray([1.0, 1.2, 3.0, 3.2])
    P = 1e-4 * np.asarray([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                           [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
    A = np.asarray([[10, 3, 17, 3.50, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8],
                    [17, 8, 0.05, 10, 0.1, 14]])
    return -np.sum(alpha * np.exp(-np.sum(A * (np.array(x) - P) ** 2, axis=1)))


CONFIGS = {
    "branin": dict(func=branin, dims=[(-5.0, 10.0), (0.0, 15.0)], acq_func="gp_hedge", acq_optimizer="sampling",
                   kwargs=dict(n_points=10000), optimum=0.397887),
    "hart6": dict(func=hart6, dims=[(0.0, 1.0)] * 6, acq_func="gp_hedge", acq_optimizer="lbfgs",
                  kwargs=dict(n_points=10000, n_restarts_optimizer=20), optimum=-3.32237),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="branin", choices=sorted(CONFIGS))
    ap.add_argument("--n-calls", type=int, default=50)
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--fit-device", default="host", choices=["host", "gpu"])
    ap.add_argument("--abi-model", action="store_true")
    args = ap.parse_args()
    if args.abi_model:
        from skopt_b200 import _lib
        from skopt_b200.optimizer import optimizer as optimizer_module
        from tests._abi_model import AbiModel
        _lib._lib = AbiModel()
        optimizer_module.device_dim_descs = lambda space: None
    cfg = CONFIGS[args.config]
    warnings.simplefilter("ignore")
    for seed in range(args.runs):
        # gp_minimize's own construction (optimizer/gp.py), kept open here so that the scoring times can be read back
        rng = np.random.RandomState(seed)
        space = skopt_b200.utils.normalize_dimensions(cfg["dims"])
        est = cook_estimator("GP", space=space, random_state=rng.randint(0, np.iinfo(np.int32).max), noise=1e-10)
        opt = Optimizer(space, est, n_initial_points=10, acq_func=cfg["acq_func"], acq_optimizer=cfg["acq_optimizer"],
                        random_state=rng, acq_optimizer_kwargs=dict(cfg["kwargs"], fit_device=args.fit_device))
        t0 = time.perf_counter()
        for _ in range(args.n_calls):
            x = opt.ask()
            res = opt.tell(x, cfg["func"](x))
        wall = time.perf_counter() - t0
        scoring = float(np.sum(opt.scoring_times_))
        vals = np.round(res.func_vals, 3)
        print(json.dumps({
            "config": args.config, "seed": seed, "n_calls": args.n_calls, "fit_device": args.fit_device,
            "abi_model": bool(args.abi_model), "wall_s": wall, "scoring_s": scoring, "other_s": wall - scoring,
            "tells_with_model": len(opt.scoring_times_), "scoring_p50_ms": 1e3 * float(np.median(opt.scoring_times_)),
            "fun": float(res.fun), "regret": float(res.fun - cfg["optimum"]), "calls_to_min": int(np.argmin(vals) + 1)}))


if __name__ == "__main__":
    main()
