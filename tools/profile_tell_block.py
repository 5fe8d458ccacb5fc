"""cProfile of the scoring block of a model-based tell() at BASELINE.json configs[0] (Branin, gp_hedge, sampling, 10 000
candidates): where the ~0.6 ms go on the host side.  Usage: python tools/profile_tell_block.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skopt_b200 import Optimizer  # noqa: E402
from bench import branin  # noqa: E402

opt = Optimizer([(-5.0, 10.0), (0.0, 15.0)], "GP", n_initial_points=10, acq_func="gp_hedge", acq_optimizer="sampling",
                acq_optimizer_kwargs=dict(n_points=10000), random_state=0)
for _ in range(30):
    x = opt.ask()
    opt.tell(x, branin(x))
est = opt.models[-1]
bounds = np.array(opt.space.transformed_bounds)
for _ in range(5):
    opt._score_and_refine(est, bounds)
t0 = time.perf_counter()
for _ in range(50):
    opt._score_and_refine(est, bounds)
print("scoring block, steady state: %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    opt._score_and_refine(est, bounds)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
