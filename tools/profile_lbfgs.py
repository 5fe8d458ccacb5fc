"""Scoring-block latency of BASELINE.json configs[1]: Hartmann-6, gp_hedge, lbfgs, 20 restarts, n_points=10000."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skopt_b200
from skopt_b200 import Optimizer
from tests.test_gpu_optimizer import hart6

opt = Optimizer([(0.0, 1.0)] * 6, "GP", n_initial_points=10, acq_func="gp_hedge", acq_optimizer="lbfgs",
                acq_optimizer_kwargs=dict(n_points=10000, n_restarts_optimizer=20), random_state=0)
t_fit = []
for i in range(40):
    x = opt.ask()
    t0 = time.perf_counter()
    opt.tell(x, hart6(x))
    t_fit.append(time.perf_counter() - t0)
s = np.asarray(opt.scoring_times_[1:]) * 1e3
print("tells with a model: %d   scoring block p50 %.2f ms  p90 %.2f ms   whole tell p50 %.1f ms   best y %.4f" % (
    s.size, np.median(s), np.percentile(s, 90), np.median(t_fit[10:]) * 1e3, min(opt.yi)))
import cProfile, pstats
est = opt.models[-1]
pr = cProfile.Profile(); pr.enable()
opt._score_and_refine(est, np.array(opt.space.transformed_bounds))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
