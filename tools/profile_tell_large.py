"""Scoring block of Optimizer.tell at large sizes: n_train = 2048, d = 20, n_points = 1e6, fixed hyper-parameters
(optimizer=None) so that the host-side fit is just one Cholesky.  Host vs device candidate generation."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skopt_b200
from skopt_b200 import Optimizer
from skopt_b200.learning import GaussianProcessRegressor
from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern

n, d, m = 2048, 20, int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.RandomState(0)
X0 = rng.rand(n, d).tolist()
w = rng.randn(d)
y0 = (np.sin(np.asarray(X0) @ w) + 0.1 * rng.randn(n)).tolist()
for flag in (False, True):
    est = GaussianProcessRegressor(ConstantKernel(1.3, "fixed") * Matern(np.full(d, 1.5), "fixed", nu=2.5),
                                   normalize_y=True, noise=1e-4, optimizer=None)
    opt = Optimizer([(0.0, 1.0)] * d, est, n_initial_points=0, acq_func="EI", acq_optimizer="sampling", random_state=1,
                    acq_optimizer_kwargs=dict(n_points=m, device_candidates=flag))
    t0 = time.perf_counter()
    opt.tell(X0, y0)
    t_first = time.perf_counter() - t0
    nxt = []
    for _ in range(3):
        x = opt.ask()
        nxt.append(x)
        opt.tell(x, float(np.sin(np.asarray(x) @ w)))
    s = np.asarray(opt.scoring_times_) * 1e3
    print("device_candidates=%-5s scoring block: first %.1f ms, then %s ms  (whole first tell %.2f s)  x[0][:3]=%s" % (
        flag, s[0], np.round(s[1:], 1).tolist(), t_first, np.round(nxt[0][:3], 6).tolist()))
