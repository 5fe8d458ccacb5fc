"""A whole model-based tell() at BASELINE.json configs[2] size - n_train = 2048, d = 20, n_points = 10^6 - with the default
cooked estimator (ARD Matern-5/2 + White, 2 optimiser restarts) and everything on the GPU: hyper-parameter search,
final factorisation, scoring-state build, candidate generation, scoring, arg-min.  Usage: python tools/bench_tell_c3.py [n d m]"""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skopt_b200 import Optimizer  # noqa: E402

n, d, m = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2048, 20, 1_000_000)
rng = np.random.RandomState(0)
X = rng.rand(n, d)
w = rng.randn(d)
y = np.sin(X.dot(w)) + 0.1 * rng.randn(n)
warnings.simplefilter("ignore")
opt = Optimizer([(0.0, 1.0)] * d, "GP", n_initial_points=n, acq_func="EI", acq_optimizer="sampling",
                acq_optimizer_kwargs=dict(n_points=m, fit_device="gpu"), random_state=1)
t0 = time.perf_counter()
opt.tell(X.tolist(), y.tolist())
first = time.perf_counter() - t0
tells = []
for _ in range(4):
    x = opt.ask()
    t0 = time.perf_counter()
    opt.tell(x, float(np.sin(np.dot(x, w))))
    tells.append(time.perf_counter() - t0)
if os.environ.get("SKB_PROFILE"):
    import cProfile
    import pstats
    x = opt.ask()
    pr = cProfile.Profile()
    pr.enable()
    opt.tell(x, float(np.sin(np.dot(x, w))))
    pr.disable()
    pstats.Stats(pr, stream=sys.stderr).sort_stats("cumtime").print_stats(28)
dev = opt.models[-1].device_state()
print(json.dumps({"config": "tell() at n_train=%d, d=%d, n_points=%d, cooked GP, fit_device=gpu" % (n, d, m),
                  "first_tell_s": first, "tell_s": tells, "scoring_block_s": [float(v) for v in opt.scoring_times_],
                  "split_info": dev.split_info(), "next_x_head": [float(v) for v in opt.ask()[:3]]}))
