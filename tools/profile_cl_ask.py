"""cProfile of Optimizer.ask(n_points=3, 'cl_min') at n_train=1024, d=10, n_points=1e6 with the GPU fit objective."""
import cProfile
import os
import pstats
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skopt_b200 import Optimizer  # noqa: E402

n, d = 1024, 10
rng = np.random.RandomState(0)
X = rng.rand(n, d)
y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n)
warnings.simplefilter("ignore")
opt = Optimizer([(0.0, 1.0)] * d, "GP", n_initial_points=n, acq_func="EI", acq_optimizer="sampling",
                acq_optimizer_kwargs=dict(n_points=1000000, fit_device="gpu"), random_state=1)
opt.tell(X.tolist(), y.tolist())
opt.ask(n_points=2, strategy="cl_min")          # warm
opt.tell(opt.ask(), 0.1)
pr = cProfile.Profile()
pr.enable()
opt.ask(n_points=3, strategy="cl_min")
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
