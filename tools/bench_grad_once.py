"""Three gradient-path passes (K1 -> K2d -> K3g) at n_train=2048, d=20 over 37,888 candidates: the ncu target."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import skopt_b200  # noqa: F401
from bench import make_problem
from tools.bench_configs import grad_dev, timed

gpr, y_opt = make_problem(2048, 20)
dev = gpr.device_state(0)
X = torch.from_numpy(np.random.RandomState(1).rand(256 * 148, 20)).cuda()
run, keep = grad_dev(dev, X, y_opt, ("EI", "LCB"))
ms = timed(run, 2)
print("gradient path: %d candidates in %.3f ms -> %.0f evals/s" % (X.shape[0], ms, X.shape[0] / ms * 1e3))
