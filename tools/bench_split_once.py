"""A few value-path passes in split-int8 mode at n_train=2048, d=20 over 75,776 candidates: the ncu target for K2o."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import skopt_b200  # noqa: F401
from bench import make_problem
from tools.bench_configs import timed

gpr, y_opt = make_problem(2048, 20)
dev = gpr.device_state(0)
X = torch.from_numpy(np.random.RandomState(1).rand(512 * 148, 20)).cuda()
def run():
    dev.score_tensors(X, y_opt=y_opt, acq_funcs=("EI",), top_k=16, precision="split_int8")
dev.set_timing(True); dev.get_timing()
ms = timed(run, 3)
tm = dev.get_timing()
print("split path: %d candidates in %.3f ms -> %.0f evals/s   k1 %.3f ms  k2o %.3f ms (per pass)" % (
    X.shape[0], ms, X.shape[0] / ms * 1e3, tm["k1"][0] / 4, tm["k2"][0] / 4))
