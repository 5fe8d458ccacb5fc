"""Where the first scoring call of a process spends its time: library load, CUDA context, state creation (kernel attribute
set-up loads every kernel it touches), first launches.  Usage: python tools/time_startup.py"""
import os
import sys
import time

t0 = time.perf_counter()
import numpy as np  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skopt_b200  # noqa: E402,F401
from skopt_b200 import _lib  # noqa: E402
from skopt_b200.learning import GaussianProcessRegressor  # noqa: E402
from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern  # noqa: E402

t1 = time.perf_counter()
lib = _lib.lib()
t2 = time.perf_counter()
rng = np.random.RandomState(0)
X = rng.rand(40, 2)
y = np.sin(X.sum(1))
gpr = GaussianProcessRegressor(ConstantKernel(1.0, "fixed") * Matern([1.0, 1.0], "fixed", nu=2.5), normalize_y=True,
                               noise=1e-3, optimizer=None).fit(X, y)
t3 = time.perf_counter()
dev = gpr.device_state(0)
t4 = time.perf_counter()
Xc = rng.rand(10000, 2)
r = dev.score(Xc, y_opt=float(y.min()), acq_funcs=("EI", "LCB", "PI"), top_k=1)
t5 = time.perf_counter()
r = dev.score(Xc, y_opt=float(y.min()), acq_funcs=("EI", "LCB", "PI"), top_k=1)
t6 = time.perf_counter()
g = dev.score_grad(Xc[:60], y_opt=float(y.min()), acq_funcs=("EI",))
t7 = time.perf_counter()
print("imports %.3f s | dlopen %.3f s | host fit %.3f s | first device_state (context + attributes + upload) %.3f s | "
      "first score %.3f s | second score %.4f s | first score_grad %.3f s" % (
          t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6))
