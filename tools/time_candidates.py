import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import skopt_b200
from skopt_b200 import engine
from skopt_b200.space import normalize_dimensions, device_dim_descs
m, d = 1_000_000, 20
space = normalize_dimensions([(0.0, 1.0)] * d)
descs = device_dim_descs(space)
rng = np.random.RandomState(0)
def t(f, n=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("device_uniform 20M      %.2f ms" % t(lambda: engine.device_uniform(rng, m * d)))
print("device_candidates 1Mx20 %.2f ms" % t(lambda: engine.device_candidates(descs, m, rng)))
print("host rvs_transformed    %.2f ms" % t(lambda: space.rvs_transformed(m, rng), 1))
print("host rng.uniform 20M    %.2f ms" % t(lambda: rng.uniform(size=m * d), 1))
