"""Candidate generation at the headline size (10^6 x 20): serial MT19937 walk against the per-dimension jump generator,
and the host path.  Usage: python tools/time_candidates.py [m d]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import skopt_b200  # noqa: E402,F401
from skopt_b200 import engine, mt_jump  # noqa: E402
from skopt_b200.space import device_dim_descs, normalize_dimensions  # noqa: E402

m, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1_000_000, 20)
space = normalize_dimensions([(0.0, 1.0)] * d)
descs = device_dim_descs(space)
rng = np.random.RandomState(0)


def t(f, n=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3


t0 = time.perf_counter()
mt_jump.column_jump_polynomials(2 * m, d)
row = {"m": m, "d": d, "jump_polynomials_once_s": time.perf_counter() - t0}
a, b = np.random.RandomState(1), np.random.RandomState(1)
row["bit_identical"] = bool(torch.equal(engine.device_candidates(descs, m, a, jump=False),
                                        engine.device_candidates(descs, m, b, jump=True)))
row["serial_ms"] = t(lambda: engine.device_candidates(descs, m, rng, jump=False))
row["jump_ms"] = t(lambda: engine.device_candidates(descs, m, rng, jump=True))
row["host_rvs_transformed_ms"] = t(lambda: space.rvs_transformed(m, rng), 1)
print(json.dumps(row))
