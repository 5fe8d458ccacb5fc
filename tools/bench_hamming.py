"""Device-resident throughput of the all-categorical value path (SKB_KERNEL_HAMMING) at the headline size: n_train = 2048
label-encoded rows of 20 categorical dimensions, 10^6 candidates, FP64 DMMA and split-integer contractions.
Usage: python tools/bench_hamming.py [n d m]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import skopt_b200  # noqa: E402,F401
from skopt_b200.learning import GaussianProcessRegressor  # noqa: E402
from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, HammingKernel  # noqa: E402

n, d, m = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2048, 20, 1_000_000)
rng = np.random.RandomState(0)
n_cats = rng.randint(3, 9, size=d)
X = np.column_stack([rng.randint(0, k, size=n) / (k - 1.0) for k in n_cats])
w = rng.uniform(0.05, 0.6, size=d)
y = np.sin(3.0 * X @ w) + 0.1 * rng.randn(n)
gpr = GaussianProcessRegressor(ConstantKernel(1.3, "fixed") * HammingKernel(w, "fixed"), normalize_y=True, noise=1e-3,
                               optimizer=None).fit(X, y)
dev = gpr.device_state(0)
Xc = torch.from_numpy(np.column_stack([rng.randint(0, k, size=m) / (k - 1.0) for k in n_cats])).cuda()
row = {"config": "Hamming value path", "n": n, "d": d, "m": m}
ref = None
for precision in ("fp64", "split_int8_p7", "split_int8"):
    def step():
        return dev.score_tensors(Xc, y_opt=float(y.min()), acq_funcs=("EI",), top_k=16, precision=precision)
    for _ in range(2):
        res = step()
    dev.set_timing(True); dev.get_timing()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        res = step()
    e1.record(); torch.cuda.synchronize()
    tm = dev.get_timing(); dev.set_timing(False)
    ms = e0.elapsed_time(e1) / 3
    idx = res["EI"]["topk_idx"].cpu().numpy()
    ref = idx if ref is None else ref
    row[precision] = {"ms": ms, "evals_per_s": m / ms * 1e3, "k1_ms": tm["k1"][0] / 3, "k2_ms": tm["k2"][0] / 3,
                      "argmin": int(idx[0]), "top16_equal_fp64": bool(np.array_equal(idx, ref))}
row["split_info"] = dev.split_info()
print(json.dumps(row))
