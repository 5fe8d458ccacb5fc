"""Host-buffer (e2e) step time of the headline workload against the chunk size (scratch limit): how much of the e2e gap is the
wait for the first chunk of the upload.  Usage: python tools/time_e2e_chunks.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

gpr, y_opt = bench.make_problem(bench.N_TRAIN, bench.N_DIMS)
dev = gpr.device_state(0)
m = bench.M_PER_GPU
X = torch.from_numpy(np.random.RandomState(1).rand(m, bench.N_DIMS)).pin_memory()
Xd = X.cuda()

for tiles in (0, 1776, 888, 444, 296, 148):
    dev.set_scratch_limit(tiles * 2048 * 128 * 8 if tiles else 0)
    for _ in range(3):
        dev.score(X.numpy(), y_opt=y_opt, acq_funcs=("EI",), top_k=16, want_posterior=False, want_values=False, precision="auto")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        dev.score(X.numpy(), y_opt=y_opt, acq_funcs=("EI",), top_k=16, want_posterior=False, want_values=False, precision="auto")
    t_host = (time.perf_counter() - t0) / 8 * 1e3
    for _ in range(2):
        dev.score_record(Xd, y_opt=y_opt, acq_funcs=("EI",), top_k=16, precision="auto")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        dev.score_record(Xd, y_opt=y_opt, acq_funcs=("EI",), top_k=16, precision="auto")
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / 8 * 1e3
    print("chunk %5s tiles: host buffers %.2f ms per step, device-resident %.2f ms, gap %.2f ms" % (tiles or "dflt", t_host, t_dev, t_host - t_dev), flush=True)
