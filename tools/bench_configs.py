"""Device-resident throughput of the value and gradient paths at the BASELINE.json config sizes (not the headline
bench; used to find the next kernel to optimise).  Usage: python tools/bench_configs.py [--quick]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import skopt_b200  # noqa: E402
from skopt_b200 import _lib, engine  # noqa: E402
from bench import make_problem, flops_value_path  # noqa: E402


def flops_grad_path(n, d):
    return n * (3 * d + 12) + (2 * n * n + 2 * n) + 6 * n * d + 60 * d


def grad_dev(dev, X, y_opt, acqs, precision="fp64"):
    m, d = X.shape
    dev._attach_kinv()
    prm = engine._acq_params(y_opt, 0.01, 1.96, acqs, 0, precision)
    out = _lib.GradOut()
    keep = []
    def buf(*shape):
        t = torch.empty(*shape, dtype=torch.float64, device=X.device); keep.append(t); return t.data_ptr()
    out.mean, out.std, out.mean_grad, out.std_grad = buf(m), buf(m), buf(m, d), buf(m, d)
    for a in acqs:
        i = _lib.ACQ_INDEX[a]
        out.acq[i] = buf(m); out.acq_grad[i] = buf(m, d)
    s = torch.cuda.current_stream()
    def run():
        _lib.check(_lib.lib().skb_score_grad_dev(dev._h, X.data_ptr(), m, C.byref(prm), C.byref(out), C.c_void_p(s.cuda_stream)))
    return run, keep


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--precision", default="fp64")
    args = ap.parse_args()
    cfgs = [("C1-like", 50, 2, 10000, 10000), ("C2-like", 200, 6, 10000, 60), ("C5", 1032, 10, 1000000, 4096),
            ("C3", 2048, 20, 1000000, 65536), ("C4", 4096, 50, 400000 if not args.quick else 100000, 100000)]
    for name, n, d, m_val, m_grad in cfgs:
        t0 = time.time()
        gpr, y_opt = make_problem(n, d)
        dev = gpr.device_state(0)
        t_state = time.time() - t0
        X = torch.from_numpy(np.random.RandomState(1).rand(m_val, d)).cuda()
        def val():
            dev.score_tensors(X, y_opt=y_opt, acq_funcs=("EI", "LCB"), top_k=16, precision=args.precision)
        dev.set_timing(True); dev.get_timing()
        ms = timed(val, 2)
        tm = dev.get_timing()
        row = {"config": name, "n": n, "d": d, "state_s": round(t_state, 2), "value": {
            "m": m_val, "ms": ms, "evals_per_s": m_val / ms * 1e3,
            "tflops": flops_value_path(n, d) * m_val / ms / 1e9,
            "k1_ms": tm["k1"][0] / 3, "k2_ms": tm["k2"][0] / 3, "topk_ms": tm["topk"][0] / 3}}
        Xg = X[:m_grad].contiguous()
        run, keep = grad_dev(dev, Xg, y_opt, ("EI", "LCB"), args.precision)
        dev.get_timing()
        ms = timed(run, 2)
        tm = dev.get_timing()
        row["grad"] = {"m": m_grad, "ms": ms, "evals_per_s": m_grad / ms * 1e3,
                       "tflops": flops_grad_path(n, d) * m_grad / ms / 1e9,
                       "k1_ms": tm["k1"][0] / 3, "k2d_ms": tm["k2"][0] / 3, "k3g_ms": tm["grad"][0] / 3}
        # latency of a small L-BFGS style batch (60 points) through the host entry point
        xs = np.random.RandomState(2).rand(60, d)
        dev.score_grad(xs, y_opt=y_opt, acq_funcs=("EI", "LCB", "PI"))
        t0 = time.perf_counter()
        for _ in range(20):
            dev.score_grad(xs, y_opt=y_opt, acq_funcs=("EI", "LCB", "PI"))
        row["grad_batch60_host_ms"] = (time.perf_counter() - t0) / 20 * 1e3
        t0 = time.perf_counter()
        for _ in range(20):
            dev.score_grad(xs[:1], y_opt=y_opt, acq_funcs=("EI",))
        row["grad_1pt_host_ms"] = (time.perf_counter() - t0) / 20 * 1e3
        xs = np.random.RandomState(3).rand(10000, d)
        dev.score(xs, y_opt=y_opt, acq_funcs=("EI", "LCB", "PI"), top_k=5, want_posterior=False, want_values=False)
        t0 = time.perf_counter()
        for _ in range(10):
            dev.score(xs, y_opt=y_opt, acq_funcs=("EI", "LCB", "PI"), top_k=5, want_posterior=False, want_values=False)
        row["score_10k_host_ms"] = (time.perf_counter() - t0) / 10 * 1e3
        print(json.dumps(row), flush=True)
        dev.close(); del X, Xg, keep
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
