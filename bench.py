#!/usr/bin/env python
"""Headline benchmark: GP-posterior + EI candidate evaluations per second (BASELINE.json metric).

Workload (BASELINE.json configs[2]): synthetic GP, n_train=2048, d=20, Matern-5/2 with fixed
hyper-parameters, 1,000,000 candidates per GPU, FP64.  A step = one pass of the hot path over the candidate set:
posterior mean/std, -EI, and the top-k (value, index) pairs; with N GPUs every rank owns its own 1M candidates
(weak scaling, GP state replicated) and the only collective is the all-gather of the per-rank top-k (NCCL).

  value : evals/s, candidates resident in HBM (device entry point, CUDA-event timed, max over ranks), with the precision
          the product picks for this workload (engine.resolve_precision("auto"): split-integer contraction on tcgen05)
  e2e   : evals/s through the host entry point (pinned host candidates -> H2D -> kernels -> D2H of the top-k)
  roofline : dominant kernel of the headline path (K2o, tcgen05 kind::i8 digit-plane contraction) - algorithmic int8
          ops / its CUDA-event time against 2 x the measured bf16 peak; the pure FP64 DMMA path (K2) is timed in the same
          run and reported beside it (fp64_dmma_path) against the measured FP64 tensor peak
  parity_sample : this run's device state against the checker (oracle on the product's fitted state), with the
          reference's own re-ordering jitter beside it
  cpu_baseline : the CPU oracle (a port of the reference's NumPy/SciPy path) on a bounded sample, this host
  ask_p50_ms, ask_p50_ms_lbfgs : scoring-block latency of Optimizer.tell at BASELINE.json configs[0] / configs[1]

`--impl reference` times the reference algorithm (oracle port, einsum contraction exactly as gpr.py:332) on the host;
its cpu_baseline.all_cores adds what one single-threaded copy of that path per core reaches together.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, N_DIMS, M_PER_GPU, TOP_K = 2048, 20, 1_000_000, 16
FP64_PEAK_TFLOPS = 37.0   # measured DMMA issue peak on this pool's B200 (profiles/r01_fp64_peaks.txt); cuBLAS DGEMM: 35.5

def split_products(planes):
    """Digit-plane products the split-integer contraction keeps (p + q <= planes - 1): 28 for 7 planes, 21 for 6."""
    return planes * (planes + 1) // 2

METRIC = "GP-EI candidate evals/sec at n_train=2048,d=20"


def flops_value_path(n, d):
    """Algorithmic flops per candidate: kernel row + mean dot, whitened solve, epilogue."""
    return n * (3 * d + 12) + (n * n + 2 * n) + 40


def flops_k2(n):
    return n * n + 2 * n + 40


def int8_ops_k2(n, planes):
    """int8 operations (2 per MAC) the split-integer scheme needs per candidate: 28 (7 planes) or 21 (6 planes) plane
    products over the lower-triangular n(n+1)/2 entries of L_inv."""
    return split_products(planes) * n * (n + 1)


def measured_bf16_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"]), "MEASURED_PEAKS.json bf16_tflops (burst)"
    except Exception:
        return 1590.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def make_problem(n, d, seed=0):
    """SURVEY.md 8(d) recipe, built with the product's own estimator class (host-side fit, hyper-parameters fixed)."""
    from skopt_b200.learning import GaussianProcessRegressor
    from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern
    rng = np.random.RandomState(seed)
    X_train = rng.rand(n, d)
    w = rng.randn(d)
    y = np.sin(X_train @ w) + 0.1 * rng.randn(n)
    kernel = ConstantKernel(1.3, "fixed") * Matern(np.full(d, 1.5), "fixed", nu=2.5)
    gpr = GaussianProcessRegressor(kernel, normalize_y=True, noise=1e-4, optimizer=None).fit(X_train, y)
    return gpr, float(y.min())


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line): NVML polled every
    2 ms from a thread (the timed region of the default run is ~200 ms, too short for `nvidia-smi -lms`, which is kept
    as the fallback when the NVML binding is unavailable)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.lines, self.proc, self.nvml = index, [], None, None
        self.sm, self.smax, self.reasons, self.stop = [], None, set(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:                         # noqa: BLE001 - any NVML problem: fall back to nvidia-smi
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip() != ""]
        return int(ids[index]) if ids and all(v.strip().isdigit() for v in ids) and index < len(ids) else index

    def __enter__(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _poll(self):
        nv = self.nvml
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.reasons.update(nm for nm, bit in bits.items() if mask & bit)
            except Exception:                     # noqa: BLE001
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.nvml is not None:
            self.stop = True
            self.thread.join(timeout=2)
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        if self.nvml is not None and self.sm:
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.smax, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "sm_mhz_min": float(np.min(self.sm)), "sm_mhz_mean": float(np.mean(self.sm)),
                    "source": "nvml, 2 ms poll"}
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(self.NAMES, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


_CPU_STATE = {}


def library_versions():
    import scipy
    import sklearn
    return {"numpy": np.__version__, "scipy": scipy.__version__, "scikit-learn": sklearn.__version__}


def cpu_baseline(n, d, sample_m, quad="einsum", seed=0, reps=1):
    """Oracle (port of the reference's NumPy/SciPy path) timed on this host on `sample_m` candidates."""
    from oracle import gp_oracle as orc
    key = (n, d, seed)
    if key not in _CPU_STATE:
        _CPU_STATE[key] = orc.make_synthetic_state(n, d, noise=1e-4, seed=seed)
    st = _CPU_STATE[key]
    Xc = np.random.RandomState(seed + 1).rand(sample_m, d)
    y_opt = float(st.meta["y"].min())
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        v = orc.gaussian_acquisition(Xc, st, y_opt, "EI", acq_func_kwargs=dict(xi=0.01, kappa=1.96), quad=quad)
        int(np.argmin(v))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample_m / best, best


def _all_cores_worker(args):
    """One worker of cpu_all_cores: the reference path on its own slice of candidates, single-threaded."""
    n, d, sample_m, passes, seed, barrier = args
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:                     # noqa: BLE001
        pass
    cpu_baseline(n, d, sample_m, quad="einsum", seed=0)              # builds the state, warms the caches
    from oracle import gp_oracle as orc
    st = _CPU_STATE[(n, d, 0)]
    Xc = np.random.RandomState(1000 + seed).rand(sample_m, d)
    y_opt = float(st.meta["y"].min())
    barrier.wait(timeout=180)             # a worker that died must not leave the others (and the bench) waiting
    t0 = time.perf_counter()
    for _ in range(passes):
        v = orc.gaussian_acquisition(Xc, st, y_opt, "EI", acq_func_kwargs=dict(xi=0.01, kappa=1.96), quad="einsum")
        int(np.argmin(v))
    return time.perf_counter() - t0


def cpu_all_cores(n, d, sample_m, passes, workers):
    """The reference path has no candidate-level parallelism (einsum runs on one thread).  What the box's cores could do
    with it: `workers` independent processes, one per core, each scoring its own `sample_m` candidates `passes` times
    after a common barrier; aggregate candidates / slowest worker's time."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Manager() as manager:
        barrier = manager.Barrier(workers)
        with ctx.Pool(workers) as pool:
            job = pool.map_async(_all_cores_worker, [(n, d, sample_m, passes, w, barrier) for w in range(workers)])
            times = job.get(timeout=600)
    return workers * passes * sample_m / max(times)


def parity_sample(gpr, dev, y_opt, precisions, sample_m=512):
    """Parity of THIS run's device state against the checker (SURVEY.md 8d gate, in the units of the gate): the oracle
    is evaluated on the product's own fitted state with the reference's einsum contraction (gpr.py:332) and, for the
    'reference vs re-ordered reference' jitter the gate asks to be shown next to it, with a GEMM-ordered contraction."""
    from oracle import gp_oracle as orc
    st = orc.state_from_estimator(gpr)
    Xs = np.random.RandomState(7).rand(sample_m, st.d)
    mu_e, sd_e = orc.predict(st, Xs, return_std=True, quad="einsum")
    mu_g, sd_g = orc.predict(st, Xs, return_std=True, quad="gemm")
    ei_e, ei_g = -orc.ei_from_posterior(mu_e, sd_e, y_opt, 0.01), -orc.ei_from_posterior(mu_g, sd_g, y_opt, 0.01)
    var_scale = st.prior_var * st.y_std ** 2

    def diff(mu, sd, ei):
        return {"mean": float(np.max(np.abs(mu - mu_e)) / max(np.max(np.abs(mu_e)), st.y_std)),
                "var": float(np.max(np.abs(sd ** 2 - sd_e ** 2)) / var_scale),
                "acq": float(np.max(np.abs(ei - ei_e)) / np.max(np.abs(ei_e))),
                "argmin_equal": bool(int(np.argmin(ei)) == int(np.argmin(ei_e)))}

    out = {"candidates": sample_m, "gate": "1e-9 on each of mean / var / acq (units of SURVEY.md 8d), identical argmin",
           "reference_vs_reordered_reference": diff(mu_g, sd_g, ei_g)}
    for prec in precisions:
        r = dev.score(Xs, y_opt=y_opt, acq_funcs=("EI",), xi=0.01, kappa=1.96, top_k=1, precision=prec)
        d = diff(r["mean"], r["std"], r["EI"]["values"])
        d["argmin_equal"] = d["argmin_equal"] and int(r["EI"]["topk_idx"][0]) == int(np.argmin(ei_e))
        out[prec] = d
    return out


def branin(x):
    a, b, c, r, s, t = 1.0, 5.1 / (4 * np.pi ** 2), 5.0 / np.pi, 6.0, 10.0, 1.0 / (8 * np.pi)
    return a * (x[1] - b * x[0] ** 2 + c * x[0] - r) ** 2 + s * (1 - t) * np.cos(x[0]) + s


def ask_p50_ms(n_tells=31):
    """Median wall time of the candidate-scoring block (optimizer.py:550-608, fit excluded) over `n_tells` model-based
    tells of BASELINE.json configs[0]: gp_minimize on Branin, acq_optimizer='sampling', n_points=10000, gp_hedge."""
    from skopt_b200 import Optimizer
    opt = Optimizer([(-5.0, 10.0), (0.0, 15.0)], "GP", n_initial_points=10, acq_func="gp_hedge",
                    acq_optimizer="sampling", acq_optimizer_kwargs=dict(n_points=10000), random_state=0,
                    acq_func_kwargs=dict(xi=0.01, kappa=1.96))
    for _ in range(10 + n_tells - 1):
        x = opt.ask()
        opt.tell(x, branin(x))
    t = np.asarray(opt.scoring_times_[1:]) * 1e3          # drop the first (allocations, module load)
    return float(np.median(t)), int(t.size)


def hart6(x):
    alpha = np.asarray([1.0, 1.2, 3.0, 3.2])
    P = 1e-4 * np.asarray([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                           [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
    A = np.asarray([[10, 3, 17, 3.50, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8],
                    [17, 8, 0.05, 10, 0.1, 14]])
    return -np.sum(alpha * np.exp(-np.sum(A * (np.array(x) - P) ** 2, axis=1)))


def ask_p50_ms_lbfgs(n_tells=31):
    """Same measurement for BASELINE.json configs[1]: Hartmann-6, gp_hedge, acq_optimizer='lbfgs', 20 restarts
    (candidate scoring + 60 lock-step L-BFGS-B searches with batched GPU gradients)."""
    from skopt_b200 import Optimizer
    opt = Optimizer([(0.0, 1.0)] * 6, "GP", n_initial_points=10, acq_func="gp_hedge", acq_optimizer="lbfgs",
                    acq_optimizer_kwargs=dict(n_points=10000, n_restarts_optimizer=20), random_state=0)
    for _ in range(10 + n_tells - 1):
        x = opt.ask()
        opt.tell(x, hart6(x))
    t = np.asarray(opt.scoring_times_[1:]) * 1e3
    return float(np.median(t)), int(t.size)


def ask_p50_ms_cpu(n_tells=5):
    """Same block with the oracle port of the reference arithmetic (3 acquisitions, posterior recomputed each time)."""
    from oracle import gp_oracle as orc
    st = orc.make_synthetic_state(40, 2, noise=1e-6, seed=0)
    rng = np.random.RandomState(0)
    times = []
    for _ in range(n_tells):
        X = rng.rand(10000, 2)
        t0 = time.perf_counter()
        for acq in ("EI", "LCB", "PI"):
            v = orc.gaussian_acquisition(X, st, -1.0, acq, acq_func_kwargs=dict(xi=0.01, kappa=1.96), quad="einsum")
            int(np.argmin(v))
        times.append(time.perf_counter() - t0)
    return float(np.median(times)) * 1e3


def run_reference(args, rank):
    """Reference arm: the reference algorithm on the host cores (bounded sample per step)."""
    if rank != 0:
        return
    sample = args.ref_sample
    cores = os.cpu_count() or 1
    try:                                  # torchrun exports OMP_NUM_THREADS=1: give the BLAS parts the host's threads back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:                     # noqa: BLE001 - the einsum contraction dominates and is single-threaded anyway
        pass
    times = []
    for i in range(args.warmup + args.steps):
        evals_s, dt = cpu_baseline(N_TRAIN, N_DIMS, sample, quad="einsum", seed=0)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    try:
        workers = min(cores, 32)          # each worker holds its own copy of the fitted state (~150 MB at n = 2048)
        all_cores = {"value": cpu_all_cores(N_TRAIN, N_DIMS, sample, max(1, args.steps), workers), "unit": "evals/s",
                     "workers": workers, "how": "one single-threaded process per core, each scoring its own candidates "
                                              "with the same path; aggregate over the slowest worker"}
    except Exception as exc:              # noqa: BLE001 - an extra, never the reason the arm fails
        all_cores = {"value": None, "error": repr(exc)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic GP posterior + EI: n_train=2048, d=20, FP64; reference algorithm "
                               "(NumPy/SciPy: cdist-style distances, Matern-5/2, dgemv mean, einsum('ki,kj,ij->k') "
                               "variance as skopt gpr.py:332) on a bounded sample of %d candidates per step" % sample,
                   "n_train": N_TRAIN, "n_dims": N_DIMS, "candidates_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": "%d candidates per step; einsum contraction is single-threaded as in the "
                                   "reference, BLAS parts may use all %d host threads" % (sample, cores),
                         "all_cores": all_cores, "versions": library_versions()},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_JSON_FD = [None]


def reserve_stdout():
    """Keep the real stdout for the single JSON line and point fd 1 at stderr for everything else (NCCL prints its
    version banner to stdout from C, torch / cuSOLVER may warn): the driver reads ONE line from stdout."""
    sys.stdout.flush()
    _JSON_FD[0] = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD[0] is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD[0], data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=M_PER_GPU, help="candidates per GPU per step")
    ap.add_argument("--ref-sample", type=int, default=256, help="candidates per step of the reference arm")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="candidates of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ask", action="store_true", help="skip the ask() p50 latency leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    reserve_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import skopt_b200
    from skopt_b200 import engine
    from skopt_b200.distributed import merge_topk_allgather
    from skopt_b200.learning.gaussian_process import gpr as gpr_mod

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    gpr_mod.set_default_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's banner / debug lines ("NCCL version ...") go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    gpr, y_opt = make_problem(N_TRAIN, N_DIMS)
    dev = gpr.device_state(local_rank)
    m = args.m
    offset = rank * m
    Xc_host = torch.from_numpy(np.random.RandomState(1000 + rank).rand(m, N_DIMS)).pin_memory()
    Xc_dev = Xc_host.cuda(non_blocking=True)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(precision="auto"):
        res = dev.score_tensors(Xc_dev, y_opt=y_opt, acq_funcs=("EI",), xi=0.01, kappa=1.96, top_k=TOP_K,
                                index_offset=offset, precision=precision)
        return merge_topk_allgather(res["EI"]["topk_val"], res["EI"]["topk_idx"], TOP_K) if world > 1 else \
            (res["EI"]["topk_val"], res["EI"]["topk_idx"])

    def step_host(precision="auto"):
        res = dev.score(Xc_host.numpy(), y_opt=y_opt, acq_funcs=("EI",), xi=0.01, kappa=1.96, top_k=TOP_K,
                        index_offset=offset, want_posterior=False, want_values=False, precision=precision)
        if world > 1:
            v = torch.from_numpy(res["EI"]["topk_val"]).cuda()
            i = torch.from_numpy(res["EI"]["topk_idx"]).cuda()
            v, i = merge_topk_allgather(v, i, TOP_K)
            return v.cpu(), i.cpu()
        return res["EI"]["topk_val"], res["EI"]["topk_idx"]

    # ---------------------------------------------------------------- device-resident throughput
    def timed_dev(precision):
        for _ in range(args.warmup):
            step_dev(precision)
        dev.set_timing(True)
        dev.get_timing()
        barrier()
        l0 = engine.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            a.record(stream)
            for _ in range(args.steps):
                best_ = step_dev(precision)
            b.record(stream)
            barrier()
        n_launch = engine.launch_count() - l0
        tm = dev.get_timing()
        dev.set_timing(False)
        return a.elapsed_time(b), tm, n_launch, clk.summary(), int(best_[1][0].item())

    def timed_host(precision):
        for _ in range(2):
            step_host(precision)
        barrier()
        t0_ = time.perf_counter()
        for _ in range(args.steps):
            step_host(precision)
        barrier()
        return time.perf_counter() - t0_

    headline = engine.resolve_precision("auto", N_TRAIN, m)          # what the product picks for this workload
    ms_total, timing, launches, clocks_summary, argmin_idx = timed_dev(headline)
    e2e_s = timed_host(headline)
    planes, probe_diff = dev.split_info() if headline == "split_int8" else (0, -1.0)
    # single-GPU context only (under torchrun every rank would have to take this branch together)
    ms_p7 = timed_dev("split_int8_p7")[0] if headline == "split_int8" and planes == 6 and world == 1 else None
    ms_fp64, timing_fp64, _, _, argmin_fp64 = timed_dev("fp64") if headline != "fp64" else (ms_total, timing, 0, None, argmin_idx)
    e2e_fp64_s = timed_host("fp64") if headline != "fp64" else e2e_s

    t = torch.tensor([ms_total, e2e_s * 1e3, ms_fp64, e2e_fp64_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, ms_fp64, e2e_fp64_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    ms_per_step = ms_total / args.steps
    value = world * m / (ms_per_step / 1e3)
    e2e_value = world * m * args.steps / (e2e_ms / 1e3)

    if rank == 0:
        def k2_stats(tm, total_ms):
            k2_ms_, k2_n = tm["k2"]
            cand = m * args.steps / max(k2_n, 1)
            avg = k2_ms_ / max(k2_n, 1)
            return cand, avg, k2_ms_ / total_ms, tm["k1"][0] / total_ms

        cand_per_k2, k2_avg_ms, k2_share, k1_share = k2_stats(timing, ms_total)
        fp64_equiv = flops_k2(N_TRAIN) * cand_per_k2 / (k2_avg_ms / 1e3) / 1e12
        if headline == "split_int8":
            bf16_peak, bf16_src = measured_bf16_peak()
            peak = 2.0 * bf16_peak
            achieved = int8_ops_k2(N_TRAIN, planes) * cand_per_k2 / (k2_avg_ms / 1e3) / 1e12
            traffic_o = None
            prof_o = os.path.join(ROOT, "profiles", "k2o_dram_bytes_per_launch.json")
            if os.path.exists(prof_o):
                traffic_o = json.load(open(prof_o)).get("dram_bytes_per_candidate", 0.0) * cand_per_k2
            roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak,
                        "traffic": traffic_o,
                        "kernel": "ozaki_trmm_kernel<0, %d> (K2o: tcgen05.mma kind::i8 over %d s8 digit planes, exact s32 "
                                  "accumulation in TMEM, FP64 recombination)" % (planes, planes),
                        "digit_planes": planes,
                        "digit_planes_self_check": "6 planes when 256 probe candidates scored with FP64 and with 6 planes "
                                                   "agree to 1e-10 of the prior variance, else 7 (csrc/skb.cu "
                                                   "decide_split_planes); measured difference: %.2e" % probe_diff,
                        "seven_plane_evals_s": (world * m / (ms_p7 / args.steps / 1e3)) if ms_p7 else None,
                        "peak_source": "int8 dense = 2 x " + bf16_src + "; MEASURED_PEAKS.json has no int8 entry, the 2:1 "
                                       "int8:bf16 issue ratio is measured in profiles/r01_umma_i8_rates.txt",
                        "algorithmic_ops_per_candidate": int8_ops_k2(N_TRAIN, planes),
                        "algorithmic_ops_definition": "2 x %d digit-plane products x n(n+1)/2 (the scheme's own minimum at "
                                                      "%d planes)" % (split_products(planes), planes),
                        "issue_peak": 2 * 8175 * 148 * (clocks_summary.get("sm_mhz_mean") or clocks_summary.get("sm_mhz")
                                                         or 1965.0) / 1e6,
                        "issue_peak_definition": "2 x 8175 MAC/clk/SM (measured kind::i8 issue rate at N=256, "
                                                 "profiles/r01_umma_i8_rates.txt) x 148 SMs x mean SM clock during the timed region",
                        "fp64_equivalent_tflops": fp64_equiv,
                        "fp64_equivalent_vs_dmma_peak": fp64_equiv / FP64_PEAK_TFLOPS,
                        "candidates_per_launch": cand_per_k2, "avg_launch_ms": k2_avg_ms,
                        "k2_share_of_step": k2_share, "k1_share_of_step": k1_share,
                        "whole_path_fp64_equivalent_tflops": flops_value_path(N_TRAIN, N_DIMS) * world * m /
                                                              (ms_total / args.steps / 1e3) / world / 1e12}
            roofline["frac_of_issue_peak"] = roofline["achieved"] / roofline["issue_peak"]
        else:
            roofline = None
        c2, a2, s2, s1 = k2_stats(timing_fp64, ms_fp64)
        achieved64 = flops_k2(N_TRAIN) * c2 / (a2 / 1e3) / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "k2_dram_bytes_per_launch.json")
        if os.path.exists(prof):
            traffic = json.load(open(prof)).get("dram_bytes_per_candidate", 0.0) * c2
        roofline_fp64 = {"bound": "tensor", "achieved": achieved64, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": achieved64 / FP64_PEAK_TFLOPS, "traffic": traffic,
                         "kernel": "trmm_kernel<0> (K2: DMMA triangular contraction + acquisition epilogue)",
                         "peak_source": "measured FP64 DMMA issue peak on this pool (tools/microbench/fp64_peaks.cu -> "
                                        "profiles/r01_fp64_peaks.txt); MEASURED_PEAKS.json has no FP64 entry; cuBLAS DGEMM "
                                        "8192^3 measured 35.5",
                         "algorithmic_flops_per_candidate": flops_k2(N_TRAIN), "candidates_per_launch": c2,
                         "avg_launch_ms": a2, "k2_share_of_step": s2, "k1_share_of_step": s1}
        if roofline is None:
            roofline = roofline_fp64
        fp64_value = world * m / (ms_fp64 / args.steps / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64" if headline == "fp64" else "f64 (variance contraction: %d s8 digit planes -> exact s32 on tcgen05 "
                                                      "-> f64)" % planes,
            "data": "synthetic",
            "config": {"workload": "synthetic GP posterior + EI: n_train=2048, d=20, %d candidates per GPU, FP64 results "
                                   "(BASELINE.json configs[2])" % m,
                       "n_train": N_TRAIN, "n_dims": N_DIMS, "candidates_per_gpu": m, "top_k": TOP_K,
                       "precision": headline + " (engine.resolve_precision('auto'); 1e-9 parity gate: tests/test_gpu_split.py)",
                       "kernel": "1.3 * Matern(nu=2.5, length_scale=1.5) + noise 1e-4, normalize_y",
                       "sharding": "candidates sharded across ranks, GP state replicated, NCCL all-gather of top-k",
                       "l2_policy": "inputs larger than L2 (candidates %.0f MB + ~1 GiB cross-covariance scratch per "
                                    "chunk vs 126 MB L2)" % (m * N_DIMS * 8 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": m * N_DIMS * 8,
                    "d2h_bytes_per_step": TOP_K * 16 + 16},
            "gpu_launches": launches,
            "clocks": clocks_summary,
            "roofline": roofline,
            "fp64_dmma_path": {"value": fp64_value, "unit": "evals/s", "ms_per_step": ms_fp64 / args.steps,
                               "e2e": world * m * args.steps / (e2e_fp64_ms / 1e3), "roofline": roofline_fp64,
                               "whole_path_tflops": flops_value_path(N_TRAIN, N_DIMS) * fp64_value / world / 1e12,
                               "argmin_index": argmin_fp64},
            "argmin_index": argmin_idx,
        }
        # the ask-latency and CPU legs are single-GPU context for the headline: rank 0 at N=1 only
        if not args.no_ask and world == 1:
            p50, cnt = ask_p50_ms()
            p50_l, cnt_l = ask_p50_ms_lbfgs()
            line["ask_p50_ms_lbfgs"] = {"value": p50_l, "tells": cnt_l, "config": "BASELINE.json configs[1]: Hartmann-6, "
                                        "gp_hedge, lbfgs, 20 restarts, n_points=10000: scoring block incl. the 60 "
                                        "lock-step L-BFGS-B searches (reference, measured in SURVEY.md section 6: "
                                        "35.0 s / 40 tells = 875 ms)"}
            line["ask_p50_ms"] = {"value": p50, "tells": cnt, "config": "BASELINE.json configs[0]: Branin, gp_hedge, "
                                  "sampling, n_points=10000, n_train 10..39: scoring block of Optimizer._tell, fit "
                                  "excluded (candidate generation + one fused GPU pass + top-k + inverse transform)",
                                  "cpu_port_ms": ask_p50_ms_cpu() if not args.no_cpu_baseline else None}
        if not args.no_cpu_baseline and world == 1:
            evals_s, dt = cpu_baseline(N_TRAIN, N_DIMS, args.cpu_sample, quad="einsum")
            evals_s_gemm, _ = cpu_baseline(N_TRAIN, N_DIMS, 4 * args.cpu_sample, quad="gemm")
            line["cpu_baseline"] = {
                "value": evals_s, "unit": "evals/s", "cores": os.cpu_count() or 1, "kind": "port",
                "sample": "%d candidates of the same workload, %.1f s; oracle port with the reference's einsum "
                          "contraction (single-threaded as in the reference; BLAS parts on all host threads)"
                          % (args.cpu_sample, dt),
                "tuned_gemm_variant_evals_s": evals_s_gemm, "versions": library_versions()}
            line["parity_sample"] = parity_sample(gpr, dev, y_opt, sorted({headline, "fp64"}))
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
