#!/usr/bin/env python
"""Headline benchmark: GP-posterior + EI candidate evaluations per second (BASELINE.json metric).

Workload (BASELINE.json configs[2]): synthetic GP, n_train=2048, d=20, Matern-5/2 with fixed
hyper-parameters, 1,000,000 candidates per GPU, FP64.  A step = one pass of the hot path over the candidate set:
posterior mean/std, -EI, and the top-k (value, index) pairs; with N GPUs every rank owns its own 1M candidates
(weak scaling, GP state replicated) and the only collective is ONE all-gather of a packed per-rank top-k record (NCCL).

  value : evals/s, candidates resident in HBM (device entry point, CUDA-event timed, max over ranks), with the precision
          the product picks for this workload (engine.resolve_precision("auto"): split-integer contraction on tcgen05);
          value_fp64 / value_p7: the strict FP64 DMMA path and the 7-plane split, timed in the same run
  e2e   : evals/s through the host entry point (pinned host candidates -> H2D -> kernels -> D2H of the top-k);
          e2e_pageable: the same call on a plain NumPy array (pageable memory, staged by skb_score_host)
  measured_peaks : tcgen05 kind::i8 issue rate and FP64 DMMA rate measured on this GPU in this run (skb_measure_peaks):
          the denominators of every roofline entry
  roofline : dominant kernel of the headline path (K2o, tcgen05 kind::i8 digit-plane contraction) - algorithmic int8
          ops / its CUDA-event time against the measured int8 peak (burst clock) and against the peak at the clock sampled
          during the timed region, plus frac_fp64_algorithmic; fp64_dmma_path.roofline: K2 on DMMA
  c4_grad : BASELINE.json configs[3] (n_train=4096, d=50, EI + LCB with gradients, 10^6 candidates per rank), K2d / K3g
          rooflines, parity_sample with gradients against the oracle's loop of single-point calls
  strong : the 10^6 candidates of configs[2] split over the N ranks (one tell()-equivalent), time per tell
  c5_cl : BASELINE.json configs[4]: ask(n_points=8, 'cl_min') at n_train=1024, d=10, 10^6 candidates per liar step
  multi_gpu_check : at N > 1, true iff every merged top-k record of this run equals, bit for bit, the top-k of the
          concatenated per-rank value vectors (computed with torch's stable sort, not with the product's kernels)
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


class Ctx:
    """What every leg needs: ranks, the CUDA stream the product launches on, barrier + max-over-ranks helpers."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.torch, self.dist = torch, dist
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def rank_max(self, *xs):
        t = self.torch.tensor([float(x) for x in xs], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        out = [float(v) for v in t.cpu()]
        return out[0] if len(out) == 1 else out

    def rank_all(self, flag):
        """True on every rank iff `flag` holds on every rank."""
        return self.rank_max(0.0 if flag else 1.0) == 0.0

    def timed(self, fn, steps, warmup):
        """CUDA-event time (ms) of `steps` calls of fn on the launching stream, barrier + synchronize on both sides."""
        for _ in range(warmup):
            fn()
        self.barrier()
        a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        last = None
        for _ in range(steps):
            last = fn()
        b.record(self.stream)
        self.barrier()
        return a.elapsed_time(b), last


def expected_record(values_by_acq, k):
    """The top-k of the CONCATENATED candidate set, computed independently of the product's kernels: torch's stable
    sort (ties -> lowest global index, NaNs last and excluded), packed in the record layout of include/skb.h."""
    import torch
    A = len(values_by_acq)
    rec = torch.empty(A * (2 * k + 1), dtype=torch.int64, device=values_by_acq[0].device)
    for j, v in enumerate(values_by_acq):
        sv, si = torch.sort(v, stable=True)
        nan = torch.isnan(v)
        good = int((~nan).sum().item())
        kk = min(k, good)
        tv = torch.full((k,), float("nan"), dtype=torch.float64, device=v.device)
        ti = torch.full((k,), -1, dtype=torch.int64, device=v.device)
        tv[:kk], ti[:kk] = sv[:kk], si[:kk]
        rec[j * k:(j + 1) * k] = tv.view(torch.int64)
        rec[A * k + j * k:A * k + (j + 1) * k] = ti
        rec[2 * A * k + j] = int(torch.nonzero(nan)[0].item()) if bool(nan.any()) else -1
    return rec


def gather_values(ctx, v):
    """All ranks' value vectors back to back (equal shard sizes)."""
    if ctx.world == 1:
        return v
    out = ctx.torch.empty(ctx.world * v.numel(), dtype=v.dtype, device=v.device)
    ctx.dist.all_gather_into_tensor(out, v.contiguous())
    return out


def leg_c4_grad(ctx, peaks):
    """BASELINE.json configs[3]: posterior + EI/LCB WITH input gradients at n_train=4096, d=50, `--m-grad` candidates per
    rank (10^6 -> 8 M at N=8), sharded with one packed top-k record exchanged per step (gpr.py:345-362,
    acquisition.py:305-319 evaluated for every candidate at once)."""
    import torch
    from skopt_b200 import engine
    from skopt_b200.distributed import exchange_records
    args, rank, world = ctx.args, ctx.rank, ctx.world
    n, d, m, acqs = 4096, 50, args.m_grad, ("EI", "LCB")
    t0 = time.perf_counter()
    gpr, y_opt = make_problem(n, d, seed=1)
    dev = gpr.device_state(ctx.local_rank)
    X = torch.from_numpy(np.random.RandomState(2000 + rank).rand(m, d)).cuda()
    setup_s = time.perf_counter() - t0
    res = {"config": "BASELINE.json configs[3]: n_train=%d, d=%d, EI + LCB with gradients, %d candidates per GPU, "
                     "NCCL arg-min (one packed top-%d record per rank)" % (n, d, m, TOP_K),
           "n_train": n, "n_dims": d, "candidates_per_gpu": m, "setup_s": setup_s}
    steps, warm = max(2, min(args.steps, 3)), 1
    buf = {"out": None}
    ok_all, merged_by_prec = True, {}
    for prec in (engine.resolve_precision("auto", n, m), "fp64"):
        if prec in merged_by_prec:
            continue

        def step():
            buf["out"] = dev.score_grad_tensors(X, y_opt=y_opt, acq_funcs=acqs, xi=0.01, kappa=1.96, precision=prec,
                                                top_k=TOP_K, index_offset=rank * m, out=buf["out"])
            return exchange_records(buf["out"]["record"], len(acqs), TOP_K)

        for _ in range(warm + 1):
            step()
        dev.set_timing(True)
        dev.get_timing()
        ms, merged = ctx.timed(step, steps, 0)
        tm = dev.get_timing()
        dev.set_timing(False)
        ms = ctx.rank_max(ms)
        merged_by_prec[prec] = merged
        # multi-GPU self-check on this leg's records: merged top-k == top-k of the concatenated value vectors, bit for bit
        exp = expected_record([gather_values(ctx, buf["out"][a]["values"]) for a in acqs], TOP_K)
        ok = ctx.rank_all(bool(torch.equal(exp, merged)))
        ok_all = ok_all and ok
        per = ms / steps
        k2_ms, k2_n = tm["k2"]
        g_ms, g_n = tm["grad"]
        k1_ms, _ = tm["k1"]
        entry = {"value": world * m / (per / 1e3), "unit": "evals/s", "ms_per_step": per, "steps": steps,
                 "topk_equals_concatenated_topk": ok,
                 "k1_share": k1_ms / ms, "k2d_share": k2_ms / ms, "k3g_share": g_ms / ms}
        cand_per = m * steps / max(k2_n, 1)
        if prec == "fp64":
            ach = 2.0 * n * n * cand_per / (k2_ms / max(k2_n, 1) / 1e3) / 1e12
            entry["roofline_k2d"] = {"bound": "tensor", "kernel": "trmm_kernel<1> (dense W = K_inv . K_trans^T on DMMA)",
                                     "achieved": ach, "peak": peaks["fp64_dmma_tflops"], "unit": "TFLOP/s",
                                     "frac": ach / peaks["fp64_dmma_tflops"], "algorithmic_flops_per_candidate": 2 * n * n,
                                     "candidates_per_launch": cand_per, "avg_launch_ms": k2_ms / max(k2_n, 1)}
        else:
            ops = 2.0 * split_products(7) * n * n
            ach = ops * cand_per / (k2_ms / max(k2_n, 1) / 1e3) / 1e12
            entry["roofline_k2d"] = {"bound": "tensor", "kernel": "ozaki_trmm_kernel<1, 7> (dense W = K_inv . K_trans^T: "
                                     "tcgen05.mma kind::i8 over 7 digit planes, FP64 recombination)",
                                     "achieved": ach, "peak": peaks["i8_tops"], "unit": "TOP/s", "frac": ach / peaks["i8_tops"],
                                     "algorithmic_ops_per_candidate": ops, "candidates_per_launch": cand_per,
                                     "avg_launch_ms": k2_ms / max(k2_n, 1),
                                     "fp64_equivalent_tflops": 2.0 * n * n * cand_per / (k2_ms / max(k2_n, 1) / 1e3) / 1e12}
            entry["roofline_k2d"]["frac_fp64_algorithmic"] = entry["roofline_k2d"]["fp64_equivalent_tflops"] / peaks["fp64_dmma_tflops"]
        fl_g = n * (3 * d + 12) + 6 * n * d + 2 * n + 60 * d
        ach_g = fl_g * (m * steps / max(g_n, 1)) / (g_ms / max(g_n, 1) / 1e3) / 1e12
        entry["roofline_k3g"] = {"bound": "tensor", "kernel": "grad_contract_kernel (K3g: FP64 pipe, DFMA + DMMA)",
                                 "achieved": ach_g, "peak": peaks["fp64_dmma_tflops"], "unit": "TFLOP/s",
                                 "frac": ach_g / peaks["fp64_dmma_tflops"], "algorithmic_flops_per_candidate": fl_g,
                                 "avg_launch_ms": g_ms / max(g_n, 1)}
        res[prec] = entry
    headline = engine.resolve_precision("auto", n, m)
    res["value"], res["unit"], res["precision"] = res[headline]["value"], "evals/s", headline
    res["multi_gpu_check"] = ok_all
    rec = engine.unpack_record(merged_by_prec[headline], acqs, TOP_K)
    res["argmin_index"] = {a: int(rec[a]["topk_idx"][0]) for a in acqs}
    if "fp64" in merged_by_prec and headline != "fp64":
        rec64 = engine.unpack_record(merged_by_prec["fp64"], acqs, TOP_K)
        res["argmin_equals_fp64_path"] = all(int(rec[a]["topk_idx"][0]) == int(rec64[a]["topk_idx"][0]) for a in acqs)
    if world == 1 and not args.no_cpu_baseline:
        # gradients against the loop of single-point reference calls (oracle), 48 candidates of this run's matrix
        from oracle import gp_oracle as orc
        st = orc.state_from_estimator(gpr)
        Xs = X[:48].cpu().numpy()
        mu, sd, gm, gs = orc.predict_with_grads_batched(st, Xs)
        par = {"candidates": 48, "gate": "1e-9 mean / var / mean_grad, 1e-8 std_grad (tests/test_gpu_parity.py)"}
        for prec in merged_by_prec:
            r = dev.score_grad(Xs, y_opt=y_opt, acq_funcs=acqs, xi=0.01, kappa=1.96, precision=prec)
            ei, gei = zip(*[orc.gaussian_acquisition(Xs[i:i + 1], st, y_opt, "EI", return_grad=True,
                                                     acq_func_kwargs=dict(xi=0.01, kappa=1.96)) for i in range(8)])
            par[prec] = {"mean": float(np.max(np.abs(r["mean"] - mu)) / max(np.max(np.abs(mu)), st.y_std)),
                         "var": float(np.max(np.abs(r["std"] ** 2 - sd ** 2)) / (st.prior_var * st.y_std ** 2)),
                         "mean_grad": float(np.max(np.abs(r["mean_grad"] - gm)) / np.max(np.abs(gm))),
                         "std_grad": float(np.max(np.abs(r["std_grad"] - gs)) / np.max(np.abs(gs))),
                         "ei_grad": float(np.max(np.abs(r["EI"]["grad"][:8] - np.asarray(gei).reshape(8, -1))) /
                                          max(np.max(np.abs(np.asarray(gei))), 1e-300))}
        res["parity_sample"] = par
    dev.close()
    del buf["out"], X
    torch.cuda.empty_cache()
    return res


def leg_strong(ctx, gpr, dev, y_opt, ms_one_gpu_step):
    """Strong scaling of ONE tell()-equivalent: the 10^6 candidates of configs[2] split over the N ranks
    (optimizer.py:552-564), one packed record exchanged; plus the multi-GPU self-check on that very set."""
    import torch
    from skopt_b200 import engine
    from skopt_b200.distributed import exchange_records, shard_bounds
    args, rank, world = ctx.args, ctx.rank, ctx.world
    m_total = args.m
    b = shard_bounds(m_total, world)
    lo, hi = b[rank], b[rank + 1]
    Xfull = np.random.RandomState(4242).rand(m_total, N_DIMS)            # the same set on every rank; rank r owns [lo, hi)
    X = torch.from_numpy(Xfull[lo:hi]).cuda()
    del Xfull
    prec = engine.resolve_precision("auto", N_TRAIN, m_total)             # resolved on the GLOBAL size
    if prec == "split_int8":
        dev.decide_split()
    vals = torch.empty(hi - lo, dtype=torch.float64, device="cuda")

    def step():
        rec = dev.score_record(X, y_opt=y_opt, acq_funcs=("EI",), xi=0.01, kappa=1.96, top_k=TOP_K, index_offset=lo,
                               precision=prec, values_out={"EI": vals})
        return exchange_records(rec, 1, TOP_K)

    ms, merged = ctx.timed(step, max(args.steps, 5), 3)
    ms = ctx.rank_max(ms) / max(args.steps, 5)
    check = None
    if m_total % world == 0:
        exp = expected_record([gather_values(ctx, vals)], TOP_K)
        check = ctx.rank_all(bool(torch.equal(exp, merged)))
    rec = engine.unpack_record(merged, ("EI",), TOP_K)
    return {"config": "configs[2] as one tell(): %d candidates in total, split over %d rank(s); n_train=%d, d=%d, EI, top-%d"
                      % (m_total, world, N_TRAIN, N_DIMS, TOP_K),
            "ms_per_tell": ms, "value": m_total / (ms / 1e3), "unit": "evals/s", "scaling": "strong",
            "candidates_per_rank": hi - lo, "precision": prec,
            "speedup_vs_one_gpu_step": (ms_one_gpu_step / ms) if ms_one_gpu_step else None,
            "one_gpu_ms": ms_one_gpu_step,
            "topk_equals_concatenated_topk": check, "argmin_index": int(rec["EI"]["topk_idx"][0])}


def leg_c5_cl(ctx):
    """BASELINE.json configs[4]: Optimizer.ask(n_points=8, strategy='cl_min') at n_train=1024, d=10, 10^6 candidates per
    liar step (optimizer.py:388-417: 8 sequential {lie -> refit -> score}); every liar step's candidates are sharded
    over the N ranks (one posterior replica per GPU, liar-updated together), fit objective on the GPU."""
    import warnings
    from skopt_b200 import Optimizer
    args, world = ctx.args, ctx.world
    n, d = 1024, 10
    rng = np.random.RandomState(0)
    X = rng.rand(n, d)
    y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n)
    out = {"config": "BASELINE.json configs[4]: ask(n_points=8, 'cl_min'), n_train=%d, d=%d, %d candidates per liar step "
                     "sharded over %d GPU(s), Matern-5/2 ARD + White refitted on every lie" % (n, d, args.m_cl, world),
           "n_gpus": world}
    # one-time, per (n_points, n_dims): the MT19937 jump polynomials of the device candidate generator (pure-Python big-integer
    # work that otherwise runs in a background thread and takes the GIL away from the first fits)
    from skopt_b200 import mt_jump
    mt_jump.column_jump_polynomials(2 * args.m_cl, d)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for mode in getattr(Optimizer, "LIAR_UPDATES", ("refit",)):
            kw = dict(n_points=args.m_cl, fit_device="gpu", shard_group=world > 1)
            if mode != "refit":
                kw["liar_update"] = mode
            opt = Optimizer([(0.0, 1.0)] * d, "GP", n_initial_points=n, acq_func="EI", acq_optimizer="sampling",
                            acq_optimizer_kwargs=kw, random_state=1)
            ctx.barrier()
            t0 = time.perf_counter()
            opt.tell(X.tolist(), y.tolist())
            ctx.barrier()
            t_tell = time.perf_counter() - t0
            t0 = time.perf_counter()
            batch = opt.ask(n_points=8, strategy="cl_min")
            ctx.barrier()
            t_ask = time.perf_counter() - t0
            t_tell, t_ask = ctx.rank_max(t_tell, t_ask)
            same = True
            if world > 1:
                pts = [None] * world
                ctx.dist.all_gather_object(pts, batch)
                same = all(p == pts[0] for p in pts)
            out[mode] = {"ask8_s": t_ask, "first_tell_s": t_tell, "ranks_agree_on_batch": same,
                         "scoring_s_per_liar_step": float(np.median(opt.scoring_times_[1:] or opt.scoring_times_)),
                         "batch_first_point": [float(v) for v in batch[0][:3]]}
    out["value"], out["unit"] = out["refit"]["ask8_s"], "s per ask(n_points=8)"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=M_PER_GPU, help="candidates per GPU per step")
    ap.add_argument("--m-grad", type=int, default=1_000_000, help="candidates per GPU of the configs[3] gradient leg")
    ap.add_argument("--m-cl", type=int, default=1_000_000, help="candidates per liar step of the configs[4] leg")
    ap.add_argument("--ref-sample", type=int, default=256, help="candidates per step of the reference arm")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="candidates of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ask", action="store_true", help="skip the ask() p50 latency leg")
    ap.add_argument("--legs", default="all", help="comma list of extra legs: c4,strong,c5,pageable (default: all)")
    args = ap.parse_args()
    legs = {"c4", "strong", "c5", "pageable"} if args.legs == "all" else {x for x in args.legs.split(",") if x}

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
    from skopt_b200.distributed import exchange_records
    from skopt_b200.learning.gaussian_process import gpr as gpr_mod

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    gpr_mod.set_default_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's banner / debug lines ("NCCL version ...") go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = Ctx(args, rank, world, local_rank)

    # roofline denominators measured on THIS device now (tcgen05 kind::i8 issue rate, DMMA rate): skb_measure_peaks
    peaks = engine.measure_peaks(local_rank)

    gpr, y_opt = make_problem(N_TRAIN, N_DIMS)
    dev = gpr.device_state(local_rank)
    m = args.m
    offset = rank * m
    Xc_np = np.random.RandomState(1000 + rank).rand(m, N_DIMS)           # pageable, what a NumPy user passes
    Xc_host = torch.from_numpy(Xc_np).pin_memory()
    Xc_dev = Xc_host.cuda(non_blocking=True)
    torch.cuda.synchronize()
    stream = ctx.stream
    barrier = ctx.barrier
    vals_dev = torch.empty(m, dtype=torch.float64, device="cuda")

    def step_dev(precision="auto", keep_values=False):
        rec = dev.score_record(Xc_dev, y_opt=y_opt, acq_funcs=("EI",), xi=0.01, kappa=1.96, top_k=TOP_K,
                               index_offset=offset, precision=precision, values_out={"EI": vals_dev} if keep_values else None)
        return exchange_records(rec, 1, TOP_K)

    def step_host(precision="auto", X=None):
        res = dev.score(Xc_host.numpy() if X is None else X, y_opt=y_opt, acq_funcs=("EI",), xi=0.01, kappa=1.96,
                        top_k=TOP_K, index_offset=offset, want_posterior=False, want_values=False, precision=precision)
        if world > 1:
            rec = np.concatenate([res["EI"]["topk_val"].view(np.int64), res["EI"]["topk_idx"],
                                  np.asarray([res["EI"]["first_nan"]], dtype=np.int64)])
            return exchange_records(torch.from_numpy(rec).cuda(), 1, TOP_K).cpu()
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
        return a.elapsed_time(b), tm, n_launch, clk.summary(), int(best_[TOP_K].item())

    def timed_host(precision, X=None):
        for _ in range(2):
            step_host(precision, X)
        barrier()
        t0_ = time.perf_counter()
        for _ in range(args.steps):
            step_host(precision, X)
        barrier()
        return time.perf_counter() - t0_

    headline = engine.resolve_precision("auto", N_TRAIN, m)          # what the product picks for this workload
    ms_total, timing, launches, clocks_summary, argmin_idx = timed_dev(headline)
    e2e_s = timed_host(headline)
    e2e_pageable_s = timed_host(headline, Xc_np) if "pageable" in legs else None
    planes, probe_diff = dev.split_info() if headline == "split_int8" else (0, -1.0)
    bound_var, bound_mean = dev.split_bounds() if headline == "split_int8" else (-1.0, -1.0)
    # every rank takes these branches together (the merge is a collective)
    planes_all = int(ctx.rank_max(planes))
    ms_p7 = timed_dev("split_int8_p7")[0] if headline == "split_int8" and planes_all == 6 else None
    ms_fp64, timing_fp64, _, _, argmin_fp64 = timed_dev("fp64") if headline != "fp64" else (ms_total, timing, 0, None, argmin_idx)
    e2e_fp64_s = timed_host("fp64") if headline != "fp64" else e2e_s

    ms_total, e2e_ms, ms_fp64, e2e_fp64_ms = ctx.rank_max(ms_total, e2e_s * 1e3, ms_fp64, e2e_fp64_s * 1e3)
    ms_p7 = ctx.rank_max(ms_p7) if ms_p7 is not None else None
    e2e_pageable_ms = ctx.rank_max(e2e_pageable_s * 1e3) if e2e_pageable_s is not None else None
    ms_per_step = ms_total / args.steps
    value = world * m / (ms_per_step / 1e3)
    e2e_value = world * m * args.steps / (e2e_ms / 1e3)

    # multi-GPU self-check (every N > 1): the merged top-k must equal the top-k of the CONCATENATED per-rank value vectors
    multi_gpu_check = None
    if world > 1:
        merged = step_dev(headline, keep_values=True)
        exp = expected_record([gather_values(ctx, vals_dev)], TOP_K)
        multi_gpu_check = ctx.rank_all(bool(torch.equal(exp, merged)))

    extra = {}

    def guarded(fn, *a):
        # an extra leg must never cost the headline line: a leg that raises (on every rank alike - the legs are
        # deterministic) is reported as {"error": ...} and the run goes on
        try:
            return fn(*a)
        except Exception as exc:          # noqa: BLE001
            return {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

    if "strong" in legs:
        # one GPU scoring the whole 10^6 set = this run's headline step when --m is the total
        extra["strong"] = guarded(leg_strong, ctx, gpr, dev, y_opt, ms_per_step)
        if world > 1 and extra["strong"].get("topk_equals_concatenated_topk") is not None:
            multi_gpu_check = bool(multi_gpu_check) and extra["strong"]["topk_equals_concatenated_topk"]
    if "c4" in legs:
        extra["c4_grad"] = guarded(leg_c4_grad, ctx, peaks)
        if world > 1:
            multi_gpu_check = bool(multi_gpu_check) and bool(extra["c4_grad"].get("multi_gpu_check"))
    if "c5" in legs:
        extra["c5_cl"] = guarded(leg_c5_cl, ctx)

    if rank == 0:
        def k2_stats(tm, total_ms):
            k2_ms_, k2_n = tm["k2"]
            cand = m * args.steps / max(k2_n, 1)
            avg = k2_ms_ / max(k2_n, 1)
            return cand, avg, k2_ms_ / total_ms, tm["k1"][0] / total_ms

        def traffic_from(profile, cand):
            path = os.path.join(ROOT, "profiles", profile)
            if not os.path.exists(path):
                return None, None
            rec = json.load(open(path))
            return rec.get("dram_bytes_per_candidate", 0.0) * cand, rec.get("source")

        run_clock = clocks_summary.get("sm_mhz_mean") or clocks_summary.get("sm_mhz") or 1965.0
        cand_per_k2, k2_avg_ms, k2_share, k1_share = k2_stats(timing, ms_total)
        fp64_equiv = flops_k2(N_TRAIN) * cand_per_k2 / (k2_avg_ms / 1e3) / 1e12
        if headline == "split_int8":
            peak = peaks["i8_tops"]
            achieved = int8_ops_k2(N_TRAIN, planes) * cand_per_k2 / (k2_avg_ms / 1e3) / 1e12
            traffic_o, traffic_src = traffic_from("r02_k2o_dram_bytes_per_launch.json", cand_per_k2)
            issue_peak_run = 2 * peaks["i8_mac_per_clk_sm"] * 148 * run_clock / 1e6
            roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak,
                        "traffic": traffic_o, "traffic_source": traffic_src,
                        "kernel": "ozaki_trmm_kernel<0, %d> (K2o: tcgen05.mma kind::i8 over %d s8 digit planes, exact s32 "
                                  "accumulation in TMEM, FP64 recombination)" % (planes, planes),
                        "digit_planes": planes,
                        "digit_planes_decision": "6 planes need BOTH the a-priori bound <= 2.5e-10 (this state: var %.2e, "
                                                 "mean %.2e) and the 256-probe difference to FP64 <= 1e-10 (measured %.2e); "
                                                 "else 7 (csrc/skb.cu decide_split_planes; no override)"
                                                 % (bound_var, bound_mean, probe_diff),
                        "peak_source": "skb_measure_peaks on this GPU in this run: tcgen05.mma kind::i8 M=128 N=256 issue "
                                       "rate, operands resident in shared memory, %.0f MAC/clk/SM, %.0f TOP/s at burst clock "
                                       "(MEASURED_PEAKS.json has no int8 entry)" % (peaks["i8_mac_per_clk_sm"], peaks["i8_tops"]),
                        "algorithmic_ops_per_candidate": int8_ops_k2(N_TRAIN, planes),
                        "algorithmic_ops_definition": "2 x %d digit-plane products x n(n+1)/2 (the scheme's own minimum at "
                                                      "%d planes)" % (split_products(planes), planes),
                        "peak_at_run_clock": issue_peak_run,
                        "frac_at_run_clock": achieved / issue_peak_run,
                        "peak_at_run_clock_definition": "2 x measured MAC/clk/SM x 148 SMs x mean SM clock sampled during "
                                                        "the timed region (%.0f MHz; the board sits at its power cap)" % run_clock,
                        "fp64_equivalent_tflops": fp64_equiv,
                        "frac_fp64_algorithmic": fp64_equiv / peaks["fp64_dmma_tflops"],
                        "frac_fp64_algorithmic_definition": "SURVEY.md 8(d): (n^2 + 2n + 40) FP64 flop per candidate / K2o "
                                                            "time, over the measured FP64 DMMA peak - above 1 because the "
                                                            "contraction does not run on the FP64 pipe; the FP64-pipe "
                                                            "figure is fp64_dmma_path.roofline",
                        "candidates_per_launch": cand_per_k2, "avg_launch_ms": k2_avg_ms,
                        "k2_share_of_step": k2_share, "k1_share_of_step": k1_share,
                        "whole_path_fp64_equivalent_tflops": flops_value_path(N_TRAIN, N_DIMS) * world * m /
                                                              (ms_total / args.steps / 1e3) / world / 1e12}
        else:
            roofline = None
        c2, a2, s2, s1 = k2_stats(timing_fp64, ms_fp64)
        achieved64 = flops_k2(N_TRAIN) * c2 / (a2 / 1e3) / 1e12
        traffic, traffic_src64 = traffic_from("k2_dram_bytes_per_launch.json", c2)
        roofline_fp64 = {"bound": "tensor", "achieved": achieved64, "peak": peaks["fp64_dmma_tflops"], "unit": "TFLOP/s",
                         "frac": achieved64 / peaks["fp64_dmma_tflops"], "traffic": traffic,
                         "kernel": "trmm_kernel<0> (K2: DMMA triangular contraction + acquisition epilogue)",
                         "peak_source": "skb_measure_peaks on this GPU in this run: mma.sync m8n8k4 f64 issue rate "
                                        "(MEASURED_PEAKS.json has no FP64 entry; cuBLAS DGEMM 8192^3 measured 35.5 in round 1)",
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
                       "sharding": "candidates sharded across ranks, GP state replicated, ONE NCCL all-gather of a packed "
                                   "top-k record per step, merged by one kernel",
                       "l2_policy": "inputs larger than L2 (candidates %.0f MB + ~1 GiB cross-covariance scratch per "
                                    "chunk vs 126 MB L2)" % (m * N_DIMS * 8 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": m * N_DIMS * 8,
                    "d2h_bytes_per_step": TOP_K * 16 + 16, "host_memory": "pinned"},
            "e2e_pageable": ({"value": world * m * args.steps / (e2e_pageable_ms / 1e3), "unit": "evals/s",
                              "host_memory": "pageable NumPy array (np.random.rand), staged through a pinned ring by "
                                             "skb_score_host"} if e2e_pageable_ms else None),
            "value_fp64": fp64_value,
            "value_p7": (world * m / (ms_p7 / args.steps / 1e3)) if ms_p7 else None,
            "value_definitions": "value: precision the product picks ('auto' -> %s, %d planes); value_fp64: strict FP64 DMMA "
                                 "path; value_p7: 7-plane split (55-bit operands)" % (headline, planes),
            "gpu_launches": launches,
            "clocks": clocks_summary,
            "measured_peaks": peaks,
            "roofline": roofline,
            "fp64_dmma_path": {"value": fp64_value, "unit": "evals/s", "ms_per_step": ms_fp64 / args.steps,
                               "e2e": world * m * args.steps / (e2e_fp64_ms / 1e3), "roofline": roofline_fp64,
                               "whole_path_tflops": flops_value_path(N_TRAIN, N_DIMS) * fp64_value / world / 1e12,
                               "argmin_index": argmin_fp64},
            "argmin_index": argmin_idx,
            "multi_gpu_check": multi_gpu_check,
        }
        line.update(extra)
        # the ask-latency and CPU legs are single-GPU context for the headline: rank 0 at N=1 only
        def ask_legs():
            p50, cnt = ask_p50_ms()
            p50_l, cnt_l = ask_p50_ms_lbfgs()
            return {"ask_p50_ms_lbfgs": {"value": p50_l, "tells": cnt_l, "config": "BASELINE.json configs[1]: Hartmann-6, "
                                         "gp_hedge, lbfgs, 20 restarts, n_points=10000: scoring block incl. the 60 "
                                         "lock-step L-BFGS-B searches (reference, measured in SURVEY.md section 6: "
                                         "35.0 s / 40 tells = 875 ms)"},
                    "ask_p50_ms": {"value": p50, "tells": cnt, "config": "BASELINE.json configs[0]: Branin, gp_hedge, "
                                   "sampling, n_points=10000, n_train 10..39: scoring block of Optimizer._tell, fit "
                                   "excluded (candidate generation + one fused GPU pass + top-k + inverse transform)",
                                   "cpu_port_ms": ask_p50_ms_cpu() if not args.no_cpu_baseline else None}}

        def cpu_legs():
            evals_s, dt = cpu_baseline(N_TRAIN, N_DIMS, args.cpu_sample, quad="einsum")
            evals_s_gemm, _ = cpu_baseline(N_TRAIN, N_DIMS, 4 * args.cpu_sample, quad="gemm")
            return {"cpu_baseline": {
                        "value": evals_s, "unit": "evals/s", "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": "%d candidates of the same workload, %.1f s; oracle port with the reference's einsum "
                                  "contraction (single-threaded as in the reference; BLAS parts on all host threads)"
                                  % (args.cpu_sample, dt),
                        "tuned_gemm_variant_evals_s": evals_s_gemm, "versions": library_versions()},
                    "parity_sample": parity_sample(gpr, dev, y_opt, sorted({headline, "fp64"}))}

        # the ask-latency and CPU legs are single-GPU context for the headline: rank 0 at N=1 only
        if not args.no_ask and world == 1:
            res = guarded(ask_legs)
            line.update(res if "error" not in res else {"ask_p50_ms": res})
        if not args.no_cpu_baseline and world == 1:
            res = guarded(cpu_legs)
            line.update(res if "error" not in res else {"cpu_baseline": res})
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
