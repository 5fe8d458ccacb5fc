"""K2o under the power cap with and without L2 -> SM operand traffic: the scratch is first filled with REAL digit planes (one
scoring call), then the contraction kernel alone runs for ~1.5 s per variant while NVML samples clock and power."""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from skopt_b200 import _lib

gpr, y_opt = bench.make_problem(2048, 20)
dev = gpr.device_state(0)
X = np.random.RandomState(1).rand(148 * 8 * 64, 20)
dev.score(X, y_opt=y_opt, acq_funcs=("EI",), top_k=1, precision="split_int8", want_posterior=False, want_values=False)
L = _lib.lib()
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for name, dbg in (("full", 8), ("cluster2", 102), ("cluster4", 104), ("operands_once", 32), ("full", 8), ("cluster2", 102), ("cluster4", 104)):
    ms, cyc = C.c_double(0.0), C.c_double(0.0)
    e0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    t0 = time.perf_counter()
    with bench.ClockSampler(0) as clk:
        _lib.check(L.skb_debug_k2o(dev._h, dbg, 148 * 8, int(os.environ.get('K2O_REPS', '800')), C.byref(ms), C.byref(cyc)))
    dt = time.perf_counter() - t0
    e1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    c = clk.summary()
    print("%-14s ms/launch %.4f  cta clk/step %.1f  mean clock %.0f MHz  power %.0f W  reasons %s" % (
        name, ms.value, cyc.value / 544, c.get("sm_mhz_mean") or 0, (e1 - e0) / 1e3 / dt, c.get("reasons")), flush=True)
