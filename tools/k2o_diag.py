"""Which of {tensor pipe, epilogue drain, operand delivery} bounds K2o?  Times the 6-plane contraction kernel alone
(skb_debug_k2o) with parts switched off; prints ms per launch and SM clocks per 32-training-point step."""
import ctypes as C
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from skopt_b200 import _lib
if os.environ.get('SKB_LIB'):
    _lib.LIB_PATH = os.environ['SKB_LIB']           # A/B builds of the library (csrc/variants/)

gpr, y_opt = bench.make_problem(2048, 20)
dev = gpr.device_state(0)
L = _lib.lib()
tiles = 148 * 8
out = {}
for name, dbg in (("full", 0), ("full_timed", 8), ("no_epilogue", 1), ("no_copies", 2), ("mma_only", 3), ("mma_only_no_handover", 7), ("mma_only_no_barriers", 23), ("copies_no_handover_no_epilogue", 5), ("full_again", 0)):
    ms, cyc = C.c_double(0.0), C.c_double(0.0)
    with bench.ClockSampler(0) as clk:
        _lib.check(L.skb_debug_k2o(dev._h, dbg, tiles, 40, C.byref(ms), C.byref(cyc)))
    c = clk.summary()
    steps = 544 * 8          # 32-point steps per SM per launch (8 waves of 148 tiles)
    out[name] = {"ms": ms.value, "cta_cycles_per_step": cyc.value / 544, "implied_mhz": cyc.value * 8 / (ms.value * 1e-3) / 1e6 if cyc.value else None, "sm_mhz_mean": c.get("sm_mhz_mean"), "clk_per_step": ms.value * 1e-3 * (c.get("sm_mhz_mean") or 0) * 1e6 / steps}
    print(os.environ.get("SKB_TAG", ""), name, "ms %.4f  cta clk/step %.1f" % (ms.value, cyc.value / 544), flush=True)
json.dump(out, open("gpurun_out/r2_k2o_diag_%s.json" % os.environ.get("SKB_TAG", "default"), "w"), indent=1)
