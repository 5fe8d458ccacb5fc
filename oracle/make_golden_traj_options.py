"""Trajectory fixtures for the less travelled options of gp_minimize / Optimizer, from the UNMODIFIED reference: x0 / y0
hand-over, no random phase, explicit noise, a bounded model queue, xi / kappa, integer-only / log-uniform / categorical
spaces, and an Optimizer ask / tell script (batch ask in the random phase, tell(fit=False), constant liar with two
strategies, update_next, copy, per-second acquisitions).  All in sampling mode, where every point must be identical.
Run in the build container only.  TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
if not hasattr(np, "int"):
    np.int = int
import skopt  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests._option_scripts import BASE, RUNS, SCRIPT_CASES, plain, run_config, run_script  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "traj_options.json")


def main():
    warnings.simplefilter("ignore")
    rec = {"runs": RUNS, "base": BASE, "script_cases": SCRIPT_CASES, "results": {}, "scripts": []}
    for name, cfg in RUNS.items():
        rec["results"][name] = run_config(skopt, cfg)
        print(name, rec["results"][name]["x"])
    for acq, okw in SCRIPT_CASES:
        rec["scripts"].append(run_script(skopt, acq, okw))
        print("script", acq, okw)
    with open(OUT, "w") as f:
        json.dump(rec, f, default=plain, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
