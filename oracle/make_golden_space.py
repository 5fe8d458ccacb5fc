"""Fixtures for candidate generation: reference `space.transform(space.rvs(n, rng))` (optimizer.py:552-553) and
`space.rvs` / `inverse_transform`, from the UNMODIFIED reference.  Run in the build container only.
TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
if not hasattr(np, "int"):
    np.int = int
from skopt.space import Space  # noqa: E402
from skopt.utils import normalize_dimensions  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
SPECS = {
    "real2": [(-5.0, 10.0), (0.0, 15.0)],
    "real6": [(0.0, 1.0)] * 6,
    "mixed": [(-2.0, 3.5), (1, 17), (1e-4, 1e2, "log-uniform"), ["a", "b", "c", "d"], (0, 1), [True, False]],
    "intlog": [(1, 1000, "log-uniform"), (2.0, 64.0, "log-uniform", 2), (-3, 3)],
}


def main():
    rec = {}
    for name, spec in SPECS.items():
        space = normalize_dimensions(spec)
        rng = np.random.RandomState(42)
        pts = space.rvs(257, random_state=rng)
        Xt = space.transform(pts)
        after = rng.randint(0, 2 ** 31 - 1)        # the stream position after the draw must match too
        back = space.inverse_transform(Xt[:20])
        rec[name + "_Xt"] = Xt
        rec[name + "_after"] = np.int64(after)
        rec[name + "_pts"] = np.array(json.dumps(pts[:20], default=lambda o: o.item() if hasattr(o, "item") else str(o)))
        rec[name + "_back"] = np.array(json.dumps(back, default=lambda o: o.item() if hasattr(o, "item") else str(o)))
        rec[name + "_bounds"] = np.array(space.transformed_bounds, dtype=float)
    rec["specs"] = np.array(json.dumps(SPECS))
    np.savez_compressed(os.path.join(OUT, "traj_space_candidates.npz"), **rec)
    print("wrote traj_space_candidates.npz")


if __name__ == "__main__":
    main()
