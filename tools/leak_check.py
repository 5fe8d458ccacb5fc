"""Device / host memory across a long ask-tell loop (GPU fit, constant-liar batches, lbfgs refinement): free device memory
and the process RSS are sampled every 25 tells; a leak shows as a steady drift.  Usage: python tools/leak_check.py [tells]"""
import os
import sys
import warnings

import numpy as np
import psutil
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skopt_b200 import Optimizer  # noqa: E402
from skopt_b200.benchmarks import branin  # noqa: E402

warnings.simplefilter("ignore")
tells = int(sys.argv[1]) if len(sys.argv) > 1 else 200
proc = psutil.Process()
for label, kw in (("sampling + gpu fit", dict(acq_optimizer="sampling", acq_optimizer_kwargs=dict(n_points=10000, fit_device="gpu"))),
                  ("lbfgs + host fit", dict(acq_optimizer="lbfgs", acq_optimizer_kwargs=dict(n_points=10000, n_restarts_optimizer=5)))):
    opt = Optimizer([(-5.0, 10.0), (0.0, 15.0)], "GP", n_initial_points=5, acq_func="gp_hedge", random_state=0, **kw)
    samples = []
    for i in range(tells):
        if i % 40 == 39:
            for x in opt.ask(n_points=3, strategy="cl_min"):
                opt.tell(x, branin(x))
        else:
            x = opt.ask()
            opt.tell(x, branin(x))
        if i % 25 == 24:
            torch.cuda.synchronize()
            free, total = torch.cuda.mem_get_info(0)
            samples.append((i + 1, (total - free) / 2 ** 20, proc.memory_info().rss / 2 ** 20))
    print(label)
    for s in samples:
        print("  tells %4d  device used %8.1f MiB  host RSS %8.1f MiB" % s)
    del opt
