"""BASELINE.json configs[4]: Optimizer.ask(n_points=8, strategy='cl_min') at n_train=1024, d=10, 10^6 candidates per
liar step (8 sequential {lie -> fit -> score} rounds).  Usage: python tools/bench_cl.py [host|gpu] [n_points]
Under torchrun (one rank per GPU) every rank drives the same optimiser and each step's candidate set is sharded across
the ranks (acq_optimizer_kwargs shard_group=True: NCCL all-gather of the per-rank top-k); rank 0 prints the record."""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skopt_b200 import Optimizer  # noqa: E402

where = sys.argv[1] if len(sys.argv) > 1 else "gpu"
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
n, d = 1024, 10
rng = np.random.RandomState(0)
X = rng.rand(n, d)
w = rng.randn(d)
y = np.sin(X.dot(w)) + 0.1 * rng.randn(n)
warnings.simplefilter("ignore")
world, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist
    from skopt_b200 import set_default_device
    torch.cuda.set_device(local_rank)
    set_default_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
opt = Optimizer([(0.0, 1.0)] * d, "GP", n_initial_points=n, acq_func="EI", acq_optimizer="sampling",
                acq_optimizer_kwargs=dict(n_points=n_points, fit_device=where, shard_group=world > 1), random_state=1)
t0 = time.perf_counter()
opt.tell(X.tolist(), y.tolist())
t_tell = time.perf_counter() - t0
t0 = time.perf_counter()
batch = opt.ask(n_points=8, strategy="cl_min")
t_ask = time.perf_counter() - t0
if world > 1:
    points = [None] * world
    dist.all_gather_object(points, batch)
    assert all(p == points[0] for p in points), "ranks disagree on the batch"
    dist.destroy_process_group()
if int(os.environ.get("RANK", "0")) == 0:
    print(json.dumps({"config": "C5 constant liar", "n_gpus": world, "fit_device": where, "n_train": n, "d": d,
                      "n_points": n_points, "tell_s": t_tell, "ask8_s": t_ask,
                      "scoring_s_per_step": float(np.median(opt.scoring_times_)), "first_point": batch[0][:3]}))
