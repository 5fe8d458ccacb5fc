"""Multi-GPU path (needs >= 2 B200s): sharded scoring must return exactly the single-GPU answer."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import skopt_b200
from skopt_b200 import Optimizer
from skopt_b200.distributed import sharded_score
from skopt_b200.learning import GaussianProcessRegressor
from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
skopt_b200.set_default_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.RandomState(0)
X = rng.rand(150, 4); y = np.sin(X @ rng.randn(4))
gpr = GaussianProcessRegressor(ConstantKernel(1.3, "fixed") * Matern(np.full(4, 1.5), "fixed", nu=2.5),
                               normalize_y=True, noise=1e-4, optimizer=None).fit(X, y)
Xc = rng.rand(5001, 4); Xc[4000] = Xc[7]                      # an exact tie that straddles the shard boundary
dev = gpr.device_state(local)
full = dev.score(Xc, y_opt=float(y.min()), acq_funcs=("EI", "LCB", "PI"), top_k=8, want_posterior=False, want_values=False)
shard = sharded_score(dev, Xc, y_opt=float(y.min()), acq_funcs=("EI", "LCB", "PI"), xi=0.01, kappa=1.96, top_k=8)
for a in ("EI", "LCB", "PI"):
    assert np.array_equal(full[a]["topk_idx"], shard[a]["topk_idx"]), a
    assert np.array_equal(full[a]["topk_val"], shard[a]["topk_val"]), a
# an Optimizer driven identically on every rank proposes identical points
opt = Optimizer([(0.0, 1.0)] * 3, "GP", n_initial_points=4, acq_func="gp_hedge", acq_optimizer="sampling",
                acq_optimizer_kwargs=dict(n_points=3000, shard_group=True), random_state=5)
for _ in range(7):
    x = opt.ask(); opt.tell(x, float(np.sum((np.array(x) - 0.3) ** 2)))
pts = torch.tensor(opt.Xi, dtype=torch.float64, device="cuda")
ref = pts.clone(); dist.broadcast(ref, 0)
assert torch.equal(pts, ref)
rank = dist.get_rank()
# the peer-memory exchange (one kernel: NVLink stores + flags + merge) against the all-gather + merge-kernel path and the
# torch restatement, over many back-to-back exchanges (the two slot sets alternate) with changing record shapes
from skopt_b200.distributed import PeerExchange, check_exchanges, exchange_records, merge_records_torch
px = PeerExchange.for_group(None, torch.device("cuda", local))
print("rank", rank, "peer exchange", "USED" if px is not None else "UNAVAILABLE")
if px is not None:
    g = torch.Generator(device="cpu"); g.manual_seed(100 + rank)
    world = dist.get_world_size()
    for step in range(40):
        A, k = [(1, 16), (3, 8), (2, 1), (3, 64), (1, 170)][step %% 5]
        vals = torch.randn(A * k, generator=g, dtype=torch.float64)
        if step %% 7 == 3:
            vals[: k // 2 + 1] = 0.25                       # ties across ranks: the lowest global index wins
        idx = torch.randperm(10 ** 6, generator=g)[: A * k].to(torch.int64) * world + rank
        if step %% 9 == 4:
            idx[-1] = -1                                    # padding entry
        nan = torch.full((A,), -1, dtype=torch.int64)
        if step %% 4 == 1:
            nan[0] = 1000 * (rank + 1) + step
        rec = torch.cat([vals.view(torch.int64), idx, nan]).cuda()
        assert px.fits(A, k)
        fused = exchange_records(rec, A, k)
        gathered = torch.empty(world * rec.numel(), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(gathered, rec)
        ref = merge_records_torch(gathered.cpu(), world, A, k)
        assert torch.equal(fused.cpu(), ref), (step, A, k)
    check_exchanges()
    assert px.seq >= 40
    # a record that does not fit a peer slot (n_acq (2 k + 1) > 1024 words) takes the all-gather + merge-kernel path
    A, k = 1, 700
    assert not px.fits(A, k)
    vals = torch.randn(A * k, generator=g, dtype=torch.float64)
    idx = torch.randperm(10 ** 6, generator=g)[: A * k].to(torch.int64) * world + rank
    rec = torch.cat([vals.view(torch.int64), idx, torch.full((A,), -1, dtype=torch.int64)]).cuda()
    seq_before = px.seq
    big = exchange_records(rec, A, k)
    gathered = torch.empty(world * rec.numel(), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(gathered, rec)
    assert torch.equal(big.cpu(), merge_records_torch(gathered.cpu(), world, A, k)) and px.seq == seq_before
# without shard_group a tell() is local even inside the process group: only rank 0 runs it here, and it must not hang
if rank == 0:
    solo = Optimizer([(0.0, 1.0)] * 3, "GP", n_initial_points=4, acq_func="EI", acq_optimizer="sampling",
                     acq_optimizer_kwargs=dict(n_points=3000), random_state=5)
    for _ in range(6):
        x = solo.ask(); solo.tell(x, float(np.sum((np.array(x) - 0.3) ** 2)))
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_scoring_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29544", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
    print(out.stdout[-600:])
