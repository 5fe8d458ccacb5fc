"""Multi-GPU plumbing: candidates are sharded by contiguous row blocks (rank r owns global rows
[offsets[r], offsets[r+1])), the GP state is replicated, and the only exchange is an all-gather of the per-rank
top-k (value, global index) pairs followed by a local lexicographic merge -- identical on every rank and identical
to np.argmin / np.argsort[:k] over the concatenated candidate set (optimizer.py:564,570).  NCCL has no MINLOC, so
all-gather + local reduce is the idiom.  Works on CUDA tensors (NCCL) and CPU tensors (gloo).
"""
import torch
import torch.distributed as dist

_I64_MAX = torch.iinfo(torch.int64).max


def merge_topk(vals, idxs, k):
    """k smallest (value, index) pairs of flat tensors; entries with index < 0 or NaN value are padding."""
    vals = vals.reshape(-1).clone()
    idxs = idxs.reshape(-1).clone()
    pad = (idxs < 0) | torch.isnan(vals)
    idxs[pad] = _I64_MAX
    vals[pad] = float("inf")
    order = torch.argsort(idxs, stable=True)
    order = order[torch.argsort(vals[order], stable=True)]
    order = order[:k]
    out_v, out_i = vals[order], idxs[order]
    gone = out_i == _I64_MAX
    out_v = torch.where(gone, torch.full_like(out_v, float("nan")), out_v)
    out_i = torch.where(gone, torch.full_like(out_i, -1), out_i)
    return out_v, out_i


def merge_topk_allgather(vals, idxs, k, group=None):
    """All-gather each rank's top-k list (k * 16 bytes per rank) and merge; every rank gets the global top-k."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_topk(vals, idxs, k)
    world = dist.get_world_size(group)
    gv = torch.empty(world * vals.numel(), dtype=vals.dtype, device=vals.device)
    gi = torch.empty(world * idxs.numel(), dtype=idxs.dtype, device=idxs.device)
    dist.all_gather_into_tensor(gv, vals.contiguous(), group=group)
    dist.all_gather_into_tensor(gi, idxs.contiguous(), group=group)
    return merge_topk(gv, gi, k)


def shard_bounds(m, world):
    """Contiguous row blocks: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(m, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b
