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


def _world(group):
    """Ranks the candidates are sharded over: `group=False` keeps the call local even inside an initialised process
    group (the Optimizer's default - collectives are opt-in), None is the default group."""
    if group is False or not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


def sharded_score(dev, X, y_opt, acq_funcs, xi, kappa, top_k, group=None, precision="fp64"):
    """Score the rows of X that belong to this rank and merge the top-k across ranks.

    `X` is the FULL candidate matrix (identical on every rank); the returned dict has, per acquisition, the GLOBAL
    'topk_val', 'topk_idx' and 'first_nan' -- the same on every rank and the same as a single-GPU run, because a
    candidate's value does not depend on where it was computed and ties resolve to the lowest global index."""
    world = _world(group)
    if world == 1:
        return dev.score(X, y_opt=y_opt, acq_funcs=acq_funcs, xi=xi, kappa=kappa, top_k=top_k,
                         want_posterior=False, want_values=False, precision=precision)
    rank = dist.get_rank(group)
    b = shard_bounds(X.shape[0], world)
    lo, hi = b[rank], b[rank + 1]
    # the precision is resolved on the FULL problem size so that every rank takes the same path
    from .engine import resolve_precision
    precision = resolve_precision(precision, dev.host.n_train, X.shape[0]) if hasattr(dev, "host") else precision
    res = dev.score(X[lo:hi], y_opt=y_opt, acq_funcs=acq_funcs, xi=xi, kappa=kappa, top_k=top_k, index_offset=lo,
                    want_posterior=False, want_values=False, precision=precision)
    backend = dist.get_backend(group)
    device = torch.device("cuda", dev.device) if backend == "nccl" else torch.device("cpu")
    for name in acq_funcs:
        r = res[name]
        v = torch.from_numpy(r["topk_val"]).to(device)
        i = torch.from_numpy(r["topk_idx"]).to(device)
        v, i = merge_topk_allgather(v, i, top_k, group=group)
        r["topk_val"], r["topk_idx"] = v.cpu().numpy(), i.cpu().numpy()
        nan = torch.tensor([r["first_nan"] if r["first_nan"] >= 0 else _I64_MAX], dtype=torch.int64, device=device)
        dist.all_reduce(nan, op=dist.ReduceOp.MIN, group=group)
        r["first_nan"] = -1 if int(nan) == _I64_MAX else int(nan)
    return res


def sharded_score_tensor(dev, X, y_opt, acq_funcs, xi, kappa, top_k, group=None, precision="fp64"):
    """sharded_score for a candidate matrix that already lives on this rank's GPU (a CUDA tensor, identical on every
    rank): no host copy of the candidates at all; only the k (value, index) pairs per acquisition come back."""
    from .engine import resolve_precision
    world = _world(group)
    rank = dist.get_rank(group) if world > 1 else 0
    m = X.shape[0]
    precision = resolve_precision(precision, dev.host.n_train, m)
    b = shard_bounds(m, world)
    lo, hi = b[rank], b[rank + 1]
    part = X[lo:hi] if world > 1 else X
    res = dev.score_tensors(part.contiguous(), y_opt=y_opt, acq_funcs=acq_funcs, xi=xi, kappa=kappa, top_k=top_k,
                            index_offset=lo, precision=precision)
    out = {"precision": precision}
    for name in acq_funcs:
        v, i, nan = res[name]["topk_val"], res[name]["topk_idx"], res[name]["first_nan"]
        if world > 1:
            v, i = merge_topk_allgather(v, i, top_k, group=group)
            nan = torch.where(nan < 0, torch.full_like(nan, _I64_MAX), nan)
            dist.all_reduce(nan, op=dist.ReduceOp.MIN, group=group)
            nan = torch.where(nan == _I64_MAX, torch.full_like(nan, -1), nan)
        out[name] = {"topk_val": v.cpu().numpy(), "topk_idx": i.cpu().numpy(), "first_nan": int(nan.cpu()[0])}
    return out
