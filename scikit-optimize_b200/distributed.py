"""Multi-GPU plumbing: candidates are sharded by contiguous row blocks (rank r owns global rows
[offsets[r], offsets[r+1])), the GP state is replicated, and the only exchange is ONE all-gather of a packed per-rank
record -- the top-k (value, global index) pairs of every requested acquisition plus its first-NaN index,
`n_acq * (2 k + 1)` 8-byte words (layout: include/skb.h, skb_topk_merge_records_dev) -- followed by one merge launch
(topk_merge_records_kernel) and one device -> host copy.  The result is identical on every rank and identical to
np.argmin / np.argsort[:k] over the concatenated candidate set (optimizer.py:564,570).  NCCL has no MINLOC, so
all-gather + local reduce is the idiom.  CUDA tensors travel over NCCL and merge on the device;
CPU tensors (gloo, the build machine's world-size-2 tests of this host logic) merge with the torch restatement
`merge_topk` below, which doubles as the independent checker of the device merge in bench.py's multi_gpu_check.
"""
import numpy as np
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
    """All-gather one (value, index) top-k list per rank and merge; every rank gets the global top-k.  Kept for callers
    that rank a single vector; the scoring paths below exchange packed records instead."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_topk(vals, idxs, k)
    rec = torch.cat([vals.contiguous().view(torch.int64), idxs.contiguous(),
                     torch.full((1,), -1, dtype=torch.int64, device=vals.device)])
    merged = exchange_records(rec, 1, vals.numel(), group=group, k_out=k)
    return merged[:k].view(torch.float64), merged[k:2 * k]


def merge_records_torch(gathered, n_records, n_acq, k, k_out=None):
    """torch restatement of topk_merge_records_kernel (CPU tensors of the gloo tests; checker of the device merge)."""
    k_out = k if k_out is None else k_out
    words = n_acq * (2 * k + 1)
    g = gathered.reshape(n_records, words)
    out = torch.empty(n_acq * (2 * k_out + 1), dtype=torch.int64, device=gathered.device)
    for a in range(n_acq):
        v, i = merge_topk(g[:, a * k:(a + 1) * k].contiguous().view(torch.float64),
                          g[:, n_acq * k + a * k:n_acq * k + (a + 1) * k], k_out)
        out[a * k_out:(a + 1) * k_out] = v.view(torch.int64)
        out[n_acq * k_out + a * k_out:n_acq * k_out + (a + 1) * k_out] = i
        nan = g[:, 2 * n_acq * k + a]
        nan = torch.where(nan < 0, torch.full_like(nan, _I64_MAX), nan).min()
        out[2 * n_acq * k_out + a] = -1 if int(nan) == _I64_MAX else nan
    return out


class PeerExchange(object):
    """Exchange + merge of the packed records as ONE kernel over NVLink peer memory (csrc/topk.cuh,
    topk_exchange_peer_kernel; include/skb.h, skb_topk_exchange_peer_dev): every rank stores its record straight into its
    peers' buffers, raises a flag, waits for the world's flags and merges - no NCCL call and no second launch on the step.
    torch's symmetric memory supplies the allocation and the peer mapping; the kernel only sees pointers.  One object per
    process group, created collectively on first use; `seq` counts the group's exchanges (every rank calls in step)."""

    _by_group = {}

    def __init__(self, group, device):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self._lib, self._C = _lib, C
        words = int(_lib.lib().skb_peer_buffer_bytes()) // 8
        self.buf = symm.empty(words, dtype=torch.int64, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, dist.group.WORLD if group is None else group)
        self.handle.barrier()                       # every rank's flags are zero before anybody's first store arrives
        self.grp = _lib.PeerGroup()
        self.grp.rank, self.grp.world = int(self.handle.rank), int(self.handle.world_size)
        for r, ptr in enumerate(self.handle.buffer_ptrs):
            self.grp.buffers[r] = int(ptr)
        self.device = device
        self.group = group                          # keeps the group object alive: for_group() keys its cache by id(group)
        self.seq = 0
        self.status = []                            # status words (device) of the exchanges not yet looked at

    def fits(self, n_acq, k):
        return self.grp.world <= 8 and n_acq * (2 * k + 1) <= 1024 and self.grp.world * (2 * k + 1) * 8 <= 48 * 1024

    def exchange(self, rec, n_acq, k, timeout_ms=5000):
        """merged record (device tensor, nothing waits).  The kernel's status word (a peer that did not arrive within
        `timeout_ms`) is kept on the device and looked at by check_exchanges() - unpack_record calls it after its own
        device -> host copy."""
        if len(self.status) >= 256:
            check_exchanges()
        self.seq += 1
        words = n_acq * (2 * k + 1)
        out = torch.empty(words + 1, dtype=torch.int64, device=rec.device)
        s = torch.cuda.current_stream(rec.device)
        self._lib.check(self._lib.lib().skb_topk_exchange_peer_dev(
            rec.device.index, self._C.byref(self.grp), self.seq, rec.contiguous().data_ptr(), int(n_acq), int(k),
            out.data_ptr(), int(timeout_ms), self._C.c_void_p(s.cuda_stream)))
        self.status.append(out[words:])
        return out[:words]

    @classmethod
    def for_group(cls, group, device):
        """The group's exchange object, or None where peer memory is not available (then NCCL's all-gather is used).
        SKB_PEER_EXCHANGE=0 switches it off."""
        import os
        key = id(group) if group is not None else None
        if key not in cls._by_group:
            px = None
            if os.environ.get("SKB_PEER_EXCHANGE", "1") != "0" and dist.get_backend(group) == "nccl":
                try:
                    px = cls(group, device)
                except Exception as exc:      # noqa: BLE001 - no P2P mapping on this machine: the collective path stays
                    import warnings
                    warnings.warn("peer-memory exchange unavailable (%s: %s); using the NCCL all-gather" % (type(exc).__name__, exc))
                    px = None
                # the decision must be the same on every rank
                ok = torch.tensor([1 if px is not None else 0], device=device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                if int(ok.item()) == 0:
                    px = None
            cls._by_group[key] = px
        return cls._by_group[key]


class PeerTimeout(RuntimeError):
    pass


def check_exchanges():
    """Raise PeerTimeout if any peer-memory exchange since the last call gave up waiting for a rank (its merged record
    is then meaningless).  One small device -> host copy per group; a no-op when no peer exchange ran."""
    for px in PeerExchange._by_group.values():
        if px is not None and px.status:
            st = torch.cat(px.status).cpu()
            px.status = []
            if bool(st.any()):
                raise PeerTimeout("a rank did not reach the peer-memory exchange in time; the merged record is invalid")


def exchange_records(rec, n_acq, k, group=None, k_out=None):
    """The one exchange of the path.  CUDA records over NCCL groups: ONE kernel that stores the record into every peer's
    symmetric buffer over NVLink and merges what arrived (PeerExchange); the returned tensor then carries a status word
    behind the record (0 = every rank arrived).  Otherwise: all-gather every rank's packed record and merge - CUDA
    records in one launch of the library's merge kernel, CPU records (gloo) through merge_records_torch."""
    world = _world(group)
    if world == 1 and (k_out is None or k_out == k):
        return rec
    if world > 1 and rec.is_cuda and (k_out is None or k_out == k):
        px = PeerExchange.for_group(group, rec.device)
        if px is not None and px.fits(n_acq, k):
            return px.exchange(rec, n_acq, k)
    if world > 1:
        gathered = torch.empty(world * rec.numel(), dtype=torch.int64, device=rec.device)
        dist.all_gather_into_tensor(gathered, rec.contiguous(), group=group)
    else:
        gathered = rec
    if rec.is_cuda and (k_out is None or k_out == k):
        from .engine import merge_records_device
        return merge_records_device(gathered, world, n_acq, k)
    return merge_records_torch(gathered, world, n_acq, k, k_out)


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


def _settle_split(dev, precision):
    """The 6-or-7-plane decision of precision="split_int8" is a property of the (replicated) state, but the library
    takes it on the first LARGE call - and a rank's shard can be below that size when the whole set is not.  Deciding up
    front on every rank keeps a candidate's value independent of the rank that scored it."""
    if precision == "split_int8" and hasattr(dev, "decide_split"):
        dev.decide_split()


def _pack_host_result(res, acq_funcs, k):
    A = len(acq_funcs)
    rec = np.empty(A * (2 * k + 1), dtype=np.int64)
    for j, name in enumerate(acq_funcs):
        r = res[name]
        rec[j * k:(j + 1) * k] = np.ascontiguousarray(r["topk_val"], dtype=np.float64).view(np.int64)
        rec[A * k + j * k:A * k + (j + 1) * k] = r["topk_idx"]
        rec[2 * A * k + j] = r["first_nan"]
    return rec


def sharded_score(dev, X, y_opt, acq_funcs, xi, kappa, top_k, group=None, precision="fp64"):
    """Score the rows of X that belong to this rank and merge the top-k across ranks.

    `X` is the FULL candidate matrix on the host (identical on every rank); the returned dict has, per acquisition, the
    GLOBAL 'topk_val', 'topk_idx' and 'first_nan' -- the same on every rank and the same as a single-GPU run, because a
    candidate's value does not depend on where it was computed and ties resolve to the lowest global index."""
    world = _world(group)
    if world == 1:
        return dev.score(X, y_opt=y_opt, acq_funcs=acq_funcs, xi=xi, kappa=kappa, top_k=top_k,
                         want_posterior=False, want_values=False, precision=precision)
    from .engine import resolve_precision, unpack_record
    rank = dist.get_rank(group)
    b = shard_bounds(X.shape[0], world)
    lo, hi = b[rank], b[rank + 1]
    # the precision is resolved on the FULL problem size so that every rank takes the same path
    precision = resolve_precision(precision, dev.host.n_train, X.shape[0]) if hasattr(dev, "host") else precision
    _settle_split(dev, precision)
    res = dev.score(X[lo:hi], y_opt=y_opt, acq_funcs=acq_funcs, xi=xi, kappa=kappa, top_k=top_k, index_offset=lo,
                    want_posterior=False, want_values=False, precision=precision)
    backend = dist.get_backend(group)
    device = torch.device("cuda", dev.device) if backend == "nccl" else torch.device("cpu")
    rec = torch.from_numpy(_pack_host_result(res, acq_funcs, top_k)).to(device)
    merged = unpack_record(exchange_records(rec, len(acq_funcs), top_k, group=group), acq_funcs, top_k)
    for name in acq_funcs:
        res[name].update(merged[name])
    return res


def sharded_score_tensor(dev, X, y_opt, acq_funcs, xi, kappa, top_k, group=None, precision="fp64", row_offset=None,
                         n_total=None):
    """sharded_score for candidates that already live on this rank's GPU: no host copy of the candidates at all; one
    packed record per rank is all-gathered, merged by one kernel, and comes back in one copy.

    Default: `X` is the full matrix (identical on every rank) and this rank scores its row block.  With `row_offset` /
    `n_total` given, `X` holds ONLY this rank's rows [row_offset, row_offset + len(X)) of an n_total-row set (a rank
    that generated just its own block)."""
    from .engine import resolve_precision, unpack_record
    world = _world(group)
    rank = dist.get_rank(group) if world > 1 else 0
    if row_offset is None:
        m = X.shape[0]
        b = shard_bounds(m, world)
        lo, hi = b[rank], b[rank + 1]
        part = X[lo:hi] if world > 1 else X
    else:
        m, lo, part = int(n_total), int(row_offset), X
    precision = resolve_precision(precision, dev.host.n_train, m)
    if world > 1:
        _settle_split(dev, precision)
    rec = dev.score_record(part.contiguous(), y_opt=y_opt, acq_funcs=acq_funcs, xi=xi, kappa=kappa, top_k=top_k,
                           index_offset=lo, precision=precision)
    out = unpack_record(exchange_records(rec, len(acq_funcs), top_k, group=group), acq_funcs, top_k)
    out["precision"] = precision
    return out
