"""Packed top-k records (include/skb.h: skb_topk_merge_records_dev): the device merge against the torch restatement and
against np.lexsort over the concatenated shards; padding of empty shards by skb_score_dev."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _records(rng, world, A, k, m_per_rank, ties=True):
    """Per-rank records built with NumPy from random value vectors (some exact ties, NaNs and an empty rank)."""
    vals = [rng.randn(A, m) for m in m_per_rank]
    if ties:
        for a in range(A):
            vals[0][a, 3] = vals[-1][a, 5] = -7.0            # the same minimum on two ranks: lowest global index wins
        r_nan = next(r for r in range(1, len(vals)) if vals[r].shape[1] > 2)   # a non-empty rank other than 0
        vals[r_nan][0, 2] = np.nan
    recs, offs = [], np.concatenate([[0], np.cumsum(m_per_rank)])
    for r, v in enumerate(vals):
        rec = np.empty(A * (2 * k + 1), dtype=np.int64)
        for a in range(A):
            m = v.shape[1]
            ok = ~np.isnan(v[a])
            order = np.lexsort((np.arange(m)[ok], v[a][ok]))[:k]
            tv = np.full(k, np.nan); ti = np.full(k, -1, dtype=np.int64)
            tv[:order.size] = v[a][ok][order]; ti[:order.size] = np.arange(m)[ok][order] + offs[r]
            nan = np.flatnonzero(~ok)
            rec[a * k:(a + 1) * k] = tv.view(np.int64)
            rec[A * k + a * k:A * k + (a + 1) * k] = ti
            rec[2 * A * k + a] = nan[0] + offs[r] if nan.size else -1
        recs.append(rec)
    return vals, np.concatenate(recs)


@pytest.mark.parametrize("world,A,k", [(2, 1, 1), (8, 3, 16), (5, 2, 7), (3, 3, 300)])
def test_device_merge_equals_numpy_and_torch(world, A, k):
    import torch
    from skopt_b200 import engine
    from skopt_b200.distributed import merge_records_torch
    rng = np.random.RandomState(world * 100 + k)
    m_per_rank = [int(x) for x in rng.randint(8, 400, size=world)]
    m_per_rank[world // 2] = 0 if world > 2 else m_per_rank[world // 2]       # an empty shard
    vals, gathered = _records(rng, world, A, k, m_per_rank)
    g = torch.from_numpy(gathered).cuda()
    merged = engine.merge_records_device(g, world, A, k)
    ref = merge_records_torch(torch.from_numpy(gathered), world, A, k)
    got = engine.unpack_record(merged, ("EI", "PI", "LCB")[:A], k)
    exp = engine.unpack_record(ref, ("EI", "PI", "LCB")[:A], k)
    for a, name in enumerate(("EI", "PI", "LCB")[:A]):
        full = np.concatenate([v[a] for v in vals])
        ok = ~np.isnan(full)
        order = np.lexsort((np.arange(full.size)[ok], full[ok]))[:k]
        idx = np.arange(full.size)[ok][order]
        assert got[name]["topk_idx"][:idx.size].tolist() == idx.tolist()
        assert np.array_equal(got[name]["topk_val"][:idx.size].view(np.int64), full[idx].view(np.int64))
        assert np.all(got[name]["topk_idx"][idx.size:] == -1)
        nan = np.flatnonzero(~ok)
        assert got[name]["first_nan"] == (int(nan[0]) if nan.size else -1)
        assert np.array_equal(got[name]["topk_idx"], exp[name]["topk_idx"])
        assert np.array_equal(got[name]["topk_val"].view(np.int64), exp[name]["topk_val"].view(np.int64))
        assert got[name]["first_nan"] == exp[name]["first_nan"]


def test_score_record_matches_score_tensors_and_pads_empty_shards():
    import torch
    from skopt_b200 import engine
    from skopt_b200.learning import GaussianProcessRegressor
    from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern
    rng = np.random.RandomState(3)
    X = rng.rand(120, 5); y = np.sin(X @ rng.randn(5))
    gpr = GaussianProcessRegressor(ConstantKernel(1.1, "fixed") * Matern(np.full(5, 0.9), "fixed", nu=2.5),
                                   normalize_y=True, noise=1e-4, optimizer=None).fit(X, y)
    dev = gpr.device_state(0)
    Xc = torch.from_numpy(rng.rand(3000, 5)).cuda()
    acqs = ("LCB", "EI")                                         # record slots follow this order
    res = dev.score_tensors(Xc, y_opt=float(y.min()), acq_funcs=acqs, top_k=6, index_offset=50)
    rec = engine.unpack_record(dev.score_record(Xc, y_opt=float(y.min()), acq_funcs=acqs, top_k=6, index_offset=50), acqs, 6)
    for name in acqs:
        assert np.array_equal(rec[name]["topk_idx"], res[name]["topk_idx"].cpu().numpy())
        assert np.array_equal(rec[name]["topk_val"], res[name]["topk_val"].cpu().numpy())
        assert rec[name]["first_nan"] == int(res[name]["first_nan"].cpu()[0]) == -1
    empty = engine.unpack_record(dev.score_record(Xc[:0], y_opt=0.0, acq_funcs=acqs, top_k=6), acqs, 6)
    for name in acqs:
        assert np.all(empty[name]["topk_idx"] == -1) and np.all(np.isnan(empty[name]["topk_val"]))
        assert empty[name]["first_nan"] == -1
    # shards of the same set, merged on the device, equal the unsharded top-k bit for bit
    cuts = [0, 700, 700, 1900, 3000]
    recs = [dev.score_record(Xc[a:b].contiguous(), y_opt=float(y.min()), acq_funcs=acqs, top_k=6, index_offset=a + 50)
            for a, b in zip(cuts[:-1], cuts[1:])]
    merged = engine.unpack_record(engine.merge_records_device(torch.cat(recs), 4, 2, 6), acqs, 6)
    for name in acqs:
        assert np.array_equal(merged[name]["topk_idx"], rec[name]["topk_idx"])
        assert np.array_equal(merged[name]["topk_val"].view(np.int64), rec[name]["topk_val"].view(np.int64))


def test_topk_dev_and_gradient_records_match_numpy():
    """skb_topk_dev ranks a value vector that already lives on the GPU; score_grad_tensors(top_k=...) uses it to pack the
    record of the acquisition values next to the gradients (bench.py c4_grad).  Against np.lexsort on the same values."""
    import torch
    from skopt_b200 import engine
    from oracle import gp_oracle as orc
    from tests.test_gpu_parity import host_state_from_oracle_state
    st = orc.make_synthetic_state(96, 5, noise=1e-4, seed=2)
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    rng = np.random.RandomState(5)
    X = rng.rand(3000, 5)
    X[7] = X[1500]                                          # an exact tie: the lower index must win
    acqs, k, off = ("EI", "LCB"), 12, 10_000
    out = dev.score_grad_tensors(torch.from_numpy(X).cuda(), y_opt=float(st.meta["y"].min()), acq_funcs=acqs, top_k=k,
                                 index_offset=off)
    torch.cuda.synchronize()
    rec = engine.unpack_record(out["record"], acqs, k)
    host = dev.score_grad(X, y_opt=float(st.meta["y"].min()), acq_funcs=acqs)
    for name in acqs:
        v = out[name]["values"].cpu().numpy()
        assert np.array_equal(v, host[name]["values"])                         # tensor and host entry points agree bit for bit
        assert np.array_equal(out[name]["grad"].cpu().numpy(), host[name]["grad"])
        order = np.lexsort((np.arange(v.size), v))[:k]
        assert rec[name]["topk_idx"].tolist() == (order + off).tolist()
        assert np.array_equal(rec[name]["topk_val"].view(np.int64), v[order].view(np.int64))
        assert rec[name]["first_nan"] == -1
    # a second call into the same output buffers (no allocation) gives the same record
    again = dev.score_grad_tensors(torch.from_numpy(X).cuda(), y_opt=float(st.meta["y"].min()), acq_funcs=acqs, top_k=k,
                                   index_offset=off, out=out)
    torch.cuda.synchronize()
    assert again is out and torch.equal(again["record"], out["record"])
    dev.close()


def test_measure_peaks_is_plausible():
    """skb_measure_peaks: the roofline denominators bench.py uses, measured on the device (tcgen05 kind::i8 issue rate at
    M=128, N=256 and the DMMA rate)."""
    from skopt_b200 import engine
    p = engine.measure_peaks(0)
    assert 7000.0 < p["i8_mac_per_clk_sm"] <= 8192.5                # 8192 MAC per clock per SM is the issue limit
    assert 2000.0 < p["i8_tops"] < 6000.0
    assert 20.0 < p["fp64_dmma_tflops"] < 45.0
