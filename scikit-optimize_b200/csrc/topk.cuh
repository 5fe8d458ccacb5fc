// K3/K4: deterministic top-k of an acquisition vector by (value, global index), ties -> lowest index.
// Element 0 is np.argmin's "first minimum" (optimizer.py:564); the k entries are argsort[:k] (optimizer.py:570)
// whenever values are distinct.  NaN never wins here (argsort semantics); the lowest NaN index is reported
// separately so the host can reproduce np.argmin's "first NaN" rule.
//
// Two launches: every block extracts the k smallest of its slice with k rounds of a block-wide lexicographic
// arg-min (warp shuffles + one shared-memory hop), then a single block merges the per-block lists the same way.
#pragma once
#include <cstdint>
#include <math.h>
#include <cuda_runtime.h>

namespace skb {

constexpr int TOPK_THREADS = 256;
constexpr int64_t IDX_NONE = INT64_MAX;

struct ValIdx {
    double v;
    int64_t i;
};

__device__ __forceinline__ bool vi_less(const ValIdx& a, const ValIdx& b) {
    return (a.v < b.v) || (a.v == b.v && a.i < b.i);
}

__device__ __forceinline__ ValIdx vi_warp_min(ValIdx x) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ValIdx y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, off);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, off);
        if (y.i != IDX_NONE && (x.i == IDX_NONE || vi_less(y, x))) x = y;
    }
    return x;
}

// k rounds over `count` entries.  idx_in == nullptr: entry e has global index base + e.
// seg_len > 0: the entries are `seg_len`-long runs that start `seg_stride` words apart (the per-rank records of
// topk_merge_records_kernel): entry e lives at (e / seg_len) * seg_stride + e % seg_len.
__device__ void block_topk_rounds(const double* __restrict__ vals, const int64_t* __restrict__ idx_in, int64_t count,
                                  int64_t base, int k, double* out_v, int64_t* out_i, int64_t* first_nan,
                                  int seg_len = 0, int64_t seg_stride = 0) {
    __shared__ ValIdx s_w[TOPK_THREADS / 32];
    __shared__ ValIdx s_win;
    __shared__ long long s_nan[TOPK_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    ValIdx prev{0.0, IDX_NONE};
    for (int r = 0; r < k; ++r) {
        ValIdx best{0.0, IDX_NONE};
        long long nan_i = INT64_MAX;
        for (int64_t e = tid; e < count; e += TOPK_THREADS) {
            const int64_t at = seg_len > 0 ? (e / seg_len) * seg_stride + e % seg_len : e;
            ValIdx cur{vals[at], idx_in ? idx_in[at] : base + e};
            if (cur.i == IDX_NONE || cur.i < 0) continue;
            if (isnan(cur.v)) {
                if (cur.i < nan_i) nan_i = cur.i;
                continue;
            }
            if (r > 0 && !vi_less(prev, cur)) continue;         // already emitted
            if (best.i == IDX_NONE || vi_less(cur, best)) best = cur;
        }
        best = vi_warp_min(best);
        if (lane == 0) s_w[warp] = best;
        if (r == 0 && first_nan) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                long long o = __shfl_xor_sync(0xffffffffu, nan_i, off);
                nan_i = o < nan_i ? o : nan_i;
            }
            if (lane == 0) s_nan[warp] = nan_i;
        }
        __syncthreads();
        if (warp == 0) {
            ValIdx x = lane < TOPK_THREADS / 32 ? s_w[lane] : ValIdx{0.0, IDX_NONE};
            x = vi_warp_min(x);
            if (lane == 0) {
                s_win = x;
                out_v[r] = x.i == IDX_NONE ? nan("") : x.v;
                out_i[r] = x.i == IDX_NONE ? -1 : x.i;
                if (r == 0 && first_nan) {
                    long long mn = INT64_MAX;
                    for (int w = 0; w < TOPK_THREADS / 32; ++w) mn = s_nan[w] < mn ? s_nan[w] : mn;
                    *first_nan = mn == INT64_MAX ? -1 : mn;
                }
            }
        }
        __syncthreads();
        prev = s_win;
        if (prev.i == IDX_NONE) {           // slice exhausted: pad the rest
            if (tid == 0)
                for (int rr = r + 1; rr < k; ++rr) { out_v[rr] = nan(""); out_i[rr] = -1; }
            break;
        }
    }
}

// grid = nblocks; block b scans vals[b*slice, min(count,(b+1)*slice))
__global__ void __launch_bounds__(TOPK_THREADS) topk_partial_kernel(const double* __restrict__ vals, int64_t count,
                                                                    int64_t slice, int64_t index_offset, int k,
                                                                    double* part_v, int64_t* part_i, int64_t* part_nan) {
    const int64_t lo = (int64_t)blockIdx.x * slice;
    int64_t n = count - lo;
    if (n > slice) n = slice;
    if (n < 0) n = 0;
    block_topk_rounds(vals + lo, nullptr, n, index_offset + lo, k, part_v + (size_t)blockIdx.x * k,
                      part_i + (size_t)blockIdx.x * k, part_nan + blockIdx.x);
}

__global__ void __launch_bounds__(TOPK_THREADS) topk_merge_kernel(const double* __restrict__ part_v,
                                                                  const int64_t* __restrict__ part_i,
                                                                  const int64_t* __restrict__ part_nan, int nparts, int k,
                                                                  double* out_v, int64_t* out_i, int64_t* out_nan) {
    block_topk_rounds(part_v, part_i, (int64_t)nparts * k, 0, k, out_v, out_i, nullptr);
    if (threadIdx.x == 0 && out_nan) {
        long long mn = INT64_MAX;
        for (int b = 0; b < nparts; ++b)
            if (part_nan[b] >= 0 && part_nan[b] < mn) mn = part_nan[b];
        *out_nan = mn == INT64_MAX ? -1 : mn;
    }
}

// Cross-rank merge of packed top-k records (one block per acquisition).  A record is `n_acq * (2 k + 1)` 8-byte words:
//   [n_acq][k] values (f64) | [n_acq][k] global indices (i64, -1 = padding) | [n_acq] lowest NaN index (i64, -1 = none)
// `gathered` holds `n_records` of them back to back (what one all-gather of the per-rank records delivers); `merged`
// receives one record of the same layout: per acquisition the k smallest (value, index) pairs over all records, ties to
// the lowest index, and the lowest NaN index of any record - np.argmin / np.argsort[:k] over the concatenated shards.
__global__ void __launch_bounds__(TOPK_THREADS) topk_merge_records_kernel(const int64_t* __restrict__ gathered, int n_records,
                                                                          int n_acq, int k, int64_t* merged) {
    const int a = blockIdx.x;
    const int64_t rec_words = (int64_t)n_acq * (2 * k + 1);
    const double* vals = reinterpret_cast<const double*>(gathered) + (int64_t)a * k;
    const int64_t* idx = gathered + (int64_t)n_acq * k + (int64_t)a * k;
    double* out_v = reinterpret_cast<double*>(merged) + (int64_t)a * k;
    int64_t* out_i = merged + (int64_t)n_acq * k + (int64_t)a * k;
    if (k > 0) block_topk_rounds(vals, idx, (int64_t)n_records * k, 0, k, out_v, out_i, nullptr, k, rec_words);
    if (threadIdx.x == 0) {
        long long mn = INT64_MAX;
        for (int r = 0; r < n_records; ++r) {
            const long long v = gathered[r * rec_words + 2 * (int64_t)n_acq * k + a];
            if (v >= 0 && v < mn) mn = v;
        }
        merged[2 * (int64_t)n_acq * k + a] = mn == INT64_MAX ? -1 : mn;
    }
}

// Padding of an empty candidate set: NaN values, index -1, no NaN seen (what the host entry points pre-fill).
__global__ void topk_fill_empty_kernel(double* out_v, int64_t* out_i, int64_t* out_nan, int k) {
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        if (out_v) out_v[j] = nan("");
        if (out_i) out_i[j] = -1;
    }
    if (threadIdx.x == 0 && out_nan) *out_nan = -1;
}

}  // namespace skb
