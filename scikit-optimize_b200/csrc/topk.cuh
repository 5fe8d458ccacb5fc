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

// The same two steps for up to three value vectors (the acquisitions of one scoring call) in ONE launch each:
// blockIdx.y selects the vector.  part_* hold [n_vec][nparts][k] / [n_vec][nparts] partial results.
struct TopkBatch {
    const double* vals[3];
    double* out_v[3];
    int64_t* out_i[3];
    int64_t* out_nan[3];
};

__global__ void __launch_bounds__(TOPK_THREADS) topk_partial_batch_kernel(const TopkBatch b, int64_t count, int64_t slice,
                                                                          int64_t index_offset, int k, double* part_v,
                                                                          int64_t* part_i, int64_t* part_nan) {
    const int a = blockIdx.y;
    const int64_t lo = (int64_t)blockIdx.x * slice;
    int64_t n = count - lo;
    if (n > slice) n = slice;
    if (n < 0) n = 0;
    const size_t slot = (size_t)a * gridDim.x + blockIdx.x;
    block_topk_rounds(b.vals[a] + lo, nullptr, n, index_offset + lo, k, part_v + slot * k, part_i + slot * k, part_nan + slot);
}

__global__ void __launch_bounds__(TOPK_THREADS) topk_merge_batch_kernel(const TopkBatch b, const double* __restrict__ part_v,
                                                                        const int64_t* __restrict__ part_i,
                                                                        const int64_t* __restrict__ part_nan, int nparts, int k) {
    const int a = blockIdx.x;
    const size_t base = (size_t)a * nparts;
    block_topk_rounds(part_v + base * k, part_i + base * k, (int64_t)nparts * k, 0, k, b.out_v[a], b.out_i[a], nullptr);
    if (threadIdx.x == 0 && b.out_nan[a]) {
        long long mn = INT64_MAX;
        for (int q = 0; q < nparts; ++q)
            if (part_nan[base + q] >= 0 && part_nan[base + q] < mn) mn = part_nan[base + q];
        *b.out_nan[a] = mn == INT64_MAX ? -1 : mn;
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

// ---------------------------------------------------------------------------------------------------------------
// Exchange + merge in ONE kernel over NVLink peer memory (the step after the per-rank top-k: SURVEY.md 8e / K4).
// Every rank owns a symmetric buffer that all peers have mapped (torch symmetric memory supplies the mapping; the
// kernel only sees pointers).  Block a (one per acquisition) of rank r
//   1. stores its slices of the local record - values, indices, first-NaN of acquisition a - into slot r of EVERY rank's
//      buffer (plain stores through the peer mapping), fences at system scope and raises flag (set, a, r) = seq there;
//   2. waits until its own buffer shows flag (set, a, p) = seq for every rank p (acquire loads at system scope, bounded:
//      a rank that never arrives sets *status instead of hanging the GPU);
//   3. copies the world's slices from its own buffer into shared memory with L2-coherent loads and runs the same
//      block_topk_rounds merge as topk_merge_records_kernel.
// Two slot sets alternate with seq: a rank can be at most one exchange ahead of a peer (it needs that peer's flag of the
// current exchange to finish), so set seq & 1 is never overwritten while somebody still reads it.
// Buffer layout in 8-byte words: [2 sets][PEER_MAX_WORLD ranks][PEER_SLOT_WORDS] records | [2][3 acq][PEER_MAX_WORLD] flags.
constexpr int PEER_MAX_WORLD = 8;
constexpr int PEER_SLOT_WORDS = 1024;                       // room for one record: n_acq (2 k + 1) <= 1024
constexpr int PEER_FLAG_BASE = 2 * PEER_MAX_WORLD * PEER_SLOT_WORDS;
constexpr int PEER_BUFFER_WORDS = PEER_FLAG_BASE + 2 * 3 * PEER_MAX_WORLD;

struct PeerExchangeParams {
    unsigned long long* peer[PEER_MAX_WORLD];   // every rank's buffer as mapped on THIS device ([world] used; [rank] is the own one)
    const int64_t* rec;                          // local packed record (device)
    int64_t* merged;                             // one record + 1 word: the merged record, then the status word
    int rank, world, n_acq, k;
    unsigned long long seq;                      // 1, 2, 3, ... identical on all ranks
    long long spin_limit;                        // clock64 ticks a rank waits for its peers
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(TOPK_THREADS) topk_exchange_peer_kernel(const PeerExchangeParams p) {
    extern __shared__ __align__(16) unsigned char smem_px[];
    int64_t* stage = reinterpret_cast<int64_t*>(smem_px);       // [world][2 k + 1]: values | indices | first NaN
    const int a = blockIdx.x, tid = threadIdx.x, k = p.k, nk = p.n_acq * p.k;
    const int set = (int)(p.seq & 1ull);
    const size_t slot_mine = ((size_t)set * PEER_MAX_WORLD + p.rank) * PEER_SLOT_WORDS;
    // 1. my slices into slot `rank` of every rank's buffer
    for (int q = 0; q < p.world; ++q) {
        unsigned long long* dst = p.peer[q] + slot_mine;
        for (int i = tid; i < k; i += TOPK_THREADS) {
            dst[a * k + i] = (unsigned long long)p.rec[a * k + i];
            dst[nk + a * k + i] = (unsigned long long)p.rec[nk + a * k + i];
        }
        if (tid == 0) dst[2 * nk + a] = (unsigned long long)p.rec[2 * nk + a];
    }
    __threadfence_system();
    __syncthreads();
    const size_t flag_off = (size_t)PEER_FLAG_BASE + ((size_t)set * 3 + a) * PEER_MAX_WORLD;
    if (tid < p.world) st_release_sys(p.peer[tid] + flag_off + p.rank, p.seq);
    // 2. the world's flags in my own buffer
    __shared__ int s_timeout;
    if (tid == 0) s_timeout = 0;
    __syncthreads();
    if (tid < p.world) {
        const unsigned long long* f = p.peer[p.rank] + flag_off + tid;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != p.seq) {
            if (clock64() - t0 > p.spin_limit) {
                s_timeout = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    // 3. stage the slices (L2-coherent loads: the lines were written by peers) and merge
    const unsigned long long* own = p.peer[p.rank] + (size_t)set * PEER_MAX_WORLD * PEER_SLOT_WORDS;
    const int row = 2 * k + 1;
    for (int e = tid; e < p.world * k; e += TOPK_THREADS) {
        const int r = e / k, i = e - r * k;
        const unsigned long long* src = own + (size_t)r * PEER_SLOT_WORDS;
        stage[r * row + i] = (int64_t)ld_relaxed_sys(src + a * k + i);
        stage[r * row + k + i] = (int64_t)ld_relaxed_sys(src + nk + a * k + i);
    }
    if (tid < p.world) stage[tid * row + 2 * k] = (int64_t)ld_relaxed_sys(own + (size_t)tid * PEER_SLOT_WORDS + 2 * nk + a);
    __syncthreads();
    double* out_v = reinterpret_cast<double*>(p.merged) + (int64_t)a * k;
    int64_t* out_i = p.merged + (int64_t)nk + (int64_t)a * k;
    if (k > 0)
        block_topk_rounds(reinterpret_cast<const double*>(stage), stage + k, (int64_t)p.world * k, 0, k, out_v, out_i, nullptr, k, row);
    if (tid == 0) {
        long long mn = INT64_MAX;
        for (int r = 0; r < p.world; ++r) {
            const long long v = stage[r * row + 2 * k];
            if (v >= 0 && v < mn) mn = v;
        }
        p.merged[2 * (int64_t)nk + a] = mn == INT64_MAX ? -1 : mn;
        if (s_timeout) atomicExch(reinterpret_cast<unsigned long long*>(p.merged + (2 * (int64_t)nk + p.n_acq)), 1ull);
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
