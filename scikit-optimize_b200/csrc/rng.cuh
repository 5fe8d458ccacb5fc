// Candidate generation on the device, bit-identical to the host path it replaces.
//
// Optimizer._tell draws its candidate set with `space.transform(space.rvs(n_points, rng))` (optimizer.py:552-553): for
// every dimension, n_points doubles from NumPy's legacy MT19937 stream (`RandomState.uniform`, two 32-bit outputs per
// double, genrand_res53), pushed through the dimension's un-warp / clip / warp arithmetic (space.py:205-209, 302-346,
// transformers.py:243-276).  On the host this costs more than scoring the candidates on the GPU.  Here
//   * mt19937_fill_kernel continues the SAME stream from the RandomState's own 624-word state (one CTA: the state
//     recurrence is sequential across regenerations but its 624 words regenerate in four parallel phases), and hands
//     the advanced state back so the host generator stays in step;
//   * candidates_transform_kernel applies the per-dimension arithmetic with IEEE round-to-nearest operations in the
//     host's order (no FMA contraction), so every coordinate has exactly the bits NumPy produces.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace skb {

constexpr int MT_N = 624, MT_M = 397;

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
// genrand_res53: (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53, every step exact in FP64
__device__ __forceinline__ double mt_double(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

// key: the generator's 624 state words (in/out), pos_io: its position (in/out), words[n_words]: the next n_words RAW
// (untempered) outputs of the generator.  The stream is inherently serial across regenerations of the 624-word state,
// so one CTA walks it and does nothing else: inside a regeneration the words depend only on the previous state and on
// words at distance 227, which gives three parallel phases (the last word joins the third); with two state buffers no
// phase reads what it writes, so a regeneration costs three barriers.  Tempering, the 53-bit conversion and the
// per-dimension arithmetic are embarrassingly parallel and run in a second, full-grid kernel.
__global__ void __launch_bounds__(256) mt19937_words_kernel(uint32_t* __restrict__ key, int* __restrict__ pos_io,
                                                            int64_t n_words, uint32_t* __restrict__ words) {
    __shared__ uint32_t buf[2][MT_N];
    const int tid = threadIdx.x;
    for (int i = tid; i < MT_N; i += 256) buf[0][i] = key[i];
    int pos = *pos_io;
    int cur = 0;
    __syncthreads();
    int64_t produced = 0;
    while (produced < n_words) {
        if (pos == MT_N) {
            const uint32_t* o = buf[cur];
            uint32_t* n = buf[cur ^ 1];
            if (tid < 227) n[tid] = o[tid + MT_M] ^ mt_twist(o[tid], o[tid + 1]);                       // i in [0, 227)
            __syncthreads();
            if (tid < 227) n[tid + 227] = n[tid] ^ mt_twist(o[tid + 227], o[tid + 228]);                // i in [227, 454)
            __syncthreads();
            if (tid < 169) n[tid + 454] = n[tid + 227] ^ mt_twist(o[tid + 454], o[tid + 455]);          // i in [454, 623)
            else if (tid == 255) n[623] = n[396] ^ mt_twist(o[623], n[0]);                              // i = 623
            __syncthreads();
            cur ^= 1;
            pos = 0;
        }
        const int64_t left = n_words - produced;
        const int take = (int)((MT_N - pos) < left ? (MT_N - pos) : left);
        // the next regeneration only READS this buffer, so no barrier is needed before it starts
        for (int i = tid; i < take; i += 256) words[produced + i] = buf[cur][pos + i];
        produced += take;
        pos += take;
    }
    __syncthreads();
    for (int i = tid; i < MT_N; i += 256) key[i] = buf[cur][i];
    if (tid == 0) *pos_io = pos;
}

// ---- parallel walk: one CTA per dimension, each started from its own point of the ONE stream --------------------------
// The stream order of a candidate set is [dimension][candidate] (space.rvs draws dimension by dimension), so column j
// starts J_j = j * words_per_column words after the generator's current position.  MT19937 is linear over GF(2): with
// w = the window of the next 624 raw words and A the one-word shift (x_{t+624} = x_{t+397} ^ twist(x_t, x_{t+1})),
// A^J w = g(A) w for g = z^J mod phi(z) - except for the low 31 bits of the window's first word, which A ignores and
// drops: so g = z^(J-1) mod phi is applied by Horner's rule (h <- A h ^ g_i w, 19937 steps, one barrier each) and ONE
// plain step follows.  The polynomials come from the host (skopt_b200/mt_jump.py, cached per (words_per_column, d)).
//   w: [624] base window; polys: [n_columns - 1][624] (bit b of word k = coefficient of z^(32 k + b));
//   keys: [n_columns][624], keys[j] = window of column j (usable as a generator key at position 0).
__global__ void __launch_bounds__(640) mt_jump_kernel(const uint32_t* __restrict__ w, const uint32_t* __restrict__ polys,
                                                      uint32_t* __restrict__ keys) {
    __shared__ uint32_t base[MT_N], poly[MT_N], h[2][MT_N];
    const int t = threadIdx.x, col = blockIdx.x;
    if (t < MT_N) {
        base[t] = w[t];
        h[0][t] = 0u;
        if (col > 0) poly[t] = polys[(size_t)(col - 1) * MT_N + t];
    }
    __syncthreads();
    if (col == 0) {
        if (t < MT_N) keys[t] = base[t];
        return;
    }
    int cur = 0;
    for (int i = MT_N * 32 - 1; i >= 0; --i) {                // leading zero coefficients only shift the zero vector
        const uint32_t* o = h[cur];
        if (t < MT_N) {
            uint32_t v = t < MT_N - 1 ? o[t + 1] : (o[MT_M] ^ mt_twist(o[0], o[1]));
            if ((poly[i >> 5] >> (i & 31)) & 1u) v ^= base[t];
            h[cur ^ 1][t] = v;
        }
        __syncthreads();
        cur ^= 1;
    }
    if (t < MT_N) {
        const uint32_t* o = h[cur];
        keys[(size_t)col * MT_N + t] = t < MT_N - 1 ? o[t + 1] : (o[MT_M] ^ mt_twist(o[0], o[1]));
    }
}

// Column j walks words_per_column raw words from keys[j] into words[j * words_per_column ...]; the last column goes on
// for 624 more words (`extra`), from which the host rebuilds the (key, position) NumPy itself would hold after the draw.
__global__ void __launch_bounds__(256) mt19937_columns_kernel(const uint32_t* __restrict__ keys, int64_t words_per_column,
                                                              uint32_t* __restrict__ words, uint32_t* __restrict__ extra) {
    __shared__ uint32_t buf[2][MT_N];
    const int tid = threadIdx.x;
    const int col = blockIdx.x;
    const bool last = col == (int)gridDim.x - 1;
    for (int i = tid; i < MT_N; i += 256) buf[0][i] = keys[(size_t)col * MT_N + i];
    uint32_t* out = words + (size_t)col * words_per_column;
    const int64_t total = words_per_column + (last ? MT_N : 0);
    int cur = 0;
    __syncthreads();
    for (int64_t produced = 0; produced < total; produced += MT_N) {
        if (produced > 0) {
            const uint32_t* o = buf[cur];
            uint32_t* n = buf[cur ^ 1];
            if (tid < 227) n[tid] = o[tid + MT_M] ^ mt_twist(o[tid], o[tid + 1]);
            __syncthreads();
            if (tid < 227) n[tid + 227] = n[tid] ^ mt_twist(o[tid + 227], o[tid + 228]);
            __syncthreads();
            if (tid < 169) n[tid + 454] = n[tid + 227] ^ mt_twist(o[tid + 454], o[tid + 455]);
            else if (tid == 255) n[623] = n[396] ^ mt_twist(o[623], n[0]);
            __syncthreads();
            cur ^= 1;
        }
        for (int i = tid; i < MT_N; i += 256) {
            const int64_t k = produced + i;
            if (k < words_per_column) out[k] = buf[cur][i];
            else if (k < total) extra[k - words_per_column] = buf[cur][i];
        }
    }
}

// out[k] = genrand_res53(words[2k], words[2k+1])  ==  RandomState.uniform(0, 1)
__global__ void mt_words_to_uniform_kernel(const uint32_t* __restrict__ words, int64_t count, double* __restrict__ out) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < count) {
        const uint2 w = reinterpret_cast<const uint2*>(words)[k];
        out[k] = mt_double(mt_temper(w.x), mt_temper(w.y));
    }
}

// Per-dimension arithmetic of the "normalize" warping (every GP run), kinds the device reproduces bit for bit:
//   0  Real, uniform prior:      v = u*(1+eps) + 0;  x = clip(v*(hi-lo)+lo, lo, hi);                t = (x-lo)/(hi-lo)
//   1  Integer / Categorical:    v as above;         x = rint(clip(rint(v*(hi-lo)+lo), lo, hi));    t = (rint(x)-lo)/(hi-lo)
// (log-uniform priors need pow()/log10(), which differ in the last bit between libm and CUDA: those stay on the host)
struct DimDesc {
    int32_t kind;
    int32_t pad;
    double lo, hi;
};

// W: raw generator words, two per uniform, uniforms in stream order [d][m] (dimension by dimension);
// X: [m][d] row-major candidates
__global__ void __launch_bounds__(256) candidates_transform_kernel(const uint32_t* __restrict__ W, const DimDesc* __restrict__ dims,
                                                                   int64_t m, int d, double* __restrict__ X) {
    __shared__ double tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int jj = ty; jj < 32; jj += 8) {
        const int j = j0 + jj;
        const int64_t i = i0 + tx;
        double t = 0.0;
        if (j < d && i < m) {
            const DimDesc dd = dims[j];
            const double width = __dsub_rn(dd.hi, dd.lo);
            const uint2 w2 = reinterpret_cast<const uint2*>(W)[(int64_t)j * m + i];
            const double u = mt_double(mt_temper(w2.x), mt_temper(w2.y));
            const double v = __dadd_rn(__dmul_rn(u, 1.0000000000000002), 0.0);
            double x = __dadd_rn(__dmul_rn(v, width), dd.lo);
            if (dd.kind == 1) x = rint(x);
            x = fmin(fmax(x, dd.lo), dd.hi);
            if (dd.kind == 1) x = rint(x);
            t = __ddiv_rn(__dsub_rn(x, dd.lo), width);
        }
        tile[jj][tx] = t;
    }
    __syncthreads();
    for (int ii = ty; ii < 32; ii += 8) {
        const int64_t i = i0 + ii;
        const int j = j0 + tx;
        if (i < m && j < d) X[i * d + j] = tile[tx][ii];
    }
}

// rows[k] = X[idx[k]]  (the few candidates the host needs back: the arg-min / the L-BFGS starts)
__global__ void gather_rows_kernel(const double* __restrict__ X, const int64_t* __restrict__ idx, int k, int d,
                                   double* __restrict__ rows) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < k * d) {
        const int r = t / d, c = t - r * d;
        rows[t] = idx[r] >= 0 ? X[idx[r] * d + c] : nan("");
    }
}

}  // namespace skb
