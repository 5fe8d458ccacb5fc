// GP fit objective on the GPU: log marginal likelihood and its gradient with respect to the log hyper-parameters
// (SURVEY.md 8(f) rank 3).  Restates sklearn GaussianProcessRegressor.log_marginal_likelihood (_gpr.py:541-657, reached
// from skopt gpr.py:195 through sklearn's fit) for the kernel grammar  amplitude * {RBF | Matern}(length_scale) + noise:
//
//   K      = amplitude * base(|| x_i/l - x_j/l ||) + (noise + alpha) I
//   L      = chol(K),  a = K^-1 y
//   lml    = -1/2 y.a - sum_i log L_ii - n/2 log 2 pi
//   dlml/dtheta_k = 1/2 sum_ij (a_i a_j - K^-1_ij) dK_ij/dtheta_k
//
// The reference materialises dK/dtheta as an n x n x (d + 2) tensor (0.7 GB at n = 2048, d = 20) and contracts it with
// einsum; here the derivative factors are recomputed from the training rows inside the contraction kernel, so the only
// n x n objects are K (overwritten by its factor and then by K^-1) - nothing of size n^2 d exists.
// The factorisation and solves are plain library calls loaded lazily by skb.cu: cuSOLVER potrf / potrs, and for
// K^-1 = L^-T L^-1 cuBLAS trsm (on the identity) + syrk.
#pragma once
#include "fastmath.cuh"
#include "kcross.cuh"

namespace skb {

constexpr int FIT_TILE = 32;            // K is processed in 32 x 32 tiles, 256 threads, 4 entries per thread
constexpr int FIT_THREADS = 256;

struct FitParams {
    const double* Xs;       // [n][d] training rows already divided by the length scales
    int n, d;
    const double* hyp;      // device: {amplitude, noise} of this evaluation (read at run time: the launch sequence of an
                            // evaluation is replayed as a CUDA graph, so nothing that changes with theta is a launch argument)
    double alpha;           // the estimator's fixed diagonal term: the diagonal is (amplitude + noise) + alpha, in that order
    double* K;              // [n][n] (symmetric; storage order is immaterial)
};

// Xs[i][k] = X[i][k] / length_scale[k]   (sklearn kernels.py: `X / length_scale` before pdist)
__global__ void fit_scale_kernel(const double* __restrict__ X, const double* __restrict__ ls, double* __restrict__ Xs,
                                 int n, int d) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < (int64_t)n * d) Xs[e] = X[e] / ls[e % d];
}

template <int KIND>
__global__ void __launch_bounds__(FIT_THREADS) fit_kbuild_kernel(const FitParams p) {
    extern __shared__ double smem_fit[];
    const int ld = p.d + 1;
    double* xi = smem_fit;                      // [32][d + 1]
    double* xj = smem_fit + FIT_TILE * ld;      // [32][d + 1]
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;                        // lower block triangle; the mirror image is written from here
    for (int e = threadIdx.x; e < FIT_TILE * p.d; e += FIT_THREADS) {
        const int r = e / p.d, k = e % p.d;
        const int gi = bi * FIT_TILE + r, gj = bj * FIT_TILE + r;
        xi[r * ld + k] = gi < p.n ? p.Xs[(size_t)gi * p.d + k] : 0.0;
        xj[r * ld + k] = gj < p.n ? p.Xs[(size_t)gj * p.d + k] : 0.0;
    }
    __syncthreads();
    const int tj = threadIdx.x & 31, ti0 = threadIdx.x >> 5;
    const double amplitude = p.hyp[0], noise = p.hyp[1];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ti = ti0 + 8 * q;
        const int i = bi * FIT_TILE + ti, j = bj * FIT_TILE + tj;
        if (i >= p.n || j >= p.n) continue;
        double sq = 0.0;
        for (int k = 0; k < p.d; ++k) {
            const double df = xi[ti * ld + k] - xj[tj * ld + k];
            sq = fma(df, df, sq);
        }
        double v = amplitude * base_kernel<KIND>(sq);
        if (i == j) v = (amplitude + noise) + p.alpha;
        p.K[(size_t)i * p.n + j] = v;
        p.K[(size_t)j * p.n + i] = v;
    }
}

// lml = -1/2 y.a - sum log diag(L) - n/2 log(2 pi); one block, fixed reduction order
__global__ void fit_lml_kernel(const double* __restrict__ L, const double* __restrict__ y, const double* __restrict__ a,
                               int n, double* __restrict__ out_lml) {
    __shared__ double red[2][256];
    double s_log = 0.0, s_dot = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        s_log += log(L[(size_t)i * n + i]);
        s_dot = fma(y[i], a[i], s_dot);
    }
    red[0][threadIdx.x] = s_log;
    red[1][threadIdx.x] = s_dot;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_lml = (-0.5 * red[1][0] - red[0][0]) - 0.5 * n * 1.8378770664093453;   // log(2 pi)
}

struct FitGradParams {
    const double* Xs;       // [n][d]
    const double* Kinv;     // [n][n]: entry (i, j) with i >= j read at Kinv[j * n + i] (cuSOLVER lower, column major)
    const double* a;        // [n] K^-1 y
    int n, d;
    const double* hyp;      // device: {amplitude, noise} (see FitParams)
    double* partial;        // [n_blocks][d + 2] per-block sums: amplitude, length scales, noise
};

// dK_ij / dlog(l_k) = amplitude * gfac(r) * (x_ik/l_k - x_jk/l_k)^2        (sklearn kernels.py Matern / RBF eval_gradient)
template <int KIND>
__device__ __forceinline__ void base_and_gfac(double sq, double& kv, double& gf) {
    if (KIND == 0) { kv = exp_nonpos(-0.5 * sq); gf = kv; return; }                     // RBF: K * D
    const double r = sqrt_nonneg(sq);
    if (KIND == 1) { kv = exp_nonpos(-r); gf = r > 0.0 ? kv / r : 0.0; return; }         // nu = 1/2: K * D / r (0 at r = 0)
    if (KIND == 2) {                                                                     // nu = 3/2: 3 D exp(-sqrt(3) r)
        const double t = r * 1.7320508075688772, e = exp_nonpos(-t);
        kv = (1.0 + t) * e;
        gf = 3.0 * e;
        return;
    }
    const double t = r * 2.23606797749979, e = exp_nonpos(-t);                           // nu = 5/2: 5/3 D (t + 1) exp(-t)
    kv = (1.0 + t + t * t * (1.0 / 3.0)) * e;
    gf = (5.0 / 3.0) * (t + 1.0) * e;
}

__device__ __forceinline__ double fit_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One block per lower-triangular 32 x 32 tile of  W = a a^T - K^-1 ; off-diagonal entries count twice (symmetry).
template <int KIND>
__global__ void __launch_bounds__(FIT_THREADS) fit_grad_kernel(const FitGradParams p) {
    extern __shared__ double smem_fit[];
    const int ld = p.d + 1;
    double* xi = smem_fit;                          // [32][d + 1]
    double* xj = xi + FIT_TILE * ld;                // [32][d + 1]
    double* wsum = xj + FIT_TILE * ld;              // [8 warps][d + 2]
    // linear block index -> (bi, bj) with bj <= bi
    const int b = blockIdx.x;
    int bi = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
    while (bi * (bi + 1) / 2 > b) --bi;
    const int bj = b - bi * (bi + 1) / 2;
    for (int e = threadIdx.x; e < FIT_TILE * p.d; e += FIT_THREADS) {
        const int r = e / p.d, k = e % p.d;
        const int gi = bi * FIT_TILE + r, gj = bj * FIT_TILE + r;
        xi[r * ld + k] = gi < p.n ? p.Xs[(size_t)gi * p.d + k] : 0.0;
        xj[r * ld + k] = gj < p.n ? p.Xs[(size_t)gj * p.d + k] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ti = lane;                            // lanes run down a column of the column-major K^-1: coalesced
    double w[4], wg[4];
    double s_amp = 0.0, s_noise = 0.0;
    const double amplitude = p.hyp[0], noise = p.hyp[1];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int tj = warp + 8 * q;
        const int i = bi * FIT_TILE + ti, j = bj * FIT_TILE + tj;
        w[q] = 0.0;
        wg[q] = 0.0;
        if (i >= p.n || j >= p.n || j > i) continue;
        const double inner = p.a[i] * p.a[j] - p.Kinv[(size_t)j * p.n + i];
        double sq = 0.0;
        for (int k = 0; k < p.d; ++k) {
            const double df = xi[ti * ld + k] - xj[tj * ld + k];
            sq = fma(df, df, sq);
        }
        double kv, gf;
        base_and_gfac<KIND>(sq, kv, gf);
        const double mult = (i == j) ? 1.0 : 2.0;
        s_amp = fma(mult * inner, amplitude * kv, s_amp);
        if (i == j) s_noise = fma(inner, noise, s_noise);
        w[q] = 1.0;
        wg[q] = mult * inner * (amplitude * gf);
    }
    s_amp = fit_warp_sum(s_amp);
    s_noise = fit_warp_sum(s_noise);
    if (lane == 0) {
        wsum[warp * (p.d + 2)] = s_amp;
        wsum[warp * (p.d + 2) + p.d + 1] = s_noise;
    }
    for (int k = 0; k < p.d; ++k) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (w[q] == 0.0) continue;
            const double df = xi[ti * ld + k] - xj[(warp + 8 * q) * ld + k];
            s = fma(wg[q], df * df, s);
        }
        s = fit_warp_sum(s);
        if (lane == 0) wsum[warp * (p.d + 2) + 1 + k] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < p.d + 2; k += FIT_THREADS) {
        double s = 0.0;
        for (int wv = 0; wv < 8; ++wv) s += wsum[wv * (p.d + 2) + k];
        p.partial[(size_t)b * (p.d + 2) + k] = s;
    }
}

// grad[k] = 1/2 sum over blocks (fixed order: one thread per hyper-parameter walks the blocks in sequence, 4-way split)
__global__ void fit_grad_reduce_kernel(const double* __restrict__ partial, int n_blocks, int n_hyper,
                                       double* __restrict__ grad) {
    __shared__ double red[256];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += 256) s += partial[(size_t)b * n_hyper + k];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) grad[k] = 0.5 * red[0];
}

__global__ void fit_identity_kernel(double* __restrict__ M, int n) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < (size_t)n * n) M[e] = (e / n == e % n) ? 1.0 : 0.0;
}

// cuSOLVER works on column-major lower triangles; the rest of this library reads row-major matrices.  `src` holds the
// lower triangle of an n x n matrix in column-major order (entry (i, j), i >= j, at src[j * n + i]; the other triangle
// is stale).  SYMMETRIC = false: dst = that triangular matrix, row-major, zeros above the diagonal (L_, L^-1).
// SYMMETRIC = true: dst = the full symmetric matrix (K^-1 after potri).  32 x 32 tiles through shared memory so that
// both the reads and the writes are coalesced.
template <bool SYMMETRIC>
__global__ void __launch_bounds__(256) fit_unpack_lower_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;             // destination tile: rows bi*32.., columns bj*32..
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bool mirror = bj > bi;                            // tile lies above the diagonal
    if (mirror && !SYMMETRIC) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = bi * 32 + ty + 8 * q, j = bj * 32 + tx;
            if (i < n && j < n) dst[(size_t)i * n + j] = 0.0;
        }
        return;
    }
    if (!mirror) {
        // dst(i, j) = src[j * n + i]: read with i fastest, transpose through the tile
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = bj * 32 + ty + 8 * q, i = bi * 32 + tx;
            tile[ty + 8 * q][tx] = (i < n && j < n) ? src[(size_t)j * n + i] : 0.0;        // tile[j_local][i_local]
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int il = ty + 8 * q, jl = tx;
            const int i = bi * 32 + il, j = bj * 32 + jl;
            if (i >= n || j >= n) continue;
            double v = tile[jl][il];
            if (j > i) v = SYMMETRIC ? tile[il][jl] : 0.0;  // only on diagonal tiles: (j, i) is in the same tile
            dst[(size_t)i * n + j] = v;
        }
        return;
    }
    // SYMMETRIC and above the diagonal: dst(i, j) = M(j, i) = src[i * n + j] - a straight copy of the stored triangle
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = bi * 32 + ty + 8 * q, j = bj * 32 + tx;
        if (i < n && j < n) dst[(size_t)i * n + j] = src[(size_t)i * n + j];
    }
}

}  // namespace skb
