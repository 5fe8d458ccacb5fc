// K2: blocked FP64 triangular contraction  U = L_inv . K_trans^T  on DMMA tensor-core tiles, fused with the
// column square-norm  q_c = || U[:, c] ||^2  (= k_c^T K_inv k_c, the einsum of gpr.py:332) and with the posterior /
// acquisition epilogue (gpr.py:316-343, acquisition.py:129-146,194-229,276-303).
//
// One CTA owns a tile of BN candidates and walks the row blocks of L_inv top to bottom; row block I only needs the
// training points j < (I+1)*BM, so only the lower-triangular tiles of L_inv are ever loaded or multiplied
// (n^2/2 MACs per candidate, plus the zero half of the diagonal blocks that is skipped at fragment granularity).
// Operand tiles arrive through a 3-stage mbarrier pipeline fed by a dedicated producer warp issuing 1-D bulk
// copies (TMA); the 8 consumer warps hold a 128 x 128 FP64 accumulator tile in registers (64 doubles/thread).
// U is never written anywhere: it is squared and folded into 8 running sums per thread as soon as a row block
// completes.  All reductions run in a fixed order, so a candidate's result does not depend on which tile, lane
// or GPU processed it.
#pragma once
#include <math.h>
#include "layout.cuh"
#include "ptx.cuh"

namespace skb {

struct AcqConsts {
    double prior_var, y_mean, y_std;
    double y_opt, xi, kappa;
    int kappa_inf, acq_mask;
};

struct TrmmParams {
    const double* Vp;         // packed lower-triangular tiles of L_inv (see pack_linv_kernel)
    const double* Kt;         // [tiles][n_pad/4][BN][4] from K1
    const double* kdot;       // [tiles*BN] K_trans . alpha from K1
    int n_pad;
    int n;                    // true number of training points: stages / fragments past it are skipped
    int64_t m;                // candidates in this launch (tiles*BN >= m)
    AcqConsts a;
    double* mean;             // [m] or nullptr
    double* std;              // [m] or nullptr
    double* acq[3];           // [m] or nullptr
    double* qout;             // [m] or nullptr: raw quadratic form (diagnostics / gradient path)
    unsigned long long* clamp_count;   // or nullptr: number of negative variances set to zero
    // MODE 1 (dense): A = packed K_inv (all tiles), result rows are stored instead of squared
    const double* Ap;         // packed dense K_inv: tile (I, kb) at (I * n_pad/BK + kb) * BM*BK
    double* Wt;               // [tiles][n_pad][BN]: W = K_inv . K_trans^T
};

constexpr size_t trmm_smem_bytes() {
    return sizeof(double) * ((size_t)TRMM_STAGES * (BM * BK + BK * BN) + 2 * BN) + 2 * TRMM_STAGES * sizeof(uint64_t);
}

// EI / PI / LCB exactly in the operation order of skopt/acquisition.py, returned as values to MINIMISE.
__device__ __forceinline__ void acquisition_values(const AcqConsts& a, double mu, double sd, double out[3]) {
    if (a.acq_mask & 4) out[2] = a.kappa_inf ? -sd : mu - a.kappa * sd;
    if (a.acq_mask & 3) {
        double ei = 0.0, pi = 0.0;
        if (sd > 0.0) {
            const double improve = a.y_opt - a.xi - mu;
            const double z = improve / sd;
            const double cdf = normcdf(z);
            pi = cdf;
            if (a.acq_mask & 1) {
                const double pdf = exp(-(z * z) / 2.0) * 0.3989422804014327;
                ei = improve * cdf + sd * pdf;
            }
        }
        out[0] = -ei;
        out[1] = -pi;
    }
}

__device__ __forceinline__ bool posterior_from_q(const AcqConsts& a, double kdot, double q, double& mu, double& sd) {
    mu = a.y_std * kdot + a.y_mean;               // gpr.py:316-318
    double var = a.prior_var - q;                  // gpr.py:331-332
    const bool clamped = var < 0.0;                // gpr.py:336-340
    if (clamped) var = 0.0;
    var = var * (a.y_std * a.y_std);               // gpr.py:342
    sd = sqrt(var);
    return clamped;
}

// MODE 0: triangular L_inv contraction + square-norm + acquisition epilogue, one CTA per candidate tile (grid.x).
// MODE 1: dense K_inv contraction, W rows stored for the gradient kernel; one CTA per (row block = grid.x, tile = grid.y).
template <int MODE>
__global__ void __launch_bounds__(TRMM_THREADS, 1) trmm_kernel(const TrmmParams p) {
    extern __shared__ __align__(128) unsigned char smem_k2[];
    double* sV = reinterpret_cast<double*>(smem_k2);              // [STAGES][BK/4][BM][4]
    double* sK = sV + TRMM_STAGES * BM * BK;                       // [STAGES][BK/4][BN][4]
    double* sRed = sK + TRMM_STAGES * BK * BN;                     // [2][BN]
    uint64_t* full = reinterpret_cast<uint64_t*>(sRed + 2 * BN);
    uint64_t* empty = full + TRMM_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TRMM_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 8);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int tile_id = MODE == 0 ? blockIdx.x : blockIdx.y;
    const int I_begin = MODE == 0 ? 0 : blockIdx.x;
    const int I_end = MODE == 0 ? p.n_pad / BM : blockIdx.x + 1;
    const int nkb_used = (p.n + BK - 1) / BK;          // contraction stages that contain real training points
    const int nkb_dense = p.n_pad / BK;
    const double* kt_tile = p.Kt + (size_t)tile_id * p.n_pad * BN;

    if (warp == 8) {
        // ---------------- producer: one lane streams operand tiles with bulk copies ----------------
        // (a dedicated warp keeps the DMMA warps free of any wait on the "empty" barriers; measured ~1-3 % faster
        //  than letting the DMMA warps take turns issuing the copies, even though 9 warps cap the kernel at 168
        //  registers per thread: each SM sub-partition owns 16K registers and one of them hosts 3 warps)
        if (lane == 0) {
            uint32_t it = 0;
            for (int I = I_begin; I < I_end; ++I) {
                const int nkb = min(MODE == 0 ? (I + 1) * (BM / BK) : nkb_dense, nkb_used);
                const double* vrow = MODE == 0 ? p.Vp + (size_t)vp_tiles_before(I) * (BM * BK)
                                               : p.Ap + (size_t)I * nkb_dense * (BM * BK);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % TRMM_STAGES;
                    const uint32_t ph = (it / TRMM_STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_arrive_expect_tx(full + s, (BM * BK + BK * BN) * 8u);
                    bulk_g2s(sV + s * BM * BK, vrow + (size_t)kb * BM * BK, BM * BK * 8u, full + s);
                    bulk_g2s(sK + s * BK * BN, kt_tile + (size_t)kb * BK * BN, BK * BN * 8u, full + s);
                }
            }
        }
        return;
    }

    // ---------------- consumers: 2 (rows) x 4 (candidates) warps, warp tile 64 x 32 ----------------
    const int wm = warp >> 2, wn = warp & 3;
    double c[8][4][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) c[mi][ni][0] = c[mi][ni][1] = 0.0;
    double qacc[4][2];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) qacc[ni][0] = qacc[ni][1] = 0.0;

    uint32_t it = 0;
    for (int I = I_begin; I < I_end; ++I) {
        const int nkb = min(MODE == 0 ? (I + 1) * (BM / BK) : nkb_dense, nkb_used);
        const int row0 = I * BM + wm * 64;                 // first global row of this warp's accumulator tile
        const bool rows_full = row0 + 64 <= p.n;           // false only in the last, partly padded row block
        for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % TRMM_STAGES;
            const uint32_t ph = (it / TRMM_STAGES) & 1;
            mbar_wait(full + s, ph);
            const double* v = sV + s * BM * BK + wm * 64 * 4 + lane;
            const double* k = sK + s * BK * BN + wn * 32 * 4 + lane;
            const int col0 = kb * BK;
            if (rows_full && (MODE == 1 || col0 + BK - 1 <= row0)) {
                // every (row, col) of this stage is at or below the diagonal
#pragma unroll
                for (int g = 0; g < BK / 4; ++g) {
                    double a[8], b[4];
#pragma unroll
                    for (int mi = 0; mi < 8; ++mi) a[mi] = v[g * BM * 4 + mi * 32];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) b[ni] = k[g * BN * 4 + ni * 32];
#pragma unroll
                    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) dmma884(c[mi][ni][0], c[mi][ni][1], a[mi], b[ni]);
                }
            } else if (row0 < p.n && (MODE == 1 || col0 <= row0 + 63)) {
                // stage straddles the diagonal (skip the 8x4 fragments of L_inv strictly above it) and / or the
                // row block runs past the last training point (skip the all-zero padded fragments)
#pragma unroll
                for (int g = 0; g < BK / 4; ++g) {
                    const int cg = col0 + g * 4;
                    if (MODE == 0 && cg > row0 + 63) break;
                    double b[4];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) b[ni] = k[g * BN * 4 + ni * 32];
#pragma unroll
                    for (int mi = 0; mi < 8; ++mi) {
                        if ((MODE == 1 || cg <= row0 + mi * 8 + 7) && row0 + mi * 8 < p.n) {
                            const double a = v[g * BM * 4 + mi * 32];
#pragma unroll
                            for (int ni = 0; ni < 4; ++ni) dmma884(c[mi][ni][0], c[mi][ni][1], a, b[ni]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        if (MODE == 1) {
            // store this warp's 64 x 32 block of W: lane holds rows (l>>2), column pair 2*(l&3)
            double* w = p.Wt + ((size_t)tile_id * p.n_pad + row0 + (lane >> 2)) * BN + wn * 32 + 2 * (lane & 3);
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    *reinterpret_cast<double2*>(w + (size_t)mi * 8 * BN + ni * 8) = make_double2(c[mi][ni][0], c[mi][ni][1]);
            continue;
        }
        // row block done: fold U^2 into the running column sums and clear the accumulators
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                qacc[ni][0] = fma(c[mi][ni][0], c[mi][ni][0], qacc[ni][0]);
                qacc[ni][1] = fma(c[mi][ni][1], c[mi][ni][1], qacc[ni][1]);
                c[mi][ni][0] = c[mi][ni][1] = 0.0;
            }
    }

    if (MODE == 1) return;

    // the 8 lanes that share (lane & 3) hold the 8 row groups of the same columns
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double x = qacc[ni][e];
            x += __shfl_xor_sync(0xffffffffu, x, 4);
            x += __shfl_xor_sync(0xffffffffu, x, 8);
            x += __shfl_xor_sync(0xffffffffu, x, 16);
            if (lane < 4) sRed[wm * BN + wn * 32 + ni * 8 + 2 * lane + e] = x;
        }
    named_bar_sync(1, 256);

    if (threadIdx.x < BN) {
        const int64_t cand = (int64_t)blockIdx.x * BN + threadIdx.x;
        if (cand < p.m) {
            const double q = sRed[threadIdx.x] + sRed[BN + threadIdx.x];
            double mu, sd, acq[3];
            if (posterior_from_q(p.a, p.kdot[cand], q, mu, sd) && p.clamp_count) atomicAdd(p.clamp_count, 1ull);
            acquisition_values(p.a, mu, sd, acq);
            if (p.qout) p.qout[cand] = q;
            if (p.mean) p.mean[cand] = mu;
            if (p.std) p.std[cand] = sd;
#pragma unroll
            for (int a = 0; a < 3; ++a)
                if (((p.a.acq_mask >> a) & 1) && p.acq[a]) p.acq[a][cand] = acq[a];
        }
    }
}


// Joint posterior covariance of a handful of points (gpr.py:320-325, `predict(return_cov=True)`): after the dense product
// W = K_inv . K_trans^T (trmm_kernel<1>) the quadratic form Q = K_trans . W is one more tile product,
//   Q[a][b] = sum_j Kt(tile ta, j, ca) * Wt(tile tb, j, cb),
// grid = (tb, ta), thread (ty, tx) owns rows ty + 16 i, columns tx + 16 j of the 128 x 128 block; 16 training points per round.
constexpr int COV_JB = 16;
__global__ void __launch_bounds__(256) cov_quad_kernel(const double* __restrict__ Kt, const double* __restrict__ Wt,
                                                        double* __restrict__ Q, int n, int n_pad, int64_t m) {
    __shared__ double sK[COV_JB][BN + 1];
    __shared__ double sW[COV_JB][BN];
    const int tb = blockIdx.x, ta = blockIdx.y, tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const double* kt = Kt + (size_t)ta * n_pad * BN;          // [n_pad / 4][BN][4]
    const double* wt = Wt + (size_t)tb * n_pad * BN;          // [n_pad][BN]
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    for (int j0 = 0; j0 < n; j0 += COV_JB) {
        __syncthreads();
        for (int e = tid; e < COV_JB * BN; e += 256) {
            const int g4 = e / (BN * 4), c = (e % (BN * 4)) >> 2, q = e & 3;
            sK[g4 * 4 + q][c] = kt[(size_t)(j0 / 4) * BN * 4 + e];
            sW[e / BN][e % BN] = wt[(size_t)j0 * BN + e];
        }
        __syncthreads();
#pragma unroll 4
        for (int jj = 0; jj < COV_JB; ++jj) {
            double a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sK[jj][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = sW[jj][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t ra = (int64_t)ta * BN + ty + 16 * i, cb = (int64_t)tb * BN + tx + 16 * j;
            if (ra < m && cb < m) Q[ra * m + cb] = acc[i][j];
        }
}

}  // namespace skb
