// K1: cross-covariance tile + posterior-mean contraction.
//
// For a tile of BN candidates: k[j, c] = amplitude * base(|| x_c/l - X_j/l ||) + bias for every training point j
// (direct differences, features summed in order, like scipy cdist reached from sklearn kernels.py:1720 / :1569),
// the running dot k[:, c] . alpha (gpr.py:316), and the k-tile written once in the k4-interleaved layout that the
// DMMA contraction (K2) consumes.  The m x n matrix K_trans of the reference is never materialised in row-major
// form and never leaves the GPU.  Training rows are staged through shared memory with 1-D bulk copies (TMA),
// double buffered on mbarriers.
#pragma once
#include "fastmath.cuh"
#include "layout.cuh"
#include "ptx.cuh"

namespace skb {

struct KCrossParams {
    const double* X;          // [m][d] candidates (un-scaled), row-major
    const double* ls;         // [d] length scales
    const double* Xp;         // packed training set: [n_pad/4][d_pad][4], already divided by ls, zero padded
    const double* alpha_p;    // [n_pad], zero padded
    double* Kt;               // [tiles][n_pad/4][BN][4]  or nullptr (mean-only)
    double* kdot;             // [tiles*BN]  K_trans . alpha
    int64_t m;
    int n, n_pad, d, d_pad;
    double amplitude, bias;
};

template <int KIND>
__device__ __forceinline__ double base_kernel(double sq) {
    if (KIND == 0) return exp_nonpos(-0.5 * sq);                  // RBF: exp(-sqeuclidean/2)
    const double r = sqrt_nonneg(sq);
    if (KIND == 1) return exp_nonpos(-r);                         // Matern 1/2
    if (KIND == 2) { const double K = r * 1.7320508075688772; return (1.0 + K) * exp_nonpos(-K); }
    const double K = r * 2.23606797749979;                        // Matern 5/2
    return (1.0 + K + K * K * (1.0 / 3.0)) * exp_nonpos(-K);
}

constexpr size_t kcross_smem_bytes(int d_pad) {
    return sizeof(double) * ((size_t)d_pad * XS_LD + 2 * (size_t)XG * d_pad * 4 + 2 * XROWS + BN) + 2 * sizeof(uint64_t);
}

template <int KIND>
__global__ void __launch_bounds__(256, 2) kcross_mean_kernel(const KCrossParams p) {
    extern __shared__ __align__(16) unsigned char smem_k1[];
    const int d_pad = p.d_pad;
    const int xt_stage = XG * d_pad * 4;
    double* xt = reinterpret_cast<double*>(smem_k1);       // [2][XG][d_pad][4]   (bulk-copy destinations first: 16B aligned)
    double* al = xt + 2 * xt_stage;                          // [2][XROWS]
    double* xs = al + 2 * XROWS;                             // [d_pad][XS_LD]
    double* mpart = xs + (size_t)d_pad * XS_LD;              // [BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(mpart + BN);

    const int tid = threadIdx.x, c = tid & (BN - 1), h = tid >> 7;
    const int64_t tile = blockIdx.x;
    const int nst = (p.n + XROWS - 1) / XROWS;      // stages holding real training rows (the rest of n_pad is padding)

    if (tid == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_fence_init();
    }
    // candidate tile, scaled by the length scale exactly like sklearn (X / length_scale, kernels.py:1714,1720)
    for (int idx = tid; idx < BN * d_pad; idx += 256) {
        const int cc = idx / d_pad, k = idx - cc * d_pad;
        const int64_t row = tile * BN + cc;
        double v = 0.0;
        if (k < p.d && row < p.m) v = p.X[row * p.d + k] / p.ls[k];
        xs[k * XS_LD + cc] = v;
    }
    __syncthreads();

    auto issue = [&](int st) {
        const int buf = st & 1;
        mbar_arrive_expect_tx(bars + buf, (uint32_t)(xt_stage + XROWS) * 8u);
        bulk_g2s(xt + buf * xt_stage, p.Xp + (size_t)st * xt_stage, (uint32_t)xt_stage * 8u, bars + buf);
        bulk_g2s(al + buf * XROWS, p.alpha_p + (size_t)st * XROWS, XROWS * 8u, bars + buf);
    };
    if (tid == 0) {
        issue(0);
        if (nst > 1) issue(1);
    }

    double macc = 0.0;
    double* kt_tile = p.Kt ? p.Kt + (size_t)tile * p.n_pad * BN : nullptr;

    for (int st = 0; st < nst; ++st) {
        const int buf = st & 1;
        mbar_wait(bars + buf, (st >> 1) & 1);
        const double* xtb = xt + buf * xt_stage;
        const double* alb = al + buf * XROWS;

        double s[XG / 2][4];
#pragma unroll
        for (int gi = 0; gi < XG / 2; ++gi) s[gi][0] = s[gi][1] = s[gi][2] = s[gi][3] = 0.0;

        for (int k0 = 0; k0 < d_pad; k0 += 4) {
            double x[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) x[kk] = xs[(k0 + kk) * XS_LD + c];
#pragma unroll
            for (int gi = 0; gi < XG / 2; ++gi) {
                const double* t = xtb + ((2 * gi + h) * d_pad + k0) * 4;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const double2 t01 = *reinterpret_cast<const double2*>(t + kk * 4);
                    const double2 t23 = *reinterpret_cast<const double2*>(t + kk * 4 + 2);
                    const double d0 = x[kk] - t01.x, d1 = x[kk] - t01.y, d2 = x[kk] - t23.x, d3 = x[kk] - t23.y;
                    s[gi][0] = fma(d0, d0, s[gi][0]);
                    s[gi][1] = fma(d1, d1, s[gi][1]);
                    s[gi][2] = fma(d2, d2, s[gi][2]);
                    s[gi][3] = fma(d3, d3, s[gi][3]);
                }
            }
        }
#pragma unroll
        for (int gi = 0; gi < XG / 2; ++gi) {
            const int g = 2 * gi + h;
            const int j0 = st * XROWS + g * 4;
            double v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                double val = fma(p.amplitude, base_kernel<KIND>(s[gi][r]), p.bias);
                if (j0 + r >= p.n) val = 0.0;          // padded training rows must contribute exact zeros downstream
                macc = fma(val, alb[g * 4 + r], macc);
                v[r] = val;
            }
            if (kt_tile) {
                double* dst = kt_tile + ((size_t)(j0 >> 2) * BN + c) * 4;
                *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
                *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
            }
        }
        __syncthreads();                                // everyone is done with buffer `buf`
        if (tid == 0 && st + 2 < nst) issue(st + 2);
    }
    if (h == 1) mpart[c] = macc;
    __syncthreads();
    if (h == 0) p.kdot[tile * BN + c] = macc + mpart[c];
}

}  // namespace skb
