// K3g: batched input gradients of the posterior and of the acquisition functions.
//
// Per candidate x (one thread each) and training point j, with W = K_inv . K_trans^T already computed by the dense
// DMMA contraction (trmm_kernel<1>):
//   k_j      = amplitude * base(r_j) + bias                                  (recomputed: cheaper than a scratch read)
//   G_jk     = d k_j / d x_k = gcoef(r_j) * (x_k - X_jk) / l_k^2              (skopt kernels.py:68-90, 93-200, 294-308)
//   mean     = y_std * sum_j k_j alpha_j + y_mean                            (gpr.py:316-318)
//   q        = sum_j k_j W_j                 -> var, std                     (gpr.py:331-343)
//   d mean_k = y_std * sum_j alpha_j G_jk                                    (gpr.py:346-349)
//   d std_k  = -y_std^2 * sum_j W_j G_jk / std, 0 when std ~ 0               (gpr.py:350-356)
// followed by the acquisition values and the chain rule of acquisition.py:129-146, 211-227, 295-319.  The reference
// evaluates one point per call; here every row gets its gradient in one launch, each row's arithmetic independent
// of its neighbours (bit-equal rows give bit-equal results).
#pragma once
#include "kcross.cuh"
#include "trmm.cuh"

namespace skb {

constexpr int GROWS = 32;            // training rows per stage
constexpr int GG = GROWS / 4;
constexpr int GRAD_THREADS = BN;     // one thread per candidate

struct GradParams {
    const double* X;          // [m][d]
    const double* ls;         // [d]
    const double* Xp;         // packed training set (see pack_xtrain_kernel)
    const double* alpha_p;    // [n_pad]
    const double* Wt;         // [tiles][n_pad][BN]
    int64_t m;
    int n, n_pad, d, d_pad;
    double amplitude, bias;
    AcqConsts a;
    double* mean;
    double* std;
    double* mean_grad;        // [m][d]
    double* std_grad;         // [m][d]
    double* acq[3];
    double* acq_grad[3];      // [m][d]
};

constexpr size_t grad_smem_bytes(int d_pad) {
    return sizeof(double) * (2 * (size_t)GG * d_pad * 4 + 2 * GROWS + (size_t)d_pad * XS_LD + 2 * (size_t)d_pad * BN) +
           2 * sizeof(uint64_t);
}

// value and radial derivative coefficient: k = amplitude*base + bias, G_jk = gcoef * (x~_k - X~_jk) / l_k
template <int KIND>
__device__ __forceinline__ void kernel_and_gcoef(double sq, double amplitude, double bias, double& kval, double& gcoef,
                                                 bool& at_zero) {
    at_zero = false;
    if (KIND == 0) {
        const double e = exp(-0.5 * sq);
        kval = fma(amplitude, e, bias);
        gcoef = -amplitude * e;
        return;
    }
    const double r = sqrt(sq);
    if (KIND == 1) {
        const double e = exp(-r);
        kval = fma(amplitude, e, bias);
        at_zero = (r == 0.0);                       // reference: gradient row = -1/length_scale (kernels.py:118-128)
        gcoef = at_zero ? 0.0 : -amplitude * e / r;
    } else if (KIND == 2) {
        const double K = r * 1.7320508075688772;
        const double e = exp(-K);
        kval = fma(amplitude, (1.0 + K) * e, bias);
        gcoef = -3.0 * amplitude * e;
    } else {
        const double K = r * 2.23606797749979;
        const double e = exp(-K);
        kval = fma(amplitude, (1.0 + K + K * K * (1.0 / 3.0)) * e, bias);
        gcoef = -(5.0 / 3.0) * amplitude * (1.0 + K) * e;
    }
}

template <int KIND>
__global__ void __launch_bounds__(GRAD_THREADS) grad_contract_kernel(const GradParams p) {
    extern __shared__ __align__(16) unsigned char smem_k3[];
    const int d_pad = p.d_pad;
    const int xt_stage = GG * d_pad * 4;
    double* xt = reinterpret_cast<double*>(smem_k3);        // [2][GG][d_pad][4]
    double* al = xt + 2 * xt_stage;                          // [2][GROWS]
    double* xs = al + 2 * GROWS;                             // [d_pad][XS_LD]
    double* gm = xs + (size_t)d_pad * XS_LD;                 // [d_pad][BN]  sum_j alpha_j gcoef_j (x~_k - X~_jk)
    double* gs = gm + (size_t)d_pad * BN;                    // [d_pad][BN]  sum_j W_j     gcoef_j (x~_k - X~_jk)
    uint64_t* bars = reinterpret_cast<uint64_t*>(gs + (size_t)d_pad * BN);

    const int c = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t cand = tile * BN + c;
    const int nst = p.n_pad / GROWS;

    if (c == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_fence_init();
    }
    for (int idx = c; idx < BN * d_pad; idx += GRAD_THREADS) {
        const int cc = idx / d_pad, k = idx - cc * d_pad;
        const int64_t row = tile * BN + cc;
        double v = 0.0;
        if (k < p.d && row < p.m) v = p.X[row * p.d + k] / p.ls[k];
        xs[k * XS_LD + cc] = v;
    }
    for (int k = 0; k < d_pad; ++k) gm[k * BN + c] = gs[k * BN + c] = 0.0;
    __syncthreads();

    auto issue = [&](int st) {
        const int buf = st & 1;
        mbar_arrive_expect_tx(bars + buf, (uint32_t)(xt_stage + GROWS) * 8u);
        bulk_g2s(xt + buf * xt_stage, p.Xp + (size_t)st * xt_stage, (uint32_t)xt_stage * 8u, bars + buf);
        bulk_g2s(al + buf * GROWS, p.alpha_p + (size_t)st * GROWS, GROWS * 8u, bars + buf);
    };
    if (c == 0) {
        issue(0);
        if (nst > 1) issue(1);
    }

    const double* wt = p.Wt + (size_t)tile * p.n_pad * BN + c;
    double kdot = 0.0, q = 0.0, z_alpha = 0.0, z_w = 0.0;      // z_*: Matern-1/2 rows with r == 0

    for (int st = 0; st < nst; ++st) {
        const int buf = st & 1;
        double w[GROWS];
#pragma unroll
        for (int r = 0; r < GROWS; ++r) w[r] = wt[(size_t)(st * GROWS + r) * BN];     // coalesced across candidates
        mbar_wait(bars + buf, (st >> 1) & 1);
        const double* xtb = xt + buf * xt_stage;
        const double* alb = al + buf * GROWS;

        double s[GROWS];
#pragma unroll
        for (int r = 0; r < GROWS; ++r) s[r] = 0.0;
        for (int k0 = 0; k0 < d_pad; k0 += 4) {
            double x[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) x[kk] = xs[(k0 + kk) * XS_LD + c];
#pragma unroll
            for (int g = 0; g < GG; ++g) {
                const double* t = xtb + (g * d_pad + k0) * 4;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const double2 t01 = *reinterpret_cast<const double2*>(t + kk * 4);
                    const double2 t23 = *reinterpret_cast<const double2*>(t + kk * 4 + 2);
                    const double d0 = x[kk] - t01.x, d1 = x[kk] - t01.y, d2 = x[kk] - t23.x, d3 = x[kk] - t23.y;
                    s[g * 4 + 0] = fma(d0, d0, s[g * 4 + 0]);
                    s[g * 4 + 1] = fma(d1, d1, s[g * 4 + 1]);
                    s[g * 4 + 2] = fma(d2, d2, s[g * 4 + 2]);
                    s[g * 4 + 3] = fma(d3, d3, s[g * 4 + 3]);
                }
            }
        }
        // s[r] <- alpha_r * gcoef_r ;  w[r] <- W_r * gcoef_r
#pragma unroll
        for (int r = 0; r < GROWS; ++r) {
            double kval, gcoef;
            bool at_zero;
            kernel_and_gcoef<KIND>(s[r], p.amplitude, p.bias, kval, gcoef, at_zero);
            const double a_r = alb[r];
            if (st * GROWS + r >= p.n) { kval = 0.0; gcoef = 0.0; at_zero = false; }
            kdot = fma(kval, a_r, kdot);
            q = fma(kval, w[r], q);
            if (KIND == 1 && at_zero) { z_alpha += a_r; z_w += w[r]; }
            s[r] = a_r * gcoef;
            w[r] = w[r] * gcoef;
        }
        for (int k = 0; k < p.d; ++k) {
            const double x = xs[k * XS_LD + c];
            double a1 = 0.0, a2 = 0.0;
#pragma unroll
            for (int g = 0; g < GG; ++g) {
                const double* t = xtb + (g * d_pad + k) * 4;
                const double2 t01 = *reinterpret_cast<const double2*>(t);
                const double2 t23 = *reinterpret_cast<const double2*>(t + 2);
                const double d0 = x - t01.x, d1 = x - t01.y, d2 = x - t23.x, d3 = x - t23.y;
                a1 = fma(s[g * 4 + 0], d0, a1); a2 = fma(w[g * 4 + 0], d0, a2);
                a1 = fma(s[g * 4 + 1], d1, a1); a2 = fma(w[g * 4 + 1], d1, a2);
                a1 = fma(s[g * 4 + 2], d2, a1); a2 = fma(w[g * 4 + 2], d2, a2);
                a1 = fma(s[g * 4 + 3], d3, a1); a2 = fma(w[g * 4 + 3], d3, a2);
            }
            gm[k * BN + c] += a1;
            gs[k * BN + c] += a2;
        }
        __syncthreads();
        if (c == 0 && st + 2 < nst) issue(st + 2);
    }

    if (cand >= p.m) return;
    const AcqConsts& A = p.a;
    const double mu = A.y_std * kdot + A.y_mean;
    double var = A.prior_var - q;
    if (var < 0.0) var = 0.0;
    var = var * (A.y_std * A.y_std);
    const double sd = sqrt(var);
    double acq[3];
    acquisition_values(A, mu, sd, acq);
    if (p.mean) p.mean[cand] = mu;
    if (p.std) p.std[cand] = sd;
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (((A.acq_mask >> a) & 1) && p.acq[a]) p.acq[a][cand] = acq[a];

    const bool sd_is_zero = !(fabs(sd) > 1e-8);         // np.allclose(y_std, 0): |std| <= atol (gpr.py:351)
    const bool mask = sd > 0.0;                          // acquisition.py:213,296
    double improve = 0.0, cdf = 0.0, pdf = 0.0;
    if (mask && (A.acq_mask & 3)) {
        improve = A.y_opt - A.xi - mu;
        const double z = improve / sd;
        cdf = normcdf(z);
        pdf = exp(-(z * z) / 2.0) * 0.3989422804014327;
    }
    for (int k = 0; k < p.d; ++k) {
        const double inv_l = 1.0 / p.ls[k];
        double tm = gm[k * BN + c], ts = gs[k * BN + c];
        if (KIND == 1) { tm -= p.amplitude * z_alpha; ts -= p.amplitude * z_w; }
        const double gmu = A.y_std * (tm * inv_l);
        const double gsd = sd_is_zero ? 0.0 : -(ts * inv_l) / sd * (A.y_std * A.y_std);
        const int64_t o = cand * p.d + k;
        if (p.mean_grad) p.mean_grad[o] = gmu;
        if (p.std_grad) p.std_grad[o] = gsd;
        if ((A.acq_mask & 4) && p.acq_grad[2]) p.acq_grad[2][o] = A.kappa_inf ? -gsd : gmu - A.kappa * gsd;
        if (A.acq_mask & 3) {
            double g_ei = 0.0, g_pi = 0.0;
            if (mask) {
                const double improve_grad = (-gmu * sd - gsd * improve) / (sd * sd);
                const double cdf_grad = improve_grad * pdf;
                g_pi = cdf_grad;
                const double pdf_grad = -improve * cdf_grad;
                g_ei = (-gmu * cdf - pdf_grad) + (gsd * pdf + pdf_grad);
            }
            if ((A.acq_mask & 1) && p.acq_grad[0]) p.acq_grad[0][o] = -g_ei;
            if ((A.acq_mask & 2) && p.acq_grad[1]) p.acq_grad[1][o] = -g_pi;
        }
    }
}

// Stateless acquisition epilogue for posteriors that were produced elsewhere (any regressor with predict(return_std),
// the duck-typed boundary of acquisition.py:134-143,198-202,280-285): values and, when gradients of mu / std are
// given, the chain rule of acquisition.py:218-227, 305-319.  One thread per row.
struct AcqOnlyParams {
    const double* mu;
    const double* sd;
    const double* mu_grad;    // [m][d] or nullptr
    const double* sd_grad;    // [m][d] or nullptr
    int64_t m;
    int d;
    AcqConsts a;
    double* acq[3];
    double* acq_grad[3];
};

__global__ void acq_from_posterior_kernel(const AcqOnlyParams p) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= p.m) return;
    const AcqConsts& A = p.a;
    const double mu = p.mu[i], sd = p.sd[i];
    double acq[3];
    acquisition_values(A, mu, sd, acq);
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (((A.acq_mask >> a) & 1) && p.acq[a]) p.acq[a][i] = acq[a];
    if (!p.mu_grad || !p.sd_grad) return;
    const bool mask = sd > 0.0;
    double improve = 0.0, cdf = 0.0, pdf = 0.0;
    if (mask && (A.acq_mask & 3)) {
        improve = A.y_opt - A.xi - mu;
        const double z = improve / sd;
        cdf = normcdf(z);
        pdf = exp(-(z * z) / 2.0) * 0.3989422804014327;
    }
    for (int k = 0; k < p.d; ++k) {
        const int64_t o = i * p.d + k;
        const double gmu = p.mu_grad[o], gsd = p.sd_grad[o];
        if ((A.acq_mask & 4) && p.acq_grad[2]) p.acq_grad[2][o] = A.kappa_inf ? -gsd : gmu - A.kappa * gsd;
        if (A.acq_mask & 3) {
            double g_ei = 0.0, g_pi = 0.0;
            if (mask) {
                const double improve_grad = (-gmu * sd - gsd * improve) / (sd * sd);
                const double cdf_grad = improve_grad * pdf;
                g_pi = cdf_grad;
                const double pdf_grad = -improve * cdf_grad;
                g_ei = (-gmu * cdf - pdf_grad) + (gsd * pdf + pdf_grad);
            }
            if ((A.acq_mask & 1) && p.acq_grad[0]) p.acq_grad[0][o] = -g_ei;
            if ((A.acq_mask & 2) && p.acq_grad[1]) p.acq_grad[1][o] = -g_pi;
        }
    }
}

}  // namespace skb
