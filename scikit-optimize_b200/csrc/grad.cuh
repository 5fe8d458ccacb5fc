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
constexpr int GRAD_THREADS = 256;    // thread (c, h): candidate c, half h of the rows of every stage
constexpr int GSC = 8;               // per-candidate scalars exchanged through shared memory

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
    int frag0;                // first 8-dimension fragment this launch accumulates gradients for (0; 8 in the second pass of n_dims > 64)
};

constexpr size_t grad_smem_bytes(int d_pad) {
    return sizeof(double) * (2 * (size_t)GG * d_pad * 4 + 2 * GROWS + 2 * (size_t)GROWS * BN + (size_t)d_pad * XS_LD +
                             (size_t)GSC * BN) + 2 * sizeof(uint64_t);
}

// value and radial derivative coefficient: k = amplitude*base + bias, G_jk = gcoef * (x~_k - X~_jk) / l_k
template <int KIND>
__device__ __forceinline__ void kernel_and_gcoef(double sq, double amplitude, double bias, double& kval, double& gcoef,
                                                 bool& at_zero) {
    at_zero = false;
    if (KIND == 0) {
        const double e = exp_nonpos(-0.5 * sq);
        kval = fma(amplitude, e, bias);
        gcoef = -amplitude * e;
        return;
    }
    const double r = sqrt_nonneg(sq);
    if (KIND == 1) {
        const double e = exp_nonpos(-r);
        kval = fma(amplitude, e, bias);
        at_zero = (r == 0.0);                       // reference: gradient row = -1/length_scale (kernels.py:118-128)
        gcoef = at_zero ? 0.0 : -amplitude * e / r;
    } else if (KIND == 2) {
        const double K = r * 1.7320508075688772;
        const double e = exp_nonpos(-K);
        kval = fma(amplitude, (1.0 + K) * e, bias);
        gcoef = -3.0 * amplitude * e;
    } else {
        const double K = r * 2.23606797749979;
        const double e = exp_nonpos(-K);
        kval = fma(amplitude, (1.0 + K + K * K * (1.0 / 3.0)) * e, bias);
        gcoef = -(5.0 / 3.0) * amplitude * (1.0 + K) * e;
    }
}

// Per stage of 32 training rows:
//   phase A (DFMA, thread = candidate x row half): distances -> k_j, gcoef_j; running k.alpha, q = k.W, and the two
//           weight tiles P1 = alpha_j gcoef_j, P2 = W_j gcoef_j written to shared memory in k4-interleaved layout;
//   phase B (DMMA): S[k, c] += sum_j X~_jk P_jc for both weight tiles -- the staged training rows already are the
//           A operand (dims x rows) in the layout the fragments need, the P tiles are the B operand.
// The gradient sums follow from  sum_j P_jc (x~_ck - X~_jk) = x~_ck * sum_j P_jc - S[k, c].
// MF = number of 8-dimension fragments (ceil(d / 8), at most 8) the accumulators are sized for.  More than 64 dimensions
// take a second launch with frag0 = 8: it repeats phase A (distances need every dimension) and accumulates the gradient
// columns 64 .. d - 1; the per-candidate scalars it writes are the first pass's, bit for bit.
template <int KIND, int MF>
__global__ void __launch_bounds__(GRAD_THREADS, 1) grad_contract_kernel(const GradParams p) {
    extern __shared__ __align__(16) unsigned char smem_k3[];
    const int d_pad = p.d_pad;
    const int xt_stage = GG * d_pad * 4;
    double* xt = reinterpret_cast<double*>(smem_k3);        // [2][GG][d_pad][4]
    double* al = xt + 2 * xt_stage;                          // [2][GROWS]
    double* P = al + 2 * GROWS;                              // [2][GG][BN][4]
    double* xs = P + 2 * GROWS * BN;                         // [d_pad][XS_LD]
    double* sc = xs + (size_t)d_pad * XS_LD;                 // [GSC][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sc + GSC * BN);

    const int tid = threadIdx.x, c = tid & (BN - 1), h = tid >> 7;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    const int nst = (p.n + GROWS - 1) / GROWS;

    if (tid == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_fence_init();
    }
    for (int idx = tid; idx < BN * d_pad; idx += GRAD_THREADS) {
        const int cc = idx / d_pad, k = idx - cc * d_pad;
        const int64_t row = tile * BN + cc;
        double v = 0.0;
        if (k < p.d && row < p.m) v = p.X[row * p.d + k] / p.ls[k];
        xs[k * XS_LD + cc] = v;
    }
    __syncthreads();

    auto issue = [&](int st) {
        const int buf = st & 1;
        mbar_arrive_expect_tx(bars + buf, (uint32_t)(xt_stage + GROWS) * 8u);
        bulk_g2s(xt + buf * xt_stage, p.Xp + (size_t)st * xt_stage, (uint32_t)xt_stage * 8u, bars + buf);
        bulk_g2s(al + buf * GROWS, p.alpha_p + (size_t)st * GROWS, GROWS * 8u, bars + buf);
    };
    if (tid == 0) {
        issue(0);
        if (nst > 1) issue(1);
    }

    const double* wt = p.Wt + (size_t)tile * p.n_pad * BN + c;
    double kdot = 0.0, q = 0.0, sum1 = 0.0, sum2 = 0.0, z_alpha = 0.0, z_w = 0.0;   // z_*: Matern-1/2 rows with r == 0
    double D[2][MF][2][2];                                   // [weight tile][dim fragment][candidate fragment][c0,c1]
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int mi = 0; mi < MF; ++mi)
#pragma unroll
            for (int nf = 0; nf < 2; ++nf) D[a][mi][nf][0] = D[a][mi][nf][1] = 0.0;

    for (int st = 0; st < nst; ++st) {
        const int buf = st & 1;
        double w[GROWS / 2];
#pragma unroll
        for (int gi = 0; gi < GG / 2; ++gi)
#pragma unroll
            for (int r = 0; r < 4; ++r) w[gi * 4 + r] = wt[(size_t)(st * GROWS + (2 * gi + h) * 4 + r) * BN];
        mbar_wait(bars + buf, (st >> 1) & 1);
        const double* xtb = xt + buf * xt_stage;
        const double* alb = al + buf * GROWS;

        // ---------------- phase A ----------------
        double s[GROWS / 2];
#pragma unroll
        for (int r = 0; r < GROWS / 2; ++r) s[r] = 0.0;
        for (int k0 = 0; k0 < d_pad; k0 += 4) {
            double x[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) x[kk] = xs[(k0 + kk) * XS_LD + c];
#pragma unroll
            for (int gi = 0; gi < GG / 2; ++gi) {
                const double* t = xtb + ((2 * gi + h) * d_pad + k0) * 4;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const double2 t01 = *reinterpret_cast<const double2*>(t + kk * 4);
                    const double2 t23 = *reinterpret_cast<const double2*>(t + kk * 4 + 2);
                    const double d0 = x[kk] - t01.x, d1 = x[kk] - t01.y, d2 = x[kk] - t23.x, d3 = x[kk] - t23.y;
                    s[gi * 4 + 0] = fma(d0, d0, s[gi * 4 + 0]);
                    s[gi * 4 + 1] = fma(d1, d1, s[gi * 4 + 1]);
                    s[gi * 4 + 2] = fma(d2, d2, s[gi * 4 + 2]);
                    s[gi * 4 + 3] = fma(d3, d3, s[gi * 4 + 3]);
                }
            }
        }
#pragma unroll
        for (int gi = 0; gi < GG / 2; ++gi) {
            const int g = 2 * gi + h;
            double p1[4], p2[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                double kval, gcoef;
                bool at_zero;
                kernel_and_gcoef<KIND>(s[gi * 4 + r], p.amplitude, p.bias, kval, gcoef, at_zero);
                const double a_r = alb[g * 4 + r];
                double w_r = w[gi * 4 + r];
                if (st * GROWS + g * 4 + r >= p.n) { kval = 0.0; gcoef = 0.0; at_zero = false; w_r = 0.0; }
                kdot = fma(kval, a_r, kdot);
                q = fma(kval, w_r, q);
                if (KIND == 1 && at_zero) { z_alpha += a_r; z_w += w_r; }
                p1[r] = a_r * gcoef;
                p2[r] = w_r * gcoef;
                sum1 += p1[r];
                sum2 += p2[r];
            }
            double* dst = P + ((size_t)g * BN + c) * 4;
            *reinterpret_cast<double2*>(dst) = make_double2(p1[0], p1[1]);
            *reinterpret_cast<double2*>(dst + 2) = make_double2(p1[2], p1[3]);
            *reinterpret_cast<double2*>(dst + GROWS * BN) = make_double2(p2[0], p2[1]);
            *reinterpret_cast<double2*>(dst + GROWS * BN + 2) = make_double2(p2[2], p2[3]);
        }
        __syncthreads();                                     // both weight tiles are complete

        // ---------------- phase B: warp owns candidate fragments 2*warp, 2*warp+1 ----------------
#pragma unroll
        for (int g = 0; g < GG; ++g) {
            double b[2][2];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int nf = 0; nf < 2; ++nf) b[a][nf] = P[((size_t)(a * GG + g) * BN + (2 * warp + nf) * 8) * 4 + lane];
#pragma unroll
            for (int mi = 0; mi < MF; ++mi) {
                const int dim = (p.frag0 + mi) * 8 + (lane >> 2);
                const double av = dim < d_pad ? xtb[(g * d_pad + dim) * 4 + (lane & 3)] : 0.0;
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int nf = 0; nf < 2; ++nf) dmma884(D[a][mi][nf][0], D[a][mi][nf][1], av, b[a][nf]);
            }
        }
        __syncthreads();                                     // everyone is done with P and with buffer `buf`
        if (tid == 0 && st + 2 < nst) issue(st + 2);
    }

    // ---------------- per-candidate scalars: combine the two row halves in a fixed order ----------------
    if (h == 1) {
        sc[0 * BN + c] = kdot; sc[1 * BN + c] = q; sc[2 * BN + c] = sum1;
        sc[3 * BN + c] = sum2; sc[4 * BN + c] = z_alpha; sc[5 * BN + c] = z_w;
    }
    __syncthreads();
    const AcqConsts& A = p.a;
    const int64_t cand = tile * BN + c;
    double kd1 = 0, q1 = 0, s11 = 0, s21 = 0, za1 = 0, zw1 = 0;
    if (h == 0) {
        kd1 = sc[0 * BN + c]; q1 = sc[1 * BN + c]; s11 = sc[2 * BN + c];
        s21 = sc[3 * BN + c]; za1 = sc[4 * BN + c]; zw1 = sc[5 * BN + c];
    }
    __syncthreads();
    if (h == 0) {
        kdot += kd1; q += q1; sum1 += s11; sum2 += s21; z_alpha += za1; z_w += zw1;
        const double mu = A.y_std * kdot + A.y_mean;
        double var = A.prior_var - q;
        if (var < 0.0) var = 0.0;
        var = var * (A.y_std * A.y_std);
        const double sd = sqrt(var);
        double acq[3];
        acquisition_values(A, mu, sd, acq);
        double improve = 0.0, cdf = 0.0, pdf = 0.0;
        if (sd > 0.0 && (A.acq_mask & 3)) {
            improve = A.y_opt - A.xi - mu;
            const double z = improve / sd;
            cdf = normcdf(z);
            pdf = exp(-(z * z) / 2.0) * 0.3989422804014327;
        }
        if (cand < p.m) {
            if (p.mean) p.mean[cand] = mu;
            if (p.std) p.std[cand] = sd;
#pragma unroll
            for (int a = 0; a < 3; ++a)
                if (((A.acq_mask >> a) & 1) && p.acq[a]) p.acq[a][cand] = acq[a];
        }
        sc[0 * BN + c] = sd;
        sc[1 * BN + c] = sum1;
        sc[2 * BN + c] = sum2;
        sc[3 * BN + c] = improve;
        sc[4 * BN + c] = cdf;
        sc[5 * BN + c] = pdf;
        sc[6 * BN + c] = p.amplitude * z_alpha;              // Matern-1/2 rows at r == 0 contribute -amplitude / l_k
        sc[7 * BN + c] = p.amplitude * z_w;
    }
    __syncthreads();

    // ---------------- gradients straight from the accumulator fragments ----------------
    // lane holds dims (frag0+mi)*8 + (lane>>2) and candidates (2*warp+nf)*8 + 2*(lane&3) + {0,1}
#pragma unroll
    for (int mi = 0; mi < MF; ++mi) {
        const int k = (p.frag0 + mi) * 8 + (lane >> 2);
        if (k >= p.d) continue;
        const double inv_l = 1.0 / p.ls[k];
#pragma unroll
        for (int nf = 0; nf < 2; ++nf)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cc = (2 * warp + nf) * 8 + 2 * (lane & 3) + e;
                const int64_t cd = tile * BN + cc;
                if (cd >= p.m) continue;
                const double x = xs[k * XS_LD + cc];
                const double sd = sc[0 * BN + cc];
                double tm = x * sc[1 * BN + cc] - D[0][mi][nf][e];
                double ts = x * sc[2 * BN + cc] - D[1][mi][nf][e];
                if (KIND == 1) { tm -= sc[6 * BN + cc]; ts -= sc[7 * BN + cc]; }
                const bool sd_is_zero = !(fabs(sd) > 1e-8);  // np.allclose(y_std, 0): |std| <= atol (gpr.py:351)
                const bool mask = sd > 0.0;                   // acquisition.py:213,296
                const double gmu = A.y_std * (tm * inv_l);
                const double gsd = sd_is_zero ? 0.0 : -(ts * inv_l) / sd * (A.y_std * A.y_std);
                const int64_t o = cd * p.d + k;
                if (p.mean_grad) p.mean_grad[o] = gmu;
                if (p.std_grad) p.std_grad[o] = gsd;
                if ((A.acq_mask & 4) && p.acq_grad[2]) p.acq_grad[2][o] = A.kappa_inf ? -gsd : gmu - A.kappa * gsd;
                if (A.acq_mask & 3) {
                    double g_ei = 0.0, g_pi = 0.0;
                    if (mask) {
                        const double improve = sc[3 * BN + cc], cdf = sc[4 * BN + cc], pdf = sc[5 * BN + cc];
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

// Per-second acquisitions EIps / PIps (acquisition.py:61-80): acq / E[t] with a log-normal time model,
//   inv_t = exp(-mu_t + std_t^2 / 2),  acq' = acq * inv_t,  grad' = grad * inv_t + acq' * (-grad mu_t + std_t grad std_t).
struct PsParams {
    const double* acq;        // [m]   values to minimise of the objective model (-EI or -PI)
    const double* acq_grad;   // [m][d] or nullptr
    const double* mu_t;       // [m]   posterior of the log-time model
    const double* sd_t;       // [m]
    const double* mu_t_grad;  // [m][d] or nullptr
    const double* sd_t_grad;  // [m][d] or nullptr
    int64_t m;
    int d;
    double* out;              // [m]
    double* out_grad;         // [m][d] or nullptr
};

__global__ void ps_combine_kernel(const PsParams p) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= p.m) return;
    const double sd = p.sd_t[i];
    const double inv_t = exp(-p.mu_t[i] + 0.5 * (sd * sd));
    const double v = p.acq[i] * inv_t;
    p.out[i] = v;
    if (!p.out_grad) return;
    for (int k = 0; k < p.d; ++k) {
        const int64_t o = i * p.d + k;
        p.out_grad[o] = p.acq_grad[o] * inv_t + v * (-p.mu_t_grad[o] + sd * p.sd_t_grad[o]);
    }
}

}  // namespace skb
