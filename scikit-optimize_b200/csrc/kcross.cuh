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
    const double* ls;         // [d] length scales (all ones for the Hamming kernel, which compares unscaled values)
    const double* hw;         // [d_pad] Hamming weights, zero padded (KIND 4 only; set by the launchers)
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
    if (KIND == 4) return exp_nonpos(-sq);                        // Hamming: sq = sum_k w_k [x_k != t_k]
    const double r = sqrt_nonneg(sq);
    if (KIND == 1) return exp_nonpos(-r);                         // Matern 1/2
    if (KIND == 2) { const double K = r * 1.7320508075688772; return (1.0 + K) * exp_nonpos(-K); }
    const double K = r * 2.23606797749979;                        // Matern 5/2
    return (1.0 + K + K * K * (1.0 / 3.0)) * exp_nonpos(-K);
}

constexpr size_t kcross_smem_bytes(int d_pad) {
    return sizeof(double) * ((size_t)d_pad * XS_LD + 2 * (size_t)XG * d_pad * 4 + 2 * XROWS + BN) + 2 * sizeof(uint64_t) +
           sizeof(double) * 2 * XROWS;                 // + [2][XROWS] squared row norms (6-plane digit kernel only)
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
                    if (KIND == 4) {       // skopt kernels.py:391-392: exp(-sum(length_scale * (x != y)))
                        const double w = __ldg(p.hw + k0 + kk);
                        s[gi][0] += x[kk] != t01.x ? w : 0.0;
                        s[gi][1] += x[kk] != t01.y ? w : 0.0;
                        s[gi][2] += x[kk] != t23.x ? w : 0.0;
                        s[gi][3] += x[kk] != t23.y ? w : 0.0;
                        continue;
                    }
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

// K1 for the split-integer contraction (ozaki.cuh): same distances, transform and running k . alpha, but the k-tile is
// emitted as P signed 8-bit digit planes (k = sK 2^-(8P-1) sum_q b_q 256^(P-1-q)) in the canonical K-major shared-memory
// layout of tcgen05.mma, [tile of 64 candidates][chunk of 32 training points][k half][plane][8 cand groups][8x16 B core],
// so a pipeline stage of the MMA kernel is again one contiguous bulk copy.  Thread (c, h) owns training rows
// [32 h, 32 h + 32) of every 64-row stage, i.e. one whole 32-point chunk, and stores 16 bytes per plane and k half.
struct KDigitParams {
    KCrossParams base;
    int8_t* Kd;               // [tiles64][n_pad/32][2 k halves][P planes][8 cand groups][8 cands][16 B]
    double inv_scale;         // 2^(8P-1) / sK
    const double* tn;         // [n_pad] squared norms of the scaled training rows (6-plane variant), zero padded
};

// sqrt(s) for finite s (negative -> 0) as in fastmath.cuh sqrt_nonneg, without its second coupled Newton step: after one
// step g and h carry a relative error e1 ~ 1.5 e0^2 (e0 <= 2^-20, the seed), and the residual correction leaves
// e1 * e1' ~ 1e-24, so the second step only matters to the last-bit reproducibility of the FP64 path, which keeps it.
// The caller clamps s to [1e-290, inf) first (no zero / negative / NaN special cases left in here).
__device__ __forceinline__ double sqrt_pos_short(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    double g = s * y;
    double h = 0.5 * y;
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, s);
    return fma(d, h, g);
}

// exp(x) for x <= 0 (or +1e-15-ish), not NaN: fastmath.cuh exp_nonpos with the degree-13 polynomial split into its even
// and odd halves (two dependency chains of 6 FMAs instead of one of 13) and without the NaN pass-through.
__device__ __forceinline__ double exp_nonpos_short(double x) {
    const double fn = fma(x, c_exp[0], c_exp[1]);
    const double n = fn - c_exp[1];
    double r = fma(n, c_exp[2], x);
    r = fma(n, c_exp[3], r);
    const double r2 = r * r;
    double po = c_exp[4], pe = c_exp[5];                      // r^13 ... r^1 coefficients / r^12 ... r^0 coefficients
#pragma unroll
    for (int i = 6; i <= 16; i += 2) {
        po = fma(po, r2, c_exp[i]);
        pe = fma(pe, r2, i + 1 <= 16 ? c_exp[i + 1] : 1.0);
    }
    const double p = fma(po, r, pe);
    const int ni = __double2loint(fn);
    const double scaled = __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
    return x < c_exp[17] ? 0.0 : scaled;
}

// kernel value from a squared distance that may be slightly negative (cancellation) but is not NaN
template <int KIND>
__device__ __forceinline__ double base_kernel_short(double sq) {
    if (KIND == 0) return exp_nonpos_short(-0.5 * sq);
    if (KIND == 4) return exp_nonpos_short(-sq);
    const double r = sqrt_pos_short(sq > 1e-290 ? sq : 1e-290);
    if (KIND == 1) return exp_nonpos_short(-r);
    if (KIND == 2) { const double K = r * 1.7320508075688772; return (1.0 + K) * exp_nonpos_short(-K); }
    const double K = r * 2.23606797749979;
    return (1.0 + K + K * K * (1.0 / 3.0)) * exp_nonpos_short(-K);
}

// Thread (cp, rq): candidates cp and cp + 64 of the 128-candidate tile (one in each 64-candidate MMA tile) x training
// rows [16 rq, 16 rq + 16) of every 64-row stage, i.e. one k-half of a 32-point chunk.  Two candidates share every
// broadcast read of a training row, which halves this kernel's shared-memory traffic -- it runs next to the
// shared-memory-bound contraction kernel (second stream), so that is what decides how much of it is hidden.
// P = number of digit planes emitted (7: 55-bit values, 6: 47-bit values; ozaki.cuh).
//
// The 6-plane variant rounds every kernel value to 2^-47 of the kernel scale (1.4e-14 absolute at scale 4), which buys
// three savings on the FP64 pipe that bounds this kernel (81 -> 59 FP64 instructions per (candidate, training point)
// at d = 20), none of them visible above that rounding:
//   * squared distances as |x|^2 + |t|^2 - 2 x.t (one FMA per feature instead of a subtraction and an FMA; the
//     cancellation error, ~1e-15 absolute in r^2, only matters for the nu = 1/2 kernel, whose value is not smooth in r^2
//     at 0 - that kernel keeps the direct differences);
//   * the square root without its redundant second Newton step (sqrt_nonneg_short);
//   * digit extraction by one FMA with 1.5 * 2^52 (the low 48 mantissa bits are then the rounded integer in two's
//     complement) instead of a multiply and a 64-bit float-to-integer conversion.
// The 7-plane variant is arithmetically the FP64 path's K1 (direct differences, fastmath.cuh).
template <int KIND, int P>
__global__ void __launch_bounds__(256, 2) kcross_digits_kernel(const KDigitParams q) {
    constexpr unsigned long long DIGIT_BIAS = 0x8080808080808080ull >> (8 * (8 - P));     // +128 in each of the P digit bytes
    constexpr bool FAST = P == 6;
    constexpr bool DOT = FAST && KIND != 1 && KIND != 4;
    const KCrossParams& p = q.base;
    extern __shared__ __align__(16) unsigned char smem_k1[];
    const int d_pad = p.d_pad;
    const int xt_stage = XG * d_pad * 4;
    double* xt = reinterpret_cast<double*>(smem_k1);
    double* al = xt + 2 * xt_stage;
    double* xs = al + 2 * XROWS;
    double* mpart = xs + (size_t)d_pad * XS_LD;              // [BN]; reused as [4][BN] staging through xt after the loop
    uint64_t* bars = reinterpret_cast<uint64_t*>(mpart + BN);
    double* tnb = reinterpret_cast<double*>(bars + 2);        // [2][XROWS] squared norms of the staged training rows (DOT)

    const int tid = threadIdx.x, cp = tid & 63, rq = tid >> 6;
    const int64_t tile = blockIdx.x;
    const int nst = (p.n + XROWS - 1) / XROWS;
    if (tid == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_fence_init();
    }
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
        mbar_arrive_expect_tx(bars + buf, (uint32_t)(xt_stage + XROWS + (DOT ? XROWS : 0)) * 8u);
        bulk_g2s(xt + buf * xt_stage, p.Xp + (size_t)st * xt_stage, (uint32_t)xt_stage * 8u, bars + buf);
        bulk_g2s(al + buf * XROWS, p.alpha_p + (size_t)st * XROWS, XROWS * 8u, bars + buf);
        if (DOT) bulk_g2s(tnb + buf * XROWS, q.tn + (size_t)st * XROWS, XROWS * 8u, bars + buf);
    };
    if (tid == 0) {
        issue(0);
        if (nst > 1) issue(1);
    }

    double macc[2] = {0.0, 0.0};
    double xn[2] = {0.0, 0.0};                                // |x|^2 of this thread's two candidates (6-plane variant)
    if (FAST) {
        for (int k = 0; k < d_pad; ++k) {
            const double a0 = xs[k * XS_LD + cp], a1 = xs[k * XS_LD + cp + 64];
            xn[0] = fma(a0, a0, xn[0]);
            xn[1] = fma(a1, a1, xn[1]);
        }
    }
    const int nkc_all = p.n_pad / 32;
    // candidate e (0/1) lives in MMA tile 2*tile + e at position cp: [k half][plane][cand group (8)][cand % 8][16 B]
    int8_t* kd_base = q.Kd + ((size_t)(tile * 2) * nkc_all) * (2 * P * 8 * 128) + (cp >> 3) * 128 + (cp & 7) * 16 +
                      (rq & 1) * (P * 8 * 128);
    const size_t tile64_stride = (size_t)nkc_all * (2 * P * 8 * 128);

    for (int st = 0; st < nst; ++st) {
        const int buf = st & 1;
        mbar_wait(bars + buf, (st >> 1) & 1);
        const double* xtb = xt + buf * xt_stage;
        const double* alb = al + buf * XROWS;

        double s[2][16];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int r = 0; r < 16; ++r) s[e][r] = 0.0;
        for (int k0 = 0; k0 < d_pad; k0 += 4) {
            double x0[4], x1[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                x0[kk] = xs[(k0 + kk) * XS_LD + cp];
                x1[kk] = xs[(k0 + kk) * XS_LD + cp + 64];
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const double* t = xtb + ((rq * 4 + g) * d_pad + k0) * 4;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const double2 t01 = *reinterpret_cast<const double2*>(t + kk * 4);
                    const double2 t23 = *reinterpret_cast<const double2*>(t + kk * 4 + 2);
                    if (DOT) {
                        s[0][g * 4 + 0] = fma(x0[kk], t01.x, s[0][g * 4 + 0]);
                        s[0][g * 4 + 1] = fma(x0[kk], t01.y, s[0][g * 4 + 1]);
                        s[0][g * 4 + 2] = fma(x0[kk], t23.x, s[0][g * 4 + 2]);
                        s[0][g * 4 + 3] = fma(x0[kk], t23.y, s[0][g * 4 + 3]);
                        s[1][g * 4 + 0] = fma(x1[kk], t01.x, s[1][g * 4 + 0]);
                        s[1][g * 4 + 1] = fma(x1[kk], t01.y, s[1][g * 4 + 1]);
                        s[1][g * 4 + 2] = fma(x1[kk], t23.x, s[1][g * 4 + 2]);
                        s[1][g * 4 + 3] = fma(x1[kk], t23.y, s[1][g * 4 + 3]);
                        continue;
                    }
                    if (KIND == 4) {
                        const double w = __ldg(p.hw + k0 + kk);
                        s[0][g * 4 + 0] += x0[kk] != t01.x ? w : 0.0;
                        s[0][g * 4 + 1] += x0[kk] != t01.y ? w : 0.0;
                        s[0][g * 4 + 2] += x0[kk] != t23.x ? w : 0.0;
                        s[0][g * 4 + 3] += x0[kk] != t23.y ? w : 0.0;
                        s[1][g * 4 + 0] += x1[kk] != t01.x ? w : 0.0;
                        s[1][g * 4 + 1] += x1[kk] != t01.y ? w : 0.0;
                        s[1][g * 4 + 2] += x1[kk] != t23.x ? w : 0.0;
                        s[1][g * 4 + 3] += x1[kk] != t23.y ? w : 0.0;
                        continue;
                    }
                    double a;
                    a = x0[kk] - t01.x; s[0][g * 4 + 0] = fma(a, a, s[0][g * 4 + 0]);
                    a = x0[kk] - t01.y; s[0][g * 4 + 1] = fma(a, a, s[0][g * 4 + 1]);
                    a = x0[kk] - t23.x; s[0][g * 4 + 2] = fma(a, a, s[0][g * 4 + 2]);
                    a = x0[kk] - t23.y; s[0][g * 4 + 3] = fma(a, a, s[0][g * 4 + 3]);
                    a = x1[kk] - t01.x; s[1][g * 4 + 0] = fma(a, a, s[1][g * 4 + 0]);
                    a = x1[kk] - t01.y; s[1][g * 4 + 1] = fma(a, a, s[1][g * 4 + 1]);
                    a = x1[kk] - t23.x; s[1][g * 4 + 2] = fma(a, a, s[1][g * 4 + 2]);
                    a = x1[kk] - t23.y; s[1][g * 4 + 3] = fma(a, a, s[1][g * 4 + 3]);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            // 16 kernel values -> digits; lo = bytes 0..3 (planes P-1..P-4), hi = bytes 4..P-1 (planes P-5..0)
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int j = st * XROWS + rq * 16 + r;
                double sq = s[e][r];
                // |x - t|^2 = (|x|^2 + |t|^2) - 2 x.t; a slightly negative result (x == t) becomes r = 0 in the root
                if (DOT) sq = fma(-2.0, sq, xn[e] + tnb[buf * XROWS + rq * 16 + r]);
                double val = fma(p.amplitude, FAST ? base_kernel_short<KIND>(sq) : base_kernel<KIND>(sq), p.bias);
                // padded training rows: L_inv's digit planes are zero there and alpha is zero padded, so the 6-plane variant
                // (whose values are always finite) needs no masking; the 7-plane variant keeps the FP64 path's exact zeros
                if (!FAST && j >= p.n) val = 0.0;
                macc[e] = fma(val, alb[rq * 16 + r], macc[e]);
                unsigned long long y;
                if (FAST)       // |val * inv_scale| <= 2^46: the sum's low 48 mantissa bits are the rounded integer
                    y = ((unsigned long long)__double_as_longlong(fma(val, q.inv_scale, 6755399441055744.0)) + DIGIT_BIAS) ^ DIGIT_BIAS;
                else
                    y = ((unsigned long long)__double2ll_rn(val * q.inv_scale) + DIGIT_BIAS) ^ DIGIT_BIAS;
                lo[r] = (uint32_t)y;
                hi[r] = (uint32_t)(y >> 32);
            }
            int8_t* dst = kd_base + e * tile64_stride + (size_t)(st * 2 + (rq >> 1)) * (2 * P * 8 * 128);
#pragma unroll
            for (int pl = 0; pl < P; ++pl) {
                const int byte = P - 1 - pl;                               // plane 0 = most significant digit = byte P-1
                uint32_t w[4];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const int r0 = k4 * 4;
                    uint32_t a0, a1, a2, a3;
                    if (byte < 4) { a0 = lo[r0]; a1 = lo[r0 + 1]; a2 = lo[r0 + 2]; a3 = lo[r0 + 3]; }
                    else          { a0 = hi[r0]; a1 = hi[r0 + 1]; a2 = hi[r0 + 2]; a3 = hi[r0 + 3]; }
                    const int b = byte & 3;
                    const uint32_t t0 = __byte_perm(a0, a1, b | ((4 + b) << 4));          // [a0.b, a1.b, -, -]
                    const uint32_t t1 = __byte_perm(a2, a3, b | ((4 + b) << 4));
                    w[k4] = __byte_perm(t0, t1, 0x5410);                                  // [a0.b, a1.b, a2.b, a3.b]
                }
                *reinterpret_cast<uint4*>(dst + pl * (8 * 128)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        __syncthreads();
        if (tid == 0 && st + 2 < nst) issue(st + 2);
    }
    // k . alpha: the four row quarters of a candidate are combined in a fixed order (staged through the xt buffers)
    double* stage = xt;                                       // [4][BN], free after the loop
    if (FAST) {
        // the 6-plane transform clamps its argument, which would swallow a NaN / Inf coordinate: hand it to the mean
        // (0 * NaN = 0 * Inf = NaN), from where it reaches every output like in the other paths
        macc[0] = fma(0.0, xn[0], macc[0]);
        macc[1] = fma(0.0, xn[1], macc[1]);
    }
    stage[rq * BN + cp] = macc[0];
    stage[rq * BN + cp + 64] = macc[1];
    __syncthreads();
    if (tid < BN)
        p.kdot[tile * BN + tid] = ((stage[tid] + stage[BN + tid]) + stage[2 * BN + tid]) + stage[3 * BN + tid];
}

// 6-plane K1 with the x.t part of the squared distances on FP64 tensor-core tiles (DMMA m8n8k4) instead of one DFMA per
// (candidate, training point, feature): the same FP64 pipe time (DMMA and DFMA share the pipe at the same flop rate), an
// eighth of the issue slots and a quarter of the shared-memory reads for that part, which is a third of the kernel's FP64
// work at d = 20.  Warp w = (row quarter rq = w >> 1, MMA tile e = w & 1) owns candidates [64 e, 64 e + 64) x training rows
// [16 rq, 16 rq + 16) of every 64-row stage as 8 x 2 accumulator tiles of 8 x 8: rows of a tile are candidates
// (A operand, from the transposed candidate tile xs), columns are training points (B operand).  The training rows of a
// stage may be taken in any order, so column c of tile ji is row 4 (c / 2) + 2 ji + c % 2 of the 16-row block: lane l then
// holds, for each of its 8 candidates 8 ci + l / 4, the FOUR CONSECUTIVE training rows 4 (l % 4) .. + 3, i.e. one aligned
// 32-bit word of the 16-byte K-major digit vector the contraction kernel reads, and a warp stores 128 contiguous bytes
// per (plane, candidate group).  Transform, digit extraction and the k . alpha reduction are those of
// kcross_digits_kernel<KIND, 6>; kinds whose distance needs direct differences (Matern 1/2, Hamming) stay there.
// sqrt(s) for s >= 1e-290 by ONE third-order step from the hardware seed y ~ 1/sqrt(s) (relative error |d| <= 2^-20):
// with g = s y and e = 1 - g y (one FMA, = 1 - (1 + d)^2), sqrt(s) = g (1 - e)^(-1/2) = g (1 + e/2 + 3 e^2/8 + O(e^3)); the
// dropped term is below 2^-58.  Five FP64 instructions instead of the seven of sqrt_pos_short, <= 1.5 ulp.
__device__ __forceinline__ double sqrt_pos_third(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double g = s * y;
    const double e = fma(-g, y, 1.0);
    return fma(g * e, fma(0.375, e, 0.5), g);
}

// exp(x) for x <= 0 (or +1e-15-ish), not NaN, by a 32-entry table: x = (32 m + j) ln2 / 32 + r with |r| <= ln2 / 64, so
// exp(x) = 2^m * 2^(j/32) * exp(r) and exp(r) needs a degree-6 polynomial (remainder r^7 / 5040 < 4e-18) instead of the
// degree-13 one of exp_nonpos_short: 11 instead of 18 FP64 instructions per kernel value (K1 spends 53 per value at d = 20).
// The table (correctly rounded 2^(j/32)) is staged in shared memory by the kernel: lanes index it independently.
__constant__ double c_exp32[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924,
    1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832, 1.4142135623730951, 1.4451808069770467,
    1.4768261459394993, 1.5091644275934228, 1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103,
    1.9152065613971474, 1.9571441241754002};
__constant__ double c_exp32k[10] = {
    46.16624130844683,                  // [0] 32 / ln2
    6755399441055744.0,                 // [1] 1.5 * 2^52
    -0.02166084938653512,               // [2] -ln2 / 32, high part (21 trailing zero bits: n * hi is exact for |n| < 2^16)
    -5.9631716539705866e-12,            // [3] -ln2 / 32, low part
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5,      // [4..8]
    -708.0                              // [9]
};
__device__ __forceinline__ double exp_nonpos_table(double x, const double* __restrict__ tab) {
    const double fn = fma(x, c_exp32k[0], c_exp32k[1]);
    const double n = fn - c_exp32k[1];
    double r = fma(n, c_exp32k[2], x);
    r = fma(n, c_exp32k[3], r);
    double p = fma(r, c_exp32k[4], c_exp32k[5]);
    p = fma(p, r, c_exp32k[6]);
    p = fma(p, r, c_exp32k[7]);
    p = fma(p, r, c_exp32k[8]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int ni = __double2loint(fn);                        // low word of (1.5 * 2^52 + n): n in two's complement
    p *= tab[ni & 31];
    const double scaled = __hiloint2double(__double2hiint(p) + ((ni >> 5) << 20), __double2loint(p));
    return x < c_exp32k[9] ? 0.0 : scaled;
}

// Distance scale folded into the squared distance: the transform receives A * r^2 with A = -1/2 (RBF), 3 (Matern 3/2),
// 5 (Matern 5/2), so that the square root IS the kernel argument K and no separate multiplication is needed.
template <int KIND>
__device__ __forceinline__ constexpr double dist_scale() { return KIND == 0 ? -0.5 : (KIND == 2 ? 3.0 : 5.0); }

// amplitude * base(r) + bias from the scaled squared distance; amp3 = amplitude / 3 (Matern 5/2 only).  The amplitude is
// folded into the polynomial: amp (1 + K + K^2/3) e^-K + bias = fma(fma(K, fma(K, amp3, amp), amp), e^-K, bias).
template <int KIND>
__device__ __forceinline__ double kernel_value_scaled(double sa, double amp, double amp3, double bias, const double* __restrict__ etab) {
    if (KIND == 0) return fma(amp, exp_nonpos_table(sa, etab), bias);
    const double K = sqrt_pos_third(sa > 1e-290 ? sa : 1e-290);
    const double e = exp_nonpos_table(-K, etab);
    if (KIND == 2) return fma(fma(K, amp, amp), e, bias);
    return fma(fma(K, fma(K, amp3, amp), amp), e, bias);
}

// CB = candidates per CTA: 128 (8 warps, two CTAs per SM: the stand-alone configuration) or 64 (4 warps, 16 K registers
// and ~40 KB of shared memory: small enough to share an SM with one CTA of the contraction kernel, skb.cu split_overlaps).
// The row stride CB + 8 of the transposed candidate tile is 8 mod 16 doubles: the A-fragment read (4 features x 8
// candidates) is conflict-free.
constexpr size_t kcross_mma_smem_bytes(int d_pad, int cb = BN) {
    return sizeof(double) * ((size_t)d_pad * (cb + 8) + 2 * (size_t)XG * d_pad * 4 + 2 * XROWS + 2 * XROWS + cb + 32) +
           2 * sizeof(uint64_t);
}

template <int KIND, int CB = BN, int NCT = 64>
__global__ void __launch_bounds__(2 * CB, 256 / CB) kcross_digits_mma_kernel(const KDigitParams q) {
    constexpr int P = 6;
    static_assert(NCT == 64 || (NCT == 80 && CB == 128), "contraction tiles of 64 candidates, or 80 with the 128-candidate CTA");
    constexpr int XS_LD_MMA = CB + 8;
    constexpr int NT = 2 * CB;
    constexpr unsigned long long DIGIT_BIAS = 0x8080808080808080ull >> (8 * (8 - P));
    const KCrossParams& p = q.base;
    extern __shared__ __align__(16) unsigned char smem_k1[];
    const int d_pad = p.d_pad;
    const int xt_stage = XG * d_pad * 4;
    double* xt = reinterpret_cast<double*>(smem_k1);          // [2][XG][d_pad][4]
    double* al = xt + 2 * xt_stage;                            // [2][XROWS]
    double* tnb = al + 2 * XROWS;                              // [2][XROWS] squared norms of the staged training rows
    double* xs = tnb + 2 * XROWS;                              // [d_pad][XS_LD_MMA]
    double* xnorm = xs + (size_t)d_pad * XS_LD_MMA;            // [CB] |x|^2 of the candidates
    double* etab = xnorm + CB;                                 // [32] 2^(j/32) for exp_nonpos_table
    uint64_t* bars = reinterpret_cast<uint64_t*>(etab + 32);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rq = CB == 128 ? warp >> 1 : warp, e = CB == 128 ? warp & 1 : 0, t4 = lane & 3, c8 = lane >> 2;
    const int64_t tile = blockIdx.x;
    const int nst = (p.n + XROWS - 1) / XROWS;
    if (tid == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_fence_init();
    }
    if (tid < 32) etab[tid] = c_exp32[tid];
    for (int idx = tid; idx < CB * d_pad; idx += NT) {
        const int cc = idx / d_pad, k = idx - cc * d_pad;
        const int64_t row = tile * CB + cc;
        double v = 0.0;
        if (k < p.d && row < p.m) v = p.X[row * p.d + k] / p.ls[k];
        xs[k * XS_LD_MMA + cc] = v;
    }
    __syncthreads();
    auto issue = [&](int st) {
        const int buf = st & 1;
        mbar_arrive_expect_tx(bars + buf, (uint32_t)(xt_stage + 2 * XROWS) * 8u);
        bulk_g2s(xt + buf * xt_stage, p.Xp + (size_t)st * xt_stage, (uint32_t)xt_stage * 8u, bars + buf);
        bulk_g2s(al + buf * XROWS, p.alpha_p + (size_t)st * XROWS, XROWS * 8u, bars + buf);
        bulk_g2s(tnb + buf * XROWS, q.tn + (size_t)st * XROWS, XROWS * 8u, bars + buf);
    };
    if (tid == 0) {
        issue(0);
        if (nst > 1) issue(1);
    }
    if (tid < CB) {
        double a = 0.0;
        for (int k = 0; k < d_pad; ++k) a = fma(xs[k * XS_LD_MMA + tid], xs[k * XS_LD_MMA + tid], a);
        xnorm[tid] = a;
    }
    __syncthreads();

    constexpr double A = dist_scale<KIND>();                   // the transform works on A * r^2 (kernel_value_scaled)
    const double amp3 = p.amplitude * (1.0 / 3.0);
    double xn[8], macc[8];                                     // xn: A |x|^2
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
        xn[ci] = A * xnorm[64 * e + 8 * ci + c8];
        macc[ci] = 0.0;
    }
    const int nkc_all = p.n_pad / 32;
    // Digit bytes go where the contraction kernel reads them: per contraction tile of NC = 8 G candidates and per chunk of 32
    // training rows, [k half][plane][candidate group of 8][candidate][16 B]; this lane's word is bytes 4 t4 .. + 3 of a
    // candidate's 16-byte vector.  The CTA's 8-candidate groups are numbered through the whole candidate stream
    // (group index / G = contraction tile, % G = group inside it), so the same kernel serves 64- and 80-candidate tiles
    // (NCT): with G = 8 a warp's eight groups fill exactly one tile, with G = 10 they straddle at most one tile boundary;
    // groups past it (ci >= ci_wrap) land `wrap` bytes further on, in the next tile (at most 8 MB: n_pad <= 16384).
    constexpr int G = NCT / 8;
    constexpr uint32_t chunk_bytes = 2 * P * G * 128;
    const int64_t grp0 = tile * (CB / 8) + 8 * e;             // first of this warp's eight groups
    const int g0 = (int)(grp0 % G);
    const int ci_wrap = G - g0;                                // >= 8 when G == 8: never
    const uint32_t wrap = (uint32_t)nkc_all * chunk_bytes - (uint32_t)G * 128;
    int8_t* kd_base = q.Kd + (size_t)(grp0 / G) * nkc_all * chunk_bytes + (size_t)(rq & 1) * (P * G * 128) + (size_t)g0 * 128 +
                      c8 * 16 + 4 * t4;
    constexpr int plane_stride = G * 128;
    const double* xa = xs + t4 * XS_LD_MMA + 64 * e + c8;      // A fragment: feature k0 + t4, candidate 8 ci + c8

    for (int st = 0; st < nst; ++st) {
        const int buf = st & 1;
        mbar_wait(bars + buf, (st >> 1) & 1);
        // B fragment: feature k0 + t4 of the training row behind column c8: group 4 rq + c8 / 2, row-in-group 2 ji + c8 % 2
        const double* xb = xt + buf * xt_stage + ((size_t)(rq * 4 + (c8 >> 1)) * d_pad + t4) * 4 + (c8 & 1);
        const double* alb = al + buf * XROWS + rq * 16 + 4 * t4;
        const double* tn4 = tnb + buf * XROWS + rq * 16 + 4 * t4;

        double acc[8][2][2];
#pragma unroll
        for (int ci = 0; ci < 8; ++ci)
#pragma unroll
            for (int ji = 0; ji < 2; ++ji) acc[ci][ji][0] = acc[ci][ji][1] = 0.0;
        for (int k0 = 0; k0 < d_pad; k0 += 4) {
            const double b0 = xb[k0 * 4], b1 = xb[k0 * 4 + 2];
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
                const double a = xa[k0 * XS_LD_MMA + 8 * ci];
                dmma884(acc[ci][0][0], acc[ci][0][1], a, b0);
                dmma884(acc[ci][1][0], acc[ci][1][1], a, b1);
            }
        }
        double tnr[4], alr[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            tnr[r] = A * tn4[r];
            alr[r] = alb[r];
        }
        int8_t* dst = kd_base + (size_t)(st * 2 + (rq >> 1)) * chunk_bytes;
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {                     // training row 4 t4 + r = tile r >> 1, accumulator r & 1
                // A |x - t|^2 = A |x|^2 + A |t|^2 - 2 A x.t; a slightly wrong-signed result (x == t) becomes K = 0
                const double sa = fma(-2.0 * A, acc[ci][r >> 1][r & 1], xn[ci] + tnr[r]);
                const double val = kernel_value_scaled<KIND>(sa, p.amplitude, amp3, p.bias, etab);
                macc[ci] = fma(val, alr[r], macc[ci]);
                const unsigned long long y =
                    ((unsigned long long)__double_as_longlong(fma(val, q.inv_scale, 6755399441055744.0)) + DIGIT_BIAS) ^ DIGIT_BIAS;
                lo[r] = (uint32_t)y;
                hi[r] = (uint32_t)(y >> 32);
            }
#pragma unroll
            for (int pl = 0; pl < P; ++pl) {
                const int byte = P - 1 - pl;                   // plane 0 = most significant digit
                const int b = byte & 3;
                const uint32_t* src = byte < 4 ? lo : hi;
                const uint32_t t0 = __byte_perm(src[0], src[1], b | ((4 + b) << 4));
                const uint32_t t1 = __byte_perm(src[2], src[3], b | ((4 + b) << 4));
                *reinterpret_cast<uint32_t*>(dst + pl * plane_stride + ci * 128 + (G != 8 && ci >= ci_wrap ? wrap : 0u)) = __byte_perm(t0, t1, 0x5410);
            }
        }
        __syncthreads();
        if (tid == 0 && st + 2 < nst) issue(st + 2);
    }
    // k . alpha: the four lanes that share a candidate (fixed xor tree), then the four row quarters in a fixed order
    double* stage = xt;                                        // [4][CB], free after the loop
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
        double v = fma(0.0, xn[ci], macc[ci]);                 // NaN / Inf coordinates reach the mean (see the DFMA variant)
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (t4 == 0) stage[rq * CB + 64 * e + 8 * ci + c8] = v;
    }
    __syncthreads();
    if (tid < CB)
        p.kdot[tile * CB + tid] = ((stage[tid] + stage[CB + tid]) + stage[2 * CB + tid]) + stage[3 * CB + tid];
}

}  // namespace skb
