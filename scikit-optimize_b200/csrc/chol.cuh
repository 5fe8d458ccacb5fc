// Dense FP64 factorisation kernels of the fit objective: what scikit-learn's
// log_marginal_likelihood does with LAPACK (sklearn/gaussian_process/_gpr.py:541-657: cholesky, cho_solve, and for the
// gradient K^-1) and what skopt's fit adds (gpr.py:223-224: L^-1, K^-1 = L^-T L^-1), written for one B200:
//
//   chol_panel_kernel  : block column k of a right-looking Cholesky.  Every CTA refactorises the 64 x 64 diagonal block
//                        in shared memory (7 us; cheaper than a launch + a trip through L2 for a broadcast), inverts it in
//                        the same sweep, and turns its own 64-row slab of the panel into L by ONE small matrix product with
//                        that inverse (no triangular solve anywhere).  The diagonal CTA keeps the inverse: tri_inverse needs it.
//   chol_update_kernel : trailing update A_ij -= L_ik L_jk^T of the lower tiles, 64 x 64 x 64 per CTA.
//   tri_inverse        : X = L^-1 by recursive doubling, X21 = -X22 (L21 X11): 2 log2(n / 64) launches of tile products
//                        instead of n / 64 dependent steps.
//   tri_gram_kernel    : K^-1 = X^T X (lower triangle), the formula of gpr.py:224.
//   tri_matvec_*       : alpha = X^T (X y) (cho_solve with the explicit inverse; X is needed anyway).
//
// All matrices are n x n column-major with the lower triangle significant (element (i, j), i >= j, at A[j n + i]) - the
// layout the kernel-matrix builder writes and the gradient kernel reads.  n is arbitrary; partial blocks are padded with
// the identity inside shared memory.  Reductions run in a fixed order: results are reproducible run to run.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace skb {

constexpr int CH_NB = 64;                 // block size
constexpr int CH_LD = CH_NB + 4;          // shared-memory row stride: 4 mod 16 doubles, so that the DMMA fragment reads (4 k x 8 rows)
                                          // and the 4 x 4 patches of the tile loader both touch 16 different 8-byte banks per half-warp
constexpr int CH_THREADS = 256;
constexpr size_t ch_smem_bytes() { return 3 * (size_t)CH_NB * CH_LD * sizeof(double); }

// dst[r][c] = src(row0 + r, col0 + c) (TRANS = false) or src(row0 + c, col0 + r) (TRANS = true); zero outside the matrix.
// A half-warp moves a 4 x 4 patch: four rows of a column are one 32-byte sector of the column-major source, and the 16
// shared-memory words land in 16 different banks in both modes (row stride CH_LD = 4 mod 16).
// The two halves of a tile load: global -> registers (16 independent loads per thread, all in flight at once) and
// registers -> shared memory.  Kernels fetch every tile they need before they store the first one (one round trip to
// L2 / HBM instead of one per tile), and the kernels that loop over k blocks fetch the next pair while the tensor-core
// tiles of the current pair run.
__device__ __forceinline__ void ch_fetch_tile(double (&v)[CH_NB / 4], const double* __restrict__ src, int n, int row0, int col0) {
    const int rsub = threadIdx.x & 3, csub = (threadIdx.x >> 2) & 3, prow = threadIdx.x >> 4;
    const int gi = row0 + 4 * prow + rsub;
#pragma unroll
    for (int pc = 0; pc < CH_NB / 4; ++pc) {
        const int gj = col0 + 4 * pc + csub;
        v[pc] = (gi < n && gj < n) ? src[(size_t)gj * n + gi] : 0.0;
    }
}
template <bool TRANS>
__device__ __forceinline__ void ch_store_tile(double* dst, const double (&v)[CH_NB / 4]) {
    const int rsub = threadIdx.x & 3, csub = (threadIdx.x >> 2) & 3, prow = threadIdx.x >> 4;
    const int r = 4 * prow + rsub;
#pragma unroll
    for (int pc = 0; pc < CH_NB / 4; ++pc) {
        const int c = 4 * pc + csub;
        if (TRANS) dst[c * CH_LD + r] = v[pc];
        else dst[r * CH_LD + c] = v[pc];
    }
}
template <bool TRANS>
__device__ __forceinline__ void ch_load_tile(double* dst, const double* __restrict__ src, int n, int row0, int col0) {
    double v[CH_NB / 4];
    ch_fetch_tile(v, src, n, row0, col0);
    ch_store_tile<TRANS>(dst, v);
}

// 64 x 64 x 64 tile product on FP64 tensor-core tiles (mma.sync m8n8k4, DMMA): warp w owns rows 8 w .. 8 w + 7 of the
// tile and all eight 8-column tiles; lane l = 4 g + t holds, per column tile j, the entries (8 w + g, 8 j + 2 t) and
// (8 w + g, 8 j + 2 t + 1).   acc[j][e] += sum_kk As[8 w + g][kk] * B(kk, 8 j + 2 t + e)
// with B(kk, c) = Bs[kk][c] (BT = false) or Bs[c][kk] (BT = true).  DMMA issues at the FP64 FMA rate, but a k-step of
// four needs 9 shared-memory loads per lane for 8 x 8 x 8 x 4 FMAs per warp where the scalar form needed 8 loads per
// 16 FMAs per lane: the scalar tile product was bound by the shared-memory pipe at about twice the FP64-pipe time.
template <bool BT>
__device__ __forceinline__ void ch_tile_mma(const double* As, const double* Bs, double (&acc)[8][2]) {
    const int w = threadIdx.x >> 5, g = (threadIdx.x & 31) >> 2, t = threadIdx.x & 3;
    const double* ap = As + (8 * w + g) * CH_LD + t;
    const double* bp = BT ? Bs + g * CH_LD + t : Bs + t * CH_LD + g;
#pragma unroll 4
    for (int k0 = 0; k0 < CH_NB; k0 += 4) {
        const double a = ap[k0];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double b = BT ? bp[8 * j * CH_LD + k0] : bp[k0 * CH_LD + 8 * j];
            dmma884(acc[j][0], acc[j][1], a, b);
        }
    }
}

// the (row, column) inside the 64 x 64 tile of acc[j][e] in ch_tile_mma
__device__ __forceinline__ int ch_mma_row() { return 8 * (threadIdx.x >> 5) + ((threadIdx.x & 31) >> 2); }
__device__ __forceinline__ int ch_mma_col(int j, int e) { return 8 * j + 2 * (threadIdx.x & 3) + e; }

// rl ~ 1 / sqrt(x), l ~ sqrt(x) for x > 0 (NaN for x <= 0 or NaN, which is what a failed factorisation should spread).
__device__ __forceinline__ void ch_rsqrt(double x, double& rl, double& l) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;                           // g ~ sqrt(x), h ~ 1 / (2 sqrt(x))
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double dd = fma(-g, g, x);                         // residual correction of the root
    g = fma(dd, h, g);
    // x <= 0 or NaN (a failed factorisation, reported through `info`): what sqrt(x) and 1 / sqrt(x) would give, by
    // selects - calling the FP64 sqrt and division here put ~80 instructions on the critical path of EVERY column
    const bool ok = x > 0.0;
    const double bad = x == 0.0 ? 0.0 : __longlong_as_double(0x7ff8000000000000ll);
    l = ok ? g : bad;
    rl = ok ? 2.0 * h : (x == 0.0 ? __longlong_as_double(0x7ff0000000000000ll) : bad);
}

// Columns 16 Q .. 16 Q + 15 of the sweep in chol_panel_kernel (see there).
template <int Q>
__device__ __forceinline__ void ch_sweep16(double (&d)[4][4], double (&e)[4][4], double* colv, double* rowE, double* pivs,
                                           int* info, int k0, bool report) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll 1
    for (int o = 0; o < 16; ++o) {
        const int c = 16 * Q + o;                            // element (c, c) lives in thread (ty = o, tx = o) at [Q][Q]
        __syncthreads();                                     // pivot of column c published; column c - 1 fully consumed
        const double piv = *pivs;
        if (!(piv > 0.0) && threadIdx.x == 0 && report) atomicCAS(info, 0, k0 + c + 1);
        // 1 / sqrt(piv) and sqrt(piv) from the hardware seed and two coupled Newton steps (<= 1 ulp): FP64 sqrt and
        // division are ~40 dependent instructions each and this is the critical path of every column
        double rl, l;
        ch_rsqrt(piv, rl, l);
        if (ty == o) {                                       // owners of column c: scale it and publish it
#pragma unroll
            for (int i = Q; i < 4; ++i) {
                const int r = tx + 16 * i;
                if (r > c) d[i][Q] *= rl;
                else if (r == c) d[i][Q] = l;
                colv[r] = d[i][Q];
            }
        }
        if (tx == o) {                                       // owners of row c of the inverse: it becomes final
#pragma unroll
            for (int j = 0; j <= Q; ++j) {
                const int col = ty + 16 * j;
                if (col <= c) e[Q][j] *= rl;
                rowE[col] = e[Q][j];
            }
        }
        __syncthreads();
        // Rank-1 updates without per-element predicates (comparisons and selects per element were most of a column's
        // instructions): the masks go into the operands - L(r, c) = 0 for r <= c, L(cc, c) = 0 for cc <= c, inverse(c, cc)
        // = 0 for cc > c - and fma(-0, x, y) = y leaves the finished entries as they are.  Entries above the diagonal of
        // a diagonal block take garbage; nothing reads them.
        double lr[4], lc[4], er[4];
#pragma unroll
        for (int i = Q; i < 4; ++i) {                        // row groups above Q are finished
            const int r = tx + 16 * i;
            lr[i] = r > c ? colv[r] : 0.0;
        }
#pragma unroll
        for (int j = Q; j < 4; ++j) {                        // block columns that can still be right of c
            const int cc = ty + 16 * j;
            lc[j] = cc > c ? colv[cc] : 0.0;
        }
#pragma unroll
        for (int j = 0; j <= Q; ++j) {                       // columns of the inverse that can be <= c
            const int cc = ty + 16 * j;
            er[j] = cc <= c ? rowE[cc] : 0.0;
        }
#pragma unroll
        for (int i = Q; i < 4; ++i) {
#pragma unroll
            for (int j = Q; j <= i; ++j) d[i][j] = fma(-lr[i], lc[j], d[i][j]);
#pragma unroll
            for (int j = 0; j <= Q; ++j) e[i][j] = fma(-lr[i], er[j], e[i][j]);
        }
        // publish the next pivot
        if (o < 15) {
            if (tx == o + 1 && ty == o + 1) *pivs = d[Q][Q];
        } else if (Q < 3) {
            if (tx == 0 && ty == 0) *pivs = d[Q < 3 ? Q + 1 : Q][Q < 3 ? Q + 1 : Q];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Cholesky
// ---------------------------------------------------------------------------------------------------------------
// grid.x = number of 64-row blocks from the diagonal block down.  Dinv: [n_blocks][64][64] row-major inverses of the
// diagonal blocks of L.  info: 0, or 1 + the index of the first non-positive pivot (LAPACK's convention).
__global__ void __launch_bounds__(CH_THREADS) chol_panel_kernel(double* __restrict__ A, int n, int k0, double* __restrict__ Dinv,
                                                                 int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    double* D = reinterpret_cast<double*>(ch_smem);          // [64][65] diagonal block -> its factor (lower)
    double* E = D + CH_NB * CH_LD;                           // [64][65] identity -> inverse of the factor (lower)
    double* S = E + CH_NB * CH_LD;                           // [64][65] this CTA's slab of the panel
    const int tid = threadIdx.x;
    {
        double vd[CH_NB / 4], vs[CH_NB / 4];
        ch_fetch_tile(vd, A, n, k0, k0);
        if (blockIdx.x > 0) ch_fetch_tile(vs, A, n, k0 + CH_NB * blockIdx.x, k0);
        ch_store_tile<false>(D, vd);
        if (blockIdx.x > 0) ch_store_tile<false>(S, vs);
    }
    __syncthreads();
    // rows past the end of the matrix: identity, so that the factor and its inverse stay defined
    if (tid < CH_NB && k0 + tid >= n) {
        for (int c = 0; c < CH_NB; ++c) D[tid * CH_LD + c] = c == tid ? 1.0 : 0.0;
    }
    __syncthreads();
    // Right-looking sweep over the 64 columns of the block with the inverse carried along (forward substitution on the
    // identity, also right-looking, so that no update is a dependent chain).  Thread (ty, tx) of a 16 x 16 grid keeps
    // its 4 x 4 share of the block and of the inverse in REGISTERS for the whole sweep - element (tx + 16 i, ty + 16 j) -
    // and only the current column / row travel through shared memory: two barriers and ~100 instructions per column
    // (ch_sweep16, one instantiation per 16 columns so that register indices stay compile-time constants without
    // unrolling all 64 columns, which would make the kernel instruction-fetch bound).
    const int tx = tid & 15, ty = tid >> 4;
    __shared__ double colv[CH_NB], rowE[CH_NB], pivs;
    double d[4][4], e[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            d[i][j] = D[(tx + 16 * i) * CH_LD + ty + 16 * j];
            e[i][j] = (tx + 16 * i == ty + 16 * j) ? 1.0 : 0.0;
        }
    if (tid == 0) pivs = d[0][0];
    const bool report = blockIdx.x == 0;
    ch_sweep16<0>(d, e, colv, rowE, &pivs, info, k0, report);
    ch_sweep16<1>(d, e, colv, rowE, &pivs, info, k0, report);
    ch_sweep16<2>(d, e, colv, rowE, &pivs, info, k0, report);
    ch_sweep16<3>(d, e, colv, rowE, &pivs, info, k0, report);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            D[(tx + 16 * i) * CH_LD + ty + 16 * j] = d[i][j];
            E[(tx + 16 * i) * CH_LD + ty + 16 * j] = e[i][j];
        }
    __syncthreads();
    if (blockIdx.x == 0) {
        // the diagonal CTA publishes the factor (lower triangle of the block) and the inverse
        double* dinv = Dinv + (size_t)(k0 / CH_NB) * CH_NB * CH_NB;
        for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
            const int c = idx >> 6, r = idx & 63;                              // r fastest: contiguous in a column
            if (r >= c && k0 + r < n) A[(size_t)(k0 + c) * n + k0 + r] = D[r * CH_LD + c];
            dinv[(idx >> 6) * CH_NB + (idx & 63)] = E[(idx >> 6) * CH_LD + (idx & 63)];
        }
        return;
    }
    // slab: L_slab = A_slab . inv(L_kk)^T, i.e. out(r, c) = sum_j S[r][j] E[c][j]  (E lower: j <= c)
    double acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
    ch_tile_mma<true>(S, E, acc);
    const int row0 = k0 + CH_NB * blockIdx.x;
    const int gi = row0 + ch_mma_row();
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int gj = k0 + ch_mma_col(j, e);
            if (gi < n && gj < n) A[(size_t)gj * n + gi] = acc[j][e];
        }
}

// Trailing update after block column k0: tile (bi, bj), bi >= bj, of the blocks below / right of the panel.
// grid.x enumerates the lower-triangular tile pairs.
__global__ void __launch_bounds__(CH_THREADS) chol_update_kernel(double* __restrict__ A, int n, int k0) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    double* As = reinterpret_cast<double*>(ch_smem);
    double* Bs = As + CH_NB * CH_LD;
    // blockIdx.x -> (bi, bj), bi >= bj >= 0 (triangular enumeration)
    int bi = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= (int)blockIdx.x) ++bi;
    while (bi * (bi + 1) / 2 > (int)blockIdx.x) --bi;
    const int bj = blockIdx.x - bi * (bi + 1) / 2;
    const int i0 = k0 + CH_NB * (bi + 1), j0 = k0 + CH_NB * (bj + 1);
    {
        double va[CH_NB / 4], vb[CH_NB / 4];
        ch_fetch_tile(va, A, n, i0, k0);
        ch_fetch_tile(vb, A, n, j0, k0);
        ch_store_tile<false>(As, va);                         // As[r][kk] = L(i0 + r, k0 + kk)
        ch_store_tile<true>(Bs, vb);                          // Bs[kk][c] = L(j0 + c, k0 + kk)
    }
    __syncthreads();
    double acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
    ch_tile_mma<false>(As, Bs, acc);
    const int gi = i0 + ch_mma_row();
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int gj = j0 + ch_mma_col(j, e);
            if (gi < n && gj < n && gi >= gj) A[(size_t)gj * n + gi] -= acc[j][e];
        }
}

// ---------------------------------------------------------------------------------------------------------------
// Triangular inverse by recursive doubling
// ---------------------------------------------------------------------------------------------------------------
// X := 0 with the inverted diagonal blocks on the diagonal.
__global__ void tri_inverse_seed_kernel(const double* __restrict__ Dinv, double* __restrict__ X, int n) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int j = (int)(idx / n), i = (int)(idx % n);
    double v = 0.0;
    if (i >= j && i / CH_NB == j / CH_NB) v = Dinv[(size_t)(i / CH_NB) * CH_NB * CH_NB + (i % CH_NB) * CH_NB + (j % CH_NB)];
    X[idx] = v;
}

// One level: the matrix is cut into diagonal blocks of 2h rows; in each, X11 (h x h) and X22 are already inverted and
//   PHASE 0:  W21 = L21 . X11          (k runs over the first half from the column block down: X11 is lower triangular)
//   PHASE 1:  X21 = -X22 . W21         (k runs over the second half up to the row block: X22 is lower triangular)
// grid = (tiles per h x h block = (h / 64)^2, number of 2h blocks).
template <int PHASE>
__global__ void __launch_bounds__(CH_THREADS) tri_inverse_level_kernel(const double* __restrict__ L, double* __restrict__ X,
                                                                        double* __restrict__ W, int n, int h) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    double* As = reinterpret_cast<double*>(ch_smem);
    double* Bs = As + CH_NB * CH_LD;
    const int hb = h / CH_NB;
    const int ti = blockIdx.x / hb, tj = blockIdx.x % hb;
    const int base = 2 * h * blockIdx.y;                      // first row / column of this 2h block
    const int i0 = base + h + CH_NB * ti, j0 = base + CH_NB * tj;
    if (i0 >= n) return;                                      // the second half does not exist (uniform per CTA)
    double acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
    const int kb_lo = PHASE == 0 ? tj : 0, kb_hi = PHASE == 0 ? hb - 1 : ti;
    const double* srcA = PHASE == 0 ? L : X;                  // L21(i, k) | X22(i, k)
    const double* srcB = PHASE == 0 ? X : W;                  // X11(k, j) | W21(k, j)
    const int kbase = PHASE == 0 ? base : base + h;
    double va[CH_NB / 4], vb[CH_NB / 4];
    ch_fetch_tile(va, srcA, n, i0, kbase + CH_NB * kb_lo);
    ch_fetch_tile(vb, srcB, n, kbase + CH_NB * kb_lo, j0);
    for (int kb = kb_lo; kb <= kb_hi; ++kb) {
        __syncthreads();
        ch_store_tile<false>(As, va);
        ch_store_tile<false>(Bs, vb);
        __syncthreads();
        if (kb < kb_hi) {                                     // the next pair travels while this one is multiplied
            ch_fetch_tile(va, srcA, n, i0, kbase + CH_NB * (kb + 1));
            ch_fetch_tile(vb, srcB, n, kbase + CH_NB * (kb + 1), j0);
        }
        ch_tile_mma<false>(As, Bs, acc);
    }
    double* out = PHASE == 0 ? W : X;
    const int gi = i0 + ch_mma_row();
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int gj = j0 + ch_mma_col(j, e);
            if (gi < n && gj < n) out[(size_t)gj * n + gi] = PHASE == 0 ? acc[j][e] : -acc[j][e];
        }
}

// out(i, j) = sum_{k >= i} X(k, i) X(k, j) for i >= j (lower triangle of X^T X); grid.x enumerates the lower tile pairs.
__global__ void __launch_bounds__(CH_THREADS) tri_gram_kernel(const double* __restrict__ X, double* __restrict__ out, int n) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    double* As = reinterpret_cast<double*>(ch_smem);
    double* Bs = As + CH_NB * CH_LD;
    int bi = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= (int)blockIdx.x) ++bi;
    while (bi * (bi + 1) / 2 > (int)blockIdx.x) --bi;
    const int bj = blockIdx.x - bi * (bi + 1) / 2;
    const int i0 = CH_NB * bi, j0 = CH_NB * bj;
    const int nb = (n + CH_NB - 1) / CH_NB;
    double acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
    double va[CH_NB / 4], vb[CH_NB / 4];
    ch_fetch_tile(va, X, n, CH_NB * bi, i0);
    ch_fetch_tile(vb, X, n, CH_NB * bi, j0);
    for (int kb = bi; kb < nb; ++kb) {
        __syncthreads();
        ch_store_tile<true>(As, va);                          // As[r][kk] = X(k0 + kk, i0 + r)
        ch_store_tile<false>(Bs, vb);                         // Bs[kk][c] = X(k0 + kk, j0 + c)
        __syncthreads();
        if (kb + 1 < nb) {                                    // the next pair travels while this one is multiplied
            ch_fetch_tile(va, X, n, CH_NB * (kb + 1), i0);
            ch_fetch_tile(vb, X, n, CH_NB * (kb + 1), j0);
        }
        ch_tile_mma<false>(As, Bs, acc);
    }
    const int gi = i0 + ch_mma_row();
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int gj = j0 + ch_mma_col(j, e);
            if (gi < n && gj < n && gi >= gj) out[(size_t)gj * n + gi] = acc[j][e];
        }
}

// z = X y for lower-triangular X: z_i = sum_{j <= i} X(i, j) y_j.  grid = (64-row blocks, chunks of CH_MV_COLS columns);
// CTA (b, ch) writes the partial sum of its columns to part[ch][i]; tri_matvec_reduce_kernel adds the chunks in order.
constexpr int CH_MV_COLS = 256;
__global__ void __launch_bounds__(CH_THREADS) tri_matvec_n_kernel(const double* __restrict__ X, const double* __restrict__ y,
                                                                   double* __restrict__ part, int n) {
    __shared__ double red[4][CH_NB];
    const int r = threadIdx.x & 63, g = threadIdx.x >> 6;
    const int i = blockIdx.x * CH_NB + r;
    const int j0 = blockIdx.y * CH_MV_COLS;
    if (j0 > blockIdx.x * CH_NB + CH_NB - 1) return;         // this chunk lies right of the diagonal for all 64 rows
    double s = 0.0;
    if (i < n) {
        const int jend = min(i, j0 + CH_MV_COLS - 1);
        double acc[4] = {0.0, 0.0, 0.0, 0.0};                // 4 loads in flight per thread; fixed combination order
        int j = j0 + g;
        for (; j + 12 <= jend; j += 16)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fma(X[(size_t)(j + 4 * u) * n + i], y[j + 4 * u], acc[u]);
        for (int u = 0; j <= jend; j += 4, ++u) acc[u & 3] = fma(X[(size_t)j * n + i], y[j], acc[u & 3]);
        s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    }
    red[g][r] = s;
    __syncthreads();
    if (g == 0 && i < n) part[(size_t)blockIdx.y * n + i] = ((red[0][r] + red[1][r]) + red[2][r]) + red[3][r];
}

__global__ void tri_matvec_reduce_kernel(const double* __restrict__ part, double* __restrict__ z, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int chunks = i / CH_MV_COLS + 1;                   // chunks left of (or on) the diagonal of row i
    double s = 0.0;
    for (int ch = 0; ch < chunks; ++ch) s += part[(size_t)ch * n + i];
    z[i] = s;
}

// a = X^T z: a_j = sum_{i >= j} X(i, j) z_i.  One warp per column (contiguous), fixed xor tree.
__global__ void __launch_bounds__(CH_THREADS) tri_matvec_t_kernel(const double* __restrict__ X, const double* __restrict__ z,
                                                                   double* __restrict__ a, int n) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (CH_THREADS / 32) + (threadIdx.x >> 5);
    if (j >= n) return;
    double s = 0.0;
    for (int i = j + lane; i < n; i += 32) s = fma(X[(size_t)j * n + i], z[i], s);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) a[j] = s;
}

}  // namespace skb
