// Split-integer ("Ozaki scheme") evaluation of the variance contraction on the 5th-generation tensor cores.
//
// tcgen05.mma has no FP64 kind, and the warp-level FP64 MMA (DMMA) runs at the FP64 FMA rate (37 TFLOP/s measured),
// which is where trmm.cuh sits.  The 8-bit integer kind runs ~60x faster and accumulates EXACTLY in int32, so the FP64
// contraction U = L_inv . K_trans^T is re-expressed in fixed point:
//
//   L_inv[i, j]  = sV_i * 2^-55 * sum_p a_p[i, j] 256^(6-p)      a_p in [-128, 127], sV_i a power of two >= 2 max_j |L_inv[i, j]|
//   k[j, c]      = sK   * 2^-55 * sum_q b_q[j, c] 256^(6-q)      b_q in [-128, 127], sK   a power of two >= 2 (|amplitude| + |bias|)
//   U[i, c]      = sV_i sK 2^-14 * sum_t 256^-t G_t[i, c],       G_t = sum_{p+q=t} sum_j a_p[i, j] b_q[j, c]   (exact int32)
//
// with the digit pairs p + q <= 6 kept (28 of 49 products; the dropped ones are below 2^-53 of full scale, i.e. the
// size of an FP64 rounding of one product).  Both operands are quantised to 2^-55 of their row / global scale, the
// integer sums are exact, and the seven partial sums are combined in FP64 -- so the result carries a few units of
// 1e-16 * (row scale) * (kernel scale) per term instead of FP64's 1e-16 * |term|: still far inside the 1e-9 parity gate
// (checked in tests/test_gpu_split.py), while the contraction runs on tcgen05 at several times the DMMA rate.
//
// One digit plane p of L_inv is multiplied with ALL admissible planes q of the candidates at once: the B operand rows
// are ordered [plane q][candidate] (64 candidates per CTA), the seven accumulator regions G_0 .. G_6 sit side by side
// in TMEM (64 columns each), so plane p accumulates into columns [64 p, 448) -- a contiguous range, issued as one or
// two MMAs of N <= 256 (10 instructions per 32 training points instead of 28).  The CTA walks the lower-triangular
// row blocks of L_inv like trmm.cuh; operand tiles arrive through a 4-stage mbarrier ring of 1-D bulk copies (TMA), one thread issues the
// MMAs, eight warps drain TMEM (tcgen05.ld), rebuild U in FP64, square and reduce it.
#pragma once
#include <math.h>
#include "layout.cuh"
#include "ptx.cuh"
#include "trmm.cuh"

namespace skb {

constexpr int OZ_P = 7;                       // digit planes per operand
constexpr int OZ_KC = 32;                     // training points per MMA (UMMA K for 8-bit operands)
constexpr int OZ_SUB = 32;                    // candidates per accumulator group
constexpr int OZ_NC = 64;                     // candidates per CTA (two groups)
constexpr int OZ_EPI_WARPS = 8;                // two warps per TMEM lane quarter, each owning 32 of the 64 candidates
constexpr int OZ_THREADS = 64 + 32 * OZ_EPI_WARPS;   // warp 0: bulk-copy producer, warp 1: MMA issuer, warps 2-9: epilogue
constexpr int OZ_A_CHUNK = OZ_P * BM * OZ_KC;                      // 28,672 B: [plane][k half][row group 16][8 rows][16 B]
constexpr int OZ_TMEM_COLS = 512;
constexpr unsigned long long OZ_BIAS = 0x0080808080808080ull;      // +128 in each of the 7 digit bytes
// The contraction can run on the leading P <= OZ_P planes of both operands.  Balanced digits make that a rounding, not
// a truncation: the first P digits of the 7-digit expansion ARE the P-digit expansion of the value rounded to
// 2^-(8P-1) of its scale (every dropped digit is at most half a unit of the one before it), so the packed L_inv planes
// serve every P and only the leading planes are fetched.  P = 6 (47-bit operands, 21 digit-plane products instead of
// 28): difference to the exact contraction 4e-12 of the prior variance at (2048, 20) - the 1e-9 gate with a factor
// of 200 to spare - against 2e-14 for P = 7 (oracle/split_emulation.py, tests/test_split_emulation.py).
__host__ __device__ constexpr int oz_a_bytes(int P) { return P * BM * OZ_KC; }                      // leading planes of one A chunk
__host__ __device__ constexpr int oz_b_half(int P, int nc = OZ_NC) { return P * (nc / 8) * 128; }                // [plane][cand group 8][8 cands][16 B] per k half
__host__ __device__ constexpr int oz_b_chunk(int P, int nc = OZ_NC) { return 2 * oz_b_half(P, nc); }                    // 14,336 B (P = 7), 12,288 B (P = 6)
__host__ __device__ constexpr int oz_stage_bytes(int P, int nc = OZ_NC) { return oz_a_bytes(P) + oz_b_chunk(P, nc); }   // 43,008 B / 36,864 B
// Build-time knobs for the co-residency experiment (profiles/r02_k2o_diagnosis.md): -DOZ_STAGES6=4 -DOZ_MAXNREG=152 leave
// room on the SM for one 64-candidate K1 CTA (SKB_SPLIT_OVERLAP=1); measured equal in speed to the defaults.
#ifndef OZ_STAGES6
#define OZ_STAGES6 5
#endif
// 168 is the ceiling for this CTA shape: its 10 warps land 3 + 3 + 2 + 2 on the SM's four sub-partitions of 16,384 registers
// each, and 3 warps x 32 lanes x 176 registers no longer fit one of them ("too many resources requested for launch").
#ifndef OZ_MAXNREG
#define OZ_MAXNREG 168
#endif
__host__ __device__ constexpr int oz_stages(int P, int nc = OZ_NC) { return P >= 7 ? 4 : OZ_STAGES6; }

// byte offset of (row, k) inside one canonical K-major, no-swizzle (rows x 32) int8 block: 8-row x 16-byte core matrices
__host__ __device__ inline int oz_canon(int rows, int row, int k) {
    return (k >> 4) * (rows >> 3) * 128 + (row >> 3) * 128 + (row & 7) * 16 + (k & 15);
}

// x (|x| <= 0.5 after scaling) -> 7 balanced base-256 digits in the low 7 bytes (byte 6 = most significant digit)
__device__ __forceinline__ unsigned long long oz_digits(double scaled_2p55) {
    const long long v = __double2ll_rn(scaled_2p55);
    return ((unsigned long long)v + OZ_BIAS) ^ OZ_BIAS;
}

__host__ __device__ inline double oz_pow2_scale(double maxabs) {      // power of two s with maxabs / s in [0.25, 0.5)
    if (!(maxabs > 0.0)) return 1.0;
    int e;
    frexp(maxabs, &e);                     // maxabs = f * 2^e, f in [0.5, 1)
    return ldexp(1.0, e + 1);
}

// ---------------------------------------------------------------------------------------------------------------
// state preparation: row scales and digit planes of L_inv (lower-triangular 128 x 32 chunks, as in pack_linv_kernel)
// ---------------------------------------------------------------------------------------------------------------
// DENSE = false: lower-triangular L_inv (entries above the diagonal ignored); DENSE = true: full symmetric K_inv
template <bool DENSE>
__global__ void oz_row_scale_kernel(const double* __restrict__ V, double* __restrict__ sV, int n, int n_pad) {
    const int row = blockIdx.x;
    __shared__ double red[256];
    double m = 0.0;
    if (row < n)
        for (int j = threadIdx.x; j <= (DENSE ? n - 1 : row); j += blockDim.x) m = fmax(m, fabs(V[(size_t)row * n + j]));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) sV[row] = oz_pow2_scale(red[0]);
}

template <bool DENSE>
__global__ void oz_pack_linv_kernel(const double* __restrict__ V, const double* __restrict__ sV, int8_t* __restrict__ Vd,
                                    int n, int n_pad) {
    const int I = blockIdx.y, kc = blockIdx.x;
    if (!DENSE && kc >= (I + 1) * (BM / OZ_KC)) return;
    int8_t* chunk = Vd + (size_t)(DENSE ? (int64_t)I * (n_pad / OZ_KC) + kc : vp_tiles_before(I) + kc) * OZ_A_CHUNK;
    for (int e = threadIdx.x; e < BM * OZ_KC; e += blockDim.x) {
        const int i = e / OZ_KC, k = e % OZ_KC;
        const int row = I * BM + i, col = kc * OZ_KC + k;
        double v = 0.0;
        if (row < n && (DENSE ? col < n : col <= row)) v = V[(size_t)row * n + col] / sV[row] * 36028797018963968.0;   // * 2^55 (exact)
        const unsigned long long y = oz_digits(v);
        const int off = oz_canon(BM, i, k);
#pragma unroll
        for (int p = 0; p < OZ_P; ++p) chunk[p * (BM * OZ_KC) + off] = (int8_t)((y >> (8 * (6 - p))) & 0xFF);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K2o: the contraction
// ---------------------------------------------------------------------------------------------------------------
struct OzakiParams {
    const int8_t* Vd;         // digit planes of L_inv, chunk (I, kc) at (vp_tiles_before(I) + kc) * OZ_A_CHUNK
    const double* sV;         // [n_pad] row scales
    const int8_t* Kd;         // digit planes of K_trans: [tile64][kc][oz_b_chunk(P)]
    const double* kdot;       // [tiles*64] K_trans . alpha from K1
    double k_scale;           // sK
    int n, n_pad;
    int64_t m;
    AcqConsts a;
    double* mean;
    double* std;
    double* acq[3];
    unsigned long long* clamp_count;
    double* Wt;               // MODE 1: [tiles128][n_pad][BN] = K_inv . K_trans^T in FP64 for the gradient kernel
    long long* dbg_cycles;    // diagnostics (skb_debug_k2o): SM cycles each CTA spent between its setup and its teardown
};

constexpr size_t oz_smem_bytes(int P, int nc = OZ_NC) {
    return (size_t)oz_stages(P, nc) * oz_stage_bytes(P, nc) + 4 * nc * sizeof(double) + 256;
}

__device__ __forceinline__ uint64_t oz_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);      // version 1, SWIZZLE_NONE
}
// kind::i8, S8 x S8 -> S32, K-major operands, M = 128
__device__ __forceinline__ uint32_t oz_idesc(int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void oz_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// issue (no wait) a load of 8 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void oz_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void oz_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// exact (double)(int32) on the FP64 pipe: the mantissa of 2^52 + 2^31 + x is x ^ 0x80000000, one subtraction removes the
// offset.  The I2F conversion unit runs at a quarter of the FP64 rate and was the epilogue's bottleneck.
__device__ __forceinline__ double oz_i2d(uint32_t g) {
    return __hiloint2double(0x43300000, (int)(g ^ 0x80000000u)) - 4503601774854144.0;
}

// Column sums over the 32 lanes of a warp for 32 columns held as v[0..31] in every lane (transposing butterfly,
// 16 + 8 + 4 + 2 + 1 exchanges); afterwards lane l holds the total of column l.  Fixed order -> deterministic.
__device__ __forceinline__ double oz_colsum32(double (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int c = 0; c < o; ++c) {
            const double send = upper ? v[c] : v[c + o];
            const double keep = upper ? v[c + o] : v[c];
            v[c] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

// MODE 0: triangular L_inv, U squared and reduced, posterior + acquisition epilogue; grid.x = 64-candidate tile.
// MODE 1: dense K_inv, W = K_inv . K_trans^T stored in FP64; grid = (row block, 64-candidate tile).
// DBG (diagnostics only, skb_debug_k2o / tools/k2o_diag.py; 0 in every product launch): bit 0 = the epilogue warps skip
// the TMEM reads and the FP64 recombination, bit 1 = the producer signals the stages full without copying anything,
// bit 2 = no accumulator hand-over between row blocks (one at the very end), bit 4 = no stage barriers at all (the MMAs
// read whatever the stages hold), bit 5 = operands are copied once per stage and re-used (real data, no L2 -> SM traffic);
// any bit: the CTA's SM cycle count goes to p.dbg_cycles.
// CL > 1 (MODE 0 only): the kernel is launched in clusters of CL CTAs along x.  All CTAs of a cluster walk the same
// (row block, chunk) sequence for different candidate tiles, so every L_inv chunk is fetched from L2 ONCE per cluster: CTA r
// copies slice r of it and multicasts the slice into the same stage of all CL CTAs; a stage is free again when the MMAs of
// ALL CL CTAs have read it (tcgen05.commit multicast onto every CTA's `empty` barrier, which counts CL arrivals).  The
// L2 -> SM operand traffic is what holds this kernel's clock down under the board's power cap (tools/k2o_power.py: the
// same MMAs on operands that stay in shared memory run 18 % faster); the L_inv planes are two thirds of it.
// NC = candidates per CTA: 64, or 80 (P = 6 only: 6 x 80 = 480 of the 512 TMEM columns).  The wider tile delivers 13 % fewer
// operand bytes per MAC into shared memory (what the power cap charges for, tools/k2o_power.py) and its instruction widths
// (240+240 | 208+192 | 160+160 | 240 | 160 | 80) run at 0.995 of the issue rate instead of 0.982 (umma_i8_shapes.cu).
template <int MODE, int P, int DBG = 0, int CL = 1, int NC = OZ_NC>
__global__ void __maxnreg__(OZ_MAXNREG) ozaki_trmm_kernel(const OzakiParams p) {
    static_assert(NC % 16 == 0 && NC * P <= OZ_TMEM_COLS, "accumulator regions must fit TMEM");
    static_assert(MODE == 0 || NC == OZ_NC, "the dense mode keeps 64-candidate tiles");
    constexpr int OZ_STAGES = oz_stages(P, NC);
    constexpr int OZ_STAGE_BYTES = oz_stage_bytes(P, NC);
    constexpr int OZ_B_HALF = oz_b_half(P, NC);
    constexpr int OZ_B_CHUNK = oz_b_chunk(P, NC);
    constexpr int A_BYTES = oz_a_bytes(P);
    extern __shared__ __align__(1024) unsigned char smem_oz[];
    unsigned char* stages = smem_oz;                                          // [OZ_STAGES][A chunk | B chunk]
    double* sRed = reinterpret_cast<double*>(stages + (size_t)OZ_STAGES * OZ_STAGE_BYTES);   // [4 row quarters][NC]
    uint64_t* full = reinterpret_cast<uint64_t*>(sRed + 4 * NC);
    uint64_t* empty = full + OZ_STAGES;
    uint64_t* tmem_full = empty + OZ_STAGES;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, CL);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, OZ_EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (CL > 1) cluster_sync_all();             // every CTA's barriers exist before a peer's copy or commit can reach them
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    const long long dbg_t0 = DBG ? clock64() : 0;
    const uint32_t cl_rank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1u);

    // a cluster launch rounds the grid up: surplus CTAs re-read the last tile's planes and store nothing (cand >= m)
    const int last_tile = MODE == 0 ? (int)((p.m + NC - 1) / NC) - 1 : 0;
    const int tile_id = MODE == 0 ? (CL > 1 ? min((int)blockIdx.x, last_tile) : (int)blockIdx.x) : blockIdx.y;
    const int I_begin = MODE == 0 ? 0 : blockIdx.x;
    const int I_end = MODE == 0 ? p.n_pad / BM : blockIdx.x + 1;
    const int nkc_used = (p.n + OZ_KC - 1) / OZ_KC;
    const int nkc_all = p.n_pad / OZ_KC;
    const int8_t* kd_tile = p.Kd + (size_t)tile_id * nkc_all * OZ_B_CHUNK;

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0 && !(DBG & 16)) {
            uint32_t it = 0;
            for (int I = I_begin; I < I_end; ++I) {
                const int nkc = MODE == 0 ? min((I + 1) * (BM / OZ_KC), nkc_used) : nkc_used;
                const int8_t* vrow = p.Vd + (size_t)(MODE == 0 ? vp_tiles_before(I) : (int64_t)I * nkc_all) * OZ_A_CHUNK;
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int s = it % OZ_STAGES;
                    const uint32_t ph = (it / OZ_STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    if ((DBG & 2) || ((DBG & 32) && it >= (uint32_t)OZ_STAGES)) {   // 32: real operands once per stage, then stale
                        mbar_arrive(full + s);
                        continue;
                    }
                    mbar_arrive_expect_tx(full + s, OZ_STAGE_BYTES);
                    unsigned char* dst = stages + (size_t)s * OZ_STAGE_BYTES;
                    if (CL == 1) {
                        bulk_g2s(dst, vrow + (size_t)kc * OZ_A_CHUNK, A_BYTES, full + s);    // leading P of the 7 packed planes
                    } else {
                        constexpr int SLICE = A_BYTES / CL;                                   // my share of the chunk, to everybody
                        bulk_g2s_multicast(dst + cl_rank * SLICE, vrow + (size_t)kc * OZ_A_CHUNK + cl_rank * SLICE, SLICE, full + s,
                                           CL_MASK);
                    }
                    bulk_g2s(dst + A_BYTES, kd_tile + (size_t)kc * OZ_B_CHUNK, OZ_B_CHUNK, full + s);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            uint32_t it = 0;
            for (int I = I_begin; I < I_end; ++I) {
                const int nkc = MODE == 0 ? min((I + 1) * (BM / OZ_KC), nkc_used) : nkc_used;
                if (I > I_begin && !(DBG & 4)) {
                    mbar_wait(tmem_empty, (uint32_t)((I - I_begin - 1) & 1));   // epilogue of the previous row block drained TMEM
                    asm volatile("tcgen05.fence::after_thread_sync;");
                }
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int s = it % OZ_STAGES;
                    const uint32_t ph = (it / OZ_STAGES) & 1;
                    if (!(DBG & 16)) mbar_wait(full + s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t a_base = smem_u32(stages + (size_t)s * OZ_STAGE_BYTES);
                    const uint32_t b_base = a_base + A_BYTES;
                    const uint32_t acc = kc > 0 ? 1u : 0u;
#pragma unroll
                    for (int pl = 0; pl < P; ++pl) {
                        const uint64_t da = oz_smem_desc(a_base + pl * (BM * OZ_KC), (BM / 8) * 128, 128);
                        // B rows [0, 64 (P - pl)) = planes 0 .. P-1-pl of all 64 candidates -> accumulator columns
                        // [64 pl, 64 P); split so that every instruction has N <= 256 (and N a multiple of 16).
                        const int rows = NC * (P - pl);
                        const int n0 = rows <= 256 ? rows : (NC == OZ_NC ? (pl == 0 ? 256 : rows / 2) : (rows / 2 + 15) / 16 * 16);
#pragma unroll
                        for (int piece = 0; piece < 2; ++piece) {
                            const int r0 = piece == 0 ? 0 : n0;
                            const int N = piece == 0 ? n0 : rows - n0;
                            if (N <= 0) continue;
                            const uint64_t db = oz_smem_desc(b_base + (r0 >> 3) * 128, OZ_B_HALF, 128);
                            // plane 0 touches every column first: it clears the accumulators on the first chunk
                            oz_mma(tmem + NC * pl + r0, da, db, oz_idesc(N), (acc | (pl > 0 ? 1u : 0u)));
                        }
                    }
                    if (!(DBG & 16)) {                      // smem slot is free once these MMAs (of every CTA of the cluster) have read it
                        if (CL == 1) oz_commit(empty + s);
                        else oz_commit_multicast(empty + s, CL_MASK);
                    }
                }
                if (!(DBG & 4) || I == I_end - 1) oz_commit(tmem_full);   // all MMAs of row block I retired -> epilogue may read
            }
        }
    } else {
        // ---------------- epilogue: 8 warps; warp w reads TMEM lanes [32 (w % 4), +32) = rows of the row block, and the
        // candidate half (w - 2) / 4.  Each thread keeps the running sum of squares of ITS row for its 32 candidates
        // (no cross-lane traffic per row block); lanes and quarters are combined once at the end. ----------------
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const double cK = p.k_scale * 6.103515625e-05;       // sK * 2^-14
        double qacc[NC / 2];
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) qacc[c] = 0.0;
        for (int I = (DBG & 4) ? I_end - 1 : I_begin; I < I_end; ++I) {
            mbar_wait(tmem_full, (DBG & 4) ? 0u : (uint32_t)((I - I_begin) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;");
            const double rs = p.sV[I * BM + quarter * 32 + lane] * cK;
#pragma unroll
            for (int jj = 0; jj < ((DBG & 1) ? 0 : NC / 16); ++jj) {
                const int j = half * (NC / 16) + jj;
                // all P regions of these 8 candidates are requested before the single wait: one TMEM round trip
                uint32_t g[P][8];
#pragma unroll
                for (int t = 0; t < P; ++t) oz_ld8_issue(tmem + ((uint32_t)(quarter * 32) << 16) + NC * t + 8 * j, g[t]);
                oz_ld_wait();
                double u[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    double acc = oz_i2d(g[P - 1][c]);
#pragma unroll
                    for (int t = P - 2; t >= 0; --t) acc = fma(acc, 0.00390625, oz_i2d(g[t][c]));   // Horner in 256^-1
                    u[c] = acc * rs;
                    if (MODE == 0) qacc[8 * jj + c] = fma(u[c], u[c], qacc[8 * jj + c]);
                }
                if (MODE == 1) {
                    // row of W owned by this lane, 8 consecutive candidates: one 64-byte store
                    double* w = p.Wt + ((size_t)(tile_id >> 1) * p.n_pad + I * BM + quarter * 32 + lane) * BN +
                                (tile_id & 1) * NC + 8 * j;
                    *reinterpret_cast<double4*>(w) = make_double4(u[0], u[1], u[2], u[3]);
                    *reinterpret_cast<double4*>(w + 4) = make_double4(u[4], u[5], u[6], u[7]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
        }
        if (MODE == 0) {
            if (NC / 2 > 32) {
                // columns 32 .. NC/2 - 1 first (xor tree over the lanes, the same fixed order for every column) ...
#pragma unroll
                for (int c = 32; c < NC / 2; ++c) {
                    double v = qacc[c];
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == c - 32) sRed[quarter * NC + half * (NC / 2) + c] = v;
                }
            }
            // ... then the transposing butterfly over the first 32 (it overwrites its argument)
            double (&q32)[32] = *reinterpret_cast<double (*)[32]>(&qacc[0]);
            sRed[quarter * NC + half * (NC / 2) + lane] = oz_colsum32(q32, lane);
        }
        named_bar_sync(1, 32 * OZ_EPI_WARPS);
        const int t = threadIdx.x - 64;                       // 0 .. 255 over the epilogue warps
        if (MODE == 0 && t < NC) {
            const int64_t cand = (int64_t)blockIdx.x * NC + t;
            if (cand < p.m) {
                const double q = ((sRed[t] + sRed[NC + t]) + sRed[2 * NC + t]) + sRed[3 * NC + t];
                double mu, sd, acq[3];
                if (posterior_from_q(p.a, p.kdot[cand], q, mu, sd) && p.clamp_count) atomicAdd(p.clamp_count, 1ull);
                acquisition_values(p.a, mu, sd, acq);
                if (p.mean) p.mean[cand] = mu;
                if (p.std) p.std[cand] = sd;
#pragma unroll
                for (int a = 0; a < 3; ++a)
                    if (((p.a.acq_mask >> a) & 1) && p.acq[a]) p.acq[a][cand] = acq[a];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (CL > 1) cluster_sync_all();             // nobody leaves while a peer's commit may still be on its way to this CTA
    if (DBG && threadIdx.x == 0 && p.dbg_cycles) p.dbg_cycles[blockIdx.x] = clock64() - dbg_t0;
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

}  // namespace skb
