// Tile sizes and packed layouts shared by the pack, cross-kernel (K1) and triangular-contraction (K2) kernels.
//
// "k4-interleaved" layout of a (rows x K) operand: element (row, k) lives at ((k/4) * rows + row) * 4 + (k % 4).
// With it every DMMA m8n8k4 fragment (8 rows x 4 k-values) is 32 consecutive doubles in lane order, so a warp
// fetches a fragment with one conflict-free 256-byte shared-memory read, and every pipeline stage of either
// operand is ONE contiguous block in global memory (a single 1-D bulk copy, no tensor map needed).
#pragma once
#include <cstdint>

namespace skb {

constexpr int BM = 128;   // rows of L_inv per CTA row block
constexpr int BN = 128;   // candidates per CTA tile
constexpr int BK = 32;    // training points (contraction depth) per pipeline stage
constexpr int TRMM_STAGES = 3;
constexpr int TRMM_THREADS = 288;          // 8 DMMA warps + 1 bulk-copy producer warp

constexpr int XROWS = 64;                  // training rows per K1 stage
constexpr int XG = XROWS / 4;              // groups of 4 rows per K1 stage
constexpr int XS_LD = BN + 1;              // padded leading dimension of the candidate tile in shared memory
// Largest n_dims: the candidate tile and two stages of training rows stay resident in shared memory per CTA
// (K1 208.7 KB, K1 DMMA variant 214.5 KB, K3g 228.7 KB at d_pad = 100, of 227 KB = 232,448 B); beyond 52 dimensions
// one K1 CTA per SM instead of two.  K3g takes the gradient columns beyond 64 in a second launch (grad.cuh).
constexpr int MAX_DIMS = 100;

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// number of (BM x BK) tiles of packed L_inv that precede row block I: sum_{i<I} (i+1)*BM/BK
__host__ __device__ inline int64_t vp_tiles_before(int I) { return (int64_t)(BM / BK) * I * (I + 1) / 2; }

}  // namespace skb
