"""NumPy emulation of the DMMA tile scheme and the trimmed transform of csrc/kcross.cuh: kcross_digits_mma_kernel.
TEST INFRASTRUCTURE ONLY.

The kernel computes the x.t part of the scaled squared distances on FP64 tensor-core tiles (mma.sync m8n8k4) and
stores 6 signed 8-bit digit planes per kernel value in the K-major layout the contraction kernel (csrc/ozaki.cuh)
reads.  This module restates, lane by lane, which element every fragment register holds, where every digit byte goes,
and the arithmetic of the transform (scale folded into the distance, third-order square root, amplitude folded into
the Matern polynomial), so that the CPU tests can pin the index scheme and the accuracy of the transform without a GPU.
"""
import numpy as np

P = 6
BN, XROWS = 128, 64                      # candidates per CTA tile, training rows per stage
CHUNK_BYTES = 2 * P * 8 * 128            # one 32-training-point chunk of a 64-candidate MMA tile: [k half][plane][8][8][16 B]


def mma_m8n8k4(a_frag, b_frag, c_frag):
    """mma.sync.aligned.m8n8k4.row.col.f64 for one warp: lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2] and
    C[l >> 2][2 (l & 3) + {0, 1}] (csrc/ptx.cuh: dmma884).  a_frag, b_frag: (32,), c_frag: (32, 2)."""
    lanes = np.arange(32)
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    A[lanes >> 2, lanes & 3] = a_frag
    B[lanes & 3, lanes >> 2] = b_frag
    C[lanes >> 2, 2 * (lanes & 3)] = c_frag[:, 0]
    C[lanes >> 2, 2 * (lanes & 3) + 1] = c_frag[:, 1]
    D = A @ B + C
    return np.stack([D[lanes >> 2, 2 * (lanes & 3)], D[lanes >> 2, 2 * (lanes & 3) + 1]], axis=1)


def stage_dot_products(xs, xt_stage, d_pad):
    """The accumulators of one 64-row stage for all 8 warps, computed exactly as the kernel issues its fragments.

    xs: (d_pad, BN) scaled candidate tile (feature-major, as in shared memory); xt_stage: (XROWS, d_pad) scaled
    training rows of the stage.  Returns (claimed, value): for warp w, lane l, candidate tile ci, accumulator r the
    kernel treats acc as x[cand] . t[row] with cand = 64 e + 8 ci + (l >> 2), row = 16 rq + 4 (l & 3) + r."""
    lanes = np.arange(32)
    t4, c8 = lanes & 3, lanes >> 2
    out = []
    for warp in range(8):
        rq, e = warp >> 1, warp & 1
        acc = np.zeros((8, 2, 32, 2))                       # [ci][ji][lane][h]
        for k0 in range(0, d_pad, 4):
            # B fragment: feature k0 + t4 of the training row behind column c8: row 16 rq + 4 (c8 >> 1) + 2 ji + (c8 & 1)
            b = [xt_stage[16 * rq + 4 * (c8 >> 1) + 2 * ji + (c8 & 1), k0 + t4] for ji in range(2)]
            for ci in range(8):
                a = xs[k0 + t4, 64 * e + 8 * ci + c8]       # A fragment: feature k0 + t4, candidate 8 ci + c8
                for ji in range(2):
                    acc[ci, ji] = mma_m8n8k4(a, b[ji], acc[ci, ji])
        for ci in range(8):
            for r in range(4):
                cand = 64 * e + 8 * ci + c8
                row = 16 * rq + 4 * t4 + r
                out.append((cand, row, acc[ci, r >> 1, :, r & 1]))
    return out


def digit_byte_offset(e, rq, st, ci, c8, t4, r, plane, nkc_all):
    """Byte offset (from the start of the 128-candidate tile's two MMA tiles) of the digit of `plane` for candidate
    64 e + 8 ci + c8 and training row 64 st + 16 rq + 4 t4 + r - the address arithmetic of the kernel's 32-bit stores."""
    return (e * nkc_all + st * 2 + (rq >> 1)) * CHUNK_BYTES + (rq & 1) * (P * 8 * 128) + plane * (8 * 128) + \
        ci * 128 + c8 * 16 + 4 * t4 + r


def canonical_byte_offset(cand, train_row, plane, nkc_all):
    """The layout csrc/ozaki.cuh consumes: [MMA tile of 64 candidates][chunk of 32 training points][k half of 16]
    [plane][candidate group of 8][candidate in group][16 bytes = training points of the k half]."""
    e, cp = divmod(cand, 64)
    chunk, in_chunk = divmod(train_row, 32)
    half, k16 = divmod(in_chunk, 16)
    return (e * nkc_all + chunk) * CHUNK_BYTES + half * (P * 8 * 128) + plane * (8 * 128) + (cp >> 3) * 128 + (cp & 7) * 16 + k16


# ---------------------------------------------------------------- transform arithmetic
def fma(a, b, c):
    """Fused multiply-add emulated in extended precision (x86 longdouble: 64-bit mantissa, enough to make the single
    rounding of the products that occur here exact to well below the effects measured)."""
    return np.asarray(np.asarray(a, dtype=np.longdouble) * np.asarray(b, dtype=np.longdouble) +
                      np.asarray(c, dtype=np.longdouble), dtype=np.float64)


def sqrt_third_order(s, seed_rel_error):
    """sqrt_pos_third: y = (1 + d) / sqrt(s) from the hardware seed, g = s y, e = 1 - g y, g (1 + e/2 + 3 e^2/8)."""
    y = (1.0 + seed_rel_error) / np.sqrt(s)
    g = s * y
    e = fma(-g, y, 1.0)
    return fma(g * e, fma(0.375, e, 0.5), g)


def matern52_scaled(s5, amp, bias, seed_rel_error=0.0):
    """kernel_value_scaled<3>: s5 = 5 r^2; amplitude folded into the polynomial."""
    K = sqrt_third_order(np.maximum(s5, 1e-290), seed_rel_error)
    return fma(fma(K, fma(K, amp * (1.0 / 3.0), amp), amp), np.exp(-K), bias)


# ---------------------------------------------------------------------------------------------------------------
# exp_nonpos_table (csrc/kcross.cuh): 32-entry table + degree-6 polynomial, restated with NumPy (no fused multiply-add on
# the host: each fma is rounded twice here, so this restatement is a few 1e-17 LESS accurate than the kernel)
# ---------------------------------------------------------------------------------------------------------------
EXP32_TABLE = np.array([
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924,
    1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832, 1.4142135623730951, 1.4451808069770467,
    1.4768261459394993, 1.5091644275934228, 1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103,
    1.9152065613971474, 1.9571441241754002])


def exp_table(x):
    x = np.asarray(x, dtype=np.float64)
    magic = 6755399441055744.0
    fn = x * 46.16624130844683 + magic
    n = fn - magic
    r = (x + n * -0.02166084938653512) + n * -5.9631716539705866e-12
    p = r * (1.0 / 720.0) + 1.0 / 120.0
    for c in (1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0):
        p = p * r + c
    ni = n.astype(np.int64)
    out = np.ldexp(p * EXP32_TABLE[ni & 31], (ni >> 5).astype(np.int32))
    return np.where(x < -708.0, 0.0, out)
