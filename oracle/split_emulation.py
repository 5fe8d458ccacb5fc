"""NumPy emulation of the split-integer contraction of csrc/ozaki.cuh.  TEST INFRASTRUCTURE ONLY.

Restates, with exact Python/NumPy integer arithmetic, what the tcgen05 kernel computes: both operands quantised to
55-bit fixed point (per-row power-of-two scale for L_inv, one global scale for the kernel values), split into 7 balanced
base-256 digits, digit-plane products with p + q <= 6 accumulated exactly, recombined in float64.  Used by the CPU
tests to pin the arithmetic of the scheme (digit extraction trick, dropped-term bound, scaling) without a GPU.
"""
import numpy as np

P = 7
BIAS = 0x0080808080808080


def pow2_scale(maxabs):
    """Power of two s with maxabs / s in [0.25, 0.5) (oz_pow2_scale)."""
    if not maxabs > 0:
        return 1.0
    _, e = np.frexp(maxabs)
    return float(np.ldexp(1.0, int(e) + 1))


def digits(x_scaled_2p55):
    """Balanced base-256 digits (most significant first) of round(x): the (v + 0x80..80) ^ 0x80..80 byte trick."""
    v = np.rint(x_scaled_2p55).astype(np.int64)
    y = (v + BIAS) ^ BIAS
    planes = [((y >> (8 * (6 - p))) & 0xFF).astype(np.uint8).view(np.int8).astype(np.int64) for p in range(P)]
    return planes, v


def split_matmul(V, K):
    """U = V @ K evaluated like the kernel: V (r x n) row-scaled, K (n x c) globally scaled."""
    V = np.asarray(V, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    sV = np.array([pow2_scale(np.max(np.abs(row))) for row in V])
    sK = pow2_scale(np.max(np.abs(K)))
    a, _ = digits(V / sV[:, None] * 2.0 ** 55)
    b, _ = digits(K / sK * 2.0 ** 55)
    G = [np.zeros((V.shape[0], K.shape[1]), dtype=np.int64) for _ in range(P)]
    for p in range(P):
        for q in range(P - p):
            G[p + q] += a[p] @ b[q]
    assert all(np.max(np.abs(g)) < 2 ** 31 for g in G), "int32 accumulators would overflow"
    acc = G[P - 1].astype(np.float64)
    for t in range(P - 2, -1, -1):                      # Horner in 256^-1, as in the kernel epilogue
        acc = acc * 0.00390625 + G[t].astype(np.float64)
    return acc * (sV[:, None] * (sK * 2.0 ** -14))
