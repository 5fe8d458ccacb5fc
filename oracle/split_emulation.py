"""NumPy emulation of the split-integer contraction of csrc/ozaki.cuh.  TEST INFRASTRUCTURE ONLY.

Restates, with exact Python/NumPy integer arithmetic, what the tcgen05 kernel computes: both operands quantised to
55-bit fixed point (per-row power-of-two scale for L_inv, one global scale for the kernel values), split into 7 balanced
base-256 digits, digit-plane products with p + q <= 6 accumulated exactly, recombined in float64 - or, with
`planes=6`, the 47-bit variant on the leading 6 digits (p + q <= 5, 21 products).  Used by the CPU tests to pin the
arithmetic of the scheme (digit extraction trick, leading digits = rounding, dropped-term bound, scaling) without a GPU.
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


def digits(x_scaled, planes=P):
    """Balanced base-256 digits (most significant first) of round(x): the (v + 0x80..80) ^ 0x80..80 byte trick.
    `x_scaled` = value / scale * 2^(8 planes - 1)."""
    v = np.rint(x_scaled).astype(np.int64)
    bias = sum(0x80 << (8 * i) for i in range(planes))
    y = (v + bias) ^ bias
    out = [((y >> (8 * (planes - 1 - p))) & 0xFF).astype(np.uint8).view(np.int8).astype(np.int64) for p in range(planes)]
    return out, v


def split_matmul(V, K, planes=P, k_scale=None):
    """U = V @ K evaluated like the kernel: V (r x n) row-scaled, K (n x c) globally scaled (k_scale: the kernel's
    a-priori scale pow2_scale(|amplitude| + |bias|); default: from the data)."""
    V = np.asarray(V, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    sV = np.array([pow2_scale(np.max(np.abs(row))) for row in V])
    sK = pow2_scale(np.max(np.abs(K))) if k_scale is None else k_scale
    a, _ = digits(V / sV[:, None] * 2.0 ** (8 * planes - 1), planes)
    b, _ = digits(K / sK * 2.0 ** (8 * planes - 1), planes)
    G = [np.zeros((V.shape[0], K.shape[1]), dtype=np.int64) for _ in range(planes)]
    for p in range(planes):
        for q in range(planes - p):
            # float64 matmul of small integers is exact (|sum| <= n 2^14 << 2^53) and runs on BLAS
            G[p + q] += np.rint(a[p].astype(np.float64) @ b[q].astype(np.float64)).astype(np.int64)
    assert all(np.max(np.abs(g)) < 2 ** 31 for g in G), "int32 accumulators would overflow"
    acc = G[planes - 1].astype(np.float64)
    for t in range(planes - 2, -1, -1):                 # Horner in 256^-1, as in the kernel epilogue
        acc = acc * 0.00390625 + G[t].astype(np.float64)
    return acc * (sV[:, None] * (sK * 2.0 ** -14))      # the 2^-14 does not depend on the number of planes


SPLIT_BOUND_GATE = 2.5e-10


def apriori_bounds(L_inv, alpha, Xs, amplitude, bias, white, planes=6):
    """Restatement of split_error_bounds (csrc/skb.cu): 8 standard deviations of the rounding model of the `planes`-plane
    contraction, as (bound on |d var| / prior variance, bound on |d mean| / y_std).  L_inv: lower-triangular inverse
    Cholesky factor; Xs: training rows divided by the length scales."""
    V = np.tril(np.asarray(L_inv, dtype=np.float64))
    n = V.shape[0]
    sV = np.array([pow2_scale(np.max(np.abs(row))) for row in V])
    sK = pow2_scale(abs(amplitude) + abs(bias))
    P = planes
    q2 = 2.0 ** -(16 * P - 2)
    r2 = 2.0 * float(np.max(np.sum(np.asarray(Xs) ** 2, axis=1)))
    eps_k2 = sK * sK * q2 / 3.0 + (abs(amplitude) * 2.0 ** -52 * r2) ** 2
    per_entry = sK * sK * q2 / 48.0 + eps_k2 / 4.0 + (P - 1) * sK * sK * 2.0 ** (-16 * P) * 0.1111
    s_max = float(np.max(np.sqrt(np.arange(1, n + 1)) * sV)) * np.sqrt(per_entry)
    prior = abs(amplitude + bias + white)
    bound_var = 8.0 * 2.0 * np.sqrt(abs(amplitude) + abs(bias)) * s_max / prior
    bound_mean = 8.0 * float(np.linalg.norm(alpha)) * np.sqrt(eps_k2)
    return bound_var, bound_mean
