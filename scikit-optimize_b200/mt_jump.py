"""Jump-ahead for NumPy's legacy MT19937 stream: the host half of the parallel device candidate generator.

`space.rvs` draws a candidate set dimension by dimension (skopt/space/space.py:874-974 via optimizer.py:552): the
stream order is [dimension][candidate], so the only way to generate it in parallel is to start every dimension's
column from its own point of the ONE stream the host generator would have walked.  MT19937 is linear over GF(2): with
w_t = (x_t, ..., x_{t+623}) the window of raw words that are output next and A the map w_t -> w_{t+1}
(x_{t+624} = x_{t+397} ^ twist(x_t, x_{t+1})), w_{t+J} = A^J w_t = g(A) w_t for g(z) = z^J mod phi(z), phi the
characteristic polynomial (degree 19937) - up to the low 31 bits of the FIRST word of the window, which A ignores and
drops.  Hence: g = z^(J-1) mod phi applied by Horner's rule, then one plain step (csrc/rng.cuh: mt_jump_kernel).

This module does the polynomial arithmetic (Python integers as GF(2)[z] elements, bit i = coefficient of z^i):
phi by Berlekamp-Massey from the generator's own output (computed once per process, ~1 s), z^J mod phi by square and
multiply, carry-less products by Kronecker substitution on ordinary big-integer multiplication.  Everything here is
checked on the CPU against numpy.random.RandomState itself (tests/test_mt_jump.py).
"""
import threading

import numpy as np

N, M = 624, 397
DEGREE = 19937
_UPPER, _LOWER, _MATRIX_A = 0x80000000, 0x7FFFFFFF, 0x9908B0DF


def _twist(u, v):
    y = (u & _UPPER) | (v & _LOWER)
    return (y >> 1) ^ np.where(v & 1, np.uint32(_MATRIX_A), np.uint32(0))


def regenerate(key):
    """The next block of 624 raw words (what genrand_int32 computes when its position reaches 624)."""
    o = np.asarray(key, dtype=np.uint32)
    n = np.empty(N, dtype=np.uint32)
    n[0:227] = o[M:N] ^ _twist(o[0:227], o[1:228])
    n[227:454] = n[0:227] ^ _twist(o[227:454], o[228:455])
    n[454:623] = n[227:396] ^ _twist(o[454:623], o[455:624])
    n[623] = n[396] ^ _twist(o[623:624], n[0:1])[0]
    return n


def window_from_state(key, pos):
    """w = the next 624 raw words of a RandomState whose get_state() holds (key, pos)."""
    key = np.asarray(key, dtype=np.uint32)
    pos = int(pos)
    if pos == 0:
        return key.copy()
    nxt = regenerate(key)
    return np.concatenate([key[pos:], nxt[:pos]])


def step(w):
    """A w."""
    out = np.empty(N, dtype=np.uint32)
    out[:-1] = w[1:]
    out[-1] = w[M] ^ _twist(w[0:1], w[1:2])[0]
    return out


def canonical_state(pos, total_words, tail):
    """NumPy's own (key, pos) after `total_words` draws from a state at position `pos`, given tail = the raw words at
    stream offsets [total_words - 624, total_words + 624) counted from the first word drawn."""
    q = int(pos) + int(total_words)
    if q <= N:
        raise ValueError("fewer words than one regeneration: the state keeps its key")
    new_pos = q - N * ((q - 1) // N)
    start = N - new_pos                                     # offset total_words - new_pos, relative to tail[0]
    return np.ascontiguousarray(tail[start:start + N], dtype=np.uint32), new_pos


# ------------------------------------------------------------------------------------------------ GF(2)[z]
def _spread(a, stride):
    """Bits of a moved to positions stride * i."""
    if a == 0:
        return 0
    return int(("0" * (stride - 1)).join(bin(a)[2:]), 2)


def poly_mod(p, phi, deg):
    while p.bit_length() > deg:
        p ^= phi << (p.bit_length() - 1 - deg)
    return p


def poly_mul(a, b):
    """Carry-less product: with the bits 16 apart no column sum (<= 19937 < 2^16) reaches the next field."""
    prod = _spread(a, 16) * _spread(b, 16)
    if prod == 0:
        return 0
    raw = prod.to_bytes((prod.bit_length() + 15) // 16 * 2, "little")
    parity = (np.frombuffer(raw, dtype="<u2") & 1).astype(np.uint8)
    return int.from_bytes(np.packbits(parity, bitorder="little").tobytes(), "little")


def poly_square(a):
    return _spread(a, 2)


def berlekamp_massey(bits):
    """Connection polynomial C (c_0 = 1, s_n = sum_{i>=1} c_i s_{n-i}) and linear complexity L of a bit sequence."""
    C, B, L, m = 1, 1, 0, 1
    window = 0                                               # bit i = s_{n-i}
    for n, s in enumerate(bits):
        window = (window << 1) | int(s)
        if (C & window).bit_count() & 1:
            T = C
            C ^= B << m
            if 2 * L <= n:
                L, B, m = n + 1 - L, T, 1
            else:
                m += 1
        else:
            m += 1
    return C, L


_PHI = [None]


def characteristic_polynomial():
    """phi(z), monic of degree 19937: the minimal polynomial of any output bit of the generator."""
    if _PHI[0] is None:
        rng = np.random.RandomState(20260101)
        words = rng.randint(0, 2 ** 32, size=2 * DEGREE + 64, dtype=np.uint64).astype(np.uint32)   # one raw draw per word
        bits = ((words >> np.uint32(7)) & np.uint32(1)).tolist()     # a bit of the TEMPERED output: still a linear functional
        C, L = berlekamp_massey(bits)
        if L != DEGREE:
            raise RuntimeError("Berlekamp-Massey found linear complexity %d, expected %d" % (L, DEGREE))
        # characteristic polynomial = reversal of the connection polynomial
        _PHI[0] = int(bin(C)[2:].zfill(DEGREE + 1)[::-1], 2)
    return _PHI[0]


def z_power(J):
    """z^J mod phi."""
    phi = characteristic_polynomial()
    result, J = 1, int(J)
    for bit in bin(J)[2:]:
        result = poly_mod(poly_square(result), phi, DEGREE)
        if bit == "1":
            result = poly_mod(result << 1, phi, DEGREE)
    return result


_JUMP_CACHE = {}
_JUMP_PENDING = {}
_JUMP_LOCK = threading.Lock()


def _compute_column_polynomials(words_per_column, n_columns):
    phi = characteristic_polynomial()
    L = int(words_per_column)
    if L < 1:
        raise ValueError("words_per_column must be positive")
    stride = z_power(L)
    # z^-1 mod phi = (phi + 1) / z (phi has constant term 1), so z^(L-1) is one product away from z^L
    g = poly_mod(poly_mul(stride, (phi ^ 1) >> 1), phi, DEGREE)
    out = np.zeros((max(n_columns - 1, 0), N), dtype=np.uint32)
    for j in range(1, n_columns):
        out[j - 1] = np.frombuffer(g.to_bytes(N * 4, "little"), dtype="<u4")
        g = poly_mod(poly_mul(g, stride), phi, DEGREE)
    return out


def _store(key, polys):
    with _JUMP_LOCK:
        if len(_JUMP_CACHE) > 8:
            _JUMP_CACHE.clear()
        _JUMP_CACHE[key] = polys
        _JUMP_PENDING.pop(key, None)


def column_jump_polynomials(words_per_column, n_columns):
    """g_j = z^(j L - 1) mod phi for the columns j = 1 .. n_columns - 1 (L = words_per_column), as an
    (n_columns - 1, 624) uint32 array: bit b of word k = coefficient of z^(32 k + b).  About 25 ms per column, cached
    per (L, n_columns)."""
    key = (int(words_per_column), int(n_columns))
    with _JUMP_LOCK:
        polys = _JUMP_CACHE.get(key)
    if polys is None:
        polys = _compute_column_polynomials(*key)
        _store(key, polys)
    return polys


def column_jump_polynomials_nowait(words_per_column, n_columns):
    """The cached polynomials, or None after starting their computation in a background thread: a caller that has a
    serial generator producing the same bits (engine.device_candidates) uses it until the polynomials are there, so
    that the first draws of an optimisation never wait for them."""
    key = (int(words_per_column), int(n_columns))
    with _JUMP_LOCK:
        polys = _JUMP_CACHE.get(key)
        if polys is not None or key in _JUMP_PENDING:
            return polys
        worker = threading.Thread(target=lambda: _store(key, _compute_column_polynomials(*key)), daemon=True,
                                  name="skb-mt-jump-%d-%d" % key)
        _JUMP_PENDING[key] = worker
    worker.start()
    return None


def apply_polynomial(words, w):
    """g(A) w by Horner's rule (the arithmetic of csrc/rng.cuh: mt_jump_kernel; CPU tests only)."""
    g = int.from_bytes(np.ascontiguousarray(words, dtype="<u4").tobytes(), "little")
    h = np.zeros(N, dtype=np.uint32)
    for i in range(g.bit_length() - 1, -1, -1):
        h = step(h)
        if (g >> i) & 1:
            h ^= w
    return h
