"""CPU pin of the DMMA tile scheme of the 6-plane K1 (oracle/k1_mma_emulation.py restates csrc/kcross.cuh:
kcross_digits_mma_kernel): every accumulator register holds the dot product the kernel believes it holds, every digit
byte lands where the contraction kernel reads it, each exactly once, and the trimmed transform keeps FP64 accuracy."""
import numpy as np

from oracle import k1_mma_emulation as em


def test_every_accumulator_is_the_dot_product_of_its_claimed_pair():
    rng = np.random.RandomState(0)
    for d_pad in (4, 20):
        xs = rng.randn(d_pad, em.BN)
        xt = rng.randn(em.XROWS, d_pad)
        exact = xs.T @ xt.T                                          # (candidate, training row)
        seen = np.zeros((em.BN, em.XROWS), dtype=int)
        for cand, row, value in em.stage_dot_products(xs, xt, d_pad):
            assert np.allclose(value, exact[cand, row], rtol=1e-13, atol=1e-13)
            seen[cand, row] += 1
        assert np.all(seen == 1)                                     # 128 x 64 pairs, each owned by exactly one register


def test_digit_stores_tile_the_canonical_layout_exactly_once():
    nkc_all, n_stages = 8, 4                                         # n_pad = 256: 8 chunks of 32 training points
    hits = np.zeros(2 * nkc_all * em.CHUNK_BYTES, dtype=int)
    for st in range(n_stages):
        for warp in range(8):
            rq, e = warp >> 1, warp & 1
            for lane in range(32):
                t4, c8 = lane & 3, lane >> 2
                for ci in range(8):
                    for plane in range(em.P):
                        for r in range(4):
                            off = em.digit_byte_offset(e, rq, st, ci, c8, t4, r, plane, nkc_all)
                            cand, row = 64 * e + 8 * ci + c8, 64 * st + 16 * rq + 4 * t4 + r
                            assert off == em.canonical_byte_offset(cand, row, plane, nkc_all)
                            hits[off] += 1
                        # the four rows of a lane are one aligned 32-bit word
                        assert em.digit_byte_offset(e, rq, st, ci, c8, t4, 0, plane, nkc_all) % 4 == 0
    assert np.all(hits == 1)


def test_warp_stores_are_contiguous_128_byte_lines():
    for plane in range(em.P):
        for ci in range(8):
            words = sorted(em.digit_byte_offset(0, 1, 2, ci, lane >> 2, lane & 3, 0, plane, 8) for lane in range(32))
            assert words[0] % 128 == 0 and words == list(range(words[0], words[0] + 128, 4))


def test_third_order_square_root_and_folded_polynomial():
    rng = np.random.RandomState(1)
    s = np.concatenate([rng.rand(50000) * 60.0, 10.0 ** rng.uniform(-14, 2, 50000)])
    exact = np.sqrt(s.astype(np.longdouble))
    for seed_err in (2.0 ** -20, -2.0 ** -20, 3e-7, 0.0):
        got = em.sqrt_third_order(s, seed_err * rng.uniform(0.5, 1.0, s.size))
        assert np.max(np.abs((got - exact) / exact)) <= 2.0 * 2.0 ** -53          # <= 2 ulp from a 2^-20 seed
    amp, bias = 1.3, 0.25
    r2 = s[:50000] / 5.0
    K = np.sqrt(5.0 * r2.astype(np.longdouble))
    reference = amp * (1 + K + K * K / 3) * np.exp(-K) + bias
    got = em.matern52_scaled(5.0 * r2, amp, bias, seed_rel_error=2.0 ** -20)
    # the 6-plane path rounds kernel values to 2^-47 of the kernel scale: the transform must be far inside that
    assert np.max(np.abs(got - reference)) <= 2.0 ** -50 * (amp + bias)
    assert em.matern52_scaled(np.array([0.0, -1e-17]), amp, bias)[0] == amp + bias  # x == t (also when cancellation leaves -0)


def test_table_exponential_accuracy():
    """exp_nonpos_table (the 6-plane K1's exponential since round 2): 32-entry table + degree-6 polynomial.  The NumPy
    restatement (double rounding instead of FMA, so slightly worse than the kernel) stays within 4e-16 relative of a
    high-precision exponential over the whole argument range of the kernels; the 6-plane path needs 3.5e-15 of the
    kernel SCALE; entries of the table are the correctly rounded 2^(j/32)."""
    from decimal import Decimal, getcontext
    from oracle import k1_mma_emulation as em
    getcontext().prec = 50
    ln2 = Decimal(2).ln()
    for j in range(32):
        assert em.EXP32_TABLE[j] == float((ln2 * j / 32).exp())
    rng = np.random.RandomState(0)
    x = -np.concatenate([rng.rand(20000) * 40.0, rng.rand(2000) * 1e-6, rng.rand(2000) * 700.0, [0.0, 1e-300, 708.0]])
    got = em.exp_table(x)
    ref = np.array([float(Decimal(float(v)).exp()) for v in x[::37]])
    rel = np.abs(got[::37] - ref) / ref
    assert rel.max() <= 4e-16, rel.max()
    assert np.max(np.abs(got - np.exp(x)) / np.exp(x)) <= 6e-16          # and against libm on all points
    assert em.exp_table(np.array([-709.0, -1e4]))[0] == 0.0 and em.exp_table(np.array([-1e4]))[0] == 0.0
    assert em.exp_table(np.array([0.0]))[0] == 1.0
