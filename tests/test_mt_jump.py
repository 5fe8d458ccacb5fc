"""CPU pin of the MT19937 jump-ahead arithmetic (skopt_b200/mt_jump.py) against numpy.random.RandomState itself: the
window / shift representation, the characteristic polynomial, z^J mod phi applied by Horner's rule (the arithmetic of
csrc/rng.cuh: mt_jump_kernel), the per-dimension column starts and the state handed back to the host generator."""
import numpy as np
import pytest

from skopt_b200 import mt_jump as mj


def _temper(y):
    y = y.copy()
    y ^= y >> 11
    y ^= (y << 7) & 0x9D2C5680
    y ^= (y << 15) & 0xEFC60000
    y ^= y >> 18
    return y


def _raw_words(rng, count):
    """The next `count` TEMPERED 32-bit outputs of a RandomState (one generator word each)."""
    return rng.randint(0, 2 ** 32, size=count, dtype=np.uint64).astype(np.uint32)


def _walk(window, count):
    out, w = np.empty(count, dtype=np.uint32), window.copy()
    for i in range(count):
        out[i] = w[0]
        w = mj.step(w)
    return out


@pytest.mark.parametrize("burn", [0, 1, 397, 623, 624, 1000])
def test_window_and_shift_reproduce_the_generator(burn):
    rng = np.random.RandomState(5)
    _raw_words(rng, burn)
    _, key, pos, _, _ = rng.get_state()
    w = mj.window_from_state(key, pos)
    assert np.array_equal(_temper(_walk(w, 1300)), _raw_words(rng, 1300))


def test_characteristic_polynomial_annihilates_every_window():
    phi = mj.characteristic_polynomial()
    assert phi.bit_length() - 1 == mj.DEGREE and phi & 1                 # monic, degree 19937, non-zero constant term
    words = np.frombuffer(phi.to_bytes(mj.N * 4, "little"), dtype="<u4")
    rng = np.random.RandomState(9)
    _, key, pos, _, _ = rng.get_state()
    z = mj.apply_polynomial(words, mj.window_from_state(key, pos))
    assert not z[1:].any() and not (int(z[0]) & 0x80000000)               # only the ignored low 31 bits of word 0 may survive


def test_polynomial_arithmetic():
    a, b = 0b1011, 0b110
    assert mj.poly_mul(a, b) == 0b111010 ^ 0 and mj.poly_square(a) == 0b1000101   # (z^3+z+1)(z^2+z), (z^3+z+1)^2
    rng = np.random.RandomState(0)
    big = [int.from_bytes(rng.bytes(2500), "little") for _ in range(3)]
    assert mj.poly_mul(big[0], big[1]) == mj.poly_mul(big[1], big[0])
    assert mj.poly_mul(big[0], big[1] ^ big[2]) == mj.poly_mul(big[0], big[1]) ^ mj.poly_mul(big[0], big[2])
    assert mj.poly_square(big[0]) == mj.poly_mul(big[0], big[0])
    phi = mj.characteristic_polynomial()
    assert mj.z_power(5) == 1 << 5 and mj.z_power(mj.DEGREE) == phi ^ (1 << mj.DEGREE)
    assert mj.poly_mod(mj.poly_mul(mj.z_power(123457), mj.z_power(76543)), phi, mj.DEGREE) == mj.z_power(200000)


@pytest.mark.parametrize("burn,words_per_column,n_columns", [(0, 624, 3), (311, 2 * 700, 4), (624, 2 * 5001, 3)])
def test_column_starts_match_the_serial_stream(burn, words_per_column, n_columns):
    rng = np.random.RandomState(21)
    _raw_words(rng, burn)
    _, key, pos, _, _ = rng.get_state()
    w = mj.window_from_state(key, pos)
    polys = mj.column_jump_polynomials(words_per_column, n_columns)
    assert polys.shape == (n_columns - 1, mj.N)
    total = words_per_column * n_columns
    stream = _raw_words(rng, total + 700)                                  # what the host generator produces
    for j in range(1, n_columns):
        start = mj.step(mj.apply_polynomial(polys[j - 1], w))               # Horner with z^(J-1), then one plain step
        assert np.array_equal(_temper(_walk(start, 650)), stream[j * words_per_column:j * words_per_column + 650])


@pytest.mark.parametrize("burn,total", [(0, 1000), (624, 624), (17, 5000), (623, 1249), (1, 1247), (300, 624 * 7 + 324)])
def test_canonical_state_is_numpys_own(burn, total):
    rng = np.random.RandomState(3)
    _raw_words(rng, burn if burn else 624)
    name, key, pos, has_gauss, cached = rng.get_state()
    raw = _walk(mj.window_from_state(key, pos), total + mj.N)
    new_key, new_pos = mj.canonical_state(pos, total, raw[total - mj.N:total + mj.N])
    _raw_words(rng, total)
    _, ref_key, ref_pos, _, _ = rng.get_state()
    assert new_pos == ref_pos and np.array_equal(new_key, ref_key)
    with pytest.raises(ValueError):
        mj.canonical_state(0, 100, raw[:2 * mj.N])


def test_background_computation_hands_over_the_same_polynomials():
    import time
    key = (2 * 4321, 3)
    with mj._JUMP_LOCK:
        mj._JUMP_CACHE.pop(key, None)
    assert mj.column_jump_polynomials_nowait(*key) is None               # started in the background: the caller goes serial
    worker = mj._JUMP_PENDING.get(key)
    mj.column_jump_polynomials_nowait(*key)                              # a second request never starts a second worker
    assert mj._JUMP_PENDING.get(key) in (worker, None)
    deadline = time.time() + 60
    polys = None
    while polys is None and time.time() < deadline:
        time.sleep(0.05)
        polys = mj.column_jump_polynomials_nowait(*key)
    assert polys is not None and np.array_equal(polys, mj._compute_column_polynomials(*key))
    assert mj.column_jump_polynomials(*key) is polys                     # the blocking call sees the cached result


def test_automatic_mode_policy():
    from skopt_b200 import engine
    assert engine._jump_wanted(None, 10 ** 6, 20) and engine._jump_wanted("auto", 52429, 20)
    assert not engine._jump_wanted(None, 10000, 20)                       # small sets: the serial walk takes < 1 ms
    assert not engine._jump_wanted(True, 311, 20) and not engine._jump_wanted(True, 10 ** 6, 1)
    assert engine._jump_wanted(True, 312, 2) and not engine._jump_wanted(False, 10 ** 7, 20)
