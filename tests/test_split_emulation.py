"""CPU pin of the split-integer scheme's arithmetic (oracle/split_emulation.py restates csrc/ozaki.cuh in NumPy)."""
import numpy as np

from oracle import split_emulation as se
from oracle import gp_oracle as orc


def test_digit_trick_reconstructs_the_integer():
    rng = np.random.RandomState(0)
    x = (rng.rand(10000) - 0.5) * 2.0 ** 55          # |x| <= 2^54, the range the kernel guarantees
    x[:4] = [0.0, 2.0 ** 54, -2.0 ** 54, 1.0]
    planes, v = se.digits(x)
    rebuilt = sum(planes[p] * 256 ** (6 - p) for p in range(se.P))
    assert np.array_equal(rebuilt, v)
    assert all(pl.min() >= -128 and pl.max() <= 127 for pl in planes)


def test_scale_puts_values_in_quarter_to_half():
    for m in (1e-300, 3e-9, 0.24, 0.25, 0.5, 1.0, 7.3, 1e12):
        s = se.pow2_scale(m)
        assert 0.25 <= m / s < 0.5 and np.log2(s) == np.round(np.log2(s))
    assert se.pow2_scale(0.0) == 1.0


def test_split_product_matches_extended_precision():
    """Error of the scheme relative to (row scale x kernel scale): a few 1e-16 per term, like an FP64 product."""
    st = orc.make_synthetic_state(96, 5, noise=1e-4, seed=1)
    from scipy.linalg import solve_triangular
    V = solve_triangular(st.L, np.eye(st.n), lower=True)
    Xc = np.random.RandomState(2).rand(40, 5)
    K = orc.kernel_cross(st, Xc).T                                   # (n, c)
    U = se.split_matmul(V, K)
    exact = (V.astype(np.longdouble) @ K.astype(np.longdouble)).astype(np.float64)
    fp64 = V @ K
    scale = np.max(np.abs(V), axis=1, keepdims=True) * np.max(np.abs(K))
    err_split = np.max(np.abs(U - exact) / scale)
    err_fp64 = np.max(np.abs(fp64 - exact) / scale)
    assert err_split <= 96 * 2.0 ** -52 and err_split <= 50 * max(err_fp64, 2.0 ** -53)
    q = np.sum(U * U, axis=0)
    q_ref = np.einsum("ci,ci->c", K.T @ st.K_inv, K.T)
    assert np.max(np.abs(q - q_ref)) <= 1e-11 * st.prior_var


def test_leading_six_digits_are_the_rounded_47_bit_value():
    """What lets the 6-plane contraction read the first 6 of the 7 packed planes of L_inv: dropping the last balanced
    digit rounds (|digit| <= 128 = half a unit of the digit before it), so the leading digits are a valid 6-digit
    expansion of the value to within half a unit of 2^-47 - the same bound as quantising to 47 bits directly."""
    rng = np.random.RandomState(3)
    x = (rng.rand(20000) - 0.5)                       # value / scale, |x| <= 0.5
    x[:3] = [0.5, -0.5, 2.0 ** -50]
    p7, _ = se.digits(x * 2.0 ** 55, 7)
    lead = sum(p7[p] * 256 ** (5 - p) for p in range(6))
    assert np.max(np.abs(lead - x * 2.0 ** 47)) <= 0.5 + 2.0 ** -8
    assert all(pl.min() >= -128 and pl.max() <= 127 for pl in p7)


def test_six_plane_contraction_meets_the_parity_gate_with_margin():
    """47-bit operands, 21 digit-plane products: the variance term stays within 1e-10 of the prior variance of the exact
    contraction (the self-check threshold of skb.cu decide_split_planes; the parity gate is 1e-9), 7 planes within 1e-12."""
    from scipy.linalg import solve_triangular
    for n, d, noise in ((384, 6, 1e-4), (256, 4, 1e-8), (512, 12, 1e-2)):
        st = orc.make_synthetic_state(n, d, noise=noise, seed=n)
        V = solve_triangular(st.L, np.eye(st.n), lower=True)
        Xc = np.random.RandomState(2).rand(64, d)
        Xc[::2] = st.X_train[:32]                                    # variance at the noise level: prior - q cancels
        K = orc.kernel_cross(st, Xc).T
        exact = V.astype(np.longdouble) @ K.astype(np.longdouble)
        q_exact = np.sum(exact * exact, axis=0).astype(np.float64)
        k_scale = se.pow2_scale(abs(st.amplitude) + abs(st.bias))
        err = {}
        for planes in (6, 7):
            U = se.split_matmul(V, K, planes=planes, k_scale=k_scale)
            err[planes] = np.max(np.abs(np.sum(U * U, axis=0) - q_exact)) / st.prior_var
        assert err[6] <= 1e-10 and err[7] <= 1e-12 and err[7] < err[6]


def _bound_case(st, m=160, seed=5):
    from scipy.linalg import solve_triangular
    V = solve_triangular(st.L, np.eye(st.n), lower=True)
    rng = np.random.RandomState(seed)
    lo, hi = st.X_train.min(0), st.X_train.max(0)
    Xc = lo + (hi - lo) * rng.rand(m, st.d)
    Xc[:m // 2] = st.X_train[rng.randint(0, st.n, m // 2)]        # at training points: prior - q cancels almost completely
    K = orc.kernel_cross(st, Xc).T
    q = np.sum((V @ K) ** 2, axis=0)
    sK = se.pow2_scale(abs(st.amplitude) + abs(st.bias))
    U6 = se.split_matmul(V, K, planes=6, k_scale=sK)
    measured = float(np.max(np.abs(np.sum(U6 * U6, axis=0) - q)) / st.prior_var)
    bound_var, bound_mean = se.apriori_bounds(V, st.alpha, st.X_train / st.length_scale, st.amplitude, st.bias, st.white)
    return measured, bound_var, bound_mean


def test_apriori_bound_covers_the_emulated_six_plane_error():
    """The a-priori bound behind the 6-plane decision (csrc/skb.cu split_error_bounds, restated in
    oracle/split_emulation.apriori_bounds) is an upper bound of what the exact-integer emulation of the 6-plane contraction
    measures, on well- and ill-conditioned states; it lets the bench-like states through (<= 2.5e-10) and sends the
    ill-conditioned ones (cooked Branin after a real fit, noise 1e-10 on clustered points) to 7 planes."""
    from tests import _golden
    allowed, refused = [], []
    for name in ("matern25_n256_d10", "rbf_aniso", "matern25_const_noise", "cooked_branin_gaussian"):
        _, _, st = _golden.load_state(name)
        measured, bv, bm = _bound_case(st)
        assert measured <= bv, (name, measured, bv)
        (allowed if max(bv, bm) <= se.SPLIT_BOUND_GATE else refused).append(name)
    for n, d, noise in ((384, 8, 1e-4), (320, 4, 1e-10)):
        st = orc.make_synthetic_state(n, d, noise=noise, seed=0)
        measured, bv, bm = _bound_case(st)
        assert measured <= bv, (n, d, noise, measured, bv)
        (allowed if max(bv, bm) <= se.SPLIT_BOUND_GATE else refused).append((n, d, noise))
    assert "matern25_n256_d10" in allowed and (384, 8, 1e-4) in allowed
    assert "cooked_branin_gaussian" in refused and (320, 4, 1e-10) in refused
