"""Split-integer (int8 digit planes on tcgen05) evaluation of the variance contraction: same parity gates as the FP64
DMMA path, checked against the reference fixtures, the oracle and the FP64 path itself."""
import numpy as np
import pytest

import skopt_b200
from skopt_b200 import engine
from oracle import gp_oracle as orc
from tests._golden import NAMES, load, load_state, var_tol
from tests.test_gpu_parity import ACQS, check_against, host_state_from_oracle_state

pytestmark = pytest.mark.gpu


SPLIT_MODES = ("split_int8", "split_int8_p7")        # self-checked 6-or-7 planes, always 7 planes


@pytest.mark.parametrize("precision", SPLIT_MODES)
@pytest.mark.parametrize("name", NAMES)
def test_split_golden_fixture(name, precision):
    rec, meta, est = load(name)
    _, _, ost = load_state(name)
    dev = engine.DeviceState.from_estimator(est)
    if precision == "split_int8":
        dev.decide_split()              # small candidate sets would otherwise never reach the 6-plane kernels
    res = dev.score(rec["Xc"], y_opt=float(rec["y_opt"]), acq_funcs=ACQS, xi=float(rec["xi"]), kappa=float(rec["kappa"]),
                    top_k=10, precision=precision)
    vtol = var_tol(ost)
    acq_ref = {a: rec["acq_" + a] for a in ACQS}
    check_against(ost, res, rec["mean"], rec["std"], acq_ref, vtol, float(rec["y_opt"]), float(rec["xi"]),
                  float(rec["kappa"]), check_argmin=vtol <= 1e-9)
    dev.close()


@pytest.mark.parametrize("n,d,m,kind,nu", [
    (300, 7, 5000, orc.KIND_MATERN, 2.5),
    (129, 3, 1000, orc.KIND_MATERN, 1.5),
    (64, 1, 257, orc.KIND_MATERN, 0.5),
    (512, 20, 4096, orc.KIND_RBF, np.inf),
    (1024, 10, 3000, orc.KIND_MATERN, 2.5),
    (37, 50, 200, orc.KIND_MATERN, 2.5),
    (1, 1, 1, orc.KIND_MATERN, 2.5),
    (600, 100, 9000, orc.KIND_MATERN, 2.5),   # widest state: the DMMA digit kernel with one CTA per SM
    (520, 77, 2000, orc.KIND_MATERN, 0.5),
])
@pytest.mark.parametrize("precision", SPLIT_MODES)
def test_split_synthetic_vs_oracle_and_fp64(n, d, m, kind, nu, precision):
    st = orc.make_synthetic_state(n, d, noise=1e-3, seed=n + d, kind=kind, nu=nu, normalize_y=n > 1)
    rng = np.random.RandomState(7)
    Xc = rng.rand(m, d)
    Xc[0] = st.X_train[0]
    y_opt = float(st.meta["y"].min())
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    if precision == "split_int8":
        dev.decide_split()
    res = dev.score(Xc, y_opt=y_opt, acq_funcs=ACQS, top_k=min(10, m), precision=precision)
    ref = dev.score(Xc, y_opt=y_opt, acq_funcs=ACQS, top_k=min(10, m), precision="fp64")
    mu, sd = orc.predict(st, Xc, return_std=True)
    acq_ref = {a: orc.gaussian_acquisition(Xc, st, y_opt, a, acq_func_kwargs=dict(xi=0.01, kappa=1.96)) for a in ACQS}
    check_against(st, res, mu, sd, acq_ref, var_tol(st), y_opt, k=min(10, m))
    # the two device formulations agree far tighter than the gate
    scale = st.prior_var * st.y_std ** 2
    print("%s (planes %s) vs fp64: max |d var| / prior = %.2e" % (precision, dev.split_info(),
                                                                   np.max(np.abs(res["std"] ** 2 - ref["std"] ** 2)) / scale))
    assert np.max(np.abs(res["std"] ** 2 - ref["std"] ** 2)) <= (2e-10 if precision == "split_int8" else 1e-11) * scale
    for a in ACQS:
        assert res[a]["topk_idx"][0] == ref[a]["topk_idx"][0]
    dev.close()


def test_split_full_size_c3():
    n, d, m = 2048, 20, 1_000_000            # the headline configuration at its full candidate count
    st = orc.make_synthetic_state(n, d, noise=1e-4, seed=0)
    rng = np.random.RandomState(123)
    Xc = rng.rand(m, d)
    Xc[150_000] = Xc[17]
    y_opt = float(st.meta["y"].min())
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    res = dev.score(Xc, y_opt=y_opt, acq_funcs=ACQS, top_k=20, precision="split_int8")
    ref = dev.score(Xc, y_opt=y_opt, acq_funcs=ACQS, top_k=20, precision="fp64")
    assert dev.split_info()[0] == 6 and max(dev.split_bounds()) <= 2.5e-10       # bound AND probe let this state take 6 planes
    sample = rng.choice(m, 1000, replace=False)
    mu, sd = orc.predict(st, Xc[sample], return_std=True)
    assert np.max(np.abs(res["mean"][sample] - mu)) <= 1e-9 * max(np.max(np.abs(mu)), st.y_std)
    assert np.max(np.abs(res["std"][sample] ** 2 - sd ** 2)) <= 1e-9 * st.prior_var * st.y_std ** 2
    assert np.max(np.abs(res["std"][sample] - sd) / sd) <= 1e-9
    err = np.max(np.abs(res["std"] ** 2 - ref["std"] ** 2)) / (st.prior_var * st.y_std ** 2)
    print("C3 split vs fp64: max |d var| / prior = %.2e, max rel d std = %.2e" % (err, np.max(np.abs(res["std"] / ref["std"] - 1))))
    assert err <= 2e-10
    for a in ACQS:
        v = res[a]["values"]
        assert np.array_equal(res[a]["topk_idx"], np.lexsort((np.arange(m), v))[:20])
        assert res[a]["topk_idx"][0] == ref[a]["topk_idx"][0]
        assert v[17] == v[150_000]
        assert np.max(np.abs(v - ref[a]["values"])) <= 1e-9 * np.max(np.abs(ref[a]["values"]))
    dev.close()


@pytest.mark.parametrize("n,d,m,kind,nu", [
    (300, 7, 200, orc.KIND_MATERN, 2.5),
    (129, 3, 130, orc.KIND_MATERN, 1.5),
    (200, 50, 100, orc.KIND_MATERN, 2.5),
    (256, 20, 150, orc.KIND_RBF, np.inf),
    (1024, 10, 300, orc.KIND_MATERN, 2.5),
])
def test_split_gradient_path(n, d, m, kind, nu):
    """Gradient path with the dense split-integer contraction W = K_inv K_trans^T: same gates as the FP64 path."""
    from tests.test_gpu_parity import check_grads
    st = orc.make_synthetic_state(n, d, noise=1e-3, seed=n + d, kind=kind, nu=nu)
    rng = np.random.RandomState(11)
    Xc = rng.rand(m, d)
    Xc[3] = st.X_train[0]
    y_opt = float(st.meta["y"].min())
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    g = dev.score_grad(Xc, y_opt=y_opt, acq_funcs=ACQS, precision="split_int8")
    g64 = dev.score_grad(Xc, y_opt=y_opt, acq_funcs=ACQS, precision="fp64")
    k = min(m, 60)
    mu, sd, gm, gs = orc.predict_with_grads_batched(st, Xc[:k])
    acq_ref = {}
    z2 = ((y_opt - 0.01 - mu) / np.maximum(sd, 1e-300)) ** 2
    for a in ACQS:
        vs, grads = np.empty(k), np.empty((k, d))
        for i in range(k):
            v, gr = orc.gaussian_acquisition_1D(Xc[i], st, y_opt, a, dict(xi=0.01, kappa=1.96))
            vs[i], grads[i] = v[0], gr
        acq_ref[a] = (vs, grads, z2)
    # K_inv (entries up to ~1/noise) is quantised to 2^-55 of its row maximum: the variance-scale noise floor of this
    # formulation is ~1e-11 of the prior instead of the FP64 path's ~1e-12 (both far below the 1e-9 gate)
    check_grads(st, g, mu, sd, gm, gs, acq_ref, var_tol(st), floor=2e-11)
    scale = st.prior_var * st.y_std ** 2
    assert np.max(np.abs(g["std"] ** 2 - g64["std"] ** 2)) <= 1e-9 * scale
    assert np.max(np.abs(g["mean_grad"] - g64["mean_grad"])) <= 1e-12 * np.max(np.abs(g64["mean_grad"]))
    gv = np.abs((g["std_grad"] - g64["std_grad"]) * g64["std"][:, None])
    assert np.max(gv) <= 1e-9 * scale / np.min(st.length_scale)
    dev.close()


def test_split_extreme_scales():
    """Operands spanning many orders of magnitude: large amplitude, short and long length scales, tiny noise
    (large entries in L_inv / K_inv).  The split path must stay inside the same conditioning-aware gate."""
    rng = np.random.RandomState(3)
    n, d, m = 400, 4, 3000
    X = rng.rand(n, d)
    y = 50.0 * np.sin(6 * X[:, 0]) + 1e3 * X[:, 1] - 3.0
    from skopt_b200.learning import GaussianProcessRegressor
    from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern
    for amp, ls, noise in ((900.0, [0.05, 0.3, 5.0, 80.0], 1e-6), (1e-3, [0.2] * 4, 1e-8), (3.0, [2.0] * 4, 1e-2)):
        gpr = GaussianProcessRegressor(ConstantKernel(amp, "fixed") * Matern(np.array(ls), "fixed", nu=2.5) + 0.25,
                                       normalize_y=True, noise=noise, optimizer=None).fit(X, y)
        ost = orc.state_from_estimator(gpr)
        Xc = rng.rand(m, d)
        Xc[:5] = X[:5]
        dev = gpr.device_state()
        planes, probe = dev.decide_split()
        bv, bm = dev.split_bounds()
        print("self-check: planes %d, probe %.2e, a-priori bounds var %.2e mean %.2e" % (planes, probe, bv, bm))
        assert planes == (6 if (max(bv, bm) <= 2.5e-10 and probe <= 1e-10) else 7)
        y_opt = float(y.min())
        res = dev.score(Xc, y_opt=y_opt, acq_funcs=ACQS, top_k=5, precision="split_int8")
        ref = dev.score(Xc, y_opt=y_opt, acq_funcs=ACQS, top_k=5, precision="fp64")
        mu, sd = orc.predict(ost, Xc, return_std=True)
        vtol = var_tol(ost)
        scale = ost.prior_var * ost.y_std ** 2
        assert np.max(np.abs(res["mean"] - mu)) <= 1e-9 * max(np.max(np.abs(mu)), ost.y_std)
        assert np.max(np.abs(res["std"] ** 2 - sd ** 2)) <= vtol * scale
        assert np.max(np.abs(res["std"] ** 2 - ref["std"] ** 2)) <= max(1e-10, 0.1 * vtol) * scale
        print("amp=%g noise=%g cond tol %.1e: split-fp64 %.2e, split-oracle %.2e" % (
            amp, noise, vtol, np.max(np.abs(res["std"] ** 2 - ref["std"] ** 2)) / scale,
            np.max(np.abs(res["std"] ** 2 - sd ** 2)) / scale))


def test_split_plane_self_check():
    """precision="split_int8" contracts 6 digit planes only on states whose self-check (256 probe candidates, FP64 against
    6 planes) agrees to 1e-10 of the prior variance; the choice, the measured difference and the two variants' accuracy."""
    st = orc.make_synthetic_state(2048, 20, noise=1e-4, seed=0)
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    assert dev.split_info() == (0, -1.0)                            # undecided until the first split-integer call
    rng = np.random.RandomState(5)
    Xc = rng.rand(20_000, 20)
    y_opt = float(st.meta["y"].min())
    p7 = dev.score(Xc, y_opt=y_opt, acq_funcs=("EI",), top_k=8, precision="split_int8_p7")
    assert dev.split_info()[0] == 0                                 # the 7-plane mode does not need the check
    small = dev.score(Xc, y_opt=y_opt, acq_funcs=("EI",), top_k=8, precision="split_int8")
    assert dev.split_info()[0] == 0                                 # nor does a call below 32768 candidates: 7 planes
    np.testing.assert_array_equal(small["std"], p7["std"])
    Xbig = np.concatenate([Xc, rng.rand(20_000, 20)])
    auto = dev.score(Xbig, y_opt=y_opt, acq_funcs=("EI",), top_k=8, precision="split_int8")
    auto = {"std": auto["std"][:20_000], "EI": auto["EI"]}
    planes, diff = dev.split_info()
    assert dev.decide_split() == (planes, diff)                     # decided: a no-op
    ref = dev.score(Xc, y_opt=y_opt, acq_funcs=("EI",), top_k=8, precision="fp64")
    scale = st.prior_var * st.y_std ** 2
    e7 = np.max(np.abs(p7["std"] ** 2 - ref["std"] ** 2)) / scale
    e_auto = np.max(np.abs(auto["std"] ** 2 - ref["std"] ** 2)) / scale
    print("planes=%d probe diff=%.2e; vs fp64: 7 planes %.2e, auto %.2e" % (planes, diff, e7, e_auto))
    assert planes == 6 and 0.0 <= diff <= 1e-10                     # this well-conditioned state takes the 6-plane path
    bv, bm = dev.split_bounds()
    assert e_auto <= bv <= 2.5e-10 and 0.0 <= bm <= 2.5e-10         # ... and the a-priori bound covers what was measured
    assert e7 <= 1e-12 and e_auto <= 1e-10 and e7 < e_auto
    ref_big = dev.score(Xbig, y_opt=y_opt, acq_funcs=("EI",), top_k=8, precision="fp64")
    assert np.array_equal(auto["EI"]["topk_idx"], ref_big["EI"]["topk_idx"])
    assert np.array_equal(p7["EI"]["topk_idx"], ref["EI"]["topk_idx"])
    mu, sd = orc.predict(st, Xc[:500], return_std=True)
    assert np.max(np.abs(auto["std"][:500] ** 2 - sd ** 2)) <= 1e-9 * scale
    dev.close()


@pytest.mark.parametrize("case", ["cooked_branin_gaussian", "clustered_noise1e-10"])
def test_split_falls_back_to_seven_planes(case):
    """States whose a-priori bound does not allow 47-bit operands MUST contract 7 planes, whatever the probe measured:
    the cooked Branin fixture (cond(K) = 7.5e8 after a real fit) and clustered points with noise 1e-10 (rows of L_inv scaled
    like 1e5).  There is no environment override (SKB_SPLIT_PLANES is gone), so split_info() is the decision."""
    import os
    os.environ["SKB_SPLIT_PLANES"] = "6"                # the round-1 override: must be ignored now
    try:
        if case == "cooked_branin_gaussian":
            rec, meta, est = load(case)
            _, _, st = load_state(case)
            dev = engine.DeviceState.from_estimator(est)
            Xc = rec["Xc"]
            y_opt = float(rec["y_opt"])
        else:
            rng = np.random.RandomState(4)
            st = orc.make_synthetic_state(320, 4, noise=1e-10, seed=0)        # 320 points in [0, 1]^4: near-duplicates
            dev = engine.DeviceState(host_state_from_oracle_state(st))
            Xc = rng.rand(2000, 4)
            y_opt = float(st.meta["y"].min())
        planes, probe = dev.decide_split()
        bv, bm = dev.split_bounds()
        print("%s: planes %d, probe %.2e, bounds var %.2e mean %.2e" % (case, planes, probe, bv, bm))
        assert max(bv, bm) > 2.5e-10
        assert planes == 7
        res = dev.score(Xc, y_opt=y_opt, acq_funcs=("EI",), top_k=4, precision="split_int8")
        p7 = dev.score(Xc, y_opt=y_opt, acq_funcs=("EI",), top_k=4, precision="split_int8_p7")
        np.testing.assert_array_equal(res["std"], p7["std"])              # what ran IS the 7-plane contraction
        np.testing.assert_array_equal(res["EI"]["values"], p7["EI"]["values"])
        dev.close()
    finally:
        os.environ.pop("SKB_SPLIT_PLANES", None)
