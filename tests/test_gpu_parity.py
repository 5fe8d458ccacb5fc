"""GPU parity: the CUDA path (through the C ABI) against the reference-generated fixtures and the CPU oracle.

Tolerances (north star: rel 1e-9 on mean, std and acquisition, identical argmin):
  mean  |d mu|   <= 1e-9 * max(|mu|, y_std)
  var   |d var|  <= max(1e-9, 1e-15*cond(K)) * prior_var * y_std^2   (sigma^2 is a cancellation residue)
  acq   normwise max|d acq| <= tol * max|acq|, and identical argmin / top-k
"""
import numpy as np
import pytest

import skopt_b200
from skopt_b200 import engine
from oracle import gp_oracle as orc
from tests._golden import NAMES, kernel_code, load, load_state, var_tol

pytestmark = pytest.mark.gpu

ACQS = ("EI", "PI", "LCB")


def host_state_from_oracle_state(st):
    from scipy.linalg import solve_triangular
    kind = kernel_code(st)
    L_inv = solve_triangular(st.L, np.eye(st.n), lower=True)
    return engine.HostState(st.X_train, st.alpha, L_inv, st.length_scale, kind, st.amplitude, st.bias, st.white,
                            st.y_mean, st.y_std, K_inv=st.K_inv)


def _acq_from_posterior(a, mu, sd, y_opt, xi, kappa):
    if a == "LCB":
        return orc.lcb_from_posterior(mu, sd, kappa)
    if a == "EI":
        return -orc.ei_from_posterior(mu, sd, y_opt, xi)
    return -orc.pi_from_posterior(mu, sd, y_opt, xi)


def check_against(st, res, mu_ref, sd_ref, acq_ref, vtol, y_opt, xi=0.01, kappa=1.96, check_argmin=True, k=10):
    assert np.max(np.abs(res["mean"] - mu_ref)) <= 1e-9 * max(np.max(np.abs(mu_ref)), st.y_std)
    var_scale = st.prior_var * st.y_std ** 2
    assert np.max(np.abs(res["std"] ** 2 - sd_ref ** 2)) <= vtol * var_scale
    for a in ACQS:
        v, ref = res[a]["values"], acq_ref[a]
        tol = 1e-9 * np.max(np.abs(ref))
        if vtol > 1e-9:
            # ill-conditioned K: the reference itself moves by eps*cond(K) on the variance when only its summation
            # order changes.  Propagate that admissible variance jitter through the acquisition.
            dvar = vtol * var_scale
            lo = np.sqrt(np.maximum(sd_ref ** 2 - dvar, 0.0))
            hi = np.sqrt(sd_ref ** 2 + dvar)
            spread = np.abs(_acq_from_posterior(a, mu_ref, hi, y_opt, xi, kappa) -
                            _acq_from_posterior(a, mu_ref, lo, y_opt, xi, kappa))
            tol = tol + 2.0 * spread
        assert np.all(np.abs(v - ref) <= tol), a
        # the device top-k must be exactly the (value, index)-sorted prefix of the device values
        order = np.lexsort((np.arange(v.size), v))[:k]
        assert np.array_equal(res[a]["topk_idx"], order), a
        assert np.array_equal(res[a]["topk_val"], v[order]), a
        if check_argmin:
            assert res[a]["topk_idx"][0] == int(np.argmin(ref)), a


@pytest.mark.parametrize("name", NAMES)
def test_golden_fixture(name):
    """Fitted state exactly as the reference produced it -> device -> compare with the reference's outputs."""
    rec, meta, est = load(name)
    _, _, ost = load_state(name)
    dev = engine.DeviceState.from_estimator(est)
    res = dev.score(rec["Xc"], y_opt=float(rec["y_opt"]), acq_funcs=ACQS, xi=float(rec["xi"]),
                    kappa=float(rec["kappa"]), top_k=10)
    vtol = var_tol(ost)
    acq_ref = {a: rec["acq_" + a] for a in ACQS}
    check_against(ost, res, rec["mean"], rec["std"], acq_ref, vtol, float(rec["y_opt"]), float(rec["xi"]),
                  float(rec["kappa"]), check_argmin=vtol <= 1e-9)
    if vtol <= 1e-9:
        for a in ACQS:
            assert res[a]["topk_idx"][0] == int(rec["argmin_" + a])
    assert np.max(np.abs(dev.predict_mean(rec["Xc"]) - rec["mean_only"])) <= 1e-9 * max(np.max(np.abs(rec["mean"])), ost.y_std)
    dev.close()


@pytest.mark.parametrize("n,d,m,kind,nu", [
    (300, 7, 5000, orc.KIND_MATERN, 2.5),      # n not a multiple of the 128-row block, m not a multiple of 128
    (129, 3, 1000, orc.KIND_MATERN, 1.5),
    (64, 1, 257, orc.KIND_MATERN, 0.5),
    (512, 20, 4096, orc.KIND_RBF, np.inf),
    (1024, 10, 3000, orc.KIND_MATERN, 2.5),
    (37, 50, 200, orc.KIND_MATERN, 2.5),
    (300, 65, 700, orc.KIND_MATERN, 2.5),      # more than 64 dimensions: one K1 CTA per SM
    (260, 100, 900, orc.KIND_RBF, np.inf),     # the largest n_dims the shared-memory tiles hold
    (150, 97, 300, orc.KIND_MATERN, 0.5),
])
def test_synthetic_vs_oracle(n, d, m, kind, nu):
    st = orc.make_synthetic_state(n, d, noise=1e-3, seed=n + d, kind=kind, nu=nu)
    rng = np.random.RandomState(7)
    Xc = rng.rand(m, d)
    Xc[3] = st.X_train[0]
    Xc[10] = Xc[9]
    y_opt = float(st.meta["y"].min())
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    res = dev.score(Xc, y_opt=y_opt, acq_funcs=ACQS, xi=0.01, kappa=1.96, top_k=10)
    mu, sd = orc.predict(st, Xc, return_std=True, quad="gemm")
    acq_ref = {a: orc.gaussian_acquisition(Xc, st, y_opt, a, acq_func_kwargs=dict(xi=0.01, kappa=1.96)) for a in ACQS}
    check_against(st, res, mu, sd, acq_ref, var_tol(st), y_opt)
    # duplicated candidate rows give bit-identical results wherever they sit
    for key in ("mean", "std"):
        assert res[key][10] == res[key][9]
    dev.close()


def test_chunked_launches_match_single_launch():
    """A small scratch limit forces several K1/K2 chunks; results must be bit-identical to the one-chunk run."""
    st = orc.make_synthetic_state(256, 6, noise=1e-3, seed=3)
    Xc = np.random.RandomState(1).rand(3000, 6)
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    a = dev.score(Xc, y_opt=0.0, acq_funcs=ACQS, top_k=5)
    dev.set_scratch_limit(4 * 256 * 128 * 8)       # 4 tiles per chunk
    b = dev.score(Xc, y_opt=0.0, acq_funcs=ACQS, top_k=5)
    assert np.array_equal(a["mean"], b["mean"]) and np.array_equal(a["std"], b["std"])
    for acq in ACQS:
        assert np.array_equal(a[acq]["values"], b[acq]["values"])
        assert np.array_equal(a[acq]["topk_idx"], b[acq]["topk_idx"])
    dev.close()


def test_kappa_inf_and_index_offset():
    st = orc.make_synthetic_state(100, 4, noise=1e-3, seed=5)
    Xc = np.random.RandomState(2).rand(500, 4)
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    res = dev.score(Xc, acq_funcs=("LCB",), kappa="inf", top_k=3, index_offset=1000)
    assert np.array_equal(res["LCB"]["values"], -res["std"])
    assert res["LCB"]["topk_idx"][0] == 1000 + int(np.argmin(-res["std"]))
    dev.close()


def test_torch_device_entry_point_matches_host_entry_point():
    import torch
    st = orc.make_synthetic_state(200, 5, noise=1e-3, seed=9)
    Xc = np.random.RandomState(4).rand(2000, 5)
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    ref = dev.score(Xc, y_opt=-1.0, acq_funcs=ACQS, top_k=4)
    Xt = torch.from_numpy(Xc).cuda()
    res = dev.score_tensors(Xt, y_opt=-1.0, acq_funcs=ACQS, top_k=4, want_posterior=True, want_values=True)
    torch.cuda.synchronize()
    assert np.array_equal(res["mean"].cpu().numpy(), ref["mean"])
    assert np.array_equal(res["std"].cpu().numpy(), ref["std"])
    for a in ACQS:
        assert np.array_equal(res[a]["values"].cpu().numpy(), ref[a]["values"])
        assert np.array_equal(res[a]["topk_idx"].cpu().numpy(), ref[a]["topk_idx"])
    # top-k only (no value vectors requested) goes through the internal scratch
    res2 = dev.score_tensors(Xt, y_opt=-1.0, acq_funcs=("EI",), top_k=4)
    torch.cuda.synchronize()
    assert np.array_equal(res2["EI"]["topk_idx"].cpu().numpy(), ref["EI"]["topk_idx"])
    dev.close()


# ----------------------------------------------------------------------------------------------------------------
# gradient path
# ----------------------------------------------------------------------------------------------------------------
def check_grads(st, g, ref_mu, ref_sd, ref_gm, ref_gs, acq_ref, vtol, floor=1e-12):
    """Gate (SURVEY 8d iv): |dg| <= 1e-9 * max|g| per candidate; sigma*grad(sigma) on the variance scale when K is
    ill-conditioned (same reasoning as tests/test_oracle_golden.py)."""
    n = ref_mu.shape[0]
    assert np.max(np.abs(g["mean"][:n] - ref_mu)) <= 1e-9 * max(np.max(np.abs(ref_mu)), st.y_std)
    var_scale = st.prior_var * st.y_std ** 2
    assert np.max(np.abs(g["std"][:n] ** 2 - ref_sd ** 2)) <= vtol * var_scale
    assert np.max(np.abs(g["mean_grad"][:n] - ref_gm)) <= 1e-9 * np.max(np.abs(ref_gm))
    gv_scale = var_scale / np.min(st.length_scale)
    assert np.max(np.abs((g["std_grad"][:n] - ref_gs) * ref_sd[:, None])) <= 10 * vtol * gv_scale
    if vtol <= 1e-9:
        for i in range(n):
            rel_var = (ref_sd[i] / st.y_std) ** 2 / st.prior_var
            # grad(sigma) = -(w . G) / sigma: an absolute error `floor` (relative to the prior) on the variance-scale
            # quantity is amplified by prior_var / sigma^2 near training points
            tol = (1e-9 + floor / max(rel_var, 1e-300)) * max(np.max(np.abs(ref_gs[i])), 1e-300)
            assert np.max(np.abs(g["std_grad"][i] - ref_gs[i])) <= tol, i
    for a, (v_ref, g_ref, z2) in acq_ref.items():
        for i in range(n):
            rel_var = (ref_sd[i] / st.y_std) ** 2 / st.prior_var
            if rel_var < 1e4 * vtol and vtol > 1e-9:
                continue
            ptol = (1e-9 + max(vtol, floor) / max(rel_var, 1e-300)) * (1.0 + z2[i])
            assert abs(g[a]["values"][i] - v_ref[i]) <= ptol * abs(v_ref[i]) + 1e-300, (a, i)
            assert np.max(np.abs(g[a]["grad"][i] - g_ref[i])) <= ptol * np.max(np.abs(g_ref[i])) + 1e-300, (a, i)


@pytest.mark.parametrize("name", NAMES)
def test_golden_gradients(name):
    rec, meta, est = load(name)
    _, _, ost = load_state(name)
    ng = meta["n_grad"]
    if ng == 0:
        pytest.skip("the reference itself raises for this kernel's gradient")
    dev = engine.DeviceState.from_estimator(est)
    y_opt, xi, kappa = float(rec["y_opt"]), float(rec["xi"]), float(rec["kappa"])
    g = dev.score_grad(rec["Xc"], y_opt=y_opt, acq_funcs=ACQS, xi=xi, kappa=kappa)      # all rows in one launch
    z2 = ((y_opt - xi - rec["mean"][:ng]) / np.maximum(rec["std"][:ng], 1e-300)) ** 2
    acq_ref = {a: (rec["acq_1d_" + a], rec["acq_grad_" + a], z2) for a in ACQS}
    check_grads(ost, g, rec["mean"][:ng], rec["std"][:ng], rec["mean_grad"], rec["std_grad"], acq_ref, var_tol(ost))
    # a one-row call returns bit-identical numbers to the same row inside the batch
    g1 = dev.score_grad(rec["Xc"][3:4], y_opt=y_opt, acq_funcs=ACQS, xi=xi, kappa=kappa)
    assert np.array_equal(g1["mean_grad"][0], g["mean_grad"][3]) and np.array_equal(g1["std_grad"][0], g["std_grad"][3])
    assert np.array_equal(g1["EI"]["grad"][0], g["EI"]["grad"][3])
    dev.close()


@pytest.mark.parametrize("n,d,m,kind,nu", [
    (300, 7, 200, orc.KIND_MATERN, 2.5),
    (129, 3, 130, orc.KIND_MATERN, 1.5),
    (200, 50, 100, orc.KIND_MATERN, 2.5),
    (256, 20, 150, orc.KIND_RBF, np.inf),
    (90, 4, 64, orc.KIND_MATERN, 0.5),
    (200, 65, 130, orc.KIND_MATERN, 2.5),      # gradient columns 64 .. d - 1 come from a second launch (frag0 = 8)
    (140, 100, 70, orc.KIND_MATERN, 1.5),
    (130, 96, 40, orc.KIND_RBF, np.inf),
    (70, 89, 33, orc.KIND_MATERN, 0.5),
])
def test_synthetic_gradients_vs_oracle_loop(n, d, m, kind, nu):
    """Batched device gradients against a loop of single-point oracle calls (the reference is 1-point only)."""
    st = orc.make_synthetic_state(n, d, noise=1e-3, seed=n + d, kind=kind, nu=nu)
    rng = np.random.RandomState(11)
    Xc = rng.rand(m, d)
    if not (kind == orc.KIND_MATERN and nu == 0.5):
        Xc[3] = st.X_train[0]                 # r = 0 corner (finite gradients, test_kernels.py:87-99)
    y_opt = float(st.meta["y"].min())
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    g = dev.score_grad(Xc, y_opt=y_opt, acq_funcs=ACQS)
    mu, sd, gm, gs = orc.predict_with_grads_batched(st, Xc)
    acq_ref = {}
    z2 = ((y_opt - 0.01 - mu) / np.maximum(sd, 1e-300)) ** 2
    for a in ACQS:
        vs, grads = np.empty(m), np.empty((m, d))
        for i in range(m):
            v, gr = orc.gaussian_acquisition_1D(Xc[i], st, y_opt, a, dict(xi=0.01, kappa=1.96))
            vs[i], grads[i] = v[0], gr
        acq_ref[a] = (vs, grads, z2)
    assert np.all(np.isfinite(g["mean_grad"])) and np.all(np.isfinite(g["std_grad"]))
    check_grads(st, g, mu, sd, gm, gs, acq_ref, var_tol(st))
    dev.close()


def test_gradient_matches_finite_differences():
    """The reference's own style of check (test_gpr.py:65-98, test_acquisition.py:98-132): analytic vs numeric."""
    st = orc.make_synthetic_state(60, 3, noise=1e-2, seed=21)
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    x = np.array([[0.31, 0.62, 0.47]])
    g = dev.score_grad(x, y_opt=-0.5, acq_funcs=ACQS)
    eps = 1e-6
    Xp = np.repeat(x, 6, axis=0)
    for k in range(3):
        Xp[2 * k, k] += eps
        Xp[2 * k + 1, k] -= eps
    f = dev.score(Xp, y_opt=-0.5, acq_funcs=ACQS)
    for k in range(3):
        assert abs((f["mean"][2 * k] - f["mean"][2 * k + 1]) / (2 * eps) - g["mean_grad"][0, k]) < 1e-5
        assert abs((f["std"][2 * k] - f["std"][2 * k + 1]) / (2 * eps) - g["std_grad"][0, k]) < 1e-5
        for a in ACQS:
            fd = (f[a]["values"][2 * k] - f[a]["values"][2 * k + 1]) / (2 * eps)
            assert abs(fd - g[a]["grad"][0, k]) < 1e-5, (a, k)
    dev.close()


# ----------------------------------------------------------------------------------------------------------------
# edge cases (ragged / minimal / maximal sizes, NaN, ties, degenerate posteriors)
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,m", [(1, 1, 1), (2, 3, 127), (128, 2, 128), (129, 5, 129), (5, 64, 3), (257, 1, 1000)])
def test_ragged_and_minimal_shapes(n, d, m):
    st = orc.make_synthetic_state(n, d, noise=1e-2, seed=n * 7 + d, normalize_y=n > 1)
    Xc = np.random.RandomState(m).rand(m, d)
    y_opt = float(st.meta["y"].min())
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    k = min(5, m + 2)                                   # top_k may exceed m: the tail is padded with (nan, -1)
    res = dev.score(Xc, y_opt=y_opt, acq_funcs=ACQS, top_k=k)
    mu, sd = orc.predict(st, Xc, return_std=True)
    acq_ref = {a: orc.gaussian_acquisition(Xc, st, y_opt, a, acq_func_kwargs=dict(xi=0.01, kappa=1.96)) for a in ACQS}
    assert np.max(np.abs(res["mean"] - mu)) <= 1e-9 * max(np.max(np.abs(mu)), st.y_std)
    assert np.max(np.abs(res["std"] ** 2 - sd ** 2)) <= var_tol(st) * st.prior_var * st.y_std ** 2
    z2 = ((y_opt - 0.01 - mu) / np.maximum(sd, 1e-300)) ** 2
    for a in ACQS:
        # normwise 1e-9, plus the pointwise (1 + z^2) allowance for deep-tail EI / PI values (SURVEY 8d iii)
        tol = 1e-9 * np.max(np.abs(acq_ref[a])) + 1e-9 * (1.0 + z2) * np.abs(acq_ref[a])
        assert np.all(np.abs(res[a]["values"] - acq_ref[a]) <= tol), a
        order = np.lexsort((np.arange(m), res[a]["values"]))[:k]
        assert np.array_equal(res[a]["topk_idx"][:len(order)], order)
        assert np.all(res[a]["topk_idx"][len(order):] == -1) and np.all(np.isnan(res[a]["topk_val"][len(order):]))
    g = dev.score_grad(Xc[: min(m, 40)], y_opt=y_opt, acq_funcs=ACQS)
    _, _, gm, gs = orc.predict_with_grads_batched(st, Xc[: min(m, 40)])
    assert np.max(np.abs(g["mean_grad"] - gm)) <= 1e-9 * max(np.max(np.abs(gm)), 1e-300)
    assert np.max(np.abs((g["std_grad"] - gs) * sd[: min(m, 40), None])) <= \
        10 * var_tol(st) * st.prior_var * st.y_std ** 2 / np.min(st.length_scale)
    dev.close()


def test_empty_candidate_set_and_bad_arguments():
    st = orc.make_synthetic_state(10, 2, noise=1e-2, seed=1)
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    res = dev.score(np.empty((0, 2)), y_opt=0.0, acq_funcs=("EI",), top_k=0)
    assert res["mean"].shape == (0,) and res["EI"]["values"].shape == (0,)
    with pytest.raises(ValueError):
        dev.score(np.zeros((3, 5)), acq_funcs=("EI",))              # wrong number of features
    with pytest.raises(ValueError):
        dev.score(np.zeros(3), acq_funcs=("EI",))                   # not 2-D
    with pytest.raises(ValueError):
        dev.score(np.zeros((3, 2)), acq_funcs=("UCB",))             # unknown acquisition
    with pytest.raises(skopt_b200._lib.SkbError):
        engine.DeviceState(engine.HostState(np.zeros((4, 101)), np.zeros(4), np.eye(4), np.ones(101), 3, 1, 0, 0, 0, 1))
    dev.close()


def test_nan_candidates_follow_numpy_argmin_and_argsort_rules():
    st = orc.make_synthetic_state(50, 3, noise=1e-2, seed=2)
    Xc = np.random.RandomState(0).rand(400, 3)
    Xc[77, 1] = np.nan
    Xc[300, 0] = np.nan
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    res = dev.score(Xc, y_opt=0.0, acq_funcs=("LCB",), top_k=4)
    v = res["LCB"]["values"]
    assert np.isnan(v[77]) and np.isnan(v[300]) and np.isfinite(np.delete(v, [77, 300])).all()
    assert res["LCB"]["first_nan"] == 77 == int(np.argmin(v))                       # np.argmin: first NaN wins
    assert np.array_equal(res["LCB"]["topk_idx"], np.argsort(v, kind="stable")[:4])  # np.argsort: NaN last
    dev.close()


def test_zero_variance_posterior_gives_zero_ei_and_pi():
    """Noise-free GP evaluated AT a training point: variance cancels to ~0 and may be clamped; EI/PI must follow the
    std > 0 mask of acquisition.py:213,296 (value exactly 0 -> -0.0 after the sign flip) without NaN."""
    st = orc.make_synthetic_state(30, 2, noise=1e-12, seed=3)
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    res = dev.score(st.X_train[:10], y_opt=float(st.meta["y"].min()), acq_funcs=ACQS)
    assert np.all(np.isfinite(res["std"])) and np.all(res["std"] >= 0)
    for a in ACQS:
        assert np.all(np.isfinite(res[a]["values"]))
    z = res["std"] == 0
    assert np.all(res["EI"]["values"][z] == 0) and np.all(res["PI"]["values"][z] == 0)
    assert res["clamped"] >= 0
    dev.close()


# ----------------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: spot checks against the oracle + size-independent properties
# ----------------------------------------------------------------------------------------------------------------
def test_full_size_c3_value_path():
    """configs[2]: n_train=2048, d=20, 1,000,000 candidates.  (a) a random sample of 1,500 candidates against the
    oracle at the 1e-9 gate; (b) the device top-k equals the (value, index) sort of the returned values, i.e. the
    argmin is np.argmin; (c) duplicated rows anywhere in the million give bit-equal results; (d) splitting the work
    into different scratch chunks does not change a single bit."""
    n, d, m = 2048, 20, 1_000_000
    st = orc.make_synthetic_state(n, d, noise=1e-4, seed=0)
    rng = np.random.RandomState(123)
    Xc = rng.rand(m, d)
    Xc[999_999] = Xc[17]
    Xc[500_000] = Xc[17]
    y_opt = float(st.meta["y"].min())
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    res = dev.score(Xc, y_opt=y_opt, acq_funcs=ACQS, top_k=20)
    sample = rng.choice(m, 1500, replace=False)
    mu, sd = orc.predict(st, Xc[sample], return_std=True)
    assert np.max(np.abs(res["mean"][sample] - mu)) <= 1e-9 * max(np.max(np.abs(mu)), st.y_std)
    assert np.max(np.abs(res["std"][sample] ** 2 - sd ** 2)) <= 1e-9 * st.prior_var * st.y_std ** 2
    assert np.max(np.abs(res["std"][sample] - sd) / sd) <= 1e-9
    for a in ACQS:
        ref = orc.gaussian_acquisition(Xc[sample], st, y_opt, a, acq_func_kwargs=dict(xi=0.01, kappa=1.96))
        z2 = ((y_opt - 0.01 - mu) / sd) ** 2
        assert np.max(np.abs(res[a]["values"][sample] - ref)) <= 1e-9 * np.max(np.abs(ref))
        assert np.all(np.abs(res[a]["values"][sample] - ref) <= 1e-9 * (1 + z2) * np.abs(ref) + 1e-300)
        v = res[a]["values"]
        order = np.lexsort((np.arange(m), v))[:20]
        assert np.array_equal(res[a]["topk_idx"], order) and res[a]["topk_idx"][0] == int(np.argmin(v))
        assert v[17] == v[500_000] == v[999_999]
    dev.set_scratch_limit(300 * 2048 * 128 * 8)          # 296 tiles per chunk instead of 888
    res2 = dev.score(Xc, y_opt=y_opt, acq_funcs=("EI",), top_k=20)
    assert np.array_equal(res2["EI"]["values"], res["EI"]["values"]) and np.array_equal(res2["std"], res["std"])
    assert np.array_equal(res2["EI"]["topk_idx"], res["EI"]["topk_idx"])
    dev.close()


def test_full_size_c4_gradient_path():
    """configs[3] sizes: n_train=4096, d=50 with gradients, a slice of 20,000 candidates: spot check of batched
    gradients against single-point oracle calls, and value-path / gradient-path consistency (two formulations of the
    same posterior: triangular L_inv contraction vs dense K_inv contraction)."""
    n, d, m = 4096, 50, 20_000
    st = orc.make_synthetic_state(n, d, noise=1e-4, seed=0)
    rng = np.random.RandomState(5)
    Xc = rng.rand(m, d)
    y_opt = float(st.meta["y"].min())
    dev = engine.DeviceState(host_state_from_oracle_state(st))
    g = dev.score_grad(Xc, y_opt=y_opt, acq_funcs=("EI", "LCB"))
    v = dev.score(Xc, y_opt=y_opt, acq_funcs=("EI", "LCB"))
    assert np.max(np.abs(g["mean"] - v["mean"])) <= 1e-12 * max(np.max(np.abs(v["mean"])), st.y_std)
    assert np.max(np.abs(g["std"] ** 2 - v["std"] ** 2)) <= 1e-9 * st.prior_var * st.y_std ** 2
    idx = rng.choice(m, 4, replace=False)
    mu, sd, gm, gs = orc.predict_with_grads_batched(st, Xc[idx])
    assert np.max(np.abs(g["mean_grad"][idx] - gm)) <= 1e-9 * np.max(np.abs(gm))
    assert np.max(np.abs(g["std_grad"][idx] - gs)) <= 1e-9 * np.max(np.abs(gs))
    for a in ("EI", "LCB"):
        for i in idx:
            val, grad = orc.gaussian_acquisition_1D(Xc[i], st, y_opt, a, dict(xi=0.01, kappa=1.96))
            z2 = ((y_opt - 0.01 - g["mean"][i]) / g["std"][i]) ** 2
            assert abs(g[a]["values"][i] - val[0]) <= 1e-9 * (1 + z2) * abs(val[0]) + 1e-300
            assert np.max(np.abs(g[a]["grad"][i] - grad)) <= 1e-9 * (1 + z2) * np.max(np.abs(grad)) + 1e-300
    dev.close()
