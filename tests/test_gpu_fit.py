"""Fit objective on the GPU (skb_fit_lml_host) against the values recorded from the unmodified reference estimator,
against this package's host path (scikit-learn) at a larger size, and through fit() / Optimizer."""
import numpy as np
import pytest

from tests._fitcases import NAMES, check_against_golden, load_case

pytestmark = pytest.mark.gpu


def _device_eval(case):
    from skopt_b200.engine import DeviceFit, FitLayout
    lay = FitLayout(case["kernel"], case["X"].shape[1])
    fit = DeviceFit(case["X"], case["y"], lay.kind, 1e-10)
    return fit, (lambda full: fit.lml(full, eval_gradient=True))


@pytest.mark.parametrize("name", NAMES)
def test_device_lml_matches_reference(name):
    case = load_case(name)
    fit, ev = _device_eval(case)
    try:
        # m52_nonoise has no noise term and alpha = 1e-10: cond(K) up to 1.5e8, so two correct Cholesky codes (LAPACK
        # in the reference, cuSOLVER here) differ by ~cond * eps in K^-1; every other case is held to 1e-9
        check_against_golden(case, ev, rtol_lml=1e-9, rtol_grad=1e-7 if name == "m52_nonoise" else 1e-9)
        # value-only entry gives the same value
        from skopt_b200.engine import FitLayout
        full = FitLayout(case["kernel"], case["X"].shape[1]).expand(case["thetas"][1])
        assert fit.lml(full, eval_gradient=False) == ev(full)[0]
    finally:
        fit.close()


def test_device_lml_not_positive_definite():
    from skopt_b200.engine import DeviceFit
    fit = DeviceFit(np.zeros((40, 2)), np.arange(40.0), 3, 0.0)
    v, g = fit.lml(np.array([0.0, 0.0, 0.0, -np.inf]))
    assert v == -np.inf and not g.any()
    fit.close()


@pytest.mark.parametrize("n,d,nu", [(700, 8, 2.5), (1100, 12, 1.5), (33, 1, 2.5), (260, 100, 2.5), (200, 97, 1.5)])
def test_estimator_lml_gpu_vs_host(n, d, nu):
    from skopt_b200.learning import GaussianProcessRegressor
    from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel
    rng = np.random.RandomState(n)
    X = rng.rand(n, d)
    y = np.sin(X.dot(rng.randn(d))) + 0.05 * rng.randn(n)
    kernel = ConstantKernel(1.0, (0.01, 1000.0)) * Matern(length_scale=np.ones(d), length_scale_bounds=[(0.01, 100)] * d,
                                                            nu=nu) + WhiteKernel(0.01, (1e-6, 10.0))
    gpr = GaussianProcessRegressor(kernel=kernel, normalize_y=True, optimizer=None).fit(X, y)
    gpr.kernel_ = kernel.clone_with_theta(kernel.theta)            # fit() silenced the White term of kernel_
    for shift in (0.0, 0.4):
        theta = kernel.theta + shift * rng.randn(kernel.theta.size)
        gpr._skb_fit_device = "host"
        v_h, g_h = gpr.log_marginal_likelihood(theta, eval_gradient=True)
        gpr._skb_fit_device = "gpu"
        v_g, g_g = gpr.log_marginal_likelihood(theta, eval_gradient=True)
        assert abs(v_g - v_h) <= 1e-9 * max(1.0, abs(v_h))
        assert np.max(np.abs(g_g - g_h)) <= 1e-9 * max(1.0, np.max(np.abs(g_h)))
        assert gpr.log_marginal_likelihood(theta) == v_g
    gpr._close_device_fit()


def test_fit_on_gpu_reaches_the_host_optimum():
    from skopt_b200.learning import GaussianProcessRegressor
    from skopt_b200.utils import cook_estimator
    rng = np.random.RandomState(4)
    X = rng.rand(160, 3)
    y = np.sin(3 * X[:, 0]) + X[:, 1] ** 2 + 0.05 * rng.randn(160)
    fits = {}
    for where in ("host", "gpu"):
        est = cook_estimator("GP", space=[(0.0, 1.0)] * 3, random_state=7)
        est._skb_fit_device = where
        est.fit(X, y)
        assert "_skb_device_fit" not in est.__dict__            # the training set left the GPU with the search
        fits[where] = est
    h, g = fits["host"], fits["gpu"]
    assert abs(g.log_marginal_likelihood_value_ - h.log_marginal_likelihood_value_) <= \
        1e-5 * abs(h.log_marginal_likelihood_value_)
    Xq = rng.rand(200, 3)
    mh, sh = h.predict(Xq, return_std=True)
    mg, sg = g.predict(Xq, return_std=True)
    assert np.max(np.abs(mg - mh)) <= 1e-3 * np.std(y) and np.max(np.abs(sg - sh)) <= 1e-3 * np.std(y)


def test_optimizer_with_gpu_fit():
    from skopt_b200 import Optimizer
    runs = {}
    for where in ("host", "gpu"):
        opt = Optimizer([(-2.0, 2.0), (-2.0, 2.0)], "GP", n_initial_points=6, acq_func="EI", acq_optimizer="sampling",
                        acq_optimizer_kwargs=dict(n_points=2000, fit_device=where), random_state=3)
        for _ in range(12):
            x = opt.ask()
            opt.tell(x, float((x[0] - 0.5) ** 2 + (x[1] + 0.3) ** 2))
        runs[where] = opt
    assert runs["gpu"].models[-1]._skb_fit_device == "gpu"
    # same initial design (RNG order untouched); later points may differ in the last digits of the hyper-parameters
    assert runs["gpu"].Xi[:6] == runs["host"].Xi[:6]
    assert min(runs["gpu"].yi) <= 0.5
    with pytest.raises(ValueError):
        Optimizer([(-2.0, 2.0)], "GP", acq_optimizer_kwargs=dict(fit_device="tpu"))


def test_batched_evaluations_are_bit_identical_to_single_ones():
    from skopt_b200.engine import DeviceFit
    rng = np.random.RandomState(0)
    n, d = 300, 6
    X = rng.rand(n, d)
    y = np.sin(X.dot(rng.randn(d)))
    y = (y - y.mean()) / y.std()
    fit = DeviceFit(X, y, 3, 1e-10)
    thetas = np.concatenate([rng.randn(11, 1), 0.5 * rng.randn(11, d), np.log(0.01) + rng.randn(11, 1)], axis=1)
    thetas[4, -1] = -np.inf                                  # no noise, alpha only
    values, grads = fit.lml_batch(thetas)                    # 11 > 8 slots: two groups
    for i, th in enumerate(thetas):
        v, g = fit.lml(th)
        assert v == values[i]
        np.testing.assert_array_equal(g, grads[i])
    fit.close()


def _fixed_kernel_problem(n, d, nu, seed):
    from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n)
    kernel = ConstantKernel(1.3, "fixed") * Matern(length_scale=np.full(d, 0.9), length_scale_bounds="fixed", nu=nu)
    return X, y, kernel, rng


# sizes around the 64-column blocks of the library's own Cholesky / recursive-doubling inverse (csrc/chol.cuh): one block,
# partial blocks, 2 / 3 / 5 blocks (odd block counts leave a pair without its second half at some level)
@pytest.mark.parametrize("n,d", [(1, 1), (2, 1), (33, 2), (63, 2), (64, 2), (65, 2), (127, 3), (128, 3), (129, 3), (200, 3),
                                 (300, 5), (321, 4), (1100, 12), (150, 100)])
def test_finalize_matches_host_factorisation(n, d):
    """skb_fit_finalize_host: L_, alpha_ and the lml at theta* against SciPy's Cholesky of the same K."""
    from scipy.linalg import cho_solve, cholesky
    from skopt_b200.engine import DeviceFit, FitLayout
    from skopt_b200.learning.gaussian_process.kernels import WhiteKernel
    X, y, kernel, _ = _fixed_kernel_problem(n, d, 2.5, n)
    y = (y - y.mean()) / (y.std() if y.std() > 0 else 1.0)
    full_kernel = kernel + WhiteKernel(1e-3, "fixed")
    lay = FitLayout(full_kernel, d)
    fit = DeviceFit(X, y, lay.kind, 1e-10)
    try:
        theta_full = lay.expand(full_kernel.theta)
        lml, alpha, L = fit.finalize(theta_full)
        assert lml == fit.lml(theta_full, eval_gradient=False)
        K = full_kernel(X)
        K[np.diag_indices_from(K)] += 1e-10
        L_ref = cholesky(K, lower=True)
        cond = np.linalg.cond(K)
        assert not np.triu(L, 1).any()
        assert np.max(np.abs(L - L_ref)) <= 1e-15 * cond * np.max(np.abs(L_ref)) + 1e-12
        a_ref = cho_solve((L_ref, True), y)
        assert np.max(np.abs(alpha - a_ref)) <= 1e-15 * cond * np.max(np.abs(a_ref)) + 1e-12
        lml2, alpha2, none = fit.finalize(theta_full, want_factor=False)
        assert none is None and lml2 == lml
        np.testing.assert_array_equal(alpha2, alpha)
    finally:
        fit.close()


def test_finalize_not_positive_definite_raises():
    from skopt_b200.engine import DeviceFit
    fit = DeviceFit(np.zeros((40, 2)), np.arange(40.0), 3, 0.0)
    with pytest.raises(np.linalg.LinAlgError):
        fit.finalize(np.array([0.0, 0.0, 0.0, -np.inf]))
    from skopt_b200._lib import SkbError
    from skopt_b200.engine import StateScalars
    with pytest.raises(SkbError):                        # no factor, no state
        fit.make_state(StateScalars(40, 2, 1.0, 3, 1.0, 0.0, 0.0, 0.0, 1.0))
    fit.close()


@pytest.mark.parametrize("n,d,nu", [(33, 2, 2.5), (65, 2, 2.5), (129, 3, 1.5), (300, 5, 1.5), (1100, 12, 2.5)])
def test_state_built_on_device_scores_like_the_host_built_state(n, d, nu):
    """fit(optimizer=None) finalised on the GPU (skb_state_create_from_fit: trtri, potri, packing on the device) against
    the same model finalised on the host (SciPy factor, solve_triangular, upload): posterior, acquisition, arg-min and
    gradients; FP64 and split-integer contraction."""
    from skopt_b200.learning import GaussianProcessRegressor
    X, y, kernel, rng = _fixed_kernel_problem(n, d, nu, 7 * n)
    est = {}
    for where in ("host", "gpu"):
        e = GaussianProcessRegressor(kernel=kernel, normalize_y=True, noise=1e-3, optimizer=None)
        e._skb_fit_device = where
        est[where] = e.fit(X, y)
    h, g = est["host"], est["gpu"]
    assert "_skb_device_state" in g.__dict__ and "K_inv_" not in g.__dict__ and "K_inv_" in h.__dict__
    assert g.noise_ == h.noise_ and g.kernel_ == h.kernel_
    cond = np.linalg.cond(h.L_) ** 2
    assert np.max(np.abs(g.L_ - h.L_)) <= 1e-15 * cond * np.max(np.abs(h.L_)) + 1e-12
    assert np.max(np.abs(g.alpha_ - h.alpha_)) <= 1e-15 * cond * np.max(np.abs(h.alpha_)) + 1e-12
    assert abs(g.log_marginal_likelihood_value_ - h.log_marginal_likelihood_value_) <= 1e-9 * abs(h.log_marginal_likelihood_value_)
    m = 9000
    Xq = rng.rand(m, d)
    y_opt = float(y.min())
    prior = h.device_state().host.prior_var * float(h.y_train_std_) ** 2
    for precision in ("fp64", "split_int8"):
        rh = h.device_state().score(Xq, y_opt=y_opt, acq_funcs=("EI", "LCB"), top_k=4, precision=precision)
        rg = g.device_state().score(Xq, y_opt=y_opt, acq_funcs=("EI", "LCB"), top_k=4, precision=precision)
        assert np.max(np.abs(rg["mean"] - rh["mean"])) <= 1e-9 * max(np.max(np.abs(rh["mean"])), float(h.y_train_std_))
        assert np.max(np.abs(rg["std"] ** 2 - rh["std"] ** 2)) <= max(1e-9, 1e-15 * cond) * prior
        for a in ("EI", "LCB"):
            # relative to the largest value, with a floor at 1e-12 of the target scale (EI can be ~1e-8 everywhere)
            assert np.max(np.abs(rg[a]["values"] - rh[a]["values"])) <= (1e-9 * np.max(np.abs(rh[a]["values"])) +
                                                                          1e-12 * float(h.y_train_std_))
            assert rg[a]["topk_idx"][0] == rh[a]["topk_idx"][0]
    gh = h.device_state().score_grad(Xq[:200], y_opt=y_opt, acq_funcs=("EI",))
    gg = g.device_state().score_grad(Xq[:200], y_opt=y_opt, acq_funcs=("EI",))
    for key in ("mean_grad", "std_grad"):
        assert np.max(np.abs(gg[key] - gh[key])) <= 1e-8 * np.max(np.abs(gh[key]))
    assert np.max(np.abs(gg["EI"]["grad"] - gh["EI"]["grad"])) <= 1e-8 * np.max(np.abs(gh["EI"]["grad"]))
    # the host copies appear on demand and agree with the host-built ones
    assert np.max(np.abs(g.K_inv_ - h.K_inv_)) <= 1e-15 * cond * np.max(np.abs(h.K_inv_)) + 1e-12
    # a pickled model comes back without device handles and rebuilds its state from L_
    import pickle
    g2 = pickle.loads(pickle.dumps(g))
    assert "_skb_device_state" not in g2.__dict__
    np.testing.assert_allclose(g2.predict(Xq[:50]), g.predict(Xq[:50]), rtol=0, atol=1e-9 * float(h.y_train_std_))
