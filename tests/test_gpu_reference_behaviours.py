"""The behaviours the reference's own test-suite checks on this path, restated against the B200 build
(reference files cited per test; the reference itself is not importable on the GPU box).  These are the
3-6 decimal / convergence / determinism level checks of SURVEY.md section 4; the 1e-9 parity lives in
test_gpu_parity.py / test_gpu_split.py / test_gpu_optimizer.py."""
import warnings

import numpy as np
import pytest
from scipy import optimize

from skopt_b200 import Optimizer, gp_minimize
from skopt_b200.learning import GaussianProcessRegressor
from skopt_b200.learning.gaussian_process.kernels import RBF, WhiteKernel
from skopt_b200.space import Categorical, Integer, Real
from skopt_b200.utils import cook_estimator

pytestmark = pytest.mark.gpu


def bench1(x):
    return x[0] ** 2


def bench2(x):
    return x[0] ** 2 if x[0] < 0 else (x[0] - 5) ** 2 - 5


def bench3(x):
    return np.sin(5 * x[0]) * (1 - np.tanh(x[0] ** 2))


def bench4(x):
    return float(x[0]) ** 2


def branin(x, a=1, b=5.1 / (4 * np.pi ** 2), c=5. / np.pi, r=6, s=10, t=1. / (8 * np.pi)):
    return a * (x[1] - b * x[0] ** 2 + c * x[0] - r) ** 2 + s * (1 - t) * np.cos(x[0]) + s


# ---------------------------------------------------------------- skopt/tests/test_gp_opt.py:15-70
@pytest.mark.parametrize("func,y_opt,bounds,margin", [
    (bench1, 0.0, [(-2.0, 2.0)], 0.05), (bench2, -5.0, [(-6.0, 6.0)], 0.05), (bench3, -0.9, [(-2.0, 2.0)], 0.05)])
@pytest.mark.parametrize("search", ["sampling", "lbfgs"])
@pytest.mark.parametrize("acq", ["LCB", "EI"])
def test_minimiser_reaches_the_reference_thresholds(func, y_opt, bounds, margin, search, acq):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = gp_minimize(func, bounds, acq_optimizer=search, acq_func=acq, n_initial_points=10, n_calls=20,
                        random_state=1, noise=1e-10)
    assert r.fun < y_opt + margin


@pytest.mark.parametrize("acq", ["LCB", "EI"])
def test_minimiser_on_a_categorical_space(acq):
    """test_gp_opt.py:63-70 (bench4): one categorical dimension of numeric strings, Hamming-kernel GP."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = gp_minimize(bench4, [("-2", "-1", "0", "1", "2")], acq_optimizer="sampling", acq_func=acq,
                        n_initial_points=10, n_calls=20, random_state=1, noise=1e-10)
    assert r.fun < 0 + 1.05
    assert all(x[0] in ("-2", "-1", "0", "1", "2") for x in r.x_iters)


def test_n_jobs_does_not_change_the_trajectory():
    """test_gp_opt.py:74-81."""
    kw = dict(acq_optimizer="lbfgs", acq_func="EI", n_calls=4, n_initial_points=2, random_state=1, noise=1e-10)
    assert gp_minimize(bench3, [(-2.0, 2.0)], **kw).x_iters == gp_minimize(bench3, [(-2.0, 2.0)], n_jobs=2, **kw).x_iters


def test_default_arguments_run():
    """test_gp_opt.py:84-88."""
    gp_minimize(branin, ((-5.0, 10.0), (0.0, 15.0)), n_initial_points=2, n_calls=2)


@pytest.mark.parametrize("queue", [None, 1])
def test_a_given_estimator_is_used(queue):
    """test_gp_opt.py:91-117."""
    domain = [(1.0, 2.0), (3.0, 4.0)]
    est = cook_estimator("GP", domain, noise=1e+5)
    res = gp_minimize(branin, domain, n_calls=4, n_initial_points=2, base_estimator=est, noise=1e-10,
                      model_queue_size=queue)
    assert res["models"][-1].noise == 1e+5
    if queue:
        assert len(res["models"]) == queue


# ---------------------------------------------------------------- skopt/learning/gaussian_process/tests/test_gpr.py
def _toy():
    rng = np.random.RandomState(0)
    X = rng.randn(5, 5)
    return rng, X, rng.randn(5)


def test_noise_gaussian_matches_an_explicit_white_kernel():
    """test_gpr.py:50-61: same mean to 4 decimals; the std differs because noise="gaussian" is silenced at predict."""
    _, X, y = _toy()
    rbf, wk = RBF(length_scale=np.arange(1, 6)), WhiteKernel(1.0)
    gpr1 = GaussianProcessRegressor(rbf + wk).fit(X, y)
    gpr2 = GaussianProcessRegressor(rbf, noise="gaussian").fit(X, y)
    assert not gpr1.noise_ and gpr2.noise_
    assert abs(gpr1.kernel_.k2.noise_level - gpr2.noise_) < 1e-4
    mean1, std1 = gpr1.predict(X, return_std=True)
    mean2, std2 = gpr2.predict(X, return_std=True)
    assert np.max(np.abs(mean1 - mean2)) < 1e-4
    assert not np.any(std1 == std2)


def test_posterior_gradients_match_finite_differences():
    """test_gpr.py:65-98: analytic mean / std gradients against approx_fprime to 3 decimals, anisotropic RBF."""
    rng = np.random.RandomState(0)
    X, y, x_new = rng.randn(10, 5), rng.randn(10), rng.randn(5)
    gpr = GaussianProcessRegressor(RBF(length_scale=np.arange(1, 6), length_scale_bounds="fixed"), random_state=0).fit(X, y)
    mean, std, mean_grad, std_grad = gpr.predict(x_new[None, :], return_std=True, return_cov=False,
                                                 return_mean_grad=True, return_std_grad=True)
    for which, analytic in ((0, mean_grad), (1, std_grad)):
        numeric = optimize.approx_fprime(x_new, lambda x: gpr.predict(x[None, :], return_std=True)[which][0], 1e-4)
        assert np.max(np.abs(analytic - numeric)) < 1.5e-3


def test_duplicated_training_points_do_not_break_the_fit():
    """test_gpr.py:101-121: the default alpha keeps K factorisable when points coincide."""
    rng = np.random.RandomState(3)
    X, y = rng.rand(8, 3), rng.rand(8)
    X[:3, :] = 0.0
    y[:3] = 1.0
    model = GaussianProcessRegressor().fit(X, y)
    mu, sd = model.predict(X, return_std=True)
    assert np.all(np.isfinite(mu)) and np.all(np.isfinite(sd))


def test_return_cov_has_positive_diagonal_with_noise():
    """skopt/tests/test_gpr.py:8-17."""
    rng = np.random.RandomState(0)
    X, y = rng.rand(12, 2), rng.rand(12)
    gpr = GaussianProcessRegressor(RBF() + WhiteKernel(0.1)).fit(X, y)
    _, cov = gpr.predict(X, return_cov=True)
    assert np.all(np.diag(cov) > 0)


# ---------------------------------------------------------------- skopt/tests/test_optimizer.py
def test_repeated_ask_is_a_no_op_and_the_model_queue_is_bounded():
    """test_optimizer.py:29-65."""
    for queue, kept in ((None, 3), (2, 2)):
        opt = Optimizer([(-2.0, 2.0)], "GP", n_initial_points=1, acq_optimizer="sampling", model_queue_size=queue,
                        random_state=2)
        opt.run(bench1, n_iter=3)
        assert len(opt.models) == kept and len(opt.Xi) == 3
        opt.ask()
        assert len(opt.models) == kept and len(opt.Xi) == 3
        assert opt.ask() == opt.ask()
        opt.update_next()
        assert opt.ask() == opt.ask()


def test_tell_validates_points_and_values():
    """test_optimizer.py:68-174: value / point mismatches, bounds and dimension checks raise ValueError before any fit."""
    low, high = -2.0, 2.0
    opt = Optimizer([(low, high)], "GP", n_initial_points=1, acq_optimizer="sampling")
    with pytest.raises(ValueError):
        opt.tell([1.0], [1.0, 1.0])                # one point, several values
    with pytest.raises(ValueError):
        opt.tell([[1.0], [2.0]], [1.0, None])      # a non-scalar value in a batch
    for x, y in (([high + 0.5], 2.0), ([low - 0.5], 2.0), ([high + 0.5, high], (2.0, 3.0)), ([low - 0.5, high], (2.0, 3.0))):
        with pytest.raises(ValueError):
            opt.tell(x, y)
    with pytest.raises(ValueError, match="Dimensions of point "):
        opt.tell([low + 1, low + 1], 2.0)          # within bounds, one coordinate too many
    opt2 = Optimizer([(low, high), (low + 4, high + 4)], "GP", n_initial_points=1, acq_optimizer="sampling")
    for x in ([high + 0.5, high + 4.5], [low - 0.5, low - 4.5], [high + 0.5, high + 0.5], [low - 0.5, high + 0.5]):
        with pytest.raises(ValueError):
            opt2.tell(x, 2.0)
    for pts in ([(high + 0.5, high + 0.5)] * 2, [(low - 0.5, high + 0.5)] * 2):
        with pytest.raises(ValueError):
            opt2.tell(pts, [2.0, 3.0])
    opt3 = Optimizer([(-2, 2), (-2, 2)])
    for x in ([-1], [-1, -1, -1]):
        with pytest.raises(ValueError, match="Dimensions of point "):
            opt3.tell(x, 2.0)
    for pts in ([[-1], [-1, 0], [-1, 1]], [[-1, -1, -1], [-1, 0], [-1, 1]]):
        with pytest.raises(ValueError, match="dimensions as the space"):
            opt3.tell(pts, 2.0)
    assert opt.Xi == [] and opt2.Xi == [] and opt3.Xi == []


def test_result_object_and_copy():
    """test_optimizer.py:177-187, 223-252."""
    opt = Optimizer([(-2.0, 2.0)], "GP", n_initial_points=1, acq_optimizer="sampling", random_state=0)
    res = opt.tell([1.5], 2.0)
    assert res.x_iters == [[1.5]] and np.array_equal(res.func_vals, [2.0]) and res.x == [1.5] and res.fun == 2.0
    opt.run(bench1, n_iter=3)
    dup = opt.copy()
    assert dup.Xi == opt.Xi and dup.yi == opt.yi and type(dup.base_estimator_) is type(opt.base_estimator_)


def test_models_appear_once_the_initial_points_are_told():
    """test_optimizer.py:255-288."""
    opt = Optimizer([(-2.0, 2.0)], "GP", n_initial_points=2, acq_optimizer="sampling", random_state=1)
    x0, x1 = opt.ask(), opt.ask()
    assert x0 != x1
    assert len(opt.tell(x1, 3.0).models) == 0
    x2 = opt.ask()
    assert x1 != x2
    assert len(opt.tell(x2, 4.0).models) == 1
    x3, x4 = opt.ask(), opt.ask()
    assert x2 != x3 and x3 == x4                   # the first model-based point; asking again adds no information
    assert len(opt.tell(x3, 1.0).models) == 2


def test_estimator_strings():
    """test_optimizer.py:292-305 (GP family only on this path)."""
    with pytest.raises(ValueError):
        Optimizer([(-2.0, 2.0)], base_estimator="rtr", n_initial_points=1)
    for name in ("GP", "gp", "DUMMY", "dummy"):
        Optimizer([(-2.0, 2.0)], base_estimator=name, n_initial_points=1).run(bench1, n_iter=3)


def test_optimizer_defaults_reproduce_gp_minimize():
    """test_optimizer.py:314-335."""
    space = [(-5.0, 10.0), (0.0, 15.0)]
    opt = Optimizer(space, random_state=1)
    for _ in range(12):
        x = opt.ask()
        res_opt = opt.tell(x, branin(x))
    res_min = gp_minimize(branin, space, n_calls=12, random_state=1)
    assert res_min.space == res_opt.space
    assert np.allclose(res_min.x_iters, res_opt.x_iters) and np.allclose(res_min.x, res_opt.x)
    again = opt.get_result()
    assert np.allclose(res_min.x_iters, again.x_iters) and np.allclose(res_min.x, again.x)


def test_dimension_names_survive():
    """test_optimizer.py:339-356."""
    opt = Optimizer([Real(0, 1, name="real"), Categorical(["a", "b", "c"], name="cat"), Integer(0, 1, name="int")],
                    n_initial_points=2)
    result = opt.tell([(0.5, "a", 0.5)], [3])
    assert [d.name for d in result.space.dimensions] == ["real", "cat", "int"]


@pytest.mark.parametrize("cats", [[2, 3, 4, 5, 6, 7, 8, 9, 10, 11], [str(v) for v in range(2, 12)]])
def test_categorical_only_space(cats):
    """test_optimizer.py:359-381: numeric and string categories, constant-liar batch afterwards."""
    opt = Optimizer([Categorical(cats), Categorical(cats)])
    for n in range(15):
        res = opt.tell(opt.ask(), 12 * n)
    assert len(res.x_iters) == 15
    pts = opt.ask(n_points=4)
    assert len(pts) == 4 and all(p[0] in cats and p[1] in cats for p in pts)


def test_categorical_space_with_a_gradient_kernel():
    """test_optimizer.py:384-401: an explicit (RBF-default) GP on categorical dimensions keeps the lbfgs optimiser."""
    opt = Optimizer([Categorical([1, 2, 3]), Categorical([4, 5, 6])], base_estimator=GaussianProcessRegressor(alpha=1e-7),
                    acq_optimizer="lbfgs", n_initial_points=10, n_jobs=2)
    for _ in range(3):
        pts = opt.ask(n_points=4)
        assert len(pts) == 4
        opt.tell(pts, [float(np.linalg.norm(p)) for p in pts])


# ---------------------------------------------------------------- skopt/tests/test_parallel_cl.py
@pytest.mark.parametrize("strategy", ["cl_min", "cl_mean", "cl_max"])
def test_constant_liar_batches(strategy):
    """test_parallel_cl.py:34-164: batches run, their points differ, repeated asks return the same set, and a run is
    exactly reproducible."""
    def make():
        return Optimizer([Real(-5.0, 10.0), Real(0.0, 15.0)], base_estimator="GP", acq_optimizer="sampling",
                         n_initial_points=5, random_state=0)
    opt = make()
    for _ in range(8):
        pts = opt.ask(n_points=3, strategy=strategy)
        assert len(pts) == 3
        opt.tell(pts, [branin(p) for p in pts])
    pts = opt.ask(n_points=5, strategy=strategy)
    for i in range(5):
        for j in range(i + 1, 5):
            assert any(abs(a - b) > 1e-3 for a, b in zip(pts[i], pts[j]))
    assert opt.ask(n_points=5, strategy=strategy) == pts
    opt2 = make()
    for _ in range(8):
        p2 = opt2.ask(n_points=3, strategy=strategy)
        opt2.tell(p2, [branin(p) for p in p2])
    assert opt2.ask(n_points=5, strategy=strategy) == pts
    with pytest.raises(ValueError):
        opt.ask(n_points=2, strategy="cl_median")


@pytest.mark.parametrize("m, block", [(1, None), (5, None), (130, None), (300, None), (300, 64), (2500, None)])
def test_return_cov_on_device_matches_the_reference_formula(m, block, monkeypatch):
    """predict(return_cov=True) (gpr.py:315-325): the quadratic form K_trans K_inv K_trans^T is computed on the GPU
    (skb_predict_cov_host: K1, dense K_inv product, one more tile product); against cho_solve on the host at 1e-9 of the
    prior variance, symmetric to rounding, and the diagonal equal to return_std squared.  Point sets larger than two
    blocks (2 x 1024 points; `block` shrinks that for the test) are assembled from device calls on pairs of blocks."""
    from scipy.linalg import cho_solve
    from skopt_b200.learning import GaussianProcessRegressor
    if block:
        monkeypatch.setattr(GaussianProcessRegressor, "_COV_BLOCK", block)
    from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern
    rng = np.random.RandomState(m)
    n, d = 200, 4
    X = rng.rand(n, d)
    y = np.sin(X.dot(rng.randn(d))) + 0.05 * rng.randn(n)
    gpr = GaussianProcessRegressor(ConstantKernel(1.7, "fixed") * Matern(np.full(d, 0.7), "fixed", nu=1.5), normalize_y=True,
                                   noise=1e-3, optimizer=None).fit(X, y)
    Xq = rng.rand(m, d)
    Xq[0] = X[3]                                             # a training point: covariance row close to the noise level
    mean, cov = gpr.predict(Xq, return_cov=True)
    K_trans = gpr.kernel_(Xq, gpr.X_train_)
    ref_mean = gpr.y_train_std_ * K_trans.dot(gpr.alpha_) + gpr.y_train_mean_
    ref = (gpr.kernel_(Xq) - K_trans.dot(cho_solve((gpr.L_, True), K_trans.T))) * gpr.y_train_std_ ** 2
    scale = 1.7 * gpr.y_train_std_ ** 2
    assert cov.shape == (m, m)
    assert np.max(np.abs(mean - ref_mean)) <= 1e-9 * max(np.max(np.abs(ref_mean)), gpr.y_train_std_)
    assert np.max(np.abs(cov - ref)) <= 1e-9 * scale
    assert np.max(np.abs(cov - cov.T)) <= 1e-12 * scale
    _, std = gpr.predict(Xq, return_std=True)
    assert np.max(np.abs(np.diag(cov) - std ** 2)) <= 1e-9 * scale
