"""GaussianProcessRegressor.append_observation (the O(n^2) liar update behind Optimizer(liar_update="rank1")): the appended
factor, its inverse, alpha_ and the target normalisation against a from-scratch fit with the SAME hyper-parameters."""
import numpy as np
import pytest


def _problem(n, d, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(X.dot(rng.randn(d))) + 0.1 * rng.randn(n)
    return X, y, rng


@pytest.mark.parametrize("n,d,noise,normalize_y", [(40, 3, 1e-4, True), (130, 6, "gaussian", True), (65, 2, 1e-2, False)])
def test_append_equals_refit_with_frozen_hyper_parameters(n, d, noise, normalize_y):
    from skopt_b200.learning import GaussianProcessRegressor
    from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern
    X, y, rng = _problem(n, d)
    kernel = ConstantKernel(1.3, "fixed") * Matern(np.full(d, 0.8), "fixed", nu=2.5)
    est = GaussianProcessRegressor(kernel, normalize_y=normalize_y, noise=noise, optimizer=None, random_state=0).fit(X, y)
    cur, Xa, ya = est, X, y
    for step in range(3):                                    # three appends in a row (a constant-liar batch)
        x_new, y_new = rng.rand(d), float(ya.min())
        cur = cur.append_observation(x_new, y_new)
        Xa, ya = np.vstack([Xa, x_new]), np.append(ya, y_new)
    ref = GaussianProcessRegressor(kernel, normalize_y=normalize_y, noise=noise, optimizer=None, random_state=0).fit(Xa, ya)
    assert cur.X_train_.shape == ref.X_train_.shape and np.array_equal(cur.X_train_, ref.X_train_)
    np.testing.assert_allclose(cur.y_train_mean_, ref.y_train_mean_, rtol=1e-13, atol=0)
    np.testing.assert_allclose(cur.y_train_std_, ref.y_train_std_, rtol=1e-13, atol=0)
    np.testing.assert_allclose(cur.y_train_, ref.y_train_, rtol=1e-12, atol=1e-13)
    scale = np.max(np.abs(ref.L_))
    assert np.max(np.abs(cur.L_ - ref.L_)) <= 1e-9 * scale
    assert np.max(np.abs(cur._L_inv_lower - ref._L_inv_lower)) <= 1e-9 * np.max(np.abs(ref._L_inv_lower))
    assert np.max(np.abs(cur.alpha_ - ref.alpha_)) <= 1e-9 * np.max(np.abs(ref.alpha_))
    assert np.max(np.abs(cur.K_inv_ - ref.K_inv_)) <= 1e-9 * np.max(np.abs(ref.K_inv_))      # lazily from the appended inverse
    assert abs(cur.log_marginal_likelihood_value_ - ref.log_marginal_likelihood_value_) <= 1e-9 * abs(ref.log_marginal_likelihood_value_)
    # the original model is untouched
    assert est.X_train_.shape[0] == n and est.L_.shape == (n, n)


def test_append_rejects_bad_input():
    from skopt_b200.learning import GaussianProcessRegressor
    est = GaussianProcessRegressor(noise=1e-3, optimizer=None)
    with pytest.raises(ValueError):
        est.append_observation([0.1, 0.2], 0.0)              # not fitted
    X, y, _ = _problem(20, 2)
    est.fit(X, y)
    with pytest.raises(ValueError):
        est.append_observation([0.1, 0.2, 0.3], 0.0)         # wrong number of features


def test_optimizer_rejects_unknown_liar_update():
    from skopt_b200 import Optimizer
    with pytest.raises(ValueError):
        Optimizer([(0.0, 1.0)], "GP", acq_optimizer_kwargs={"liar_update": "woodbury"})
    assert Optimizer.LIAR_UPDATES == ("refit", "rank1")
