"""Fit objective: the oracle restatement of sklearn's log_marginal_likelihood against values
recorded from the unmodified reference estimator, and the theta layout logic.  CPU only."""
import numpy as np
import pytest

from oracle import gp_oracle as orc
from skopt_b200.engine import FitLayout
from skopt_b200.learning.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

from tests._fitcases import NAMES, check_against_golden, load_case


@pytest.mark.parametrize("name", NAMES)
def test_oracle_lml_matches_reference(name):
    case = load_case(name)
    check_against_golden(case, lambda full: orc.log_marginal_likelihood(case["X"], case["y"], case["nu"], full, 1e-10))


def test_oracle_lml_not_positive_definite():
    X = np.zeros((5, 2))                                   # five identical points, no noise, no jitter: singular K
    v, g = orc.log_marginal_likelihood(X, np.arange(5.0), 2.5, np.array([0.0, 0.0, 0.0, -np.inf]), alpha=0.0)
    assert v == -np.inf and not g.any()


def test_fit_layout_orders_and_fixed_parameters():
    d = 3
    k = ConstantKernel(2.0, (0.01, 1000)) * Matern(length_scale=[0.5, 1.0, 2.0], nu=2.5) + WhiteKernel(0.3)
    lay = FitLayout(k, d)
    assert lay.n_theta == 5 and lay.slots == [[0], [1], [2], [3], [4]]
    np.testing.assert_allclose(lay.expand(k.theta), np.log([2.0, 0.5, 1.0, 2.0, 0.3]))
    k = ConstantKernel(2.0, "fixed") * RBF(length_scale=0.7) + WhiteKernel(0.3, "fixed")
    lay = FitLayout(k, d)
    assert lay.kind == 0 and lay.slots == [[1, 2, 3]]
    np.testing.assert_allclose(lay.expand([0.1]), [np.log(2.0), 0.1, 0.1, 0.1, np.log(0.3)])
    np.testing.assert_allclose(lay.select(np.array([9.0, 1.0, 2.0, 3.0, 7.0])), [6.0])
    lay = FitLayout(Matern(length_scale=[0.5, 1.0, 2.0], nu=1.5), d)          # no amplitude, no noise
    assert lay.full[0] == 0.0 and lay.full[-1] == -np.inf and lay.n_theta == 3
    k = WhiteKernel(0.3) + Matern(length_scale=1.0, nu=0.5) * ConstantKernel(1.5)      # operand order reversed
    lay = FitLayout(k, d)
    assert lay.kind == 1 and lay.n_theta == 3
    np.testing.assert_allclose(lay.expand(k.theta), np.log([1.5, 1.0, 1.0, 1.0, 0.3]))
    with pytest.raises(NotImplementedError):
        FitLayout(k + ConstantKernel(1.0), d)
    with pytest.raises(NotImplementedError):
        FitLayout(RBF(1.0) * RBF(2.0), d)


class _OracleFit:
    """Stand-in for engine.DeviceFit (same interface) that evaluates the objective with the CPU oracle: lets the
    lock-step restart logic of GaussianProcessRegressor.fit run without a GPU."""

    def __init__(self, X, y, nu):
        self.X, self.y, self.nu, self.batches = X, y, nu, []

    def lml(self, theta_full, eval_gradient=True):
        v, g = orc.log_marginal_likelihood(self.X, self.y, self.nu, theta_full, 1e-10)
        return (v, g) if eval_gradient else v

    def lml_batch(self, thetas_full):
        self.batches.append(len(thetas_full))
        out = [self.lml(t) for t in thetas_full]
        return np.array([o[0] for o in out]), np.array([o[1] for o in out])

    def finalize(self, theta_full, want_factor=True):
        from scipy.linalg import cho_solve, cholesky
        amplitude, ls, noise = np.exp(theta_full[0]), np.exp(theta_full[1:-1]), np.exp(theta_full[-1])
        K = amplitude * Matern(length_scale=ls, nu=self.nu)(self.X)
        K[np.diag_indices_from(K)] = (amplitude + noise) + 1e-10
        L = cholesky(K, lower=True)
        self.finalized = theta_full.copy()
        return self.lml(theta_full, eval_gradient=False), cho_solve((L, True), self.y), L

    def make_state(self, scalars):
        self.scalars = scalars
        return None                      # no GPU in this test: the estimator then builds its state lazily

    def close(self):
        pass


def test_lockstep_restarts_equal_sequential_restarts(monkeypatch):
    """fit() with the device objective advances all L-BFGS-B restarts in lock step; the result must be bit-identical
    to scikit-learn's one-after-the-other loop over the same objective, and the estimator's RNG must end in the same
    state."""
    from sklearn.gaussian_process import GaussianProcessRegressor as SkGPR
    from skopt_b200 import engine
    from skopt_b200.learning import GaussianProcessRegressor

    rng = np.random.RandomState(2)
    X = rng.rand(40, 2)
    y = np.sin(4 * X[:, 0]) + X[:, 1] + 0.05 * rng.randn(40)
    kernel = ConstantKernel(1.0, (0.01, 1000.0)) * Matern(length_scale=[1.0, 1.0], length_scale_bounds=[(0.01, 100)] * 2,
                                                            nu=2.5) + WhiteKernel(1.0, (1e-5, 1e5))
    fakes = []

    def fake_device_fit(self):
        cached = self.__dict__.get("_skb_device_fit")
        if cached is None or cached[0] is not self.X_train_:
            fakes.append(_OracleFit(self.X_train_, self.y_train_, 2.5))
            cached = (self.X_train_, fakes[-1], engine.FitLayout(self.kernel_, 2))
            self.__dict__["_skb_device_fit"] = cached
            self.__dict__.pop("_skb_lockstep", None)
        return cached[1], cached[2]

    monkeypatch.setattr(GaussianProcessRegressor, "_device_fit", fake_device_fit)
    from skopt_b200.learning.gaussian_process import gpr as gpr_module
    monkeypatch.setattr(gpr_module, "_FIT_DEVICE", ["gpu"])

    class Sequential(GaussianProcessRegressor):          # same objective, scikit-learn's own restart loop
        def _constrained_optimization(self, obj_func, initial_theta, bounds):
            return SkGPR._constrained_optimization(self, obj_func, initial_theta, bounds)

    a = GaussianProcessRegressor(kernel=kernel, normalize_y=True, n_restarts_optimizer=3, random_state=5).fit(X, y)
    assert max(fakes[-1].batches) == 4 and len(fakes[-1].batches) > 5          # four searches really ran together
    b = Sequential(kernel=kernel, normalize_y=True, n_restarts_optimizer=3, random_state=5).fit(X, y)
    np.testing.assert_array_equal(a.kernel_.theta, b.kernel_.theta)
    assert a.log_marginal_likelihood_value_ == b.log_marginal_likelihood_value_
    np.testing.assert_array_equal(a.alpha_, b.alpha_)
    assert a._rng.get_state()[2] == b._rng.get_state()[2]
    np.testing.assert_array_equal(a._rng.get_state()[1], b._rng.get_state()[1])
    # the tail of the fit ran through the device objective's finalize at the chosen theta (noise still in it), and the
    # state scalars describe the prediction kernel (noise silenced: here the user's own WhiteKernel stays, noise=None)
    lay = engine.FitLayout(kernel, 2)
    np.testing.assert_array_equal(fakes[0].finalized, lay.expand(a.kernel_.theta))
    assert fakes[0].scalars.n_train == 40 and fakes[0].scalars.white == a.kernel_.k2.noise_level
    assert a.L_.shape == (40, 40) and not np.triu(a.L_, 1).any()
    assert "K_inv_" not in a.__dict__                       # host copy only on demand ...
    np.testing.assert_allclose(a.K_inv_.dot(a.L_.dot(a.L_.T)), np.eye(40), atol=1e-6)
    assert "K_inv_" in a.__dict__ and "_L_inv_lower" in a.__dict__
    # and both sit at scikit-learn's host optimum up to the optimiser's tolerance
    monkeypatch.undo()
    c = GaussianProcessRegressor(kernel=kernel, normalize_y=True, n_restarts_optimizer=3, random_state=5).fit(X, y)
    assert abs(a.log_marginal_likelihood_value_ - c.log_marginal_likelihood_value_) <= 1e-6 * abs(c.log_marginal_likelihood_value_)
