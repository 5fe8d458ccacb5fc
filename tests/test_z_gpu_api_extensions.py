"""API extensions of the reference boundary, on a B200: regressors of other families under the Optimizer (value route
of the duck-typed boundary) and the public `kernel.gradient_x`.

Added late in the round: both compose entry points the rest of the GPU suite verifies (skb_acq_from_posterior_host +
skb_topk_host; skb_score_grad_host on an n = 1 state, cf. tests/test_gpu_parity.py::test_ragged_and_minimal_shapes).  First
run on a B200: 13 passed (profiles/r01_gpu_api_extensions_run.json); the same tests also run through the software model of
the ABI on the CPU (tests/test_host_optimizer.py)."""
import numpy as np
import pytest

import skopt_b200  # noqa: F401
from skopt_b200 import Optimizer

pytestmark = pytest.mark.gpu


def test_a_regressor_of_another_family_is_scored_through_its_own_predict():
    """optimizer.py:225-227: any sklearn regressor is accepted.  One that is not this package's GP keeps the search space
    un-normalised and is scored through predict(X, return_std=True), with the acquisition arithmetic and the ranking on the
    GPU (skb_acq_from_posterior_host, skb_topk_host); the suggestion is the arg-min of the same values on the same
    candidates."""
    from sklearn.gaussian_process import GaussianProcessRegressor as SklearnGPR
    from sklearn.gaussian_process.kernels import RBF as SklearnRBF
    from skopt_b200.acquisition import _gaussian_acquisition
    from skopt_b200.utils import has_gradients
    est = SklearnGPR(SklearnRBF(1.0, "fixed"), alpha=1e-6, optimizer=None)
    opt = Optimizer([(-2.0, 2.0), (0.0, 3.0)], est, n_initial_points=4, acq_func="gp_hedge", acq_optimizer="sampling",
                    acq_optimizer_kwargs=dict(n_points=500), random_state=3)
    assert opt.space.get_transformer() == ["identity", "identity"] and has_gradients(est)
    for _ in range(6):
        x = opt.ask()
        res = opt.tell(x, (x[0] - 0.5) ** 2 + (x[1] - 1.0) ** 2)
    assert len(res.models) == 3 and isinstance(res.models[-1], SklearnGPR) and len(opt.next_xs_) == 3
    # replay the last scoring block: same RNG state -> same candidates -> the arg-min of each acquisition
    twin = Optimizer([(-2.0, 2.0), (0.0, 3.0)], est, n_initial_points=4, acq_func="gp_hedge", acq_optimizer="sampling",
                     acq_optimizer_kwargs=dict(n_points=500), random_state=3)
    for x, y in zip(opt.Xi[:-1], opt.yi[:-1]):
        twin.tell(x, y)
    state = twin.rng.get_state()
    twin.tell(opt.Xi[-1], opt.yi[-1])
    rng = np.random.RandomState(0)
    rng.set_state(state)
    X = twin.space.rvs_transformed(500, rng)
    for acq, got in zip(["EI", "LCB", "PI"], twin.next_xs_):
        values = _gaussian_acquisition(X, twin.models[-1], np.min(twin.yi), acq)
        assert np.array_equal(got, X[np.argmin(values)])
    assert np.array_equal(np.asarray(twin.next_xs_), np.asarray(opt.next_xs_))


def _kernels_with_gradients():
    from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern, RBF, WhiteKernel
    ls = np.array([0.7, 1.9, 0.4])
    return [RBF(1.3), RBF(ls), Matern(0.8, nu=0.5), Matern(ls, nu=0.5), Matern(ls, nu=1.5), Matern(1.1, nu=2.5),
            Matern(ls, nu=2.5), ConstantKernel(2.5) * Matern(ls, nu=2.5) + WhiteKernel(0.3),
            1.7 * RBF(ls) + 0.4, ConstantKernel(3.0) + WhiteKernel(1.0)]


@pytest.mark.parametrize("index", range(10))
def test_kernel_gradient_x_matches_finite_differences(index):
    """learning/gaussian_process/tests/test_kernels.py:79-99 restated: kernel.gradient_x(x, X_train) against central
    differences of kernel(x, X_train), and finite everywhere."""
    kernel = _kernels_with_gradients()[index]
    rng = np.random.RandomState(index)
    x, X_train = rng.randn(3), rng.randn(7, 3)
    grad = kernel.gradient_x(x, X_train)
    assert grad.shape == (7, 3) and np.all(np.isfinite(grad))
    step = 1e-6
    numeric = np.empty((7, 3))
    for k in range(3):
        e = np.zeros(3)
        e[k] = step
        numeric[:, k] = (kernel(np.atleast_2d(x + e), X_train)[0] - kernel(np.atleast_2d(x - e), X_train)[0]) / (2 * step)
    assert np.max(np.abs(grad - numeric)) <= 1e-6 * max(1.0, np.max(np.abs(numeric)))


def test_kernel_gradient_x_conventions():
    from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, HammingKernel, Matern, RBF
    X_train = np.array([[0.2, -0.4], [1.0, 0.5], [0.2, -0.4]])
    x = X_train[0].copy()
    ls = np.array([0.5, 2.0])
    grad = (ConstantKernel(3.0) * Matern(ls, nu=0.5)).gradient_x(x, X_train)
    assert np.allclose(grad[0], -3.0 / ls, rtol=1e-14, atol=0) and np.array_equal(grad[2], grad[0])   # zero distance
    for kernel in (RBF(ls), Matern(ls, nu=1.5), Matern(ls, nu=2.5)):
        g = kernel.gradient_x(x, X_train)
        # smooth kernels: zero at zero distance (the device forms x * sum(P) - S, so "zero" is a rounding residue)
        assert np.all(np.abs(g[0]) <= 1e-15) and np.array_equal(g[2], g[0]) and np.all(np.abs(g[1]) > 1e-3)
    with pytest.raises(NotImplementedError):
        HammingKernel(np.ones(2)).gradient_x(x, X_train)
    with pytest.raises(ValueError):
        RBF(1.0).gradient_x(np.zeros(3), X_train)
