"""Partial dependence (skopt/plots.py:896-1055): the batched evaluation against a grid-value-by-grid-value loop, and
against results recorded from the unmodified reference (tests/golden/traj_partial_dependence.npz, made by
oracle/make_golden_pd.py) with the CPU oracle as the model; on the GPU with the CUDA posterior as the model."""
import numpy as np
import pytest

import skopt_b200  # noqa: F401
from oracle import gp_oracle as orc
from skopt_b200.plots import _evenly_sample, partial_dependence_1D, partial_dependence_2D
from skopt_b200.space import Space, normalize_dimensions
from tests._golden import load

DIMS = [(-5.0, 10.0), (0, 15), ["a", "b", "c"]]


class _Model:
    def __init__(self, fn):
        self.predict, self.calls = self._predict, 0
        self._fn = fn

    def _predict(self, X):
        self.calls += 1
        return self._fn(np.asarray(X))


def _loop_1d(space, model, i, samples, n_points):
    starts = np.cumsum([0] + [d.transformed_size for d in space.dimensions])
    xi, xt = _evenly_sample(space.dimensions[i], n_points)
    out = []
    for x in xt:
        rows = np.array(samples)
        rows[:, starts[i]:starts[i + 1]] = x
        out.append(np.mean(model.predict(rows)))
    return xi, out


def _loop_2d(space, model, i, j, samples, n_points):
    starts = np.cumsum([0] + [d.transformed_size for d in space.dimensions])
    xi, xt = _evenly_sample(space.dimensions[j], n_points)
    yi, yt = _evenly_sample(space.dimensions[i], n_points)
    zi = np.empty((len(yi), len(xi)))
    for a, y in enumerate(yt):
        for b, x in enumerate(xt):
            rows = np.array(samples)
            rows[:, starts[j]:starts[j + 1]] = x
            rows[:, starts[i]:starts[i + 1]] = y
            zi[a, b] = np.mean(model.predict(rows))
    return xi, yi, zi


@pytest.mark.parametrize("onehot", [False, True])
def test_batched_partial_dependence_equals_the_loop(onehot):
    space = Space(DIMS) if onehot else normalize_dimensions(DIMS)        # one-hot: the category spans three columns
    rng = np.random.RandomState(0)
    samples = space.transform(space.rvs(30, random_state=rng))
    w = rng.randn(samples.shape[1])
    fn = lambda X: np.sin(X.dot(w)) + X[:, 0] ** 2                       # noqa: E731
    for i in range(3):
        model = _Model(fn)
        xi, yi = partial_dependence_1D(space, model, i, samples, n_points=7)
        assert model.calls == 1
        xr, yr = _loop_1d(space, _Model(fn), i, samples, 7)
        np.testing.assert_array_equal(xi, xr)
        np.testing.assert_array_equal(yi, yr)
    for i, j in ((0, 1), (2, 0), (1, 2), (1, 1)):
        model = _Model(fn)
        xi, yi, zi = partial_dependence_2D(space, model, i, j, samples, n_points=5)
        assert model.calls == 1
        xr, yr, zr = _loop_2d(space, _Model(fn), i, j, samples, 5)
        np.testing.assert_array_equal(xi, xr)
        np.testing.assert_array_equal(yi, yr)
        np.testing.assert_array_equal(zi, zr)


def _check_against_fixture(rec, space, model, rtol):
    scale = float(rec["y_train_std"])
    for i in range(3):
        xi, yi = partial_dependence_1D(space, model, i, rec["samples"], n_points=9)
        np.testing.assert_array_equal(np.asarray(xi, dtype=float), rec["pd1_x_%d" % i])
        assert np.max(np.abs(np.asarray(yi) - rec["pd1_y_%d" % i])) <= rtol * scale
    for i, j in ((0, 1), (2, 0), (1, 2)):
        xi, yi, zi = partial_dependence_2D(space, model, i, j, rec["samples"], n_points=6)
        np.testing.assert_array_equal(np.asarray(xi, dtype=float), rec["pd2_x_%d%d" % (i, j)])
        np.testing.assert_array_equal(np.asarray(yi, dtype=float), rec["pd2_y_%d%d" % (i, j)])
        assert np.max(np.abs(zi - rec["pd2_z_%d%d" % (i, j)])) <= rtol * scale


def test_partial_dependence_matches_reference_with_oracle_model():
    rec, _, est = load("traj_partial_dependence")
    st = orc.state_from_estimator(est)
    _check_against_fixture(rec, normalize_dimensions(DIMS), _Model(lambda X: orc.predict(st, X)), 1e-9)


@pytest.mark.gpu
def test_partial_dependence_matches_reference_on_gpu():
    from skopt_b200.engine import DeviceState
    rec, _, est = load("traj_partial_dependence")
    dev = DeviceState.from_estimator(est)
    _check_against_fixture(rec, normalize_dimensions(DIMS), _Model(dev.predict_mean), 1e-9)
    # a 40 x 40 grid over 250 samples (the reference's plot_objective defaults) is one 400 000-row mean-only pass
    space = normalize_dimensions(DIMS)
    samples = space.transform(space.rvs(250, random_state=np.random.RandomState(1)))
    model = _Model(dev.predict_mean)
    xi, yi, zi = partial_dependence_2D(space, model, 0, 1, samples, n_points=40)
    assert model.calls == 1 and zi.shape == (40, 40) and np.isfinite(zi).all()
