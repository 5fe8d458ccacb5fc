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
