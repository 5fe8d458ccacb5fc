"""All-categorical GP (HammingKernel, skopt/learning/gaussian_process/kernels.py:317-407; cook_estimator
utils.py:377-378; has_gradients utils.py:320-330) against values recorded from the UNMODIFIED reference
(oracle/make_golden_hamming.py -> tests/golden/hamming_*.npz, traj_categorical.npz).

CPU part: the kernel class that serves scikit-learn's marginal-likelihood fit (bit-identical K and dK/dtheta, hence
bit-identical fitted hyper-parameters) and the host plumbing.  GPU part: posterior / acquisition parity of the
hamming_* fixtures runs through the parametrised tests of test_gpu_parity.py / test_gpu_split.py; here the gp_minimize
trajectories (every suggested point identical) and the error behaviour of the gradient entry points."""
import json
import os
import warnings

import numpy as np
import pytest

from oracle import gp_oracle as orc
from skopt_b200 import engine
from skopt_b200.learning import GaussianProcessRegressor
from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, HammingKernel, Matern, WhiteKernel
from skopt_b200.space import Space, normalize_dimensions
from skopt_b200.utils import cook_estimator, has_gradients
from tests._golden import load

GOLD = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "traj_categorical.npz")))
DIMS = json.loads(str(GOLD["dims"]))
RUNS = json.loads(str(GOLD["runs"]))


def objective(x):
    score = {"a": 0.0, "b": 1.0, "c": -0.5, "d": 0.3}[x[0]] + {"x": 0.2, "y": -0.4, "z": 0.9}[x[1]]
    return score * (1.5 if x[2] == "hi" else 0.7) + 0.1 * len(x[3])


@pytest.mark.parametrize("tag,ls", [("aniso", np.array([0.4, 1.3, 2.2])), ("iso", 0.8), ("iso1", [1.7])])
def test_kernel_class_bit_identical_to_reference(tag, ls):
    k = HammingKernel(ls)
    K, G = k(GOLD["hk_X"], eval_gradient=True)
    assert np.array_equal(K, GOLD["hk_%s_K" % tag])
    assert G.shape == GOLD["hk_%s_grad" % tag].shape and np.array_equal(G, GOLD["hk_%s_grad" % tag])
    assert np.array_equal(k(GOLD["hk_X"], GOLD["hk_Y"]), GOLD["hk_%s_cross" % tag])
    assert np.array_equal(k.diag(GOLD["hk_X"]), np.ones(len(GOLD["hk_X"]))) and k.is_stationary()
    assert k.theta.shape == ((3,) if tag == "aniso" else (1,))


def test_kernel_class_errors_and_algebra():
    k = HammingKernel([1.0, 2.0])
    with pytest.raises(ValueError, match="Expected X to have 2 features"):
        k(np.zeros((3, 3)))
    with pytest.raises(ValueError, match="gradient can be evaluated only"):
        k(np.zeros((3, 2)), np.zeros((2, 2)), eval_gradient=True)
    prod = 2.0 * k + WhiteKernel(0.1)
    assert type(prod).__module__ == HammingKernel.__module__          # operators stay inside this package
    X = np.array([[0.0, 1.0], [0.0, 0.5], [1.0, 0.5]])
    expect = 2.0 * np.exp(-np.array([[0, 2, 3], [2, 0, 1], [3, 1, 0]], dtype=float)) + 0.1 * np.eye(3)
    assert np.allclose(prod(X), expect, rtol=1e-15)
    assert np.array_equal(k.clone_with_theta(k.theta).length_scale, k.length_scale)


def test_cooked_estimator_and_gradient_dispatch():
    space = normalize_dimensions(DIMS)
    assert space.is_categorical
    est = cook_estimator("GP", space=space, random_state=5, noise="gaussian")
    assert isinstance(est.kernel.k2, HammingKernel) and np.array_equal(est.kernel.k2.length_scale, np.ones(4))
    assert est.kernel.k2.length_scale_bounds == (1e-5, 1e5)
    assert not has_gradients(est)
    assert not has_gradients(GaussianProcessRegressor(kernel=HammingKernel() + WhiteKernel()))
    assert has_gradients(cook_estimator("GP", space=Space([(0.0, 1.0), ["a", "b"]])))      # mixed space: Matern
    assert has_gradients(GaussianProcessRegressor(kernel=ConstantKernel() * Matern()))
    from skopt_b200 import Optimizer
    with pytest.raises(ValueError, match="should run with acq_optimizer='sampling'"):
        Optimizer(DIMS, "GP", acq_optimizer="lbfgs")
    assert Optimizer(DIMS, "GP", acq_optimizer="auto").acq_optimizer == "sampling"
    assert Optimizer([(0.0, 1.0), ["a", "b"]], "GP", acq_optimizer="auto").acq_optimizer == "lbfgs"


def test_host_fit_reproduces_the_reference_state():
    """Same data, same seed: scikit-learn's search over our kernel class must land on the reference's hyper-parameters
    and produce the reference's alpha_ bit for bit (the fit never leaves the host in the default mode); K_inv_ is formed
    from the same factor in a different order here (L_inv^T L_inv), so it agrees to rounding, as for the other kernels."""
    rec, meta, ref = load("hamming_cooked_categorical")
    space = normalize_dimensions(DIMS)
    est = cook_estimator("GP", space=space, random_state=5, noise="gaussian")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        est.fit(rec["X_train"], rec["y"])
    p = meta["kernel_params"]
    assert est.kernel_.k1.k1.constant_value == p["k1__k1__constant_value"]
    assert np.array_equal(est.kernel_.k1.k2.length_scale, np.asarray(p["k1__k2__length_scale"]))
    assert est.noise_ == float(rec["noise_"])
    assert np.array_equal(est.alpha_, rec["alpha"])
    assert np.allclose(est.K_inv_, rec["K_inv"], rtol=1e-7, atol=1e-9 * np.abs(rec["K_inv"]).max())
    # flattening: product walker and checker walker agree, and give the reference's cross-covariances
    hs, st = engine.host_state_from_estimator(est), orc.state_from_estimator(est)
    assert hs.kernel == 4 and st.kind == orc.KIND_HAMMING and np.array_equal(hs.length_scale, st.length_scale)
    assert hs.amplitude == st.amplitude and hs.white == 0.0
    assert np.max(np.abs(orc.kernel_cross(st, rec["Xc"]) - rec["K_trans"])) <= 1e-12 * hs.amplitude
    with pytest.raises(NotImplementedError):          # the device fit objective is RBF / Matern only: host fit stays
        engine.FitLayout(est.kernel_, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(RUNS))
def test_categorical_trajectory_identical(name):
    from skopt_b200 import gp_minimize
    res = gp_minimize(objective, DIMS, **RUNS[name])
    assert res.x_iters == json.loads(str(GOLD[name + "_x_iters"]))          # every suggested point
    assert np.array_equal(np.asarray(res.func_vals), GOLD[name + "_func_vals"])
    assert repr(res.models[-1].kernel_) == str(GOLD[name + "_kernel"])


@pytest.mark.gpu
def test_gradient_entry_points_refuse_the_hamming_kernel():
    from skopt_b200 import _lib
    rec, meta, est = load("hamming_aniso_const")
    dev = engine.DeviceState.from_estimator(est)
    with pytest.raises(_lib.SkbError) as e:
        dev.score_grad(rec["Xc"][:3], y_opt=float(rec["y_opt"]), acq_funcs=("EI",))
    assert e.value.code == -2 and "no input gradient" in str(e.value)
    res = dev.score(rec["Xc"], y_opt=float(rec["y_opt"]), acq_funcs=("EI",), top_k=1)
    assert res["EI"]["topk_idx"][0] == int(rec["argmin_EI"])
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_cats", [[5, 4, 7, 3, 6, 2, 8, 4], [3, 4, 5, 2] * 10])
def test_large_categorical_candidate_set_all_precisions_agree(n_cats):
    """n_train = 600 categorical rows (duplicates included), 20 000 candidates: the FP64 path, the 7-plane and the
    self-checked split paths against the oracle; ties between identical candidates resolve to the lowest index.  The
    40-dimensional case needs more than 48 KB of dynamic shared memory in K1 (opt-in attribute of every kernel kind)."""
    rng = np.random.RandomState(0)
    X = np.column_stack([rng.randint(0, k, size=600) / (k - 1.0) for k in n_cats])
    w = rng.uniform(0.05, 1.5, size=len(n_cats))
    y = np.cos(4.0 * X @ w) + 0.05 * rng.randn(600)
    est = GaussianProcessRegressor(ConstantKernel(1.4, "fixed") * HammingKernel(w, "fixed"), normalize_y=True,
                                   noise=1e-2, optimizer=None).fit(X, y)
    Xc = np.column_stack([rng.randint(0, k, size=20000) / (k - 1.0) for k in n_cats])
    st = orc.state_from_estimator(est)
    mu, sd = orc.predict(st, Xc, return_std=True, quad="gemm")
    ref = -orc.ei_from_posterior(mu, sd, float(y.min()), 0.01)
    dev = engine.DeviceState.from_estimator(est)
    for precision in ("fp64", "split_int8_p7", "split_int8"):
        res = dev.score(Xc, y_opt=float(y.min()), acq_funcs=("EI",), xi=0.01, top_k=5, precision=precision)
        assert np.max(np.abs(res["mean"] - mu)) <= 1e-9 * max(np.max(np.abs(mu)), st.y_std), precision
        assert np.max(np.abs(res["std"] ** 2 - sd ** 2)) <= 1e-9 * st.prior_var * st.y_std ** 2, precision
        v = res["EI"]["values"]
        assert np.max(np.abs(v - ref)) <= 1e-9 * np.max(np.abs(ref)), precision
        assert np.array_equal(res["EI"]["topk_idx"], np.lexsort((np.arange(v.size), v))[:5]), precision
    dev.close()
