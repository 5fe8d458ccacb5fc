"""End-to-end drop-in check on a B200: gp_minimize / Optimizer trajectories against runs of the UNMODIFIED reference
stored in tests/golden/traj_gp_minimize.npz (oracle/make_golden_traj.py).

Sampling optimiser: every suggested point must be IDENTICAL (north star: bit-identical suggested points).
L-BFGS optimiser: start points identical; refined points agree within 1e-6 while the trajectories have not been
separated by line-search noise -- checked on the first model-based suggestion, which depends on a
single refinement."""
import json
import os

import numpy as np
import pytest

import skopt_b200
from skopt_b200 import Optimizer, gp_minimize

pytestmark = pytest.mark.gpu

GOLD = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "traj_gp_minimize.npz")))
RUNS = json.loads(str(GOLD["runs"]))


def branin(x, a=1, b=5.1 / (4 * np.pi ** 2), c=5. / np.pi, r=6, s=10, t=1. / (8 * np.pi)):
    return a * (x[1] - b * x[0] ** 2 + c * x[0] - r) ** 2 + s * (1 - t) * np.cos(x[0]) + s


def hart6(x):
    alpha = np.asarray([1.0, 1.2, 3.0, 3.2])
    P = 1e-4 * np.asarray([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                           [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
    A = np.asarray([[10, 3, 17, 3.50, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8],
                    [17, 8, 0.05, 10, 0.1, 14]])
    return -np.sum(alpha * np.exp(-np.sum(A * (np.array(x) - P) ** 2, axis=1)))


def mixed(x):
    return (x[0] - 0.5) ** 2 + 0.3 * (x[1] - 3) ** 2 + {"a": 0.0, "b": 1.0, "c": -0.5}[x[2]]


FUNCS = {"branin": branin, "hart6": hart6, "mixed": mixed}


def _run(name):
    cfg = RUNS[name]
    kw = {k: v for k, v in cfg.items() if k not in ("func", "dims")}
    dims = [tuple(d) if not isinstance(d[0], str) else list(d) for d in cfg["dims"]]
    res = gp_minimize(FUNCS[cfg["func"]], dims, **kw)
    return res, json.loads(str(GOLD[name + "_x_iters"])), GOLD[name + "_func_vals"], cfg


@pytest.mark.parametrize("name", ["branin_sampling_ei", "branin_sampling_hedge", "mixed_sampling_pi"])
def test_sampling_trajectory_identical(name):
    res, ref_x, ref_y, cfg = _run(name)
    assert res.x_iters == ref_x                      # every suggested point, bit for bit
    assert np.array_equal(np.asarray(res.func_vals), ref_y)
    assert len(res.models) == cfg["n_calls"] - cfg["n_initial_points"] + 1


WIDE = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "traj_wide.npz")))
WIDE_RUNS = json.loads(str(WIDE["runs"]))
_CATS = list("abcdefghi")


def sphere70(x):
    x = np.asarray(x, dtype=float)
    return float(np.sum((x - 0.3) ** 2) + np.sin(5.0 * x[0]) * x[3])


def mixed66(x):
    real = sum((v - 0.1 * (i % 7)) ** 2 for i, v in enumerate(x[:40]))
    ints = sum(0.02 * (v - 3) ** 2 for v in x[40:56])
    cat = sum((_CATS.index(c) - 4) ** 2 * (0.05 + 0.01 * i) for i, c in enumerate(x[56:]))
    return float(real + ints + cat)


@pytest.mark.parametrize("name", ["real70_sampling_ei", "mixed66_sampling_hedge"])
def test_wide_space_trajectory_identical(name):
    """More than 64 model dimensions (oracle/make_golden_wide.py: 70 Real dimensions; 40 Real + 16 Integer + 10
    Categorical): one K1 CTA per SM, candidates drawn on the device - every suggested point as in the reference."""
    cfg = WIDE_RUNS[name]
    kw = {k: v for k, v in cfg.items() if k not in ("func", "dims")}
    dims = [tuple(d) if not isinstance(d[0], str) else list(d) for d in cfg["dims"]]
    res = gp_minimize({"sphere70": sphere70, "mixed66": mixed66}[cfg["func"]], dims, **kw)
    assert res.x_iters == json.loads(str(WIDE[name + "_x_iters"]))
    assert np.array_equal(np.asarray(res.func_vals), WIDE[name + "_func_vals"])
    assert res.models[-1].X_train_.shape[1] > 64


def test_sampling_trajectory_tiny_noise():
    """noise=1e-10 (benchmarks/bench_*.py): K is nearly singular, where the reference's own variance moves by
    eps*cond(K) with summation order; the run must still reach an equally good minimum and agree on the first
    model-based suggestions."""
    res, ref_x, ref_y, cfg = _run("branin_sampling_lcb_noise1e-10")
    n0 = cfg["n_initial_points"]
    assert res.x_iters[:n0 + 1] == ref_x[:n0 + 1]
    assert res.fun <= ref_y.min() + 1.0


@pytest.mark.parametrize("name", ["hart6_lbfgs_ei", "hart6_lbfgs_hedge"])
def test_lbfgs_trajectory(name):
    res, ref_x, ref_y, cfg = _run(name)
    n0 = cfg["n_initial_points"]
    assert res.x_iters[:n0] == ref_x[:n0]                                   # random initial points: same RNG stream
    first = np.asarray(res.x_iters[n0]), np.asarray(ref_x[n0])
    assert np.max(np.abs(first[0] - first[1])) <= 1e-6                      # one refinement: same optimum
    # later suggestions: L-BFGS line searches amplify 1e-12 differences, so only a prefix can agree closely
    agree = 0
    for a, b in zip(res.x_iters, ref_x):
        if np.max(np.abs(np.asarray(a) - np.asarray(b))) > 1e-5:
            break
        agree += 1
    assert agree >= n0 + 2, agree


@pytest.mark.parametrize("strategy", ["cl_min", "cl_mean", "cl_max"])
def test_constant_liar_batch_identical(strategy):
    opt = Optimizer([(-5.0, 10.0), (0.0, 15.0)], "GP", n_initial_points=5, acq_func="EI", acq_optimizer="sampling",
                    acq_optimizer_kwargs=dict(n_points=1500), random_state=11)
    for _ in range(7):
        x = opt.ask()
        opt.tell(x, branin(x))
    pts = opt.ask(n_points=4, strategy=strategy)
    assert np.array_equal(np.asarray(pts, dtype=float), GOLD["cl_" + strategy])
    assert opt.ask(n_points=4, strategy=strategy) is pts or opt.ask(n_points=4, strategy=strategy) == pts   # cached
    assert np.array_equal(np.asarray(opt.ask(), dtype=float), GOLD["cl_" + strategy + "_next_after"])


def test_estimator_api_on_device():
    """predict() signature / shapes / exceptions of gpr.py:239-368 and the acquisition module-level functions."""
    from skopt_b200.acquisition import (_gaussian_acquisition, gaussian_acquisition_1D, gaussian_ei, gaussian_lcb,
                                        gaussian_pi)
    from skopt_b200.learning import GaussianProcessRegressor
    from skopt_b200.learning.gaussian_process.kernels import Matern, WhiteKernel
    rng = np.random.RandomState(0)
    X = rng.rand(40, 3)
    y = np.sin(3 * X[:, 0]) + X[:, 1] * X[:, 2]
    gpr = GaussianProcessRegressor(Matern(nu=2.5) + WhiteKernel(0.01), normalize_y=True, random_state=0).fit(X, y)
    Xq = rng.rand(17, 3)
    mu = gpr.predict(Xq)
    mu2, sd = gpr.predict(Xq, return_std=True)
    assert mu.shape == (17,) and sd.shape == (17,) and np.array_equal(mu, mu2) and np.all(sd > 0)
    m1, s1, gm, gs = gpr.predict(Xq[:1], return_std=True, return_mean_grad=True, return_std_grad=True)
    assert gm.shape == (3,) and gs.shape == (3,)
    with pytest.raises(RuntimeError):
        gpr.predict(Xq, return_std=True, return_cov=True)
    with pytest.raises(ValueError):
        gpr.predict(Xq[:1], return_std_grad=True)
    _, cov = gpr.predict(Xq[:4], return_cov=True)
    assert cov.shape == (4, 4) and np.allclose(np.sqrt(np.diag(cov)), sd[:4], rtol=1e-6)
    ei = gaussian_ei(Xq, gpr, y_opt=y.min())
    assert np.all(ei >= 0) and np.array_equal(_gaussian_acquisition(Xq, gpr, y.min(), "EI"), -ei)
    assert np.allclose(gaussian_lcb(Xq, gpr, kappa=1.5), mu - 1.5 * sd, rtol=1e-12)
    assert np.all((gaussian_pi(Xq, gpr, y_opt=y.min()) >= 0) & (gaussian_pi(Xq, gpr, y_opt=y.min()) <= 1))
    v, g = gaussian_acquisition_1D(Xq[0], gpr, y.min(), "EI")
    assert v.shape == (1,) and g.shape == (3,)
    with pytest.raises(ValueError):
        _gaussian_acquisition(Xq[0], gpr)                       # 1-D input
    with pytest.raises(ValueError):
        _gaussian_acquisition(Xq, gpr, acq_func="XX")

    # known answers of skopt/tests/test_acquisition.py:54-82 through the duck-typed (foreign model) route
    class ConstSurrogate:
        def predict(self, X, return_std=True):
            X = np.array(X)
            return np.zeros(X.shape[0]), np.ones(X.shape[0])
    one = [[1.0]]
    assert abs(gaussian_ei(one, ConstSurrogate(), -0.5, xi=0.)[0] - 0.1977966) < 1e-6
    assert abs(gaussian_pi(one, ConstSurrogate(), -0.5, xi=0.)[0] - 0.308538) < 1e-6
    assert gaussian_lcb(one, ConstSurrogate(), kappa="inf")[0] == -1.0
    assert abs(gaussian_lcb(one, ConstSurrogate(), kappa=0.3)[0] + 0.3) < 1e-12


# ----------------------------------------------------------------------------------------------------------------
# per-second acquisitions (EIps / PIps): acquisition.py:38-40, 61-80; optimizer.py:230-234, 485-491
# ----------------------------------------------------------------------------------------------------------------
PS = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "traj_per_second.npz")))


def _objective_ps(x):
    return branin(x), 1.0 + 0.1 * (x[0] + 5.0) + 0.05 * x[1]


@pytest.mark.parametrize("acq,optm", [("EIps", "sampling"), ("PIps", "sampling"), ("EIps", "lbfgs")])
def test_per_second_acquisition(acq, optm):
    from skopt_b200.acquisition import _gaussian_acquisition, gaussian_acquisition_1D
    key = "%s_%s" % (acq, optm)
    opt = Optimizer([(-5.0, 10.0), (0.0, 15.0)], "GP", n_initial_points=5, acq_func=acq, acq_optimizer=optm,
                    acq_optimizer_kwargs=dict(n_points=1500, n_restarts_optimizer=4), random_state=3)
    for _ in range(10):
        x = opt.ask()
        opt.tell(x, _objective_ps(x))
    Xi = np.asarray(opt.Xi, dtype=float)
    if optm == "sampling":
        assert np.array_equal(Xi, PS[key + "_Xi"])                 # every suggested point, bit for bit
        assert np.allclose(np.asarray(opt.yi, dtype=float), PS[key + "_yi"], rtol=0, atol=0)
    else:
        assert np.array_equal(Xi[:5], PS[key + "_Xi"][:5])
        assert np.max(np.abs(Xi[5] - PS[key + "_Xi"][5])) <= 1e-5
    # values and single-point gradients on a model pair refitted from the REFERENCE's own trajectory
    from sklearn.base import clone
    est = clone(opt.base_estimator_)
    est.fit(opt.space.transform(PS[key + "_Xi"].tolist()), PS[key + "_yi"].tolist())
    y_opt = np.min(PS[key + "_yi"])
    vals = _gaussian_acquisition(PS[key + "_Xc"], est, y_opt, acq)
    # the LML fit of these tiny data sets is ill-conditioned; the reference's own values move by ~eps*cond(K)
    cond = max(np.linalg.cond(e.K_inv_) for e in est.estimators_)
    rtol = max(1e-8, 1e-11 * cond)
    assert np.max(np.abs(vals - PS[key + "_vals"])) <= rtol * np.max(np.abs(PS[key + "_vals"]))
    assert int(np.argmin(vals)) == int(np.argmin(PS[key + "_vals"]))
    for i in range(8):
        v, g = gaussian_acquisition_1D(PS[key + "_Xc"][i], est, y_opt, acq)
        assert v.shape == (1,) and g.shape == (2,)
        assert abs(v[0] - PS[key + "_v1d"][i]) <= 10 * rtol * np.max(np.abs(PS[key + "_v1d"])) + 1e-300
        assert np.max(np.abs(g - PS[key + "_g1d"][i])) <= 100 * rtol * np.max(np.abs(PS[key + "_g1d"])) + 1e-300
    # batched gradients (extension) agree with the one-row calls
    vb, gb = _gaussian_acquisition(PS[key + "_Xc"][:8], est, y_opt, acq, return_grad=True)
    v0, g0 = gaussian_acquisition_1D(PS[key + "_Xc"][3], est, y_opt, acq)
    assert vb[3] == v0[0] and np.array_equal(gb[3], g0)
    pts = opt.ask(n_points=2)                                       # constant liar with (value, log-time) lies
    assert len(pts) == 2 and pts[0] != pts[1]


@pytest.mark.gpu
def test_expected_minimum_and_persistence(tmp_path):
    """utils.expected_minimum / expected_minimum_random_sampling / dump / load on a finished run (reference
    tests/test_utils.py::test_expected_minimum, test_dump_and_load)."""
    import skopt_b200
    from skopt_b200 import gp_minimize

    def branin(x):
        a, b, c, r, s, t = 1.0, 5.1 / (4 * np.pi ** 2), 5.0 / np.pi, 6.0, 10.0, 1.0 / (8 * np.pi)
        return a * (x[1] - b * x[0] ** 2 + c * x[0] - r) ** 2 + s * (1 - t) * np.cos(x[0]) + s

    res = gp_minimize(branin, [(-5.0, 10.0), (0.0, 15.0)], n_calls=20, n_initial_points=10, random_state=1,
                      acq_optimizer="sampling", n_points=2000, noise=1e-10)
    x_min, f_min = skopt_b200.expected_minimum(res, n_random_starts=20, random_state=1)
    x_min2, f_min2 = skopt_b200.expected_minimum(res, n_random_starts=20, random_state=1)
    assert x_min == x_min2 and f_min == f_min2
    assert all(lo <= v <= hi for v, (lo, hi) in zip(x_min, res.space.bounds))
    model = res.models[-1]
    assert f_min == pytest.approx(model.predict(res.space.transform([x_min]))[0], rel=1e-9, abs=1e-9)
    x_rs, f_rs = skopt_b200.expected_minimum_random_sampling(res, n_random_starts=100000, random_state=1)
    assert f_min <= f_rs + 1e-6 * max(1.0, abs(f_rs))          # local refinement beats dense random sampling
    assert f_rs == pytest.approx(model.predict(res.space.transform([x_rs]))[0], rel=1e-9, abs=1e-9)
    # the gradient handed to L-BFGS is the derivative of the function it minimises
    x0 = np.array([2.5, 7.0])
    h = 1e-3           # alpha is large at noise=1e-10: smaller steps drown in the rounding of k . alpha
    def mean_at(x):
        return model.predict(res.space.transform([list(x)]))[0]
    dev = model.device_state()
    g = dev.score_grad(res.space.transform([list(x0)]), acq_funcs=())["mean_grad"][0]
    slopes = np.array([skopt_b200.utils._transform_slope(d, v) for d, v in zip(res.space.dimensions, x0)])
    for j in range(2):
        e = np.zeros(2); e[j] = h
        fd = (mean_at(x0 + e) - mean_at(x0 - e)) / (2 * h)
        assert g[j] * slopes[j] == pytest.approx(fd, rel=1e-3, abs=1e-5)

    path = tmp_path / "res.pkl"
    skopt_b200.dump(res, path, store_objective=False)
    back = skopt_b200.load(path)
    assert "func" not in back.specs["args"] and "func" in res.specs["args"]
    assert back.x == res.x and back.fun == res.fun
    Xq = np.random.RandomState(0).rand(50, 2)
    np.testing.assert_array_equal(back.models[-1].predict(Xq), model.predict(Xq))


def test_constant_liar_rank1_batch_equals_manual_frozen_kernel_loop():
    """Optimizer(liar_update="rank1").ask(n_points, 'cl_min'): every lie is appended to the current model's factor
    (hyper-parameters frozen); the batch equals a hand-written loop that refits a fixed-kernel GP from scratch on the
    lied data and scores the same candidate stream on the GPU."""
    import numpy as np
    from skopt_b200 import Optimizer
    from skopt_b200.learning import GaussianProcessRegressor
    rng = np.random.RandomState(3)
    d, n0 = 3, 30
    X0 = rng.rand(n0, d)
    y0 = np.sin(X0.dot(rng.randn(d))) + 0.05 * rng.randn(n0)
    kw = dict(n_points=5000, liar_update="rank1")
    opt = Optimizer([(0.0, 1.0)] * d, "GP", n_initial_points=n0, acq_func="EI", acq_optimizer="sampling",
                    acq_optimizer_kwargs=kw, random_state=7)
    opt.tell(X0.tolist(), y0.tolist())
    state = opt.rng.get_state()
    batch = opt.ask(n_points=4, strategy="cl_min")
    assert len(batch) == 4 and batch[0] == opt._next_x and len({tuple(b) for b in batch}) == 4
    assert opt.ask(n_points=4, strategy="cl_min") is batch                     # cached, like the reference
    # manual loop: same sandbox seed, same candidate streams, from-scratch fits with the fitted kernel frozen
    model = opt.models[-1]
    sandbox_rng = np.random.RandomState()
    sandbox_rng.set_state(state)
    sandbox_rng = np.random.RandomState(sandbox_rng.randint(0, np.iinfo(np.int32).max))
    Xi, yi = [list(x) for x in X0.tolist()], list(y0)
    expect = [batch[0]]
    for step in range(3):
        Xi.append(expect[-1])
        yi.append(float(np.min(yi)))
        frozen = GaussianProcessRegressor(kernel=model.kernel_, noise=model.noise_, normalize_y=model.normalize_y,
                                          alpha=model.alpha, optimizer=None).fit(opt.space.transform(Xi), yi)
        cand = opt.space.rvs_transformed(5000, sandbox_rng)
        res = frozen.device_state().score(cand, y_opt=float(np.min(yi)), acq_funcs=("EI",), top_k=1)
        x = np.clip(cand[int(res["EI"]["topk_idx"][0])], 0.0, 1.0)
        expect.append(opt.space.inverse_transform(x.reshape(1, -1))[0])
    assert batch == expect
