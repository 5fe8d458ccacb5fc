"""Host side above the C ABI, exercised on the CPU: the tests of tests/test_gpu_optimizer.py and
tests/test_gpu_reference_behaviours.py that only use HOST entry points run here against the software model of the ABI
(tests/_abi_model.py, `abi_model` fixture) - same fixtures from the unmodified reference, same assertions.  What this
pins without a GPU: the ctypes marshalling (struct layouts, pointer / NULL conventions, top-k and first-NaN protocol),
`GaussianProcessRegressor.predict`, the acquisition module, the Optimizer's RNG draw order, hedge bookkeeping,
constant-liar loop and the lock-step L-BFGS driver.  The arithmetic behind the ABI is the CPU oracle here; the CUDA
kernels are checked by the `-m gpu` originals."""
import numpy as np
import pytest

import skopt_b200  # noqa: F401
from tests import test_gpu_optimizer as gpu_opt


def test_model_is_what_the_package_talks_to(abi_model):
    from skopt_b200 import _lib, engine
    from oracle import gp_oracle as orc
    if abi_model.on_device:
        pytest.skip("the real library is loaded (SKB_TESTS_ON_DEVICE=1)")
    assert _lib.lib() is abi_model
    st = orc.make_synthetic_state(24, 3)
    L_inv = np.linalg.inv(np.linalg.cholesky(np.linalg.inv(st.K_inv)))        # K_inv = L_inv^T L_inv, L_inv lower
    hs = engine.HostState(st.X_train, st.alpha, L_inv, st.length_scale, 3, st.amplitude,
                          st.bias, st.white, st.y_mean, st.y_std, K_inv=st.K_inv)
    dev = engine.DeviceState(hs)
    X = np.random.RandomState(0).rand(50, 3)
    res = dev.score(X, y_opt=-0.3, acq_funcs=("EI", "LCB"), top_k=4)
    want = orc.score_candidates(X, st, -0.3, ("EI", "LCB"), top_k=4)
    for acq in ("EI", "LCB"):
        assert np.allclose(res[acq]["values"], want[acq]["values"], rtol=1e-6, atol=1e-9)
        assert res[acq]["topk_idx"].tolist() == want[acq]["topk"].tolist() and res[acq]["first_nan"] == -1
    assert np.allclose(dev.predict_mean(X), res["mean"])
    assert abi_model.calls[:2] == ["skb_state_create", "skb_score_host"]
    dev.close()


def _mirror(module, skip=()):
    """Re-export every test function of a `-m gpu` module as a CPU test that runs it inside the `abi_model` fixture,
    keeping its parametrisation and fixtures (the wrapper gets the original's signature plus `abi_model`)."""
    import inspect
    made = {}
    for name, func in sorted(vars(module).items()):
        if not (name.startswith("test_") and inspect.isfunction(func)) or name in skip:
            continue
        args = list(inspect.signature(func).parameters)
        scope = {"_orig": func}
        exec("def {0}(abi_model{1}):\n    return _orig({2})".format(
            name, "".join(", " + a for a in args), ", ".join("%s=%s" % (a, a) for a in args)), scope)
        wrapper = scope[name]
        wrapper.__doc__ = func.__doc__
        wrapper.__module__ = __name__
        wrapper.pytestmark = [m for m in getattr(func, "pytestmark", []) if m.name == "parametrize"]
        made[name] = wrapper
    return made


# tests/test_gpu_optimizer.py: trajectories of the unmodified reference (sampling: identical points; L-BFGS: same
# optimum), constant liar, estimator / acquisition API, per-second acquisitions, expected_minimum, dump / load
globals().update(_mirror(gpu_opt))

# tests/test_gpu_reference_behaviours.py: what the reference's own test files check on this path (convergence
# thresholds, determinism, validation, result object, constant liar, categorical spaces), restated for this package
from tests import test_gpu_reference_behaviours as gpu_behaviours  # noqa: E402

globals().update({"test_behaviour_" + name[5:]: func for name, func in _mirror(gpu_behaviours).items()})

# tests/test_z_gpu_api_extensions.py: regressors of other families under the Optimizer, kernel.gradient_x
from tests import test_z_gpu_api_extensions as gpu_extensions  # noqa: E402

globals().update({"test_extension_" + name[5:]: func for name, func in _mirror(gpu_extensions).items()})

# all-categorical GPs and partial dependence: the `-m gpu` members of the mixed modules that use host entry points only
# (test_large_categorical_candidate_set_all_precisions_agree compares precisions of the CUDA kernels: GPU only)
from tests import test_hamming as hamming_tests  # noqa: E402
from tests import test_partial_dependence as pd_tests  # noqa: E402

globals().update({"test_hamming_" + name[5:]: func for name, func in _mirror(hamming_tests).items()
                  if name in ("test_categorical_trajectory_identical",
                              "test_gradient_entry_points_refuse_the_hamming_kernel")})
globals().update({"test_pd_" + name[5:]: func for name, func in _mirror(pd_tests).items()
                  if name == "test_partial_dependence_matches_reference_on_gpu"})


def test_foreign_regressor_with_gradients_is_refined_by_lbfgs(abi_model):
    """A regressor of another family that returns posterior gradients for one row (the reference's convention) runs with
    acq_optimizer="lbfgs": ranking through skb_topk_host, every L-BFGS evaluation through its own one-row predict() and
    the chain rule of skb_acq_from_posterior_host.  The model is analytic so the optimum of LCB is known:
    mu = |x - c|^2, std = 0.5 + 0.1 x_0  ->  argmin(mu - kappa std) = c + 0.05 kappa e_0."""
    from sklearn.base import BaseEstimator, RegressorMixin
    from skopt_b200 import Optimizer
    from skopt_b200.utils import has_gradients

    class Bowl(RegressorMixin, BaseEstimator):
        def __init__(self, centre=(0.3, -0.4)):
            self.centre = centre

        def fit(self, X, y):
            self.n_seen_ = len(y)
            return self

        def predict(self, X, return_std=False, return_mean_grad=False, return_std_grad=False):
            X = np.asarray(X, dtype=float)
            if (return_mean_grad or return_std_grad) and X.shape[0] != 1:
                raise ValueError("Not implemented for n_samples > 1")
            c = np.asarray(self.centre)
            mu, std = np.sum((X - c) ** 2, axis=1), 0.5 + 0.1 * X[:, 0]
            if return_mean_grad:
                return mu, std, 2.0 * (X[0] - c), np.array([0.1, 0.0])
            return (mu, std) if return_std else mu

    model = Bowl()
    assert has_gradients(model)
    opt = Optimizer([(-1.0, 1.0), (-1.0, 1.0)], model, n_initial_points=2, acq_func="LCB", acq_optimizer="lbfgs",
                    acq_optimizer_kwargs=dict(n_points=200, n_restarts_optimizer=3),
                    acq_func_kwargs=dict(kappa=2.0), random_state=0)
    for _ in range(3):
        x = opt.ask()
        opt.tell(x, float(np.sum(np.square(x))))
    assert np.allclose(opt.ask(), [0.3 + 0.05 * 2.0, -0.4], atol=1e-6)
    if not abi_model.on_device:
        assert {"skb_acq_from_posterior_host", "skb_topk_host"} <= set(abi_model.calls)
        assert "skb_state_create" not in abi_model.calls        # no GP state: the model is not ours


def test_tree_families_are_routed_to_sampling():
    from sklearn.ensemble import GradientBoostingRegressor, RandomForestRegressor
    from skopt_b200 import Optimizer
    from skopt_b200.utils import cook_estimator, has_gradients
    assert not has_gradients(RandomForestRegressor()) and not has_gradients(GradientBoostingRegressor())
    assert Optimizer([(0.0, 1.0)], RandomForestRegressor(), acq_func="EI").acq_optimizer == "sampling"
    with pytest.raises(ValueError, match="should run with acq_optimizer='sampling'"):
        Optimizer([(0.0, 1.0)], RandomForestRegressor(), acq_optimizer="lbfgs")
    with pytest.raises(ValueError, match="'RF', 'ET', 'GP', 'GBRT' or 'DUMMY'"):
        cook_estimator("rtr")
    with pytest.raises(ValueError, match="not built by this package"):
        cook_estimator("ET", space=[(0.0, 1.0)])


_SHARDED_WORKER = """
import sys
sys.path.insert(0, %(root)r)
import numpy as np, torch.distributed as dist
import skopt_b200
from skopt_b200 import Optimizer, _lib
from skopt_b200.optimizer import optimizer as optimizer_module
from tests._abi_model import AbiModel
_lib._lib = AbiModel()                                    # test stand-in for libskb.so (no GPU on the build machine)
optimizer_module.device_dim_descs = lambda space: None
dist.init_process_group("gloo")
rank = dist.get_rank()

def objective(x):
    return (x[0] - 0.3) ** 2 + (x[1] + 0.2) ** 2 + 0.1 * np.sin(5 * x[2])

def run(shard, search):
    opt = Optimizer([(-1.0, 1.0)] * 3, "GP", n_initial_points=4, acq_func="gp_hedge", acq_optimizer=search,
                    acq_optimizer_kwargs=dict(n_points=301, n_restarts_optimizer=3, shard_group=shard), random_state=5)
    for _ in range(7):
        x = opt.ask()
        opt.tell(x, objective(x))
    batch = opt.ask(n_points=3, strategy="cl_min")
    return opt.Xi + batch, [r.tolist() for r in opt._last_scoring["EI"]["topk_idx"].reshape(1, -1)]

for search in ("sampling", "lbfgs"):
    sharded = run(True, search)            # every tell() is a collective: rank r scores its block of the 301 candidates
    local = run(False, search)             # the same optimiser without the collective
    assert sharded == local, (search, rank)
    everyone = [None, None]
    dist.all_gather_object(everyone, sharded)
    assert everyone[0] == everyone[1], search
dist.destroy_process_group()
open(%(done)r + str(rank), "w").write("ok")
"""


def test_sharded_optimizer_two_ranks_gloo(tmp_path):
    """The N > 1 path of the Optimizer on the CPU (gloo, world size 2): with `shard_group=True` every rank drives an
    identical optimiser, scores its own block of the candidate set and merges the per-rank top-k; asked points (hedge,
    sampling and L-BFGS, constant-liar batch included) equal the unsharded run's and agree across ranks."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "sharded_worker.py"
    done = str(tmp_path / "done_rank")
    script.write_text(_SHARDED_WORKER % {"root": root, "done": done})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29547", str(script)],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert os.path.exists(done + "0") and os.path.exists(done + "1")       # both ranks reached the end of the script


def test_expected_minimum_of_a_foreign_model():
    """utils.expected_minimum on a result whose model is not this package's GP minimises the model's own predict()."""
    from sklearn.base import BaseEstimator, RegressorMixin
    from skopt_b200.space import Space
    from skopt_b200.utils import create_result, expected_minimum

    class Bowl(RegressorMixin, BaseEstimator):
        def predict(self, X, return_std=False):
            return np.sum((np.asarray(X) - np.array([0.25, -0.5])) ** 2, axis=1) + 3.0

    space = Space([(-1.0, 1.0), (-1.0, 1.0)])
    res = create_result([[0.9, 0.9], [0.0, 0.0]], [5.0, 4.0], space, models=[Bowl()])
    x, fun = expected_minimum(res, n_random_starts=3, random_state=0)
    assert np.allclose(x, [0.25, -0.5], atol=1e-5) and abs(fun - 3.0) < 1e-9


def test_foreign_regressor_on_a_mixed_unnormalised_space(abi_model):
    """A regressor of another family keeps the reference's default warpings (one-hot categoricals, identity reals):
    the candidates, the model's training matrix and the suggested point all live in that space."""
    from sklearn.gaussian_process import GaussianProcessRegressor as SklearnGPR
    from skopt_b200 import Optimizer
    dims = [(-1.0, 1.0), ["red", "green", "blue"], (1, 5)]
    opt = Optimizer(dims, SklearnGPR(alpha=1e-6), n_initial_points=3, acq_func="EI", acq_optimizer="sampling",
                    acq_optimizer_kwargs=dict(n_points=200), random_state=2)
    assert opt.space.get_transformer() == ["identity", "onehot", "identity"] and opt.space.transformed_n_dims == 5
    for _ in range(5):
        x = opt.ask()
        assert x in opt.space and isinstance(x[1], str)
        opt.tell(x, x[0] ** 2 + {"red": 0.0, "green": 1.0, "blue": 2.0}[x[1]] + 0.1 * x[2])
    assert opt.models[-1].X_train_.shape == (5, 5)
    assert set(np.unique(opt.models[-1].X_train_[:, 1:4])) <= {0.0, 1.0}
    assert opt.next_xs_[0].shape == (5,) and opt.ask() in opt.space


def test_dummy_minimize_is_random_search():
    """skopt/optimizer/dummy.py: no surrogate, every point a draw from the space; nothing reaches the C ABI."""
    from skopt_b200 import Space, dummy_minimize
    dims = [(-1.0, 1.0), ["a", "b"], (1, 9)]
    calls = []

    def objective(x):
        calls.append(list(x))
        return (x[0] - 0.3) ** 2 + {"a": 0.0, "b": 1.0}[x[1]] + 0.1 * x[2]

    res = dummy_minimize(objective, dims, n_calls=12, random_state=3)
    assert res.x_iters == calls and len(calls) == 12 and res.models == [] and res.fun == min(res.func_vals)
    assert res.x_iters == [p for p in _draws(Space(dims), 12, 3)]
    calls.clear()
    res = dummy_minimize(objective, dims, n_calls=5, random_state=1, x0=[[0.1, "a", 2], [0.5, "b", 3]], y0=[1.0, 2.0])
    assert len(res.x_iters) == 7 and len(calls) == 5 and res.x_iters[:2] == [[0.1, "a", 2], [0.5, "b", 3]]
    res = dummy_minimize(objective, dims, n_calls=6, random_state=1, x0=[[0.1, "a", 2]])
    assert len(res.x_iters) == 6 and res.x_iters[0] == [0.1, "a", 2]


def _draws(space, n, seed):
    rng = np.random.RandomState(seed)
    rng.randint(0, np.iinfo(np.int32).max)          # the Optimizer seeds its (absent) estimator first, like the reference
    return [space.rvs(random_state=rng)[0] for _ in range(n)]


# ---- less travelled options, against fixtures from the unmodified reference (oracle/make_golden_traj_options.py) --------
def _options_fixture():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "traj_options.json")) as f:
        return json.load(f)


def _jsonable(value):
    import json
    return json.loads(json.dumps(value, default=lambda o: o.item() if hasattr(o, "item") else str(o)))


_OPTIONS = _options_fixture()


@pytest.mark.parametrize("name", sorted(_OPTIONS["runs"]))
def test_option_trajectory_identical(abi_model, name):
    """x0 / y0 hand-over, no random phase, explicit noise, bounded model queue, xi / kappa, integer-only / log-uniform /
    all-categorical spaces: every asked point, value, minimiser and model count equals the reference's run."""
    from tests._option_scripts import run_config
    assert _jsonable(run_config(skopt_b200, _OPTIONS["runs"][name])) == _OPTIONS["results"][name]


@pytest.mark.parametrize("index", range(len(_OPTIONS["script_cases"])))
def test_optimizer_script_identical(abi_model, index):
    """Batch ask in the random phase, tell(fit=False), constant liar with two strategies and its cache, update_next, run,
    copy, get_result - one script per acquisition (EI, hedge, LCB, EIps, PIps, eta, model queue), step by step."""
    from tests._option_scripts import run_script
    acq, okw = _OPTIONS["script_cases"][index]
    got, want = _jsonable(run_script(skopt_b200, acq, okw)), _OPTIONS["scripts"][index]
    assert len(got) == len(want)
    for step_got, step_want in zip(got, want):
        assert step_got == step_want, step_want[0]


def test_optimizer_pickles_without_its_device_handles(abi_model, tmp_path):
    """test_utils.py::test_dump_and_load_optimizer restated for the GP: an Optimizer with fitted models survives pickle
    and dump / load (device states are dropped and rebuilt on first use) and continues on the same trajectory."""
    import pickle
    from skopt_b200 import Optimizer, dump, load

    def objective(x):
        return x[0] ** 2 + (x[1] == "b")

    opt = Optimizer([(-1.0, 1.0), ["a", "b"]], "GP", n_initial_points=3, acq_optimizer="sampling",
                    acq_optimizer_kwargs=dict(n_points=100), random_state=0)
    for _ in range(5):
        x = opt.ask()
        opt.tell(x, objective(x))
    twin = pickle.loads(pickle.dumps(opt))
    assert twin.ask() == opt.ask() and len(twin.models) == len(opt.models) == 3
    assert "_skb_device_state" not in twin.models[-1].__dict__
    x = opt.ask()
    a, b = opt.tell(x, objective(x)), twin.tell(x, objective(x))
    assert a.x_iters == b.x_iters and opt.ask() == twin.ask()
    dump(opt, str(tmp_path / "opt.pkl"))
    restored = load(str(tmp_path / "opt.pkl"))
    assert restored.ask() == opt.ask()
    mean, std = restored.models[-1].predict(np.array([[0.3, 1.0]]), return_std=True)
    assert mean.shape == std.shape == (1,)


# tests/test_gpu_fit.py: the opt-in GPU fit (objective, lock-step restarts, finalisation, state from the resident factor)
# through the model of skb_fit_* - the host logic of gpr._search_on_device / _finalize_on_device and engine.DeviceFit
from tests import test_gpu_fit as gpu_fit  # noqa: E402

globals().update({"test_fit_" + name[5:]: func for name, func in _mirror(gpu_fit).items()})


def test_verbose_prints_the_reference_progress_lines(abi_model, capsys):
    """verbose=True (benchmarks/bench_branin.py runs with it): the announcements of skopt.callbacks.VerboseCallback."""
    import re
    from skopt_b200 import gp_minimize
    gp_minimize(lambda x: (x[0] - 0.3) ** 2, [(-1.0, 1.0)], n_calls=4, n_initial_points=1, x0=[[0.1]], verbose=True,
                acq_optimizer="sampling", n_points=50, random_state=0)
    lines = [re.sub(r": -?[0-9][0-9.e-]*$", ": #", line) for line in capsys.readouterr().out.splitlines()]
    assert lines[:5] == ["Iteration No: 1 started. Evaluating function at provided point.",
                         "Iteration No: 1 ended. Evaluation done at provided point.",
                         "Time taken: #", "Function value obtained: #", "Current minimum: #"]
    assert lines[5] == "Iteration No: 2 started. Evaluating function at random point."
    # the reference counts the provided point into the random phase as well (base.py:249-274), so iteration 3 is "random"
    assert lines[10] == "Iteration No: 3 started. Evaluating function at random point."
    assert lines[15] == "Iteration No: 4 started. Searching for the next optimal point."
    assert lines[16] == "Iteration No: 4 ended. Search finished for the next optimal point."
    assert len(lines) == 20 and lines[-1] == "Current minimum: #"
