"""CPU tests: C-ABI surface, host-side flattening of fitted models, top-k merge (incl. 2-rank gloo), L-BFGS batching."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
from scipy.optimize import fmin_l_bfgs_b

import skopt_b200
from skopt_b200 import _lib, engine
from skopt_b200.distributed import merge_topk, shard_bounds
from skopt_b200.optimizer.lbfgs import minimize_batch
from oracle import gp_oracle as orc
from tests._golden import NAMES, kernel_code, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/skb.h is the contract: each declared function must be exported by libskb.so and bound in _lib.py."""
    header = open(os.path.join(ROOT, "include", "skb.h")).read()
    declared = set(re.findall(r"\b(skb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound
    assert _lib.lib().skb_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    st = orc.make_synthetic_state(16, 2)
    hs = engine.HostState(st.X_train, st.alpha, np.eye(16), st.length_scale, 3, 1.0, 0.0, 0.0, 0.0, 1.0)
    with pytest.raises(_lib.SkbError) as e:
        engine.DeviceState(hs)
    assert e.value.code == -3


@pytest.mark.parametrize("name", NAMES)
def test_kernel_flattening_matches_oracle_walker(name):
    """Two independent walkers (product: engine._flatten_kernel, checker: oracle._reduce_kernel) must agree, and the
    golden K_trans / diag pin both to the reference."""
    rec, meta, est = load(name)
    hs = engine.host_state_from_estimator(est)
    st = orc.state_from_estimator(est)
    kind = kernel_code(st)
    assert hs.kernel == kind
    assert hs.amplitude == pytest.approx(st.amplitude, rel=1e-15) and hs.bias == pytest.approx(st.bias, rel=1e-15)
    assert hs.white == pytest.approx(st.white, rel=1e-15, abs=1e-300)
    assert np.array_equal(hs.length_scale, st.length_scale)
    assert np.allclose(hs.prior_var, rec["diag"], rtol=1e-15)
    assert np.allclose(hs.L_inv.T @ hs.L_inv, rec["K_inv"], rtol=1e-6, atol=1e-8 * np.abs(rec["K_inv"]).max())


def test_unsupported_kernels_raise():
    from sklearn.gaussian_process import kernels as skk

    class E:
        pass
    e = E()
    e.X_train_ = np.zeros((3, 2)); e.alpha_ = np.zeros(3); e.L_ = np.eye(3)
    e.y_train_mean_ = 0.0; e.y_train_std_ = 1.0
    for k in (skk.RationalQuadratic(), skk.RBF() * skk.Matern(), skk.RBF() + skk.Matern(), skk.Matern(nu=3.5),
              skk.DotProduct()):
        e.kernel_ = k
        with pytest.raises(NotImplementedError):
            engine.host_state_from_estimator(e)


def test_product_fit_reproduces_reference_state():
    """fit() post-processing (gpr.py:182-237): noise_, K_inv_, zeroed WhiteKernel, y statistics."""
    from skopt_b200.learning import GaussianProcessRegressor
    from skopt_b200.learning.gaussian_process.kernels import ConstantKernel, Matern
    rec, meta, est = load("matern25_const_noise")
    c = meta["ctor"]
    gpr = GaussianProcessRegressor(ConstantKernel(c["amplitude"], "fixed") * Matern(np.array(c["length_scale"]), "fixed",
                                                                                     nu=c["nu"]),
                                   normalize_y=c["normalize_y"], noise=c["noise"], optimizer=None)
    gpr.fit(rec["X_train"], rec["y"])
    assert gpr.noise_ == float(rec["noise_"])
    assert np.allclose(gpr.alpha_, rec["alpha"], rtol=1e-9, atol=1e-12 * np.abs(rec["alpha"]).max())
    assert np.allclose(gpr.K_inv_, rec["K_inv"], rtol=1e-7, atol=1e-9 * np.abs(rec["K_inv"]).max())
    assert float(gpr.y_train_mean_) == float(rec["y_train_mean"]) and float(gpr.y_train_std_) == float(rec["y_train_std"])
    hs = engine.host_state_from_estimator(gpr)
    assert hs.white == 0.0 and hs.amplitude == pytest.approx(c["amplitude"])
    import pickle
    from sklearn.base import clone
    pickle.loads(pickle.dumps(gpr))
    assert clone(gpr).noise == c["noise"]


def test_merge_topk_is_lexicographic():
    import torch
    v = torch.tensor([0.5, -1.0, -1.0, float("nan"), 0.0, -0.0, 3.0], dtype=torch.float64)
    i = torch.tensor([10, 7, 3, 1, 5, 4, -1], dtype=torch.int64)
    ov, oi = merge_topk(v, i, 5)
    assert oi.tolist() == [3, 7, 4, 5, 10]
    ov, oi = merge_topk(v, i, 7)
    assert oi.tolist()[-2:] == [-1, -1] and torch.isnan(ov[-1])
    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import skopt_b200
from skopt_b200.distributed import merge_topk, merge_topk_allgather, shard_bounds
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.RandomState(0)
vals = rng.randn(1001); vals[17] = vals[900] = vals.min() - 1.0          # an exact tie across ranks
b = shard_bounds(vals.size, world)
lo, hi = b[rank], b[rank + 1]
k = 6
local = np.lexsort((np.arange(lo, hi), vals[lo:hi]))[:k]
v, i = merge_topk_allgather(torch.from_numpy(vals[lo:hi][local]), torch.from_numpy(local + lo), k)
expect = np.lexsort((np.arange(vals.size), vals))[:k]
assert i.tolist() == expect.tolist(), (i.tolist(), expect.tolist())
assert i[0].item() == int(np.argmin(vals)) == 17

# sharded_score: the N>1 host logic with a stand-in for the device (same interface as engine.DeviceState.score)
from skopt_b200.distributed import sharded_score
class FakeDev:
    device = 0
    def score(self, X, y_opt=None, acq_funcs=("EI",), xi=0.01, kappa=1.96, top_k=0, index_offset=0, **kw):
        assert kw.get("precision", "fp64") in ("fp64", "split_int8", "auto")
        out = {}
        for j, a in enumerate(acq_funcs):
            v = np.sin(7 * X[:, 0] + j) * X[:, 1]
            order = np.lexsort((np.arange(len(v)), v))[:top_k]
            out[a] = dict(topk_val=v[order], topk_idx=order + index_offset, first_nan=-1)
        return out
X = np.random.RandomState(1).rand(777, 2); X[600] = X[5]
full = FakeDev().score(X, acq_funcs=("EI", "PI"), top_k=5)
sh = sharded_score(FakeDev(), X, y_opt=0.0, acq_funcs=("EI", "PI"), xi=0.01, kappa=1.96, top_k=5)
for a in ("EI", "PI"):
    assert np.array_equal(full[a]["topk_idx"], sh[a]["topk_idx"]) and np.array_equal(full[a]["topk_val"], sh[a]["topk_val"])
# packed records: every acquisition's top-k + first NaN in ONE all-gather; rank 1's shard is EMPTY (padding record)
from skopt_b200.distributed import exchange_records, merge_records_torch
from skopt_b200.engine import unpack_record, record_words
A, k2 = 3, 4
assert record_words(A, k2) == 27
rec = np.empty(27, dtype=np.int64)
if rank == 0:
    v = np.array([[0.5, 0.5, 2.0, np.nan], [-1.0, 0.0, 1.0, 2.0], [3.0, 4.0, 5.0, 6.0]])
    i = np.array([[4, 9, 2, -1], [0, 1, 2, 3], [7, 6, 5, 8]])
    nan = np.array([11, -1, -1])
else:
    v = np.full((A, k2), np.nan); i = np.full((A, k2), -1); nan = np.array([-1, -1, -1])
rec[:12] = v.reshape(-1).view(np.int64); rec[12:24] = i.reshape(-1); rec[24:] = nan
got = unpack_record(exchange_records(torch.from_numpy(rec), A, k2), ("EI", "PI", "LCB"), k2)
assert got["EI"]["topk_idx"].tolist() == [4, 9, 2, -1] and got["EI"]["first_nan"] == 11
assert got["PI"]["topk_idx"].tolist() == [0, 1, 2, 3] and got["PI"]["first_nan"] == -1
assert got["LCB"]["topk_idx"].tolist() == [7, 6, 5, 8] and got["LCB"]["topk_val"].tolist() == [3.0, 4.0, 5.0, 6.0]
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_topk_allgather_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_lockstep_lbfgs_equals_sequential_runs():
    """Batching must not change any search: same iterates as running fmin_l_bfgs_b one start at a time."""
    A = np.diag([1.0, 4.0, 9.0])
    c = np.array([0.3, 0.9, 0.1])

    def f_and_g(x):
        r = x - c
        return float(r @ A @ r + np.sin(3 * x[0])), 2 * A @ r + np.array([3 * np.cos(3 * x[0]), 0.0, 0.0])

    calls = []

    def evaluate(tasks, X):
        calls.append(len(tasks))
        out = [f_and_g(x) for x in X]
        return np.array([o[0] for o in out]), np.vstack([o[1] for o in out])

    from skopt_b200.optimizer import lbfgs
    starts = np.random.RandomState(3).rand(7, 3)
    starts[2] = [0.0, 1.0, 0.5]                        # a start on the boundary
    bounds = [(0.0, 1.0)] * 3
    drivers = ["threads"] + (["stepper"] if lbfgs._SETULB is not None else [])
    assert "stepper" in drivers, "SciPy's compiled stepper should be usable with the SciPy in this image"
    for driver in drivers:
        del calls[:]
        for maxiter in (20, 3):
            res = minimize_batch(starts, evaluate, bounds, maxiter=maxiter, driver=driver)
            for x0, (x, f, info) in zip(starts, res):
                xs, fs, infos = fmin_l_bfgs_b(f_and_g, x0, bounds=bounds, approx_grad=False, maxiter=maxiter)
                assert np.array_equal(x, xs) and f == fs, driver
                assert info["funcalls"] == infos["funcalls"] and info["nit"] == infos["nit"], driver
        assert calls[0] == 7 and max(calls) == 7      # all searches advance together


def test_optimizer_host_logic_and_initial_points():
    """Validation and the random initial phase need no GPU; the points must equal the reference's (same RNG order)."""
    import json
    from skopt_b200 import Optimizer
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "traj_gp_minimize.npz")))
    runs = json.loads(str(gold["runs"]))
    cfg = runs["hart6_lbfgs_ei"]
    rng = np.random.RandomState(cfg["random_state"])
    from skopt_b200.utils import cook_estimator
    dims = [tuple(d) for d in cfg["dims"]]
    est = cook_estimator("GP", space=dims, random_state=rng.randint(0, np.iinfo(np.int32).max))   # gp.py:254-257
    opt = Optimizer(dims, est, n_initial_points=cfg["n_initial_points"], random_state=rng)
    ref = json.loads(str(gold["hart6_lbfgs_ei_x_iters"]))
    for i in range(cfg["n_initial_points"] - 1):
        x = opt.ask()
        assert x == ref[i]
        opt.tell(x, float(i))
    assert opt.acq_optimizer == "lbfgs" and opt.cand_acq_funcs_ == ["EI", "LCB", "PI"] and len(opt.models) == 0
    with pytest.raises(ValueError):
        Optimizer([(0.0, 1.0)], "GP", acq_func="nope")
    with pytest.raises(ValueError):
        Optimizer([(0.0, 1.0)], "GP", acq_optimizer="newton")
    with pytest.raises(ValueError):
        opt.tell([2.0] * 6, 1.0)                         # outside the space
    with pytest.raises(ValueError):
        opt.tell([0.5] * 6, "bad")
    with pytest.raises(ValueError):
        opt.ask(n_points=0)
    with pytest.raises(ValueError):
        opt.ask(n_points=2, strategy="cl_median")


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """The boundary is a C ABI: include/skb.h must compile as C99 (no C++ or torch types) and a C program must link
    against libskb.so and reach the entry points that need no GPU."""
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "skb.h"\n'
                   'int main(void) {\n'
                   '    skb_state* st = NULL;\n'
                   '    skb_state_desc desc;\n'
                   '    memset(&desc, 0, sizeof desc);\n'
                   '    int rc = skb_state_create(&desc, 0, &st);          /* n_train = 0: refused before any CUDA call */\n'
                   '    printf("%d %d %d %s\\n", skb_abi_version(), SKB_ABI_VERSION, rc, skb_last_error());\n'
                   '    return (skb_abi_version() == SKB_ABI_VERSION && rc == SKB_ERR_INVALID && st == NULL) ? 0 : 1;\n'
                   '}\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    build = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                            str(src), "-o", str(exe), "-L", libdir, "-lskb", "-Wl,-rpath," + libdir],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "n_train and n_dims must be positive" in run.stdout


def test_argument_validation_needs_no_device():
    """Error behaviour of the C ABI (include/skb.h: negative skb_status + skb_last_error, nothing thrown): arguments are
    validated before any CUDA call, so these refusals are observable on the build machine through the real library."""
    import ctypes as C
    L = _lib.lib()
    launches_before = L.skb_launch_count()

    def last():
        return L.skb_last_error().decode()

    handle = C.c_void_p()
    assert L.skb_state_create(None, 0, C.byref(handle)) == -1 and "NULL" in last()
    desc = _lib.StateDesc()
    assert L.skb_state_create(C.byref(desc), 0, C.byref(handle)) == -1 and "must be positive" in last()
    x = np.ones((2, 3)); a = np.ones(2); li = np.eye(2); ls = np.ones(3)
    desc.n_train, desc.n_dims, desc.kernel = 2, 3, 3
    assert L.skb_state_create(C.byref(desc), 0, C.byref(handle)) == -1 and "must not be NULL" in last()
    desc.length_scale, desc.X_train, desc.alpha, desc.L_inv = ls.ctypes.data, x.ctypes.data, a.ctypes.data, li.ctypes.data
    desc.kernel = 9
    assert L.skb_state_create(C.byref(desc), 0, C.byref(handle)) == -2 and "outside the hot-path grammar" in last()
    desc.kernel, desc.n_dims = 3, 101
    assert L.skb_state_create(C.byref(desc), 0, C.byref(handle)) == -2 and "n_dims 101 > 100" in last()
    desc.n_dims = 3
    ls[1] = 0.0
    assert L.skb_state_create(C.byref(desc), 0, C.byref(handle)) == -1 and "length_scale[1] must be > 0" in last()
    assert handle.value is None
    # entry points that take a state refuse NULL handles; the stateless ones refuse bad sizes
    prm, out = _lib.AcqParams(), _lib.ScoreOut()
    assert L.skb_score_host(None, x.ctypes.data, 2, 0, C.byref(prm), C.byref(out)) == -1
    assert L.skb_predict_mean_host(None, x.ctypes.data, 2, a.ctypes.data) == -1
    assert L.skb_state_attach_kinv(None, li.ctypes.data) == -1
    assert L.skb_state_sync(None) == -1 and L.skb_state_set_timing(None, 1) == -1
    tv, ti = np.empty(4), np.empty(4, dtype=np.int64)
    assert L.skb_topk_host(0, a.ctypes.data, 2, 0, 0, tv.ctypes.data, ti.ctypes.data, None) == -1 and "top_k" in last()
    assert L.skb_topk_host(0, a.ctypes.data, -1, 0, 2, tv.ctypes.data, ti.ctypes.data, None) == -1
    acq = (C.c_void_p * 3)()
    assert L.skb_acq_from_posterior_host(0, -1, 3, a.ctypes.data, a.ctypes.data, None, None, C.byref(prm), acq, acq) == -1
    assert L.skb_acq_from_posterior_host(0, 2, 3, None, a.ctypes.data, None, None, C.byref(prm), acq, acq) == -1
    L.skb_state_destroy(None)                                   # destroying nothing is allowed
    assert L.skb_launch_count() == launches_before              # nothing was launched by any of this


def test_kernels_outside_the_grammar_load_but_have_no_device_form():
    """skopt.learning.gaussian_process.kernels exports RationalQuadratic, ExpSineSquared, DotProduct and Exponentiation;
    here they import, combine, clone and fit on the host, and predict / gradient_x refuse them by name (no CPU route)."""
    from sklearn.base import clone
    from skopt_b200.learning import GaussianProcessRegressor
    from skopt_b200.learning.gaussian_process import kernels as K
    combo = 2.0 * K.RationalQuadratic(1.0, 0.5) + K.DotProduct(1.0)
    assert type(combo).__module__ == K.__name__ and isinstance(K.RBF(1.0) ** 2, K.Exponentiation)
    assert repr(clone(combo)) == repr(combo)
    rng = np.random.RandomState(0)
    for kernel in (K.RationalQuadratic(1.0), K.ExpSineSquared(1.0, 2.0), K.DotProduct(1.0), K.RBF(1.0) ** 2, combo):
        est = GaussianProcessRegressor(kernel, optimizer=None).fit(rng.rand(5, 2), rng.rand(5))
        with pytest.raises(NotImplementedError, match="outside the GPU hot-path grammar"):
            est.predict(rng.rand(3, 2), return_std=True)
        with pytest.raises(NotImplementedError, match="outside the GPU hot-path grammar"):
            kernel.gradient_x(np.zeros(2), np.zeros((3, 2)))


def test_benchmark_objectives_have_their_known_minima():
    from skopt_b200 import benchmarks as B
    for x in ([-np.pi, 12.275], [np.pi, 2.275], [9.42478, 2.475]):
        assert abs(B.branin(x) - 0.397887) < 1e-5
    assert abs(B.hart6([0.20169, 0.15001, 0.476874, 0.275332, 0.311652, 0.6573]) + 3.32237) < 1e-5
    assert B.bench1([0.0]) == 0.0 and B.bench2([5.0]) == -5.0 and B.bench1_with_time([2.0]) == (4.0, 2.22)
    assert abs(B.bench3([-0.3]) + 0.9) < 0.1 and B.bench4(["3"]) == 9.0 and B.bench5(["3", 4]) == 25.0


def test_default_precision_setter_and_estimator_override():
    """engine.set_default_precision: the precision of the calls without an argument for it (predict(return_std=True),
    gaussian_ei, an Optimizer without acq_optimizer_kwargs['precision']); an estimator attribute overrides it."""
    from skopt_b200 import engine

    class Est:
        precision = None

    old = engine.set_default_precision("fp64")
    try:
        assert old == "auto"
        assert engine.resolve_precision(None, 2048, 10 ** 6) == "fp64"
        assert engine.default_precision(Est()) == "fp64"
        e = Est()
        e.precision = "split_int8_p7"
        assert engine.default_precision(e) == "split_int8_p7"
        with pytest.raises(ValueError):
            engine.set_default_precision("fp16")
    finally:
        engine.set_default_precision(old)
    assert engine.resolve_precision(None, 2048, 10 ** 6) == "split_int8"
    assert engine.resolve_precision(None, 100, 10 ** 6) == "fp64" and engine.resolve_precision("auto", 2048, 100) == "fp64"
