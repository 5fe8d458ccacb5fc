"""Generate tests/golden/*.npz by running the UNMODIFIED reference (skopt 0.9.0 under /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py
Each fixture holds the fitted state exactly as the reference produced it (X_train_, alpha_, K_inv_, L_,
y_train_mean_/std_, kernel_ parameters), the candidate matrix, and the reference outputs of
  kernel_(X, X_train_), kernel_.diag(X)                       (sklearn kernels via skopt kernels)
  GaussianProcessRegressor.predict(X, return_std=True)        (gpr.py:239-343)
  predict(x, return_std, return_mean_grad, return_std_grad)   (gpr.py:345-362, one point at a time)
  _gaussian_acquisition(X, est, y_opt, acq)  for EI / PI / LCB (acquisition.py:20-87)
  gaussian_acquisition_1D(x, est, y_opt, acq)  value+gradient  (acquisition.py:7-17)
  np.argmin / np.argsort of the acquisition values            (optimizer.py:564,570)
TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
if not hasattr(np, "int"):
    np.int = int   # reference uses the removed alias in space/transformers.py:262,275

import scipy, sklearn, skopt  # noqa: E402
from skopt.acquisition import _gaussian_acquisition, gaussian_acquisition_1D  # noqa: E402
from skopt.learning import GaussianProcessRegressor  # noqa: E402
from skopt.learning.gaussian_process.kernels import ConstantKernel, Matern, RBF, WhiteKernel  # noqa: E402
from skopt.utils import cook_estimator  # noqa: E402
from skopt.benchmarks import branin, hart6  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
VERSIONS = dict(skopt=skopt.__version__, sklearn=sklearn.__version__, scipy=scipy.__version__, numpy=np.__version__)


def kernel_params(k):
    out = {}
    for key, v in k.get_params().items():
        if key.endswith("_bounds"):
            continue
        if hasattr(v, "get_params"):
            out[key] = type(v).__name__
        elif isinstance(v, np.ndarray):
            out[key] = v.tolist()
        elif isinstance(v, (tuple, list)):
            out[key] = [float(x) for x in v]
        else:
            out[key] = v if isinstance(v, str) else float(v)
    return out


def emit(name, est, X_train, y, Xc, n_grad, xi=0.01, kappa=1.96, ctor=None):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        est.fit(X_train, y)
    y_opt = float(np.min(y))
    kw = dict(xi=xi, kappa=kappa)
    rec = dict(
        X_train=est.X_train_, y=np.asarray(y, dtype=float), alpha=est.alpha_, K_inv=est.K_inv_, L=est.L_,
        y_train_mean=np.float64(np.ravel(est.y_train_mean_)[0]), y_train_std=np.float64(np.ravel(est.y_train_std_)[0]),
        noise_=np.float64(np.nan if est.noise_ is None else est.noise_),
        Xc=Xc, y_opt=np.float64(y_opt), xi=np.float64(xi), kappa=np.float64(kappa),
        K_trans=est.kernel_(Xc, est.X_train_), diag=est.kernel_.diag(Xc),
    )
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mu, sd = est.predict(Xc, return_std=True)
        rec["mean"], rec["std"] = mu, sd
        rec["mean_only"] = est.predict(Xc)
        for acq in ("EI", "PI", "LCB"):
            v = _gaussian_acquisition(Xc, est, y_opt, acq, acq_func_kwargs=kw)
            rec["acq_" + acq] = v
            rec["argmin_" + acq] = np.int64(np.argmin(v))
            rec["argsort_" + acq] = np.argsort(v)[:10].astype(np.int64)
        d = Xc.shape[1]
        gm, gs = np.empty((n_grad, d)), np.empty((n_grad, d))
        ga = {a: np.empty((n_grad, d)) for a in ("EI", "PI", "LCB")}
        va = {a: np.empty(n_grad) for a in ("EI", "PI", "LCB")}
        grad_ok = True
        try:
            for i in range(n_grad):
                _, _, gm[i], gs[i] = est.predict(Xc[i:i + 1], return_std=True, return_mean_grad=True,
                                                 return_std_grad=True)
                for a in ga:
                    v, g = gaussian_acquisition_1D(Xc[i], est, y_opt, a, kw)
                    va[a][i], ga[a][i] = v[0], g
        except Exception as e:   # e.g. Matern(nu=0.5) inside a Product raises in the reference
            grad_ok = False
            rec["grad_error"] = np.array(repr(e))
        if grad_ok:
            rec["mean_grad"], rec["std_grad"] = gm, gs
            for a in ga:
                rec["acq_grad_" + a], rec["acq_1d_" + a] = ga[a], va[a]
    meta = dict(name=name, versions=VERSIONS, kernel_repr=repr(est.kernel_), kernel_params=kernel_params(est.kernel_),
                kernel_class=type(est.kernel_).__name__, ctor=ctor or {}, n_grad=n_grad if grad_ok else 0)
    rec["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print("%-28s n=%4d d=%2d m=%5d  kernel_=%s" % (name, X_train.shape[0], X_train.shape[1], Xc.shape[0], est.kernel_))


def candidates(rng, m, X_train, lo=0.0, hi=1.0):
    Xc = lo + (hi - lo) * rng.rand(m, X_train.shape[1])
    k = min(4, X_train.shape[0])
    Xc[5:5 + k] = X_train[:k]          # r == 0 corner (test_kernels.py:87-99)
    Xc[12] = Xc[11]                    # duplicated candidate rows -> exact ties
    return Xc


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.RandomState(1234)

    # 1. the synthetic benchmark recipe of SURVEY 8(d) at toy size
    n, d = 48, 5
    X = rng.rand(n, d); w = rng.randn(d); y = np.sin(X @ w) + 0.1 * rng.randn(n)
    emit("matern25_const_noise", GaussianProcessRegressor(
        ConstantKernel(1.3, "fixed") * Matern(np.full(d, 1.5), "fixed", nu=2.5), normalize_y=True, noise=1e-4,
        optimizer=None), X, y, candidates(rng, 300, X), 12,
        ctor=dict(kernel="c*matern", amplitude=1.3, length_scale=[1.5] * d, nu=2.5, normalize_y=True, noise=1e-4))

    # 2. anisotropic Matern 1.5 with an explicit WhiteKernel left in kernel_ (diag keeps the noise level)
    n, d = 40, 3
    X = rng.rand(n, d); y = np.cos(3 * X[:, 0]) + X[:, 1] ** 2 - X[:, 2] + 0.05 * rng.randn(n)
    emit("matern15_explicit_white", GaussianProcessRegressor(
        ConstantKernel(0.8, "fixed") * Matern([0.4, 0.9, 1.7], "fixed", nu=1.5) + WhiteKernel(0.05, "fixed"),
        normalize_y=False, noise=None, optimizer=None), X, y, candidates(rng, 257, X), 12,
        ctor=dict(kernel="c*matern+white", amplitude=0.8, length_scale=[0.4, 0.9, 1.7], nu=1.5, white=0.05,
                  normalize_y=False, noise=None))

    # 3. bare Matern 0.5 (gradient_x works only outside a Product in the reference)
    n, d = 33, 2
    X = rng.rand(n, d); y = np.sin(5 * X[:, 0]) * X[:, 1]
    emit("matern05_bare", GaussianProcessRegressor(Matern(0.6, "fixed", nu=0.5), normalize_y=True, noise=1e-3,
                                                   optimizer=None), X, y, candidates(rng, 130, X), 12,
         ctor=dict(kernel="matern", length_scale=0.6, nu=0.5, normalize_y=True, noise=1e-3))

    # 4. anisotropic RBF, no y normalisation
    n, d = 64, 4
    X = 2 * rng.rand(n, d) - 1; y = np.sum(X ** 2, axis=1) + 0.01 * rng.randn(n)
    emit("rbf_aniso", GaussianProcessRegressor(ConstantKernel(2.0, "fixed") * RBF([0.5, 1.0, 1.5, 2.0], "fixed"),
                                               normalize_y=False, noise=1e-3, optimizer=None),
         X, y, candidates(rng, 400, X, -1, 1), 12,
         ctor=dict(kernel="c*rbf", amplitude=2.0, length_scale=[0.5, 1.0, 1.5, 2.0], normalize_y=False, noise=1e-3))

    # 5. additive constant (bias) + scaled isotropic RBF
    n, d = 30, 2
    X = rng.rand(n, d); y = 3 + np.sin(4 * X[:, 0]) + X[:, 1]
    emit("bias_plus_rbf", GaussianProcessRegressor(ConstantKernel(0.5, "fixed") + ConstantKernel(2.0, "fixed") * RBF(0.7, "fixed"),
                                                   normalize_y=True, noise=1e-2, optimizer=None),
         X, y, candidates(rng, 200, X), 8,
         ctor=dict(kernel="bias+c*rbf", bias=0.5, amplitude=2.0, length_scale=0.7, normalize_y=True, noise=1e-2))

    # 6. the cooked default estimator of gp_minimize (utils.py:366-387) after a real LML fit, Branin
    from skopt.space import Space
    from skopt.utils import normalize_dimensions
    space = Space(normalize_dimensions([(-5.0, 10.0), (0.0, 15.0)]))
    Xo = space.rvs(25, random_state=3)
    y = np.array([branin(x) for x in Xo])
    Xt = space.transform(Xo)
    est = cook_estimator("GP", space=space, random_state=7, noise="gaussian")
    emit("cooked_branin_gaussian", est, Xt, y, candidates(rng, 512, Xt), 12,
         ctor=dict(kernel="cooked", noise="gaussian", random_state=7, dims=2))

    # 7. cooked estimator with the tiny fixed noise used by benchmarks/bench_*.py (ill-conditioned on purpose), Hart6
    space = Space(normalize_dimensions([(0.0, 1.0)] * 6))
    Xo = space.rvs(60, random_state=5)
    y = np.array([hart6(x) for x in Xo])
    Xt = space.transform(Xo)
    est = cook_estimator("GP", space=space, random_state=11, noise=1e-10)
    emit("cooked_hart6_noise1e-10", est, Xt, y, candidates(rng, 512, Xt), 12,
         ctor=dict(kernel="cooked", noise=1e-10, random_state=11, dims=6))

    # 8. a medium, well-conditioned state closer to the bench recipe (n=256, d=10)
    n, d = 256, 10
    X = rng.rand(n, d); w = rng.randn(d); y = np.sin(X @ w) + 0.1 * rng.randn(n)
    emit("matern25_n256_d10", GaussianProcessRegressor(
        ConstantKernel(1.3, "fixed") * Matern(np.full(d, 1.5), "fixed", nu=2.5), normalize_y=True, noise=1e-4,
        optimizer=None), X, y, candidates(rng, 1024, X), 6,
        ctor=dict(kernel="c*matern", amplitude=1.3, length_scale=[1.5] * d, nu=2.5, normalize_y=True, noise=1e-4))


if __name__ == "__main__":
    main()
