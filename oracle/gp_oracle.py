"""CPU oracle for the GP-posterior + acquisition hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy FP64 *restatement* of the reference algorithm
(scikit-optimize 0.9.0 on top of scikit-learn / SciPy / NumPy).  It is the
checker the CUDA path is compared against; it is never the thing shipped or
measured as the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Parity pinning: ``tests/golden/*.npz`` were produced by running the unmodified
reference (``/root/reference``, skopt 0.9.0, scikit-learn 1.9.0, SciPy 1.18.1,
NumPy 2.3.5) through ``oracle/make_golden.py``; ``tests/test_oracle_golden.py``
checks every function below against them.

Third-party arithmetic that the reference reaches (not vendored under
/root/reference; versions above) and that is restated here:
  * sklearn.gaussian_process.kernels  Matern.__call__ (kernels.py:1713-1729),
    RBF.__call__ (:1558-1570), ConstantKernel (:1278-1315), WhiteKernel
    (:1418-1438), Sum (:871,:890), Product (:971,:990)
  * scipy.spatial.distance.cdist 'euclidean' / 'sqeuclidean' (direct
    differences, sequential sum over features)
  * scipy.stats.norm.cdf / pdf  == scipy.special.ndtr, exp(-z^2/2)/sqrt(2 pi)

Reference call sites restated (paths relative to /root/reference):
  * skopt/learning/gaussian_process/gpr.py:299-368     -> predict()
  * skopt/learning/gaussian_process/kernels.py:68-200  -> kernel_gradient_x()
  * skopt/learning/gaussian_process/kernels.py:317-407 -> kernel_cross() for the HammingKernel
  * skopt/acquisition.py:20-87, 90-146, 149-229, 232-321 -> gaussian_*()
  * skopt/optimizer/optimizer.py:557-570               -> score_candidates()
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
from scipy.special import ndtr

KIND_RBF = 0
KIND_MATERN = 1
KIND_HAMMING = 2      # skopt HammingKernel: length_scale holds the per-dimension mismatch weights

_SQRT3 = math.sqrt(3)
_SQRT5 = math.sqrt(5)
_INV_SQRT_2PI = 1.0 / math.sqrt(2 * math.pi)


@dataclass
class GPState:
    """Flat description of a fitted GP: everything predict() reads (gpr.py:315-356)."""
    X_train: np.ndarray            # (n, d)  == est.X_train_
    alpha: np.ndarray              # (n,)    == est.alpha_
    K_inv: np.ndarray              # (n, n)  == est.K_inv_  (gpr.py:223-224)
    length_scale: np.ndarray       # (d,) broadcast of kernel length_scale
    kind: int = KIND_MATERN
    nu: float = 2.5
    amplitude: float = 1.0         # product of ConstantKernels multiplying the base kernel
    bias: float = 0.0              # additive ConstantKernel (k(x,y)=bias)
    white: float = 0.0             # WhiteKernel noise_level still present in kernel_ (diag only)
    y_mean: float = 0.0            # est.y_train_mean_
    y_std: float = 1.0             # est.y_train_std_
    L: Optional[np.ndarray] = None  # (n, n) lower Cholesky factor est.L_
    meta: dict = field(default_factory=dict)

    @property
    def n(self):
        return self.X_train.shape[0]

    @property
    def d(self):
        return self.X_train.shape[1]

    @property
    def prior_var(self):
        """kernel_.diag(x): constant for the stationary grammar (Sum.diag / Product.diag)."""
        return self.amplitude + self.bias + self.white


# --------------------------------------------------------------------------
# kernel tree  ->  (amplitude, bias, white, base)        independent walker
# --------------------------------------------------------------------------
def _reduce_kernel(k):
    """Return dict(a=coef of base, c=const, w=white, base=(kind, nu, length_scale) | None).

    Mirrors the algebra of sklearn Sum/Product for cross-covariances k(X, Y!=X)
    (White contributes 0, kernels.py:1418-1419) and for diag(X) (White
    contributes noise_level, :1438).
    """
    name = type(k).__name__
    if name == "ConstantKernel":
        return dict(a=0.0, c=float(k.constant_value), w=0.0, base=None)
    if name == "WhiteKernel":
        return dict(a=0.0, c=0.0, w=float(k.noise_level), base=None)
    if name == "RBF":
        return dict(a=1.0, c=0.0, w=0.0, base=(KIND_RBF, math.inf, np.asarray(k.length_scale, dtype=float)))
    if name == "Matern":
        return dict(a=1.0, c=0.0, w=0.0, base=(KIND_MATERN, float(k.nu), np.asarray(k.length_scale, dtype=float)))
    if name == "HammingKernel":
        return dict(a=1.0, c=0.0, w=0.0, base=(KIND_HAMMING, math.nan, np.asarray(k.length_scale, dtype=float)))
    if name == "Sum":
        p, q = _reduce_kernel(k.k1), _reduce_kernel(k.k2)
        if p["base"] is not None and q["base"] is not None:
            raise NotImplementedError("sum of two stationary base kernels is outside the hot-path grammar")
        return dict(a=p["a"] + q["a"], c=p["c"] + q["c"], w=p["w"] + q["w"],
                    base=p["base"] if p["base"] is not None else q["base"])
    if name == "Product":
        p, q = _reduce_kernel(k.k1), _reduce_kernel(k.k2)
        if p["base"] is not None and q["base"] is not None:
            raise NotImplementedError("product of two stationary base kernels is outside the hot-path grammar")
        a = p["a"] * q["c"] + p["c"] * q["a"]
        c = p["c"] * q["c"]
        diag = (p["a"] + p["c"] + p["w"]) * (q["a"] + q["c"] + q["w"])
        w = diag - (p["a"] + p["c"]) * (q["a"] + q["c"])
        return dict(a=a, c=c, w=w, base=p["base"] if p["base"] is not None else q["base"])
    raise NotImplementedError("kernel %s is outside the hot-path grammar" % name)


def state_from_estimator(est) -> GPState:
    """Flatten a fitted skopt-style GaussianProcessRegressor (gpr.py:166-237 outputs)."""
    red = _reduce_kernel(est.kernel_)
    X_train = np.ascontiguousarray(est.X_train_, dtype=np.float64)
    d = X_train.shape[1]
    if red["base"] is None:
        kind, nu, ls = KIND_RBF, math.inf, np.ones(d)
        amplitude = 0.0
    else:
        kind, nu, ls = red["base"]
        amplitude = red["a"]
    ls = np.broadcast_to(np.asarray(ls, dtype=np.float64), (d,)).copy()
    return GPState(
        X_train=X_train,
        alpha=np.ascontiguousarray(est.alpha_, dtype=np.float64),
        K_inv=np.ascontiguousarray(est.K_inv_, dtype=np.float64),
        length_scale=ls, kind=kind, nu=nu, amplitude=amplitude, bias=red["c"], white=red["w"],
        y_mean=float(np.ravel(est.y_train_mean_)[0]), y_std=float(np.ravel(est.y_train_std_)[0]),
        L=np.ascontiguousarray(est.L_, dtype=np.float64) if hasattr(est, "L_") else None,
    )


# --------------------------------------------------------------------------
# kernels
# --------------------------------------------------------------------------
def _scaled_sqdist(Xs, Ys):
    """cdist(Xs, Ys, 'sqeuclidean'): direct differences, features summed in order."""
    acc = np.zeros((Xs.shape[0], Ys.shape[0]))
    for k in range(Xs.shape[1]):
        diff = Xs[:, k, None] - Ys[None, :, k]
        acc += diff * diff
    return acc


def base_kernel_from_sqdist(state: GPState, sq):
    """Matern/RBF transform of scaled squared distances (sklearn kernels.py:1722-1731, :1570)."""
    if state.kind == KIND_RBF or state.nu == math.inf:
        return np.exp(-0.5 * sq)
    dists = np.sqrt(sq)
    if state.nu == 0.5:
        return np.exp(-dists)
    if state.nu == 1.5:
        K = dists * _SQRT3
        return (1.0 + K) * np.exp(-K)
    if state.nu == 2.5:
        K = dists * _SQRT5
        return (1.0 + K + K ** 2 / 3.0) * np.exp(-K)
    raise NotImplementedError("Matern nu=%r is outside the hot-path grammar" % state.nu)


def kernel_cross(state: GPState, X):
    """kernel_(X, X_train_): amplitude * base(r) + bias; the White term is 0 off-diagonal."""
    X = np.asarray(X, dtype=np.float64)
    if state.kind == KIND_HAMMING:
        # skopt kernels.py:391-392: exp(-sum_j length_scale_j * [x_j != y_j]) on the unscaled coordinates
        differs = X[:, None, :] != state.X_train[None, :, :]
        K = np.exp(-np.sum(state.length_scale * differs, axis=2))
        return state.amplitude * K + state.bias
    sq = _scaled_sqdist(X / state.length_scale, state.X_train / state.length_scale)
    K = base_kernel_from_sqdist(state, sq)
    return state.amplitude * K + state.bias


def kernel_gradient_x(state: GPState, x):
    """d kernel_(x, X_train_) / dx, shape (n, d)  (skopt kernels.py:68-90, 93-200, 294-308)."""
    x = np.asarray(x, dtype=np.float64)
    if state.kind == KIND_HAMMING:
        raise AttributeError("HammingKernel has no gradient_x (skopt kernels.py:317-407; utils.py:320-330)")
    ls = state.length_scale
    diff = (x - state.X_train) / ls            # (n, d)
    sq = np.sum(diff ** 2, axis=1)
    if state.kind == KIND_RBF or state.nu == math.inf:
        coef = -np.exp(-0.5 * sq)
    else:
        r = np.sqrt(sq)
        if state.nu == 2.5:
            coef = -(5.0 / 3.0) * (1.0 + _SQRT5 * r) * np.exp(-_SQRT5 * r)
        elif state.nu == 1.5:
            coef = -3.0 * np.exp(-_SQRT3 * r)
        elif state.nu == 0.5:
            # reference: -exp(-r)/r * diff/ls, and -1/ls at r == 0 (kernels.py:118-128)
            g = np.empty_like(diff)
            nz = r != 0.0
            g[nz] = (-np.exp(-r[nz]) / r[nz])[:, None] * diff[nz] / ls
            g[~nz] = -1.0 / ls
            return state.amplitude * g
        else:
            raise NotImplementedError
    return state.amplitude * coef[:, None] * diff / ls


# --------------------------------------------------------------------------
# posterior  (gpr.py:299-368)
# --------------------------------------------------------------------------
def quad_form(K_trans, K_inv, how="einsum"):
    """k^T K_inv k per row.  'einsum' is the reference's own contraction (gpr.py:332)."""
    if how == "einsum":
        return np.einsum("ki,kj,ij->k", K_trans, K_trans, K_inv)
    if how == "gemm":
        return np.einsum("ki,ki->k", K_trans @ K_inv, K_trans)
    raise ValueError(how)


def predict(state: GPState, X, return_std=False, return_mean_grad=False, return_std_grad=False,
            quad="gemm", chunk=4096):
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("X must be 2-dimensional")
    if return_std_grad and not return_std:
        raise ValueError("Not returning std_gradient without returning the std.")
    if X.shape[0] != 1 and (return_mean_grad or return_std_grad):
        raise ValueError("Not implemented for n_samples > 1")
    m = X.shape[0]
    y_mean = np.empty(m)
    y_std = np.empty(m) if return_std else None
    K_last = None
    for s in range(0, m, chunk):
        K_trans = kernel_cross(state, X[s:s + chunk])
        y_mean[s:s + chunk] = state.y_std * K_trans.dot(state.alpha) + state.y_mean
        if return_std:
            y_var = np.full(K_trans.shape[0], state.prior_var)
            y_var -= quad_form(K_trans, state.K_inv, quad)
            y_var[y_var < 0] = 0.0
            y_var = y_var * state.y_std ** 2
            y_std[s:s + chunk] = np.sqrt(y_var)
        K_last = K_trans
    if return_mean_grad:
        grad = kernel_gradient_x(state, X[0])
        grad_mean = grad.T.dot(state.alpha) * state.y_std
        if return_std_grad:
            grad_std = np.zeros(X.shape[1])
            if not np.allclose(y_std, grad_std):
                grad_std = -np.dot(K_last, np.dot(state.K_inv, grad))[0] / y_std
                grad_std = grad_std * state.y_std ** 2
            return y_mean, y_std, grad_mean, grad_std
        if return_std:
            return y_mean, y_std, grad_mean
        return y_mean, grad_mean
    if return_std:
        return y_mean, y_std
    return y_mean


def predict_with_grads_batched(state: GPState, X, quad="gemm"):
    """Loop of single-point reference calls: the oracle for the batched-gradient extension."""
    X = np.asarray(X, dtype=np.float64)
    m, d = X.shape
    mu, sd = np.empty(m), np.empty(m)
    gmu, gsd = np.empty((m, d)), np.empty((m, d))
    for i in range(m):
        a, b, c, e = predict(state, X[i:i + 1], True, True, True, quad=quad)
        mu[i], sd[i], gmu[i], gsd[i] = a[0], b[0], c, e
    return mu, sd, gmu, gsd


# --------------------------------------------------------------------------
# acquisition  (acquisition.py)
# --------------------------------------------------------------------------
def _norm_pdf(z):
    return np.exp(-z ** 2 / 2.0) * _INV_SQRT_2PI


def lcb_from_posterior(mu, std, kappa=1.96, mu_grad=None, std_grad=None):
    if mu_grad is not None:
        if kappa == "inf":
            return -std, -std_grad
        return mu - kappa * std, mu_grad - kappa * std_grad
    if kappa == "inf":
        return -std
    return mu - kappa * std


def pi_from_posterior(mu, std, y_opt=0.0, xi=0.01, mu_grad=None, std_grad=None):
    values = np.zeros_like(mu)
    mask = std > 0
    improve = y_opt - xi - mu[mask]
    scaled = improve / std[mask]
    values[mask] = ndtr(scaled)
    if mu_grad is not None:
        if not np.all(mask):
            return values, np.zeros_like(std_grad)
        improve_grad = (-mu_grad * std - std_grad * improve) / std ** 2
        return values, improve_grad * _norm_pdf(scaled)
    return values


def ei_from_posterior(mu, std, y_opt=0.0, xi=0.01, mu_grad=None, std_grad=None):
    values = np.zeros_like(mu)
    mask = std > 0
    improve = y_opt - xi - mu[mask]
    scaled = improve / std[mask]
    cdf = ndtr(scaled)
    pdf = _norm_pdf(scaled)
    values[mask] = improve * cdf + std[mask] * pdf
    if mu_grad is not None:
        if not np.all(mask):
            return values, np.zeros_like(std_grad)
        improve_grad = (-mu_grad * std - std_grad * improve) / std ** 2
        cdf_grad = improve_grad * pdf
        pdf_grad = -improve * cdf_grad
        grad = (-mu_grad * cdf - pdf_grad) + (std_grad * pdf + pdf_grad)
        return values, grad
    return values


def gaussian_acquisition(X, state: GPState, y_opt=None, acq_func="LCB", return_grad=False,
                         acq_func_kwargs=None, quad="gemm"):
    """_gaussian_acquisition (acquisition.py:20-87): everything is returned as something to minimise."""
    X = np.asarray(X)
    if X.ndim != 2:
        raise ValueError("X is {}-dimensional, however, it must be 2-dimensional.".format(X.ndim))
    kw = acq_func_kwargs or {}
    xi = kw.get("xi", 0.01)
    kappa = kw.get("kappa", 1.96)
    if return_grad:
        mu, std, gm, gs = predict(state, X, True, True, True, quad=quad)
    else:
        mu, std = predict(state, X, True, quad=quad)
        gm = gs = None
    if acq_func == "LCB":
        return lcb_from_posterior(mu, std, kappa, gm, gs)
    if acq_func == "EI":
        out = ei_from_posterior(mu, std, y_opt, xi, gm, gs)
    elif acq_func == "PI":
        out = pi_from_posterior(mu, std, y_opt, xi, gm, gs)
    else:
        raise ValueError("Acquisition function not implemented.")
    if return_grad:
        return -out[0], -out[1]
    return -out


def gaussian_acquisition_1D(x, state, y_opt=None, acq_func="LCB", acq_func_kwargs=None, return_grad=True):
    return gaussian_acquisition(np.expand_dims(x, 0), state, y_opt, acq_func, return_grad, acq_func_kwargs)


def score_candidates(X, state: GPState, y_opt, acq_funcs=("EI",), acq_func_kwargs=None, top_k=1, quad="gemm"):
    """Scoring block of Optimizer._tell (optimizer.py:556-570): values, argmin and top-k per acquisition."""
    out = {}
    for acq in acq_funcs:
        v = gaussian_acquisition(X, state, y_opt, acq, False, acq_func_kwargs, quad=quad)
        order = np.lexsort((np.arange(v.size), v))[:top_k]   # (value, index) ascending, ties -> low index
        out[acq] = dict(values=v, argmin=int(np.argmin(v)), topk=order)
    return out


# --------------------------------------------------------------------------
# synthetic benchmark state, built without the reference
# --------------------------------------------------------------------------
def make_synthetic_state(n, d, noise=1e-4, seed=0, nu=2.5, amplitude=1.3, length_scale=1.5,
                         kind=KIND_MATERN, normalize_y=True):
    """Same recipe as the fit of GaussianProcessRegressor(optimizer=None, normalize_y=True, noise=noise)
    (sklearn _gpr.py:275-280, 349-367; skopt gpr.py:223-235), written out with NumPy/SciPy."""
    from scipy.linalg import cholesky, cho_solve, solve_triangular
    rng = np.random.RandomState(seed)
    X_train = rng.rand(n, d)
    w = rng.randn(d)
    y = np.sin(X_train @ w) + 0.1 * rng.randn(n)
    if normalize_y:
        y_mean, y_std = float(np.mean(y)), float(np.std(y))
        yn = (y - y_mean) / y_std
    else:
        y_mean, y_std, yn = 0.0, 1.0, y
    ls = np.full(d, float(length_scale))
    st = GPState(X_train=X_train, alpha=np.zeros(n), K_inv=np.zeros((1, 1)), length_scale=ls, kind=kind,
                 nu=nu, amplitude=amplitude, bias=0.0, white=0.0, y_mean=y_mean, y_std=y_std)
    Xs = X_train / ls
    K = amplitude * base_kernel_from_sqdist(st, _scaled_sqdist(Xs, Xs))
    K[np.diag_indices_from(K)] = amplitude + noise + 1e-10   # kernel diag + White + sklearn's alpha=1e-10
    L = cholesky(K, lower=True)
    st.alpha = cho_solve((L, True), yn)
    L_inv = solve_triangular(L.T, np.eye(n))
    st.K_inv = L_inv.dot(L_inv.T)
    st.L = L
    st.meta = dict(y=y, noise=noise, seed=seed, rng=rng)
    return st


# --------------------------------------------------------------------------
# fit objective: log marginal likelihood and gradient
# --------------------------------------------------------------------------
def log_marginal_likelihood(X_train, y_train, nu, theta_full, alpha=1e-10):
    """sklearn GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True) (_gpr.py:541-657) for the
    kernel  amplitude * {RBF (nu=inf) | Matern(nu)}(length_scale[d]) + WhiteKernel(noise)  with
    theta_full = [log amplitude, log length_scale[d], log noise] (log noise = -inf: no noise term).

    Kernel gradients follow sklearn kernels.py: ConstantKernel :1296-1307 (dK/dlog c = c), Product :975-983,
    Sum :876-880, WhiteKernel :1406-1416 (noise * I), RBF :1578-1589 (K * D), Matern :1737-1763
    (nu=1/2: K * D / r with 0 where r = 0; 3/2: 3 D exp(-sqrt(3) r); 5/2: 5/3 D (sqrt(5) r + 1) exp(-sqrt(5) r)),
    D = (x_i - x_j)^2 / l^2 per dimension.  Returns (lml, grad[d + 2]); (-inf, zeros) when K is not positive definite.
    The n x n x (d + 2) gradient tensor of the reference is formed here too (small cases only)."""
    from scipy.linalg import cho_solve, cholesky
    X = np.asarray(X_train, dtype=np.float64)
    y = np.asarray(y_train, dtype=np.float64)
    theta_full = np.asarray(theta_full, dtype=np.float64)
    n, d = X.shape
    amplitude, ls, noise = np.exp(theta_full[0]), np.exp(theta_full[1:d + 1]), np.exp(theta_full[d + 1])
    Xs = X / ls
    sq = _scaled_sqdist(Xs, Xs)
    D = (X[:, None, :] - X[None, :, :]) ** 2 / (ls ** 2)              # kernels.py:1738-1741
    if nu == math.inf:
        base = np.exp(-0.5 * sq)
        G = base[..., None] * D
    else:
        r = np.sqrt(sq)
        if nu == 0.5:
            base = np.exp(-r)
            G = np.zeros_like(D)
            np.divide(D, r[..., None], out=G, where=r[..., None] != 0)
            G = base[..., None] * G
        elif nu == 1.5:
            t = r * _SQRT3
            base = (1.0 + t) * np.exp(-t)
            G = 3.0 * D * np.exp(-t)[..., None]
        elif nu == 2.5:
            t = r * _SQRT5
            base = (1.0 + t + t ** 2 / 3.0) * np.exp(-t)
            G = 5.0 / 3.0 * D * ((t + 1.0) * np.exp(-t))[..., None]
        else:
            raise NotImplementedError("Matern nu=%r" % nu)
    K = amplitude * base + noise * np.eye(n)
    K_gradient = np.dstack([(amplitude * base)[..., None], amplitude * G, (noise * np.eye(n))[..., None]])
    K[np.diag_indices_from(K)] += alpha
    try:
        L = cholesky(K, lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        return -np.inf, np.zeros(d + 2)
    a = cho_solve((L, True), y, check_finite=False)
    lml = -0.5 * y.dot(a) - np.log(np.diag(L)).sum() - n / 2.0 * np.log(2.0 * np.pi)
    inner = np.outer(a, a) - cho_solve((L, True), np.eye(n), check_finite=False)
    grad = 0.5 * np.einsum("ij,jik->k", inner, K_gradient)
    return lml, grad
