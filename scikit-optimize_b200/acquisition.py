"""Acquisition functions evaluated on the GPU; same names, signatures, sign conventions and exceptions as
skopt/acquisition.py (gaussian_acquisition_1D :7-17, _gaussian_acquisition :20-87, gaussian_lcb :90-146,
gaussian_pi :149-229, gaussian_ei :232-321).

Two device routes, no host arithmetic:
  * `model` is this package's fitted GaussianProcessRegressor -> ONE fused scoring call (posterior + acquisition
    [+ gradients]) on the resident GP state (skb_score_host / skb_score_grad_host);
  * any other regressor exposing `predict(X, return_std=True[, return_mean_grad, return_std_grad])` (the reference's
    duck-typed boundary) -> its own predict(), then the acquisition / chain-rule arithmetic on the GPU
    (skb_acq_from_posterior_host).
Batched gradients are an extension: the reference accepts one row when return_grad=True; here any number of rows
is accepted and one row returns the reference's shapes.
"""
import ctypes as C
import warnings

import numpy as np

from . import _lib, engine
from .engine import _acq_params


def _is_device_model(model):
    return hasattr(model, "device_state") and hasattr(model, "X_train_")


def _default_device():
    from .learning.gaussian_process.gpr import _DEFAULT_DEVICE
    return _DEFAULT_DEVICE[0]


def _check_X(X):
    X = np.asarray(X)
    if X.ndim != 2:
        raise ValueError("X is {}-dimensional, however, it must be 2-dimensional.".format(X.ndim))
    return X


def _from_foreign_model(X, model, acq, y_opt, xi, kappa, return_grad):
    """model.predict on the model's own terms, acquisition arithmetic on the GPU."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if return_grad:
            mu, std, mu_grad, std_grad = model.predict(X, return_std=True, return_mean_grad=True, return_std_grad=True)
        else:
            mu, std = model.predict(X, return_std=True)
            mu_grad = std_grad = None
    mu, std = np.asarray(mu, dtype=np.float64), np.asarray(std, dtype=np.float64)
    if acq != "LCB" and (mu.ndim != 1 or std.ndim != 1):
        raise ValueError("mu and std are {}-dimensional and {}-dimensional, however both must be 1-dimensional. "
                         "Did you train your model with an (N, 1) vector instead of an (N,) vector?"
                         .format(mu.ndim, std.ndim))
    mu, std = np.ascontiguousarray(mu.ravel()), np.ascontiguousarray(std.ravel())
    m = mu.shape[0]
    d = X.shape[1]
    prm = _acq_params(y_opt, xi, kappa, (acq,), 0)
    a = _lib.ACQ_INDEX[acq]
    vals = np.empty(m)
    acq_ptrs = (C.c_void_p * 3)()
    grad_ptrs = (C.c_void_p * 3)()
    acq_ptrs[a] = vals.ctypes.data
    gm = gs = grad = None
    if return_grad:
        gm = np.ascontiguousarray(np.asarray(mu_grad, dtype=np.float64).reshape(m, d))
        gs = np.ascontiguousarray(np.asarray(std_grad, dtype=np.float64).reshape(m, d))
        grad = np.empty((m, d))
        grad_ptrs[a] = grad.ctypes.data
    _lib.check(_lib.lib().skb_acq_from_posterior_host(
        _default_device(), m, d, mu.ctypes.data, std.ctypes.data,
        gm.ctypes.data if gm is not None else None, gs.ctypes.data if gs is not None else None,
        C.byref(prm), acq_ptrs, grad_ptrs))
    if return_grad:
        return vals, (grad[0] if m == 1 else grad)
    return vals


def _evaluate(X, model, acq, y_opt, xi, kappa, return_grad):
    """Value(s) to MINIMISE (-EI, -PI, LCB) [and gradient]."""
    X = _check_X(X)
    if not _is_device_model(model):
        return _from_foreign_model(X, model, acq, y_opt, xi, kappa, return_grad)
    dev = model.device_state()
    if return_grad:
        g = dev.score_grad(X, y_opt=y_opt, acq_funcs=(acq,), xi=xi, kappa=kappa)
        grad = g[acq]["grad"]
        return g[acq]["values"], (grad[0] if X.shape[0] == 1 else grad)
    res = dev.score(X, y_opt=y_opt, acq_funcs=(acq,), xi=xi, kappa=kappa, want_posterior=False,
                    precision=engine.default_precision(model))
    return res[acq]["values"]


def gaussian_lcb(X, model, kappa=1.96, return_grad=False):
    """Lower confidence bound mu - kappa*std (kappa='inf': -std), acquisition.py:90-146."""
    return _evaluate(X, model, "LCB", None, 0.01, kappa, return_grad)


def gaussian_pi(X, model, y_opt=0.0, xi=0.01, return_grad=False):
    """Probability of improvement (to be MAXIMISED, like the reference's gaussian_pi), acquisition.py:149-229."""
    out = _evaluate(X, model, "PI", y_opt, xi, 1.96, return_grad)
    if return_grad:
        return -out[0], -out[1]
    return -out


def gaussian_ei(X, model, y_opt=0.0, xi=0.01, return_grad=False):
    """Expected improvement (to be MAXIMISED, like the reference's gaussian_ei), acquisition.py:232-321."""
    out = _evaluate(X, model, "EI", y_opt, xi, 1.96, return_grad)
    if return_grad:
        return -out[0], -out[1]
    return -out


def _gaussian_acquisition(X, model, y_opt=None, acq_func="LCB", return_grad=False, acq_func_kwargs=None):
    """Dispatch + sign flip so that everything is minimised (acquisition.py:20-87)."""
    X = _check_X(X)
    if acq_func_kwargs is None:
        acq_func_kwargs = dict()
    xi = acq_func_kwargs.get("xi", 0.01)
    kappa = acq_func_kwargs.get("kappa", 1.96)
    if acq_func in ("EIps", "PIps"):
        return _per_second(X, model, acq_func[:2], y_opt, xi, kappa, return_grad)
    if acq_func not in ("LCB", "EI", "PI"):
        raise ValueError("Acquisition function not implemented.")
    return _evaluate(X, model, acq_func, y_opt, xi, kappa, return_grad)


def _time_posterior(X, time_model, return_grad):
    """(mu, std[, mu_grad, std_grad]) of the log-time model, batched."""
    if _is_device_model(time_model):
        dev = time_model.device_state()
        if return_grad:
            g = dev.score_grad(X, acq_funcs=())
            return g["mean"], g["std"], g["mean_grad"], g["std_grad"]
        r = dev.score(X, acq_funcs=(), want_posterior=True, precision=engine.default_precision(time_model))
        return r["mean"], r["std"], None, None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if return_grad:
            mu, std, gm, gs = time_model.predict(X, return_std=True, return_mean_grad=True, return_std_grad=True)
            m, d = X.shape
            return (np.asarray(mu, dtype=np.float64).ravel(), np.asarray(std, dtype=np.float64).ravel(),
                    np.asarray(gm, dtype=np.float64).reshape(m, d), np.asarray(gs, dtype=np.float64).reshape(m, d))
        mu, std = time_model.predict(X, return_std=True)
        return np.asarray(mu, dtype=np.float64).ravel(), np.asarray(std, dtype=np.float64).ravel(), None, None


def _per_second(X, model, base_acq, y_opt, xi, kappa, return_grad):
    """EIps / PIps (acquisition.py:38-40, 61-80): `model.estimators_` = (objective model, log-time model); the
    acquisition of the first is divided by the expected evaluation time exp(mu_t - std_t^2/2) of the second.  Both
    posteriors and the composition (value and chain rule) are evaluated on the GPU."""
    obj_model, time_model = model.estimators_
    out = _evaluate(X, obj_model, base_acq, y_opt, xi, kappa, return_grad)
    m, d = X.shape
    if return_grad:
        vals, grad = out
        grad = np.ascontiguousarray(np.asarray(grad, dtype=np.float64).reshape(m, d))
    else:
        vals, grad = out, None
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    mu_t, sd_t, gm_t, gs_t = _time_posterior(np.ascontiguousarray(X, dtype=np.float64), time_model, return_grad)
    mu_t, sd_t = np.ascontiguousarray(mu_t), np.ascontiguousarray(sd_t)
    res = np.empty(m)
    res_grad = np.empty((m, d)) if return_grad else None
    if return_grad:
        gm_t, gs_t = np.ascontiguousarray(gm_t), np.ascontiguousarray(gs_t)
    _lib.check(_lib.lib().skb_ps_combine_host(
        _default_device(), m, d, vals.ctypes.data, grad.ctypes.data if return_grad else None,
        mu_t.ctypes.data, sd_t.ctypes.data, gm_t.ctypes.data if return_grad else None,
        gs_t.ctypes.data if return_grad else None, res.ctypes.data, res_grad.ctypes.data if return_grad else None))
    if return_grad:
        return res, (res_grad[0] if m == 1 else res_grad)
    return res


def device_topk(values, k, index_offset=0):
    """(values, indices, first_nan) of the k smallest entries by (value, index) -- np.argmin / np.argsort[:k] on
    the GPU for a value vector that is already on the host (per-second acquisitions)."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    tv, ti, nan = np.empty(k), np.empty(k, dtype=np.int64), np.full(1, -1, dtype=np.int64)
    _lib.check(_lib.lib().skb_topk_host(_default_device(), values.ctypes.data, values.shape[0], int(index_offset), int(k),
                                        tv.ctypes.data, ti.ctypes.data, nan.ctypes.data))
    return tv, ti, int(nan[0])


def gaussian_acquisition_1D(X, model, y_opt=None, acq_func="LCB", acq_func_kwargs=None, return_grad=True):
    """The L-BFGS-B callback: one point in, (value array of shape (1,), gradient of shape (d,)) out
    (acquisition.py:7-17)."""
    return _gaussian_acquisition(np.expand_dims(X, axis=0), model, y_opt, acq_func=acq_func,
                                 acq_func_kwargs=acq_func_kwargs, return_grad=return_grad)
