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

from . import _lib
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
    res = dev.score(X, y_opt=y_opt, acq_funcs=(acq,), xi=xi, kappa=kappa, want_posterior=False)
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
        raise NotImplementedError("per-second acquisitions (EIps / PIps) are not on the GPU path yet")
    if acq_func not in ("LCB", "EI", "PI"):
        raise ValueError("Acquisition function not implemented.")
    return _evaluate(X, model, acq_func, y_opt, xi, kappa, return_grad)


def gaussian_acquisition_1D(X, model, y_opt=None, acq_func="LCB", acq_func_kwargs=None, return_grad=True):
    """The L-BFGS-B callback: one point in, (value array of shape (1,), gradient of shape (d,)) out
    (acquisition.py:7-17)."""
    return _gaussian_acquisition(np.expand_dims(X, axis=0), model, y_opt, acq_func=acq_func,
                                 acq_func_kwargs=acq_func_kwargs, return_grad=return_grad)
