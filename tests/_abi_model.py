"""Software model of the HOST entry points of the C ABI (include/skb.h), for CPU tests of everything ABOVE the ABI.

TEST INFRASTRUCTURE ONLY.  The product has no CPU path: without a B200 `skb_state_create` fails (tests/test_host_logic.py
::test_no_cpu_fallback_without_device).  To still exercise the Python host side on the build machine - the ctypes
marshalling of engine.py / acquisition.py, `GaussianProcessRegressor.predict`, the Optimizer's state machine, hedge
bookkeeping, constant-liar loop and lock-step L-BFGS driver - the `abi_model` fixture (tests/conftest.py) swaps the loaded
library for this object inside ONE test.  It takes the same arguments the shared library takes (struct pointers, raw
addresses, sizes), reads and writes the caller's buffers through those addresses, and computes with the CPU oracle
(oracle/gp_oracle.py).  Nothing under scikit-optimize_b200/ imports it.

The `*_host` entry points, state management, the small host-pointer helpers and the fit objective (skb_fit_*) are
modelled; the `*_dev` entry points (device pointers, streams) and candidate generation on the device are not - tests that
need them are `-m gpu` tests.
"""
import ctypes as C

import numpy as np

from oracle import gp_oracle as orc

OK, ERR_INVALID, ERR_UNSUPPORTED = 0, -1, -2
_ACQ = ("EI", "PI", "LCB")


def _obj(arg):
    """The ctypes object behind byref(x) / pointer(x) / x."""
    if hasattr(arg, "_obj"):
        return arg._obj
    if hasattr(arg, "contents"):
        return arg.contents
    return arg


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    if isinstance(p, C.c_void_p):
        return p.value or 0
    return C.cast(p, C.c_void_p).value or 0


def _view(p, shape, dtype=np.float64):
    """NumPy view of caller memory at address `p` (None when the pointer is NULL)."""
    addr = _addr(p)
    if addr == 0:
        return None
    count = int(np.prod(shape))
    if count == 0:
        return np.empty(shape, dtype=dtype)
    ctype = {np.float64: C.c_double, np.int64: C.c_int64}[dtype]
    return np.ctypeslib.as_array((ctype * count).from_address(addr)).reshape(shape)


def lexicographic_topk(values, k, index_offset=0):
    """(values, global indices, first_nan) as skb_score_out documents them: ascending by (value, index), NaN last."""
    values = np.asarray(values, dtype=np.float64)
    nan = np.isnan(values)
    first_nan = int(np.argmax(nan)) + index_offset if nan.any() else -1
    keyed = np.where(nan, np.inf, values)
    order = np.lexsort((np.arange(values.size), nan, keyed))[:k]
    tv, ti = np.full(k, np.nan), np.full(k, -1, dtype=np.int64)
    tv[:order.size] = values[order]
    ti[:order.size] = order + index_offset
    return tv, ti, first_nan


class _State:
    def __init__(self, desc):
        n, d = desc.n_train, desc.n_dims
        L_inv = np.tril(_view(desc.L_inv, (n, n)).copy())
        kind = {0: (orc.KIND_RBF, 2.5), 1: (orc.KIND_MATERN, 0.5), 2: (orc.KIND_MATERN, 1.5), 3: (orc.KIND_MATERN, 2.5),
                4: (orc.KIND_HAMMING, 2.5)}[desc.kernel]
        K_inv = _view(desc.K_inv, (n, n))
        self.has_kinv = K_inv is not None
        self.is_hamming = desc.kernel == 4
        self.gp = orc.GPState(X_train=_view(desc.X_train, (n, d)).copy(), alpha=_view(desc.alpha, (n,)).copy(),
                              K_inv=K_inv.copy() if self.has_kinv else L_inv.T @ L_inv,
                              length_scale=_view(desc.length_scale, (d,)).copy(), kind=kind[0], nu=kind[1],
                              amplitude=desc.amplitude, bias=desc.bias, white=desc.white, y_mean=desc.y_mean,
                              y_std=desc.y_std)


def _acquisitions(prm, mu, std, mu_grad=None, std_grad=None):
    """{acq index: values or (values, grad)} to MINIMISE, for the acquisitions prm.acq_mask requests."""
    kappa = "inf" if prm.kappa_inf else prm.kappa
    out = {}
    for a, name in enumerate(_ACQ):
        if not prm.acq_mask & (1 << a):
            continue
        if name == "LCB":
            out[a] = orc.lcb_from_posterior(mu, std, kappa, mu_grad, std_grad)
            continue
        fn = orc.ei_from_posterior if name == "EI" else orc.pi_from_posterior
        if mu_grad is None:
            out[a] = -fn(mu, std, prm.y_opt, prm.xi)
        else:                                     # the reference's gradient branch is per point
            vals, grads = np.empty(mu.shape[0]), np.empty(mu_grad.shape)
            for i in range(mu.shape[0]):
                v, g = fn(mu[i:i + 1], std[i:i + 1], prm.y_opt, prm.xi, mu_grad[i], std_grad[i])
                vals[i], grads[i] = -v[0], -g
            out[a] = (vals, grads)
    return out


class AbiModel:
    """Drop-in for the object `_lib.lib()` returns, restricted to the host entry points."""
    on_device = False

    def __init__(self):
        self._states = {}
        self._fits = {}
        self._next = 1
        self._error = b""
        self.calls = []                 # names of the entry points called, in order (tests assert on it)

    def _fail(self, code, message):
        self._error = message.encode()
        return code

    # ---- bookkeeping ------------------------------------------------------------------------------------------------
    def skb_abi_version(self):
        return 1

    def skb_last_error(self):
        return self._error

    def skb_device_count(self):
        return 1

    def skb_launch_count(self):
        return len(self.calls)

    def skb_state_create(self, desc, device, out):
        desc = _obj(desc)
        self.calls.append("skb_state_create")
        if desc.n_train <= 0 or desc.n_dims <= 0 or not desc.X_train or not desc.alpha or not desc.L_inv:
            return self._fail(ERR_INVALID, "bad state descriptor")
        if desc.n_dims > 100:
            return self._fail(ERR_UNSUPPORTED, "n_dims %d > 100" % desc.n_dims)
        handle = self._next
        self._next += 1
        self._states[handle] = _State(desc)
        _obj(out).value = handle
        return OK

    def skb_state_destroy(self, handle):
        self._states.pop(_addr(handle), None)

    def skb_state_attach_kinv(self, handle, K_inv):
        st = self._states[_addr(handle)]
        self.calls.append("skb_state_attach_kinv")
        n = st.gp.n
        st.gp.K_inv = _view(K_inv, (n, n)).copy()
        st.has_kinv = True
        return OK

    def skb_state_split_info(self, handle, planes, diff):
        _obj(planes).value, _obj(diff).value = 0, -1.0
        return OK

    def skb_state_split_bounds(self, handle, bound_var, bound_mean):
        _obj(bound_var).value, _obj(bound_mean).value = -1.0, -1.0
        return OK

    def skb_state_decide_split(self, handle):
        return OK

    def skb_state_stream(self, handle):
        return None

    def skb_state_sync(self, handle):
        return OK

    def skb_state_set_scratch_limit(self, handle, nbytes):
        return OK

    def skb_state_set_timing(self, handle, enabled):
        return OK

    # ---- scoring ----------------------------------------------------------------------------------------------------
    def skb_predict_mean_host(self, handle, X, m, mean):
        st = self._states[_addr(handle)]
        self.calls.append("skb_predict_mean_host")
        _view(mean, (m,))[:] = orc.predict(st.gp, _view(X, (m, st.gp.d)))
        return OK

    def skb_predict_cov_host(self, handle, X, m, mean, quad):
        st = self._states[_addr(handle)]
        self.calls.append("skb_predict_cov_host")
        Xv = _view(X, (m, st.gp.d))
        K_trans = orc.kernel_cross(st.gp, Xv)
        if mean:
            _view(mean, (m,))[:] = orc.predict(st.gp, Xv)
        _view(quad, (m, m))[:] = K_trans.dot(st.gp.K_inv).dot(K_trans.T)
        return OK

    def skb_score_host(self, handle, X, m, index_offset, prm, out):
        st = self._states[_addr(handle)]
        prm, out = _obj(prm), _obj(out)
        self.calls.append("skb_score_host")
        if m < 0 or prm.top_k < 0:
            return self._fail(ERR_INVALID, "negative size")
        mu, std = orc.predict(st.gp, _view(X, (m, st.gp.d)), return_std=True) if m else (np.empty(0), np.empty(0))
        for ptr, src in ((out.mean, mu), (out.std, std)):
            if ptr:
                _view(ptr, (m,))[:] = src
        if out.n_var_clamped:
            _view(out.n_var_clamped, (1,), np.int64)[0] = 0
        for a, vals in _acquisitions(prm, mu, std).items():
            if out.acq[a]:
                _view(out.acq[a], (m,))[:] = vals
            if prm.top_k > 0:
                tv, ti, first_nan = lexicographic_topk(vals, prm.top_k, index_offset)
                if out.topk_val[a]:
                    _view(out.topk_val[a], (prm.top_k,))[:] = tv
                if out.topk_idx[a]:
                    _view(out.topk_idx[a], (prm.top_k,), np.int64)[:] = ti
                if out.first_nan[a]:
                    _view(out.first_nan[a], (1,), np.int64)[0] = first_nan
        return OK

    def skb_score_grad_host(self, handle, X, m, prm, out):
        st = self._states[_addr(handle)]
        prm, out = _obj(prm), _obj(out)
        self.calls.append("skb_score_grad_host")
        if st.is_hamming:
            return self._fail(ERR_UNSUPPORTED, "the Hamming kernel has no input gradient")
        if not st.has_kinv:
            return self._fail(ERR_UNSUPPORTED, "state was created without K_inv: gradient entry points are unavailable")
        d = st.gp.d
        mu, std, gm, gs = orc.predict_with_grads_batched(st.gp, _view(X, (m, d)))
        for ptr, src, shape in ((out.mean, mu, (m,)), (out.std, std, (m,)), (out.mean_grad, gm, (m, d)),
                                (out.std_grad, gs, (m, d))):
            if ptr:
                _view(ptr, shape)[:] = src
        for a, (vals, grads) in _acquisitions(prm, mu, std, gm, gs).items():
            if out.acq[a]:
                _view(out.acq[a], (m,))[:] = vals
            if out.acq_grad[a]:
                _view(out.acq_grad[a], (m, d))[:] = grads
        return OK

    def skb_acq_from_posterior_host(self, device, m, n_dims, mu, std, mu_grad, std_grad, prm, acq, acq_grad):
        prm = _obj(prm)
        self.calls.append("skb_acq_from_posterior_host")
        mu, std = _view(mu, (m,)), _view(std, (m,))
        gm, gs = _view(mu_grad, (m, n_dims)), _view(std_grad, (m, n_dims))
        for a, res in _acquisitions(prm, mu, std, gm, gs if gm is not None else None).items():
            vals, grads = res if gm is not None else (res, None)
            if acq[a]:
                _view(acq[a], (m,))[:] = vals
            if grads is not None and acq_grad[a]:
                _view(acq_grad[a], (m, n_dims))[:] = grads
        return OK

    def skb_ps_combine_host(self, device, m, n_dims, acq, acq_grad, mu_t, std_t, mu_t_grad, std_t_grad, out, out_grad):
        """acquisition.py:61-80: value / expected time, and its gradient by the product rule."""
        self.calls.append("skb_ps_combine_host")
        acq, mu_t, std_t = _view(acq, (m,)), _view(mu_t, (m,)), _view(std_t, (m,))
        inv_t = np.exp(-mu_t + 0.5 * std_t ** 2)
        _view(out, (m,))[:] = acq * inv_t
        if _addr(out_grad):
            g = _view(acq_grad, (m, n_dims))
            gm, gs = _view(mu_t_grad, (m, n_dims)), _view(std_t_grad, (m, n_dims))
            inv_t_grad = inv_t[:, None] * (std_t[:, None] * gs - gm)
            _view(out_grad, (m, n_dims))[:] = g * inv_t[:, None] + acq[:, None] * inv_t_grad
        return OK

    def skb_topk_host(self, device, values, m, index_offset, top_k, topk_val, topk_idx, first_nan):
        self.calls.append("skb_topk_host")
        tv, ti, nan = lexicographic_topk(_view(values, (m,)), top_k, index_offset)
        _view(topk_val, (top_k,))[:] = tv
        _view(topk_idx, (top_k,), np.int64)[:] = ti
        if _addr(first_nan):
            _view(first_nan, (1,), np.int64)[0] = nan
        return OK

    # ---- fit objective (opt-in GPU fit): log marginal likelihood, finalisation, state from the resident factor ------------
    def skb_fit_create(self, desc, device, out):
        desc = _obj(desc)
        self.calls.append("skb_fit_create")
        if desc.n_train <= 0 or desc.n_dims <= 0 or not desc.X_train or not desc.y_train:
            return self._fail(ERR_INVALID, "bad fit descriptor")
        if desc.kernel == 4:
            return self._fail(ERR_UNSUPPORTED, "the Hamming kernel has no device fit objective")
        handle = self._next
        self._next += 1
        n, d = desc.n_train, desc.n_dims
        self._fits[handle] = {"X": _view(desc.X_train, (n, d)).copy(), "y": _view(desc.y_train, (n,)).copy(),
                              "nu": {0: np.inf, 1: 0.5, 2: 1.5, 3: 2.5}[desc.kernel], "alpha": desc.alpha,
                              "kernel": desc.kernel, "factor": None}
        _obj(out).value = handle
        return OK

    def skb_fit_destroy(self, handle):
        self._fits.pop(_addr(handle), None)

    def _lml(self, fit, theta, want_grad):
        with np.errstate(all="ignore"):
            value, grad = orc.log_marginal_likelihood(fit["X"], fit["y"], fit["nu"], theta, alpha=fit["alpha"])
        if not np.isfinite(theta[-1]):             # log noise = -inf: no noise term, its gradient entry is 0
            grad = np.where(np.isfinite(grad), grad, 0.0)
        return value, (grad if want_grad else np.zeros_like(grad))

    def skb_fit_lml_host(self, handle, theta, want_grad, lml, grad):
        fit = self._fits[_addr(handle)]
        self.calls.append("skb_fit_lml_host")
        d = fit["X"].shape[1]
        value, g = self._lml(fit, _view(theta, (d + 2,)).copy(), want_grad)
        _obj(lml).value = value
        if _addr(grad):
            _view(grad, (d + 2,))[:] = g
        return OK

    def skb_fit_lml_batch_host(self, handle, count, thetas, want_grad, lml, grad):
        fit = self._fits[_addr(handle)]
        self.calls.append("skb_fit_lml_batch_host")
        d = fit["X"].shape[1]
        T = _view(thetas, (count, d + 2))
        for i in range(count):
            value, g = self._lml(fit, T[i].copy(), want_grad)
            _view(lml, (count,))[i] = value
            if _addr(grad):
                _view(grad, (count, d + 2))[i] = g
        return OK

    def skb_fit_finalize_host(self, handle, theta, lml, alpha_out, L_out):
        """Tail of sklearn's fit (_gpr.py:341-367) at theta: K = kernel(X) + alpha I, L = chol(K), alpha_ = K^-1 y."""
        from scipy.linalg import cho_solve, cholesky
        fit = self._fits[_addr(handle)]
        self.calls.append("skb_fit_finalize_host")
        X, y = fit["X"], fit["y"]
        n, d = X.shape
        theta = _view(theta, (d + 2,)).copy()
        amplitude, ls, noise = np.exp(theta[0]), np.exp(theta[1:d + 1]), np.exp(theta[d + 1])
        state = orc.GPState(X_train=X, alpha=np.zeros(n), K_inv=np.eye(n), length_scale=ls,
                            kind=orc.KIND_RBF if fit["kernel"] == 0 else orc.KIND_MATERN, nu=fit["nu"], amplitude=amplitude)
        K = orc.kernel_cross(state, X) + (noise + fit["alpha"]) * np.eye(n)
        try:
            L = cholesky(K, lower=True, check_finite=False)
        except np.linalg.LinAlgError:
            return self._fail(-6, "K(theta) is not positive definite")
        a = cho_solve((L, True), y, check_finite=False)
        _obj(lml).value = -0.5 * y.dot(a) - np.log(np.diag(L)).sum() - n / 2.0 * np.log(2.0 * np.pi)
        _view(alpha_out, (n,))[:] = a
        if _addr(L_out):
            _view(L_out, (n, n))[:] = L
        fit["factor"], fit["alpha_vec"] = L, a
        return OK

    def skb_state_create_from_fit(self, handle, desc, out):
        fit = self._fits[_addr(handle)]
        desc = _obj(desc)
        self.calls.append("skb_state_create_from_fit")
        if fit["factor"] is None:
            return self._fail(ERR_INVALID, "skb_fit_finalize_host has not succeeded on this handle")
        n, d = fit["X"].shape
        L_inv = np.linalg.solve(fit["factor"], np.eye(n))
        st = _State.__new__(_State)
        st.has_kinv, st.is_hamming = True, False
        kind = {0: (orc.KIND_RBF, 2.5), 1: (orc.KIND_MATERN, 0.5), 2: (orc.KIND_MATERN, 1.5), 3: (orc.KIND_MATERN, 2.5)}
        st.gp = orc.GPState(X_train=fit["X"].copy(), alpha=fit["alpha_vec"].copy(), K_inv=L_inv.T @ L_inv,
                            length_scale=_view(desc.length_scale, (d,)).copy(), kind=kind[desc.kernel][0],
                            nu=kind[desc.kernel][1], amplitude=desc.amplitude, bias=desc.bias, white=desc.white,
                            y_mean=desc.y_mean, y_std=desc.y_std)
        new = self._next
        self._next += 1
        self._states[new] = st
        _obj(out).value = new
        return OK
