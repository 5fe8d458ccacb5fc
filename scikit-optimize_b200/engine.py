"""Host side of the device boundary: flatten a fitted GP into the POD the C ABI takes, and drive the scoring calls.

Reference interfaces mirrored (paths relative to the reference tree):
  * attribute reads of GaussianProcessRegressor.predict       skopt/learning/gaussian_process/gpr.py:315-356
  * kernel algebra of sklearn Sum / Product / Constant / White sklearn kernels.py:871,890,971,990,1278,1315,1418,1438
  * _gaussian_acquisition + argmin / argsort[:k]               skopt/acquisition.py:20-87, optimizer.py:557-570
"""
import ctypes as C
import math
import os
import weakref

import numpy as np
from scipy.linalg import solve_triangular

from . import _lib

ACQ_NAMES = ("EI", "PI", "LCB")


# ----------------------------------------------------------------------------------------------
# kernel tree -> amplitude * base(r) + bias (+ white on the diagonal)
# ----------------------------------------------------------------------------------------------
class _Affine:
    """k(x, y) = a * base(x, y) + c  off the diagonal;  diag(x) = a + c + w."""
    __slots__ = ("a", "c", "w", "base")

    def __init__(self, a=0.0, c=0.0, w=0.0, base=None):
        self.a, self.c, self.w, self.base = a, c, w, base


def _flatten_kernel(kernel):
    cls = type(kernel).__name__
    if cls == "ConstantKernel":
        return _Affine(c=float(kernel.constant_value))
    if cls == "WhiteKernel":
        return _Affine(w=float(kernel.noise_level))
    if cls == "RBF":
        return _Affine(a=1.0, base=(_lib.KERNEL_RBF, np.asarray(kernel.length_scale, dtype=np.float64)))
    if cls == "Matern":
        kinds = {0.5: _lib.KERNEL_MATERN12, 1.5: _lib.KERNEL_MATERN32, 2.5: _lib.KERNEL_MATERN52}
        if kernel.nu == math.inf:
            kind = _lib.KERNEL_RBF          # sklearn kernels.py:1730-1731
        elif kernel.nu in kinds:
            kind = kinds[kernel.nu]
        else:
            raise NotImplementedError("Matern(nu=%r): only nu in {0.5, 1.5, 2.5, inf} run on the GPU path" % kernel.nu)
        return _Affine(a=1.0, base=(kind, np.asarray(kernel.length_scale, dtype=np.float64)))
    if cls == "HammingKernel":          # length_scale = per-dimension mismatch weights, coordinates stay unscaled
        return _Affine(a=1.0, base=(_lib.KERNEL_HAMMING, np.asarray(kernel.length_scale, dtype=np.float64)))
    if cls in ("Sum", "Product"):
        l, r = _flatten_kernel(kernel.k1), _flatten_kernel(kernel.k2)
        if l.base is not None and r.base is not None:
            raise NotImplementedError("%s of two stationary kernels is outside the GPU hot-path grammar "
                                      "([Constant *] {RBF|Matern|Hamming} [+ Constant] [+ White])" % cls)
        base = l.base if l.base is not None else r.base
        if cls == "Sum":
            return _Affine(l.a + r.a, l.c + r.c, l.w + r.w, base)
        off_l, off_r = l.a + l.c, r.a + r.c                 # value at r = 0 without the white term
        return _Affine(l.a * r.c + l.c * r.a, l.c * r.c, (off_l + l.w) * (off_r + r.w) - off_l * off_r, base)
    raise NotImplementedError("kernel %s is outside the GPU hot-path grammar "
                              "([Constant *] {RBF|Matern|Hamming} [+ Constant] [+ White])" % cls)


class HostState:
    """Plain arrays describing a fitted GP (the argument of skb_state_create)."""

    def __init__(self, X_train, alpha, L_inv, length_scale, kernel, amplitude, bias, white, y_mean, y_std,
                 K_inv=None):
        self.X_train = np.ascontiguousarray(X_train, dtype=np.float64)
        self.alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        self.L_inv = np.ascontiguousarray(L_inv, dtype=np.float64)
        self.K_inv = None if K_inv is None else np.ascontiguousarray(K_inv, dtype=np.float64)
        if self.K_inv is not None and self.K_inv.shape != self.L_inv.shape:
            raise ValueError("K_inv must have shape (n_train, n_train)")
        n, d = self.X_train.shape
        if self.alpha.shape != (n,):
            raise ValueError("alpha must have shape (n_train,); multi-output targets are not supported on the GPU path")
        if self.L_inv.shape != (n, n):
            raise ValueError("L_inv must have shape (n_train, n_train)")
        self.length_scale = np.ascontiguousarray(np.broadcast_to(np.asarray(length_scale, dtype=np.float64), (d,)))
        self.kernel = int(kernel)
        self.amplitude, self.bias, self.white = float(amplitude), float(bias), float(white)
        self.y_mean, self.y_std = float(y_mean), float(y_std)

    @property
    def n_train(self):
        return self.X_train.shape[0]

    @property
    def n_dims(self):
        return self.X_train.shape[1]

    @property
    def prior_var(self):
        return self.amplitude + self.bias + self.white


def host_state_from_estimator(est):
    """Flatten a fitted skopt-style GaussianProcessRegressor (outputs of gpr.py:166-237)."""
    if not hasattr(est, "X_train_"):
        raise ValueError("estimator is not fitted")
    aff = _flatten_kernel(est.kernel_)
    d = est.X_train_.shape[1]
    if aff.base is None:
        kind, ls, amp = _lib.KERNEL_RBF, np.ones(d), 0.0
    else:
        (kind, ls), amp = aff.base, aff.a
    L_inv = getattr(est, "_L_inv_lower", None)
    if L_inv is None:
        L_inv = solve_triangular(est.L_, np.eye(est.L_.shape[0]), lower=True)
    return HostState(est.X_train_, est.alpha_, L_inv, ls, kind, amp, aff.c, aff.w,
                     np.ravel(est.y_train_mean_)[0], np.ravel(est.y_train_std_)[0], K_inv=getattr(est, "K_inv_", None))


# ----------------------------------------------------------------------------------------------
# device state
# ----------------------------------------------------------------------------------------------
def _mask(acq_funcs):
    m = 0
    for a in acq_funcs:
        if a not in _lib.ACQ_INDEX:
            raise ValueError("Acquisition function not implemented.")
        m |= 1 << _lib.ACQ_INDEX[a]
    return m


# "auto": the split-integer tensor-core contraction pays off once the O(n^2)-per-candidate term dominates; below these
# sizes the FP64 DMMA path is used (it is also the path whose small-n trajectories are pinned bit-for-bit to the reference)
AUTO_SPLIT_MIN_TRAIN = 512
AUTO_SPLIT_MAX_TRAIN = 16384          # exact int32 accumulation bound of the split-integer path
AUTO_SPLIT_MIN_CANDIDATES = 8192


_DEFAULT_PRECISION = ["auto"]


def set_default_precision(precision):
    """Precision of the calls that have no argument for it - predict(return_std=True), gaussian_ei / pi / lcb,
    _gaussian_acquisition and an Optimizer built without acq_optimizer_kwargs["precision"]: "auto" (default; the
    split-integer contraction once n_train >= 512 and the call has >= 8192 candidates, FP64 DMMA below), "fp64" (every
    call on the FP64 path, bit-pinned against the reference fixtures at every size), "split_int8", "split_int8_p7".
    An estimator attribute `precision` overrides it for that estimator.  Returns the previous setting."""
    if precision != "auto" and precision not in _lib.PRECISION:
        raise ValueError("precision must be 'auto' or one of %s" % sorted(_lib.PRECISION))
    old, _DEFAULT_PRECISION[0] = _DEFAULT_PRECISION[0], precision
    return old


def default_precision(estimator=None):
    own = getattr(estimator, "precision", None) if estimator is not None else None
    return own if own is not None else _DEFAULT_PRECISION[0]


def resolve_precision(precision, n_train, m):
    if precision is None:
        precision = _DEFAULT_PRECISION[0]
    if precision == "auto":
        return "split_int8" if (AUTO_SPLIT_MIN_TRAIN <= n_train <= AUTO_SPLIT_MAX_TRAIN and
                                m >= AUTO_SPLIT_MIN_CANDIDATES) else "fp64"
    if precision not in _lib.PRECISION:
        raise ValueError("precision must be 'auto' or one of %s" % sorted(_lib.PRECISION))
    return precision


def _acq_params(y_opt, xi, kappa, acq_funcs, top_k, precision="fp64"):
    p = _lib.AcqParams()
    if precision not in _lib.PRECISION:
        raise ValueError("precision must be one of %s" % sorted(_lib.PRECISION))
    p.precision = _lib.PRECISION[precision]
    p.y_opt = 0.0 if y_opt is None else float(y_opt)
    p.xi = float(xi)
    p.kappa_inf = 1 if isinstance(kappa, str) and kappa == "inf" else 0
    p.kappa = 0.0 if p.kappa_inf else float(kappa)
    p.acq_mask = _mask(acq_funcs)
    p.top_k = int(top_k)
    return p


class StateScalars:
    """The part of a HostState that is not an n-sized array: what skb_state_create_from_fit needs from the host when the
    training rows, alpha and the Cholesky factor are already on the device."""

    def __init__(self, n_train, n_dims, length_scale, kernel, amplitude, bias, white, y_mean, y_std):
        self.n_train, self.n_dims = int(n_train), int(n_dims)
        self.length_scale = np.ascontiguousarray(np.broadcast_to(np.asarray(length_scale, dtype=np.float64), (self.n_dims,)))
        self.kernel = int(kernel)
        self.amplitude, self.bias, self.white = float(amplitude), float(bias), float(white)
        self.y_mean, self.y_std = float(y_mean), float(y_std)
        self.K_inv = None

    @property
    def prior_var(self):
        return self.amplitude + self.bias + self.white


def state_scalars_from_estimator(est):
    """Kernel scalars of the PREDICTION kernel est.kernel_ (after gpr.py:199-220 silenced the noise term)."""
    aff = _flatten_kernel(est.kernel_)
    n, d = est.X_train_.shape
    if aff.base is None:
        kind, ls, amp = _lib.KERNEL_RBF, np.ones(d), 0.0
    else:
        (kind, ls), amp = aff.base, aff.a
    return StateScalars(n, d, ls, kind, amp, aff.c, aff.w, np.ravel(est.y_train_mean_)[0], np.ravel(est.y_train_std_)[0])


def _fill_desc(desc, hs):
    desc.n_train, desc.n_dims, desc.kernel = hs.n_train, hs.n_dims, hs.kernel
    desc.amplitude, desc.bias, desc.white = hs.amplitude, hs.bias, hs.white
    desc.y_mean, desc.y_std = hs.y_mean, hs.y_std
    desc.length_scale = hs.length_scale.ctypes.data


class DeviceState:
    """A fitted GP resident on one B200 (owner of a skb_state handle)."""

    def __init__(self, host_state, device=0, _handle=None):
        L = _lib.lib()
        self.host = host_state
        self.device = int(device)
        if _handle is None:
            desc = _lib.StateDesc()
            _fill_desc(desc, host_state)
            desc.X_train = host_state.X_train.ctypes.data
            desc.alpha = host_state.alpha.ctypes.data
            desc.L_inv = host_state.L_inv.ctypes.data
            desc.K_inv = None                   # attached lazily: only the gradient entry points read it
            self._kinv_attached = False
            handle = C.c_void_p()
            _lib.check(L.skb_state_create(C.byref(desc), self.device, C.byref(handle)))
        else:
            handle = _handle                    # built on the device by skb_state_create_from_fit, K_inv included
            self._kinv_attached = True
        self._h = handle
        self._finalizer = weakref.finalize(self, L.skb_state_destroy, handle)

    @classmethod
    def from_estimator(cls, est, device=0):
        return cls(host_state_from_estimator(est), device)

    def close(self):
        self._finalizer()

    @property
    def n_dims(self):
        return self.host.n_dims

    @property
    def stream(self):
        return _lib.lib().skb_state_stream(self._h)

    def sync(self):
        _lib.check(_lib.lib().skb_state_sync(self._h))

    def set_scratch_limit(self, nbytes):
        _lib.check(_lib.lib().skb_state_set_scratch_limit(self._h, int(nbytes)))

    def _attach_kinv(self):
        """Upload K_inv (and pack its FP64 tiles / digit planes) the first time a gradient is requested."""
        if not self._kinv_attached:
            if self.host.K_inv is None:
                raise ValueError("this state was built without K_inv: gradients are unavailable")
            _lib.check(_lib.lib().skb_state_attach_kinv(self._h, self.host.K_inv.ctypes.data))
            self._kinv_attached = True

    def split_info(self):
        """(planes, probe_diff): digit planes precision="split_int8" contracts on this state - 6 when the self-check (run on
        the first split-integer call with >= 32768 candidates, or by decide_split()) agreed with the FP64 path to 1e-10 on
        variance and mean, else 7; 0 while undecided (such calls contract 7 planes) - and the difference it measured."""
        planes, diff = C.c_int32(0), C.c_double(-1.0)
        _lib.check(_lib.lib().skb_state_split_info(self._h, C.byref(planes), C.byref(diff)))
        return int(planes.value), float(diff.value)

    def split_bounds(self):
        """(bound_var, bound_mean): the a-priori error bounds of the 6-plane contraction for this state (-1 before the
        decision ran); 6 planes need both <= 2.5e-10 and the measured probe difference <= 1e-10."""
        bv, bm = C.c_double(-1.0), C.c_double(-1.0)
        _lib.check(_lib.lib().skb_state_split_bounds(self._h, C.byref(bv), C.byref(bm)))
        return float(bv.value), float(bm.value)

    def decide_split(self):
        """Run the 6-or-7-plane self-check now (it otherwise runs on the first split-integer call with >= 32768
        candidates); returns split_info()."""
        _lib.check(_lib.lib().skb_state_decide_split(self._h))
        return self.split_info()

    def set_timing(self, enabled):
        _lib.check(_lib.lib().skb_state_set_timing(self._h, 1 if enabled else 0))

    def get_timing(self):
        """{kind: (milliseconds, spans)} accumulated since the last call; kinds: k1, k2, topk, grad."""
        ms = (C.c_double * 4)()
        cnt = (C.c_int64 * 4)()
        _lib.check(_lib.lib().skb_state_get_timing(self._h, ms, cnt, 4))
        return {name: (ms[i], int(cnt[i])) for i, name in enumerate(("k1", "k2", "topk", "grad"))}

    # ---------------------------------------------------------------- host (NumPy) entry points
    def _check_X(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("X is {}-dimensional, however, it must be 2-dimensional.".format(X.ndim))
        if X.shape[1] != self.n_dims:
            raise ValueError("X has %d features, the model was fitted with %d" % (X.shape[1], self.n_dims))
        return X

    def predict_mean(self, X):
        X = self._check_X(X)
        mean = np.empty(X.shape[0])
        _lib.check(_lib.lib().skb_predict_mean_host(self._h, X.ctypes.data, X.shape[0], mean.ctypes.data))
        return mean

    def predict_cov(self, X):
        """(mean [m], quad [m, m]) with quad = K_trans K_inv K_trans^T in kernel units: the device half of
        predict(return_cov=True) (gpr.py:315-325).  Raises SkbError (unsupported) when m exceeds the state's scratch."""
        X = self._check_X(X)
        m = X.shape[0]
        self._attach_kinv()
        mean, quad = np.empty(m), np.empty((m, m))
        _lib.check(_lib.lib().skb_predict_cov_host(self._h, X.ctypes.data, m, mean.ctypes.data, quad.ctypes.data))
        return mean, quad

    def score(self, X, y_opt=None, acq_funcs=("EI",), xi=0.01, kappa=1.96, top_k=0, index_offset=0,
              want_posterior=True, want_values=True, precision="fp64"):
        """One pass over the candidates: posterior, every requested acquisition, and their top-k.

        Returns a dict with 'mean', 'std' (if want_posterior), per acquisition name 'values' (if want_values),
        'topk_val', 'topk_idx', 'first_nan' (if top_k > 0).
        """
        X = self._check_X(X)
        m = X.shape[0]
        precision = resolve_precision(precision, self.host.n_train, m)
        prm = _acq_params(y_opt, xi, kappa, acq_funcs, top_k, precision)
        out = _lib.ScoreOut()
        res = {"precision": precision}
        clamped = np.zeros(1, dtype=np.int64)
        out.n_var_clamped = clamped.ctypes.data
        if want_posterior:
            res["mean"], res["std"] = np.empty(m), np.empty(m)
            out.mean, out.std = res["mean"].ctypes.data, res["std"].ctypes.data
        for name in acq_funcs:
            a = _lib.ACQ_INDEX[name]
            r = res.setdefault(name, {})
            if want_values:
                r["values"] = np.empty(m)
                out.acq[a] = r["values"].ctypes.data
            if top_k > 0:
                r["topk_val"] = np.full(top_k, np.nan)
                r["topk_idx"] = np.full(top_k, -1, dtype=np.int64)
                r["_nan"] = np.full(1, -1, dtype=np.int64)
                out.topk_val[a] = r["topk_val"].ctypes.data
                out.topk_idx[a] = r["topk_idx"].ctypes.data
                out.first_nan[a] = r["_nan"].ctypes.data
        _lib.check(_lib.lib().skb_score_host(self._h, X.ctypes.data, m, int(index_offset), C.byref(prm), C.byref(out)))
        for name in acq_funcs:
            r = res[name]
            if "_nan" in r:
                r["first_nan"] = int(r.pop("_nan")[0])
        res["clamped"] = int(clamped[0])
        return res

    def score_grad(self, X, y_opt=None, acq_funcs=("EI",), xi=0.01, kappa=1.96, precision="fp64"):
        """Posterior, acquisition values and their input gradients for every row of X (batched extension of
        gaussian_acquisition_1D / predict(..., return_mean_grad, return_std_grad); the reference takes one row)."""
        X = self._check_X(X)
        m, d = X.shape
        self._attach_kinv()
        prm = _acq_params(y_opt, xi, kappa, acq_funcs, 0, resolve_precision(precision, self.host.n_train, m))
        out = _lib.GradOut()
        res = {"mean": np.empty(m), "std": np.empty(m), "mean_grad": np.empty((m, d)), "std_grad": np.empty((m, d))}
        out.mean, out.std = res["mean"].ctypes.data, res["std"].ctypes.data
        out.mean_grad, out.std_grad = res["mean_grad"].ctypes.data, res["std_grad"].ctypes.data
        for name in acq_funcs:
            a = _lib.ACQ_INDEX[name]
            res[name] = {"values": np.empty(m), "grad": np.empty((m, d))}
            out.acq[a] = res[name]["values"].ctypes.data
            out.acq_grad[a] = res[name]["grad"].ctypes.data
        _lib.check(_lib.lib().skb_score_grad_host(self._h, X.ctypes.data, m, C.byref(prm), C.byref(out)))
        return res

    # ---------------------------------------------------------------- device (torch tensor) entry points
    def score_tensors(self, X, y_opt=None, acq_funcs=("EI",), xi=0.01, kappa=1.96, top_k=0, index_offset=0,
                      want_posterior=False, want_values=False, stream=None, precision="fp64"):
        """Same as score() for a CUDA float64 tensor already resident on this state's GPU.  Asynchronous:
        work is enqueued on `stream` (a torch.cuda.Stream; default: torch's current stream)."""
        import torch
        if not (X.is_cuda and X.dtype == torch.float64 and X.dim() == 2 and X.is_contiguous()):
            raise ValueError("X must be a contiguous 2-D float64 CUDA tensor")
        if X.device.index != self.device or X.shape[1] != self.n_dims:
            raise ValueError("X is on the wrong device or has the wrong number of features")
        m = X.shape[0]
        dev = X.device
        precision = resolve_precision(precision, self.host.n_train, m)
        prm = _acq_params(y_opt, xi, kappa, acq_funcs, top_k, precision)
        out = _lib.ScoreOut()
        res = {}
        if want_posterior:
            res["mean"] = torch.empty(m, dtype=torch.float64, device=dev)
            res["std"] = torch.empty(m, dtype=torch.float64, device=dev)
            out.mean, out.std = res["mean"].data_ptr(), res["std"].data_ptr()
        for name in acq_funcs:
            a = _lib.ACQ_INDEX[name]
            r = res.setdefault(name, {})
            if want_values:
                r["values"] = torch.empty(m, dtype=torch.float64, device=dev)
                out.acq[a] = r["values"].data_ptr()
            if top_k > 0:
                r["topk_val"] = torch.empty(top_k, dtype=torch.float64, device=dev)
                r["topk_idx"] = torch.empty(top_k, dtype=torch.int64, device=dev)
                r["first_nan"] = torch.empty(1, dtype=torch.int64, device=dev)
                out.topk_val[a] = r["topk_val"].data_ptr()
                out.topk_idx[a] = r["topk_idx"].data_ptr()
                out.first_nan[a] = r["first_nan"].data_ptr()
        s = stream if stream is not None else torch.cuda.current_stream(dev)
        _lib.check(_lib.lib().skb_score_dev(self._h, X.data_ptr(), m, int(index_offset), C.byref(prm), C.byref(out),
                                            C.c_void_p(s.cuda_stream)))
        return res


    def score_record(self, X, y_opt=None, acq_funcs=("EI",), xi=0.01, kappa=1.96, top_k=1, index_offset=0, stream=None,
                     precision="fp64", values_out=None):
        """score_tensors() whose top-k outputs for ALL requested acquisitions land in one packed record (an int64 CUDA
        tensor of len(acq_funcs) * (2 top_k + 1) words, layout of skb_topk_merge_records_dev in include/skb.h): what a
        rank contributes to the single all-gather of a sharded scoring call.  `X` may have zero rows (the record is
        then padding).  values_out: optional {acq name: float64 CUDA tensor [m]} to also keep the full value vectors."""
        import torch
        if not (X.is_cuda and X.dtype == torch.float64 and X.dim() == 2 and X.is_contiguous()):
            raise ValueError("X must be a contiguous 2-D float64 CUDA tensor")
        if X.device.index != self.device or X.shape[1] != self.n_dims:
            raise ValueError("X is on the wrong device or has the wrong number of features")
        m, A, k = X.shape[0], len(acq_funcs), int(top_k)
        prm = _acq_params(y_opt, xi, kappa, acq_funcs, k, resolve_precision(precision, self.host.n_train, m))
        rec = torch.empty(record_words(A, k), dtype=torch.int64, device=X.device)
        out = _lib.ScoreOut()
        base = rec.data_ptr()
        for j, name in enumerate(acq_funcs):               # record slots follow the ORDER of acq_funcs
            a = _lib.ACQ_INDEX[name]
            out.topk_val[a] = base + 8 * (j * k)
            out.topk_idx[a] = base + 8 * (A * k + j * k)
            out.first_nan[a] = base + 8 * (2 * A * k + j)
            if values_out is not None and name in values_out:
                out.acq[a] = values_out[name].data_ptr()
        s = stream if stream is not None else torch.cuda.current_stream(X.device)
        _lib.check(_lib.lib().skb_score_dev(self._h, X.data_ptr() if m else None, m, int(index_offset), C.byref(prm),
                                            C.byref(out), C.c_void_p(s.cuda_stream)))
        return rec


    def score_grad_tensors(self, X, y_opt=None, acq_funcs=("EI",), xi=0.01, kappa=1.96, precision="fp64", top_k=0,
                           index_offset=0, stream=None, out=None):
        """score_grad() for a CUDA float64 tensor resident on this state's GPU; asynchronous on `stream`.  Returns CUDA
        tensors ('mean', 'std', 'mean_grad', 'std_grad', per acquisition 'values' / 'grad') and, with top_k > 0, 'record':
        the packed top-k record of the acquisition VALUES (slots in the order of acq_funcs) for the cross-rank arg-min.
        `out`: a dict returned by an earlier call of the same shape, whose tensors are then overwritten (no allocation)."""
        import torch
        if not (X.is_cuda and X.dtype == torch.float64 and X.dim() == 2 and X.is_contiguous()):
            raise ValueError("X must be a contiguous 2-D float64 CUDA tensor")
        if X.device.index != self.device or X.shape[1] != self.n_dims:
            raise ValueError("X is on the wrong device or has the wrong number of features")
        m, d = X.shape
        A, k = len(acq_funcs), int(top_k)
        self._attach_kinv()
        prm = _acq_params(y_opt, xi, kappa, acq_funcs, 0, resolve_precision(precision, self.host.n_train, m))
        g = _lib.GradOut()
        if out is None:
            new = lambda *shape: torch.empty(*shape, dtype=torch.float64, device=X.device)   # noqa: E731
            out = {"mean": new(m), "std": new(m), "mean_grad": new(m, d), "std_grad": new(m, d)}
            for name in acq_funcs:
                out[name] = {"values": new(m), "grad": new(m, d)}
            if k > 0:
                out["record"] = torch.empty(record_words(A, k), dtype=torch.int64, device=X.device)
        g.mean, g.std = out["mean"].data_ptr(), out["std"].data_ptr()
        g.mean_grad, g.std_grad = out["mean_grad"].data_ptr(), out["std_grad"].data_ptr()
        for name in acq_funcs:
            a = _lib.ACQ_INDEX[name]
            g.acq[a], g.acq_grad[a] = out[name]["values"].data_ptr(), out[name]["grad"].data_ptr()
        s = stream if stream is not None else torch.cuda.current_stream(X.device)
        sp = C.c_void_p(s.cuda_stream)
        L = _lib.lib()
        _lib.check(L.skb_score_grad_dev(self._h, X.data_ptr() if m else None, m, C.byref(prm), C.byref(g), sp))
        if k > 0:
            base = out["record"].data_ptr()
            for j, name in enumerate(acq_funcs):
                _lib.check(L.skb_topk_dev(self._h, out[name]["values"].data_ptr() if m else None, m, int(index_offset), k,
                                          base + 8 * (j * k), base + 8 * (A * k + j * k), base + 8 * (2 * A * k + j), sp))
        return out


def record_words(n_acq, top_k):
    """8-byte words of a packed top-k record (include/skb.h, skb_topk_merge_records_dev)."""
    return int(n_acq) * (2 * int(top_k) + 1)


def unpack_record(rec, acq_funcs, top_k):
    """{acq name: {'topk_val', 'topk_idx', 'first_nan'}} (NumPy) from a packed record (int64 tensor or array, host or
    device): ONE device -> host copy for everything a sharded scoring call returns."""
    words = rec.cpu().numpy() if hasattr(rec, "cpu") else np.asarray(rec)
    if hasattr(rec, "is_cuda") and rec.is_cuda:
        from .distributed import check_exchanges
        check_exchanges()                 # a peer-memory exchange that gave up waiting for a rank raises here
    A, k = len(acq_funcs), int(top_k)
    out = {}
    for j, name in enumerate(acq_funcs):
        out[name] = {"topk_val": words[j * k:(j + 1) * k].view(np.float64).copy(),
                     "topk_idx": words[A * k + j * k:A * k + (j + 1) * k].copy(),
                     "first_nan": int(words[2 * A * k + j])}
    return out


def merge_records_device(gathered, n_records, n_acq, top_k, stream=None):
    """Merge `n_records` packed records (one int64 CUDA tensor, back to back) into one: a single launch of
    topk_merge_records_kernel through skb_topk_merge_records_dev."""
    import torch
    merged = torch.empty(record_words(n_acq, top_k), dtype=torch.int64, device=gathered.device)
    s = stream if stream is not None else torch.cuda.current_stream(gathered.device)
    _lib.check(_lib.lib().skb_topk_merge_records_dev(gathered.device.index, gathered.data_ptr(), int(n_records), int(n_acq),
                                                     int(top_k), merged.data_ptr(), C.c_void_p(s.cuda_stream)))
    return merged


def kernel_gradient_x(kernel, x, X_train, device=None):
    """`kernel.gradient_x(x, X_train)`: d k(x, X_i) / dx for every training row, shape (n_samples, n_features)
    (skopt/learning/gaussian_process/kernels.py:48-65 and the per-class forms :68-308), on the GPU.

    There is one device implementation of the radial derivatives - the gradient kernel K3g behind skb_score_grad_host -
    and this reaches it without a second code path: for a stationary kernel d k(x, X_i) / dx = -d k(z, x) / dz at z = X_i,
    and the right-hand side is the posterior-mean gradient of a one-point GP (training set {x}, alpha = 1, no target
    scaling) evaluated at the rows of X_train.  The only value that is not antisymmetric is the reference's convention for
    Matern(nu=0.5) at zero distance (the row -amplitude / length_scale, kernels.py:118-128): those rows keep their sign."""
    x = np.asarray(x, dtype=np.float64).ravel()
    X_train = np.atleast_2d(np.asarray(X_train, dtype=np.float64))
    n, d = X_train.shape
    if x.shape[0] != d:
        raise ValueError("x has %d features, X_train has %d" % (x.shape[0], d))
    aff = _flatten_kernel(kernel)
    if aff.base is None:                           # constants and white noise do not depend on x (kernels.py:260-269)
        return np.zeros((n, d))
    (kind, ls), amplitude = aff.base, aff.a
    if kind == _lib.KERNEL_HAMMING:
        raise NotImplementedError("HammingKernel has no gradient_x (categorical inputs)")
    if device is None:
        from .learning.gaussian_process.gpr import _DEFAULT_DEVICE
        device = _DEFAULT_DEVICE[0]
    one = np.ones((1, 1))
    dev = DeviceState(HostState(x[None, :], np.ones(1), one, ls, kind, amplitude, aff.c, 0.0, 0.0, 1.0, K_inv=one), device)
    try:
        mirrored = dev.score_grad(X_train, acq_funcs=())["mean_grad"]
    finally:
        dev.close()
    grad = -mirrored
    if kind == _lib.KERNEL_MATERN12:
        same = np.all(X_train == x, axis=1)
        grad[same] = mirrored[same]
    return grad


def measure_peaks(device=0):
    """{'i8_mac_per_clk_sm', 'i8_tops', 'fp64_dmma_tflops'} measured on the device now (skb_measure_peaks)."""
    a, b, c = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
    _lib.check(_lib.lib().skb_measure_peaks(int(device), C.byref(a), C.byref(b), C.byref(c)))
    return {"i8_mac_per_clk_sm": float(a.value), "i8_tops": float(b.value), "fp64_dmma_tflops": float(c.value)}


def launch_count():
    return int(_lib.lib().skb_launch_count())


# ----------------------------------------------------------------------------------------------
# candidate generation on the device (bit-identical to Space.rvs_transformed)
# ----------------------------------------------------------------------------------------------
def device_uniform(rng, count, device=0):
    """`rng.uniform(0, 1, count)` computed on the GPU from the RandomState's own MT19937 state; the RandomState is left
    exactly where the host call would have left it.  Returns a float64 CUDA tensor."""
    import torch
    name, key, pos, has_gauss, cached = rng.get_state()
    if name != "MT19937":
        raise ValueError("device candidate generation needs a legacy MT19937 RandomState, got %s" % name)
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    pos_c = C.c_int32(int(pos))
    out = torch.empty(int(count), dtype=torch.float64, device=torch.device("cuda", device))
    s = torch.cuda.current_stream(out.device)
    _lib.check(_lib.lib().skb_mt19937_uniform_dev(int(device), key.ctypes.data, C.byref(pos_c), int(count),
                                                  out.data_ptr(), C.c_void_p(s.cuda_stream)))
    rng.set_state((name, key, int(pos_c.value), has_gauss, cached))
    return out


# the parallel (jump-ahead) generator needs one polynomial per dimension from the host (~25 ms each, cached per
# (n_points, n_dims)) and ~1.5 ms of jump kernel per call: worth it once the serial walk takes several ms.  In the
# automatic mode the polynomials are computed in a background thread and the serial walk (same bits) serves until
# they are there.  Measured at 10^6 x 20: 22.5 ms serial, 4.9 ms with the jump (tools/time_candidates.py).
JUMP_MIN_DOUBLES = 1 << 20
_JUMP_DEFAULT = "auto"


def _jump_wanted(jump, n_samples, d):
    if 2 * int(n_samples) < 624 or d < 2:
        return False
    if jump is None:
        jump = os.environ.get("SKB_MT_JUMP", _JUMP_DEFAULT)
    if jump in (True, "1", "always"):
        return True
    if jump in (False, "0", "never"):
        return False
    return int(n_samples) * d >= JUMP_MIN_DOUBLES


def device_candidates(dim_descs, n_samples, rng, device=0, jump=None):
    """Candidate matrix (n_samples x n_dims, CUDA float64) for dimensions described as (kind, low, high) triples, with
    the draw order and arithmetic of `space.transform(space.rvs(n, rng))`; `rng` is advanced as the host call would.

    jump: None = automatic (one CTA per dimension, each started from its own point of the stream by an MT19937 jump, for
    large candidate sets - mt_jump.py, csrc/rng.cuh; the serial walk otherwise), True / False to force either; the
    environment variable SKB_MT_JUMP (auto / 1 / 0) sets the default.  Both produce the same bits and leave `rng` in
    the state NumPy itself would be in."""
    import torch
    name, key, pos, has_gauss, cached = rng.get_state()
    if name != "MT19937":
        raise ValueError("device candidate generation needs a legacy MT19937 RandomState, got %s" % name)
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    d = len(dim_descs)
    X = torch.empty((int(n_samples), d), dtype=torch.float64, device=torch.device("cuda", device))
    arr = (_lib.DimDesc * d)()
    for j, (kind, lo, hi) in enumerate(dim_descs):
        arr[j].kind, arr[j].low, arr[j].high = int(kind), float(lo), float(hi)
    s = torch.cuda.current_stream(X.device)
    polys = None
    if _jump_wanted(jump, n_samples, d):
        from . import mt_jump
        forced = jump in (True, "1", "always") or os.environ.get("SKB_MT_JUMP") in ("1", "always")
        polys = (mt_jump.column_jump_polynomials if forced else mt_jump.column_jump_polynomials_nowait)(2 * int(n_samples), d)
    if polys is not None:
        window = mt_jump.window_from_state(key, pos)
        tail = np.empty(2 * mt_jump.N, dtype=np.uint32)
        _lib.check(_lib.lib().skb_candidates_generate_jump_dev(int(device), window.ctypes.data, polys.ctypes.data,
                                                               int(n_samples), d, arr, X.data_ptr(), tail.ctypes.data,
                                                               C.c_void_p(s.cuda_stream)))
        key, new_pos = mt_jump.canonical_state(pos, 2 * int(n_samples) * d, tail)
        rng.set_state((name, key, new_pos, has_gauss, cached))
        return X
    pos_c = C.c_int32(int(pos))
    _lib.check(_lib.lib().skb_candidates_generate_dev(int(device), key.ctypes.data, C.byref(pos_c), int(n_samples), d, arr,
                                                      X.data_ptr(), C.c_void_p(s.cuda_stream)))
    rng.set_state((name, key, int(pos_c.value), has_gauss, cached))
    return X


# ---------------------------------------------------------------------------------------------------------------
# GP fit objective on the device (opt-in): log marginal likelihood + gradient, include/skb.h skb_fit_*
# ---------------------------------------------------------------------------------------------------------------
class FitLayout:
    """How sklearn's `theta` (the log of the NON-fixed hyper-parameters, in `kernel.hyperparameters` order) maps onto
    the device objective's full vector [log amplitude, log length_scale[d], log noise]."""

    def __init__(self, kernel, n_dims):
        self.kind = None
        self.n_dims = n_dims
        self.full = np.empty(n_dims + 2)            # fixed entries keep their value, free ones are overwritten
        self.full[0], self.full[1:n_dims + 1], self.full[n_dims + 1] = 0.0, 0.0, -np.inf
        self.slots = []                             # per theta entry: indices of `full` it drives
        self._walk(kernel)
        if self.kind is None:
            raise NotImplementedError("no stationary base kernel")
        params = kernel.get_params()
        for hp in kernel.hyperparameters:
            leaf = hp.name.rsplit("__", 1)[-1]
            value = np.atleast_1d(np.asarray(params[hp.name], dtype=np.float64))
            if leaf == "constant_value":
                idx = [[0]]
            elif leaf == "length_scale":
                idx = [[1 + k] for k in range(n_dims)] if value.size == n_dims and hp.n_elements == n_dims \
                    else [list(range(1, n_dims + 1))]
                if value.size not in (1, n_dims):
                    raise NotImplementedError("length_scale with %d entries for %d dimensions" % (value.size, n_dims))
            elif leaf == "noise_level":
                idx = [[n_dims + 1]]
            else:
                raise NotImplementedError("hyper-parameter %s" % hp.name)
            with np.errstate(divide="ignore"):
                logs = np.log(value)
            for j, group in enumerate(idx):
                self.full[group] = logs[j] if logs.size > 1 else logs[0]
            if not hp.fixed:
                self.slots.extend(idx)
        self.n_theta = len(self.slots)

    def _walk(self, kernel):
        """Accept exactly  [Constant *] {RBF | Matern} [+ White]  (either operand order)."""
        cls = type(kernel).__name__
        if cls == "Sum":
            kinds = sorted([type(kernel.k1).__name__, type(kernel.k2).__name__])
            if "WhiteKernel" not in kinds or kinds.count("WhiteKernel") != 1:
                raise NotImplementedError("Sum without exactly one WhiteKernel")
            self._walk(kernel.k2 if type(kernel.k1).__name__ == "WhiteKernel" else kernel.k1)
        elif cls == "Product":
            kinds = [type(kernel.k1).__name__, type(kernel.k2).__name__]
            if kinds.count("ConstantKernel") != 1:
                raise NotImplementedError("Product without exactly one ConstantKernel")
            self._walk(kernel.k2 if kinds[0] == "ConstantKernel" else kernel.k1)
        elif cls in ("RBF", "Matern"):
            if self.kind is not None:
                raise NotImplementedError("more than one stationary kernel")
            self.kind = _flatten_kernel(kernel).base[0]
        else:
            raise NotImplementedError("kernel %s" % cls)

    def expand(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        if theta.shape != (self.n_theta,):
            raise ValueError("theta has %s entries, the kernel has %d free hyper-parameters" % (theta.shape, self.n_theta))
        full = self.full.copy()
        for t, group in zip(theta, self.slots):
            full[group] = t
        return full

    def select(self, grad_full):
        return np.array([grad_full[group].sum() for group in self.slots], dtype=np.float64)


class DeviceFit:
    """Training set resident on the GPU for the duration of one hyper-parameter search."""

    def __init__(self, X_train, y_train, kernel_kind, alpha, device=0):
        self.X = np.ascontiguousarray(X_train, dtype=np.float64)
        self.y = np.ascontiguousarray(y_train, dtype=np.float64)
        if self.y.shape != (self.X.shape[0],):
            raise ValueError("y_train must be 1-D with one entry per training row")
        desc = _lib.FitDesc(self.X.shape[0], self.X.shape[1], int(kernel_kind), 0, float(alpha),
                            self.X.ctypes.data, self.y.ctypes.data)
        self._h = C.c_void_p()
        self.device = int(device)
        _lib.check(_lib.lib().skb_fit_create(C.byref(desc), self.device, C.byref(self._h)))
        self.n_evals = 0

    def lml(self, theta_full, eval_gradient=True):
        theta_full = np.ascontiguousarray(theta_full, dtype=np.float64)
        if theta_full.shape != (self.X.shape[1] + 2,):
            raise ValueError("theta_full must have n_dims + 2 entries")
        value = C.c_double()
        grad = np.zeros(theta_full.shape[0])
        _lib.check(_lib.lib().skb_fit_lml_host(self._h, theta_full.ctypes.data, 1 if eval_gradient else 0,
                                               C.byref(value), grad.ctypes.data))
        self.n_evals += 1
        return (value.value, grad) if eval_gradient else value.value

    def lml_batch(self, thetas_full):
        """Values and gradients for several hyper-parameter vectors, evaluated concurrently on the GPU; row i is
        bit-identical to lml(thetas_full[i])."""
        thetas_full = np.ascontiguousarray(thetas_full, dtype=np.float64)
        if thetas_full.ndim != 2 or thetas_full.shape[1] != self.X.shape[1] + 2:
            raise ValueError("thetas_full must have shape (count, n_dims + 2)")
        count = thetas_full.shape[0]
        values = np.empty(count)
        grads = np.zeros_like(thetas_full)
        _lib.check(_lib.lib().skb_fit_lml_batch_host(self._h, count, thetas_full.ctypes.data, 1, values.ctypes.data,
                                                     grads.ctypes.data))
        self.n_evals += count
        return values, grads

    def finalize(self, theta_full, want_factor=True):
        """Tail of sklearn's fit at the chosen hyper-parameters (_gpr.py:341-367) on the device: returns
        (lml, alpha_, L_) with L_ lower triangular like SciPy's (None unless want_factor).  Raises LinAlgError when K is
        not positive definite.  The factor stays on the device for make_state()."""
        theta_full = np.ascontiguousarray(theta_full, dtype=np.float64)
        n = self.X.shape[0]
        if theta_full.shape != (self.X.shape[1] + 2,):
            raise ValueError("theta_full must have n_dims + 2 entries")
        value = C.c_double()
        alpha = np.empty(n)
        factor = np.empty((n, n)) if want_factor else None
        rc = _lib.lib().skb_fit_finalize_host(self._h, theta_full.ctypes.data, C.byref(value), alpha.ctypes.data,
                                              factor.ctypes.data if want_factor else None)
        if rc == _lib.SKB_ERR_NOT_POSDEF:
            raise np.linalg.LinAlgError(_lib.lib().skb_last_error().decode("utf-8", "replace"))
        _lib.check(rc)
        self.n_evals += 1
        return value.value, alpha, factor

    def make_state(self, scalars):
        """DeviceState built on the GPU from the finalised factor (skb_state_create_from_fit): L^-1, K^-1 and their
        packed forms never exist on the host."""
        desc = _lib.StateDesc()
        _fill_desc(desc, scalars)
        handle = C.c_void_p()
        _lib.check(_lib.lib().skb_state_create_from_fit(self._h, C.byref(desc), C.byref(handle)))
        return DeviceState(scalars, self.device, _handle=handle)

    def close(self):
        if self._h is not None and self._h.value:
            _lib.lib().skb_fit_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:     # noqa: BLE001 - interpreter shutdown
            pass
