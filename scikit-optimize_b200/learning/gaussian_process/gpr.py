"""GaussianProcessRegressor whose posterior (mean, std, input gradients) is evaluated on a B200.

Drop-in for skopt.learning.GaussianProcessRegressor (skopt/learning/gaussian_process/gpr.py:39-368): same
constructor, same fitted attributes (noise_, K_inv_, y_train_mean_, y_train_std_), same predict() signature, return
shapes and exceptions.  fit() stays on the host (scikit-learn's marginal-likelihood optimisation) and
produces the state; predict() ships that state to the GPU once and runs the CUDA kernels.  There is no CPU fallback
for the posterior: without the CUDA library or a B200, predict() on a fitted model raises.
"""
import copy
import warnings

import numpy as np
from scipy.linalg import solve_triangular
from sklearn.base import clone
from sklearn.gaussian_process import GaussianProcessRegressor as _SkGPR
from sklearn.preprocessing._data import _handle_zeros_in_scale
from sklearn.utils import check_array, check_random_state
from sklearn.utils.validation import validate_data

from ... import engine
from .kernels import RBF, ConstantKernel, Sum, WhiteKernel

_DEFAULT_DEVICE = [0]


def set_default_device(index):
    """CUDA device new models are placed on (one process per GPU: pass LOCAL_RANK)."""
    _DEFAULT_DEVICE[0] = int(index)


_FIT_DEVICE = ["host"]


def set_fit_device(where):
    """Where the hyper-parameter search evaluates the log marginal likelihood: "host" (default: scikit-learn's own
    code, bit-identical hyper-parameters and hence trajectories to the reference) or "gpu" (skb_fit_lml_host: same
    objective within ~1e-12, hundreds of times faster at n >= 1000; the optimiser may then stop at hyper-parameters that
    differ in the last digits).  An estimator attribute `_skb_fit_device` overrides the global setting."""
    if where not in ("host", "gpu"):
        raise ValueError("fit device must be 'host' or 'gpu', got %r" % (where,))
    _FIT_DEVICE[0] = where


def _white_kernel_key(kernel, prefix=""):
    """get_params() key of the first WhiteKernel that is a direct operand in a tree of Sums (gpr.py:17-36)."""
    if isinstance(kernel, Sum):
        for name in ("k1", "k2"):
            child = getattr(kernel, name)
            if isinstance(child, WhiteKernel):
                return prefix + name
            found = _white_kernel_key(child, prefix + name + "__")
            if found is not None:
                return found
    return None


def _param_for_white_kernel_in_Sum(kernel, kernel_str=""):
    """The reference's spelling of the lookup above: (found, key), with "_" as the key of a miss (gpr.py:17-36)."""
    key = _white_kernel_key(kernel, kernel_str + "__" if kernel_str else "")
    return (True, key) if key is not None else (False, "_")


class GaussianProcessRegressor(_SkGPR):
    def __init__(self, kernel=None, alpha=1e-10, optimizer="fmin_l_bfgs_b", n_restarts_optimizer=0,
                 normalize_y=False, copy_X_train=True, random_state=None, noise=None):
        self.noise = noise
        super().__init__(kernel=kernel, alpha=alpha, optimizer=optimizer, n_restarts_optimizer=n_restarts_optimizer,
                         normalize_y=normalize_y, copy_X_train=copy_X_train, random_state=random_state)

    # device handles must not be pickled / cloned (checkpoint row)
    def __getstate__(self):
        state = super().__getstate__()
        state = dict(state)
        state.pop("_skb_device_state", None)
        state.pop("_skb_device_fit", None)
        state.pop("_skb_lockstep", None)
        return state

    def _close_device_fit(self):
        cached = self.__dict__.pop("_skb_device_fit", None)
        self.__dict__.pop("_skb_lockstep", None)
        if cached is not None and cached[1] is not None:
            cached[1].close()

    def _device_fit(self):
        """(DeviceFit, FitLayout) for the current training set when the fit objective runs on the GPU, else None."""
        where = self.__dict__.get("_skb_fit_device", _FIT_DEVICE[0])
        if where != "gpu" or np.ndim(self.alpha) != 0 or np.ndim(self.y_train_) != 1:
            return None
        cached = self.__dict__.get("_skb_device_fit")
        if cached is None or cached[0] is not self.X_train_:
            self._close_device_fit()
            try:
                layout = engine.FitLayout(self.kernel_, self.X_train_.shape[1])
            except NotImplementedError:
                layout = None
            fit = None if layout is None else engine.DeviceFit(self.X_train_, self.y_train_, layout.kind, self.alpha,
                                                               _DEFAULT_DEVICE[0])
            cached = (self.X_train_, fit, layout)
            self.__dict__["_skb_device_fit"] = cached
            self.__dict__.pop("_skb_lockstep", None)
        return None if cached[1] is None else (cached[1], cached[2])

    def log_marginal_likelihood(self, theta=None, eval_gradient=False, clone_kernel=True):
        """sklearn _gpr.py:541-657.  With the fit device set to "gpu" (set_fit_device / `_skb_fit_device`) and a kernel
        of the form [Constant *] {RBF | Matern} [+ White], the value and gradient come from the GPU objective; any
        other kernel, multi-output targets or an array-valued alpha use scikit-learn's host code."""
        dev = None if theta is None else self._device_fit()
        if dev is None:
            return super().log_marginal_likelihood(theta, eval_gradient=eval_gradient, clone_kernel=clone_kernel)
        fit, layout = dev
        if not clone_kernel:
            self.kernel_.theta = theta              # the side effect sklearn's fit relies on (_gpr.py:568-572)
        out = fit.lml(layout.expand(theta), eval_gradient=eval_gradient)
        if eval_gradient:
            return out[0], layout.select(out[1])
        return out

    def _constrained_optimization(self, obj_func, initial_theta, bounds):
        """sklearn _gpr.py:658-677.  scikit-learn runs the L-BFGS-B search from the kernel's theta and then from
        `n_restarts_optimizer` random starts, one after the other (_gpr.py:316-337).  With the GPU objective all starts
        advance in lock step instead: the random starts are pre-drawn from a COPY of the estimator's generator (the
        draws do not depend on the search results, so scikit-learn later draws the very same vectors), every round
        evaluates the pending points of all searches concurrently (skb_fit_lml_batch_host), and each call returns its
        search's result.  Each search sees exactly the iterates it would see alone (optimizer/lbfgs.py)."""
        dev = self._device_fit() if self.optimizer == "fmin_l_bfgs_b" else None
        if dev is None:
            return super()._constrained_optimization(obj_func, initial_theta, bounds)
        fit, layout = dev
        cache = self.__dict__.get("_skb_lockstep")
        if cache is None:
            import copy

            from ...optimizer.lbfgs import minimize_batch
            rng = copy.deepcopy(self._rng)
            starts = [np.array(initial_theta, dtype=np.float64)]
            if np.isfinite(bounds).all():
                starts += [rng.uniform(bounds[:, 0], bounds[:, 1]) for _ in range(self.n_restarts_optimizer)]

            def evaluate(tasks, thetas):
                values, grads = fit.lml_batch(np.array([layout.expand(t) for t in thetas]))
                return -values, -np.array([layout.select(g) for g in grads])

            results = minimize_batch(starts, evaluate, [tuple(b) for b in bounds], maxiter=15000)
            cache = {"starts": starts, "results": results, "next": 0}
            self.__dict__["_skb_lockstep"] = cache
        i = cache["next"]
        cache["next"] += 1
        if i >= len(cache["starts"]) or not np.array_equal(cache["starts"][i], initial_theta):
            return super()._constrained_optimization(obj_func, initial_theta, bounds)      # not the sequence fit() issues
        theta_opt, func_min, info = cache["results"][i]
        if info["warnflag"] != 0:
            from sklearn.exceptions import ConvergenceWarning
            warnings.warn("lbfgs failed to converge after %d iteration(s) (warnflag=%d)." % (info["nit"], info["warnflag"]),
                          ConvergenceWarning)
        return theta_opt, func_min

    def _search_on_device(self, X, y):
        """scikit-learn's fit up to the choice of the hyper-parameters (_gpr.py:233-340) with the objective on the GPU.
        Same sequence of calls as scikit-learn makes - the search from the kernel's theta, then `n_restarts_optimizer`
        searches from starts drawn from `self._rng`, each through `_constrained_optimization` - so the lock-step
        machinery (and anything a subclass overrides) sees exactly what it would see under scikit-learn's fit.
        Returns the full device hyper-parameter vector at the optimum, or None when this model is outside the device
        objective's grammar (the caller then runs scikit-learn's fit)."""
        where = self.__dict__.get("_skb_fit_device", _FIT_DEVICE[0])
        if where != "gpu" or np.ndim(self.alpha) != 0 or self.optimizer not in (None, "fmin_l_bfgs_b"):
            return None
        self._validate_params()
        X, y = validate_data(self, X, y, multi_output=True, y_numeric=True, ensure_2d=True, dtype="numeric")
        if y.ndim != 1 or (self.n_targets is not None and self.n_targets != 1):
            return None
        self.kernel_ = clone(self.kernel)
        self._rng = check_random_state(self.random_state)
        if self.normalize_y:
            self._y_train_mean = np.mean(y, axis=0)
            self._y_train_std = _handle_zeros_in_scale(np.std(y, axis=0), copy=False)
            y = (y - self._y_train_mean) / self._y_train_std
        else:
            self._y_train_mean = np.zeros(shape=1)
            self._y_train_std = np.ones(shape=1)
        self.X_train_ = np.copy(X) if self.copy_X_train else X
        self.y_train_ = np.copy(y) if self.copy_X_train else y
        dev = self._device_fit()
        if dev is None:                         # kernel outside [Constant *] {RBF | Matern} [+ White]
            return None
        layout = dev[1]
        self.__dict__.pop("log_marginal_likelihood_value_", None)
        if self.optimizer is not None and self.kernel_.n_dims > 0:
            def obj_func(theta, eval_gradient=True):
                if eval_gradient:
                    lml, grad = self.log_marginal_likelihood(theta, eval_gradient=True, clone_kernel=False)
                    return -lml, -grad
                return -self.log_marginal_likelihood(theta, clone_kernel=False)

            bounds = self.kernel_.bounds
            found = [self._constrained_optimization(obj_func, self.kernel_.theta, bounds)]
            if self.n_restarts_optimizer > 0:
                if not np.isfinite(bounds).all():
                    raise ValueError("Multiple optimizer restarts (n_restarts_optimizer>0) requires that all bounds are "
                                     "finite.")
                for _ in range(self.n_restarts_optimizer):
                    start = self._rng.uniform(bounds[:, 0], bounds[:, 1])
                    found.append(self._constrained_optimization(obj_func, start, bounds))
            best = int(np.argmin([f[1] for f in found]))
            self.kernel_.theta = found[best][0]
            self.kernel_._check_bounds_params()
            self.log_marginal_likelihood_value_ = -found[best][1]
        return layout.expand(self.kernel_.theta)

    def _finalize_on_device(self, theta_full):
        """Tail of the fit on the GPU: L_, alpha_ (sklearn _gpr.py:341-367) and the scoring state (gpr.py:223-224) from
        the training set that is still resident from the search."""
        fit = self._device_fit()[0]
        lml, self.alpha_, self.L_ = fit.finalize(theta_full)
        if "log_marginal_likelihood_value_" not in self.__dict__:
            self.log_marginal_likelihood_value_ = lml
        self.y_train_std_ = self._y_train_std
        self.y_train_mean_ = self._y_train_mean
        state = fit.make_state(engine.state_scalars_from_estimator(self))
        if state is not None:
            self.__dict__["_skb_device_state"] = state

    def __getattr__(self, name):
        # K_inv_ and the triangular inverse are fitted attributes of the reference (gpr.py:223-224).  After a fit that
        # was finalised on the GPU they only exist there; the host copies are produced from L_ on first access.
        if name in ("K_inv_", "_L_inv_lower") and "L_" in self.__dict__:
            L_inv = self.__dict__.get("_L_inv_lower")
            if L_inv is None:
                L_inv = solve_triangular(self.L_, np.eye(self.L_.shape[0]), lower=True)
                self.__dict__["_L_inv_lower"] = L_inv
            if name == "K_inv_":
                self.__dict__["K_inv_"] = L_inv.T.dot(L_inv)
            return self.__dict__[name]
        raise AttributeError("%r object has no attribute %r" % (type(self).__name__, name))

    def fit(self, X, y):
        """gpr.py:166-237: add the noise kernel, fit with scikit-learn, silence the noise for prediction,
        precompute K_inv_ (here together with the triangular factor the GPU kernels contract with).  With the fit
        device set to "gpu" the whole fit runs on the device: the search (log_marginal_likelihood), the final
        factorisation and the triangular inverse / K^-1 the scoring kernels read; K_inv_ is then materialised on the
        host only if somebody reads it."""
        if isinstance(self.noise, str) and self.noise != "gaussian":
            raise ValueError("expected noise to be 'gaussian', got %s" % self.noise)
        if self.kernel is None:
            self.kernel = ConstantKernel(1.0, constant_value_bounds="fixed") * RBF(1.0, length_scale_bounds="fixed")
        if self.noise == "gaussian":
            self.kernel = self.kernel + WhiteKernel()
        elif self.noise:
            self.kernel = self.kernel + WhiteKernel(noise_level=self.noise, noise_level_bounds="fixed")
        for stale in ("_skb_device_state", "K_inv_", "_L_inv_lower"):
            self.__dict__.pop(stale, None)
        try:
            theta_full = self._search_on_device(X, y)
            if theta_full is None:
                super().fit(X, y)

            self.noise_ = None
            if self.noise:
                # K(X*, X*) must not contain the observation noise (GPML eq. 2.24); alpha_ already has it.
                if isinstance(self.kernel_, WhiteKernel):
                    self.kernel_.set_params(noise_level=0.0)
                else:
                    key = _white_kernel_key(self.kernel_)
                    if key is not None:
                        self.noise_ = self.kernel_.get_params()[key].noise_level
                        self.kernel_.set_params(**{key: WhiteKernel(noise_level=0.0)})
            if theta_full is not None:
                self._finalize_on_device(theta_full)
        finally:
            self._close_device_fit()            # the training set only stays on the GPU during the fit

        if theta_full is None:
            n = self.L_.shape[0]
            self._L_inv_lower = solve_triangular(self.L_, np.eye(n), lower=True)
            self.K_inv_ = self._L_inv_lower.T.dot(self._L_inv_lower)
            self.y_train_std_ = self._y_train_std
            self.y_train_mean_ = self._y_train_mean
        return self

    def append_observation(self, x, y):
        """A fitted copy of this model with ONE more observation (x, y) and the SAME hyper-parameters: the Cholesky factor
        grows by one row, L' = [[L, 0], [l^T, lam]] with l = L^-1 k(X, x), lam^2 = k(x, x) + noise + alpha - l.l, its
        inverse by the row [-l^T L^-1 / lam, 1 / lam], and alpha_ = L'^-T L'^-1 y' for the re-normalised targets -
        O(n^2) instead of the O(n^3) per evaluation of a new hyper-parameter search.  This is what
        GaussianProcessRegressor(kernel=self.kernel_ [+ the fitted noise], optimizer=None).fit(X', y') computes (tests/
        test_rank1_append.py); the reference's constant liar (optimizer.py:388-417) instead refits the hyper-parameters on
        every lie, so the Optimizer only uses this with acq_optimizer_kwargs={"liar_update": "rank1"}.
        x: one point in the model's (transformed) input space; y: its raw target value."""
        if not hasattr(self, "X_train_") or "L_" not in self.__dict__:
            raise ValueError("append_observation needs a fitted model")
        if np.ndim(self.alpha) != 0 or np.ndim(self.y_train_) != 1:
            raise NotImplementedError("append_observation supports a scalar alpha and a single target")
        x = np.asarray(x, dtype=np.float64).reshape(1, -1)
        if x.shape[1] != self.X_train_.shape[1]:
            raise ValueError("x has %d features, the model was fitted with %d" % (x.shape[1], self.X_train_.shape[1]))
        n = self.X_train_.shape[0]
        L_inv = self._L_inv_lower                           # materialised from L_ on first use (__getattr__)
        k_vec = self.kernel_(self.X_train_, x)[:, 0]        # kernel_ carries no observation noise any more (fit)
        kappa = float(self.kernel_(x)[0, 0]) + (self.noise_ or 0.0) + float(self.alpha)
        l = L_inv.dot(k_vec)
        lam2 = kappa - float(l.dot(l))
        if not lam2 > 0.0:
            raise np.linalg.LinAlgError("the kernel matrix with the appended point is not positive definite")
        lam = np.sqrt(lam2)
        new = copy.copy(self)                               # shares kernel_ and the constructor parameters
        new.__dict__ = dict(self.__dict__)
        for stale in ("_skb_device_state", "K_inv_", "_skb_device_fit", "log_marginal_likelihood_value_"):
            new.__dict__.pop(stale, None)
        L_new = np.zeros((n + 1, n + 1))
        L_new[:n, :n] = self.L_
        L_new[n, :n] = l
        L_new[n, n] = lam
        Li_new = np.zeros((n + 1, n + 1))
        Li_new[:n, :n] = L_inv
        Li_new[n, :n] = -(l.dot(L_inv)) / lam
        Li_new[n, n] = 1.0 / lam
        y_raw = np.append(self.y_train_ * self._y_train_std + self._y_train_mean, float(y))
        if self.normalize_y:
            new._y_train_mean = np.mean(y_raw, axis=0)
            new._y_train_std = _handle_zeros_in_scale(np.std(y_raw, axis=0), copy=False)
        y_new = (y_raw - new._y_train_mean) / new._y_train_std
        new.y_train_mean_, new.y_train_std_ = new._y_train_mean, new._y_train_std
        new.X_train_ = np.vstack([self.X_train_, x])
        new.y_train_ = y_new
        new.L_ = L_new
        new._L_inv_lower = Li_new
        new.alpha_ = Li_new.T.dot(Li_new.dot(y_new))
        new.log_marginal_likelihood_value_ = float(-0.5 * y_new.dot(new.alpha_) - np.log(np.diag(L_new)).sum()
                                                   - 0.5 * (n + 1) * np.log(2.0 * np.pi))
        return new

    _COV_BLOCK = 1024          # points per block of a blocked joint covariance: one device call holds two blocks

    def _cov_quadratic_on_device(self, X):
        """(mean [m], K_trans K_inv K_trans^T [m, m] in kernel units) from DeviceState.predict_cov.  One call takes what the
        state's scratch holds (2048 points up to n_train = 65536); larger point sets are assembled from calls on pairs
        of 1024-point blocks, so every entry still comes from the device kernels."""
        dev = self.device_state()
        m = X.shape[0]
        B = self._COV_BLOCK
        if m <= 2 * B:
            return dev.predict_cov(X)
        mean, quad = np.empty(m), np.empty((m, m))
        starts = list(range(0, m, B))
        for bi, i0 in enumerate(starts):
            i1 = min(i0 + B, m)
            for j0 in starts[bi + 1:]:
                j1 = min(j0 + B, m)
                mu, q = dev.predict_cov(np.vstack([X[i0:i1], X[j0:j1]]))
                ni = i1 - i0
                mean[i0:i1], mean[j0:j1] = mu[:ni], mu[ni:]
                quad[i0:i1, i0:i1], quad[i0:i1, j0:j1] = q[:ni, :ni], q[:ni, ni:]
                quad[j0:j1, i0:i1], quad[j0:j1, j0:j1] = q[ni:, :ni], q[ni:, ni:]
        return mean, quad

    def device_state(self, device=None):
        """The fitted state resident on the GPU (created on first use, rebuilt after every fit)."""
        dev = self.__dict__.get("_skb_device_state")
        want = _DEFAULT_DEVICE[0] if device is None else int(device)
        if dev is None or dev.device != want:
            dev = engine.DeviceState.from_estimator(self, want)
            self.__dict__["_skb_device_state"] = dev
        return dev

    def predict(self, X, return_std=False, return_cov=False, return_mean_grad=False, return_std_grad=False):
        """gpr.py:239-368.  The gradient outputs are also available for more than one row (an extension: the
        reference raises "Not implemented for n_samples > 1"); for one row the shapes are the reference's."""
        if return_std and return_cov:
            raise RuntimeError("Not returning standard deviation of predictions when returning full covariance.")
        if return_std_grad and not return_std:
            raise ValueError("Not returning std_gradient without returning the std.")
        X = check_array(X)

        if not hasattr(self, "X_train_"):       # unfitted: GP prior (gpr.py:303-312), no arithmetic worth a launch
            if X.shape[0] != 1 and (return_mean_grad or return_std_grad):
                raise ValueError("Not implemented for n_samples > 1")
            y_mean = np.zeros(X.shape[0])
            if return_cov:
                return y_mean, self.kernel(X)
            if return_std:
                return y_mean, np.sqrt(self.kernel.diag(X))
            return y_mean

        if return_cov:
            # gpr.py:320-325.  The O(n^2 m) product K_inv K_trans^T and the O(n m^2) quadratic form run on the GPU; kernel_(X)
            # (m^2 d) is formed on the host like the reference's.  There is no host route for the quadratic form.
            y_mean, quad = self._cov_quadratic_on_device(X)
            return y_mean, (self.kernel_(X) - quad) * self.y_train_std_ ** 2

        dev = self.device_state()
        if return_mean_grad or return_std_grad:
            g = dev.score_grad(X, acq_funcs=())
            one = X.shape[0] == 1
            mean_grad = g["mean_grad"][0] if one else g["mean_grad"]
            std_grad = g["std_grad"][0] if one else g["std_grad"]
            if return_std_grad:
                return g["mean"], g["std"], mean_grad, std_grad
            if return_std:
                return g["mean"], g["std"], mean_grad
            return g["mean"], mean_grad
        if return_std:
            res = dev.score(X, acq_funcs=(), want_posterior=True, precision=engine.default_precision(self))
            if np.any(res["clamped"]):
                warnings.warn("Predicted variances smaller than 0. Setting those variances to 0.")
            return res["mean"], res["std"]
        return dev.predict_mean(X)
