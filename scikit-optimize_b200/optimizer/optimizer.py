"""Ask/tell engine whose candidate-scoring block runs on the GPU.

Mirror of skopt/optimizer/optimizer.py: Optimizer.__init__ :163-305, copy :307-333, ask :335-421, _ask :423-453,
tell :455-493, _tell :495-615, run :642-651, update_next :653-661, get_result :663-676.  Everything that is not
arithmetic on the candidate set (state, validation, RNG draw order, constant-liar loop, hedge bookkeeping) follows
the reference so that trajectories stay comparable; the scoring block (:550-608) is replaced:

  reference                                                    here
  ---------                                                    ----
  X = space.transform(space.rvs(n_points, rng))  (lists)       space.rvs_transformed: same bits, vectorised
  for acq in cand_acq_funcs_:  (posterior recomputed 3x)       ONE fused pass: posterior + EI + LCB + PI + top-k
      values = _gaussian_acquisition(X, est, ...)
      argmin / argsort[:n_restarts]                            device top-k by (value, index)
      n_restarts x fmin_l_bfgs_b(gaussian_acquisition_1D)      lock-step batched restarts, one launch per L-BFGS step
  gains_ -= est.predict(next_xs_)                              device mean-only predict
"""
import copy
import sys
import time
import warnings
from math import log
from numbers import Number

import numpy as np
from sklearn.base import clone, is_regressor
from sklearn.multioutput import MultiOutputRegressor
from sklearn.utils import check_random_state

from ..learning import GaussianProcessRegressor
from ..space import Categorical, Space
from ..utils import (check_x_in_space, cook_estimator, create_result, has_gradients, is_2Dlistlike, is_listlike,
                     normalize_dimensions)
from ..acquisition import _gaussian_acquisition, _is_device_model, device_topk
from ..distributed import sharded_score, sharded_score_tensor
from ..engine import device_candidates
from ..space import device_dim_descs
from .lbfgs import minimize_batch

_INT32_MAX = np.iinfo(np.int32).max


class _DeviceRows(object):
    """Candidate matrix living on the GPU; `X[i]` brings one row to the host (NumPy), as the scoring block needs."""

    def __init__(self, tensor):
        self.tensor = tensor
        self.shape = tuple(tensor.shape)

    def __getitem__(self, i):
        return self.tensor[int(i)].cpu().numpy()

    def rows(self, idx):
        """The rows `idx` as one (len(idx), n_dims) NumPy array: one gather on the device, one copy to the host."""
        import torch
        sel = torch.as_tensor(np.asarray(idx, dtype=np.int64), device=self.tensor.device)
        return self.tensor.index_select(0, sel).cpu().numpy()


class Optimizer(object):
    def __init__(self, dimensions, base_estimator="gp", n_random_starts=None, n_initial_points=10,
                 initial_point_generator="random", n_jobs=1, acq_func="gp_hedge", acq_optimizer="auto",
                 random_state=None, model_queue_size=None, acq_func_kwargs=None, acq_optimizer_kwargs=None):
        args = locals().copy()
        del args["self"]
        self.specs = {"args": args, "function": "Optimizer"}
        self.rng = check_random_state(random_state)

        self.acq_func = acq_func
        self.acq_func_kwargs = acq_func_kwargs
        allowed = ["gp_hedge", "EI", "LCB", "PI", "EIps", "PIps"]
        if self.acq_func not in allowed:
            raise ValueError("expected acq_func to be in %s, got %s" % (",".join(allowed), self.acq_func))
        if self.acq_func == "gp_hedge":
            self.cand_acq_funcs_ = ["EI", "LCB", "PI"]
            self.gains_ = np.zeros(3)
        else:
            self.cand_acq_funcs_ = [self.acq_func]
        if acq_func_kwargs is None:
            acq_func_kwargs = dict()
        self.eta = acq_func_kwargs.get("eta", 1.0)

        if n_random_starts is not None:
            warnings.warn("n_random_starts will be removed in favour of n_initial_points.", DeprecationWarning)
            n_initial_points = n_random_starts
        if n_initial_points < 0:
            raise ValueError("Expected `n_initial_points` >= 0, got %d" % n_initial_points)
        self._n_initial_points = n_initial_points
        self.n_initial_points_ = n_initial_points

        if isinstance(base_estimator, str):
            base_estimator = cook_estimator(base_estimator, space=dimensions,
                                            random_state=self.rng.randint(0, _INT32_MAX), n_jobs=n_jobs)
        if base_estimator is not None and not is_regressor(base_estimator):
            raise ValueError("%s has to be a regressor." % base_estimator)
        # per-second acquisitions model (objective, log time) with one surrogate each (optimizer.py:230-234)
        if "ps" in self.acq_func and base_estimator is not None and not isinstance(base_estimator, MultiOutputRegressor):
            self.base_estimator_ = MultiOutputRegressor(base_estimator)
        else:
            self.base_estimator_ = base_estimator

        if acq_optimizer == "auto":
            acq_optimizer = "lbfgs" if has_gradients(self.base_estimator_) else "sampling"
        if acq_optimizer not in ["lbfgs", "sampling"]:
            raise ValueError("Expected acq_optimizer to be 'lbfgs' or 'sampling', got {0}".format(acq_optimizer))
        if not has_gradients(self.base_estimator_) and acq_optimizer != "sampling":
            raise ValueError("The regressor {0} should run with acq_optimizer='sampling'.".format(type(base_estimator)))
        self.acq_optimizer = acq_optimizer

        if acq_optimizer_kwargs is None:
            acq_optimizer_kwargs = dict()
        self.n_points = acq_optimizer_kwargs.get("n_points", 10000)
        self.n_restarts_optimizer = acq_optimizer_kwargs.get("n_restarts_optimizer", 5)
        self.n_jobs = acq_optimizer_kwargs.get("n_jobs", 1)      # accepted for compatibility; restarts are batched
        # arithmetic of the variance contraction: "auto" (default), "fp64" or "split_int8" (see engine.resolve_precision)
        self.precision = acq_optimizer_kwargs.get("precision", None)      # None: engine.default_precision() at call time
        # draw the candidate set on the GPU when the space allows a bit-exact device version (engine.device_candidates)
        self.device_candidates = acq_optimizer_kwargs.get("device_candidates", True)
        # where the GP hyper-parameter search evaluates its objective: None = the package setting
        # (learning.gaussian_process.gpr.set_fit_device, default "host"), or "host" / "gpu" for this optimizer
        self.fit_device = acq_optimizer_kwargs.get("fit_device", None)
        if self.fit_device not in (None, "host", "gpu"):
            raise ValueError("fit_device must be None, 'host' or 'gpu', got %r" % (self.fit_device,))
        # multi-GPU scoring is opt-in: shard_group=True (default process group) or a torch.distributed group makes
        # every tell() a collective that ALL ranks of the group must enter with identical Optimizer state; the default
        # keeps tell() local so independent optimisers can live inside an initialised process group
        # constant-liar batches: "refit" (the reference, optimizer.py:388-417: hyper-parameters re-optimised on every lie)
        # or "rank1" (hyper-parameters of the current model frozen, every lie appended to its Cholesky factor in O(n^2):
        # GaussianProcessRegressor.append_observation)
        self.liar_update = acq_optimizer_kwargs.get("liar_update", "refit")
        if self.liar_update not in self.LIAR_UPDATES:
            raise ValueError("liar_update must be one of %s, got %r" % (self.LIAR_UPDATES, self.liar_update))
        shard = acq_optimizer_kwargs.get("shard_group", None)
        self._shard_group = False if shard is None or shard is False else (None if shard is True else shard)
        self.acq_optimizer_kwargs = acq_optimizer_kwargs

        if isinstance(self.base_estimator_, GaussianProcessRegressor):
            dimensions = normalize_dimensions(dimensions)
        self.space = Space(dimensions)

        if initial_point_generator not in ("random", None):
            raise NotImplementedError("only the 'random' initial point generator is available on this path "
                                      "(sobol / lhs / halton / hammersly / grid samplers are out of scope)")
        self._initial_point_generator = None
        self._initial_samples = None

        self._cat_inds = [i for i, dim in enumerate(self.space.dimensions) if isinstance(dim, Categorical)]
        self._non_cat_inds = [i for i, dim in enumerate(self.space.dimensions) if not isinstance(dim, Categorical)]

        if not isinstance(model_queue_size, (int, type(None))):
            raise TypeError("model_queue_size should be an int or None, got {}".format(type(model_queue_size)))
        self.max_model_queue_size = model_queue_size
        self.models = []
        self.Xi = []
        self.yi = []
        self.cache_ = {}
        self.scoring_times_ = []

    # ------------------------------------------------------------------------------------------------ copy / ask
    def copy(self, random_state=None):
        optimizer = Optimizer(dimensions=self.space.dimensions, base_estimator=self.base_estimator_,
                              n_initial_points=self.n_initial_points_,
                              initial_point_generator="random", acq_func=self.acq_func,
                              acq_optimizer=self.acq_optimizer, acq_func_kwargs=self.acq_func_kwargs,
                              acq_optimizer_kwargs=self.acq_optimizer_kwargs, random_state=random_state)
        optimizer._initial_samples = self._initial_samples
        if hasattr(self, "gains_"):
            optimizer.gains_ = np.copy(self.gains_)
        if self.Xi:
            optimizer._tell(self.Xi, self.yi)
        return optimizer

    LIAR_UPDATES = ("refit", "rank1")

    def _ask_batch_rank1(self, n_points, strategy):
        """ask(n_points, strategy) with liar_update="rank1": the lies are appended to the CURRENT model (no refit, not
        even of the throw-away copy): n_points x {score the candidate set, append (x, lie) to the factor}."""
        sandbox = copy.copy(self)                           # shares the space and the estimator; own lists / rng
        sandbox.Xi, sandbox.yi = list(self.Xi), list(self.yi)
        sandbox.models = list(self.models)
        sandbox.rng = np.random.RandomState(self.rng.randint(0, _INT32_MAX))
        sandbox.cache_ = {}
        sandbox.scoring_times_ = self.scoring_times_       # one history: the caller reads the per-step scoring times
        est = sandbox.models[-1]
        transformed_bounds = np.array(self.space.transformed_bounds)
        X = [self._next_x]
        pick = {"cl_min": np.min, "cl_mean": np.mean, "cl_max": np.max}[strategy]
        for step in range(n_points):
            x = X[-1]
            y_lie = float(pick(sandbox.yi))
            sandbox.Xi.append(x)
            sandbox.yi.append(y_lie)
            if step + 1 == n_points:
                break
            est = est.append_observation(self.space.transform([x])[0], y_lie)
            t_score = time.perf_counter()
            next_xs = sandbox._score_and_refine(est, transformed_bounds)
            sandbox.scoring_times_.append(time.perf_counter() - t_score)
            est.__dict__.pop("_skb_device_state", None)     # one resident state at a time
            X.append(self.space.inverse_transform(np.asarray(next_xs[0]).reshape((1, -1)))[0])
        return X

    def ask(self, n_points=None, strategy="cl_min"):
        """One point, or `n_points` points by the constant-liar strategy (cl_min / cl_mean / cl_max)."""
        if n_points is None:
            return self._ask()
        supported = ["cl_min", "cl_mean", "cl_max"]
        if not (isinstance(n_points, int) and n_points > 0):
            raise ValueError("n_points should be int > 0, got " + str(n_points))
        if strategy not in supported:
            raise ValueError("Expected parallel_strategy to be one of " + str(supported) + ", " + "got %s" % strategy)
        if (n_points, strategy) in self.cache_:
            return self.cache_[(n_points, strategy)]
        if (self.liar_update == "rank1" and self.models and self._n_initial_points <= 0 and hasattr(self, "_next_x")
                and self.acq_func in ("EI", "PI", "LCB") and hasattr(self.models[-1], "append_observation")):
            X = self._ask_batch_rank1(n_points, strategy)
            self.cache_ = {(n_points, strategy): X}
            return X

        # the lies are told to a throw-away copy (one rng draw seeds it, as in the reference)
        opt = self.copy(random_state=self.rng.randint(0, _INT32_MAX))
        X = []
        for _ in range(n_points):
            x = opt.ask()
            X.append(x)
            ti_available = "ps" in self.acq_func and len(opt.yi) > 0
            ti = [t for (_, t) in opt.yi] if ti_available else None
            if strategy == "cl_min":
                y_lie = np.min(opt.yi) if opt.yi else 0.0
                t_lie = np.min(ti) if ti is not None else log(sys.float_info.max)
            elif strategy == "cl_mean":
                y_lie = np.mean(opt.yi) if opt.yi else 0.0
                t_lie = np.mean(ti) if ti is not None else log(sys.float_info.max)
            else:
                y_lie = np.max(opt.yi) if opt.yi else 0.0
                t_lie = np.max(ti) if ti is not None else log(sys.float_info.max)
            if "ps" in self.acq_func:
                opt._tell(x, (y_lie, t_lie))       # _tell: the time is already a logarithm
            else:
                opt._tell(x, y_lie)
        self.cache_ = {(n_points, strategy): X}
        return X

    def _ask(self):
        if self._n_initial_points > 0 or self.base_estimator_ is None:
            if self._initial_samples is None:
                return self.space.rvs(random_state=self.rng)[0]
            return self._initial_samples[len(self._initial_samples) - self._n_initial_points]
        if not self.models:
            raise RuntimeError("Random evaluations exhausted and no model has been fit.")
        next_x = self._next_x
        min_delta_x = self.space.min_distance(next_x, self.Xi)
        if abs(min_delta_x) <= 1e-8:
            warnings.warn("The objective has been evaluated at this point before.")
        return next_x

    # ------------------------------------------------------------------------------------------------ tell
    def tell(self, x, y, fit=True):
        check_x_in_space(x, self.space)
        self._check_y_is_valid(x, y)
        if "ps" in self.acq_func:                  # (value, seconds) -> (value, log seconds), optimizer.py:485-491
            if is_2Dlistlike(x):
                y = [[val, log(t)] for (val, t) in y]
            elif is_listlike(x):
                y = list(y)
                y[1] = log(y[1])
        return self._tell(x, y, fit=fit)

    def _tell(self, x, y, fit=True):
        if "ps" in self.acq_func:
            if is_2Dlistlike(x):
                self.Xi.extend(x)
                self.yi.extend(y)
                self._n_initial_points -= len(y)
            elif is_listlike(x):
                self.Xi.append(x)
                self.yi.append(y)
                self._n_initial_points -= 1
        elif is_listlike(y) and is_2Dlistlike(x):
            self.Xi.extend(x)
            self.yi.extend(y)
            self._n_initial_points -= len(y)
        elif is_listlike(x):
            self.Xi.append(x)
            self.yi.append(y)
            self._n_initial_points -= 1
        else:
            raise ValueError("Type of arguments `x` (%s) and `y` (%s) not compatible." % (type(x), type(y)))
        self.cache_ = {}

        if fit and self._n_initial_points <= 0 and self.base_estimator_ is not None:
            transformed_bounds = np.array(self.space.transformed_bounds)
            est = clone(self.base_estimator_)
            if self.fit_device is not None and hasattr(est, "log_marginal_likelihood"):
                est._skb_fit_device = self.fit_device
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                est.fit(self.space.transform(self.Xi), self.yi)

            if hasattr(self, "next_xs_") and self.acq_func == "gp_hedge":
                self.gains_ -= est.predict(np.vstack(self.next_xs_))

            if self.max_model_queue_size is None or len(self.models) < self.max_model_queue_size:
                self.models.append(est)
            else:
                self.models.pop(0)
                self.models.append(est)

            # only the newest model needs to stay resident on the GPU; older ones re-upload lazily if predict()ed
            # (per-second models keep theirs on the two regressors inside MultiOutputRegressor)
            for old in self.models[:-1]:
                old.__dict__.pop("_skb_device_state", None)
                for sub in getattr(old, "estimators_", []):
                    sub.__dict__.pop("_skb_device_state", None)

            t_score = time.perf_counter()
            self.next_xs_ = self._score_and_refine(est, transformed_bounds)
            # wall time of the scoring block alone (fit excluded): the "ask() p50" of BASELINE.json -- in this
            # version of scikit-optimize the work ask() returns is done here (optimizer.py:550-608)
            self.scoring_times_.append(time.perf_counter() - t_score)

            if self.acq_func == "gp_hedge":
                logits = np.array(self.gains_)
                logits -= np.max(logits)
                exp_logits = np.exp(self.eta * logits)
                probs = exp_logits / np.sum(exp_logits)
                next_x = self.next_xs_[np.argmax(self.rng.multinomial(1, probs))]
            else:
                next_x = self.next_xs_[0]
            self._next_x = self.space.inverse_transform(next_x.reshape((1, -1)))[0]

        result = create_result(self.Xi, self.yi, self.space, self.rng, models=self.models)
        result.specs = self.specs
        return result

    def _score_and_refine(self, est, transformed_bounds):
        """The scoring block (optimizer.py:550-594) on the GPU: returns one proposed point per candidate acquisition."""
        kw = self.acq_func_kwargs or {}
        xi, kappa = kw.get("xi", 0.01), kw.get("kappa", 1.96)
        y_opt = np.min(self.yi)
        acqs = tuple(self.cand_acq_funcs_)
        if "ps" in self.acq_func or not hasattr(est, "device_state"):
            # per-second acquisitions (two models), or a regressor that is not this package's GP: the reference's
            # duck-typed boundary (optimizer.py:225-227) - posterior from the model's own predict(), acquisition
            # arithmetic and ranking on the GPU
            X = self.space.rvs_transformed(self.n_points, self.rng)
            return self._score_and_refine_through_predict(est, X, y_opt, transformed_bounds)
        dev = est.device_state()
        k = 1 if self.acq_optimizer == "sampling" else max(1, min(int(self.n_restarts_optimizer), self.n_points))
        # candidate matrix: generated ON the GPU (same MT19937 stream, same arithmetic, same bits as the host path) when
        # every dimension allows it, so neither the draw nor the upload of n_points x n_dims doubles happens on the host
        # (tiny candidate sets are cheaper to draw on the host than the device round trip of the generator state)
        big_enough = self.n_points * self.space.transformed_n_dims >= 30000 or self.device_candidates == "always"
        descs = device_dim_descs(self.space) if (self.device_candidates and big_enough) else None
        if descs is not None and isinstance(self.rng, np.random.RandomState):
            X = _DeviceRows(device_candidates(descs, self.n_points, self.rng, dev.device))
            res = sharded_score_tensor(dev, X.tensor, y_opt=y_opt, acq_funcs=acqs, xi=xi, kappa=kappa, top_k=k,
                                       group=self._shard_group, precision=self.precision)
        else:
            X = self.space.rvs_transformed(self.n_points, self.rng)
            # with shard_group set (one process per GPU): every rank holds the same RNG stream, hence the same candidate matrix; rank r scores
            # its contiguous block and the per-rank top-k lists are all-gathered (k*16 bytes per acquisition) and merged
            res = sharded_score(dev, X, y_opt=y_opt, acq_funcs=acqs, xi=xi, kappa=kappa, top_k=k,
                                group=self._shard_group, precision=self.precision)
        self._last_scoring = res          # kept for inspection / tests (top-k per acquisition)

        next_xs = []
        if self.acq_optimizer == "sampling":
            for acq in acqs:
                r = res[acq]
                # np.argmin: the first NaN wins, otherwise the first minimum
                idx = r["first_nan"] if r["first_nan"] >= 0 else int(r["topk_idx"][0])
                next_xs.append(X[idx])
        else:
            start_idx, owner = [], []
            for a_i, acq in enumerate(acqs):
                for idx in res[acq]["topk_idx"]:
                    if idx >= 0:
                        start_idx.append(int(idx))
                        owner.append(a_i)
            # all restart points of all acquisitions in one fetch (device candidates: one gather + one copy)
            rows = X.rows(start_idx) if isinstance(X, _DeviceRows) else np.asarray(X)[start_idx]
            starts = [rows[i] for i in range(len(start_idx))]

            owner_arr = np.asarray(owner, dtype=np.intp)

            def evaluate(tasks, Xq):
                g = dev.score_grad(Xq, y_opt=y_opt, acq_funcs=acqs, xi=xi, kappa=kappa)
                own, row = owner_arr[np.asarray(tasks, dtype=np.intp)], np.arange(len(tasks))
                if len(acqs) == 1:
                    return g[acqs[0]]["values"], g[acqs[0]]["grad"]
                f = np.stack([g[a]["values"] for a in acqs])[own, row]           # row i belongs to acquisition own[i]
                G = np.stack([g[a]["grad"] for a in acqs])[own, row]
                return f, G

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                results = minimize_batch(starts, evaluate, self.space.transformed_bounds, maxiter=20)
            for a_i, acq in enumerate(acqs):
                mine = [r for r, o in zip(results, owner) if o == a_i]
                cand_xs = np.array([r[0] for r in mine])
                cand_acqs = np.array([r[1] for r in mine])
                next_xs.append(cand_xs[np.argmin(cand_acqs)])

        if not self.space.is_categorical:
            next_xs = [np.clip(x, transformed_bounds[:, 0], transformed_bounds[:, 1]) for x in next_xs]
        return next_xs

    def _score_and_refine_through_predict(self, est, X, y_opt, transformed_bounds):
        """The scoring block for models reached through `predict`: EIps / PIps (`est.estimators_` = objective and
        log-time model, two resident GPs when they are this package's) and regressors of other families.  Values come
        from acquisition._gaussian_acquisition (our GPs: fused GPU pass; others: their predict(), then
        skb_acq_from_posterior_host / skb_ps_combine_host), ranking from acquisition.device_topk (skb_topk_host)."""
        models = list(est.estimators_) if "ps" in self.acq_func else [est]
        batched = all(_is_device_model(m) for m in models)      # foreign predict() returns gradients for one row only
        k = 1 if self.acq_optimizer == "sampling" else max(1, min(int(self.n_restarts_optimizer), X.shape[0]))
        next_xs = []
        for acq in self.cand_acq_funcs_:
            values = _gaussian_acquisition(X, est, y_opt, acq, acq_func_kwargs=self.acq_func_kwargs)
            tv, ti, first_nan = device_topk(values, k)
            if self.acq_optimizer == "sampling":
                next_x = X[first_nan if first_nan >= 0 else int(ti[0])]
            else:
                starts = [X[int(i)] for i in ti if i >= 0]

                def evaluate(tasks, Xq, acq=acq):
                    if batched:
                        f, G = _gaussian_acquisition(Xq, est, y_opt, acq, return_grad=True,
                                                     acq_func_kwargs=self.acq_func_kwargs)
                        return f, np.asarray(G).reshape(len(tasks), -1)
                    rows = [_gaussian_acquisition(Xq[i:i + 1], est, y_opt, acq, return_grad=True,
                                                  acq_func_kwargs=self.acq_func_kwargs) for i in range(len(tasks))]
                    return np.array([f[0] for f, _ in rows]), np.vstack([np.ravel(g) for _, g in rows])

                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    results = minimize_batch(starts, evaluate, self.space.transformed_bounds, maxiter=20)
                cand_xs = np.array([r[0] for r in results])
                cand_acqs = np.array([r[1] for r in results])
                next_x = cand_xs[np.argmin(cand_acqs)]
            if not self.space.is_categorical:
                next_x = np.clip(next_x, transformed_bounds[:, 0], transformed_bounds[:, 1])
            next_xs.append(next_x)
        return next_xs

    def _check_y_is_valid(self, x, y):
        if "ps" in self.acq_func:
            if is_2Dlistlike(x):
                if not (np.ndim(y) == 2 and np.shape(y)[1] == 2):
                    raise TypeError("expected y to be a list of (func_val, t)")
            elif is_listlike(x):
                if not (np.ndim(y) == 1 and len(y) == 2):
                    raise TypeError("expected y to be (func_val, t)")
        elif is_listlike(y) and is_2Dlistlike(x):
            for y_value in y:
                if not isinstance(y_value, Number):
                    raise ValueError("expected y to be a list of scalars")
        elif is_listlike(x):
            if not isinstance(y, Number):
                raise ValueError("`func` should return a scalar")
        else:
            raise ValueError("Type of arguments `x` (%s) and `y` (%s) not compatible." % (type(x), type(y)))

    # ------------------------------------------------------------------------------------------------ helpers
    def run(self, func, n_iter=1):
        for _ in range(n_iter):
            x = self.ask()
            self.tell(x, func(x))
        return self.get_result()

    def update_next(self):
        self.cache_ = {}
        if hasattr(self, "_next_x"):
            opt = self.copy(random_state=self.rng)
            self._next_x = opt._next_x

    def get_result(self):
        result = create_result(self.Xi, self.yi, self.space, self.rng, models=self.models)
        result.specs = self.specs
        return result
