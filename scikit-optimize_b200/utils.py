"""Host glue used around the hot path; mirrors the relevant pieces of skopt/utils.py: create_result :29-72,
eval_callbacks :75-102, has_gradients :301-330, cook_estimator :333-405 (GP branch only), normalize_dimensions
:569-607, the list-shape helpers used by Optimizer.tell (is_listlike / is_2Dlistlike / check_x_in_space) and the
helpers callers wrap their objective and search space with (dimensions_aslist :453, point_asdict :489, point_aslist
:529, check_list_types :610, check_dimension_names :641, use_named_args :663)."""
from collections import OrderedDict
from functools import wraps

import numpy as np
from scipy.optimize import OptimizeResult
from sklearn.base import is_regressor

from .learning import GaussianProcessRegressor
from .learning.gaussian_process.kernels import ConstantKernel, HammingKernel, Matern
from .space import Dimension, Space, normalize_dimensions  # noqa: F401


def is_listlike(x):
    return isinstance(x, (list, tuple))


def is_2Dlistlike(x):
    return np.all([is_listlike(xi) for xi in x])


def check_x_in_space(x, space):
    if is_2Dlistlike(x):
        if not np.all([p in space for p in x]):
            raise ValueError("Not all points are within the bounds of the space.")
        if any([len(p) != len(space.dimensions) for p in x]):
            raise ValueError("Not all points have the same dimensions as the space.")
    elif is_listlike(x):
        if x not in space:
            raise ValueError("Point (%s) is not within the bounds of the space (%s)." % (x, space.bounds))
        if len(x) != len(space.dimensions):
            raise ValueError("Dimensions of point (%s) and space (%s) do not match" % (x, space.bounds))


def create_result(Xi, yi, space=None, rng=None, specs=None, models=None):
    """scipy OptimizeResult with the fields gp_minimize documents (x, fun, func_vals, x_iters, models, space, ...)."""
    res = OptimizeResult()
    yi = np.asarray(yi)
    if np.ndim(yi) == 2:
        res.log_time = np.ravel(yi[:, 1])
        yi = np.ravel(yi[:, 0])
    best = np.argmin(yi)
    res.x = Xi[best]
    res.fun = yi[best]
    res.func_vals = yi
    res.x_iters = Xi
    res.models = models
    res.space = space
    res.random_state = rng
    res.specs = specs
    return res


def eval_callbacks(callbacks, result):
    """True when any callback asks to stop; every callback runs, `None` answers do not count (utils.py:75-102)."""
    stop = False
    for callback in callbacks or ():
        decision = callback(result)
        if decision is not None:
            stop = stop or decision
    return stop


def dimensions_aslist(search_space):
    """Dimensions of a {name: Dimension} search space in the order of the sorted names."""
    return [search_space[name] for name in sorted(search_space)]


def point_asdict(search_space, point_as_list):
    """A point given in sorted-name order -> OrderedDict {name: value}."""
    return OrderedDict(zip(sorted(search_space), point_as_list))


def point_aslist(search_space, point_as_dict):
    """{name: value} -> the values in sorted-name order."""
    return [point_as_dict[name] for name in sorted(search_space)]


def check_list_types(x, types):
    wrong = [item for item in x if not isinstance(item, types)]
    if wrong:
        raise ValueError("All elements in list must be instances of {}, but found: {}".format(types, wrong))


def check_dimension_names(dimensions):
    unnamed = [dim for dim in dimensions if dim.name is None]
    if unnamed:
        raise ValueError("All dimensions must have names, but found: {}".format(unnamed))


def use_named_args(dimensions):
    """Decorator: an objective written with named arguments, `f(lr=..., depth=...)`, becomes the `f(x)` the minimisers
    call, `x` being the point as a list in the order of `dimensions` (all of which must be named Dimensions)."""
    def decorator(func):
        check_list_types(dimensions, Dimension)
        check_dimension_names(dimensions)

        @wraps(func)
        def wrapper(x):
            if len(x) != len(dimensions):
                raise ValueError("Mismatch in number of search-space dimensions. len(dimensions)=={} and len(x)=={}"
                                 .format(len(dimensions), len(x)))
            return func(**{dim.name: value for dim, value in zip(dimensions, x)})
        return wrapper
    return decorator


def _tree_families():
    from sklearn.ensemble import BaseEnsemble
    from sklearn.ensemble._gb import BaseGradientBoosting
    from sklearn.tree import BaseDecisionTree
    return (BaseEnsemble, BaseGradientBoosting, BaseDecisionTree)


_TREE_FAMILIES = _tree_families()


def has_gradients(estimator):
    """utils.py:301-330: True when predict() can return input gradients - every GP this package builds (the GPU
    computes them) except one whose kernel contains a HammingKernel (categorical inputs have no gradient)."""
    if estimator is None or not hasattr(estimator, "predict"):
        return False
    if isinstance(estimator, _TREE_FAMILIES) or type(estimator).__name__ == "GradientBoostingQuantileRegressor":
        return False
    if hasattr(estimator, "kernel"):
        params = estimator.get_params()
        if isinstance(estimator.kernel, HammingKernel) or any(isinstance(v, HammingKernel) for v in params.values()):
            return False
    return True


def cook_estimator(base_estimator, space=None, **kwargs):
    """Default surrogate of gp_minimize: ConstantKernel(1, (0.01, 1000)) * Matern(nu=2.5, ARD length scales in
    (0.01, 100)), normalize_y, noise="gaussian", 2 optimiser restarts (utils.py:366-387).  Only the GP family
    An all-categorical space gets HammingKernel(ones) in place of the Matern factor (utils.py:377-378).  Only the GP
    family exists here; tree / GBRT surrogates are outside the GPU hot path."""
    if isinstance(base_estimator, str):
        name = base_estimator.upper()
        if name not in ("GP", "RF", "ET", "GBRT", "DUMMY"):
            raise ValueError("Valid strings for the base_estimator parameter  are: 'RF', 'ET', 'GP', 'GBRT' or 'DUMMY' "
                             "not %s." % base_estimator)
        if name in ("RF", "ET", "GBRT"):
            raise ValueError("base_estimator=%r: tree / GBRT surrogates are not built by this package (the GPU path "
                             "is the GP family); pass a fitted-able regressor instance instead, it is scored through "
                             "its own predict()." % base_estimator)
        if name == "DUMMY":
            return None
        if space is None:
            raise ValueError("Expected a Space instance, not None.")
        space = normalize_dimensions(Space(space).dimensions)
        n_dims = space.transformed_n_dims
        if space.is_categorical:            # only special if *all* dimensions are categorical
            other = HammingKernel(length_scale=np.ones(n_dims))
        else:
            other = Matern(length_scale=np.ones(n_dims), length_scale_bounds=[(0.01, 100)] * n_dims, nu=2.5)
        kernel = ConstantKernel(1.0, (0.01, 1000.0)) * other
        base_estimator = GaussianProcessRegressor(kernel=kernel, normalize_y=True, noise="gaussian",
                                                  n_restarts_optimizer=2)
    elif not is_regressor(base_estimator):
        raise ValueError("base_estimator has to be a regressor.")
    if "n_jobs" in kwargs and not hasattr(base_estimator, "n_jobs"):
        del kwargs["n_jobs"]
    base_estimator.set_params(**kwargs)
    return base_estimator


def dump(res, filename, store_objective=True, **kwargs):
    """utils.py:103-143: joblib-persist an OptimizeResult; store_objective=False drops specs['args']['func'] from a
    deep copy first.  Fitted estimators pickle without their GPU handle (gpr.__getstate__); the device state is
    rebuilt on the first predict after load()."""
    from copy import deepcopy

    from joblib import dump as dump_
    if not store_objective and "func" in res.specs["args"]:
        res = deepcopy(res)
        del res.specs["args"]["func"]
    dump_(res, filename, **kwargs)


def load(filename, **kwargs):
    """utils.py:146-170."""
    from joblib import load as load_
    return load_(filename, **kwargs)


def expected_minimum_random_sampling(res, n_random_starts=100000, random_state=None):
    """utils.py:259-296: the minimum of the last model's mean over `n_random_starts` random points of the space -
    one mean-only pass of the hot path (skb_predict_mean_host)."""
    random_samples = res.space.rvs(n_random_starts, random_state=random_state)
    y_random = res.models[-1].predict(res.space.transform(random_samples))
    best = np.argmin(y_random)
    return random_samples[best], y_random[best]


def _transform_slope(dim, x):
    """d transform(x) / dx of one numeric dimension; 0 for Integer dimensions, whose transform rounds (the
    reference's finite-difference search never moves them either)."""
    from .space import Real
    if not isinstance(dim, Real):
        return 0.0
    slope = 1.0 if dim.prior == "uniform" else 1.0 / (x * np.log(dim.base))
    if dim.transform_ == "normalize":
        slope /= dim._whigh - dim._wlow
    return slope


def expected_minimum(res, n_random_starts=20, random_state=None):
    """utils.py:203-256: minimise the last model's posterior mean from res.x and `n_random_starts` random points.

    The reference runs one scipy `minimize` per start with finite-difference gradients of a one-row predict
    (about 2*(d+1) predicts per iteration); here all starts advance in lock step and every round is one batched
    skb_score_grad_host call returning the mean and its exact input gradient (chain rule through the space
    transform).  Same search space, starts and optimiser (L-BFGS-B with the space bounds); the minimiser agrees
    with the reference's to the optimiser's tolerance, not bit for bit."""
    from .optimizer.lbfgs import minimize_batch
    if res.space.is_partly_categorical:
        return expected_minimum_random_sampling(res, n_random_starts=100000, random_state=random_state)
    space, reg = res.space, res.models[-1]
    xs = [res.x]
    if n_random_starts > 0:
        xs.extend(space.rvs(n_random_starts, random_state=random_state))
    if not hasattr(reg, "device_state"):
        # a regressor of another family has no resident state and no gradients to batch: minimise its own predict() from
        # every start, the way the reference does for all models (scipy.optimize.minimize, finite differences)
        from scipy.optimize import minimize as sp_minimize
        best_x, best_fun = None, np.inf
        for x0 in xs:
            r = sp_minimize(lambda x: reg.predict(space.transform([list(x)]))[0], x0=x0, bounds=space.bounds)
            if r.fun < best_fun:
                best_x, best_fun = r.x, r.fun
        return [v for v in best_x], best_fun
    dev = reg.device_state()

    def evaluate(tasks, X):
        Xt = space.transform([list(row) for row in X])
        out = dev.score_grad(Xt, acq_funcs=())
        slopes = np.array([[_transform_slope(dim, v) for dim, v in zip(space.dimensions, row)] for row in X])
        return out["mean"], out["mean_grad"] * slopes

    best_x, best_fun = None, np.inf
    for x, f, _ in minimize_batch([np.asarray(x, dtype=float) for x in xs], evaluate, space.bounds, maxiter=15000):
        if f < best_fun:
            best_x, best_fun = x, f
    return [v for v in best_x], best_fun
