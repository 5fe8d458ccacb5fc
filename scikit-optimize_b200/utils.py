"""Host glue used around the hot path; mirrors the relevant pieces of skopt/utils.py: create_result :29-72,
has_gradients :301-330, cook_estimator :333-405 (GP branch only), normalize_dimensions :569-607, and the list-shape
helpers used by Optimizer.tell (is_listlike / is_2Dlistlike / check_x_in_space)."""
import numpy as np
from scipy.optimize import OptimizeResult
from sklearn.base import is_regressor

from .learning import GaussianProcessRegressor
from .learning.gaussian_process.kernels import ConstantKernel, Matern
from .space import Space, normalize_dimensions  # noqa: F401


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


def has_gradients(estimator):
    """True when predict() can return input gradients: every GP this package builds (the GPU computes them)."""
    return estimator is not None and hasattr(estimator, "predict")


def cook_estimator(base_estimator, space=None, **kwargs):
    """Default surrogate of gp_minimize: ConstantKernel(1, (0.01, 1000)) * Matern(nu=2.5, ARD length scales in
    (0.01, 100)), normalize_y, noise="gaussian", 2 optimiser restarts (utils.py:366-387).  Only the GP family
    exists here; tree / GBRT surrogates and the all-categorical HammingKernel GP are outside the GPU hot path."""
    if isinstance(base_estimator, str):
        name = base_estimator.upper()
        if name not in ("GP", "DUMMY"):
            raise ValueError("Valid strings for the base_estimator parameter of the GPU path are 'GP' or 'DUMMY', "
                             "not %s." % base_estimator)
        if name == "DUMMY":
            return None
        if space is None:
            raise ValueError("Expected a Space instance, not None.")
        space = normalize_dimensions(Space(space).dimensions)
        if space.is_categorical:
            raise NotImplementedError("all-categorical spaces use the HammingKernel in scikit-optimize, which is "
                                      "outside the GPU hot-path grammar (RBF / Matern)")
        n_dims = space.transformed_n_dims
        kernel = ConstantKernel(1.0, (0.01, 1000.0)) * Matern(length_scale=np.ones(n_dims),
                                                              length_scale_bounds=[(0.01, 100)] * n_dims, nu=2.5)
        base_estimator = GaussianProcessRegressor(kernel=kernel, normalize_y=True, noise="gaussian",
                                                  n_restarts_optimizer=2)
    elif not is_regressor(base_estimator):
        raise ValueError("base_estimator has to be a regressor.")
    if "n_jobs" in kwargs and not hasattr(base_estimator, "n_jobs"):
        del kwargs["n_jobs"]
    base_estimator.set_params(**kwargs)
    return base_estimator
