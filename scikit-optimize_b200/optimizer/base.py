"""Ask/tell loop shared by the minimisers (mirror of skopt/optimizer/base.py:22-305)."""
import numbers
import time
import warnings
from collections.abc import Iterable

from ..utils import eval_callbacks
from .optimizer import Optimizer


def _as_callbacks(callback):
    if callback is None:
        return []
    if callable(callback):
        return [callback]
    if isinstance(callback, list) and all(callable(c) for c in callback):
        return list(callback)
    raise ValueError("callback should be either a callable or a list of callables.")


class _Progress(object):
    """verbose=True: the lines skopt.callbacks.VerboseCallback prints (callbacks.py:38-113) - an announcement when an
    iteration starts (provided point / random point / model-based search), and at its end the time taken, the value
    obtained and the current minimum.  The callback runs after every tell(), so "start" of iteration i + 1 is printed
    right after the "end" of iteration i."""

    def __init__(self, n_provided, n_random, n_total):
        self.n_provided, self.n_random, self.n_total = n_provided, n_random, n_total
        self.iteration = 1
        self.started = time.time()
        self._announce(True)

    def _announce(self, start):
        i = self.iteration
        status, doing = ("started", "Evaluating function") if start else ("ended", "Evaluation done")
        if i <= self.n_provided:
            print("Iteration No: %d %s. %s at provided point." % (i, status, doing))
        elif i <= self.n_provided + self.n_random:
            print("Iteration No: %d %s. %s at random point." % (i, status, doing))
        else:
            print("Iteration No: %d %s. %s" % (i, status, "Searching for the next optimal point." if start
                                               else "Search finished for the next optimal point."))

    def __call__(self, res):
        elapsed = time.time() - self.started
        self._announce(False)
        print("Time taken: %0.4f" % elapsed)
        print("Function value obtained: %0.4f" % res.func_vals[-1])
        print("Current minimum: %0.4f" % res.fun)
        self.iteration += 1
        if self.iteration <= self.n_total:
            self._announce(True)
            self.started = time.time()


def base_minimize(func, dimensions, base_estimator, n_calls=100, n_random_starts=None, n_initial_points=10,
                  initial_point_generator="random", acq_func="EI", acq_optimizer="lbfgs", x0=None, y0=None,
                  random_state=None, verbose=False, callback=None, n_points=10000, n_restarts_optimizer=5,
                  xi=0.01, kappa=1.96, n_jobs=1, model_queue_size=None):
    specs = {"args": locals(), "function": "base_minimize"}
    acq_optimizer_kwargs = {"n_points": n_points, "n_restarts_optimizer": n_restarts_optimizer, "n_jobs": n_jobs}
    acq_func_kwargs = {"xi": xi, "kappa": kappa}

    if x0 is None:
        x0 = []
    elif not isinstance(x0[0], (list, tuple)):
        x0 = [x0]
    if not isinstance(x0, list):
        raise ValueError("`x0` should be a list, but got %s" % type(x0))
    if n_random_starts is not None:
        warnings.warn("n_random_starts will be removed in favour of n_initial_points. It overwrites n_initial_points.",
                      DeprecationWarning)
        n_initial_points = n_random_starts
    if n_initial_points <= 0 and not x0:
        raise ValueError("Either set `n_initial_points` > 0, or provide `x0`")
    if isinstance(y0, Iterable):
        y0 = list(y0)
    elif isinstance(y0, numbers.Number):
        y0 = [y0]
    required_calls = n_initial_points + (len(x0) if not y0 else 0)
    if n_calls < required_calls:
        raise ValueError("Expected `n_calls` >= %d, got %d" % (required_calls, n_calls))
    n_initial_points = n_initial_points + len(x0)

    optimizer = Optimizer(dimensions, base_estimator, n_initial_points=n_initial_points,
                          initial_point_generator=initial_point_generator, n_jobs=n_jobs, acq_func=acq_func,
                          acq_optimizer=acq_optimizer, random_state=random_state, model_queue_size=model_queue_size,
                          acq_optimizer_kwargs=acq_optimizer_kwargs, acq_func_kwargs=acq_func_kwargs)
    if not all(isinstance(p, Iterable) for p in x0):
        raise ValueError("every point of `x0` must be iterable")
    if not all(len(p) == optimizer.space.n_dims for p in x0):
        raise RuntimeError("Optimization space (%s) and initial points in x0 use inconsistent dimensions."
                           % optimizer.space)
    callbacks = _as_callbacks(callback)
    if verbose:
        callbacks.append(_Progress(n_provided=len(x0) if not y0 else 0, n_random=n_initial_points, n_total=n_calls))

    result = None
    if x0 and y0 is None:
        y0 = list(map(func, x0))
        n_calls -= len(y0)
    if x0:
        if not (isinstance(y0, Iterable) or isinstance(y0, numbers.Number)):
            raise ValueError("`y0` should be an iterable or a scalar, got %s" % type(y0))
        if len(x0) != len(y0):
            raise ValueError("`x0` and `y0` should have the same length")
        result = optimizer.tell(x0, y0)
        result.specs = specs
        if eval_callbacks(callbacks, result):
            return result

    for _ in range(n_calls):
        next_x = optimizer.ask()
        next_y = func(next_x)
        result = optimizer.tell(next_x, next_y)
        result.specs = specs
        if eval_callbacks(callbacks, result):
            break
    return result
