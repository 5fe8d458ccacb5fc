"""dummy_minimize: random search, the model-free baseline every BO run is compared with (mirror of
skopt/optimizer/dummy.py:6-118).  It is the ask / tell loop with no surrogate: every point is a draw from the space, so
nothing here touches the GPU - it exists so that code written against the reference's minimisers keeps its baseline."""
from .base import base_minimize


def dummy_minimize(func, dimensions, n_calls=100, initial_point_generator="random", x0=None, y0=None, random_state=None,
                   verbose=False, callback=None, model_queue_size=None, init_point_gen_kwargs=None):
    # every call is an "initial" (random) point; points given without values still have to be evaluated
    n_initial_points = n_calls - len(x0) if (x0 is not None and y0 is None) else n_calls
    return base_minimize(func, dimensions, base_estimator="dummy", acq_optimizer="sampling", n_calls=n_calls,
                         n_initial_points=n_initial_points, initial_point_generator=initial_point_generator, x0=x0,
                         y0=y0, random_state=random_state, verbose=verbose, callback=callback,
                         model_queue_size=model_queue_size)
