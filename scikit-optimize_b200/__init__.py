"""B200-native GP-posterior + acquisition engine with the scikit-optimize API for that path.

    from skopt_b200 import gp_minimize, Optimizer
    from skopt_b200.acquisition import gaussian_ei, gaussian_pi, gaussian_lcb
    from skopt_b200.learning import GaussianProcessRegressor

(the package directory is `scikit-optimize_b200/`; `import skopt_b200` resolves to it through the shim at the
repository root).
"""
from . import acquisition, benchmarks, plots, space, utils  # noqa: F401
from .engine import (DeviceState, HostState, host_state_from_estimator, launch_count,  # noqa: F401
                     set_default_precision)
from .learning.gaussian_process.gpr import set_default_device, set_fit_device  # noqa: F401
from .optimizer import Optimizer, base_minimize, dummy_minimize, gp_minimize  # noqa: F401
from .space import Space  # noqa: F401
from .utils import dump, expected_minimum, expected_minimum_random_sampling, load  # noqa: F401

__version__ = "0.1.0"
