from . import kernels  # noqa: F401
from .gpr import GaussianProcessRegressor  # noqa: F401
