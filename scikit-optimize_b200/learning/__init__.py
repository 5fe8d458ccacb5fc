"""Surrogate models of the hot path (mirror of skopt.learning for the GP family)."""
from .gaussian_process import GaussianProcessRegressor  # noqa: F401
