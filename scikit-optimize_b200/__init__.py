"""B200-native GP-posterior + acquisition engine with the scikit-optimize API for that path."""
from .engine import DeviceState, HostState, host_state_from_estimator, launch_count  # noqa: F401

__version__ = "0.1.0"
