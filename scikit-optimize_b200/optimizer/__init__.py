from .base import base_minimize  # noqa: F401
from .dummy import dummy_minimize  # noqa: F401
from .gp import gp_minimize  # noqa: F401
from .optimizer import Optimizer  # noqa: F401
