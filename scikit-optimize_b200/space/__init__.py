"""Search space: dimensions, their warpings and candidate generation (mirror of skopt/space/__init__.py)."""
from .space import *  # noqa: F401,F403
from .space import (Categorical, Dimension, Integer, Real, Space, check_dimension, device_dim_descs,  # noqa: F401
                    normalize_dimensions)
from .transformers import (CategoricalEncoder, Identity, LabelEncoder, LogN, Normalize, Pipeline,  # noqa: F401
                           StringEncoder, Transformer)
