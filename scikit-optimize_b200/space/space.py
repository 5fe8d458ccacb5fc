"""Search-space plumbing that FEEDS the hot path: dimensions, the "normalize" warping used by every GP run, and
candidate generation.

Host-side mirror of skopt/space/space.py (Real :212, Integer :421, Categorical :578, Space :745, check_dimension
:31-136) and of the Normalize / LogN / LabelEncoder transformers (skopt/space/transformers.py:221-276), written
array-first: `Space.rvs_transformed(n, rng)` produces the (n x d) candidate matrix that `Optimizer._tell` scores
(optimizer.py:552-553: `space.transform(space.rvs(n_points, rng))`) with vectorised NumPy instead of Python lists,
consuming the RandomState in the same order (one `uniform(0, 1, n)` draw per dimension, column by column,
space.py:899-900) and applying the same floating-point operations (u*(1+eps) -> *(hi-lo)+lo -> clip -> (x-lo)/(hi-lo)),
so the matrix is bit-identical to the reference's (pinned in tests/test_space.py).
"""
import numbers

import numpy as np
from sklearn.utils import check_random_state

from .transformers import (CategoricalEncoder, Identity, LabelEncoder, LogN, Normalize, Pipeline,  # noqa: F401
                           StringEncoder)

_INCLUSIVE_ONE = np.nextafter(1.0, 2.0)       # scipy uniform(0, nextafter(1, 2)): samples include the upper edge
_EPS = 1e-8


def _unit_inclusive(rng, n):
    """scipy.stats.uniform(loc=0, scale=nextafter(1,2)).rvs(n, random_state=rng): U * scale + loc (space.py:205-209).
    U = rng.uniform(0, 1) is 0.0 + 1.0 * random_sample(), i.e. random_sample() bit for bit (same stream consumption), and
    adding loc = 0.0 to a non-negative product changes nothing: two array passes less per column."""
    return rng.random_sample(n) * _INCLUSIVE_ONE


def _affine_inclusive(rng, n, loc, width):
    return rng.uniform(0.0, 1.0, size=n) * np.nextafter(width, width + 1.0) + loc


class _Ellipsis(object):
    def __repr__(self):
        return "..."


class Dimension(object):
    """Base class: a dimension knows how to sample, warp and un-warp 1-D arrays of values."""
    prior = None
    transform_ = "identity"
    _name = None

    @property
    def name(self):
        return self._name

    @name.setter
    def name(self, value):
        if not (isinstance(value, str) or value is None):
            raise ValueError("Dimension's name must be either string or None.")
        self._name = value

    @property
    def size(self):
        return 1

    @property
    def transformed_size(self):
        return 1

    def rvs(self, n_samples=1, random_state=None):
        rng = check_random_state(random_state)
        return self.inverse_transform(self._draw(rng, n_samples))

    def _sample_column(self, rng, n):
        """`n` samples in the original space as one column for Space.rvs (an array, or a list of categories)."""
        return self._from_unit(self._draw(rng, n))

    # subclasses: _draw (sample in the warped space), _from_unit (warped -> original on whole arrays: the form
    # Space.rvs_transformed / Space.inverse_transform use), inverse_transform (the same values with the reference's
    # return types), transform


class Real(Dimension):
    def __init__(self, low, high, prior="uniform", base=10, transform=None, name=None, dtype=float):
        if high <= low:
            raise ValueError("the lower bound {} has to be less than the upper bound {}".format(low, high))
        if prior not in ("uniform", "log-uniform"):
            raise ValueError("prior should be 'uniform' or 'log-uniform' got {}".format(prior))
        if isinstance(dtype, str) and dtype not in ("float", "float16", "float32", "float64"):
            raise ValueError("dtype must be 'float', 'float16', 'float32' or 'float64' got {}".format(dtype))
        if isinstance(dtype, type) and not np.issubdtype(dtype, np.floating):
            raise ValueError("dtype must be a np.floating subtype; got {}".format(dtype))
        self.low, self.high, self.prior, self.base = low, high, prior, base
        self.log_base = np.log10(base)
        self.name, self.dtype = name, dtype
        self.set_transformer(transform or "identity")

    def set_transformer(self, transform="identity"):
        if transform not in ("normalize", "identity"):
            raise ValueError("transform should be 'normalize' or 'identity' got {}".format(transform))
        self.transform_ = transform
        if self.prior == "uniform":
            self._wlow, self._whigh = float(self.low), float(self.high)
        else:
            self._wlow = np.log10(self.low) / self.log_base
            self._whigh = np.log10(self.high) / self.log_base

    def _warp(self, x):
        return x if self.prior == "uniform" else np.log10(x) / self.log_base

    def _unwarp(self, w):
        return w if self.prior == "uniform" else np.power(self.base, w)

    def _draw(self, rng, n):
        if self.transform_ == "normalize":
            return _unit_inclusive(rng, n)
        return _affine_inclusive(rng, n, self._wlow, self._whigh - self._wlow)

    def transform(self, X):
        X = np.asarray(X, dtype=float)
        w = self._warp(X)
        if self.transform_ != "normalize":
            return w
        if np.any(w > self._whigh + _EPS):
            raise ValueError("All values should be less than %f" % self._whigh)
        if np.any(w < self._wlow - _EPS):
            raise ValueError("All values should be greater than %f" % self._wlow)
        return (w - self._wlow) / (self._whigh - self._wlow)

    def _from_unit(self, Xt):
        Xt = np.asarray(Xt, dtype=float)
        if self.transform_ == "normalize":
            if np.any(Xt > 1.0 + _EPS):
                raise ValueError("All values should be less than 1.0")
            if np.any(Xt < 0.0 - _EPS):
                raise ValueError("All values should be greater than 0.0")
            Xt = Xt * (self._whigh - self._wlow) + self._wlow
        out = np.clip(self._unwarp(Xt), self.low, self.high)
        return out if self.dtype in (float, "float") else out.astype(self.dtype)

    def _sample_transformed_column(self, rng, n):
        """transform(_sample_column(rng, n)) for the case the scoring block draws ten thousand times per tell - uniform
        prior, "normalize", default dtype - with the same draws and the same arithmetic, minus the four range checks
        (`_from_unit` and `transform` validate caller-supplied values; a fresh draw lies in [0, nextafter(1, 2)] and its
        clipped image in [low, high], so they cannot fire): 4 array passes instead of 12.  None = not this case."""
        if self.prior != "uniform" or self.transform_ != "normalize" or self.dtype not in (float, "float"):
            return None
        width = self._whigh - self._wlow
        x = np.clip(_unit_inclusive(rng, n) * width + self._wlow, self.low, self.high)
        return (x - self._wlow) / width

    def inverse_transform(self, Xt):
        """Warped -> original; Python floats (a list, or one float for a scalar) for the default dtype, an array of
        `dtype` otherwise (space.py:337-351)."""
        out = self._from_unit(Xt)
        return out.tolist() if self.dtype in (float, "float") else out

    @property
    def transformer(self):
        """The reference's transformer object for this dimension (space.py:299-323), built on demand."""
        warp = Identity() if self.prior == "uniform" else LogN(self.base)
        if self.transform_ != "normalize":
            return warp
        return Pipeline([warp, Normalize(self._wlow, self._whigh)])

    @property
    def bounds(self):
        return (self.low, self.high)

    @property
    def is_constant(self):
        return self.low == self.high

    @property
    def transformed_bounds(self):
        if self.transform_ == "normalize":
            return 0.0, 1.0
        if self.prior == "uniform":
            return self.low, self.high
        return np.log10(self.low), np.log10(self.high)

    def __contains__(self, point):
        if isinstance(point, list):
            point = np.array(point)
        return self.low <= point <= self.high

    def distance(self, a, b):
        if not (a in self and b in self):
            raise RuntimeError("Can only compute distance for values within the space, not %s and %s." % (a, b))
        return abs(a - b)

    def __eq__(self, other):
        return (type(self) is type(other) and np.allclose([self.low], [other.low]) and
                np.allclose([self.high], [other.high]) and self.prior == other.prior and
                self.transform_ == other.transform_)

    __hash__ = None

    def __repr__(self):
        return "Real(low={}, high={}, prior='{}', transform='{}')".format(self.low, self.high, self.prior, self.transform_)


class Integer(Dimension):
    _INT_TYPES = (int, np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64)

    def __init__(self, low, high, prior="uniform", base=10, transform=None, name=None, dtype=np.int64):
        if high <= low:
            raise ValueError("the lower bound {} has to be less than the upper bound {}".format(low, high))
        if prior not in ("uniform", "log-uniform"):
            raise ValueError("prior should be 'uniform' or 'log-uniform' got {}".format(prior))
        if isinstance(dtype, str) and dtype not in ("int", "int8", "int16", "int32", "int64", "uint8", "uint16",
                                                    "uint32", "uint64"):
            raise ValueError("dtype must be an integer type name, but got {}".format(dtype))
        if isinstance(dtype, type) and dtype not in self._INT_TYPES:
            raise ValueError("dtype must be an integer type, but got {}".format(dtype))
        self.low, self.high, self.prior, self.base = low, high, prior, base
        self.log_base = np.log10(base)
        self.name, self.dtype = name, dtype
        self.set_transformer(transform or "identity")

    def set_transformer(self, transform="identity"):
        if transform not in ("normalize", "identity"):
            raise ValueError("transform should be 'normalize' or 'identity' got {}".format(transform))
        self.transform_ = transform
        if self.prior == "uniform":
            self._wlow, self._whigh = float(self.low), float(self.high)
        else:
            self._wlow = np.log10(self.low) / self.log_base
            self._whigh = np.log10(self.high) / self.log_base

    def _draw(self, rng, n):
        if self.transform_ == "normalize":
            return _unit_inclusive(rng, n)
        if self.prior == "uniform":
            return rng.randint(self.low, self.high + 1, size=n)
        return _affine_inclusive(rng, n, self._wlow, self._whigh - self._wlow)

    def transform(self, X):
        X = np.asarray(X)
        if self.prior == "uniform":
            if self.transform_ != "normalize":
                return X
            if np.any(np.round(X) > self.high):
                raise ValueError("All integer values should be less than %f" % self.high)
            if np.any(np.round(X) < self.low):
                raise ValueError("All integer values should be greater than %f" % self.low)
            return (np.round(X).astype(int) - self._wlow) / (self._whigh - self._wlow)
        w = np.log10(np.asarray(X, dtype=float)) / self.log_base
        if self.transform_ != "normalize":
            return w
        if np.any(w > self._whigh + _EPS):
            raise ValueError("All values should be less than %f" % self._whigh)
        if np.any(w < self._wlow - _EPS):
            raise ValueError("All values should be greater than %f" % self._wlow)
        return (w - self._wlow) / (self._whigh - self._wlow)

    def _from_unit(self, Xt):
        Xt = np.asarray(Xt, dtype=float)
        if self.transform_ == "normalize":
            if np.any(Xt > 1.0 + _EPS):
                raise ValueError("All values should be less than 1.0")
            if np.any(Xt < 0.0 - _EPS):
                raise ValueError("All values should be greater than 0.0")
            Xt = Xt * (self._whigh - self._wlow) + self._wlow
            if self.prior == "uniform":
                Xt = np.round(Xt).astype(int)
        if self.prior != "uniform":
            Xt = np.power(self.base, Xt)
        out = np.clip(Xt, self.low, self.high)
        return np.round(out).astype(int if self.dtype in (int, "int") else self.dtype)

    def inverse_transform(self, Xt):
        """Warped -> original; Python ints (a list, or one int for a scalar) when dtype is `int`, an array of `dtype`
        otherwise (space.py:522-538)."""
        out = self._from_unit(Xt)
        return out.tolist() if self.dtype in (int, "int") else out

    @property
    def transformer(self):
        """The reference's transformer object for this dimension (space.py:489-511), built on demand."""
        if self.transform_ != "normalize":
            return Identity() if self.prior == "uniform" else LogN(self.base)
        if self.prior == "uniform":
            return Pipeline([Identity(), Normalize(self.low, self.high, is_int=True)])
        return Pipeline([LogN(self.base), Normalize(self._wlow, self._whigh)])

    @property
    def bounds(self):
        return (self.low, self.high)

    @property
    def is_constant(self):
        return self.low == self.high

    @property
    def transformed_bounds(self):
        if self.transform_ == "normalize":
            return 0.0, 1.0
        return (self.low, self.high)

    def __contains__(self, point):
        if isinstance(point, list):
            point = np.array(point)
        return self.low <= point <= self.high

    def distance(self, a, b):
        if not (a in self and b in self):
            raise RuntimeError("Can only compute distance for values within the space, not %s and %s." % (a, b))
        return abs(a - b)

    def __eq__(self, other):
        return (type(self) is type(other) and np.allclose([self.low], [other.low]) and
                np.allclose([self.high], [other.high]))

    __hash__ = None

    def __repr__(self):
        return "Integer(low={}, high={}, prior='{}', transform='{}')".format(self.low, self.high, self.prior,
                                                                              self.transform_)


class Categorical(Dimension):
    """Categories are label-encoded; under "normalize" (every GP run) the label is scaled to [0, 1] in ONE column
    (space.py:646-649), so the Matern kernel sees it as an ordinary coordinate on a lattice."""

    def __init__(self, categories, prior=None, transform=None, name=None):
        self.categories = tuple(categories)
        self.name = name
        self.prior = prior
        self.prior_ = np.tile(1.0 / len(self.categories), len(self.categories)) if prior is None else prior
        cats = np.asarray(self.categories)
        if cats.dtype == object:
            ordered = list(self.categories)                      # LabelEncoder.fit: given order for object arrays
        else:                                                    # ... np.unique (sorted) order otherwise
            ordered = [self.categories[int(np.where(cats == u)[0][0])] for u in np.unique(cats)]
        self._order = ordered                                    # label -> category (np.unique order for plain types)
        self._label = {c: i for i, c in enumerate(ordered)}
        self._onehot_pos = {c: i for i, c in enumerate(self.categories)}
        self.set_transformer(transform or "onehot")

    def set_transformer(self, transform="onehot"):
        if transform not in ("identity", "onehot", "string", "normalize", "label"):
            raise ValueError("Expected transform to be 'identity', 'string', 'label' or 'onehot' got {}".format(transform))
        self.transform_ = transform

    @property
    def _n(self):
        return len(self.categories)

    def _draw(self, rng, n):
        if self.transform_ == "normalize":
            return _unit_inclusive(rng, n)
        return rng.choice(self._n, size=n, p=self.prior_)

    def rvs(self, n_samples=None, random_state=None):
        """Categories drawn from the prior.  Return types follow space.py:688-698: without `n_samples` one category
        (an array holding one under "normalize"), otherwise a list (an array under "normalize")."""
        rng = check_random_state(random_state)
        scalar = n_samples is None
        draws = self._draw(rng, 1 if scalar else n_samples)
        if self.transform_ == "normalize":
            return self.inverse_transform(draws)
        out = [self.categories[int(c)] for c in draws]
        return out[0] if scalar else out

    def _sample_column(self, rng, n):
        draws = self._draw(rng, n)
        if self.transform_ == "normalize":
            return self._from_unit(draws)
        return [self.categories[int(c)] for c in draws]

    def labels(self, X):
        return np.array([self._label[v] for v in X], dtype=float)

    def transform(self, X):
        if self.transform_ == "identity":
            return X
        if self.transform_ == "string":
            return [str(v) for v in X]
        if self.transform_ == "label":
            return self.labels(X).astype(int)
        if self.transform_ == "normalize":
            lab = self.labels(X)
            if self._n == 1:
                return lab * 0.0
            return (np.round(lab).astype(int) - 0.0) / (float(self._n - 1) - 0.0)
        # one-hot over the categories in the GIVEN order (CategoricalEncoder, transformers.py:113-117); two categories
        # give one 0/1 column and a single category a column of zeros, like the LabelBinarizer behind the reference
        pos = np.fromiter((self._onehot_pos[v] for v in X), dtype=np.int64)
        if self._n <= 2:
            return (pos if self._n == 2 else pos * 0).reshape(-1, 1)
        onehot = np.zeros((len(pos), self._n), dtype=np.int64)
        onehot[np.arange(len(pos)), pos] = 1
        return onehot

    def inverse_transform(self, Xt):
        """Warped -> categories, as an array like the reference's (space.py:679-687)."""
        return np.array(self._from_unit(Xt))

    @property
    def transformer(self):
        """The reference's transformer object for this dimension (space.py:636-652), built on demand."""
        if self.transform_ == "normalize":
            return Pipeline([LabelEncoder(list(self.categories)), Normalize(0, self._n - 1, is_int=True)])
        enc = {"onehot": CategoricalEncoder, "string": StringEncoder, "label": LabelEncoder,
               "identity": Identity}[self.transform_]()
        enc.fit(self.categories)
        return enc

    def _from_unit(self, Xt):
        if self.transform_ == "identity":
            return list(Xt)
        if self.transform_ == "string":
            typ = type(self.categories[0])
            return [typ(v) for v in Xt]
        Xt = np.asarray(Xt, dtype=float)
        if self.transform_ == "normalize":
            if np.any(Xt > 1.0 + _EPS):
                raise ValueError("All values should be less than 1.0")
            if np.any(Xt < 0.0 - _EPS):
                raise ValueError("All values should be greater than 0.0")
            lab = np.round(Xt * (float(self._n - 1) - 0.0) + 0.0).astype(int)
        elif self.transform_ == "label":
            lab = np.round(Xt).astype(int)
        else:
            Xt = Xt.reshape(len(Xt), -1)
            if self._n <= 2:
                pos = (Xt[:, 0] > 0.5).astype(int) if self._n == 2 else np.zeros(len(Xt), dtype=int)
            else:
                pos = np.argmax(Xt, axis=1)
            return [self.categories[int(i)] for i in pos]
        return [self._order[int(i)] for i in np.ravel(lab)]

    @property
    def transformed_size(self):
        if self.transform_ == "onehot":
            return self._n if self._n != 2 else 1
        return 1

    @property
    def bounds(self):
        return self.categories

    @property
    def is_constant(self):
        return self._n <= 1

    @property
    def transformed_bounds(self):
        if self.transformed_size == 1:
            return 0.0, 1.0
        return [(0.0, 1.0) for _ in range(self.transformed_size)]

    def __contains__(self, point):
        return point in self.categories

    def distance(self, a, b):
        if not (a in self and b in self):
            raise RuntimeError("Can only compute distance for values within the space, not {} and {}.".format(a, b))
        return 1 if a != b else 0

    def __eq__(self, other):
        return (type(self) is type(other) and self.categories == other.categories and
                np.allclose(self.prior_, other.prior_))

    __hash__ = None

    def __repr__(self):
        cats, prior = self.categories, self.prior
        if len(cats) > 7:
            cats = cats[:3] + (_Ellipsis(),) + cats[-3:]
        if prior is not None and len(prior) > 7:
            prior = list(prior[:3]) + [_Ellipsis()] + list(prior[-3:])
        return "Categorical(categories={}, prior={})".format(cats, prior)


def check_dimension(dimension, transform=None):
    """Turn a tuple / list specification into a Dimension (same dispatch rules as space.py:31-136)."""
    if isinstance(dimension, Dimension):
        return dimension
    if not isinstance(dimension, (list, tuple, np.ndarray)):
        raise ValueError("Dimension has to be a list or tuple.")
    if len(dimension) == 1:
        return Categorical(dimension, transform=transform)
    if len(dimension) == 2:
        if any(isinstance(d, (str, bool, np.bool_)) for d in dimension):
            return Categorical(dimension, transform=transform)
        if all(isinstance(d, numbers.Integral) for d in dimension):
            return Integer(*dimension, transform=transform)
        if any(isinstance(d, numbers.Real) for d in dimension):
            return Real(*dimension, transform=transform)
        raise ValueError("Invalid dimension {}. Read the documentation for supported types.".format(dimension))
    if len(dimension) == 3:
        if any(isinstance(d, int) for d in dimension[:2]) and dimension[2] in ("uniform", "log-uniform"):
            return Integer(*dimension, transform=transform)
        if any(isinstance(d, (float, int)) for d in dimension[:2]) and dimension[2] in ("uniform", "log-uniform"):
            return Real(*dimension, transform=transform)
        return Categorical(dimension, transform=transform)
    if len(dimension) == 4:
        if (any(isinstance(d, int) for d in dimension[:2]) and dimension[2] == "log-uniform" and
                isinstance(dimension[3], int)):
            return Integer(*dimension, transform=transform)
        if (any(isinstance(d, (float, int)) for d in dimension[:2]) and dimension[2] == "log-uniform" and
                isinstance(dimension[3], int)):
            return Real(*dimension, transform=transform)
    if len(dimension) > 3:
        return Categorical(dimension, transform=transform)
    raise ValueError("Invalid dimension {}. Read the documentation for supported types.".format(dimension))


def _rows(columns):
    """Columns (one sequence per dimension) -> list of per-point lists with native Python scalars."""
    cols = [c.tolist() if isinstance(c, np.ndarray) else list(c) for c in columns]
    return [list(row) for row in zip(*cols)]


class Space(object):
    def __init__(self, dimensions):
        if isinstance(dimensions, Space):
            dimensions = dimensions.dimensions
        self.dimensions = [check_dimension(dim) for dim in dimensions]

    def __eq__(self, other):
        return all(a == b for a, b in zip(self.dimensions, other.dimensions))

    def __iter__(self):
        return iter(self.dimensions)

    def __repr__(self):
        dims = self.dimensions
        if len(dims) > 31:
            dims = dims[:15] + [_Ellipsis()] + dims[-15:]
        return "Space([{}])".format(",\n       ".join(map(str, dims)))

    @property
    def n_dims(self):
        return len(self.dimensions)

    @property
    def transformed_n_dims(self):
        return sum(dim.transformed_size for dim in self.dimensions)

    @property
    def dimension_names(self):
        return [dim.name if dim.name is not None else "X_%d" % i for i, dim in enumerate(self.dimensions)]

    @property
    def is_real(self):
        return all(isinstance(dim, Real) for dim in self.dimensions)

    @property
    def is_categorical(self):
        return all(isinstance(dim, Categorical) for dim in self.dimensions)

    @property
    def is_partly_categorical(self):
        return any(isinstance(dim, Categorical) for dim in self.dimensions)

    @property
    def bounds(self):
        b = []
        for dim in self.dimensions:
            if dim.size == 1:
                b.append(dim.bounds)
            else:
                b.extend(dim.bounds)
        return b

    @property
    def transformed_bounds(self):
        b = []
        for dim in self.dimensions:
            if dim.transformed_size == 1:
                b.append(dim.transformed_bounds)
            else:
                b.extend(dim.transformed_bounds)
        return b

    def set_transformer(self, transform):
        for j, dim in enumerate(self.dimensions):
            dim.set_transformer(transform[j] if isinstance(transform, list) else transform)

    def set_transformer_by_type(self, transform, dim_type):
        for dim in self.dimensions:
            if isinstance(dim, dim_type):
                dim.set_transformer(transform)

    def get_transformer(self):
        return [dim.transform_ for dim in self.dimensions]

    @property
    def n_constant_dimensions(self):
        return sum(1 for dim in self.dimensions if dim.is_constant)

    def __getitem__(self, dimension_names):
        """`space['foo']` / `space[2]` -> (index, Dimension) or (None, None); a list of names -> a list of such pairs
        (space.py:1039-1089)."""
        def find(key):
            for index, dim in enumerate(self.dimensions):
                if key == dim.name or key == index:
                    return index, dim
            return None, None

        if isinstance(dimension_names, (str, int)):
            return find(dimension_names)
        if isinstance(dimension_names, (list, tuple)):
            return [find(key) for key in dimension_names]
        raise ValueError("Dimension name should be either string orlist of strings, but got {}.".format(
            type(dimension_names)))

    @classmethod
    def from_yaml(cls, yml_path, namespace=None):
        """Space from a YAML file: a list of one-key mappings {Real | Integer | Categorical: constructor arguments},
        either at the top level or under a namespace (the first one unless `namespace` names another); keys are
        case-insensitive and unknown dimension kinds are skipped (space.py:808-872)."""
        import yaml
        with open(yml_path, "rb") as f:
            config = yaml.safe_load(f)
        if isinstance(config, dict):
            options = next(iter(config.values())) if namespace is None else config[namespace]
        elif isinstance(config, list):
            options = config
        else:
            raise TypeError("YaML does not specify a list or dictionary")
        kinds = {"real": Real, "integer": Integer, "categorical": Categorical}
        dimensions = []
        for option in options:
            kind, arguments = next(iter(option.items()))
            if kind.lower() in kinds:
                dimensions.append(kinds[kind.lower()](**{k.lower(): v for k, v in arguments.items()}))
        return cls(dimensions=dimensions)

    def __contains__(self, point):
        return all(component in dim for component, dim in zip(point, self.dimensions))

    def distance(self, point_a, point_b):
        return sum(dim.distance(a, b) for a, b, dim in zip(point_a, point_b, self.dimensions))

    def min_distance(self, point, points):
        """min(self.distance(point, p) for p in points) (the duplicate check of Optimizer._ask, optimizer.py:350-353),
        one NumPy pass per dimension instead of one Python call per (point, dimension): the per-point sums accumulate
        in dimension order like the reference's sum(), so the result is the same float."""
        if len(points) == 0:
            raise ValueError("min_distance needs at least one point")
        if point not in self or len(point) != len(self.dimensions):
            return min(self.distance(point, p) for p in points)          # raises the reference's error
        acc = 0
        for j, dim in enumerate(self.dimensions):
            col = [p[j] for p in points]
            if isinstance(dim, Categorical):
                if not all(v in dim for v in col):
                    return min(self.distance(point, p) for p in points)
                term = np.fromiter((1 if v != point[j] else 0 for v in col), dtype=np.int64, count=len(col))
            else:
                arr = np.asarray(col)
                if arr.dtype == object or not np.all((arr >= dim.low) & (arr <= dim.high)):
                    return min(self.distance(point, p) for p in points)
                term = np.abs(point[j] - arr)
            acc = acc + term
        return acc.min()

    # ------------------------------------------------------------------ sampling and warping
    def rvs(self, n_samples=1, random_state=None):
        """Points in the ORIGINAL space as a list of lists (space.py:874-903): one draw per dimension, in order."""
        rng = check_random_state(random_state)
        return _rows([dim._sample_column(rng, n_samples) for dim in self.dimensions])

    def transform(self, X):
        """Original space -> warped space, (n, transformed_n_dims) float array (space.py:942-974)."""
        n = len(X)
        cols = []
        for j, dim in enumerate(self.dimensions):
            cols.append(_as_column(dim.transform([X[i][j] for i in range(n)]), n))
        return np.hstack(cols) if cols else np.empty((n, 0))

    def inverse_transform(self, Xt):
        """Warped space -> list of points in the original space (space.py:976-1007)."""
        Xt = np.asarray(Xt)
        cols, start = [], 0
        for dim in self.dimensions:
            width = dim.transformed_size
            block = Xt[:, start] if width == 1 else Xt[:, start:start + width]
            cols.append(dim._from_unit(block))
            start += width
        return _rows(cols)

    def rvs_transformed(self, n_samples, random_state=None):
        """`transform(rvs(n_samples, rng))` as ONE vectorised pass: the candidate matrix the GPU scores.

        Same RandomState consumption and the same floating-point operations as the reference's list-based path
        (sample -> inverse_transform -> clip -> transform), so the result is bit-identical to it."""
        rng = check_random_state(random_state)
        cols = []
        for dim in self.dimensions:
            fast = getattr(dim, "_sample_transformed_column", None)
            col = fast(rng, n_samples) if fast is not None else None
            if col is None:
                col = dim.transform(dim._sample_column(rng, n_samples))
            cols.append(_as_column(col, n_samples))
        return np.ascontiguousarray(np.hstack(cols))


def _as_column(col, n):
    """One dimension's transformed values as an (n, width) block: float for every numeric transform; a Categorical with
    transform "identity" / "string" keeps its labels (NumPy then makes the stacked matrix a string / object array, as
    the reference's np.hstack does, space.py:967-974)."""
    arr = np.asarray(col)
    if arr.dtype.kind in "biuf":
        arr = arr.astype(float, copy=False)
    return arr.reshape((n, -1))


def device_dim_descs(space):
    """(kind, low, high) per dimension if every dimension's "normalize" arithmetic has a bit-exact device version
    (Real with a uniform prior, Integer with a uniform prior, Categorical with >= 2 categories), else None."""
    descs = []
    for dim in space.dimensions:
        if dim.transform_ != "normalize":
            return None
        if isinstance(dim, Real) and dim.prior == "uniform" and dim.dtype in (float, "float"):
            descs.append((0, float(dim.low), float(dim.high)))
        elif isinstance(dim, Integer) and dim.prior == "uniform":
            descs.append((1, float(dim.low), float(dim.high)))
        elif isinstance(dim, Categorical) and len(dim.categories) >= 2:
            descs.append((1, 0.0, float(len(dim.categories) - 1)))
        else:
            return None
    return descs


def normalize_dimensions(dimensions):
    """Space whose dimensions all use the "normalize" warping (skopt/utils.py:569-607)."""
    space = Space(dimensions)
    for dim in space.dimensions:
        if not isinstance(dim, Dimension):
            raise RuntimeError("Unknown dimension type (%s)" % type(dim))
        dim.set_transformer("normalize")
    return Space(space.dimensions)
