"""1-D warpings between the original search space and the space the GP sees.

Public counterparts of skopt/space/transformers.py (Identity :20, StringEncoder :31, LogN :85, CategoricalEncoder :98,
LabelEncoder :154, Normalize :221, Pipeline :279).  The dimensions in `space.py` apply the same arithmetic inline on
whole columns (that is what keeps `Space.rvs_transformed` one vectorised pass); these classes expose it under the
reference's names for code that builds or inspects transformers directly.  Every class works on whole arrays and
performs the reference's floating-point operations in the reference's order, so results are the same bits.
"""
import numpy as np

_EPS = 1e-8


class Transformer(object):
    def fit(self, X):
        return self

    def transform(self, X):
        raise NotImplementedError

    def inverse_transform(self, X):
        raise NotImplementedError


class Identity(Transformer):
    def transform(self, X):
        return X

    def inverse_transform(self, Xt):
        return Xt


class StringEncoder(Transformer):
    """str() on the way in, the type of the first fitted category on the way back (transformers.py:31-82)."""

    def __init__(self, dtype=str):
        self.dtype = dtype

    def fit(self, X):
        if len(X) > 0:
            self.dtype = type(X[0])

    def transform(self, X):
        return list(map(str, X))

    def inverse_transform(self, Xt):
        return list(map(self.dtype, Xt))


class LogN(Transformer):
    """log_base(x) as log10(x) / log10(base) (two roundings, like transformers.py:91-95)."""

    def __init__(self, base):
        self._base = base

    def transform(self, X):
        return np.log10(np.asarray(X, dtype=float)) / np.log10(self._base)

    def inverse_transform(self, Xt):
        return self._base ** np.asarray(Xt, dtype=float)


def _positions(categories, sort_plain):
    """category -> code.  Object arrays (mixed types) keep the given order, plain arrays take np.unique's sorted
    order when `sort_plain` (LabelEncoder.fit, transformers.py:168-177); duplicates keep their last position."""
    arr = np.asarray(categories)
    if sort_plain and arr.dtype != object:
        return {v: i for i, v in enumerate(np.unique(arr))}
    return {v: i for i, v in enumerate(categories)}


class CategoricalEncoder(Transformer):
    """One-hot rows, except that two categories give ONE 0/1 column and a single category a column of zeros
    (sklearn LabelBinarizer's conventions, which the reference delegates to: transformers.py:101-151)."""

    def fit(self, X):
        self.mapping_ = _positions(X, sort_plain=False)
        self.inverse_mapping_ = {i: v for v, i in self.mapping_.items()}
        self._codes = np.array(sorted(self.inverse_mapping_))          # classes_ of the binarizer
        self.n_classes = len(self._codes)
        return self

    def transform(self, X):
        codes = np.fromiter((self.mapping_[v] for v in X), dtype=np.int64)
        col = np.searchsorted(self._codes, codes)
        if self.n_classes == 1:
            return np.zeros((len(codes), 1), dtype=np.int64)
        if self.n_classes == 2:
            return col.reshape(-1, 1).astype(np.int64)
        onehot = np.zeros((len(codes), self.n_classes), dtype=np.int64)
        onehot[np.arange(len(codes)), col] = 1
        return onehot

    def inverse_transform(self, Xt):
        Xt = np.asarray(Xt)
        if self.n_classes <= 2:                  # one column, thresholded at 0.5 like LabelBinarizer
            col = (Xt.reshape(len(Xt), -1)[:, 0] > 0.5).astype(int) if self.n_classes == 2 else np.zeros(len(Xt), int)
        else:
            col = np.argmax(Xt.reshape(len(Xt), -1), axis=1)
        return [self.inverse_mapping_[int(self._codes[c])] for c in col]


class LabelEncoder(Transformer):
    """Integer codes; the inverse rounds (transformers.py:154-218)."""

    def __init__(self, X=None):
        if X is not None:
            self.fit(X)

    def fit(self, X):
        self.mapping_ = _positions(X, sort_plain=True)
        self.inverse_mapping_ = {i: v for v, i in self.mapping_.items()}
        return self

    def transform(self, X):
        return [self.mapping_[v] for v in np.asarray(X)]

    def inverse_transform(self, Xt):
        Xt = [Xt] if isinstance(Xt, (float, np.floating)) else np.asarray(Xt)
        return [self.inverse_mapping_[int(np.round(i))] for i in Xt]


class Normalize(Transformer):
    """[low, high] -> [0, 1] with the reference's bound checks (1e-8 slack) and operation order
    (transformers.py:236-276).  `is_int` rounds first on the way in and rounds + casts on the way back."""

    def __init__(self, low, high, is_int=False):
        self.low = float(low)
        self.high = float(high)
        self.is_int = is_int
        self._eps = _EPS

    def transform(self, X):
        X = np.asarray(X)
        probe, slack = (np.round(X), 0.0) if self.is_int else (X, self._eps)
        if np.any(probe > self.high + slack):
            raise ValueError("All %svalues shouldbe less than %f" % ("integer " if self.is_int else "", self.high))
        if np.any(probe < self.low - slack):
            raise ValueError("All %svalues shouldbe greater than %f" % ("integer " if self.is_int else "", self.low))
        width = self.high - self.low
        if width == 0.0:
            return X * 0.0
        if self.is_int:
            return (np.round(X).astype(int) - self.low) / width
        return (X - self.low) / width

    def inverse_transform(self, X):
        X = np.asarray(X)
        if np.any(X > 1.0 + self._eps):
            raise ValueError("All values should be less than 1.0")
        if np.any(X < 0.0 - self._eps):
            raise ValueError("All values should be greater than 0.0")
        X_orig = X * (self.high - self.low) + self.low
        return np.round(X_orig).astype(int) if self.is_int else X_orig


class Pipeline(Transformer):
    def __init__(self, transformers):
        self.transformers = list(transformers)
        for t in self.transformers:
            if not isinstance(t, Transformer):
                raise ValueError("Provided transformers should be a Transformer instance. Got %s" % t)

    def fit(self, X):
        for t in self.transformers:
            t.fit(X)
        return self

    def transform(self, X):
        for t in self.transformers:
            X = t.transform(X)
        return X

    def inverse_transform(self, X):
        for t in reversed(self.transformers):
            X = t.inverse_transform(X)
        return X
