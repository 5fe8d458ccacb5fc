"""Numeric half of skopt/plots.py that consumes the GP posterior: partial dependence (plots.py:896-1055) and the
evenly spaced grid it is evaluated on (plots.py:1298-1331).  No drawing code (matplotlib is not part of the hot path).

The reference evaluates `np.mean(model.predict(samples with dimension i pinned to x))` once per grid value - 40 predict
calls per dimension and 1 600 per pair of dimensions, each over `n_samples` (250) rows.  Here every (grid value,
sample) row of a dimension or of a pair goes into ONE mean-only pass of the hot path (K1, skb_predict_mean_host):
400 000 rows for a pair.  The averages are taken per grid value with `np.mean` over the same rows in the same order,
so the only difference to the reference is the 1e-13-level difference of the predictions themselves.
"""
import numpy as np


def _evenly_sample(dim, n_points):
    """`n_points` evenly spaced values of a dimension and their transformed (model-space) form; for a Categorical
    dimension evenly spaced category indices, at most one per category (plots.py:1298-1331)."""
    cats = np.array(getattr(dim, "categories", []), dtype=object)
    if len(cats):
        xi = np.linspace(0, len(cats) - 1, min(len(cats), n_points), dtype=int)
        xi_transformed = dim.transform(cats[xi])
    else:
        low, high = dim.bounds
        xi = np.linspace(low, high, n_points)
        xi_transformed = dim.transform(xi)
    return xi, xi_transformed


def _column_ranges(space):
    starts = np.cumsum([0] + [dim.transformed_size for dim in space.dimensions])
    return [(int(starts[k]), int(starts[k + 1])) for k in range(len(space.dimensions))]


def _pinned_rows(samples, pins):
    """Stack one copy of `samples` per grid value; `pins` = [(column range, values per copy)], later entries written
    last (the reference writes dimension j first, then i: for i == j the value of i wins)."""
    samples = np.array(samples, dtype=np.float64)
    count = len(pins[0][1])
    rows = np.tile(samples, (count, 1)).reshape(count, samples.shape[0], samples.shape[1])
    for (lo, hi), values in pins:
        values = np.asarray(values, dtype=np.float64).reshape(count, -1)
        rows[:, :, lo:hi] = values[:, None, :]
    return rows.reshape(count * samples.shape[0], samples.shape[1]), samples.shape[0]


def partial_dependence_1D(space, model, i, samples, n_points=40):
    """plots.py:896-971.  Average predicted objective with dimension `i` pinned to each of `n_points` evenly spaced
    values, the other dimensions taken from `samples` (transformed points, shape (n_samples, transformed dims)).
    Returns (xi, yi) like the reference (yi as a list of floats)."""
    ranges = _column_ranges(space)
    xi, xi_transformed = _evenly_sample(space.dimensions[i], n_points)
    rows, n_samples = _pinned_rows(samples, [(ranges[i], xi_transformed)])
    pred = np.asarray(model.predict(rows)).reshape(len(xi), n_samples)
    return xi, [np.mean(row) for row in pred]


def partial_dependence_2D(space, model, i, j, samples, n_points=40):
    """plots.py:974-1055.  zi[a, b] = average prediction with dimension `i` pinned to yi[a] and dimension `j` to xi[b].
    Returns (xi, yi, zi): xi are the values of dimension j, yi those of dimension i (the reference's order)."""
    ranges = _column_ranges(space)
    xi, xi_transformed = _evenly_sample(space.dimensions[j], n_points)
    yi, yi_transformed = _evenly_sample(space.dimensions[i], n_points)
    xt = np.asarray(xi_transformed, dtype=np.float64).reshape(len(xi), -1)
    yt = np.asarray(yi_transformed, dtype=np.float64).reshape(len(yi), -1)
    # grid in row-major (a over yi, b over xi) order
    x_rep = np.tile(xt, (len(yi), 1))
    y_rep = np.repeat(yt, len(xi), axis=0)
    rows, n_samples = _pinned_rows(samples, [(ranges[j], x_rep), (ranges[i], y_rep)])
    pred = np.asarray(model.predict(rows)).reshape(len(yi) * len(xi), n_samples)
    zi = np.array([np.mean(row) for row in pred]).reshape(len(yi), len(xi))
    return xi, yi, zi
