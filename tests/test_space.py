"""Candidate generation: the vectorised Space must reproduce the reference's list-based
`space.transform(space.rvs(n, rng))` bit for bit and leave the RandomState at the same position."""
import json
import os

import numpy as np
import pytest

import skopt_b200  # noqa: F401
from skopt_b200.space import Categorical, Integer, Real, Space, check_dimension, normalize_dimensions

GOLD = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "traj_space_candidates.npz")))
SPECS = json.loads(str(GOLD["specs"]))


def _spec(name):
    return [tuple(d) if not isinstance(d[0], (str, bool)) else list(d) for d in SPECS[name]]


@pytest.mark.parametrize("name", sorted(SPECS))
def test_candidate_matrix_bit_identical(name):
    space = normalize_dimensions(_spec(name))
    rng = np.random.RandomState(42)
    Xt = space.rvs_transformed(257, rng)
    assert Xt.dtype == np.float64 and Xt.flags["C_CONTIGUOUS"]
    assert np.array_equal(Xt, GOLD[name + "_Xt"])
    assert rng.randint(0, 2 ** 31 - 1) == int(GOLD[name + "_after"])
    assert np.array_equal(np.array(space.transformed_bounds, dtype=float), GOLD[name + "_bounds"])


@pytest.mark.parametrize("name", sorted(SPECS))
def test_rvs_transform_inverse_roundtrip(name):
    space = normalize_dimensions(_spec(name))
    rng = np.random.RandomState(42)
    pts = space.rvs(257, random_state=rng)
    assert pts[:20] == json.loads(str(GOLD[name + "_pts"]))
    assert np.array_equal(space.transform(pts), GOLD[name + "_Xt"])
    assert space.inverse_transform(GOLD[name + "_Xt"][:20]) == json.loads(str(GOLD[name + "_back"]))
    assert all(p in space for p in pts)


def test_dimension_dispatch_and_validation():
    assert isinstance(check_dimension((1, 5)), Integer)
    assert isinstance(check_dimension((1.0, 5)), Real)
    assert isinstance(check_dimension((1e-3, 1.0, "log-uniform")), Real)
    assert isinstance(check_dimension(["a", "b"]), Categorical)
    assert isinstance(check_dimension((True, False)), Categorical)
    with pytest.raises(ValueError):
        Real(1.0, 1.0)
    with pytest.raises(ValueError):
        Real(0.0, 1.0, prior="gaussian")
    with pytest.raises(ValueError):
        check_dimension(3)
    sp = Space([(0.0, 1.0), (1, 3)])
    assert sp.n_dims == 2 and sp.transformed_n_dims == 2 and not sp.is_categorical
    assert sp.distance([0.25, 1], [0.75, 3]) == 2.5
    with pytest.raises(ValueError):
        normalize_dimensions([(0.0, 1.0)]).transform([[1.5]])


def test_min_distance_equals_reference_loop():
    """Space.min_distance (used by Optimizer._ask) returns exactly min(space.distance(x, p) for p in points)."""
    from skopt_b200.space import Categorical, Integer, Real, Space
    rng = np.random.RandomState(0)
    for dims in ([Real(-3.0, 5.0), Real(1e-3, 10.0, prior="log-uniform"), Real(0.0, 1.0)],
                 [Integer(-5, 50), Real(0.0, 2.0), Categorical(["a", "b", "c"]), Integer(0, 3)]):
        space = Space(dims)
        pts = space.rvs(200, random_state=rng)
        for x in space.rvs(5, random_state=rng) + [pts[17]]:
            ref = min(space.distance(x, p) for p in pts)
            got = space.min_distance(x, pts)
            assert got == ref and type(float(got)) is float
    space = Space([Real(0.0, 1.0)])
    with pytest.raises(RuntimeError):
        space.min_distance([0.5], [[2.0]])
    with pytest.raises(RuntimeError):
        space.min_distance([3.0], [[0.5]])
