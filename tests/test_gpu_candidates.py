"""Device-side candidate generation: NumPy's MT19937 stream continued on the GPU and the Space warping arithmetic,
both bit for bit."""
import numpy as np
import pytest

import skopt_b200
from skopt_b200 import engine
from skopt_b200.space import Space, device_dim_descs, normalize_dimensions

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,burn,count", [(0, 0, 1), (1, 0, 312), (2, 1, 313), (3, 623, 5000), (4, 311, 2), (5, 7, 100001),
                                             (6, 0, 624 * 40), (7, 3, 0)])
def test_mt19937_stream_continues_bit_for_bit(seed, burn, count):
    ref = np.random.RandomState(seed)
    dev = np.random.RandomState(seed)
    if burn:                                    # move to an arbitrary (also odd) position inside the 624-word state
        ref.randint(0, 2 ** 31 - 1, size=burn)
        dev.randint(0, 2 ** 31 - 1, size=burn)
    expect = ref.uniform(0.0, 1.0, size=count)
    got = engine.device_uniform(dev, count).cpu().numpy()
    assert np.array_equal(got, expect)
    # the host generator was advanced exactly as if it had produced the numbers itself
    assert np.array_equal(ref.uniform(size=7), dev.uniform(size=7))
    assert ref.randint(0, 2 ** 31 - 1) == dev.randint(0, 2 ** 31 - 1)


SPECS = {
    "real2": [(-5.0, 10.0), (0.0, 15.0)],
    "real20": [(0.0, 1.0)] * 20,
    "mixed": [(-2.0, 3.5), (1, 17), ["a", "b", "c", "d"], (0, 1), [True, False], (1e-3, 7.0)],
    "wide": [(-1e6, 1e6), (3, 4), (0.1, 0.1000001)],
}


@pytest.mark.parametrize("name", sorted(SPECS))
@pytest.mark.parametrize("n", [1, 257, 10000])
def test_device_candidates_equal_host_candidates(name, n):
    space = normalize_dimensions(SPECS[name])
    descs = device_dim_descs(space)
    assert descs is not None
    r_host, r_dev = np.random.RandomState(42), np.random.RandomState(42)
    X_host = space.rvs_transformed(n, r_host)
    X_dev = engine.device_candidates(descs, n, r_dev).cpu().numpy()
    assert X_dev.shape == X_host.shape and np.array_equal(X_dev, X_host)
    assert r_host.randint(0, 2 ** 31 - 1) == r_dev.randint(0, 2 ** 31 - 1)


@pytest.mark.parametrize("name", sorted(SPECS))
@pytest.mark.parametrize("n,burn", [(312, 0), (1000, 5), (10007, 623), (70001, 311)])
def test_jump_generator_equals_serial_generator_and_host(name, n, burn):
    """One CTA per dimension, each started by an MT19937 jump (mt_jump.py, csrc/rng.cuh): same matrix, bit for bit, and
    the host generator ends in NumPy's own state (key AND position)."""
    space = normalize_dimensions(SPECS[name])
    descs = device_dim_descs(space)
    rngs = [np.random.RandomState(77) for _ in range(3)]
    for r in rngs:
        r.randint(0, 2 ** 31 - 1, size=burn)
    X_host = space.rvs_transformed(n, rngs[0])
    X_serial = engine.device_candidates(descs, n, rngs[1], jump=False).cpu().numpy()
    X_jump = engine.device_candidates(descs, n, rngs[2], jump=True).cpu().numpy()
    assert np.array_equal(X_serial, X_host) and np.array_equal(X_jump, X_host)
    ref = rngs[0].get_state()
    for r in rngs[1:]:
        st = r.get_state()
        assert st[2] == ref[2] and np.array_equal(st[1], ref[1])
    assert rngs[0].randint(0, 2 ** 31 - 1) == rngs[2].randint(0, 2 ** 31 - 1)


def test_unsupported_dimensions_fall_back_to_the_host():
    assert device_dim_descs(normalize_dimensions([(1e-4, 1e2, "log-uniform"), (0.0, 1.0)])) is None
    assert device_dim_descs(Space([(0.0, 1.0)])) is None            # identity transform: not the GP path
    assert device_dim_descs(normalize_dimensions([["only"], (0.0, 1.0)])) is None


def test_optimizer_same_points_with_and_without_device_candidates():
    from skopt_b200 import Optimizer
    pts = {}
    for flag in ("always", False):
        opt = Optimizer([(-5.0, 10.0), (0, 7), ["x", "y", "z"]], "GP", n_initial_points=5, acq_func="gp_hedge",
                        acq_optimizer="sampling", random_state=3,
                        acq_optimizer_kwargs=dict(n_points=4000, device_candidates=flag))
        for _ in range(9):
            x = opt.ask()
            opt.tell(x, float((x[0] - 1.0) ** 2 + 0.1 * x[1] + {"x": 0.0, "y": 1.0, "z": -1.0}[x[2]]))
        pts[flag] = opt.Xi
    assert pts["always"] == pts[False]
