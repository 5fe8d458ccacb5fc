"""Objective functions of the reference's benchmarks and tests (skopt/benchmarks.py): the two BASELINE.json objectives
(Branin :69, Hartmann-6 :83) and the small test functions bench1 ... bench5 (:7-66).  Points are sequences; Branin and
Hartmann use the standard constants."""
import numpy as np


def bench1(x):
    """x^2: one minimum, 0 at x = 0."""
    return x[0] ** 2


def bench1_with_time(x):
    """bench1 together with a constant evaluation time, for the per-second acquisitions."""
    return x[0] ** 2, 2.22


def bench2(x):
    """x^2 left of zero, (x - 5)^2 - 5 right of it: global minimum -5 at x = 5."""
    return x[0] ** 2 if x[0] < 0 else (x[0] - 5) ** 2 - 5


def bench3(x):
    """sin(5x) (1 - tanh(x^2)): global minimum about -0.9 near x = -0.3."""
    return np.sin(5 * x[0]) * (1 - np.tanh(x[0] ** 2))


def bench4(x):
    """float(x)^2 for a string-valued x (categorical spaces)."""
    return float(x[0]) ** 2


def bench5(x):
    """float(x0)^2 + x1^2 (mixed categorical / numeric spaces)."""
    return float(x[0]) ** 2 + x[1] ** 2


def branin(x, a=1, b=5.1 / (4 * np.pi ** 2), c=5. / np.pi, r=6, s=10, t=1. / (8 * np.pi)):
    """Branin-Hoo on [-5, 10] x [0, 15]: three global minima of 0.397887 at (-pi, 12.275), (pi, 2.275), (9.42478, 2.475)."""
    return a * (x[1] - b * x[0] ** 2 + c * x[0] - r) ** 2 + s * (1 - t) * np.cos(x[0]) + s


_HART6_ALPHA = np.array([1.0, 1.2, 3.0, 3.2])
_HART6_P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                            [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
_HART6_A = np.array([[10, 3, 17, 3.50, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8],
                     [17, 8, 0.05, 10, 0.1, 14]])


def hart6(x, alpha=_HART6_ALPHA, P=_HART6_P, A=_HART6_A):
    """Hartmann 6-D on [0, 1]^6: global minimum -3.32237 at (0.20169, 0.15001, 0.476874, 0.275332, 0.311652, 0.6573)."""
    return -np.sum(alpha * np.exp(-np.sum(A * (np.array(x) - P) ** 2, axis=1)))
