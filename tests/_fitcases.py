"""Shared by the CPU and GPU tests of the fit objective: rebuild the kernels of tests/golden/lml.npz
(oracle/make_golden_lml.py) with this package's kernel classes."""
import os

import numpy as np

from skopt_b200.learning.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fit_lml.npz")
NAMES = ["m52_ard", "m32_ard", "m12_ard", "rbf_ard", "m52_iso", "m52_nonoise", "m52_big"]


def load_case(name):
    z = np.load(GOLDEN)
    nu, iso = float(z[name + "_nu"]), bool(z[name + "_iso"])
    ls = z[name + "_length_scale"]
    ls = float(ls[0]) if iso else ls
    base = RBF(length_scale=ls, length_scale_bounds=(0.01, 100)) if nu == np.inf else \
        Matern(length_scale=ls, length_scale_bounds=(0.01, 100), nu=nu)
    k = ConstantKernel(1.7, "fixed" if bool(z[name + "_fixed_amp"]) else (0.01, 1000.0)) * base
    noise_kind = str(z[name + "_noise_kind"])
    if noise_kind == "free":
        k = k + WhiteKernel(0.05, (1e-6, 10.0))
    elif noise_kind == "fixed":
        k = k + WhiteKernel(0.02, "fixed")
    assert np.allclose(k.theta, z[name + "_theta0"], rtol=1e-14, atol=0)
    return dict(X=z[name + "_X"], y=z[name + "_y"], nu=nu, kernel=k, thetas=z[name + "_thetas"],
                lml=z[name + "_lml"], grad=z[name + "_grad"])


def check_against_golden(case, evaluate, rtol_lml=1e-10, rtol_grad=1e-9):
    """evaluate(theta_full) -> (lml, grad_full[d + 2]); compared with the reference's values for every stored theta."""
    from skopt_b200.engine import FitLayout
    lay = FitLayout(case["kernel"], case["X"].shape[1])
    for th, v_ref, g_ref in zip(case["thetas"], case["lml"], case["grad"]):
        v, g_full = evaluate(lay.expand(th))
        g = lay.select(np.asarray(g_full))
        assert abs(v - v_ref) <= rtol_lml * max(1.0, abs(v_ref)), (v, v_ref)
        assert np.max(np.abs(g - g_ref)) <= rtol_grad * max(1.0, np.max(np.abs(g_ref))), (g, g_ref)
