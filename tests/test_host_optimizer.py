"""Host side above the C ABI, exercised on the CPU: the tests of tests/test_gpu_optimizer.py and
tests/test_gpu_reference_behaviours.py that only use HOST entry points run here against the software model of the ABI
(tests/_abi_model.py, `abi_model` fixture) - same fixtures from the unmodified reference, same assertions.  What this
pins without a GPU: the ctypes marshalling (struct layouts, pointer / NULL conventions, top-k and first-NaN protocol),
`GaussianProcessRegressor.predict`, the acquisition module, the Optimizer's RNG draw order, hedge bookkeeping,
constant-liar loop and the lock-step L-BFGS driver.  The arithmetic behind the ABI is the CPU oracle here; the CUDA
kernels are checked by the `-m gpu` originals."""
import numpy as np
import pytest

import skopt_b200  # noqa: F401
from tests import test_gpu_optimizer as gpu_opt


def test_model_is_what_the_package_talks_to(abi_model):
    from skopt_b200 import _lib, engine
    from oracle import gp_oracle as orc
    assert _lib.lib() is abi_model
    st = orc.make_synthetic_state(24, 3)
    L_inv = np.linalg.inv(np.linalg.cholesky(np.linalg.inv(st.K_inv)))        # K_inv = L_inv^T L_inv, L_inv lower
    hs = engine.HostState(st.X_train, st.alpha, L_inv, st.length_scale, 3, st.amplitude,
                          st.bias, st.white, st.y_mean, st.y_std, K_inv=st.K_inv)
    dev = engine.DeviceState(hs)
    X = np.random.RandomState(0).rand(50, 3)
    res = dev.score(X, y_opt=-0.3, acq_funcs=("EI", "LCB"), top_k=4)
    want = orc.score_candidates(X, st, -0.3, ("EI", "LCB"), top_k=4)
    for acq in ("EI", "LCB"):
        assert np.allclose(res[acq]["values"], want[acq]["values"], rtol=1e-6, atol=1e-9)
        assert res[acq]["topk_idx"].tolist() == want[acq]["topk"].tolist() and res[acq]["first_nan"] == -1
    assert np.allclose(dev.predict_mean(X), res["mean"])
    assert abi_model.calls[:2] == ["skb_state_create", "skb_score_host"]
    dev.close()


def _mirror(module, skip=()):
    """Re-export every test function of a `-m gpu` module as a CPU test that runs it inside the `abi_model` fixture,
    keeping its parametrisation and fixtures (the wrapper gets the original's signature plus `abi_model`)."""
    import inspect
    made = {}
    for name, func in sorted(vars(module).items()):
        if not (name.startswith("test_") and inspect.isfunction(func)) or name in skip:
            continue
        args = list(inspect.signature(func).parameters)
        scope = {"_orig": func}
        exec("def {0}(abi_model{1}):\n    return _orig({2})".format(
            name, "".join(", " + a for a in args), ", ".join("%s=%s" % (a, a) for a in args)), scope)
        wrapper = scope[name]
        wrapper.__doc__ = func.__doc__
        wrapper.__module__ = __name__
        wrapper.pytestmark = [m for m in getattr(func, "pytestmark", []) if m.name == "parametrize"]
        made[name] = wrapper
    return made


# tests/test_gpu_optimizer.py: trajectories of the unmodified reference (sampling: identical points; L-BFGS: same
# optimum), constant liar, estimator / acquisition API, per-second acquisitions, expected_minimum, dump / load
globals().update(_mirror(gpu_opt))

# tests/test_gpu_reference_behaviours.py: what the reference's own test files check on this path (convergence
# thresholds, determinism, validation, result object, constant liar, categorical spaces), restated for this package
from tests import test_gpu_reference_behaviours as gpu_behaviours  # noqa: E402

globals().update({"test_behaviour_" + name[5:]: func for name, func in _mirror(gpu_behaviours).items()})

# all-categorical GPs and partial dependence: the `-m gpu` members of the mixed modules that use host entry points only
# (test_large_categorical_candidate_set_all_precisions_agree compares precisions of the CUDA kernels: GPU only)
from tests import test_hamming as hamming_tests  # noqa: E402
from tests import test_partial_dependence as pd_tests  # noqa: E402

globals().update({"test_hamming_" + name[5:]: func for name, func in _mirror(hamming_tests).items()
                  if name in ("test_categorical_trajectory_identical",
                              "test_gradient_entry_points_refuse_the_hamming_kernel")})
globals().update({"test_pd_" + name[5:]: func for name, func in _mirror(pd_tests).items()
                  if name == "test_partial_dependence_matches_reference_on_gpu"})
