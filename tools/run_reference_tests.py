"""Run the reference's OWN test files against this package (drop-in check of the path's Python boundary).

Build container only: the tests are read from /root/reference, copied into a scratch directory (gpurun_out/reftests,
git-ignored) with their `skopt` imports for the GP path redirected to `skopt_b200`, and run with pytest.  Nothing of the
reference is copied into the repository.  Imports that stay on the reference: benchmarks (objective functions),
callbacks, tree / GBRT learners and their minimisers (out of scope here; most of those tests fail in the reference itself
under scikit-learn >= 1.2 because of criterion='mse').

    python tools/run_reference_tests.py            # no GPU: the C ABI's host entry points are the software model
                                                   # tests/_abi_model.py (TEST INFRASTRUCTURE; arithmetic by the oracle)
    python tools/run_reference_tests.py --device   # on a box with a B200 and /root/reference: the real libskb.so

Deselected (out of scope): tree / GBRT surrogates, the sobol / lhs / halton /
hammersly / grid initial-point generators, and tests that cannot run on this pytest / scikit-learn at all
(`pytest.warns(None)`, tree estimators with criterion='mse').
"""
import argparse
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/skopt"
FILES = ["tests/test_space.py", "tests/test_transformers.py", "tests/test_utils.py", "tests/test_acquisition.py",
         "tests/test_optimizer.py", "tests/test_gp_opt.py", "tests/test_common.py", "tests/test_parallel_cl.py",
         "learning/gaussian_process/tests/test_gpr.py", "learning/gaussian_process/tests/test_kernels.py"]
OURS = {"gp_minimize", "dummy_minimize", "Optimizer", "Space", "load", "dump", "expected_minimum", "expected_minimum_random_sampling"}
REDIRECT = [
    (r"^(\s*)from skopt\.optimizer import Optimizer", r"\1from skopt_b200 import Optimizer"),
    (r"^(\s*)from skopt\.acquisition import", r"\1from skopt_b200.acquisition import"),
    (r"^(\s*)from skopt\.learning import GaussianProcessRegressor$", r"\1from skopt_b200.learning import GaussianProcessRegressor"),
    (r"^(\s*)from skopt\.learning\.gaussian_process import", r"\1from skopt_b200.learning.gaussian_process import"),
    (r"^(\s*)from skopt\.learning\.gaussian_process\.(kernels|gpr) import", r"\1from skopt_b200.learning.gaussian_process.\2 import"),
    (r"^(\s*)from skopt\.space(\.space|\.transformers)? import", r"\1from skopt_b200.space import"),
    (r"^(\s*)from skopt\.utils import", r"\1from skopt_b200.utils import"),
    (r"^(\s*)from skopt\.benchmarks import", r"\1from skopt_b200.benchmarks import"),
]
# node ids (regular expressions) of tests about components outside the path
OUT_OF_SCOPE = [r"forest_minimize", r"gbrt_minimize", r"RandomForest", r"ExtraTrees", r"GradientBoosting",
                r"\[(.*-)?(RF|ET|GBRT|rf|et|gbrt)(-.*)?\]", r"initgen", r"\[(.*-)?(lhs|halton|hammersly|sobol|grid)(-.*)?\]",
                r"test_common.py::.*\[(optimizer[23]|minimizer[1-9])", r"base_estimator[0-9]",
                # test_kernels.py KERNELS list: RationalQuadratic, ExpSineSquared, Exponentiation, RBF + Matern, RBF * Matern,
                # DotProduct - outside the hot-path grammar
                r"test_kernels.py::test_gradient_(correctness|finiteness)\[kernel(4|5|8|9|10|11|16|17|20|21|22|23)[\]-]"]
BROKEN_IN_REFERENCE = ["test_common.py::test_repeated_x", "test_utils.py::test_dump_and_load_optimizer",
                       "test_optimizer.py::test_multiple_asks", "test_optimizer.py::test_model_queue_size",
                       "test_optimizer.py::test_returns_result_object", "test_optimizer.py::test_optimizer_copy",
                       "test_optimizer.py::test_acq_optimizer", "test_optimizer.py::test_acq_optimizer_with_time_api",
                       "test_optimizer.py::test_optimizer_base_estimator_string_smoke_njobs"]

CONFTEST = '''import sys
import numpy as np
import pytest
if not hasattr(np, "int"):
    np.int = int                      # the reference predates NumPy 1.24
sys.dont_write_bytecode = True
sys.path.insert(0, %(root)r)
sys.path.insert(1, "/root/reference")
USE_MODEL = %(model)r


def pytest_configure(config):
    for mark in ("fast_test", "slow_test"):
        config.addinivalue_line("markers", mark + ": reference marker")


def pytest_collection_modifyitems(config, items):
    import re
    patterns = [re.compile(p) for p in %(out_of_scope)r]
    keep, drop = [], []
    for item in items:
        (drop if any(p.search(item.nodeid) for p in patterns) else keep).append(item)
    config.hook.pytest_deselected(items=drop)
    items[:] = keep


@pytest.fixture(autouse=True)
def _abi(monkeypatch):
    if not USE_MODEL:
        return None
    import skopt_b200  # noqa: F401
    from skopt_b200 import _lib
    from skopt_b200.optimizer import optimizer as optimizer_module
    from tests._abi_model import AbiModel
    model = AbiModel()
    monkeypatch.setattr(_lib, "_lib", model)
    monkeypatch.setattr(optimizer_module, "device_dim_descs", lambda space: None)
    return model
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", action="store_true", help="use the real libskb.so (needs a B200) instead of the ABI model")
    ap.add_argument("--keep", action="store_true", help="keep the scratch directory")
    args, extra = ap.parse_known_args()
    if not os.path.isdir(REF):
        sys.exit("the reference tree is not present on this machine")
    scratch = os.path.join(ROOT, "gpurun_out", "reftests")
    shutil.rmtree(scratch, ignore_errors=True)
    os.makedirs(scratch)
    for rel in FILES:
        text = open(os.path.join(REF, rel)).read()
        for pattern, repl in REDIRECT:
            text = re.sub(pattern, repl, text, flags=re.M)

        def split_import(match):        # `from skopt import a, b`: the GP path's names come from this package
            indent, names = match.group(1), [n.strip() for n in match.group(2).split(",")]
            mine, theirs = [n for n in names if n in OURS], [n for n in names if n not in OURS]
            lines = ["%sfrom skopt_b200 import %s" % (indent, ", ".join(mine))] if mine else []
            lines += ["%sfrom skopt import %s" % (indent, ", ".join(theirs))] if theirs else []
            return "\n".join(lines)
        text = re.sub(r"^(\s*)from skopt import ([A-Za-z_0-9, ]+)$", split_import, text, flags=re.M)
        # test_common builds its list of minimisers at import time: keep the GP one
        text = re.sub(r"^for est, acq in product\(\[\"ET\", \"RF\"\], ACQUISITION\):.*?(?=^def )", "", text, flags=re.M | re.S)
        open(os.path.join(scratch, os.path.basename(rel)), "w").write(text)
    open(os.path.join(scratch, "conftest.py"), "w").write(
        CONFTEST % {"root": ROOT, "model": not args.device, "out_of_scope": OUT_OF_SCOPE})
    cmd = [sys.executable, "-m", "pytest", ".", "-q", "--no-header", "-p", "no:cacheprovider", "--rootdir", ".", "-W", "ignore",
           "--tb=line"]
    for item in BROKEN_IN_REFERENCE:
        cmd += ["--deselect", item]
    rc = subprocess.call(cmd + extra, cwd=scratch, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    if not args.keep:
        shutil.rmtree(scratch, ignore_errors=True)
    sys.exit(rc)


if __name__ == "__main__":
    main()
