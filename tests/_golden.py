"""Helpers to load the reference-generated fixtures in tests/golden (see oracle/make_golden.py)."""
import glob
import json
import os

import numpy as np
from sklearn.gaussian_process import kernels as skk

from oracle import gp_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
               if not os.path.basename(p).startswith(("traj_", "fit_")))


def _build_kernel(params, cls, prefix=""):
    """Rebuild the fitted kernel_ tree from the flat get_params() dump stored in the fixture."""
    if cls in ("Sum", "Product"):
        k1 = _build_kernel(params, params[prefix + "k1"], prefix + "k1__")
        k2 = _build_kernel(params, params[prefix + "k2"], prefix + "k2__")
        return getattr(skk, cls)(k1, k2)
    if cls == "ConstantKernel":
        return skk.ConstantKernel(params[prefix + "constant_value"], "fixed")
    if cls == "WhiteKernel":
        return skk.WhiteKernel(params[prefix + "noise_level"], "fixed")
    if cls == "RBF":
        return skk.RBF(np.asarray(params[prefix + "length_scale"]), "fixed")
    if cls == "Matern":
        return skk.Matern(np.asarray(params[prefix + "length_scale"]), "fixed", nu=params[prefix + "nu"])
    if cls == "HammingKernel":
        from skopt_b200.learning.gaussian_process.kernels import HammingKernel
        return HammingKernel(np.asarray(params[prefix + "length_scale"]), "fixed")
    raise NotImplementedError(cls)


def kernel_code(st):
    """skb_kernel_kind (include/skb.h) of an oracle state."""
    if st.kind == orc.KIND_HAMMING:
        return 4
    return {(orc.KIND_RBF, np.inf): 0, (orc.KIND_MATERN, 0.5): 1, (orc.KIND_MATERN, 1.5): 2,
            (orc.KIND_MATERN, 2.5): 3}[(st.kind, st.nu)]


class _Est:
    pass


def load(name):
    rec = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    meta = json.loads(str(rec["meta"]))
    est = _Est()
    est.kernel_ = _build_kernel(meta["kernel_params"], meta["kernel_class"])
    est.X_train_, est.alpha_, est.K_inv_, est.L_ = rec["X_train"], rec["alpha"], rec["K_inv"], rec["L"]
    est.y_train_mean_, est.y_train_std_ = float(rec["y_train_mean"]), float(rec["y_train_std"])
    return rec, meta, est


def load_state(name):
    rec, meta, est = load(name)
    return rec, meta, orc.state_from_estimator(est)


def var_tol(st, floor=1e-9):
    """Tolerance on the posterior variance relative to the prior variance.

    1e-9 on well-conditioned states; on ill-conditioned ones the reference's own result moves by ~eps*cond(K)
    when only the summation order of k^T K^-1 k changes (hard part 1), so the gate widens.
    """
    cond = float(np.linalg.cond(st.K_inv))
    return max(floor, 1e-15 * cond)
