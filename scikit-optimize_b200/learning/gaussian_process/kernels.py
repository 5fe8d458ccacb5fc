"""Kernel classes with the names of skopt.learning.gaussian_process.kernels.

On the GPU path the kernel values AND their input gradients are computed by the CUDA kernels from the flattened
description (engine._flatten_kernel), so these classes only have to build, clone, tune (sklearn's marginal
likelihood) and describe the kernel (`gradient_x`, public in the reference, is answered by the device too); they are sklearn's kernels with operator overloads that keep the result inside
this module.  Kernels outside the hot-path grammar (RationalQuadratic, ExpSineSquared, DotProduct, Exponentiation;
SURVEY.md section 2 row 3) exist under their names so that code importing them keeps loading and can build and fit
them on the host, but they have no device form: predict / gradient_x on a model that contains one raises
NotImplementedError naming the grammar (engine._flatten_kernel) - there is no CPU route to fall back to.
Mirrors skopt/learning/gaussian_process/kernels.py:20-46 (operators), :68, :93, :260, :266, :285, :294, :317 (classes).
"""
import numpy as np
from sklearn.gaussian_process import kernels as _sk


class Kernel(_sk.Kernel):
    def __add__(self, b):
        return Sum(self, b if isinstance(b, _sk.Kernel) else ConstantKernel(b))

    def __radd__(self, b):
        return Sum(b if isinstance(b, _sk.Kernel) else ConstantKernel(b), self)

    def __mul__(self, b):
        return Product(self, b if isinstance(b, _sk.Kernel) else ConstantKernel(b))

    def __rmul__(self, b):
        return Product(b if isinstance(b, _sk.Kernel) else ConstantKernel(b), self)

    def __pow__(self, b):
        return Exponentiation(self, b)

    def gradient_x(self, x, X_train):
        """Gradient of K(x, X_train) with respect to the single point x, shape (n_samples, n_features)
        (kernels.py:48-65; RBF :68-90, Matern :93-200, Constant / White :260-269, Sum :285-291, Product :294-308),
        evaluated on the GPU by the same device code the posterior gradients use (engine.kernel_gradient_x)."""
        from ... import engine
        return engine.kernel_gradient_x(self, x, X_train)


class RBF(Kernel, _sk.RBF):
    pass


class Matern(Kernel, _sk.Matern):
    pass


class ConstantKernel(Kernel, _sk.ConstantKernel):
    pass


class WhiteKernel(Kernel, _sk.WhiteKernel):
    pass


class Sum(Kernel, _sk.Sum):
    pass


class Product(Kernel, _sk.Product):
    pass


class Exponentiation(Kernel, _sk.Exponentiation):
    pass


class RationalQuadratic(Kernel, _sk.RationalQuadratic):
    pass


class ExpSineSquared(Kernel, _sk.ExpSineSquared):
    pass


class DotProduct(Kernel, _sk.DotProduct):
    pass


class HammingKernel(_sk.StationaryKernelMixin, _sk.NormalizedKernelMixin, Kernel):
    """k(x, y) = exp(-sum_j length_scale_j * [x_j != y_j]) for categorical inputs (skopt kernels.py:317-407): the
    kernel cook_estimator gives an all-categorical space (utils.py:377-378).  It has no gradient_x, so such models run
    with acq_optimizer="sampling" (utils.py:320-330).  On the host this class only serves scikit-learn's marginal-
    likelihood fit (value + d/d log length_scale, the same NumPy expressions and order as the reference so the fitted
    hyper-parameters are bit-identical); prediction goes to the GPU as SKB_KERNEL_HAMMING (csrc/kcross.cuh)."""

    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-5, 1e5)):
        self.length_scale = length_scale
        self.length_scale_bounds = length_scale_bounds

    def _weights(self):
        """(per-dimension weights or scalar, anisotropic?)"""
        ls = self.length_scale
        if np.iterable(ls):
            if len(ls) > 1:
                return np.asarray(ls, dtype=float), True
            return float(ls[0]), False
        return float(ls), False

    @property
    def hyperparameter_length_scale(self):
        w, aniso = self._weights()
        if aniso:
            return _sk.Hyperparameter("length_scale", "numeric", self.length_scale_bounds, len(w))
        return _sk.Hyperparameter("length_scale", "numeric", self.length_scale_bounds)

    def __call__(self, X, Y=None, eval_gradient=False):
        w, aniso = self._weights()
        X = np.atleast_2d(X)
        if aniso and X.shape[1] != len(w):
            raise ValueError("Expected X to have %d features, got %d" % (len(w), X.shape[1]))
        if Y is None:
            Y = X
        elif eval_gradient:
            raise ValueError("gradient can be evaluated only when Y != X")
        else:
            Y = np.atleast_2d(Y)
        differs = X[:, np.newaxis, :] != Y                      # (n_X, n_Y, d) mismatch indicator
        K = np.exp(-np.sum(w * differs, axis=2))
        if not eval_gradient:
            return K
        # theta = log(length_scale): dK/dtheta_j = -K * [x_j != y_j] * length_scale_j (one shared theta when isotropic)
        if aniso:
            grad = -K[:, :, np.newaxis] * np.array(differs, dtype=np.float32)
        else:
            grad = -(K * np.sum(differs, axis=2))[:, :, np.newaxis]
        grad *= w
        return K, grad
