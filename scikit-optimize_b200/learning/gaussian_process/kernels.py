"""Kernel classes with the names of skopt.learning.gaussian_process.kernels.

On the GPU path the kernel values AND their input gradients are computed by the CUDA kernels from the flattened
description (engine._flatten_kernel), so these classes only have to build, clone, tune (sklearn's marginal
likelihood) and describe the kernel; they are sklearn's kernels with operator overloads that keep the result inside
this module.  Kernels outside the hot-path grammar (RationalQuadratic, ExpSineSquared, DotProduct, Exponentiation,
HammingKernel) are intentionally not provided: asking the GPU path for them raises.
Mirrors skopt/learning/gaussian_process/kernels.py:20-46 (operators), :68, :93, :260, :266, :285, :294 (classes).
"""
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
