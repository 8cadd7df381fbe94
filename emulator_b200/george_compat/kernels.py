"""
Minimal ``george.kernels`` look-alike: just enough kernel algebra for the expression the
reference builds, ``scalar * ExpSquaredKernel(metric=np.ones(d), ndim=d)``
(/root/reference/swiftemulator/emulators/gaussian_process.py:181-183,
gaussian_process_bins.py:216-219, gaussian_process_mcmc.py:213-215,
gaussian_process_one_dim.py:180-182), which must be deep-copyable and carry its own
parameter vector.  Only that product maps onto the CUDA path; any other kernel expression is
rejected when the GP is built (there is no generic CPU evaluator to fall back on).
"""

from __future__ import annotations

import numpy as np


class Kernel:
    is_kernel = True
    ndim = 1

    def __mul__(self, other):
        if isinstance(other, Kernel):
            if other.ndim != self.ndim:
                raise ValueError("Dimension mismatch")
            return Product(self, other)
        # george: scalar b -> ConstantKernel(log_constant=log(b / ndim)); the constant kernel is
        # summed over its ndim axes when evaluated, so the product amplitude is b.
        return Product(ConstantKernel(log_constant=np.log(float(other) / self.ndim), ndim=self.ndim), self)

    __rmul__ = __mul__

    def __add__(self, other):
        raise NotImplementedError("kernel sums are not on the accelerated path")

    __radd__ = __add__

    # modeling protocol
    def __len__(self):
        return len(self.get_parameter_vector())

    def get_parameter_vector(self) -> np.ndarray:
        raise NotImplementedError

    def set_parameter_vector(self, vector):
        raise NotImplementedError

    def get_parameter_names(self):
        raise NotImplementedError


class ConstantKernel(Kernel):
    def __init__(self, log_constant: float, ndim: int = 1, axes=None):
        if axes is not None:
            raise NotImplementedError("axes subsets are not on the accelerated path")
        self.ndim = int(ndim)
        self.log_constant = float(log_constant)

    def get_parameter_vector(self):
        return np.array([self.log_constant])

    def set_parameter_vector(self, vector):
        (self.log_constant,) = (float(v) for v in vector)

    def get_parameter_names(self):
        return ("log_constant",)


class ExpSquaredKernel(Kernel):
    """Axis-aligned squared-exponential: exp(-0.5 * sum_i (x_i - y_i)^2 / M_i)."""

    def __init__(self, metric, ndim: int = 1, axes=None, **kwargs):
        if axes is not None or kwargs:
            raise NotImplementedError("only a full axis-aligned metric is on the accelerated path")
        self.ndim = int(ndim)
        metric = np.atleast_1d(np.asarray(metric, dtype=np.float64))
        if metric.ndim != 1:
            raise NotImplementedError("general (non axis-aligned) metrics are not on the accelerated path")
        if len(metric) == 1 and self.ndim > 1:
            raise NotImplementedError("isotropic metrics in ndim > 1 are not on the accelerated path")
        if len(metric) != self.ndim:
            raise ValueError("Dimension mismatch")
        if np.any(metric <= 0):
            raise ValueError("metric must be positive")
        self.log_metric = np.log(metric)

    def get_parameter_vector(self):
        return self.log_metric.copy()

    def set_parameter_vector(self, vector):
        vector = np.asarray(vector, dtype=np.float64)
        if vector.shape != self.log_metric.shape:
            raise ValueError("Dimension mismatch")
        self.log_metric = vector.copy()

    def get_parameter_names(self):
        return tuple(f"metric:log_M_{i}_{i}" for i in range(self.ndim))


class Product(Kernel):
    def __init__(self, k1: Kernel, k2: Kernel):
        self.k1, self.k2 = k1, k2
        self.ndim = k1.ndim

    def get_parameter_vector(self):
        return np.concatenate([self.k1.get_parameter_vector(), self.k2.get_parameter_vector()])

    def set_parameter_vector(self, vector):
        n1 = len(self.k1)
        self.k1.set_parameter_vector(vector[:n1])
        self.k2.set_parameter_vector(vector[n1:])

    def get_parameter_names(self):
        return tuple(f"k1:{n}" for n in self.k1.get_parameter_names()) + tuple(
            f"k2:{n}" for n in self.k2.get_parameter_names()
        )


def device_form(kernel: Kernel):
    """
    Validate that ``kernel`` is ``ConstantKernel * ExpSquaredKernel`` and return its kernel
    dimension.  Raises NotImplementedError otherwise.
    """
    if (
        isinstance(kernel, Product)
        and isinstance(kernel.k1, ConstantKernel)
        and isinstance(kernel.k2, ExpSquaredKernel)
        and kernel.k1.ndim == kernel.k2.ndim
    ):
        return kernel.k2.ndim
    raise NotImplementedError(
        "emulator_b200 accelerates `scalar * ExpSquaredKernel(metric=vector, ndim=d)` only "
        "(the kernel every swiftemulator GP emulator builds); other kernels have no device path"
    )
