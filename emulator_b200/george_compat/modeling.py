"""
``george.modeling`` look-alike for the mean-model plumbing
(/root/reference/swiftemulator/mean_models/base.py:5, :80-88).
"""

from __future__ import annotations

import numpy as np


class Model:
    def get_value(self, x):
        raise NotImplementedError

    def __len__(self):
        return 0


class ConstantModel(Model):
    def __init__(self, value: float = 0.0):
        self.value = float(value)

    def get_value(self, x):
        return self.value + np.zeros(len(x))


class CallableModel(Model):
    """Wraps a function f(x) -> mean values; has no parameters of its own."""

    def __init__(self, function, gradient=None):
        self.function = function
        self.gradient = gradient

    def get_value(self, x):
        return self.function(x)
