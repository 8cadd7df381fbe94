"""
Mean models with the reference's protocol (/root/reference/swiftemulator/mean_models/):
``train(independent, dependent)``, ``predict(independent)``, ``george_model`` (a
``CallableModel`` around a frozen copy), ``copy()``.  The GP sees the residual y - mean(X) and
the host adds mean(t) back after the device prediction.
"""

from copy import deepcopy
from typing import Optional

import attr
import numpy as np
import sklearn.linear_model as lm
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import PolynomialFeatures

from ..george_compat.modeling import CallableModel


@attr.s
class MeanModel(object):
    """Protocol of mean_models/base.py:13-97."""

    def train(self, independent: np.ndarray, dependent: np.ndarray) -> None:
        raise NotImplementedError

    def predict(self, independent: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    @property
    def george_model(self) -> CallableModel:
        # a frozen copy: later re-training of this object must not change an existing GP
        return CallableModel(self.copy().predict)

    def copy(self) -> "MeanModel":
        return deepcopy(self)


@attr.s
class LinearMeanModel(MeanModel):
    """mean_models/linear.py:16: LinearRegression, or Lasso when lasso_model_alpha != 0."""

    lasso_model_alpha: float = attr.ib(default=0.0)
    model: Optional[object] = None

    def train(self, independent, dependent):
        self.model = (
            lm.LinearRegression(fit_intercept=True)
            if self.lasso_model_alpha == 0.0
            else lm.Lasso(alpha=self.lasso_model_alpha)
        )
        self.model.fit(independent, dependent)

    def predict(self, independent):
        return self.model.predict(independent)


@attr.s
class PolynomialMeanModel(MeanModel):
    """mean_models/polynomial.py:19: polynomial features of `degree` + linear regression."""

    degree: int = attr.ib(default=1)
    model: Optional[object] = None

    def train(self, independent, dependent):
        self.model = Pipeline(
            [("poly", PolynomialFeatures(degree=self.degree)), ("linear", lm.LinearRegression(fit_intercept=True))]
        )
        self.model.fit(independent, dependent)

    def predict(self, independent):
        return self.model.predict(independent)


@attr.s
class OffsetMeanModel(MeanModel):
    """mean_models/offset.py:14: the mean of the training values, in the inputs' dtype."""

    model: Optional[float] = None

    def train(self, independent, dependent):
        self.model = np.mean(dependent)

    def predict(self, independent):
        return np.ones(independent.shape[0], dtype=independent.dtype) * self.model


@attr.s
class FixedMeanModel(MeanModel):
    """mean_models/fixed.py:14: a fixed, untrained constant."""

    model: float = attr.ib(default=1.0)

    def train(self, independent, dependent):
        return

    def predict(self, independent):
        return np.ones(independent.shape[0], dtype=independent.dtype) * self.model
