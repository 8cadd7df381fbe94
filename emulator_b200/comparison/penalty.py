"""
Penalty functions over emulator outputs -- the calculators of
/root/reference/swiftemulator/comparison/penalty.py:19-1100 (plots excluded), the step that follows the hot path
in a calibration loop: a prediction sweep produces mean[n_theta, n_ind], a penalty
calculator turns every row into a number between 0 and 1 against an observational data set.

Same classes, constructor parameters, `register_observation` / `penalty` / `penalties` methods and formulas as
the reference.  Two things differ:

  * the reference takes a `velociraptor.observations.ObservationalData` with `unyt` quantities; neither package
    is needed here: any object with `.x`, `.y` (and `.y_scatter` for the Gaussian calculators) works -- unyt
    arrays are converted with `.to(units)` when units are given, plain arrays are taken as they are (`Observation`
    below is such a plain holder);
  * `penalties_grid(independent, dependent[n_theta, n_ind])` evaluates ALL rows of a sweep in one vectorised
    pass -- the reference's `penalties()` loops over a ModelValues container one model at a time (:149-197).

Every calculator reduces to an element-wise function of (independent, dependent) on the valid points
`lower <= independent < upper`; `penalty` returns those per-point penalties (0.0 when no point is valid, as the
reference does), capped at 1.
"""

from __future__ import annotations

from typing import Callable, Dict, Hashable, Optional

import attr
import numpy as np


@attr.s
class Observation(object):
    """Plain stand-in for velociraptor's ObservationalData: x, y and (optionally) y_scatter ([n] or [2, n])."""

    x = attr.ib()
    y = attr.ib()
    y_scatter = attr.ib(default=None)


def _plain(values, units=None) -> np.ndarray:
    """unyt array -> numbers in `units`; anything else -> float array."""
    if units is not None and hasattr(values, "to"):
        values = values.to(units)
    return np.asarray(getattr(values, "value", values), dtype=np.float64)


def _number(value, units=None, convert=True) -> float:
    if convert and units is not None and hasattr(value, "to"):
        value = value.to(units)
    return float(getattr(value, "value", value))


class _Interpolator:
    """Linear interpolation clamped to the end values: what the reference builds with
    `interp1d(kind="linear", bounds_error=False, fill_value=(y[x.argmin()], y[x.argmax()]))`."""

    def __init__(self, x, y):
        order = np.argsort(x, kind="stable")
        self.x, self.y = np.asarray(x, dtype=np.float64)[order], np.asarray(y, dtype=np.float64)[order]

    def __call__(self, at):
        return np.interp(at, self.x, self.y)


@attr.s
class PenaltyCalculator(object):
    observation = None
    interpolator_values = None
    interpolator_error = None
    log_independent: bool = False
    log_dependent: bool = False
    independent_units = None
    dependent_units = None

    # ---- reference API ----------------------------------------------------------------------------
    def register_observation(self, observation, log_independent: bool, log_dependent: bool, independent_units=None,
                             dependent_units=None) -> None:
        self.observation = observation
        self.log_independent = log_independent
        self.log_dependent = log_dependent
        self.independent_units = independent_units
        self.dependent_units = dependent_units
        self.observation_interpolation()

    def _observed_xy(self):
        x = _plain(self.observation.x, self.independent_units)
        y = _plain(self.observation.y, self.dependent_units)
        if self.log_independent:
            x = np.log10(x)
        if self.log_dependent:
            y = np.log10(y)
        return x, y

    def observation_interpolation(self):
        x, y = self._observed_xy()
        self.interpolator_values = _Interpolator(x, y)
        self.interpolator_error = None
        self._convert_limits()

    def _convert_limits(self):
        """Limits and offsets given with units are taken in the model's units (in log space they are plain numbers)."""
        for name in ("lower", "upper", "offset_transition"):
            if hasattr(self, name):
                setattr(self, "_" + name, _number(getattr(self, name), self.independent_units, convert=not self.log_independent))
        for name in ("offset", "offset_lower", "offset_upper", "offset_squeeze", "offset_normal"):
            if hasattr(self, name):
                setattr(self, "_" + name, _number(getattr(self, name), self.dependent_units, convert=not self.log_dependent))
        if hasattr(self, "transition_width"):
            self._transition_width = _number(self.transition_width, self.independent_units, convert=not self.log_independent)

    def _point_penalties(self, independent, dependent):
        """Element-wise penalties before the cap at 1 (arrays of equal, broadcastable shape)."""
        raise NotImplementedError

    def _valid(self, independent):
        return np.logical_and(independent >= self._lower, independent < self._upper)

    def penalty(self, independent: np.ndarray, dependent: np.ndarray, dependent_error: Optional[np.ndarray] = None):
        independent, dependent = np.asarray(independent, dtype=np.float64), np.asarray(dependent, dtype=np.float64)
        mask = self._valid(independent)
        if mask.sum() == 0:
            return 0.0
        return np.minimum(self._point_penalties(independent[mask], dependent[mask]), 1.0)

    def penalties(self, model, collate_with: Callable) -> Dict[Hashable, float]:
        out = {}
        for unique_id, data in model.model_values.items():
            out[unique_id] = collate_with(
                self.penalty(independent=data["independent"], dependent=data["dependent"],
                             dependent_error=data.get("dependent_error", None))
            )
        return out

    # ---- sweeps -----------------------------------------------------------------------------------------
    def penalties_grid(self, independent: np.ndarray, dependent: np.ndarray, collate_with: Callable = np.mean) -> np.ndarray:
        """
        The penalty of every row of a sweep at once: `independent` [n_ind] are the values every row was predicted
        at, `dependent` [n_theta, n_ind] the predictions (`mock_sweep`, `mock_hypercube`, `sobol.sweep_predict`).
        Returns [n_theta]: `collate_with(per-point penalties, axis=-1)` over the valid points, 0 where none is valid
        -- the same numbers as `penalties()` over the equivalent ModelValues container, in one vectorised pass.
        """
        independent = np.asarray(independent, dtype=np.float64)
        dependent = np.atleast_2d(np.asarray(dependent, dtype=np.float64))
        mask = self._valid(independent)
        if mask.sum() == 0:
            return np.zeros(dependent.shape[0])
        points = np.minimum(self._point_penalties(independent[mask][None, :], dependent[:, mask]), 1.0)
        return np.asarray(collate_with(points, axis=-1), dtype=np.float64)

    def plot_penalty(self, *args, **kwargs):
        raise NotImplementedError("plotting (matplotlib) is outside the accelerated hot path")


@attr.s
class L1PenaltyCalculator(PenaltyCalculator):
    """|observed - model| / offset, capped at 1 (reference :268-331)."""

    offset = attr.ib()
    lower = attr.ib()
    upper = attr.ib()

    def _point_penalties(self, independent, dependent):
        return np.abs(self.interpolator_values(independent) - dependent) / self._offset


@attr.s
class L1PenaltyCalculatorOneSided(PenaltyCalculator):
    """L1, with the maximal penalty on one side of the data (reference :334-416)."""

    offset = attr.ib()
    maximum_penalty: str = attr.ib()
    lower = attr.ib()
    upper = attr.ib()

    @maximum_penalty.validator
    def _check_maximum_penalty(self, attribute, value):
        if value not in ("above", "below"):
            raise AttributeError("maximum_penalty must be one of above or below.")

    def _point_penalties(self, independent, dependent):
        observed = self.interpolator_values(independent)
        penalties = np.abs(observed - dependent) / self._offset
        side = dependent > observed if self.maximum_penalty == "above" else dependent < observed
        return np.where(side, 1.0, penalties)


@attr.s
class L2PenaltyCalculator(PenaltyCalculator):
    """(observed - model)^2 / offset^2, capped at 1 (reference :419-482)."""

    offset = attr.ib()
    lower = attr.ib()
    upper = attr.ib()

    def _point_penalties(self, independent, dependent):
        return (self.interpolator_values(independent) - dependent) ** 2 / self._offset ** 2


def _one_sided_offsets(offsets, dependent, observed, ratio):
    return np.where(dependent < observed, offsets * ratio, offsets)


@attr.s
class L1VariablePenaltyCalculator(PenaltyCalculator):
    """L1 whose offset moves from offset_lower to offset_upper through a logistic transition (reference :485-589)."""

    offset_lower = attr.ib()
    offset_upper = attr.ib()
    offset_transition = attr.ib()
    transition_width = attr.ib()
    lower = attr.ib()
    upper = attr.ib()
    offset_below_above_ratio: float = attr.ib(default=1.0)

    def _point_penalties(self, independent, dependent):
        offsets = self._offset_lower + (self._offset_upper - self._offset_lower) * (
            1.0 / (1.0 + np.exp((self._offset_transition - independent) / self._transition_width))
        )
        observed = self.interpolator_values(independent)
        offsets = _one_sided_offsets(np.broadcast_to(offsets, np.broadcast(offsets, dependent).shape), dependent, observed,
                                     self.offset_below_above_ratio)
        return np.abs(observed - dependent) / offsets


@attr.s
class L1SqueezePenaltyCalculator(PenaltyCalculator):
    """L1 whose offset is squeezed to offset_squeeze in a Gaussian window around offset_transition (reference :592-698)."""

    offset_squeeze = attr.ib()
    offset_normal = attr.ib()
    offset_transition = attr.ib()
    transition_width = attr.ib()
    lower = attr.ib()
    upper = attr.ib()
    offset_below_above_ratio: float = attr.ib(default=1.0)

    def _point_penalties(self, independent, dependent):
        offsets = self._offset_normal - (self._offset_normal - self._offset_squeeze) * np.exp(
            -0.5 * ((independent - self._offset_transition) / self._transition_width) ** 2
        )
        observed = self.interpolator_values(independent)
        offsets = _one_sided_offsets(np.broadcast_to(offsets, np.broadcast(offsets, dependent).shape), dependent, observed,
                                     self.offset_below_above_ratio)
        return np.abs(observed - dependent) / offsets


class _GaussianBase(PenaltyCalculator):
    """1 - exp(-chi^2 / 2) with everything beyond sigma_max set to 1."""

    uses_data_errors = False

    def observation_interpolation(self):
        super().observation_interpolation()
        if not self.uses_data_errors:
            return
        x = _plain(self.observation.x, self.independent_units)
        scatter = _plain(self.observation.y_scatter, self.dependent_units)
        if self.log_independent:
            x = np.log10(x)
            # (the reference propagates the errors to log space under the same switch, :737-741)
            y = _plain(self.observation.y, self.dependent_units)
            scatter = np.abs(scatter / (y * np.log(10)))
        if scatter.ndim > 1:  # unsymmetric errors: their mean
            scatter = np.mean(scatter, axis=0)
        self.error_interpolator_values = _Interpolator(x, scatter)

    def _variance(self, independent, observed):
        raise NotImplementedError

    def _point_penalties(self, independent, dependent):
        observed = self.interpolator_values(independent)
        penalties = 1.0 - np.exp(-0.5 * ((dependent - observed) ** 2 / self._variance(independent, observed)))
        cap = 1.0 - np.exp(-0.5 * (float(self.sigma_max) ** 2))
        return np.where(penalties > cap, 1.0, penalties)


@attr.s
class GaussianDataErrorsPenaltyCalculator(_GaussianBase):
    """Gaussian around the data with the data's own errors (reference :701-799)."""

    sigma_max = attr.ib()
    lower = attr.ib()
    upper = attr.ib()
    uses_data_errors = True

    def _variance(self, independent, observed):
        return self.error_interpolator_values(independent) ** 2


@attr.s
class GaussianPercentErrorsPenaltyCalculator(_GaussianBase):
    """Gaussian around the data with a fixed percentage error (reference :802-873)."""

    percent_error: float = attr.ib()
    sigma_max: float = attr.ib()
    lower = attr.ib()
    upper = attr.ib()

    def _variance(self, independent, observed):
        return (observed * (self.percent_error / 100)) ** 2


@attr.s
class GaussianDataErrorsPercentFloorPenaltyCalculator(_GaussianBase):
    """Data errors with a percentage floor added in quadrature (reference :876-983)."""

    percent_error: float = attr.ib()
    sigma_max: float = attr.ib()
    lower = attr.ib()
    upper = attr.ib()
    uses_data_errors = True

    def _variance(self, independent, observed):
        return self.error_interpolator_values(independent) ** 2 + (observed * (self.percent_error / 100)) ** 2


@attr.s
class GaussianWeightedDataErrorsPercentFloorPenaltyCalculator(_GaussianBase):
    """As above, carrying a weight for the caller's collation (the reference stores it and does not use it in
    `penalty`, :986-1100)."""

    percent_error: float = attr.ib()
    sigma_max: float = attr.ib()
    weight: float = attr.ib()
    lower = attr.ib()
    upper = attr.ib()
    uses_data_errors = True

    def _variance(self, independent, observed):
        return self.error_interpolator_values(independent) ** 2 + (observed * (self.percent_error / 100)) ** 2
