"""
Several GP emulators over (possibly overlapping) regions of the independent axis, blended
linearly in the overlaps -- API of
/root/reference/swiftemulator/emulators/multi_gaussian_process.py:24-468.

Deviations from the reference, all where the reference is broken:
the right-hand error in an overlap comes from the right-hand emulator (the reference reads it
from the left one, :312-314), and `predict_values_no_error` works inside overlaps (the
reference raises NameError there, :459-462).
"""

from __future__ import annotations

from typing import Dict, List, Optional

import attr
import numpy as np

from ..backend import ModelValues
from .base import BaseEmulator
from .gaussian_process import GaussianProcessEmulator


def _inside(independent, low, high):
    mask = np.ones(len(independent), dtype=bool)
    if low is not None:
        mask &= independent > low
    if high is not None:
        mask &= independent < high
    return mask


@attr.s
class MultipleGaussianProcessEmulator(BaseEmulator):
    kernel = attr.ib(default=None)
    mean_model = attr.ib(default=None)
    independent_regions: List[List[Optional[float]]] = attr.ib(default=None)
    device: Optional[int] = attr.ib(default=None)  # None: this rank's GPU under torchrun, else GPU 0 (dist.resolve_device)

    model_values_regions: Optional[List[ModelValues]] = None
    emulators: Optional[List[GaussianProcessEmulator]] = None

    @independent_regions.validator
    def _check_regions(self, attribute, value):
        if value is None:
            return
        for region in value:
            if len(region) != 2:
                raise AttributeError("each independent region must be a [low, high] pair")

    def __attrs_post_init__(self):
        if self.independent_regions is None:
            self.independent_regions = [[None, None]]

    def _build_arrays(self, model_specification, model_parameters, model_values):
        per_region = [{} for _ in self.independent_regions]
        for uid in list(model_values.keys()):
            raw = model_values[uid]
            independent, dependent = raw["independent"], raw["dependent"]
            errors = raw.get("dependent_error", None)
            for index, (low, high) in enumerate(self.independent_regions):
                mask = _inside(independent, low, high)
                entry = {"independent": independent[mask], "dependent": dependent[mask]}
                if errors is not None:
                    entry["dependent_error"] = errors[mask]
                per_region[index][uid] = entry
        self.model_values_regions = [ModelValues(x) for x in per_region]

    def fit_model(self, model_specification, model_parameters, model_values):
        if self.model_values_regions is None:
            self._build_arrays(model_specification, model_parameters, model_values)
        self.model_specification = model_specification
        self.model_parameters = model_parameters
        self.model_values = model_values
        self.emulators = [
            GaussianProcessEmulator(kernel=self.kernel, mean_model=self.mean_model, device=self.device)
            for _ in self.independent_regions
        ]
        for emulator, values in zip(self.emulators, self.model_values_regions):
            emulator.fit_model(model_specification, model_parameters, values)

    def _overlaps(self):
        """{left region index: (low, high)} for consecutive regions that overlap."""
        found = {}
        for index in range(1, len(self.independent_regions)):
            left = self.independent_regions[index][0]
            right = self.independent_regions[index - 1][1]
            if left is None or right is None:
                continue
            if right > left:
                found[index - 1] = (left, right)
        return found

    def _predict(self, independent, model_parameters, with_error: bool):
        if self.emulators is None:
            raise AttributeError("Please train the emulator with fit_model before attempting to make predictions.")
        independent = np.asarray(independent)
        n_regions = len(self.independent_regions)
        values = np.full((n_regions, len(independent)), np.nan)
        errors = np.full((n_regions, len(independent)), np.nan)
        member = np.zeros((n_regions, len(independent)), dtype=bool)
        for index, (low, high) in enumerate(self.independent_regions):
            mask = _inside(independent, low, high)
            member[index] = mask
            if not mask.any():
                continue
            if with_error:
                values[index, mask], errors[index, mask] = self.emulators[index].predict_values(
                    independent[mask], model_parameters
                )
            else:
                values[index, mask] = self.emulators[index].predict_values_no_error(independent[mask], model_parameters)
        overlaps = self._overlaps()
        out = np.empty(len(independent), dtype=np.float64)
        out_err = np.empty(len(independent), dtype=np.float64)
        for i, x in enumerate(independent):
            regions = np.where(member[:, i])[0]
            if len(regions) == 0:
                raise AttributeError(f"independent value {x} lies in none of the regions {self.independent_regions}")
            left = regions[0]
            if len(regions) > 1 and left in overlaps and regions[1] == left + 1:
                low, high = overlaps[left]
                wl, wr = (high - x) / (high - low), (x - low) / (high - low)
                out[i] = values[left, i] * wl + values[left + 1, i] * wr
                if with_error:
                    out_err[i] = np.sqrt(wl * errors[left, i] ** 2 + wr * errors[left + 1, i] ** 2)
            else:
                out[i] = values[left, i]
                if with_error:
                    out_err[i] = errors[left, i]
        return (out, out_err) if with_error else out

    def predict_values(self, independent: np.array, model_parameters: Dict[str, float]):
        return self._predict(independent, model_parameters, with_error=True)

    def predict_values_no_error(self, independent: np.array, model_parameters: Dict[str, float]):
        return self._predict(independent, model_parameters, with_error=False)
