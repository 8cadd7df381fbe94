"""
Leave-one-out cross checks -- API of
/root/reference/swiftemulator/sensitivity/cross_check.py:30-345 (plots excluded) and
cross_check_bins.py:26-113.  The S refits are independent: they are assigned round-robin to
ranks (one process per GPU) and only the fitted hyperparameter vectors are all-gathered, so
every rank ends with a usable emulator for every left-out simulation.

The reference pops the left-out entry from the caller's dict and re-inserts it, which rotates
the row order from refit to refit (:99, :114).  Here each refit sees the original order with
the one simulation removed and the caller's container is never mutated.
"""

from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Hashable, List, Optional

import attr
import numpy as np

from .. import dist
from ..backend import ModelParameters, ModelSpecification, ModelValues
from ..emulators.base import build_gp, default_kernel, optimise_gps_lockstep
from ..emulators.gaussian_process import GaussianProcessEmulator
from ..emulators.gaussian_process_bins import GaussianProcessEmulatorBins


def _without(model_values: ModelValues, uid) -> ModelValues:
    return ModelValues(model_values={k: v for k, v in model_values.model_values.items() if k != uid})


@attr.s
class CrossCheck(object):
    kernel = attr.ib(default=None)
    mean_model = attr.ib(default=None)
    hide_progress: bool = attr.ib(default=True)
    device: Optional[int] = attr.ib(default=None)  # None: this rank's GPU under torchrun, else GPU 0 (dist.resolve_device)
    #: leave-one-out refits of this rank run concurrently from this many host threads (each refit
    #: owns its device handle and stream; the Cholesky's dependency chain of one refit leaves SMs
    #: idle that another refit's kernels use)
    workers: int = attr.ib(default=2)
    #: "scipy": every refit is optimised by scipy's BFGS as in the reference; "lockstep": the refits of a rank
    #: advance together through emulator_b200.optimise (same optima, different iterates; pays for small
    #: problems, where the per-evaluation Python overhead dominates)
    fit_mode: str = attr.ib(default="scipy")
    lockstep_result = None

    model_specification: ModelSpecification = None
    model_parameters: ModelParameters = None
    model_values: ModelValues = None

    leave_out_order: Optional[List[Hashable]] = None
    cross_emulators: Optional[Dict[Hashable, GaussianProcessEmulator]] = None
    cross_fit_parameters: Optional[np.ndarray] = None

    def build_emulators(self, model_specification, model_parameters, model_values):
        self.model_specification = model_specification
        self.model_parameters = model_parameters
        self.model_values = model_values
        self.leave_out_order = list(model_values.model_values.keys())
        count = len(self.leave_out_order)
        n_hyper = model_specification.number_of_parameters + 2 if self.kernel is None else len(self.kernel)

        mine = dist.my_indices(count)
        fitted = np.zeros((count, n_hyper))
        emulators = {}
        def refit(i):
            uid = self.leave_out_order[i]
            emulator = GaussianProcessEmulator(kernel=self.kernel, mean_model=self.mean_model, device=self.device)
            emulator.fit_model(model_specification, model_parameters, _without(model_values, uid))
            return i, uid, emulator

        if self.fit_mode not in ("scipy", "lockstep"):
            raise ValueError("fit_mode must be 'scipy' or 'lockstep'")
        if self.fit_mode == "lockstep" and self.kernel is None and self.mean_model is None and len(mine) > 0:
            # all refits of this rank advance together (emulator_b200/optimise.py): one batched device call
            # per optimiser round.  Every refit has the same shape (one simulation left out).
            done = []
            for i in mine:
                uid = self.leave_out_order[i]
                emulator = GaussianProcessEmulator(device=self.device)
                emulator._build_arrays(model_specification, model_parameters, _without(model_values, uid))
                emulator.kernel = default_kernel(model_specification.number_of_parameters + 1)
                emulator.emulator = build_gp(
                    emulator.kernel, None, emulator.independent_variables, emulator.dependent_variables,
                    emulator.dependent_variable_errors, device=self.device,
                )
                done.append((i, uid, emulator))
            self.lockstep_result = optimise_gps_lockstep(
                [e.emulator for _, _, e in done], [e.dependent_variables for _, _, e in done]
            )
        elif self.kernel is None and self.mean_model is None and self.workers > 1 and len(mine) > 1:
            with ThreadPoolExecutor(max_workers=self.workers) as pool:
                done = list(pool.map(refit, mine))
        else:  # shared kernel / mean-model objects are mutated by fit_model: keep it sequential
            done = [refit(i) for i in mine]
        for i, uid, emulator in done:
            fitted[i] = emulator.emulator.get_parameter_vector()
            emulators[uid] = emulator
        fitted = dist.all_gather_rows(fitted, mine, count)
        self.cross_fit_parameters = fitted
        # refits done on other ranks: rebuild the GP at the gathered optimum (no optimisation)
        for i, uid in enumerate(self.leave_out_order):
            if uid in emulators:
                continue
            emulator = GaussianProcessEmulator(kernel=self.kernel, mean_model=self.mean_model, device=self.device)
            emulator._build_arrays(model_specification, model_parameters, _without(model_values, uid))
            if emulator.kernel is None:
                emulator.kernel = default_kernel(model_specification.number_of_parameters + 1)
            gp = build_gp(
                emulator.kernel, emulator.mean_model, emulator.independent_variables, emulator.dependent_variables,
                emulator.dependent_variable_errors, device=self.device,
            )
            gp.set_parameter_vector(fitted[i])
            emulator.emulator = gp
            emulators[uid] = emulator
        self.cross_emulators = {uid: emulators[uid] for uid in self.leave_out_order}

    def _require_built(self):
        if self.cross_emulators is None:
            raise AttributeError("You need to build the emulators before the prediction step.")

    def build_mocked_model_values_original_independent(self) -> ModelValues:
        self._require_built()
        out = {}
        for uid in self.model_values.keys():
            independent = self.model_values[uid]["independent"]
            emulated, variance = self.cross_emulators[uid].predict_values(
                independent=independent, model_parameters=self.model_parameters[uid]
            )
            out[uid] = {"independent": independent, "dependent": emulated, "dependent_error": np.sqrt(variance)}
        return ModelValues(out)

    def build_mocked_model_values(self, emulate_at: np.array) -> ModelValues:
        self._require_built()
        out = {}
        for uid in self.model_values.keys():
            emulated, variance = self.cross_emulators[uid].predict_values(
                independent=emulate_at, model_parameters=self.model_parameters[uid]
            )
            out[uid] = {"independent": emulate_at, "dependent": emulated, "dependent_error": np.sqrt(variance)}
        return ModelValues(out)

    def get_mean_squared(self, use_dependent_error: bool = False, use_y_as_error: bool = False,
                         use_squared_difference: bool = True):
        self._require_built()
        per_model, everything = {}, []
        for uid in self.cross_emulators.keys():
            x_model = self.model_values.model_values[uid]["independent"]
            y_model = self.model_values.model_values[uid]["dependent"]
            emulated, _ = self.cross_emulators[uid].predict_values(
                independent=x_model, model_parameters=self.model_parameters.model_parameters[uid]
            )
            if use_y_as_error:
                weight = y_model
            else:
                weight = self.model_values.model_values[uid].get("dependent_error", None)
            difference = (y_model - emulated) / weight if use_dependent_error else y_model - emulated
            if use_squared_difference:
                difference = difference ** 2
            per_model[uid] = difference
            everything.extend(difference)
        return np.mean(everything), per_model

    def plot_results(self, *args, **kwargs):
        raise NotImplementedError("plotting (matplotlib) is outside the accelerated hot path")


@attr.s
class CrossCheckBins(object):
    kernel = attr.ib(default=None)
    mean_model = attr.ib(default=None)
    hide_progress: bool = attr.ib(default=True)
    device: Optional[int] = attr.ib(default=None)  # None: this rank's GPU under torchrun, else GPU 0 (dist.resolve_device)

    model_specification: ModelSpecification = None
    model_parameters: ModelParameters = None
    model_values: ModelValues = None

    leave_out_order: Optional[List[Hashable]] = None
    cross_emulators: Optional[Dict[Hashable, GaussianProcessEmulatorBins]] = None

    def build_emulators(self, model_specification, model_parameters, model_values):
        self.model_specification = model_specification
        self.model_parameters = model_parameters
        self.model_values = model_values
        self.leave_out_order = list(model_values.model_values.keys())
        emulators = {}
        for uid in self.leave_out_order:
            # each binned refit shards its own bins over the ranks
            emulator = GaussianProcessEmulatorBins(kernel=self.kernel, mean_model=self.mean_model, device=self.device)
            emulator.fit_model(model_specification, model_parameters, _without(model_values, uid))
            emulators[uid] = emulator
        self.cross_emulators = emulators

    def get_mean_squared(self, use_dependent_error: bool = False, use_y_as_error: bool = False,
                         use_squared_difference: bool = True):
        """Evident intent of cross_check_bins.py:190-249 (broken there against the 2-tuple predict)."""
        if self.cross_emulators is None:
            raise AttributeError("You need to build the emulators before the prediction step.")
        per_model, everything = {}, []
        for uid, emulator in self.cross_emulators.items():
            x_model = self.model_values.model_values[uid]["independent"]
            y_model = self.model_values.model_values[uid]["dependent"]
            valid = np.isin(x_model, emulator.bin_centers)
            emulated, _ = emulator.predict_values(x_model[valid], self.model_parameters.model_parameters[uid])
            weight = y_model[valid] if use_y_as_error else self.model_values.model_values[uid].get(
                "dependent_error", np.ones_like(y_model))[valid]
            difference = (y_model[valid] - emulated) / weight if use_dependent_error else y_model[valid] - emulated
            if use_squared_difference:
                difference = difference ** 2
            per_model[uid] = difference
            everything.extend(difference)
        return np.mean(everything), per_model
