"""
GP emulator for one scalar per model (no independent axis) -- API of
/root/reference/swiftemulator/emulators/gaussian_process_one_dim.py:23-330.
"""

from __future__ import annotations

from typing import Dict

import attr
import numpy as np

from .base import BaseEmulator, build_gp, default_kernel, flatten_model_values, optimise_gp


@attr.s
class GaussianProcessEmulator1D(BaseEmulator):
    kernel = attr.ib(default=None)
    mean_model = attr.ib(default=None)
    device: Optional[int] = attr.ib(default=None)  # None: this rank's GPU under torchrun, else GPU 0 (dist.resolve_device)

    def _build_arrays(self, model_specification, model_parameters, model_values):
        self.model_specification = model_specification
        self.model_parameters = model_parameters
        self.model_values = model_values
        self.parameter_order = model_specification.parameter_names
        (
            self.independent_variables,
            self.dependent_variables,
            self.dependent_variable_errors,
            self.ordering,
        ) = flatten_model_values(model_specification, model_parameters, model_values, with_independent=False)

    def fit_model(self, model_specification, model_parameters, model_values):
        if self.independent_variables is None:
            self._build_arrays(model_specification, model_parameters, model_values)
        if self.kernel is None:
            self.kernel = default_kernel(model_specification.number_of_parameters)
        gp = build_gp(
            self.kernel, self.mean_model, self.independent_variables, self.dependent_variables,
            self.dependent_variable_errors, device=self.device,
        )
        self.fit_result = optimise_gp(gp, self.dependent_variables)
        self.emulator = gp

    def _query(self, model_parameters):
        if self.emulator is None:
            raise AttributeError("Please train the emulator with fit_model before attempting to make predictions.")
        theta = np.array([model_parameters[name] for name in self.parameter_order])
        # the reference duplicates the row because george wants two points; row 0 is returned
        t = np.empty((2, len(theta)), dtype=np.float32)
        t[0] = theta
        t[1] = theta
        return t

    def predict_values(self, model_parameters: Dict[str, float]):
        t = self._query(model_parameters)
        model, errors = self.emulator.predict(y=self.dependent_variables, t=t, return_cov=False, return_var=True)
        return model[0], errors[0]

    def predict_values_no_error(self, model_parameters: Dict[str, float]):
        t = self._query(model_parameters)
        model = self.emulator.predict(y=self.dependent_variables, t=t, return_cov=False, return_var=False)
        return model[0]
