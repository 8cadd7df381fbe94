"""
Gaussian Process Emulator -- same API and data handling as
/root/reference/swiftemulator/emulators/gaussian_process.py:23-348, with the george calls
going to the CUDA-backed `george_compat.GP`.
"""

from __future__ import annotations

from typing import Dict, Optional

import attr
import numpy as np

from .base import (
    BaseEmulator, build_gp, default_kernel, flatten_model_values, optimise_gp, optimise_gp_with_restarts, query_matrix,
)


@attr.s
class GaussianProcessEmulator(BaseEmulator):
    kernel = attr.ib(default=None)
    mean_model = attr.ib(default=None)
    device: Optional[int] = attr.ib(default=None)  # None: this rank's GPU under torchrun, else GPU 0 (dist.resolve_device)
    #: extra BFGS runs from starts drawn around the initial vector, the best optimum wins (extension; 0 = the
    #: reference's single run).  Under torchrun the starts are shared out over the ranks.
    restarts: int = attr.ib(default=0)
    restart_seed: int = attr.ib(default=0)

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
        ) = flatten_model_values(model_specification, model_parameters, model_values, with_independent=True)

    def fit_model(self, model_specification, model_parameters, model_values):
        if self.independent_variables is None:
            self._build_arrays(model_specification, model_parameters, model_values)
        if self.kernel is None:
            self.kernel = default_kernel(model_specification.number_of_parameters + 1)
        gp = build_gp(
            self.kernel, self.mean_model, self.independent_variables, self.dependent_variables,
            self.dependent_variable_errors, device=self.device,
        )
        if self.restarts > 0:
            self.fit_result = optimise_gp_with_restarts(
                gp, self.kernel, self.mean_model, self.independent_variables, self.dependent_variables,
                self.dependent_variable_errors, self.restarts, seed=self.restart_seed,
            )
        else:
            self.fit_result = optimise_gp(gp, self.dependent_variables)
        self.emulator = gp

    def _require_trained(self):
        if self.emulator is None:
            raise AttributeError("Please train the emulator with fit_model before attempting to make predictions.")

    def predict_values(self, independent: np.array, model_parameters: Dict[str, float]):
        self._require_trained()
        t = query_matrix(independent, model_parameters, self.parameter_order)
        model, errors = self.emulator.predict(y=self.dependent_variables, t=t, return_cov=False, return_var=True)
        return model, errors

    def predict_values_no_error(self, independent: np.array, model_parameters: Dict[str, float]):
        self._require_trained()
        t = query_matrix(independent, model_parameters, self.parameter_order)
        return self.emulator.predict(y=self.dependent_variables, t=t, return_cov=False, return_var=False)

    def predict_grid(self, theta_rows: np.ndarray, independent: np.ndarray, want_var: bool = True):
        """
        Every parameter row of `theta_rows` (columns in `parameter_order`) at every value of `independent`,
        with the query rows formed on the device: mean[n_theta, n_ind] (and variance).  Same values as
        predict_values row by row (not in the reference API; what the sweeps are built on).
        """
        self._require_trained()
        return self.emulator.predict_grid(self.dependent_variables, theta_rows, independent, return_var=want_var)

    def predict_grid_device(self, theta_t, independent_t, mean_out, var_out=None):
        """`predict_grid` on CUDA tensors, results left in device memory (see `GP.predict_grid_device`)."""
        self._require_trained()
        self.emulator.predict_grid_device(self.dependent_variables, theta_t, independent_t, mean_out, var_out)

    @property
    def supports_device_grid(self) -> bool:
        """Whether `predict_grid_device` applies: a zero-mean GP (mean models are host callables)."""
        return self.emulator is not None and self.mean_model is None and hasattr(self.emulator, "predict_grid_device")

    def predict_batch(self, query_rows: np.ndarray, want_var: bool = True):
        """
        Batched prediction at arbitrary rows [independent, theta...] in ONE device call (the
        primitive behind the prediction sweeps; not in the reference API).
        """
        self._require_trained()
        t = np.asarray(query_rows, dtype=np.float32)
        if want_var:
            return self.emulator.predict(y=self.dependent_variables, t=t, return_cov=False, return_var=True)
        return self.emulator.predict(y=self.dependent_variables, t=t, return_cov=False, return_var=False)
