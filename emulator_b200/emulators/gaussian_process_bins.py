"""
One independent GP per unique `independent` value (bin), kernel over the model parameters
only -- API and data handling of
/root/reference/swiftemulator/emulators/gaussian_process_bins.py:24-462.  The bins are
independent problems: they are sharded round-robin over ranks (emulator_b200.dist) and, on
one GPU, each owns its own device handle and stream.
"""

from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional

import attr
import numpy as np

from .. import dist
from .base import BaseEmulator, build_gp, default_kernel, optimise_gp, optimise_gps_lockstep


@attr.s
class GaussianProcessEmulatorBins(BaseEmulator):
    kernel = attr.ib(default=None)
    mean_model = attr.ib(default=None)
    device: Optional[int] = attr.ib(default=None)  # None: this rank's GPU under torchrun, else GPU 0 (dist.resolve_device)
    #: independent per-bin fits run concurrently from this many host threads; every GP owns its
    #: device handle and stream, and the C ABI calls release the GIL (results do not depend on it)
    workers: int = attr.ib(default=8)
    #: "scipy": every bin is fitted by scipy's BFGS, as in the reference (from `workers` host threads);
    #: "lockstep": all bins advance together through emulator_b200.optimise (one batched device call per
    #: round; a different line search, so iterates differ while optima agree to the optimiser's tolerance)
    fit_mode: str = attr.ib(default="scipy")
    lockstep_result = None

    n_bins: int = None
    bin_centers: List[float] = None
    bin_model_values: List[Dict[str, np.array]] = None
    bin_gaussian_process: List = None
    bin_fit_parameters: Optional[np.ndarray] = None

    def _build_arrays(self, model_specification, model_parameters, model_values):
        self.model_specification = model_specification
        self.model_parameters = model_parameters
        self.model_values = model_values
        self.parameter_order = model_specification.parameter_names
        names = self.parameter_order
        n_par = model_specification.number_of_parameters
        identifiers = list(model_values.model_values.keys())

        centers = np.unique([v for uid in identifiers for v in model_values.model_values[uid]["independent"]])
        self.n_bins = len(centers)
        self.bin_centers = centers

        per_bin = []
        for center in centers:
            thetas, dependent, errors = [], [], []
            for uid in identifiers:
                relation = model_values.model_values[uid]
                independent = relation["independent"]
                hit = independent == center  # exact float match, as in the reference (:121)
                if not hit.any():
                    continue
                dependent.append(relation["dependent"][hit][0])
                errors.append(relation.get("dependent_error", np.zeros(len(independent)))[hit][0])
                thetas.append([model_parameters.model_parameters[uid][name] for name in names])
                if np.ndim(errors) != 1:
                    raise AttributeError("Multiple dimensional errors are not currently supported in GPE mode")
            per_bin.append(
                {
                    "independent": np.array(thetas).reshape((len(dependent), n_par)),  # float64 (:146-156)
                    "dependent": np.array(dependent).flatten(),
                    "dependent_errors": np.array(errors).flatten(),
                }
            )
        self.bin_model_values = per_bin

    def fit_model(self, model_specification, model_parameters, model_values):
        if self.bin_model_values is None:
            self._build_arrays(model_specification, model_parameters, model_values)
        if self.kernel is None:
            self.kernel = default_kernel(self.model_specification.number_of_parameters)

        n_hyper = len(self.kernel)
        mine = dist.my_indices(self.n_bins)
        gps = [None] * self.n_bins
        fitted = np.zeros((self.n_bins, n_hyper))
        def fit_bin(b):
            values = self.bin_model_values[b]
            gp = build_gp(
                self.kernel, self.mean_model, values["independent"], values["dependent"], values["dependent_errors"],
                device=self.device,
            )
            optimise_gp(gp, values["dependent"])
            return b, gp

        if self.fit_mode not in ("scipy", "lockstep"):
            raise ValueError("fit_mode must be 'scipy' or 'lockstep'")
        if self.fit_mode == "lockstep" and self.mean_model is None and len(mine) > 0:
            # all bins of this rank advance together: one batched device call per optimiser round
            mine_gps = [
                build_gp(self.kernel, None, self.bin_model_values[b]["independent"], self.bin_model_values[b]["dependent"],
                         self.bin_model_values[b]["dependent_errors"], device=self.device)
                for b in mine
            ]
            self.lockstep_result = optimise_gps_lockstep(mine_gps, [self.bin_model_values[b]["dependent"] for b in mine])
            results = list(zip(mine, mine_gps))
        elif self.mean_model is not None or self.workers <= 1 or len(mine) <= 1:
            # mean_model.train mutates the shared mean-model object: keep the reference's sequential order
            results = [fit_bin(b) for b in mine]
        else:
            with ThreadPoolExecutor(max_workers=self.workers) as pool:
                results = list(pool.map(fit_bin, mine))
        for b, gp in results:
            gps[b] = gp
            fitted[b] = gp.get_parameter_vector()
        # collect every bin's optimum on every rank; bins fitted elsewhere get a GP at that optimum
        fitted = dist.all_gather_rows(fitted, mine, self.n_bins)
        for b in range(self.n_bins):
            if gps[b] is None:
                values = self.bin_model_values[b]
                gp = build_gp(
                    self.kernel, self.mean_model, values["independent"], values["dependent"],
                    values["dependent_errors"], device=self.device,
                )
                gp.set_parameter_vector(fitted[b])
                gps[b] = gp
        self.bin_fit_parameters = fitted
        self.bin_gaussian_process = gps

    def _bin_indices(self, independent):
        if self.bin_gaussian_process is None:
            raise AttributeError("Please train the emulator with fit_model before attempting to make predictions.")
        centers = np.array(self.bin_centers)
        order = []
        for value in independent:
            where = np.where(centers == value)[0]
            if len(where) == 0:
                raise AttributeError(
                    f"Requested independent variable {independent} not valid, ",
                    f"this instance of GPE Bins is only valid at {centers}.",
                )
            order.append(where[0])
        return order

    def _query(self, model_parameters):
        theta = np.array([model_parameters[name] for name in self.parameter_order])
        # the reference pads with 0.98 theta and 1.02 theta "so george predicts > 1 point" and
        # keeps row 1 (:338-345, :361-363); rows are independent, so one row is enough here
        return theta.reshape(1, -1).astype(np.float64)

    def predict_values(self, independent: np.array, model_parameters: Dict[str, float]):
        order = self._bin_indices(independent)
        t = self._query(model_parameters)
        means, variances = [], []
        for b in order:
            mu, var = self.bin_gaussian_process[b].predict(
                y=self.bin_model_values[b]["dependent"], t=t, return_cov=False, return_var=True
            )
            means.append(mu[0])
            variances.append(var[0])
        return np.array(means), np.array(variances)

    def predict_grid(self, theta_rows: np.ndarray, independent: np.ndarray, want_var: bool = True):
        """
        Every parameter row of `theta_rows` (columns in `parameter_order`) at every requested bin: one
        batched device call per bin instead of one per (parameter row, bin) -- what the sweeps
        (mocking.mock_hypercube / mock_sweep) would otherwise issue through predict_values.  Returns
        mean[n_theta, n_bins] (and variance); same values as predict_values row by row.
        """
        order = self._bin_indices(independent)
        theta = np.ascontiguousarray(np.atleast_2d(theta_rows), dtype=np.float64)
        mean = np.empty((len(theta), len(order)))
        var = np.empty((len(theta), len(order))) if want_var else None
        for col, b in enumerate(order):
            out = self.bin_gaussian_process[b].predict(
                y=self.bin_model_values[b]["dependent"], t=theta, return_cov=False, return_var=want_var
            )
            if want_var:
                mean[:, col], var[:, col] = out
            else:
                mean[:, col] = out
        return (mean, var) if want_var else mean

    def predict_values_no_error(self, independent: np.array, model_parameters: Dict[str, float]):
        order = self._bin_indices(independent)
        t = self._query(model_parameters)
        means = []
        for b in order:
            mu = self.bin_gaussian_process[b].predict(
                y=self.bin_model_values[b]["dependent"], t=t, return_cov=False, return_var=False
            )
            means.append(mu[0])
        return np.array(means)
