"""
GP emulator whose hyperparameters are the mean of an ensemble-MCMC chain started at the BFGS
optimum -- API of /root/reference/swiftemulator/emulators/gaussian_process_mcmc.py:26-505.

Deviations, all at documented reference defects:
  * burn-in is discarded along the *step* axis (the reference slices the walker axis of
    emcee 3's (steps, walkers, ndim) chain, :267, which with its own defaults selects nothing);
  * the hyperparameter-error loop indexes samples and columns the evident way (the reference
    swaps them, :420-426, and can write out of range);
  * the random generator is seeded through `seed` (the reference uses the unseeded global RNG).
Sampling evaluates the likelihood value only (no gradient): each half-ensemble of proposals is one
batched device call (`GP.log_likelihood_batch`, Cholesky factor alone, factorisations sharing launches).
"""

from __future__ import annotations

from typing import Dict, Optional

import attr
import numpy as np

from .. import dist
from .base import BaseEmulator, build_gp, default_kernel, flatten_model_values, optimise_gp, query_matrix
from .ensemble import EnsembleSampler


@attr.s
class GaussianProcessEmulatorMCMC(BaseEmulator):
    kernel = attr.ib(default=None)
    mean_model = attr.ib(default=None)
    burn_in_steps: int = attr.ib(default=50)
    mcmc_steps: int = attr.ib(default=100)
    walkers: int = attr.ib(default=40)
    use_hyperparameter_error: bool = attr.ib(default=False)
    samples_for_error: int = attr.ib(default=100)
    seed: Optional[int] = attr.ib(default=None)
    #: under torchrun the walkers' likelihood evaluations are split over the ranks (needs a seed: every rank runs the
    #: same sampler); off: every rank evaluates all walkers itself
    shard_walkers: bool = attr.ib(default=True)
    device: Optional[int] = attr.ib(default=None)  # None: this rank's GPU under torchrun, else GPU 0 (dist.resolve_device)

    hyperparameter_samples: Optional[np.ndarray] = None
    hyperparameter_best_fit: Optional[np.ndarray] = None

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
        y = self.dependent_variables
        self.fit_result = optimise_gp(gp, y)
        start = self.fit_result.x

        def log_likelihood(positions):  # a block of walker positions per call (value-only, batched on the device)
            world = dist.world_size()
            if world == 1 or not self.shard_walkers or len(positions) < world:
                return gp.log_likelihood_batch(y, positions, quiet=True)
            # every rank runs the same seeded sampler, evaluates its slice of the proposals, and one all-gather per
            # half-step gives every rank all values (deterministic kernels: bitwise-equal replicas, same decisions)
            start, stop = dist.contiguous_slice(len(positions))
            mine = gp.log_likelihood_batch(y, positions[start:stop], quiet=True)
            return dist.all_gather_concat(mine.reshape(-1, 1), len(positions)).ravel()

        ndim = len(gp)
        seed = self.seed
        if seed is None and self.shard_walkers and dist.world_size() > 1:
            # unseeded run on several ranks: rank 0 draws the seed, everyone uses it (the ranks must propose alike)
            drawn = float(np.random.SeedSequence().entropy % (1 << 52)) if dist.rank() == 0 else 0.0
            seed = int(dist.all_reduce_sum(np.array([drawn]))[0])
        rng = np.random.default_rng(seed)
        gp.eager_gradient = False  # value-only evaluations while sampling
        try:
            sampler = EnsembleSampler(self.walkers, ndim, log_likelihood, seed=rng, vectorize=True)
            p0 = start + 1e-4 * rng.standard_normal((self.walkers, ndim))
            p0, _, _ = sampler.run_mcmc(p0, self.burn_in_steps)
            sampler.run_mcmc(p0, self.mcmc_steps)
        finally:
            gp.eager_gradient = True
        samples = sampler.get_chain()[self.burn_in_steps :, :, :].reshape((-1, ndim))
        result = np.mean(samples, axis=0)
        self.hyperparameter_samples = samples
        self.hyperparameter_best_fit = result
        self._rng = rng
        gp.set_parameter_vector(result)
        self.emulator = gp

    def plot_hyperparameter_distribution(self, filename=None, labels=None):
        raise NotImplementedError("plotting (corner/matplotlib) is outside the accelerated hot path")

    def _require_trained(self):
        if self.emulator is None:
            raise AttributeError("Please train the emulator with fit_model before attempting to make predictions.")

    def predict_values(self, independent: np.array, model_parameters: Dict[str, float]):
        self._require_trained()
        if len(self.hyperparameter_samples[:, 0]) < self.samples_for_error and self.use_hyperparameter_error:
            raise ValueError("Number of subsamples must be less then the total number of samples")
        t = query_matrix(independent, model_parameters, self.parameter_order)
        y = self.dependent_variables
        model, variance = self.emulator.predict(y=y, t=t, return_cov=False, return_var=True)
        if self.use_hyperparameter_error:
            picks = self._rng.choice(len(self.hyperparameter_samples), self.samples_for_error, replace=False)
            if hasattr(self.emulator, "predict_mean_batch"):
                # the samples are factorised in groups on sibling handles (one native call per group)
                spread = self.emulator.predict_mean_batch(y, t, self.hyperparameter_samples[picks, :]).T
            else:
                spread = np.empty((len(independent), self.samples_for_error), dtype=np.float64)
                for column, sample in enumerate(picks):
                    self.emulator.set_parameter_vector(self.hyperparameter_samples[sample, :])
                    spread[:, column] = self.emulator.predict(y=y, t=t, return_cov=False, return_var=False)
                self.emulator.set_parameter_vector(self.hyperparameter_best_fit)
            variance = variance + np.var(spread, axis=1)
        return model, variance

    def predict_values_no_error(self, independent: np.array, model_parameters: Dict[str, float]):
        self._require_trained()
        t = query_matrix(independent, model_parameters, self.parameter_order)
        return self.emulator.predict(y=self.dependent_variables, t=t, return_cov=False, return_var=False)
