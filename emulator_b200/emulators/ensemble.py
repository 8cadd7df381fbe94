"""
Affine-invariant ensemble sampler (Goodman & Weare 2010 stretch move) with the structure of
``emcee.EnsembleSampler`` 3.x, which the reference drives at
/root/reference/swiftemulator/emulators/gaussian_process_mcmc.py:259-267 but which is not
installed in this image.  One step updates the two half-ensembles in turn (emcee's red-blue
split), so every step issues two batches of ``walkers / 2`` value-only likelihood
evaluations.  Seeded explicitly (the reference uses the unseeded global numpy RNG, :262).
"""

from __future__ import annotations

import numpy as np


class EnsembleSampler:
    def __init__(self, nwalkers: int, ndim: int, log_prob_fn, a: float = 2.0, seed=None, vectorize: bool = False):
        if nwalkers < 2 or nwalkers % 2:
            raise ValueError("the stretch move needs an even number (>= 2) of walkers")
        self.nwalkers, self.ndim, self.log_prob_fn, self.a = nwalkers, ndim, log_prob_fn, a
        #: emcee's option of the same name: log_prob_fn takes the whole (n, ndim) block of positions
        self.vectorize = vectorize
        self.rng = np.random.default_rng(seed)
        self._chain = []
        self._log_prob = []
        self.naccepted = np.zeros(nwalkers)
        self._last = None

    def _evaluate(self, coords):
        if self.vectorize:
            values = np.asarray(self.log_prob_fn(coords), dtype=np.float64)
            if values.shape != (len(coords),):
                raise ValueError("a vectorized probability function must return one value per position")
        else:
            values = np.array([self.log_prob_fn(c) for c in coords], dtype=np.float64)
        if np.any(np.isnan(values)):
            raise ValueError("Probability function returned NaN")
        return values

    def run_mcmc(self, initial_state, nsteps: int):
        coords = np.array(initial_state, dtype=np.float64)
        if coords.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions")
        if self._last is not None and np.array_equal(coords, self._last[0]):
            log_prob = self._last[1].copy()
        else:
            log_prob = self._evaluate(coords)
        if not np.all(np.isfinite(log_prob[~np.isneginf(log_prob)])):
            raise ValueError("The initial log probability is not finite")
        half = self.nwalkers // 2
        for _ in range(nsteps):
            for split in (0, 1):
                first = np.arange(half) + split * half
                second = np.arange(half) + (1 - split) * half
                # z ~ g(z) proportional to 1/sqrt(z) on [1/a, a]
                zz = ((self.a - 1.0) * self.rng.random(half) + 1.0) ** 2 / self.a
                partners = coords[second][self.rng.integers(half, size=half)]
                proposal = partners - (partners - coords[first]) * zz[:, None]
                new_log_prob = self._evaluate(proposal)
                log_accept = (self.ndim - 1.0) * np.log(zz) + new_log_prob - log_prob[first]
                accept = np.log(self.rng.random(half)) < log_accept
                idx = first[accept]
                coords[idx] = proposal[accept]
                log_prob[idx] = new_log_prob[accept]
                self.naccepted[idx] += 1
            self._chain.append(coords.copy())
            self._log_prob.append(log_prob.copy())
        self._last = (coords.copy(), log_prob.copy())
        return coords.copy(), log_prob.copy(), None

    def get_chain(self, flat: bool = False, discard: int = 0):
        chain = np.array(self._chain)[discard:]  # (nsteps, nwalkers, ndim), emcee 3 layout
        return chain.reshape(-1, self.ndim) if flat else chain

    def get_log_prob(self):
        return np.array(self._log_prob)
