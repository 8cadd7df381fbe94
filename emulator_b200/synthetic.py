"""
Seeded synthetic data for BASELINE.json's configurations: parameters
theta ~ U[0,1]^n_par rescaled to the limits, bins x = linspace(0, 1, B) shared by all sims,
a smooth non-linear response in the style of the reference docs' log-Schechter mock
(/root/reference/docs/source/getting_started/index.rst:64-67, :127-134) generalised to
n_par dimensions, and 2 % errors floored at 1e-3.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np

from .backend import ModelParameters, ModelSpecification, ModelValues

CONFIGS = {
    # name: (n_par, sims, bins)
    "C1": (2, 50, 20),
    "C2": (8, 256, 16),
    "C3": (8, 256, 64),
    "C4": (8, 512, 16),
}


def response(x: np.ndarray, theta: np.ndarray) -> np.ndarray:
    """Smooth non-linear f(x, theta): a Schechter-like knee whose position/slope move with theta."""
    n_par = len(theta)
    w = np.cos(0.7 * np.arange(1, n_par + 1))
    knee = 0.5 + 0.3 * np.tanh(np.dot(w, theta - 0.5))
    slope = -1.0 - 0.5 * theta[0] + 0.25 * np.sin(3.0 * theta[-1])
    norm = 1.0 + 0.5 * np.sum((theta - 0.5) ** 2)
    return norm + slope * (x - knee) - np.exp(2.0 * (x - knee)) + 0.2 * np.sin(4.0 * x + theta[n_par // 2])


def make_model(n_par: int, sims: int, bins: int, seed: int = 1234, error_fraction: float = 0.02,
               with_errors: bool = True) -> Tuple[ModelSpecification, ModelParameters, ModelValues]:
    rng = np.random.default_rng(seed)
    names = [f"p{i}" for i in range(n_par)]
    limits = [[0.0, 1.0] for _ in range(n_par)]
    spec = ModelSpecification(number_of_parameters=n_par, parameter_names=names, parameter_limits=limits)
    x = np.linspace(0.0, 1.0, bins)
    thetas = rng.random((sims, n_par))
    parameters, values = {}, {}
    for uid in range(sims):
        theta = thetas[uid]
        parameters[uid] = {name: float(theta[i]) for i, name in enumerate(names)}
        y = response(x, theta)
        entry = {"independent": x, "dependent": y}
        if with_errors:
            entry["dependent_error"] = np.maximum(error_fraction * np.abs(y), 1e-3)
        values[uid] = entry
    return spec, ModelParameters(model_parameters=parameters), ModelValues(model_values=values)


def make_config(name: str, seed_offset: int = 0):
    n_par, sims, bins = CONFIGS[name]
    return make_model(n_par, sims, bins, seed=1234 + int(name[1:]) + seed_offset)


def design_arrays(n: int, d: int, seed: int = 0, error_fraction: float = 0.02):
    """
    Flat (X [n, d], y, yerr) arrays, float32-rounded like `_build_arrays` does
    (/root/reference/swiftemulator/emulators/gaussian_process.py:89-95), for kernel-level
    parity tests and the bench: d = n_par + 1, column 0 is the independent variable.
    """
    n_par = d - 1
    bins = 16 if n % 16 == 0 else 1
    sims = n // bins
    rng = np.random.default_rng(seed)
    x = np.empty((n, d), dtype=np.float32)
    y = np.empty(n, dtype=np.float32)
    xs = np.linspace(0.0, 1.0, bins) if bins > 1 else np.array([0.5])
    thetas = rng.random((sims, max(n_par, 1)))
    row = 0
    for s in range(sims):
        th = thetas[s][:n_par] if n_par > 0 else np.zeros(0)
        for b in range(bins):
            x[row, 0] = xs[b] if bins > 1 else rng.random()
            if n_par > 0:
                x[row, 1:] = th
            y[row] = response(np.float64(x[row, 0]), th if n_par > 0 else np.array([0.5]))
            row += 1
    yerr = np.maximum(error_fraction * np.abs(y), 1e-3).astype(np.float32)
    return x, y, yerr


def perturbed_parameter_vectors(d: int, count: int = 32, seed: int = 7, scale: float = 0.5) -> np.ndarray:
    """x0 followed by `count` vectors drawn N(x0, scale^2) -- the LML+grad timing set."""
    x0 = np.concatenate([[np.log(1.0 / d)], np.zeros(d)])
    rng = np.random.default_rng(seed)
    return np.vstack([x0, x0 + scale * rng.standard_normal((count, d + 1))])
