"""
Sobol sensitivity sweep on top of batched prediction (BASELINE.json configuration 5; an
extension: the reference only has SALib RBD-FAST on raw data,
/root/reference/swiftemulator/sensitivity/basic.py:23-104, and SALib is absent here).

Saltelli design from a scrambled Sobol sequence (scipy.stats.qmc) over
`ModelSpecification.salib_problem["bounds"]` (backend/model_specification.py:71-83):
matrices A, B and the n_par matrices AB_i; every parameter row is predicted (mean and
variance) at every requested independent value in batched device calls; first-order and
total-order indices per independent value use the Saltelli (2010) / Jansen estimators.
Rows are split contiguously over ranks and all-gathered once.
"""

from __future__ import annotations

from typing import Dict

import numpy as np
from scipy.stats import qmc

from .. import dist
from ..backend import ModelSpecification


def saltelli_design(model_specification: ModelSpecification, base_samples: int, seed: int = 11) -> np.ndarray:
    """Rows [A; B; AB_0; ...; AB_{k-1}], each block `base_samples` x n_par, scaled to the bounds."""
    k = model_specification.number_of_parameters
    sampler = qmc.Sobol(d=2 * k, scramble=True, seed=seed)
    unit = sampler.random(base_samples)
    bounds = np.array(model_specification.salib_problem["bounds"], dtype=np.float64)
    lo, hi = bounds[:, 0], bounds[:, 1]
    a = lo + unit[:, :k] * (hi - lo)
    b = lo + unit[:, k:] * (hi - lo)
    blocks = [a, b]
    for i in range(k):
        ab = a.copy()
        ab[:, i] = b[:, i]
        blocks.append(ab)
    return np.vstack(blocks)


def _sweep_predict_device(emulator, theta_rows, independent, want_var):
    """The sweep with the results kept in HBM until they are gathered: this rank's parameter rows go to the device
    once, the (theta x independent) query rows are formed there, mean and variance are written to device buffers,
    all-gathered over NVLink (NCCL) and copied to the host once."""
    import torch

    n_theta, n_ind = len(theta_rows), len(independent)
    world = dist.world_size()
    per_rank = (n_theta + world - 1) // world
    start, stop = dist.contiguous_slice(n_theta)
    dev = torch.device("cuda", emulator.emulator.device)
    with torch.cuda.device(dev):
        theta_t = torch.from_numpy(np.ascontiguousarray(theta_rows[start:stop], dtype=np.float64)).to(dev)
        ind_t = torch.from_numpy(np.ascontiguousarray(independent, dtype=np.float64)).to(dev)
        mean_t = torch.zeros((per_rank, n_ind), dtype=torch.float64, device=dev)
        var_t = torch.zeros((per_rank, n_ind), dtype=torch.float64, device=dev) if want_var else None
        if stop > start:
            emulator.predict_grid_device(theta_t, ind_t, mean_t[: stop - start], var_t[: stop - start] if want_var else None)
        mean = dist.all_gather_concat_device(mean_t, n_theta)
        if want_var:
            return mean, dist.all_gather_concat_device(var_t, n_theta)
        return mean


def sweep_predict(emulator, theta_rows: np.ndarray, independent: np.ndarray, want_var: bool = True,
                  chunk_rows: int = 1 << 20):
    """
    Predict every theta row at every independent value: returns mean[n_theta, n_ind]
    (and var).  This rank handles a contiguous slice of the theta rows; results are
    all-gathered.
    """
    n_theta, n_ind = len(theta_rows), len(independent)
    if getattr(emulator, "supports_device_grid", False) and dist.device_collectives():
        return _sweep_predict_device(emulator, np.asarray(theta_rows), np.asarray(independent, dtype=np.float64), want_var)
    start, stop = dist.contiguous_slice(n_theta)
    mine = theta_rows[start:stop]
    mean = np.empty((len(mine), n_ind))
    var = np.empty((len(mine), n_ind)) if want_var else None
    thetas_per_chunk = max(1, chunk_rows // n_ind)
    on_device = hasattr(emulator, "predict_grid")  # query rows formed on the device: only theta crosses the bus
    for c0 in range(0, len(mine), thetas_per_chunk):
        block = mine[c0 : c0 + thetas_per_chunk]
        if on_device:
            out = emulator.predict_grid(block, independent, want_var=want_var)
            m, v = out if want_var else (out, None)
        else:
            rows = np.empty((len(block) * n_ind, theta_rows.shape[1] + 1), dtype=np.float32)
            rows[:, 0] = np.tile(independent, len(block))
            rows[:, 1:] = np.repeat(block, n_ind, axis=0)
            out = emulator.predict_batch(rows, want_var=want_var)
            m, v = out if want_var else (out, None)
        if want_var:
            var[c0 : c0 + len(block)] = v.reshape(len(block), n_ind)
        mean[c0 : c0 + len(block)] = m.reshape(len(block), n_ind)
    mean = dist.all_gather_concat(mean, n_theta)
    if want_var:
        var = dist.all_gather_concat(var, n_theta)
        return mean, var
    return mean


def sobol_indices(emulator, model_specification: ModelSpecification, independent: np.ndarray,
                  base_samples: int = 1024, seed: int = 11) -> Dict[str, np.ndarray]:
    """First-order (S1) and total-order (ST) indices of the emulated mean, per independent value."""
    k = model_specification.number_of_parameters
    design = saltelli_design(model_specification, base_samples, seed)
    mean = sweep_predict(emulator, design, np.asarray(independent, dtype=np.float64), want_var=False)
    n = base_samples
    fa, fb = mean[:n], mean[n : 2 * n]
    variance = np.var(np.concatenate([fa, fb]), axis=0)
    s1 = np.empty((k, mean.shape[1]))
    st = np.empty((k, mean.shape[1]))
    for i in range(k):
        fab = mean[(2 + i) * n : (3 + i) * n]
        s1[i] = np.mean(fb * (fab - fa), axis=0) / variance
        st[i] = 0.5 * np.mean((fa - fab) ** 2, axis=0) / variance
    return {"S1": s1, "ST": st, "names": list(model_specification.parameter_names), "independent": np.asarray(independent)}
