"""
Prediction sweeps: `mock_hypercube` and `mock_sweep` with the reference's signatures and
outputs (/root/reference/swiftemulator/mocking/__init__.py:17-200).  Where the reference calls
`emulator.predict_values` once per parameter point in a Python loop (:84-95, :187-198), all
(theta x independent) query rows are predicted in one batched device call -- formed on the device
when the emulator offers `predict_grid`, on the host with `predict_batch` -- and other emulators take
the per-point path.
"""

from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from .backend import ModelParameters, ModelSpecification, ModelValues


def _unique_independents(emulator) -> np.ndarray:
    values = emulator.model_values.model_values
    return np.unique([item for uid in values.keys() for item in values[uid]["independent"]])


def _predict_all(emulator, parameters: ModelParameters, independent: np.ndarray, kwargs) -> ModelValues:
    out = {}
    uids = list(parameters.model_parameters.keys())
    if hasattr(emulator, "predict_grid") and not kwargs:
        # one device call; the (independent, theta) query rows are formed on the device
        order = emulator.parameter_order
        theta = np.array([[parameters.model_parameters[uid][name] for name in order] for uid in uids], dtype=np.float64)
        mean, var = emulator.predict_grid(theta.reshape(len(uids), len(order)), independent, want_var=True)
        for k, uid in enumerate(uids):
            out[uid] = {"independent": independent, "dependent": mean[k], "dependent_error": var[k]}
    elif hasattr(emulator, "predict_batch") and not kwargs:
        order = emulator.parameter_order
        rows = np.empty((len(uids) * len(independent), len(order) + 1), dtype=np.float32)
        for k, uid in enumerate(uids):
            block = rows[k * len(independent) : (k + 1) * len(independent)]
            block[:, 0] = independent
            block[:, 1:] = np.array([parameters.model_parameters[uid][name] for name in order])
        mean, var = emulator.predict_batch(rows, want_var=True)
        for k, uid in enumerate(uids):
            sl = slice(k * len(independent), (k + 1) * len(independent))
            out[uid] = {"independent": independent, "dependent": mean[sl], "dependent_error": var[sl]}
    else:
        for uid in uids:
            dep, dep_err = emulator.predict_values(
                independent=independent, model_parameters=parameters.model_parameters[uid], **kwargs
            )
            out[uid] = {"independent": independent, "dependent": dep, "dependent_error": dep_err}
    return ModelValues(model_values=out)


def random_cube(model_specification: ModelSpecification, samples: int, prefix: str = "", rng=None) -> ModelParameters:
    """Uniform random cube scaled to the limits (design/random.py:14-61, transform.py:13-57)."""
    draw = np.random.rand(samples, model_specification.number_of_parameters) if rng is None else rng.random(
        (samples, model_specification.number_of_parameters)
    )
    limits = np.array(model_specification.parameter_limits, dtype=np.float64)
    scaled = limits[:, 0] + draw * (limits[:, 1] - limits[:, 0])
    return ModelParameters(
        model_parameters={
            f"{prefix}{i}": {name: float(scaled[i, j]) for j, name in enumerate(model_specification.parameter_names)}
            for i in range(samples)
        }
    )


def mock_hypercube(
    emulator, model_specification: ModelSpecification, samples: int,
    predict_values_kwargs: Optional[Dict[str, Any]] = None, rng=None,
) -> Tuple[ModelValues, ModelParameters]:
    kwargs = {} if predict_values_kwargs is None else predict_values_kwargs
    parameters = random_cube(model_specification, samples, prefix="emulated_", rng=rng)
    return _predict_all(emulator, parameters, _unique_independents(emulator), kwargs), parameters


def mock_sweep(
    emulator, model_specification: ModelSpecification, samples: int, sweep_parameter: str,
    center_point: Dict[str, float], predict_values_kwargs: Optional[Dict[str, Any]] = None,
) -> Tuple[ModelValues, ModelParameters]:
    kwargs = {} if predict_values_kwargs is None else predict_values_kwargs
    index = model_specification.parameter_names.index(sweep_parameter)
    sweep = np.linspace(*model_specification.parameter_limits[index], samples)
    parameters = ModelParameters(
        {
            f"emulated_{key}": {k: (v if k != sweep_parameter else sweep[key]) for k, v in center_point.items()}
            for key in range(samples)
        }
    )
    return _predict_all(emulator, parameters, _unique_independents(emulator), kwargs), parameters
