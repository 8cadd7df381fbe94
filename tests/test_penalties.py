"""
The penalty calculators against golden vectors produced by the REFERENCE's own comparison/penalty.py
(tests/golden/make_golden.py `penalties`: the reference classes run in this container on seeded sweeps, in linear
and in log space), point by point, and the vectorised sweep form against the per-model form.
"""
import os

import numpy as np
import pytest

from emulator_b200 import comparison
from emulator_b200.backend import ModelValues
from emulator_b200.comparison import Observation

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "penalties.npz"))

# (tests/golden/make_golden.py PENALTY_CASES)
CASES = [
    ("L1PenaltyCalculator", dict(offset=5.0, lower=1.0, upper=4.0)),
    ("L1PenaltyCalculatorOneSided", dict(offset=5.0, maximum_penalty="above", lower=1.0, upper=4.0)),
    ("L1PenaltyCalculatorOneSided", dict(offset=4.0, maximum_penalty="below", lower=1.5, upper=4.2)),
    ("L2PenaltyCalculator", dict(offset=5.0, lower=1.0, upper=4.0)),
    ("L1VariablePenaltyCalculator", dict(offset_lower=2.0, offset_upper=6.0, offset_transition=2.5, transition_width=0.5, lower=1.0,
                                         upper=4.0, offset_below_above_ratio=0.5)),
    ("L1SqueezePenaltyCalculator", dict(offset_squeeze=1.0, offset_normal=5.0, offset_transition=2.5, transition_width=0.5, lower=1.0,
                                        upper=4.0, offset_below_above_ratio=2.0)),
    ("GaussianDataErrorsPenaltyCalculator", dict(sigma_max=3.0, lower=1.0, upper=4.0)),
    ("GaussianPercentErrorsPenaltyCalculator", dict(percent_error=10.0, sigma_max=3.0, lower=1.0, upper=4.0)),
    ("GaussianDataErrorsPercentFloorPenaltyCalculator", dict(percent_error=10.0, sigma_max=2.0, lower=1.0, upper=4.0)),
    ("GaussianWeightedDataErrorsPercentFloorPenaltyCalculator", dict(percent_error=10.0, sigma_max=3.0, weight=2.0, lower=1.0, upper=4.0)),
]


def _setup(index, log_space):
    cls, kwargs = CASES[index]
    independent, dependent = GOLDEN["independent"], GOLDEN["dependent"]
    if log_space:  # the transformation make_golden.py applies
        kw = {k: (float(np.log10(v)) if k in ("lower", "upper", "offset_transition") else v) for k, v in kwargs.items()}
        kw = {k: (0.3 * v / 5.0 if k.startswith("offset") and k not in ("offset_transition", "offset_below_above_ratio") else v)
              for k, v in kw.items()}
        independent, dependent = np.log10(independent), np.log10(np.abs(dependent) + 1.0)
    else:
        kw = dict(kwargs)
    calc = getattr(comparison, cls)(**kw)
    calc.register_observation(Observation(GOLDEN["obs_x"], GOLDEN["obs_y"], GOLDEN["scatter"]), log_space, log_space)
    return calc, independent, dependent


@pytest.mark.parametrize("index", range(len(CASES)))
@pytest.mark.parametrize("log_space", [False, True])
def test_penalties_match_the_reference_point_by_point(index, log_space):
    calc, independent, dependent = _setup(index, log_space)
    golden = GOLDEN[f"case{index}_{'log' if log_space else 'lin'}"]
    valid = ~np.isnan(golden[0])
    assert valid.sum() > 0
    for r in range(len(dependent)):
        got = calc.penalty(independent, dependent[r])
        assert np.allclose(got, golden[r][valid], rtol=1e-13, atol=1e-15)
    # the sweep form: every row at once, any collation
    for collate in (np.mean, np.max, np.median):
        grid = calc.penalties_grid(independent, dependent, collate)
        assert np.allclose(grid, collate(golden[:, valid], axis=1), rtol=1e-13, atol=1e-15)
    # and the container form of the reference API
    model = ModelValues({r: {"independent": independent, "dependent": dependent[r]} for r in range(4)})
    per_model = calc.penalties(model, np.mean)
    assert np.allclose([per_model[r] for r in range(4)], np.mean(golden[:4][:, valid], axis=1), rtol=1e-13, atol=1e-15)


def test_no_valid_point_gives_zero_and_bad_options_raise():
    calc = comparison.L1PenaltyCalculator(offset=1.0, lower=100.0, upper=200.0)
    calc.register_observation(Observation(GOLDEN["obs_x"], GOLDEN["obs_y"]), False, False)
    assert calc.penalty(GOLDEN["independent"], GOLDEN["dependent"][0]) == 0.0
    assert np.array_equal(calc.penalties_grid(GOLDEN["independent"], GOLDEN["dependent"]), np.zeros(len(GOLDEN["dependent"])))
    with pytest.raises(AttributeError):
        comparison.L1PenaltyCalculatorOneSided(offset=1.0, maximum_penalty="sideways", lower=0.0, upper=1.0)
    with pytest.raises(NotImplementedError):
        calc.plot_penalty()
