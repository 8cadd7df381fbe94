from .penalty import (  # noqa: F401
    GaussianDataErrorsPenaltyCalculator, GaussianDataErrorsPercentFloorPenaltyCalculator,
    GaussianPercentErrorsPenaltyCalculator, GaussianWeightedDataErrorsPercentFloorPenaltyCalculator,
    L1PenaltyCalculator, L1PenaltyCalculatorOneSided, L1SqueezePenaltyCalculator, L1VariablePenaltyCalculator,
    L2PenaltyCalculator, Observation, PenaltyCalculator,
)
