from .cross_check import CrossCheck, CrossCheckBins  # noqa: F401
from .sobol import saltelli_design, sobol_indices, sweep_predict  # noqa: F401
