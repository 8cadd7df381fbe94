from .base import BaseEmulator  # noqa: F401
from .gaussian_process import GaussianProcessEmulator  # noqa: F401
from .gaussian_process_bins import GaussianProcessEmulatorBins  # noqa: F401
from .gaussian_process_mcmc import GaussianProcessEmulatorMCMC  # noqa: F401
from .gaussian_process_one_dim import GaussianProcessEmulator1D  # noqa: F401
from .multi_gaussian_process import MultipleGaussianProcessEmulator  # noqa: F401
