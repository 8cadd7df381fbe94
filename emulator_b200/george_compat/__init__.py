"""
Drop-in stand-in for the parts of ``george`` that swiftemulator uses, backed by the CUDA
library.  ``from emulator_b200 import george_compat as george`` is the whole integration on
the reference side.
"""

from . import kernels, modeling  # noqa: F401
from .gp import GP, TINY, pinned_evaluator  # noqa: F401
