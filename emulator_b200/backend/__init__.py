from .model_parameters import ModelParameters  # noqa: F401
from .model_specification import ModelSpecification  # noqa: F401
from .model_values import ModelValues  # noqa: F401
