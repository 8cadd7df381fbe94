"""
Model specification container; same fields and behaviour as
/root/reference/swiftemulator/backend/model_specification.py:13-83.
"""

from typing import List, Optional

import attr


@attr.s
class ModelSpecification(object):
    number_of_parameters: int = attr.ib(validator=attr.validators.instance_of(int))
    parameter_names: List[str] = attr.ib()
    parameter_limits: List[List[float]] = attr.ib()
    parameter_printable_names: Optional[List[str]] = attr.ib(default=None)

    @parameter_limits.validator
    def _limits_match_names(self, attribute, value):
        if len(value) != len(self.parameter_names):
            raise AttributeError("Parameter limits and parameter names not of same length.")

    def __attrs_post_init__(self):
        if self.number_of_parameters != len(self.parameter_names):
            raise AttributeError("Number of parameters does not match the length of parameter names")
        if self.parameter_printable_names is None:
            self.parameter_printable_names = self.parameter_names

    @property
    def salib_problem(self):
        """The SALib-style problem dictionary (bounds drive the Sobol sweep)."""
        return {
            "num_vars": self.number_of_parameters,
            "names": self.parameter_names,
            "bounds": self.parameter_limits,
        }
