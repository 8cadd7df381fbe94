"""
Model parameters container; same data model as
/root/reference/swiftemulator/backend/model_parameters.py:23-146, :206-301 (the corner-plot
helper is out of scope: plotting is not on the hot path).
"""

import json
from pathlib import Path
from typing import Dict, Hashable, List, Tuple

import attr
import numpy as np
import yaml
from sklearn.neighbors import KDTree


@attr.s
class ModelParameters(object):
    model_parameters: Dict[Hashable, Dict[str, float]] = attr.ib()

    @model_parameters.validator
    def _same_parameter_set(self, attribute, value):
        reference_keys = set(next(iter(value.values())).keys())
        for parameters in value.values():
            if set(parameters.keys()) != reference_keys:
                raise AttributeError("Models do not all have the same set of parameters.")

    def items(self):
        return self.model_parameters.items()

    def keys(self):
        return self.model_parameters.keys()

    def values(self):
        return self.model_parameters.values()

    def __getitem__(self, key):
        return self.model_parameters[key]

    def __len__(self):
        return len(self.model_parameters)

    def find_closest_model(
        self, comparison_parameters: Dict[str, float], number_of_close_models: int = 1
    ) -> Tuple[List[Hashable], List[Dict[str, float]]]:
        names = list(comparison_parameters.keys())
        identifiers = list(self.model_parameters.keys())
        table = np.array([[self.model_parameters[uid][name] for name in names] for uid in identifiers])
        point = np.array([comparison_parameters[name] for name in names]).reshape(1, -1)
        nearest = KDTree(table, leaf_size=2).query(point, k=number_of_close_models, return_distance=False)[0]
        closest = [identifiers[i] for i in nearest]
        return closest, [self.model_parameters[uid] for uid in closest]

    # ---- serialisation -----------------------------------------------------------------------
    def to_yaml(self, filename: Path):
        with open(filename, "w") as handle:
            yaml.safe_dump(_plain(self.model_parameters), handle)

    @classmethod
    def from_yaml(cls, filename: Path) -> "ModelParameters":
        with open(filename, "r") as handle:
            return cls(model_parameters=dict(yaml.safe_load(handle)))

    def to_json(self, filename: Path):
        with open(filename, "w") as handle:
            json.dump(_plain(self.model_parameters), handle)

    @classmethod
    def from_json(cls, filename: Path) -> "ModelParameters":
        with open(filename, "r") as handle:
            return cls(model_parameters=dict(json.load(handle)))


def _plain(mapping):
    return {uid: {name: float(value) for name, value in pars.items()} for uid, pars in mapping.items()}
