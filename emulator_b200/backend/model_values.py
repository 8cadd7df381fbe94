"""
Model values container; same data model and validation as
/root/reference/swiftemulator/backend/model_values.py:16-241.
"""

import json
from pathlib import Path
from typing import Dict, Hashable, Optional

import attr
import numpy as np
import yaml

_ALLOWED = ("independent", "dependent", "dependent_error")


@attr.s
class ModelValues(object):
    model_values: Dict[Hashable, Dict[str, Optional[np.array]]] = attr.ib()

    @model_values.validator
    def _check(self, attribute, value):
        for uid, relation in value.items():
            # the reference tests `or` here (model_values.py:54-57); kept for identical behaviour
            if not ("independent" in relation.keys() or "dependent" in relation.keys()):
                raise AttributeError(
                    f"independent and dependent must be available for all models, missing for model {uid}."
                )
            if "dependent_error" in relation.keys():
                if len(relation["dependent_error"]) != len(relation["dependent"]):
                    raise AttributeError(
                        f"dependent_error present but not the same length as dependent for model {uid}."
                    )
            for key in relation.keys():
                if key not in _ALLOWED:
                    raise AttributeError(
                        f"Invalid key {key} in model values container. Choose from one of {list(_ALLOWED)}."
                    )

    def items(self):
        return self.model_values.items()

    def keys(self):
        return self.model_values.keys()

    def values(self):
        return self.model_values.values()

    def __getitem__(self, key):
        return self.model_values[key]

    def __len__(self):
        return len(self.model_values)

    @property
    def number_of_variables(self) -> int:
        return sum(len(relation["independent"]) for relation in self.model_values.values())

    # ---- serialisation -------------------------------------------------------------------------
    def _plain(self):
        return {
            uid: {key: [float(v) for v in np.asarray(arr)] for key, arr in relation.items()}
            for uid, relation in self.model_values.items()
        }

    @staticmethod
    def _arrays(raw):
        return {uid: {key: np.array(arr) for key, arr in relation.items()} for uid, relation in raw.items()}

    def to_yaml(self, filename: Path):
        with open(filename, "w") as handle:
            yaml.safe_dump(self._plain(), handle)

    @classmethod
    def from_yaml(cls, filename: Path) -> "ModelValues":
        with open(filename, "r") as handle:
            return cls(model_values=cls._arrays(dict(yaml.safe_load(handle))))

    def to_json(self, filename: Path):
        with open(filename, "w") as handle:
            json.dump(self._plain(), handle)

    @classmethod
    def from_json(cls, filename: Path) -> "ModelValues":
        with open(filename, "r") as handle:
            return cls(model_values=cls._arrays(dict(json.load(handle))))
