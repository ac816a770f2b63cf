"""Nested configuration namespace.  API mirror of the reference's ``hparams/hp.py:9-66`` (modules are
constructed from ``Hparams`` sub-trees and call ``get`` / ``has`` / ``check_arg_in_hparams``)."""
from __future__ import annotations

import json
from types import SimpleNamespace
from typing import Any, Mapping

import yaml


def _to_hp(obj: Any) -> Any:
    if isinstance(obj, Mapping):
        return Hparams(**{k: _to_hp(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [_to_hp(v) for v in obj]
    return obj


def _to_plain(obj: Any) -> Any:
    if isinstance(obj, Hparams):
        return {k: _to_plain(v) for k, v in vars(obj).items()}
    if isinstance(obj, (list, tuple)):
        return [_to_plain(v) for v in obj]
    return obj


class Hparams(SimpleNamespace):
    def __init__(self, *args, **kwargs):
        super().__init__(**kwargs)

    # -- queries used by every module constructor
    def check_arg_in_hparams(self, *names: str) -> None:
        missing = [n for n in names if n not in vars(self)]
        if missing:
            raise ValueError(f"{missing[0]} not specifed in the hyperapramer: {self}")

    def get(self, name: str, default=None):
        return vars(self).get(name, default)

    def has(self, name: str) -> bool:
        return name in vars(self)

    def merge(self, other: "Hparams") -> "Hparams":
        return Hparams(**vars(self), **vars(other))

    def __eq__(self, other):
        return vars(self) == vars(other)

    def __repr__(self):
        return repr(vars(self))

    # -- (de)serialisation
    def to_dict(self) -> Mapping[str, Any]:
        return _to_plain(self)

    @classmethod
    def from_dict(cls, d: Mapping[str, Any]) -> "Hparams":
        return _to_hp(d)

    @classmethod
    def from_json(cls, text: str) -> "Hparams":
        return _to_hp(json.loads(text))

    @classmethod
    def from_jsonfile(cls, path: str) -> "Hparams":
        with open(path) as f:
            return _to_hp(json.load(f))

    @classmethod
    def from_argparse(cls, args) -> "Hparams":
        return _to_hp(dict(vars(args)))

    @classmethod
    def from_yamlfile(cls, path: str) -> "Hparams":
        with open(path) as f:
            return _to_hp(yaml.safe_load(f))

    def save(self, path: str) -> None:
        with open(path, "w") as f:
            yaml.dump(self.to_dict(), f)
