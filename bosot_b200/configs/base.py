"""Minimal stand-in for the reference's pydantic ``BaseSettings`` configs (bosot/configs/**): plain
classes whose class attributes are the defaults, keyword overrides in the constructor, and
``.dict()`` returning every field so factories can splat it into constructors exactly as the
reference does (e.g. bosot/kernels/kernel_factory.py:52-54)."""
from __future__ import annotations

import copy


class BaseConfig:
    def __init__(self, **overrides):
        fields = self._fields()
        for k, v in overrides.items():
            if k not in fields:
                raise TypeError("%s has no field %r" % (type(self).__name__, k))
        for k, v in fields.items():
            setattr(self, k, copy.deepcopy(overrides.get(k, v)))
        missing = [k for k, v in fields.items() if v is REQUIRED and k not in overrides]
        if missing:
            raise TypeError("%s missing required field(s) %s" % (type(self).__name__, missing))

    @classmethod
    def _fields(cls) -> dict:
        out = {}
        for klass in reversed(cls.__mro__):
            for k, v in vars(klass).items():
                if not k.startswith("_") and not callable(v) and not isinstance(v, (classmethod, staticmethod, property)):
                    out[k] = v
        return out

    def dict(self) -> dict:
        return {k: getattr(self, k) for k in self._fields()}

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join("%s=%r" % kv for kv in self.dict().items()))


class _Required:
    def __repr__(self):
        return "REQUIRED"


REQUIRED = _Required()
