"""TEST INFRASTRUCTURE ONLY -- import shims that make the *unmodified* reference importable on CPU.

The reference (``/root/reference`` in the build container) depends on packages that are not installed
(``param``, ``mashumaro``, ``matplotlib``, ``ray``, ``natsort``, ``dask``, ``implicit``, ``rmet``).  None of them
carries arithmetic of the hot path except ``rmet`` (metrics), which is restated in :mod:`oracle.rmet_restated`
("parity unpinned").  Everything here is non-arithmetic glue: dataclass ``from_dict`` parsing,
empty modules, a typing alias.

Only ``oracle/make_golden.py`` (fixture generator, run in the build container) and ``bench.py --impl reference``
(when a reference tree is present) use this module.  The product package never imports it.
"""
from __future__ import annotations

import dataclasses
import enum
import os
import sys
import types
import typing

REFERENCE_CANDIDATES = ("/root/reference", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "baseline", "_ref"))


def find_reference() -> str | None:
    for p in REFERENCE_CANDIDATES:
        if os.path.isfile(os.path.join(p, "algorithms", "sgd_alg.py")):
            return p
    return None


# ----------------------------------------------------------------------------- mashumaro
class MissingField(Exception):
    pass


def _convert(tp, v):
    if v is None:
        return None
    origin = typing.get_origin(tp)
    if origin is typing.Union or origin is types.UnionType:
        last = None
        for arm in typing.get_args(tp):
            if arm is type(None):
                continue
            try:
                return _convert(arm, v)
            except Exception as e:  # try next arm
                last = e
        if last is not None:
            raise last
        return v
    if origin in (list, typing.List):
        (a,) = typing.get_args(tp) or (typing.Any,)
        return [_convert(a, x) for x in v]
    if origin in (set, typing.Set):
        (a,) = typing.get_args(tp) or (typing.Any,)
        return {_convert(a, x) for x in v}
    if origin in (dict, typing.Dict):
        return dict(v)
    if isinstance(tp, type):
        if dataclasses.is_dataclass(tp):
            return tp.from_dict(v) if isinstance(v, dict) else v
        if issubclass(tp, enum.Enum):
            return v if isinstance(v, tp) else tp(v)
    return v


class DataClassDictMixin:
    @classmethod
    def from_dict(cls, d: dict):
        hints = typing.get_type_hints(cls)
        vals = {}
        for f in dataclasses.fields(cls):
            if not f.init:
                continue
            if f.name in d:
                des = f.metadata.get("deserialize") if f.metadata else None
                vals[f.name] = des(d[f.name]) if des is not None else _convert(hints.get(f.name, typing.Any),
                                                                                 d[f.name])
            elif f.default is dataclasses.MISSING and f.default_factory is dataclasses.MISSING:
                raise MissingField(f'Field "{f.name}" missing for {cls.__name__}')
        return cls(**vals)

    def to_dict(self):
        return dataclasses.asdict(self)


class DataClassYAMLMixin(DataClassDictMixin):
    pass


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install(reference_root: str | None = None) -> str:
    """Install the shims and put the reference on ``sys.path``.  Returns the reference root."""
    global _installed
    root = reference_root or find_reference()
    if root is None:
        raise FileNotFoundError("reference tree not found (looked in %s)" % (REFERENCE_CANDIDATES,))
    if _installed:
        return root
    sys.dont_write_bytecode = True  # reference tree is read-only

    # param: Parameterized base + descriptors that collapse to their default
    _module("param", Parameterized=type("Parameterized", (), {}),
            Selector=lambda default=None, **kw: default,
            Integer=lambda default=None, **kw: default,
            Number=lambda default=None, **kw: default)

    ma = _module("mashumaro", DataClassDictMixin=DataClassDictMixin)
    ma.exceptions = _module("mashumaro.exceptions", MissingField=MissingField)
    ma.mixins = _module("mashumaro.mixins")
    ma.mixins.yaml = _module("mashumaro.mixins.yaml", DataClassYAMLMixin=DataClassYAMLMixin)

    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _module("matplotlib")
        mpl.pyplot = _module("matplotlib.pyplot")
        mpl.colors = _module("matplotlib.colors")
    for name in ("dask", "dask.dataframe", "implicit"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _module(name)
    if "implicit.als" not in sys.modules:
        try:
            __import__("implicit.als")
        except Exception:
            _module("implicit.als", AlternatingLeastSquares=object)
    try:
        import ray.air  # noqa: F401
    except Exception:
        ray = _module("ray")
        ray.air = _module("ray.air", session=types.SimpleNamespace(report=lambda *a, **k: None))
    try:
        import natsort  # noqa: F401
    except Exception:
        _module("natsort", natsorted=sorted)
    try:
        import wandb  # noqa: F401
    except Exception:
        _module("wandb", log=lambda *a, **k: None)
    try:
        import rmet  # noqa: F401
    except Exception:
        from oracle import rmet_restated
        sys.modules["rmet"] = rmet_restated

    import torch.utils.data.dataloader as tdl
    if not hasattr(tdl, "T_co"):
        tdl.T_co = typing.TypeVar("T_co", covariant=True)

    if root not in sys.path:
        sys.path.insert(0, root)
    _installed = True
    return root
