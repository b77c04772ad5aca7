"""Lenient placeholders: let the reference's config modules import and build their cfg trees when isaaclab is absent.

Anything the B200 backend actually consumes is implemented for real next to this file (configclass, the cfg classes the
Flat task reads, the env, the wrapper, the runner).  Everything else -- USD spawners, terrain generators, managers of
task ids that SURVEY.md section 8 marks out of scope -- resolves to an inert, subclassable, attribute-bag placeholder."""
from __future__ import annotations

import copy
import sys
import types


class PlaceholderMeta(type):
    """Class-level access to an unknown CamelCase attribute (RayCasterCfg.OffsetCfg, ...) yields a nested placeholder."""

    def __getattr__(cls, name):
        if name.startswith("_") or not name[:1].isupper():
            raise AttributeError(name)
        sub = PlaceholderMeta(name, (Placeholder,), {"__module__": cls.__module__})
        setattr(cls, name, sub)
        return sub


class Placeholder(metaclass=PlaceholderMeta):
    """Accepts any constructor arguments, any attribute assignment; unknown attributes read as None."""

    def __init__(self, *args, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return None

    def replace(self, **kwargs):
        new = copy.deepcopy(self)
        for k, v in kwargs.items():
            setattr(new, k, v)
        return new

    def copy(self):
        return copy.deepcopy(self)

    def to_dict(self):
        return {k: (v.to_dict() if hasattr(v, "to_dict") else v) for k, v in vars(self).items()}

    def __call__(self, *args, **kwargs):
        return None


def _make_func(name: str):
    def _f(*args, **kwargs):
        raise NotImplementedError(f"{name} is a placeholder: the B200 backend computes this term inside its fused kernel")
    _f.__name__ = _f.__qualname__ = name
    return _f


class LenientModule(types.ModuleType):
    """Module whose unknown CamelCase attributes are placeholder classes and unknown snake_case ones named functions."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        if full in sys.modules:
            return sys.modules[full]
        if name[:1].isupper():
            obj = PlaceholderMeta(name, (Placeholder,), {"__module__": self.__name__})
        elif name.isupper():
            obj = Placeholder()
        else:
            obj = _make_func(name)
        setattr(self, name, obj)
        return obj


def make_lenient(name: str, **attrs) -> LenientModule:
    mod = sys.modules.get(name)
    if not isinstance(mod, LenientModule):
        mod = LenientModule(name)
        mod.__path__ = []  # behave like a package so that "import a.b.c" works for arbitrary depth
        sys.modules[name] = mod
        parent, _, child = name.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], child, mod)
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


class _LenientFinder:
    """Resolve any still-unknown submodule of the shimmed top-level packages to a lenient module."""

    ROOTS = ("isaaclab", "isaaclab_tasks", "isaaclab_rl", "isaaclab_assets", "isaacsim", "omni", "pxr", "carb")

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        root = fullname.split(".")[0]
        if root in self.ROOTS and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return make_lenient(spec.name)

    def exec_module(self, module):
        return None


def install_finder():
    if not any(isinstance(f, _LenientFinder) for f in sys.meta_path):
        sys.meta_path.append(_LenientFinder())
