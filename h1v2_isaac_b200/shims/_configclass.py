"""configclass: restatement of isaaclab.utils.configclass semantics that the reference's cfg modules rely on
(reference use: packages/biped_tasks/.../velocity_env_cfg.py:34-324, config/h12_12dof/rough_env_cfg.py:17-125):
class attributes (annotated or not) become per-instance fields, mutable defaults are deep-copied per instance,
__post_init__ runs after construction, and instances offer replace / copy / to_dict / from_dict / validate."""
from __future__ import annotations

import copy
import dataclasses
import types

MISSING = dataclasses.MISSING


def _is_field(name, value):
    if name.startswith("_"):
        return False
    if isinstance(value, (types.FunctionType, classmethod, staticmethod, property)):
        return False
    if isinstance(value, type) and name[:1].isupper() and name != "class_type":
        return False  # nested cfg classes (e.g. PolicyCfg, Ranges) stay class attributes
    return True


def _fields_of(cls):
    out = {}
    for klass in reversed(cls.__mro__):
        if klass is object:
            continue
        ann = klass.__dict__.get("__annotations__", {})
        for name in ann:
            if not name.startswith("_") and name not in klass.__dict__:
                out.setdefault(name, MISSING)
        for name, value in klass.__dict__.items():
            if _is_field(name, value):
                out[name] = value
    return out


def _to_dict(v):
    if hasattr(v, "__configclass_fields__"):
        return {k: _to_dict(getattr(v, k)) for k in v.__configclass_fields__()}
    if isinstance(v, dict):
        return {k: _to_dict(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return type(v)(_to_dict(x) for x in v)
    if isinstance(v, (types.FunctionType, type)):
        return f"{getattr(v, '__module__', '')}:{getattr(v, '__qualname__', repr(v))}"
    if hasattr(v, "to_dict") and callable(v.to_dict) and not isinstance(v, type):
        try:
            return v.to_dict()
        except Exception:
            return repr(v)
    return v


def configclass(cls=None, **kwargs):
    def wrap(cls):
        user_post_init = cls.__dict__.get("__post_init__")

        def __init__(self, **kw):
            fields = _fields_of(type(self))
            for name, default in fields.items():
                object.__setattr__(self, name, copy.deepcopy(default))
            for k, v in kw.items():
                if k not in fields:  # upstream's configclass is a dataclass: a misspelled field is a TypeError at construction, not a silent attribute
                    raise TypeError(f"{type(self).__name__}.__init__() got an unexpected keyword argument '{k}'")
                setattr(self, k, v)
            post = getattr(self, "__post_init__", None)
            if post is not None:
                post()

        cls.__init__ = __init__
        if user_post_init is None and not any("__post_init__" in k.__dict__ for k in cls.__mro__[1:]):
            cls.__post_init__ = lambda self: None
        cls.__configclass_fields__ = lambda self: list(_fields_of(type(self)).keys())
        cls.to_dict = lambda self: _to_dict(self)
        cls.copy = lambda self: copy.deepcopy(self)

        def replace(self, **kw):
            new = copy.deepcopy(self)
            for k, v in kw.items():
                setattr(new, k, v)
            return new

        cls.replace = replace

        def from_dict(self, data):
            for k, v in data.items():
                cur = getattr(self, k, None)
                if isinstance(v, dict) and hasattr(cur, "from_dict"):
                    cur.from_dict(v)
                else:
                    setattr(self, k, v)

        cls.from_dict = from_dict

        def validate(self, prefix=""):
            missing = []
            for k in self.__configclass_fields__():
                v = getattr(self, k)
                if v is MISSING:
                    missing.append(prefix + k)
                elif hasattr(v, "validate") and hasattr(v, "__configclass_fields__"):
                    try:
                        v.validate(prefix + k + ".")
                    except TypeError as e:
                        missing.append(str(e))
            if missing:
                raise TypeError("Missing values detected in object: " + ", ".join(missing))

        cls.validate = validate

        def __repr__(self):
            return f"{type(self).__name__}({', '.join(f'{k}={getattr(self, k)!r}' for k in self.__configclass_fields__())})"

        cls.__repr__ = __repr__
        return cls

    return wrap if cls is None else wrap(cls)
