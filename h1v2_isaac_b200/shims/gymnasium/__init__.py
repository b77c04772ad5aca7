"""Minimal gymnasium stand-in (registry, make, spaces, vector.utils.batch_space) -- only what the reference's
train/play scripts and env class touch (scripts/rsl_rl/train.py:47,102-113; utils/cat/cat_env.py:250-275)."""
from __future__ import annotations

import importlib
import sys
import types

import numpy as np

__version__ = "1.2.1+h1v2_b200_shim"


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape, self.dtype = (tuple(shape) if shape is not None else None), dtype


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        super().__init__(shape, dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape)

    def __repr__(self):
        return f"Box({self.shape}, {np.dtype(self.dtype).name})"


class Dict(Space, dict):
    def __init__(self, spaces=None, **kw):
        Space.__init__(self, None, None)
        dict.__init__(self, spaces or {}, **kw)

    @property
    def spaces(self):
        return self


spaces = types.ModuleType("gymnasium.spaces")
spaces.Space, spaces.Box, spaces.Dict = Space, Box, Dict
sys.modules["gymnasium.spaces"] = spaces


def batch_space(space, n=1):
    if isinstance(space, Dict):
        return Dict({k: batch_space(v, n) for k, v in space.items()})
    if isinstance(space, Box):
        return Box(np.broadcast_to(space.low, (n,) + space.shape), np.broadcast_to(space.high, (n,) + space.shape), (n,) + space.shape, space.dtype)
    raise TypeError(type(space))


vector = types.ModuleType("gymnasium.vector")
vector.utils = types.ModuleType("gymnasium.vector.utils")
vector.utils.batch_space = batch_space
sys.modules["gymnasium.vector"] = vector
sys.modules["gymnasium.vector.utils"] = vector.utils


class Env:
    metadata: dict = {}
    render_mode = None
    spec = None

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class _RecordVideo(Wrapper):
    def __init__(self, env, **kwargs):
        super().__init__(env)
        print("[gymnasium shim] RecordVideo is a no-op (no renderer in the B200 backend)")


wrappers = types.ModuleType("gymnasium.wrappers")
wrappers.RecordVideo = _RecordVideo
sys.modules["gymnasium.wrappers"] = wrappers


class EnvSpec:
    def __init__(self, id, entry_point, kwargs, disable_env_checker=True, **extra):
        self.id, self.entry_point, self.kwargs, self.disable_env_checker = id, entry_point, dict(kwargs or {}), disable_env_checker
        for k, v in extra.items():
            setattr(self, k, v)


registry: dict[str, EnvSpec] = {}


def register(id, entry_point=None, disable_env_checker=True, kwargs=None, **extra):
    registry[id] = EnvSpec(id, entry_point, kwargs, disable_env_checker, **extra)


def spec(id):
    if id not in registry:
        raise KeyError(f"No registered env with id: {id}")
    return registry[id]


def make(id, **kwargs):
    sp = spec(id)
    ep = sp.entry_point
    if isinstance(ep, str):
        mod, _, attr = ep.partition(":")
        ep = getattr(importlib.import_module(mod), attr)
    kw = dict(sp.kwargs)
    kw.update(kwargs)
    env = ep(**kw)
    try:
        env.spec = sp
    except Exception:
        pass
    return env


envs = types.ModuleType("gymnasium.envs")
envs.registry = registry
sys.modules["gymnasium.envs"] = envs
