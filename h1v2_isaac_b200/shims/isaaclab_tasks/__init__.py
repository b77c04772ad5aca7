"""isaaclab_tasks stand-in: package walker, registry cfg loader, hydra-style CLI overrides, and the upstream
locomotion-velocity cfg/mdp namespaces."""
from __future__ import annotations

import ast
import functools
import importlib
import os
import pkgutil
import re
import sys

from h1v2_isaac_b200.shims._lenient import install_finder, make_lenient

install_finder()
_me = sys.modules[__name__]


def import_packages(package_name: str, blacklist_pkgs: list[str] | None = None):
    """Import every sub-module so that gym.register calls run (reference: biped_tasks/tasks/__init__.py:13).
    Unlike upstream this walker is tolerant: task variants outside the B200 backend's scope that need
    un-shimmed isaaclab internals are skipped with a one-line notice instead of aborting the import."""
    blacklist_pkgs = blacklist_pkgs or []
    package = importlib.import_module(package_name)
    for _, name, _ in pkgutil.walk_packages(package.__path__, package.__name__ + ".", onerror=lambda n: None):
        if any(b in name for b in blacklist_pkgs):
            continue
        try:
            importlib.import_module(name)
        except Exception as e:  # noqa: BLE001
            if os.environ.get("H1V2_SHIM_VERBOSE"):
                print(f"[isaaclab_tasks shim] skipped {name}: {type(e).__name__}: {e}")


def load_cfg_from_registry(task_name: str, entry_point_key: str):
    import gymnasium as gym
    cfg_entry_point = gym.spec(task_name.split(":")[-1]).kwargs.get(entry_point_key)
    if cfg_entry_point is None:
        raise ValueError(f"Could not find configuration for the environment: '{task_name}' (key {entry_point_key}).")
    if isinstance(cfg_entry_point, str) and cfg_entry_point.endswith(".yaml"):
        import yaml
        mod, _, fname = cfg_entry_point.partition(":")
        path = os.path.join(os.path.dirname(importlib.import_module(mod).__file__), fname)
        with open(path, encoding="utf-8") as f:
            return yaml.full_load(f)
    if callable(cfg_entry_point):
        cfg_cls = cfg_entry_point
    else:
        mod, _, attr = cfg_entry_point.partition(":")
        cfg_cls = getattr(importlib.import_module(mod), attr)
    return cfg_cls() if callable(cfg_cls) else cfg_cls


def parse_env_cfg(task_name: str, device: str = "cuda:0", num_envs: int | None = None, use_fabric: bool | None = None):
    cfg = load_cfg_from_registry(task_name, "env_cfg_entry_point")
    cfg.sim.device = device
    if num_envs is not None:
        cfg.scene.num_envs = num_envs
    return cfg


def get_checkpoint_path(log_path: str, run_dir: str = ".*", checkpoint: str = ".*", other_dirs=None, sort_alpha: bool = True) -> str:
    runs = [os.path.join(log_path, r.name) for r in os.scandir(log_path) if r.is_dir() and re.match(run_dir, r.name)]
    runs.sort() if sort_alpha else runs.sort(key=os.path.getmtime)
    if not runs:
        raise ValueError(f"No runs present in the directory: '{log_path}' match: '{run_dir}'.")
    run_path = os.path.join(runs[-1], *(other_dirs or []))
    files = [f for f in os.listdir(run_path) if re.match(checkpoint, f)]
    if not files:
        raise ValueError(f"No checkpoints in the directory: '{run_path}' match '{checkpoint}'.")
    files.sort(key=lambda m: f"{m:0>15}")
    return os.path.join(run_path, files[-1])


def _apply_overrides(root_objs: dict, argv: list[str]):
    """hydra-style `env.a.b=value` / `agent.x=value` overrides (reference use: slurm/base_job.sh:37-39)."""
    rest = []
    for arg in argv:
        m = re.match(r"^(env|agent)\.([\w\.]+)=(.*)$", arg)
        if not m:
            rest.append(arg)
            continue
        obj = root_objs[m.group(1)]
        path = m.group(2).split(".")
        for key in path[:-1]:
            obj = obj[key] if isinstance(obj, dict) else getattr(obj, key)
        try:
            val = ast.literal_eval(m.group(3))
        except Exception:
            val = {"true": True, "false": False, "null": None, "none": None}.get(m.group(3).lower(), m.group(3))
        if isinstance(obj, dict):
            obj[path[-1]] = val
        else:
            setattr(obj, path[-1], val)
    return rest


def hydra_task_config(task_name: str, agent_cfg_entry_point: str):
    def decorator(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            env_cfg = load_cfg_from_registry(task_name, "env_cfg_entry_point")
            agent_cfg = load_cfg_from_registry(task_name, agent_cfg_entry_point) if agent_cfg_entry_point else None
            sys.argv = [sys.argv[0]] + _apply_overrides({"env": env_cfg, "agent": agent_cfg}, sys.argv[1:])
            return func(env_cfg, agent_cfg, *args, **kwargs)
        return wrapper
    return decorator


utils = make_lenient("isaaclab_tasks.utils", import_packages=import_packages, get_checkpoint_path=get_checkpoint_path, parse_env_cfg=parse_env_cfg,
                     load_cfg_from_registry=load_cfg_from_registry)
make_lenient("isaaclab_tasks.utils.hydra", hydra_task_config=hydra_task_config)
make_lenient("isaaclab_tasks.utils.parse_cfg", load_cfg_from_registry=load_cfg_from_registry, parse_env_cfg=parse_env_cfg,
             get_checkpoint_path=get_checkpoint_path)
make_lenient("isaaclab_tasks.utils.importer", import_packages=import_packages)
make_lenient("isaaclab_tasks.manager_based")
make_lenient("isaaclab_tasks.manager_based.locomotion")
_vel = make_lenient("isaaclab_tasks.manager_based.locomotion.velocity")


def _build_velocity_namespace():
    """isaaclab_tasks.manager_based.locomotion.velocity.{mdp, velocity_env_cfg}.
    mdp = isaaclab.envs.mdp + the four locomotion-specific term names the H1-2 cfgs use.
    velocity_env_cfg re-exports the reference's OWN in-tree mirror (biped_tasks/tasks/locomotion/velocity/
    velocity_env_cfg.py) when biped_tasks is importable, so no upstream file is restated here."""
    import isaaclab.envs as ienvs
    base = ienvs.mdp
    names = {n: getattr(base, n) for n in base.__all__}
    for extra in ("feet_air_time", "feet_air_time_positive_biped", "feet_slide", "track_lin_vel_xy_yaw_frame_exp",
                  "track_ang_vel_z_world_exp", "terrain_levels_vel"):
        names[extra] = ienvs._term(extra)
    mdp = make_lenient("isaaclab_tasks.manager_based.locomotion.velocity.mdp", **names)
    mdp.__all__ = list(names)
    _vel.mdp = mdp


_build_velocity_namespace()


class _VelocityEnvCfgModule(type(_me)):
    """Lazy module: resolves names from the reference's in-tree mirror at first use (avoids an import cycle)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        mirror = importlib.import_module("biped_tasks.tasks.locomotion.velocity.velocity_env_cfg")
        return getattr(mirror, name)


_vcfg = _VelocityEnvCfgModule("isaaclab_tasks.manager_based.locomotion.velocity.velocity_env_cfg")
sys.modules[_vcfg.__name__] = _vcfg
_vel.velocity_env_cfg = _vcfg
