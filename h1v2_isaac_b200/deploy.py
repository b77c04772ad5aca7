"""Train -> deploy wire format: the `env.yaml` the reference's deployment stack reads next to a policy
(scripts/deploy/policies/<name>/env.yaml, consumed by packages/biped_deploy/biped_deploy/controllers/rl.py:34-176).

The reference writes it with utils/mdp/config_exporter.py:27-58 `get_deploy_config`, which only runs on cfgs whose observation
group carries the vendored manager's `history_step` (the CaT cfg); on the Flat / Rsl cfgs it raises AttributeError.  This module
produces the same document from the flattened kernel config, so a policy trained on this backend for any supported id ships
with the file the reference's unchanged sim2sim.py / sim2real.py expect.  Checked against the env.yaml the reference ships for
its Rsl-trained policy (scripts/deploy/policies/demo_rsl/env.yaml) in tests/test_boundary.py.
"""
from __future__ import annotations

from ._capi import H1v2Config
from .env import JOINT_NAMES, OBS_LAYOUT, flatten_cfg


def _f(x: float) -> float:
    """float32 config value -> the shortest decimal that round-trips (0.25, not 0.25000000000000006)."""
    return float(f"{float(x):.7g}")


def deploy_config(cfg) -> dict:
    """Deployment description of an env cfg tree (or an already flattened H1v2Config): control period, history, action scale,
    command ranges and dead zone, observation terms with scales, and per-joint gains / default pose in ACTION order."""
    c = cfg if isinstance(cfg, H1v2Config) else flatten_cfg(cfg)
    scales = [c.scale_ang_vel, c.scale_gravity, c.scale_cmd, c.scale_joint_pos, c.scale_joint_vel, c.scale_action]
    order = [int(c.joint_perm[i]) for i in range(len(JOINT_NAMES))]
    if order != list(range(len(JOINT_NAMES))):
        # rl.py:139-176 indexes joints in the MJCF / real-robot order (A/robots/h12.py:40-53): the cfg must preserve it
        raise ValueError("deploy_config: the action term must list the joints in the robot's own order (preserve_order=True), "
                         "as the reference's exporter asserts (config_exporter.py:28)")
    return {
        "control_dt": _f(c.sim_dt * c.decimation),
        "history_length": int(c.history_length),
        "history_step": 1,
        "action_scale": _f(c.action_scale),
        "velocity_deadzone": _f(c.velocity_deadzone) if c.command_class == 1 else 0.0,
        "command_ranges": {"lin_vel_x": [_f(v) for v in c.cmd_lin_x], "lin_vel_y": [_f(v) for v in c.cmd_lin_y],
                           "ang_vel_z": [_f(v) for v in c.cmd_ang_z]},
        "observations": [{"name": n, "scale": _f(s) if s != 1.0 else 1} for n, s in zip(OBS_LAYOUT, scales)],
        "joints": [{"name": JOINT_NAMES[j], "kp": _f(c.kp[j]), "kd": _f(c.kd[j]), "default_joint_pos": _f(c.default_joint_pos[j]), "enabled": True}
                   for j in order],
    }


def write_deploy_config(cfg, path: str) -> dict:
    import yaml
    d = deploy_config(cfg)
    with open(path, "w") as f:
        yaml.safe_dump(d, f, sort_keys=False)
    return d
