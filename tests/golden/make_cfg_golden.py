"""Generate tests/golden/flat_cfg_resolved.json and rsl_cfg_resolved.json: the reference's OWN resolved cfg trees of
Isaac-Velocity-Flat-H12_12dof-v0 (packages/biped_tasks/.../config/h12_12dof/flat_env_cfg.py:13-48 and parents, robot from
biped_assets/robots/h12.py:18-114), Isaac-Velocity-Rough-H12_12dof-v0 (.../config/h12_12dof/rough_env_cfg.py:65-125 on velocity_env_cfg.py:36-324; its
terrain generator resolves to the reference's in-tree utils/mdp/terrains.py:11-28 through the shim) and Isaac-Velocity-Rsl-H12_12dof-v0 (.../config/h12_12dof/rsl_env_cfg.py:44-540, robot
h12.py:117-206), imported from /root/reference through the isaaclab shims and flattened by h1v2_isaac_b200.env.flatten_cfg.
Run in the build container (the reference is not on the GPU box):  python tests/golden/make_cfg_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "h1v2_isaac_b200", "shims"), os.path.join(REF, "packages", "biped_tasks"), os.path.join(REF, "packages", "biped_assets")]

import biped_tasks.tasks  # noqa: E402,F401  (runs the reference's gym.register calls)
from isaaclab_tasks.utils import load_cfg_from_registry  # noqa: E402

from h1v2_isaac_b200.env import config_to_dict, constraint_curriculum, constraint_terms, curriculum_schedule, flatten_cfg, reward_slots  # noqa: E402

for TASK, name in (("Isaac-Velocity-Flat-H12_12dof-v0", "flat_cfg_resolved.json"), ("Isaac-Velocity-Rsl-H12_12dof-v0", "rsl_cfg_resolved.json"),
                   ("Isaac-Velocity-CaT-Flat-H12_12dof-v0", "cat_cfg_resolved.json"), ("Isaac-Velocity-Rough-H12_12dof-v0", "rough_cfg_resolved.json")):
    env_cfg = load_cfg_from_registry(TASK, "env_cfg_entry_point")
    agent_cfg = load_cfg_from_registry(TASK, "rsl_rl_cfg_entry_point" if "CaT" not in TASK else "clean_rl_cfg_entry_point")
    out = {"task": TASK, "num_envs": env_cfg.scene.num_envs, "kernel_config": config_to_dict(flatten_cfg(env_cfg)),
           "agent": agent_cfg.to_dict() if hasattr(agent_cfg, "to_dict") else {k: v for k, v in vars(agent_cfg).items() if isinstance(v, (int, float, str, bool, list, tuple, type(None)))},
           "reward_slots": reward_slots(env_cfg), "curriculum": curriculum_schedule(env_cfg),
           "constraint_terms": constraint_terms(env_cfg), "constraint_curriculum": constraint_curriculum(env_cfg)}
    with open(os.path.join(ROOT, "tests", "golden", name), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", name)

# the three Play ids (C12/flat_env_cfg.py:51-66, C12/rsl_env_cfg.py:543-564, C12/cat_env_cfg.py:568-583): flattened kernel config only
play = {}
for TASK in ("Isaac-Velocity-Flat-H12_12dof-Play-v0", "Isaac-Velocity-Rsl-H12_12dof-Play-v0", "Isaac-Velocity-CaT-Flat-H12_12dof-Play-v0",
             "Isaac-Velocity-Rough-H12_12dof-Play-v0"):
    env_cfg = load_cfg_from_registry(TASK, "env_cfg_entry_point")
    play[TASK] = {"num_envs": env_cfg.scene.num_envs, "kernel_config": config_to_dict(flatten_cfg(env_cfg))}
with open(os.path.join(ROOT, "tests", "golden", "play_cfg_resolved.json"), "w") as f:
    json.dump(play, f, indent=1, sort_keys=True)
print("wrote play_cfg_resolved.json")
