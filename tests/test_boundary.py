"""Drop-in boundary, host side only: cfg-tree flattening against the golden fixture generated from the
reference's own cfg classes, the self-contained task registration, and the reference's UNMODIFIED train.py reaching the
backend through the import shims (it must then fail loudly here: no CUDA device, no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "flat_cfg_resolved.json")))
GOLD_RSL = json.load(open(os.path.join(ROOT, "tests", "golden", "rsl_cfg_resolved.json")))
GOLD_ROUGH = json.load(open(os.path.join(ROOT, "tests", "golden", "rough_cfg_resolved.json")))
GOLD_PLAY = json.load(open(os.path.join(ROOT, "tests", "golden", "play_cfg_resolved.json")))
has_ref = os.path.isdir(os.path.join(REF, "packages", "biped_tasks"))


def _close(a, b):
    if isinstance(a, list):
        return len(a) == len(b) and all(_close(x, y) for x, y in zip(a, b))
    return a == b or abs(a - b) <= 1e-7 * max(1.0, abs(a))


def test_default_config_equals_reference_cfg_golden(cfg):
    """h1v2_default_config (the C restatement of the resolved Flat cfg) == the reference's cfg tree, value by value."""
    from h1v2_isaac_b200.env import config_to_dict
    mine = config_to_dict(cfg)
    gold = GOLD["kernel_config"]
    assert set(mine) == set(gold)
    for k in gold:
        assert _close(mine[k], gold[k]), k


def test_rsl_config_equals_reference_cfg_golden():
    """h1v2_rsl_config (C restatement of config/h12_12dof/rsl_env_cfg.py:44-540) == the reference's Rsl cfg tree, value by value;
    the 12 modify_reward_weight curriculum terms (:447-497) target the weights the terms already have."""
    from h1v2_isaac_b200._capi import rsl_config
    from h1v2_isaac_b200.env import config_to_dict
    mine = config_to_dict(rsl_config())
    gold = GOLD_RSL["kernel_config"]
    assert set(mine) == set(gold)
    for k in gold:
        assert _close(mine[k], gold[k]), k
    assert mine["history_length"] == 6 and mine["command_class"] == 1 and mine["max_delay"] == 0
    assert len(GOLD_RSL["curriculum"]) == 12
    for slot, weight, num_steps in GOLD_RSL["curriculum"]:
        assert num_steps == 24 * 5000 and _close(weight, mine["rew_weight"][slot])


def test_rough_config_equals_reference_cfg_golden():
    """h1v2_rough_config (C restatement of config/h12_12dof/rough_env_cfg.py:65-125 on velocity_env_cfg.py:36-324 with the in-tree terrain
    generator cfg utils/mdp/terrains.py:11-28) == the reference's own Rough cfg tree flattened, value by value: 10 x 20 tiles of 8 m at
    0.1 m / 5 mm, levels 0..4, rim 3 cells, max init level 5, terrain_levels_vel on; base_lin_vel + 45 + 17 x 11 height scan, no history;
    and the Play id (rough_env_cfg.py:128-156) equals tasks.rough_play_env_cfg."""
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200._capi import obs_dim_of, rough_config
    from h1v2_isaac_b200.env import config_to_dict, flatten_cfg, reward_slots
    mine = config_to_dict(rough_config())
    gold = GOLD_ROUGH["kernel_config"]
    assert set(mine) == set(gold)
    for k in gold:
        assert _close(mine[k], gold[k]), k
    assert mine["terrain_enable"] == 1 and mine["terrain_rows"] == 10 and mine["terrain_cols"] == 20 and mine["terrain_curriculum"] == 1
    assert obs_dim_of(rough_config()) == 3 + 45 + 17 * 11 == 235
    # the self-contained trees invert flatten_cfg, with the reference's reward term names
    tree = tasks.rough_env_cfg(32)
    assert config_to_dict(flatten_cfg(tree)) == mine
    assert reward_slots(tree) == GOLD_ROUGH["reward_slots"]
    play = config_to_dict(flatten_cfg(tasks.rough_play_env_cfg()))
    gp = GOLD_PLAY["Isaac-Velocity-Rough-H12_12dof-Play-v0"]
    assert gp["num_envs"] == 50 and set(play) == set(gp["kernel_config"])
    for k in play:
        assert _close(play[k], gp["kernel_config"][k]), k


def test_flatten_refuses_what_the_rough_kernel_cannot_do():
    """Terrain / scanner settings outside the built path raise NotImplementedError naming the entry (never approximated): upstream's mesh
    sub-terrains, a second sub-terrain, a scanner that is not yaw-aligned, height_scan on a history cfg."""
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200.env import flatten_cfg
    from h1v2_isaac_b200.shims._lenient import Placeholder

    class MeshPyramidStairsTerrainCfg(Placeholder):
        pass

    t = tasks.rough_env_cfg(16)
    t.scene.terrain.terrain_generator.sub_terrains["pyramid_stairs"] = MeshPyramidStairsTerrainCfg(proportion=0.2)
    with pytest.raises(NotImplementedError, match="HfRandomUniformTerrainCfg"):
        flatten_cfg(t)
    t = tasks.rough_env_cfg(16)
    t.scene.height_scanner.attach_yaw_only = False
    with pytest.raises(NotImplementedError, match="attach_yaw_only"):
        flatten_cfg(t)
    t = tasks.rough_env_cfg(16)
    t.observations.policy.history_length = 5
    with pytest.raises(NotImplementedError, match="history"):
        flatten_cfg(t)
    t = tasks.rough_env_cfg(16)
    t.scene.terrain.terrain_type = "usd"
    with pytest.raises(NotImplementedError, match="terrain_type"):
        flatten_cfg(t)


def test_self_contained_rsl_task_roundtrip():
    """tasks.rsl_env_cfg() inverts flatten_cfg on the Rsl config, with the reference's term names, slots and curriculum."""
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200._capi import rsl_config
    from h1v2_isaac_b200.env import config_to_dict, curriculum_schedule, flatten_cfg, reward_slots
    tree = tasks.rsl_env_cfg(32)
    assert config_to_dict(flatten_cfg(tree)) == config_to_dict(rsl_config())
    assert reward_slots(tree) == GOLD_RSL["reward_slots"]
    mine, gold = sorted(map(list, curriculum_schedule(tree))), sorted(GOLD_RSL["curriculum"])
    assert len(mine) == len(gold) and all(_close(a, b) for a, b in zip(mine, gold))
    tasks.register()
    import gymnasium as gym
    assert gym.spec(tasks.RSL_TASK_ID).kwargs["env_cfg_entry_point"]


def test_cat_config_equals_reference_cfg_golden():
    """tasks.cat_config / cat_env_cfg against the reference's own CaT cfg class (config/h12_12dof/cat_env_cfg.py:44-557, flattened in
    tests/golden/cat_cfg_resolved.json): kernel config value by value, reward slots, the ten constraint terms and their
    modify_constraint_p curriculum; the curriculum starts every scheduled term at max_p = 1 / 20 (curriculums.py:27-34)."""
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200.env import (config_to_dict, constraint_curriculum, constraint_max_p, constraint_terms, flatten_cfg, reward_slots)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "cat_cfg_resolved.json")))
    tree = tasks.cat_env_cfg(16)
    mine = config_to_dict(flatten_cfg(tree))
    assert mine == config_to_dict(tasks.cat_config())
    for k, v in gold["kernel_config"].items():
        assert _close(mine[k], v), k
    assert mine["cat_enable"] == 1 and mine["mass_recompute_inertia"] == 0 and abs(mine["velocity_deadzone"] - 0.2) < 1e-7
    assert constraint_terms(tree) == gold["constraint_terms"] and reward_slots(tree) == gold["reward_slots"]
    sched = constraint_curriculum(tree)
    assert sorted(map(list, sched)) == sorted(gold["constraint_curriculum"]) and len(sched) == 9
    p0 = constraint_max_p(sched, mine["cat_max_p"], 0)
    assert p0[0] == 1.0 and all(abs(p - 0.05) < 1e-12 for p in p0[1:])
    p1 = constraint_max_p(sched, mine["cat_max_p"], 10 ** 7)
    assert all(abs(p - 0.25) < 1e-12 for p in p1[1:])
    tasks.register()
    import gymnasium as gym
    assert gym.spec(tasks.CAT_TASK_ID).entry_point.endswith("H1v2CaTEnv") and "clean_rl_cfg_entry_point" in gym.spec(tasks.CAT_TASK_ID).kwargs
    agent = vars(tasks.cat_agent_cfg())
    assert agent == gold["agent"]  # the CleanRL PPO cfg of the reference (agents/clean_rl_ppo_cfg.py), value by value


def test_self_contained_play_ids_equal_the_reference_play_cfgs():
    """tasks.flat_play_env_cfg / rsl_play_env_cfg against the reference's own Play cfg classes (C12/flat_env_cfg.py:51-66,
    C12/rsl_env_cfg.py:543-564), flattened in tests/golden/play_cfg_resolved.json: scene size, noise off, pushes and friction
    randomisation off, fixed forward command for the Rsl one."""
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200.env import config_to_dict, flatten_cfg
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "play_cfg_resolved.json")))
    tasks.register()
    import gymnasium as gym
    for tid, make in (("Isaac-Velocity-Flat-H12_12dof-Play-v0", tasks.flat_play_env_cfg), ("Isaac-Velocity-Rsl-H12_12dof-Play-v0", tasks.rsl_play_env_cfg),
                      ("Isaac-Velocity-CaT-Flat-H12_12dof-Play-v0", tasks.cat_play_env_cfg)):
        tree = make()
        mine = config_to_dict(flatten_cfg(tree))
        assert tree.scene.num_envs == gold[tid]["num_envs"]
        for k, v in gold[tid]["kernel_config"].items():
            assert _close(mine[k], v), (tid, k)
        if "CaT" in tid:  # C12/cat_env_cfg.py:568-583 only shrinks the scene and zeroes the command ranges
            assert mine["cat_enable"] == 1 and mine["cmd_lin_x"] == [0.0, 0.0] and mine["cmd_ang_z"] == [0.0, 0.0]
            assert gym.spec(tid).entry_point.endswith("H1v2CaTEnv")
        else:
            assert mine["enable_corruption"] == 0 and mine["push_enable"] == 0
        assert gym.spec(tid).kwargs["env_cfg_entry_point"]


def test_self_contained_task_roundtrip(cfg):
    """tasks.default_env_cfg() is flatten_cfg's inverse on the default; registration uses the reference's id and kwargs."""
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200.env import config_to_dict, flatten_cfg
    tree = tasks.default_env_cfg(128)
    assert config_to_dict(flatten_cfg(tree)) == config_to_dict(cfg)
    assert tree.scene.num_envs == 128
    tasks.register()
    import gymnasium as gym
    spec = gym.spec(tasks.TASK_ID)
    assert {"env_cfg_entry_point", "rsl_rl_cfg_entry_point"} <= set(spec.kwargs)
    agent = tasks.default_agent_cfg().to_dict()
    for k in ("num_steps_per_env", "save_interval", "experiment_name", "empirical_normalization"):
        assert agent[k] == GOLD["agent"][k], k
    assert agent["policy"] == GOLD["agent"]["policy"] and agent["algorithm"] == GOLD["agent"]["algorithm"]


def test_unsupported_cfg_is_rejected_loudly():
    """A cfg the fused kernel cannot express must raise, not be silently approximated."""
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200.env import flatten_cfg
    tree = tasks.default_env_cfg(8)
    tree.observations.policy.joint_vel.clip = (-1.0, 1.0)
    with pytest.raises(NotImplementedError, match="clip"):
        flatten_cfg(tree)
    tree = tasks.default_env_cfg(8)
    tree.scene.robot.actuators["knees"].max_delay = 2
    with pytest.raises(NotImplementedError, match="min_delay"):
        flatten_cfg(tree)
    tree = tasks.rsl_env_cfg(8)
    tree.observations.policy.history_step = 2
    with pytest.raises(NotImplementedError, match="history_step"):
        flatten_cfg(tree)
    tree = tasks.rsl_env_cfg(8)
    tree.commands.base_velocity.velocity_deadzone = 0.1  # the class default (commands.py:104): supported, balanced per step
    assert abs(flatten_cfg(tree).velocity_deadzone - 0.1) < 1e-7
    tree.commands.base_velocity.velocity_deadzone = -0.1
    with pytest.raises(NotImplementedError, match="velocity_deadzone"):
        flatten_cfg(tree)
    tree = tasks.rsl_env_cfg(8)
    tree.curriculum.feet_slide.func = tree.rewards.feet_slide.func  # anything but modify_reward_weight
    with pytest.raises(NotImplementedError, match="curriculum.feet_slide"):
        flatten_cfg(tree)
    tree = tasks.rsl_env_cfg(8)
    tree.observations.policy.joint_vel.params = {}  # articulation order while the action term preserves the MJCF order
    with pytest.raises(NotImplementedError, match="joint order"):
        flatten_cfg(tree)


def test_env_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from h1v2_isaac_b200 import tasks
    tasks.register()
    import gymnasium as gym
    with pytest.raises(RuntimeError, match="no CUDA device"):
        gym.make(tasks.TASK_ID, cfg=tasks.default_env_cfg(8))


@pytest.mark.skipif(not has_ref, reason="reference tree not present (GPU box)")
def test_reference_cfg_tree_flattens_to_golden():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_cfg_golden.py")], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    again = json.load(open(os.path.join(ROOT, "tests", "golden", "flat_cfg_resolved.json")))
    assert again == GOLD  # regenerating from the reference changes nothing
    assert json.load(open(os.path.join(ROOT, "tests", "golden", "rsl_cfg_resolved.json"))) == GOLD_RSL
    assert json.load(open(os.path.join(ROOT, "tests", "golden", "cat_cfg_resolved.json")))["kernel_config"]["cat_enable"] == 1
    assert set(json.load(open(os.path.join(ROOT, "tests", "golden", "play_cfg_resolved.json")))) == {
        "Isaac-Velocity-Flat-H12_12dof-Play-v0", "Isaac-Velocity-Rsl-H12_12dof-Play-v0", "Isaac-Velocity-CaT-Flat-H12_12dof-Play-v0",
        "Isaac-Velocity-Rough-H12_12dof-Play-v0"}
    assert json.load(open(os.path.join(ROOT, "tests", "golden", "rough_cfg_resolved.json"))) == GOLD_ROUGH


@pytest.mark.skipif(not has_ref, reason="reference tree not present (GPU box)")
def test_unmodified_train_py_reaches_the_backend(tmp_path):
    """scripts/rsl_rl/train.py, untouched, through argparse -> AppLauncher -> hydra cfg -> gym.make -> our env class."""
    import torch
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "h1v2_isaac_b200", "shims"), ROOT, os.path.join(REF, "packages", "biped_tasks"),
                                         os.path.join(REF, "packages", "biped_assets"), os.path.join(REF, "scripts", "rsl_rl")])
    cmd = [sys.executable, os.path.join(REF, "scripts", "rsl_rl", "train.py"), "--task", "Isaac-Velocity-Flat-H12_12dof-v0", "--num_envs", "64",
           "--max_iterations", "1", "--headless", "env.episode_length_s=10.0"]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp_path, env=env, timeout=600)
    if torch.cuda.is_available():
        assert out.returncode == 0, out.stderr[-2000:]
    else:
        assert "H1v2ManagerBasedRLEnv" in out.stderr or "h1v2_isaac_b200/env.py" in out.stderr
        assert "no CUDA device visible" in out.stderr


@pytest.mark.skipif(not has_ref, reason="reference tree not present (GPU box)")
def test_reference_variants_are_accepted_or_refused_by_name():
    """Which of the reference's H1-2 12-dof ids the backend takes: Flat and Flat-Play flatten, also with the
    H12_12DOF_IDEAL robot (IdealPD -> no delay line), and so do Rsl and Rsl-Play (dead-zone command class, second joint-set
    terms, modify_reward_weight curriculum), CaT (constraint manager) and Rough (height field, height scan, base_lin_vel, terrain
    curriculum)."""
    code = r'''
import gymnasium as gym
import biped_tasks.tasks
from isaaclab_tasks.utils import load_cfg_from_registry
from biped_assets.robots.h12 import H12_12DOF_IDEAL
from h1v2_isaac_b200.env import flatten_cfg
out = {}
for tid in ("Isaac-Velocity-Flat-H12_12dof-v0", "Isaac-Velocity-Flat-H12_12dof-Play-v0", "Isaac-Velocity-CaT-Flat-H12_12dof-v0",
            "Isaac-Velocity-Rsl-H12_12dof-v0", "Isaac-Velocity-Rsl-H12_12dof-Play-v0", "Isaac-Velocity-Rough-H12_12dof-v0"):
    cfg = load_cfg_from_registry(tid, "env_cfg_entry_point")
    try:
        c = flatten_cfg(cfg); out[tid] = "ok corruption=%d" % c.enable_corruption + (" class=%d H=%d" % (c.command_class, c.history_length) if "Rsl" in tid else "") + (" terrain=%dx%d scan=%d" % (c.terrain_rows, c.terrain_cols, c.obs_height_scan) if "Rough" in tid else "")
    except NotImplementedError as e:
        out[tid] = "refused: " + str(e)[:40]
cfg = load_cfg_from_registry("Isaac-Velocity-Flat-H12_12dof-v0", "env_cfg_entry_point")
cfg.scene.robot = H12_12DOF_IDEAL.replace(prim_path="{ENV_REGEX_NS}/Robot")
c = flatten_cfg(cfg); out["ideal"] = "ok delays %d %d" % (c.min_delay, c.max_delay)
import json; print("RESULT" + json.dumps(out))
'''
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "h1v2_isaac_b200", "shims"), ROOT, os.path.join(REF, "packages", "biped_tasks"),
                                         os.path.join(REF, "packages", "biped_assets")])
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.split("RESULT")[1])
    assert res["Isaac-Velocity-Flat-H12_12dof-v0"] == "ok corruption=1"
    assert res["Isaac-Velocity-Flat-H12_12dof-Play-v0"] == "ok corruption=0"
    assert res["ideal"] == "ok delays 0 0"
    assert res["Isaac-Velocity-CaT-Flat-H12_12dof-v0"] == "ok corruption=1"  # constraints group -> the CaT tail (h1v2_cat_step)
    assert res["Isaac-Velocity-Rsl-H12_12dof-v0"] == "ok corruption=1 class=1 H=6"
    assert res["Isaac-Velocity-Rsl-H12_12dof-Play-v0"] == "ok corruption=0 class=1 H=6"
    assert res["Isaac-Velocity-Rough-H12_12dof-v0"] == "ok corruption=1 terrain=10x20 scan=1"


@pytest.mark.skipif(not has_ref, reason="reference tree not present (GPU box)")
def test_reference_deploy_exporter_on_shim_cfg_matches_shipped_yaml():
    """Train -> deploy wire format: the reference's own get_deploy_config (utils/mdp/config_exporter.py:27-58)
    runs unmodified on a cfg tree built from the shim classes and reproduces the env.yaml the reference ships for its deployed
    policy (scripts/deploy/policies/demo_rsl/env.yaml): control rate, history, action scale, observation list with scales,
    per-joint kp / kd / default pose."""
    code = r'''
import json, yaml
import gymnasium as gym
import biped_tasks.tasks
from isaaclab_tasks.utils import load_cfg_from_registry
from biped_tasks.utils.mdp.config_exporter import get_deploy_config
d = get_deploy_config(load_cfg_from_registry("Isaac-Velocity-CaT-Flat-H12_12dof-v0", "env_cfg_entry_point"))
y = yaml.load(open("/root/reference/scripts/deploy/policies/demo_rsl/env.yaml"), Loader=yaml.UnsafeLoader)
print("RESULT" + json.dumps({"mine": d, "shipped": y}))
'''
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "h1v2_isaac_b200", "shims"), ROOT, os.path.join(REF, "packages", "biped_tasks"),
                                         os.path.join(REF, "packages", "biped_assets")])
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.split("RESULT")[1])
    mine, shipped = r["mine"], r["shipped"]
    for k in ("control_dt", "history_length", "history_step", "action_scale", "command_ranges"):
        assert mine[k] == shipped[k], k
    assert [(o["name"], o.get("scale") or 1) for o in mine["observations"]] == [(o["name"], o.get("scale") or 1) for o in shipped["observations"]]
    ship_j = {j["name"]: j for j in shipped["joints"]}
    assert len(mine["joints"]) == 12
    for j in mine["joints"]:
        s = ship_j[j["name"]]
        assert (j["kp"], j["kd"], j["default_joint_pos"]) == (s["kp"], s["kd"], s["default_joint_pos"]), j["name"]


def test_deploy_config_of_the_rsl_task():
    """h1v2_isaac_b200.deploy.deploy_config: the env.yaml document for a policy trained on the Rsl id (the reference's own exporter
    raises AttributeError on that cfg: no history_step).  Structure and values; the Flat id (articulation joint order) is refused
    like the reference's exporter refuses cfgs without preserve_order."""
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200.deploy import deploy_config
    d = deploy_config(tasks.rsl_env_cfg(8))
    assert (d["control_dt"], d["history_length"], d["history_step"], d["action_scale"], d["velocity_deadzone"]) == (0.02, 6, 1, 0.25, 0.0)
    assert d["command_ranges"] == {"lin_vel_x": [-1.0, 1.0], "lin_vel_y": [-1.0, 1.0], "ang_vel_z": [-1.0, 1.0]}
    assert [(o["name"], o["scale"]) for o in d["observations"]] == [("base_ang_vel", 0.25), ("projected_gravity", 1), ("generated_commands", 1),
                                                                   ("joint_pos_rel", 1), ("joint_vel_rel", 0.05), ("last_action", 1)]
    assert [j["name"] for j in d["joints"]][:4] == ["left_hip_yaw_joint", "left_hip_pitch_joint", "left_hip_roll_joint", "left_knee_joint"]
    assert d["joints"][3] == {"name": "left_knee_joint", "kp": 300.0, "kd": 4.0, "default_joint_pos": 0.36, "enabled": True}
    with pytest.raises(ValueError, match="preserve_order"):
        deploy_config(tasks.default_env_cfg(8))


@pytest.mark.skipif(not has_ref, reason="reference tree not present (GPU box)")
def test_deploy_config_matches_the_env_yaml_the_reference_ships(tmp_path):
    """... and it is the document the reference ships next to its Rsl-trained policy (scripts/deploy/policies/demo_rsl/env.yaml):
    every key equal; the shipped file additionally lists the 15 disabled upper-body joints of the real robot."""
    import yaml
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200.deploy import write_deploy_config
    mine = write_deploy_config(tasks.rsl_env_cfg(8), str(tmp_path / "env.yaml"))
    assert yaml.safe_load(open(tmp_path / "env.yaml")) == mine
    shipped = yaml.load(open(os.path.join(REF, "scripts", "deploy", "policies", "demo_rsl", "env.yaml")), Loader=yaml.UnsafeLoader)
    for k in ("control_dt", "history_length", "history_step", "action_scale", "velocity_deadzone", "command_ranges"):
        assert mine[k] == shipped[k], k
    assert [(o["name"], o.get("scale") or 1) for o in mine["observations"]] == [(o["name"], o.get("scale") or 1) for o in shipped["observations"]]
    assert mine["joints"] == shipped["joints"][:12]
    assert all(not j["enabled"] for j in shipped["joints"][12:])


def test_shim_cfg_classes_reject_misspelled_fields():
    """The shim's configclass is as strict as upstream's dataclass-based one: a misspelled constructor keyword raises instead of becoming a
    silent attribute that flatten_cfg would never read (round-1 review: lenient placeholders)."""
    from h1v2_isaac_b200 import shims
    shims.install()
    from isaaclab.managers import RewardTermCfg
    from isaaclab.sensors import RayCasterCfg, patterns
    from isaaclab.terrains import HfRandomUniformTerrainCfg, TerrainGeneratorCfg
    with pytest.raises(TypeError, match="wieght"):
        RewardTermCfg(func=None, wieght=1.0)
    with pytest.raises(TypeError, match="noise_rnage"):
        HfRandomUniformTerrainCfg(noise_rnage=(0.0, 0.02), noise_step=0.005)
    with pytest.raises(TypeError, match="num_row"):
        TerrainGeneratorCfg(size=(8.0, 8.0), num_row=10, sub_terrains={})
    with pytest.raises(TypeError, match="resolutoin"):
        patterns.GridPatternCfg(resolutoin=0.1, size=[1.6, 1.0])
    ok = RayCasterCfg(prim_path="{ENV_REGEX_NS}/Robot/torso_link", offset=RayCasterCfg.OffsetCfg(pos=(0.0, 0.0, 20.0)), attach_yaw_only=True,
                      pattern_cfg=patterns.GridPatternCfg(resolution=0.1, size=[1.6, 1.0]), mesh_prim_paths=["/World/ground"])
    assert ok.offset.pos == (0.0, 0.0, 20.0) and ok.pattern_cfg.ordering == "xy"
