"""Self-contained registration of the Flat and Rsl H1-2 tasks for machines that have neither the reference's `biped_tasks`
package nor isaaclab (the GPU test box): the same gym ids, the same entry-point kwargs, cfg trees of the same shape.

The trees are generated FROM the kernel's resolved configs (`h1v2_default_config`, `h1v2_rsl_config`, every value of which
cites the reference line that pins it) by inverting `env.flatten_cfg`, so `flatten_cfg(default_env_cfg()) == default_config()`
(and the same for Rsl) by construction; asserted in tests/test_boundary.py.  When `biped_tasks` is importable its own
registration (packages/biped_tasks/.../config/h12_12dof/__init__.py:40-49,85-93) wins and this module does nothing.
"""
from __future__ import annotations

TASK_ID = "Isaac-Velocity-Flat-H12_12dof-v0"
RSL_TASK_ID = "Isaac-Velocity-Rsl-H12_12dof-v0"
CAT_TASK_ID = "Isaac-Velocity-CaT-Flat-H12_12dof-v0"
ROUGH_TASK_ID = "Isaac-Velocity-Rough-H12_12dof-v0"
# reward term names of C12/cat_env_cfg.py:306-333 by kernel slot
CAT_REW_NAMES = {14: "track_lin_vel_xy_exp", 15: "track_ang_vel_z_exp", 8: "dof_torques_l2", 9: "joint_acc_l2", 17: "joint_vel_l2", 10: "action_rate_l2",
                 6: "joint_deviation_l1"}
# reward term names of C12/rsl_env_cfg.py:278-407 by kernel slot (the Flat tree uses _capi.REW_NAMES)
RSL_REW_NAMES = {14: "track_lin_vel_xy_exp", 15: "track_ang_vel_z_exp", 3: "feet_air_time", 4: "feet_slide", 11: "flat_orientation",
                 18: "base_height_l2", 8: "joint_torques_l2", 17: "joint_vel_l2", 9: "dof_acc_l2", 6: "joint_deviation_hip",
                 21: "joint_deviation_ankle", 5: "joint_pos_limits_ankle", 20: "joint_pos_limits_hip", 10: "action_rate_l2",
                 19: "contact_forces", 0: "termination_penalty"}
# C12/rsl_env_cfg.py:447-497: modify_reward_weight on these terms after 24 * 5000 steps, to the weight they already have
RSL_CURRICULUM = ["flat_orientation", "joint_torques_l2", "joint_vel_l2", "dof_acc_l2", "joint_deviation_hip", "joint_deviation_ankle",
                  "joint_pos_limits_ankle", "joint_pos_limits_hip", "contact_forces", "feet_air_time", "feet_slide", "base_height_l2"]


class UniformVelocityCommandWithDeadzone:
    """Name-only stand-in for T/utils/mdp/commands.py:19 (flatten_cfg keys on class_type.__name__; the logic is in the kernel)."""


def default_env_cfg(num_envs: int = 4096, device: str = "cuda:0"):
    from ._capi import default_config
    return env_cfg_from_config(default_config(), num_envs, device)


def rsl_env_cfg(num_envs: int = 4096, device: str = "cuda:0"):
    from ._capi import rsl_config
    return env_cfg_from_config(rsl_config(), num_envs, device, rew_names=RSL_REW_NAMES, curriculum=RSL_CURRICULUM, curriculum_steps=24 * 5000)


def flat_play_env_cfg(num_envs: int = 50, device: str = "cuda:0"):
    """Isaac-Velocity-Flat-H12_12dof-Play-v0 (C12/flat_env_cfg.py:51-66): 50 envs, no observation noise, no pushes."""
    from ._capi import default_config
    c = default_config()
    c.enable_corruption = 0
    c.push_enable = 0
    return env_cfg_from_config(c, num_envs, device)


def rsl_play_env_cfg(num_envs: int = 100, device: str = "cuda:0"):
    """Isaac-Velocity-Rsl-H12_12dof-Play-v0 (C12/rsl_env_cfg.py:543-564): 100 envs, no noise, no pushes, no friction
    randomisation (the asset's own material), forward command 0.5 m/s."""
    from ._capi import default_config, rsl_config
    c, d = rsl_config(), default_config()
    c.enable_corruption = 0
    c.push_enable = 0
    c.friction, c.solver_iterations = d.friction, d.solver_iterations
    c.friction_range[0] = c.friction_range[1] = d.friction
    c.cmd_lin_x[0] = c.cmd_lin_x[1] = 0.5
    c.cmd_lin_y[0] = c.cmd_lin_y[1] = 0.0
    c.cmd_ang_z[0] = c.cmd_ang_z[1] = 0.0
    return env_cfg_from_config(c, num_envs, device, rew_names=RSL_REW_NAMES, curriculum=RSL_CURRICULUM, curriculum_steps=24 * 5000)


def rough_terrains_cfg(c=None):
    """The reference's in-tree terrain generator cfg (packages/biped_tasks/biped_tasks/utils/mdp/terrains.py:11-28) rebuilt from the
    resolved kernel config, for machines without biped_tasks."""
    from . import shims
    shims.install()
    from isaaclab.terrains import HfRandomUniformTerrainCfg, TerrainGeneratorCfg
    if c is None:
        from ._capi import rough_config
        c = rough_config()
    hs, vs = float(c.terrain_hscale), float(c.terrain_vscale)
    return TerrainGeneratorCfg(
        size=(c.terrain_tile_size, c.terrain_tile_size), border_width=20.0, num_rows=c.terrain_rows, num_cols=c.terrain_cols, horizontal_scale=hs,
        vertical_scale=vs, slope_threshold=0.75, use_cache=False, curriculum=bool(c.terrain_curriculum),
        sub_terrains={"random_rough": HfRandomUniformTerrainCfg(
            proportion=1.0, noise_range=(c.terrain_level_min * vs, c.terrain_level_max * vs), noise_step=c.terrain_level_step * vs,
            border_width=(c.terrain_border_px - 0.5) * hs if c.terrain_border_px > 0 else 0.0)})


def rough_env_cfg(num_envs: int = 4096, device: str = "cuda:0"):
    """Isaac-Velocity-Rough-H12_12dof-v0 (C12/rough_env_cfg.py:65-125) with the in-tree terrain generator cfg."""
    from ._capi import rough_config
    return env_cfg_from_config(rough_config(), num_envs, device)


def rough_play_env_cfg(num_envs: int = 50, device: str = "cuda:0"):
    """Isaac-Velocity-Rough-H12_12dof-Play-v0 (C12/rough_env_cfg.py:128-156): 50 envs, 40 s episodes, 5 x 5 tiles without curriculum and
    envs spread over all levels, forward command 1 m/s, heading 0, no observation noise."""
    from ._capi import rough_config
    c = rough_config()
    c.episode_length_s = 40.0
    c.terrain_max_init_level = -1
    c.terrain_rows = c.terrain_cols = 5
    c.terrain_curriculum = 0
    c.cmd_lin_x[0] = c.cmd_lin_x[1] = 1.0
    c.cmd_heading[0] = c.cmd_heading[1] = 0.0
    c.enable_corruption = 0
    return env_cfg_from_config(c, num_envs, device)


def cat_config():
    """Resolved cfg of Isaac-Velocity-CaT-Flat-H12_12dof-v0: h1v2_cat_config (csrc/h1v2_config.cpp cites C12/cat_env_cfg.py line by line);
    checked value by value against the reference's own cfg class in tests/test_boundary.py (tests/golden/cat_cfg_resolved.json)."""
    from ._capi import cat_config as _cat_config
    return _cat_config()


def cat_env_cfg(num_envs: int = 4096, device: str = "cuda:0"):
    return env_cfg_from_config(cat_config(), num_envs, device, rew_names=CAT_REW_NAMES)


def cat_play_env_cfg(num_envs: int = 100, device: str = "cuda:0"):
    """Isaac-Velocity-CaT-Flat-H12_12dof-Play-v0 (C12/cat_env_cfg.py:568-583): 100 envs, every command range (0, 0)."""
    c = cat_config()
    for r in (c.cmd_lin_x, c.cmd_lin_y, c.cmd_ang_z):
        r[0] = r[1] = 0.0
    return env_cfg_from_config(c, num_envs, device, rew_names=CAT_REW_NAMES)


def env_cfg_from_config(c, num_envs: int = 4096, device: str = "cuda:0", rew_names=None, curriculum=(), curriculum_steps=0):
    """Inverse of env.flatten_cfg: an H1v2Config -> a ManagerBasedRLEnvCfg-shaped tree built from the shim cfg classes."""
    from . import shims
    shims.install()
    import isaaclab.envs as ienvs
    from isaaclab.actuators import DelayedPDActuatorCfg, IdealPDActuatorCfg
    from isaaclab.assets import ArticulationCfg
    from isaaclab.managers import EventTermCfg, ObservationGroupCfg, ObservationTermCfg, RewardTermCfg, SceneEntityCfg, TerminationTermCfg
    from isaaclab.scene import InteractiveSceneCfg
    from isaaclab.utils.noise import AdditiveUniformNoiseCfg as Unoise

    from .env import JOINT_NAMES, REW_FUNC_SLOT, REW_FUNC_SLOT_B, SLOT_BODIES
    from .shims._lenient import Placeholder

    mdp = ienvs.mdp
    import isaaclab_tasks  # noqa: F401  (builds the locomotion mdp namespace)
    import isaaclab_tasks.manager_based.locomotion.velocity.mdp as lmdp

    def names(mask, table):
        return [table[i] for i in range(len(table)) if (mask >> i) & 1]

    class bag:
        """Ordered attribute group with the two configclass methods the env reads."""

        def __init__(self, **kw):
            self.__dict__.update(kw)

        def __configclass_fields__(self):
            return list(self.__dict__)

        def to_dict(self):
            return {k: (v.to_dict() if hasattr(v, "to_dict") else v) for k, v in self.__dict__.items()}

    groups = {"legs": [0, 1, 2, 6, 7, 8], "knees": [3, 9], "feet": [4, 5, 10, 11]}
    delayed = c.max_delay > 0  # A/robots/h12.py:18-114 DelayedPDActuatorCfg (Flat) vs :117-206 IdealPDActuatorCfg (Rsl)
    actuators = {
        g: (DelayedPDActuatorCfg if delayed else IdealPDActuatorCfg)(
            joint_names_expr=[JOINT_NAMES[i] for i in ids], effort_limit={JOINT_NAMES[i]: c.effort_limit[i] for i in ids},
            velocity_limit=(c.joint_vel_limit if c.joint_vel_limit > 0 else None), stiffness={JOINT_NAMES[i]: c.kp[i] for i in ids}, damping={JOINT_NAMES[i]: c.kd[i] for i in ids},
            armature=c.dof_armature[6 + ids[0]], friction=0.0, **({"min_delay": c.min_delay, "max_delay": c.max_delay} if delayed else {}))
        for g, ids in groups.items()}
    robot = ArticulationCfg(prim_path="{ENV_REGEX_NS}/Robot", init_state=ArticulationCfg.InitialStateCfg(
        pos=(0.0, 0.0, c.init_root_height), joint_pos={JOINT_NAMES[i]: c.default_joint_pos[i] for i in range(12)}, joint_vel={".*": 0.0}),
        soft_joint_pos_limit_factor=c.soft_limit_factor, actuators=actuators)
    scene = InteractiveSceneCfg(num_envs=num_envs, env_spacing=c.env_spacing)
    object.__setattr__(scene, "robot", robot)
    ground = Placeholder(static_friction=1.0, dynamic_friction=1.0)
    if c.terrain_enable:  # V/velocity_env_cfg.py:40-58
        from isaaclab.terrains import TerrainImporterCfg
        object.__setattr__(scene, "terrain", TerrainImporterCfg(
            prim_path="/World/ground", terrain_type="generator", terrain_generator=rough_terrains_cfg(c),
            max_init_terrain_level=(None if c.terrain_max_init_level < 0 else c.terrain_max_init_level), physics_material=ground))
    else:
        object.__setattr__(scene, "terrain", Placeholder(terrain_type="plane", physics_material=ground))
    if c.obs_height_scan:  # V/velocity_env_cfg.py:61-68, C12/rough_env_cfg.py:74-75
        from isaaclab.sensors import RayCasterCfg, patterns
        object.__setattr__(scene, "height_scanner", RayCasterCfg(
            prim_path="{ENV_REGEX_NS}/Robot/torso_link", offset=RayCasterCfg.OffsetCfg(pos=(0.0, 0.0, 20.0)), attach_yaw_only=True,
            pattern_cfg=patterns.GridPatternCfg(resolution=c.scan_resolution, size=[c.scan_size[0], c.scan_size[1]]), mesh_prim_paths=["/World/ground"]))

    def obs(func, n=0.0, s=1.0, joints=False):
        t = ObservationTermCfg(func=func, noise=Unoise(n_min=-n, n_max=n) if n else None, scale=None if s == 1.0 else s)
        if joints and list(c.joint_perm) != [0, 6, 1, 7, 2, 8, 3, 9, 4, 10, 5, 11]:  # not the articulation's own order: name it (C12/rsl_env_cfg.py:148-192)
            t.params = {"asset_cfg": SceneEntityCfg("robot", joint_names=[JOINT_NAMES[j] for j in c.joint_perm], preserve_order=True)}
        return t

    rough_obs = bool(c.obs_base_lin_vel or c.obs_height_scan)
    policy = ObservationGroupCfg(concatenate_terms=True, enable_corruption=bool(c.enable_corruption), history_length=(0 if rough_obs else c.history_length))
    lead, tail = [], []
    if c.obs_base_lin_vel:  # V/velocity_env_cfg.py:123
        lead = [("base_lin_vel", obs(mdp.base_lin_vel, c.noise_lin_vel, c.scale_lin_vel))]
    if c.obs_height_scan:   # V/velocity_env_cfg.py:133-138
        t = obs(mdp.height_scan, c.noise_height_scan, c.scale_height_scan)
        t.params = {"sensor_cfg": SceneEntityCfg("height_scanner")}
        t.clip = (c.scan_clip[0], c.scan_clip[1])
        tail = [("height_scan", t)]
    for k, v in lead + [("base_ang_vel", obs(mdp.base_ang_vel, c.noise_ang_vel, c.scale_ang_vel)),
                 ("projected_gravity", obs(mdp.projected_gravity, c.noise_gravity, c.scale_gravity)),
                 ("velocity_commands", obs(mdp.generated_commands, 0.0, c.scale_cmd)),
                 ("joint_pos", obs(mdp.joint_pos_rel, c.noise_joint_pos, c.scale_joint_pos, True)),
                 ("joint_vel", obs(mdp.joint_vel_rel, c.noise_joint_vel, c.scale_joint_vel, True)),
                 ("actions", obs(mdp.last_action, 0.0, c.scale_action))] + tail:
        object.__setattr__(policy, k, v)
    policy.__configclass_fields__ = lambda: ["concatenate_terms", "enable_corruption", "history_length"] + [k for k, _ in lead] + [
        "base_ang_vel", "projected_gravity", "velocity_commands", "joint_pos", "joint_vel", "actions"] + [k for k, _ in tail]

    feet = SceneEntityCfg("contact_forces", body_names=SLOT_BODIES[:2])
    rew_params = {
        "track_lin_vel_xy_yaw_frame_exp": {"command_name": "base_velocity", "std": c.track_std},
        "track_ang_vel_z_world_exp": {"command_name": "base_velocity", "std": c.track_std},
        "track_lin_vel_xy_exp": {"command_name": "base_velocity", "std": c.track_std},
        "track_ang_vel_z_exp": {"command_name": "base_velocity", "std": c.track_std},
        "feet_air_time_positive_biped": {"command_name": "base_velocity", "sensor_cfg": feet, "threshold": c.feet_air_threshold},
        "feet_air_time": {"command_name": "base_velocity", "sensor_cfg": feet, "threshold": c.feet_air_threshold},
        "feet_slide": {"sensor_cfg": feet, "asset_cfg": SceneEntityCfg("robot", body_names=SLOT_BODIES[:2])},
        "joint_pos_limits": {"asset_cfg": SceneEntityCfg("robot", joint_names=names(c.mask_pos_limits, JOINT_NAMES))},
        "joint_deviation_l1": {"asset_cfg": SceneEntityCfg("robot", joint_names=names(c.mask_joint_dev, JOINT_NAMES))},
        "joint_torques_l2": {"asset_cfg": SceneEntityCfg("robot", joint_names=names(c.mask_torques, JOINT_NAMES))},
        "undesired_contacts": {"sensor_cfg": SceneEntityCfg("contact_forces", body_names=names(c.mask_undesired_slots, SLOT_BODIES)), "threshold": 1.0},
        "contact_forces": {"sensor_cfg": SceneEntityCfg("contact_forces", body_names=names(c.mask_contact_forces_slots, SLOT_BODIES)), "threshold": c.contact_forces_threshold},
        "base_height_l2": {"target_height": c.base_height_target},
    }
    rew_params_b = {
        "joint_pos_limits": {"asset_cfg": SceneEntityCfg("robot", joint_names=names(c.mask_pos_limits_b, JOINT_NAMES))},
        "joint_deviation_l1": {"asset_cfg": SceneEntityCfg("robot", joint_names=names(c.mask_joint_dev_b, JOINT_NAMES))},
    }
    slot_func = {s: (f, rew_params) for f, s in REW_FUNC_SLOT.items()}
    slot_func.update({s: (f, rew_params_b) for f, s in REW_FUNC_SLOT_B.items()})
    from ._capi import REW_NAMES
    rew = {}
    for s, w in enumerate(c.rew_weight):  # ascending slots: a function's first slot precedes its second, as reward_slots assigns them
        if w != 0.0:
            f, params = slot_func[s]
            rew[(rew_names or {}).get(s, REW_NAMES[s])] = RewardTermCfg(func=getattr(lmdp, f), weight=w, params=params.get(f, {}))
    rewards = bag(**rew)
    cur = bag(**{n: ienvs_managers().CurriculumTermCfg(func=mdp.modify_reward_weight, params={
        "term_name": n, "weight": rew[n].weight, "num_steps": curriculum_steps}) for n in curriculum}) if curriculum else None

    if c.terrain_enable and c.terrain_curriculum:  # V/velocity_env_cfg.py:275
        cur = bag(**{**(cur.__dict__ if cur is not None else {}), "terrain_levels": ienvs_managers().CurriculumTermCfg(func=lmdp.terrain_levels_vel)})

    terminations = bag(
        time_out=TerminationTermCfg(func=mdp.time_out, time_out=True),
        base_contact=TerminationTermCfg(func=mdp.illegal_contact, params={
            "sensor_cfg": SceneEntityCfg("contact_forces", body_names=names(c.mask_illegal_slots, SLOT_BODIES)), "threshold": c.contact_threshold}))
    ranges = ienvs.UniformVelocityCommandCfg.Ranges(lin_vel_x=tuple(c.cmd_lin_x), lin_vel_y=tuple(c.cmd_lin_y), ang_vel_z=tuple(c.cmd_ang_z),
                                                    heading=tuple(c.cmd_heading))
    base_velocity = ienvs.UniformVelocityCommandCfg(
        asset_name="robot", resampling_time_range=tuple(c.cmd_resample_time), rel_standing_envs=c.rel_standing_envs, rel_heading_envs=c.rel_heading_envs,
        heading_command=bool(c.heading_command), heading_control_stiffness=c.heading_stiffness, ranges=ranges)
    if c.command_class == 1:  # C12/rsl_env_cfg.py:86-99 UniformVelocityCommandWithDeadzoneCfg
        object.__setattr__(base_velocity, "class_type", UniformVelocityCommandWithDeadzone)
        object.__setattr__(base_velocity, "velocity_deadzone", c.velocity_deadzone)
    commands = bag(base_velocity=base_velocity)
    axes = ["x", "y", "z", "roll", "pitch", "yaw"]
    ev = dict(
        physics_material=EventTermCfg(func=mdp.randomize_rigid_body_material, mode="startup", params={
            "static_friction_range": tuple(c.friction_range), "dynamic_friction_range": tuple(c.friction_range), "restitution_range": (0.0, 0.0)}),
        reset_base=EventTermCfg(func=mdp.reset_root_state_uniform, mode="reset", params={
            "pose_range": {a: tuple(c.reset_pose_range[i]) for i, a in enumerate(axes)},
            "velocity_range": {a: tuple(c.reset_vel_range[i]) for i, a in enumerate(axes)}}),
        reset_robot_joints=EventTermCfg(func=mdp.reset_joints_by_scale, mode="reset", params={
            "position_range": tuple(c.reset_joint_pos_scale), "velocity_range": tuple(c.reset_joint_vel_scale)}))
    if c.mass_add_range[0] != 0.0 or c.mass_add_range[1] != 0.0:
        ev["add_base_mass"] = EventTermCfg(func=mdp.randomize_rigid_body_mass, mode="startup", params={
            "mass_distribution_params": tuple(c.mass_add_range), "operation": "add", "recompute_inertia": bool(c.mass_recompute_inertia)})
    if c.push_enable:
        ev["push_robot"] = EventTermCfg(func=mdp.push_by_setting_velocity, mode="interval", interval_range_s=tuple(c.push_interval_s),
                                        params={"velocity_range": {"x": tuple(c.push_vel_xy), "y": tuple(c.push_vel_xy)}})
    perm = list(c.joint_perm)
    action = ienvs.JointPositionActionCfg(asset_name="robot", joint_names=[JOINT_NAMES[j] for j in perm], scale=c.action_scale,
                                          use_default_offset=True, preserve_order=True)
    cons = None
    if c.cat_enable:  # the ten ConstraintTerms of C12/cat_env_cfg.py:336-431 and their modify_constraint_p curriculum (:462-519)
        import isaaclab_tasks.manager_based.locomotion.velocity.mdp as cmdp  # named stand-ins for utils/cat/constraints.py functions
        from ._capi import CSTR_NAMES
        robot_all = SceneEntityCfg("robot", joint_names=[".*"])
        feet_s = SceneEntityCfg("contact_forces", body_names=SLOT_BODIES[:2])
        cparams = {
            "contact": {"asset_cfg": SceneEntityCfg("contact_forces", body_names=names(c.cat_contact_slots, SLOT_BODIES))},
            "joint_position_limits": {"asset_cfg": robot_all}, "joint_velocity_limits": {"asset_cfg": robot_all}, "joint_torque_limits": {"asset_cfg": robot_all},
            "foot_contact_force": {"limit": c.cat_foot_force_limit, "asset_cfg": feet_s},
            "no_move": {"velocity_deadzone": c.cat_no_move_deadzone, "joint_vel_limit": c.cat_no_move_vel_limit, "asset_cfg": robot_all},
            "base_orientation": {"limit": c.cat_orientation_limit, "asset_cfg": SceneEntityCfg("robot")},
            "base_height": {"asset_cfg": SceneEntityCfg("robot"), "height": c.cat_height, "std": c.cat_height_std},
            "foot_contact": {"asset_cfg": feet_s},
            "foot_clearance": {"min_height": c.cat_clearance_min_height, "velocity_deadzone": c.cat_clearance_deadzone,
                               "pos_asset_cfg": SceneEntityCfg("robot", body_names=SLOT_BODIES[:2]), "contact_asset_cfg": feet_s}}
        cons = bag(**{n: Placeholder(func=getattr(cmdp, n), max_p=float(c.cat_max_p[i]), params=cparams[n]) for i, n in enumerate(CSTR_NAMES) if c.cat_max_p[i] > 0})
        sched = {n: Placeholder(func=cmdp.modify_constraint_p, params={"term_name": n, "num_steps": 24 * 5000, "init_max_p": float(c.cat_max_p[i])})
                 for i, n in enumerate(CSTR_NAMES) if i > 0 and c.cat_max_p[i] > 0}
        cur = bag(**{**(cur.__dict__ if cur is not None else {}), **sched})
    cfg = ienvs.ManagerBasedRLEnvCfg(decimation=c.decimation, episode_length_s=c.episode_length_s, scene=scene,
                                     observations=bag(policy=policy), actions=bag(joint_pos=action), rewards=rewards,
                                     terminations=terminations, commands=commands, events=bag(**ev), curriculum=cur)
    if cons is not None:
        object.__setattr__(cfg, "constraints", cons)
    cfg.sim.dt = c.sim_dt
    cfg.sim.device = device
    cfg.sim.gravity = (0.0, 0.0, -c.gravity)
    return cfg


def default_agent_cfg():
    """rsl_rl runner cfg of the Flat id (values: config/h12_12dof/agents/rsl_rl_ppo_cfg.py:11-47)."""
    from . import shims
    shims.install()
    from isaaclab_rl.rsl_rl import RslRlOnPolicyRunnerCfg, RslRlPpoActorCriticCfg, RslRlPpoAlgorithmCfg
    return RslRlOnPolicyRunnerCfg(
        num_steps_per_env=24, max_iterations=3000, save_interval=100, experiment_name="h12_12dof_flat", empirical_normalization=False,
        policy=RslRlPpoActorCriticCfg(init_noise_std=1.0, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128], activation="elu"),
        algorithm=RslRlPpoAlgorithmCfg(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.0081, num_learning_epochs=5,
                                       num_mini_batches=4, learning_rate=1.0e-3, schedule="adaptive", gamma=0.99, lam=0.95, desired_kl=0.01,
                                       max_grad_norm=1.0))


def rough_agent_cfg():
    """rsl_rl runner cfg of the Rough id (config/h12_12dof/agents/rsl_rl_ppo_cfg.py:10-37): the Flat cfg's values, experiment h12_12dof_rough."""
    a = default_agent_cfg()
    a.experiment_name = "h12_12dof_rough"
    return a


def cat_agent_cfg():
    """CleanRL PPO cfg of the CaT id (values: config/h12_12dof/agents/clean_rl_ppo_cfg.py, schema utils/cleanrl/rl_cfg.py:11-38)."""
    import types
    return types.SimpleNamespace(
        seed=42, save_interval=200, learning_rate=3.0e-4, num_steps=24, num_iterations=50000, gamma=0.99, gae_lambda=0.95, updates_epochs=5,
        minibatch_size=16384, clip_coef=0.2, ent_coef=0.0081, vf_coef=2.0, max_grad_norm=1.0, norm_adv=True, clip_vloss=True, anneal_lr=True,
        experiment_name="h12_12dof_flat", logger="tensorboard", wandb_project="h12_12dof_flat", load_run=".*", load_checkpoint="model_.*.pt")


def ienvs_managers():
    import isaaclab.managers as m
    return m


def register() -> bool:
    """gym.register the Flat and Rsl ids with this package's own entry points unless something registered them already."""
    from . import shims
    shims.install()
    import gymnasium as gym
    done = False
    for tid, env_cfg in ((TASK_ID, "default_env_cfg"), (RSL_TASK_ID, "rsl_env_cfg"), (CAT_TASK_ID, "cat_env_cfg"),
                         (ROUGH_TASK_ID, "rough_env_cfg"), (ROUGH_TASK_ID.replace("-v0", "-Play-v0"), "rough_play_env_cfg"),  # C12/__init__.py:17-36
                         (TASK_ID.replace("-v0", "-Play-v0"), "flat_play_env_cfg"), (RSL_TASK_ID.replace("-v0", "-Play-v0"), "rsl_play_env_cfg"),
                         (CAT_TASK_ID.replace("-v0", "-Play-v0"), "cat_play_env_cfg")):  # C12/__init__.py:76-82
        try:
            gym.spec(tid)
            continue
        except Exception:
            pass
        # both ids use the same runner cfg (C12/__init__.py:47,91: rsl_rl_ppo_cfg:H12_12dof_FlatPPORunnerCfg)
        gym.register(id=tid, entry_point="h1v2_isaac_b200.env:" + ("H1v2CaTEnv" if "CaT" in tid else "H1v2ManagerBasedRLEnv"), disable_env_checker=True,
                     kwargs={"env_cfg_entry_point": f"h1v2_isaac_b200.tasks:{env_cfg}",
                             **({"clean_rl_cfg_entry_point": "h1v2_isaac_b200.tasks:cat_agent_cfg"} if "CaT" in tid  # C12/__init__.py:63-82
                                else {"rsl_rl_cfg_entry_point": "h1v2_isaac_b200.tasks:" + ("rough_agent_cfg" if "Rough" in tid else "default_agent_cfg")})})
        done = True
    return done
