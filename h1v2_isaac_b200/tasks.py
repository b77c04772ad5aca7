"""Self-contained registration of the Flat H1-2 task for machines that have neither the reference's `biped_tasks`
package nor isaaclab (the GPU test box): the same gym id, the same entry-point kwargs, a cfg tree of the same shape.

The tree is generated FROM the kernel's resolved default (`h1v2_default_config`, every value of which cites the reference
line that pins it) by inverting `env.flatten_cfg`, so `flatten_cfg(default_env_cfg()) == default_config()` by
construction and is asserted in tests/test_boundary.py.  When `biped_tasks` is importable its own registration
(packages/biped_tasks/.../config/h12_12dof/__init__.py:40-49) wins and this module does nothing.
"""
from __future__ import annotations

TASK_ID = "Isaac-Velocity-Flat-H12_12dof-v0"


def default_env_cfg(num_envs: int = 4096, device: str = "cuda:0"):
    from . import shims
    shims.install()
    import isaaclab.envs as ienvs
    from isaaclab.actuators import DelayedPDActuatorCfg
    from isaaclab.assets import ArticulationCfg
    from isaaclab.managers import EventTermCfg, ObservationGroupCfg, ObservationTermCfg, RewardTermCfg, SceneEntityCfg, TerminationTermCfg
    from isaaclab.scene import InteractiveSceneCfg
    from isaaclab.utils.noise import AdditiveUniformNoiseCfg as Unoise

    from ._capi import default_config
    from .env import JOINT_NAMES, REW_FUNC_SLOT, SLOT_BODIES
    from .shims._lenient import Placeholder

    c = default_config()
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
    actuators = {
        g: DelayedPDActuatorCfg(joint_names_expr=[JOINT_NAMES[i] for i in ids], effort_limit={JOINT_NAMES[i]: c.effort_limit[i] for i in ids},
                                velocity_limit=(c.joint_vel_limit if c.joint_vel_limit > 0 else None), stiffness={JOINT_NAMES[i]: c.kp[i] for i in ids}, damping={JOINT_NAMES[i]: c.kd[i] for i in ids},
                                armature=c.dof_armature[6 + ids[0]], friction=0.0, min_delay=c.min_delay, max_delay=c.max_delay)
        for g, ids in groups.items()}
    robot = ArticulationCfg(prim_path="{ENV_REGEX_NS}/Robot", init_state=ArticulationCfg.InitialStateCfg(
        pos=(0.0, 0.0, c.init_root_height), joint_pos={JOINT_NAMES[i]: c.default_joint_pos[i] for i in range(12)}, joint_vel={".*": 0.0}),
        soft_joint_pos_limit_factor=c.soft_limit_factor, actuators=actuators)
    scene = InteractiveSceneCfg(num_envs=num_envs, env_spacing=c.env_spacing)
    object.__setattr__(scene, "robot", robot)
    object.__setattr__(scene, "terrain", Placeholder(terrain_type="plane", physics_material=Placeholder(static_friction=1.0, dynamic_friction=1.0)))

    def obs(func, n=0.0, s=1.0):
        return ObservationTermCfg(func=func, noise=Unoise(n_min=-n, n_max=n) if n else None, scale=None if s == 1.0 else s)

    policy = ObservationGroupCfg(concatenate_terms=True, enable_corruption=bool(c.enable_corruption), history_length=c.history_length)
    for k, v in (("base_ang_vel", obs(mdp.base_ang_vel, c.noise_ang_vel, c.scale_ang_vel)),
                 ("projected_gravity", obs(mdp.projected_gravity, c.noise_gravity, c.scale_gravity)),
                 ("velocity_commands", obs(mdp.generated_commands, 0.0, c.scale_cmd)),
                 ("joint_pos", obs(mdp.joint_pos_rel, c.noise_joint_pos, c.scale_joint_pos)),
                 ("joint_vel", obs(mdp.joint_vel_rel, c.noise_joint_vel, c.scale_joint_vel)),
                 ("actions", obs(mdp.last_action, 0.0, c.scale_action))):
        object.__setattr__(policy, k, v)
    policy.__configclass_fields__ = lambda: ["concatenate_terms", "enable_corruption", "history_length", "base_ang_vel", "projected_gravity",
                                             "velocity_commands", "joint_pos", "joint_vel", "actions"]

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
        "contact_forces": {"sensor_cfg": SceneEntityCfg("contact_forces", body_names=names(c.mask_undesired_slots, SLOT_BODIES)), "threshold": 1.0},
        "base_height_l2": {"target_height": c.base_height_target},
    }
    slot_func = {s: f for f, s in REW_FUNC_SLOT.items()}
    from ._capi import REW_NAMES
    rew = {}
    for s, w in enumerate(c.rew_weight):
        if w != 0.0:
            f = slot_func[s]
            rew[REW_NAMES[s]] = RewardTermCfg(func=getattr(lmdp, f), weight=w, params=rew_params.get(f, {}))
    rewards = bag(**rew)

    terminations = bag(
        time_out=TerminationTermCfg(func=mdp.time_out, time_out=True),
        base_contact=TerminationTermCfg(func=mdp.illegal_contact, params={
            "sensor_cfg": SceneEntityCfg("contact_forces", body_names=names(c.mask_illegal_slots, SLOT_BODIES)), "threshold": c.contact_threshold}))
    ranges = ienvs.UniformVelocityCommandCfg.Ranges(lin_vel_x=tuple(c.cmd_lin_x), lin_vel_y=tuple(c.cmd_lin_y), ang_vel_z=tuple(c.cmd_ang_z),
                                                    heading=tuple(c.cmd_heading))
    commands = bag(base_velocity=ienvs.UniformVelocityCommandCfg(
        asset_name="robot", resampling_time_range=tuple(c.cmd_resample_time), rel_standing_envs=c.rel_standing_envs, rel_heading_envs=c.rel_heading_envs,
        heading_command=bool(c.heading_command), heading_control_stiffness=c.heading_stiffness, ranges=ranges))
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
            "mass_distribution_params": tuple(c.mass_add_range), "operation": "add"})
    if c.push_enable:
        ev["push_robot"] = EventTermCfg(func=mdp.push_by_setting_velocity, mode="interval", interval_range_s=tuple(c.push_interval_s),
                                        params={"velocity_range": {"x": tuple(c.push_vel_xy), "y": tuple(c.push_vel_xy)}})
    perm = list(c.joint_perm)
    action = ienvs.JointPositionActionCfg(asset_name="robot", joint_names=[JOINT_NAMES[j] for j in perm], scale=c.action_scale,
                                          use_default_offset=True, preserve_order=True)
    cfg = ienvs.ManagerBasedRLEnvCfg(decimation=c.decimation, episode_length_s=c.episode_length_s, scene=scene,
                                     observations=bag(policy=policy), actions=bag(joint_pos=action), rewards=rewards,
                                     terminations=terminations, commands=commands, events=bag(**ev), curriculum=None)
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


def register() -> bool:
    """gym.register the Flat id with this package's own entry points unless something registered it already."""
    from . import shims
    shims.install()
    import gymnasium as gym
    try:
        gym.spec(TASK_ID)
        return False
    except Exception:
        pass
    gym.register(id=TASK_ID, entry_point="h1v2_isaac_b200.env:H1v2ManagerBasedRLEnv", disable_env_checker=True,
                 kwargs={"env_cfg_entry_point": "h1v2_isaac_b200.tasks:default_env_cfg", "rsl_rl_cfg_entry_point": "h1v2_isaac_b200.tasks:default_agent_cfg"})
    return True
