"""isaaclab.envs stand-in: cfg classes + the env class itself (the B200 backend) + the mdp term namespace."""
from __future__ import annotations

import sys

from h1v2_isaac_b200.shims._configclass import MISSING, configclass
from h1v2_isaac_b200.shims._lenient import Placeholder, make_lenient
from isaaclab import SimulationCfg  # noqa: F401

from h1v2_isaac_b200.env import H1v2ManagerBasedRLEnv


@configclass
class ViewerCfg:
    eye: tuple = (7.5, 7.5, 7.5)
    lookat: tuple = (0.0, 0.0, 0.0)
    cam_prim_path: str = "/OmniverseKit_Persp"
    resolution: tuple = (1280, 720)
    origin_type: str = "world"
    env_index: int = 0
    asset_name = None
    body_name = None


@configclass
class ManagerBasedEnvCfg:
    viewer = ViewerCfg()
    sim = SimulationCfg()
    ui_window_class_type = None
    seed = None
    decimation: int = MISSING
    scene = MISSING
    recorders = Placeholder()
    observations = MISSING
    actions = MISSING
    events = Placeholder()
    rerender_on_reset: bool = False
    wait_for_textures: bool = True


@configclass
class ManagerBasedRLEnvCfg(ManagerBasedEnvCfg):
    is_finite_horizon: bool = False
    episode_length_s: float = MISSING
    rewards = MISSING
    terminations = MISSING
    curriculum = None
    commands = None


@configclass
class DirectRLEnvCfg:
    pass


@configclass
class DirectMARLEnvCfg:
    pass


class DirectRLEnv:
    pass


class DirectMARLEnv:
    pass


def multi_agent_to_single_agent(env, *a, **k):
    return env


ManagerBasedRLEnv = H1v2ManagerBasedRLEnv
ManagerBasedEnv = H1v2ManagerBasedRLEnv
RLTaskEnv = H1v2ManagerBasedRLEnv
VecEnvStepReturn = tuple
VecEnvObs = dict

_envs = sys.modules[__name__]
make_lenient("isaaclab.envs.common", VecEnvStepReturn=tuple, VecEnvObs=dict)
make_lenient("isaaclab.envs.manager_based_rl_env", ManagerBasedRLEnv=H1v2ManagerBasedRLEnv)
make_lenient("isaaclab.envs.manager_based_env", ManagerBasedEnv=H1v2ManagerBasedRLEnv)


# ---------------------------------------------------------------- mdp: names only -- the arithmetic lives in the fused kernel
def _term(name):
    def f(env, *args, **kwargs):
        raise NotImplementedError(f"mdp.{name}: evaluated inside the fused B200 step kernel, not callable on the host")
    f.__name__ = f.__qualname__ = name
    return f


_TERM_NAMES = [
    # observations
    "base_pos_z", "base_lin_vel", "base_ang_vel", "projected_gravity", "root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w",
    "joint_pos", "joint_pos_rel", "joint_pos_limit_normalized", "joint_vel", "joint_vel_rel", "last_action", "generated_commands",
    "height_scan", "body_incoming_wrench", "imu_orientation", "imu_ang_vel", "imu_lin_acc",
    # rewards
    "is_alive", "is_terminated", "lin_vel_z_l2", "ang_vel_xy_l2", "flat_orientation_l2", "base_height_l2", "body_lin_acc_l2",
    "joint_torques_l2", "joint_vel_l1", "joint_vel_l2", "joint_acc_l2", "joint_deviation_l1", "joint_pos_limits", "joint_vel_limits",
    "applied_torque_limits", "action_rate_l2", "action_l2", "undesired_contacts", "contact_forces", "track_lin_vel_xy_exp",
    "track_ang_vel_z_exp",
    # terminations
    "time_out", "command_resample", "bad_orientation", "root_height_below_minimum", "joint_pos_out_of_limit",
    "joint_pos_out_of_manual_limit", "joint_vel_out_of_limit", "joint_vel_out_of_manual_limit", "joint_effort_out_of_limit", "illegal_contact",
    # events
    "randomize_rigid_body_scale", "randomize_rigid_body_material", "randomize_rigid_body_mass", "randomize_rigid_body_com",
    "randomize_rigid_body_collider_offsets", "randomize_physics_scene_gravity", "randomize_actuator_gains", "randomize_joint_parameters",
    "randomize_fixed_tendon_parameters", "apply_external_force_torque", "push_by_setting_velocity", "reset_root_state_uniform",
    "reset_root_state_with_random_orientation", "reset_root_state_from_terrain", "reset_joints_by_scale", "reset_joints_by_offset",
    "reset_nodal_state_uniform", "reset_scene_to_default",
    # curriculums
    "modify_reward_weight", "modify_env_param", "modify_term_cfg",
]


@configclass
class JointActionCfg:
    class_type = None
    asset_name: str = MISSING
    debug_vis: bool = False
    clip = None
    joint_names: list = MISSING
    scale = 1.0
    offset = 0.0
    preserve_order: bool = False


@configclass
class JointPositionActionCfg(JointActionCfg):
    use_default_offset: bool = True


@configclass
class JointEffortActionCfg(JointActionCfg):
    pass


@configclass
class UniformVelocityCommandCfg:
    @configclass
    class Ranges:
        lin_vel_x: tuple = MISSING
        lin_vel_y: tuple = MISSING
        ang_vel_z: tuple = MISSING
        heading = None

    class_type = None
    resampling_time_range = MISSING
    debug_vis: bool = False
    asset_name: str = MISSING
    heading_command: bool = False
    heading_control_stiffness: float = 1.0
    rel_standing_envs: float = 0.0
    rel_heading_envs: float = 1.0
    ranges = MISSING
    goal_vel_visualizer_cfg = Placeholder()
    current_vel_visualizer_cfg = Placeholder()


class UniformVelocityCommand:
    """Placeholder base so that reference subclasses (utils/mdp/commands.py:19) import; command logic is in the kernel."""

    cfg = None

    def __init__(self, cfg=None, env=None):
        self.cfg = cfg


UniformVelocityCommandCfg.class_type = UniformVelocityCommand

mdp = make_lenient("isaaclab.envs.mdp", JointPositionActionCfg=JointPositionActionCfg, JointEffortActionCfg=JointEffortActionCfg,
                   JointActionCfg=JointActionCfg, UniformVelocityCommandCfg=UniformVelocityCommandCfg, UniformVelocityCommand=UniformVelocityCommand)
for _n in _TERM_NAMES:
    setattr(mdp, _n, _term(_n))
mdp.__all__ = _TERM_NAMES + ["JointPositionActionCfg", "JointEffortActionCfg", "JointActionCfg", "UniformVelocityCommandCfg", "UniformVelocityCommand"]
_envs.mdp = mdp
make_lenient("isaaclab.envs.mdp.actions")
make_lenient("isaaclab.envs.mdp.commands")
