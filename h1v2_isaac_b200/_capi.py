"""ctypes mirror of include/h1v2_b200.h and the loader of the in-tree CUDA library.

The library is the product: if it is missing or does not export a symbol the header declares, loading
fails loudly -- there is no CPU or PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

NJ = 12
NUM_REW = 22
NUM_SLOT = 6
OBS_TERM_DIM = 45
MAX_HISTORY = 10
LOG_DIM = 32
NUM_CSTR, CSTR_COLS = 10, 56
CSTR_NAMES = ["contact", "joint_position_limits", "joint_velocity_limits", "joint_torque_limits", "foot_contact_force", "no_move",
              "base_orientation", "base_height", "foot_contact", "foot_clearance"]
CSTR_COL0 = [0, 1, 13, 25, 37, 39, 51, 52, 53, 54, 56]  # first column of every term (and the end)

REW_NAMES = [
    "termination_penalty", "track_lin_vel_xy_exp", "track_ang_vel_z_exp", "feet_air_time", "feet_slide",
    "dof_pos_limits", "joint_deviation_hip", "ang_vel_xy_l2", "dof_torques_l2", "dof_acc_l2", "action_rate_l2",
    "flat_orientation_l2", "lin_vel_z_l2", "undesired_contacts", "track_lin_vel_xy_exp_base",
    "track_ang_vel_z_exp_base", "feet_air_time_l2", "joint_vel_l2", "base_height_l2", "contact_forces",
    "dof_pos_limits_b", "joint_deviation_b",
]
LOG_COUNT, LOG_REW0, LOG_TERM_TIMEOUT, LOG_TERM_CONTACT, LOG_ERR_XY, LOG_ERR_YAW = 0, 1, 23, 24, 25, 26
LOG_NAN_RESETS, LOG_MAX_ITERS, LOG_CAP_HITS, LOG_SUM_ITERS, LOG_CONTACT_OVERFLOW = 27, 28, 29, 30, 31

f32, i32, u32, i64 = C.c_float, C.c_int32, C.c_uint32, C.c_int64


class H1v2Config(C.Structure):
    _fields_ = [
        ("sim_dt", f32), ("decimation", i32), ("episode_length_s", f32),
        ("action_scale", f32), ("default_joint_pos", f32 * NJ), ("joint_perm", i32 * NJ),
        ("kp", f32 * NJ), ("kd", f32 * NJ), ("effort_limit", f32 * NJ), ("min_delay", i32), ("max_delay", i32),
        ("gravity", f32), ("friction", f32),
        ("dof_damping", f32 * 18), ("dof_armature", f32 * 18), ("dof_frictionloss", f32 * 18),
        ("act_frc_limit", f32 * NJ), ("joint_range", (f32 * 2) * NJ),
        ("contact_solref", f32 * 2), ("contact_solimp", f32 * 5),
        ("floss_solref", f32 * 2), ("floss_solimp", f32 * 5),
        ("limit_solref", f32 * 2), ("limit_solimp", f32 * 5),
        ("solver_iterations", i32), ("solver_tolerance", f32), ("solver_step_tolerance", f32), ("solver_ls_tolerance", f32),
        ("history_length", i32), ("enable_corruption", i32),
        ("noise_ang_vel", f32), ("noise_gravity", f32), ("noise_joint_pos", f32), ("noise_joint_vel", f32),
        ("scale_ang_vel", f32), ("scale_gravity", f32), ("scale_cmd", f32), ("scale_joint_pos", f32),
        ("scale_joint_vel", f32), ("scale_action", f32),
        ("rew_weight", f32 * NUM_REW), ("track_std", f32), ("feet_air_threshold", f32),
        ("contact_threshold", f32), ("soft_limit_factor", f32), ("base_height_target", f32),
        ("mask_pos_limits", u32), ("mask_joint_dev", u32), ("mask_torques", u32), ("mask_undesired_slots", u32),
        ("mask_illegal_slots", u32),
        ("cmd_lin_x", f32 * 2), ("cmd_lin_y", f32 * 2), ("cmd_ang_z", f32 * 2), ("cmd_heading", f32 * 2),
        ("cmd_resample_time", f32 * 2),
        ("rel_standing_envs", f32), ("rel_heading_envs", f32), ("heading_stiffness", f32), ("heading_command", i32),
        ("reset_pose_range", (f32 * 2) * 6), ("reset_vel_range", (f32 * 2) * 6),
        ("reset_joint_pos_scale", f32 * 2), ("reset_joint_vel_scale", f32 * 2), ("init_root_height", f32),
        ("push_enable", i32), ("push_interval_s", f32 * 2), ("push_vel_xy", f32 * 2),
        ("mass_add_range", f32 * 2), ("friction_range", f32 * 2),
        ("env_id_offset", i64), ("env_spacing", f32), ("joint_vel_limit", f32),
        ("mask_pos_limits_b", u32), ("mask_joint_dev_b", u32), ("mask_contact_forces_slots", u32), ("contact_forces_threshold", f32),
        ("command_class", i32), ("velocity_deadzone", f32), ("ang_vel_flip_prob", f32),
        ("root_link_com", f32 * 3), ("body_vel_at_com", i32),
        ("mass_recompute_inertia", i32), ("cat_enable", i32), ("cat_tau", f32), ("cat_min_p", f32), ("cat_max_p", f32 * 10), ("cat_contact_slots", u32),
        ("cat_foot_force_limit", f32), ("cat_no_move_deadzone", f32), ("cat_no_move_vel_limit", f32), ("cat_orientation_limit", f32),
        ("cat_height", f32), ("cat_height_std", f32), ("cat_clearance_min_height", f32), ("cat_clearance_deadzone", f32),
        ("runaway_vel", f32), ("solver_vel_tolerance", f32),
        ("terrain_enable", i32), ("terrain_rows", i32), ("terrain_cols", i32), ("terrain_tile_size", f32), ("terrain_hscale", f32), ("terrain_vscale", f32),
        ("terrain_level_min", i32), ("terrain_level_max", i32), ("terrain_level_step", i32), ("terrain_border_px", i32),
        ("terrain_max_init_level", i32), ("terrain_curriculum", i32),
        ("obs_base_lin_vel", i32), ("noise_lin_vel", f32), ("scale_lin_vel", f32),
        ("obs_height_scan", i32), ("scan_size", f32 * 2), ("scan_resolution", f32), ("scan_offset", f32),
        ("noise_height_scan", f32), ("scale_height_scan", f32), ("scan_clip", f32 * 2),
        ("reserved", i32 * 7),
    ]

    def copy(self) -> "H1v2Config":
        out = H1v2Config()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(H1v2Config))
        return out


# (name, per-env element count as a function of history H, ctype)
STATE_FIELDS = [
    ("root_pos", 3, f32), ("root_quat", 4, f32), ("root_lin_vel", 3, f32), ("root_ang_vel", 3, f32),
    ("joint_pos", NJ, f32), ("joint_vel", NJ, f32), ("last_action", NJ, f32), ("target_hist", 2 * NJ, f32),
    ("lag", 1, i32), ("fresh", 1, i32), ("command", 3, f32), ("heading_target", 1, f32), ("time_left", 1, f32),
    ("is_standing", 1, i32), ("is_heading", 1, i32), ("cmd_metrics", 2, f32), ("feet_timers", 8, f32),
    ("episode_sums", NUM_REW, f32), ("obs_history", None, f32), ("friction", 1, f32), ("mass_add", 1, f32),
    ("push_time_left", 1, f32),
    ("slot_force", NUM_SLOT * 3, f32), ("slot_force_hist", NUM_SLOT * 3, f32), ("applied_torque", NJ, f32),
    ("joint_acc", NJ, f32), ("reward_terms", NUM_REW, f32), ("foot_vel", 6, f32), ("solver_iters", 3, f32),
    ("pre_reset_qpos", 19, f32), ("pre_reset_qvel", 18, f32), ("pre_reset_timers", 8, f32),
    ("terrain_level", 1, i32), ("terrain_type", 1, i32),
]
READ_ONLY_STATE = {"terrain_type", "slot_force", "slot_force_hist", "applied_torque", "joint_acc", "reward_terms", "foot_vel", "solver_iters", "pre_reset_qpos", "pre_reset_qvel", "pre_reset_timers"}


class H1v2State(C.Structure):
    _fields_ = [(name, C.POINTER(ct)) for name, _, ct in STATE_FIELDS]


def state_field_count(name: str, history: int) -> int:
    for n, cnt, _ in STATE_FIELDS:
        if n == name:
            return history * OBS_TERM_DIM if cnt is None else cnt
    raise KeyError(name)


_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libh1v2_b200.so")

_SYMBOLS = {
    "h1v2_default_config": (C.c_int, [C.POINTER(H1v2Config)]),
    "h1v2_rsl_config": (C.c_int, [C.POINTER(H1v2Config)]),
    "h1v2_cat_config": (C.c_int, [C.POINTER(H1v2Config)]),
    "h1v2_rough_config": (C.c_int, [C.POINTER(H1v2Config)]),
    "h1v2_terrain_dims": (C.c_int, [C.c_void_p, C.POINTER(i32)]),
    "h1v2_get_terrain": (C.c_int, [C.c_void_p, C.c_void_p]),
    "h1v2_set_terrain": (C.c_int, [C.c_void_p, C.c_void_p]),
    "h1v2_get_terrain_log": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "h1v2_create": (C.c_int, [C.POINTER(H1v2Config), i32, i32, C.c_uint64, C.POINTER(C.c_void_p)]),
    "h1v2_destroy": (None, [C.c_void_p]),
    "h1v2_last_error": (C.c_char_p, []),
    "h1v2_check_guards": (i64, [C.c_void_p]),
    "h1v2_obs_dim": (C.c_int, [C.c_void_p]),
    "h1v2_num_envs": (C.c_int, [C.c_void_p]),
    "h1v2_bind_episode_length": (C.c_int, [C.c_void_p, C.c_void_p]),
    "h1v2_reset": (C.c_int, [C.c_void_p, C.c_void_p, i32, C.c_void_p]),
    "h1v2_observe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "h1v2_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "h1v2_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "h1v2_host_path_info": (C.c_int, [C.c_void_p, C.POINTER(i32), C.POINTER(i32)]),
    "h1v2_host_path_rows": (C.c_int, [C.c_void_p]),
    "h1v2_set_reward_weights": (C.c_int, [C.c_void_p, C.POINTER(f32)]),
    "h1v2_cat_step": (C.c_int, [C.c_void_p] * 7),
    "h1v2_cat_step_host": (C.c_int, [C.c_void_p] * 6),
    "h1v2_set_constraint_max_p": (C.c_int, [C.c_void_p, C.POINTER(f32)]),
    "h1v2_cat_debug": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "h1v2_get_cat_log_host": (C.c_int, [C.c_void_p, C.POINTER(f32)]),
    "h1v2_get_cat_log": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "h1v2_get_state": (C.c_int, [C.c_void_p, C.POINTER(H1v2State), C.c_void_p]),
    "h1v2_set_state": (C.c_int, [C.c_void_p, C.POINTER(H1v2State), C.c_void_p]),
    "h1v2_get_log": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "h1v2_get_log_host": (C.c_int, [C.c_void_p, C.POINTER(f32)]),
    "h1v2_envs_per_warp": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "h1v2_debug_iter_hist": (C.c_int, [C.c_void_p, C.POINTER(f32)]),
    "h1v2_launch_count": (i64, [C.c_void_p]),
    "h1v2_measure_fp32_peak": (C.c_int, [i32, C.POINTER(f32)]),
    "h1v2_random_actions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load libh1v2_b200.so and bind every symbol of include/h1v2_b200.h; raises if anything is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("H1V2_LIB") or LIB_PATH  # H1V2_LIB: build variants under test (tools/diag_variants.py)
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: the CUDA extension is the product and there is no fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a)."
        )
    lib = C.CDLL(p)
    for name, (res, args) in _SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def default_config() -> H1v2Config:
    cfg = H1v2Config()
    rc = load_library().h1v2_default_config(C.byref(cfg))
    if rc != 0:
        raise RuntimeError("h1v2_default_config failed")
    return cfg


def rsl_config() -> H1v2Config:
    """Resolved cfg of Isaac-Velocity-Rsl-H12_12dof-v0 (config/h12_12dof/rsl_env_cfg.py)."""
    cfg = H1v2Config()
    rc = load_library().h1v2_rsl_config(C.byref(cfg))
    if rc != 0:
        raise RuntimeError("h1v2_rsl_config failed")
    return cfg


def rough_config() -> H1v2Config:
    """Resolved cfg of Isaac-Velocity-Rough-H12_12dof-v0 (config/h12_12dof/rough_env_cfg.py, utils/mdp/terrains.py)."""
    cfg = H1v2Config()
    rc = load_library().h1v2_rough_config(C.byref(cfg))
    if rc != 0:
        raise RuntimeError("h1v2_rough_config failed")
    return cfg


def scan_count(size: float, resolution: float) -> int:
    """Rays along one axis of a GridPatternCfg: len(torch.arange(-size / 2, size / 2 + 1e-9, resolution))."""
    import math
    return int(math.floor(size / resolution + 1e-4)) + 1  # 1e-4: the fp32 rounding of the config values (0.1f > 0.1)


def obs_dim_of(cfg: H1v2Config) -> int:
    if cfg.obs_base_lin_vel or cfg.obs_height_scan:
        scan = scan_count(cfg.scan_size[0], cfg.scan_resolution) * scan_count(cfg.scan_size[1], cfg.scan_resolution) if cfg.obs_height_scan else 0
        return (3 if cfg.obs_base_lin_vel else 0) + OBS_TERM_DIM + scan
    return cfg.history_length * OBS_TERM_DIM


def cat_config() -> H1v2Config:
    """Resolved cfg of Isaac-Velocity-CaT-Flat-H12_12dof-v0 (config/h12_12dof/cat_env_cfg.py)."""
    cfg = H1v2Config()
    rc = load_library().h1v2_cat_config(C.byref(cfg))
    if rc != 0:
        raise RuntimeError("h1v2_cat_config failed")
    return cfg
