"""H1v2ManagerBasedRLEnv: the ManagerBasedRLEnv contract of the reference, backed by the fused B200 step kernel.

Drop-in seam: `gym.make("Isaac-Velocity-Flat-H12_12dof-v0", cfg=env_cfg)` in the reference's
scripts/rsl_rl/train.py:102 constructs this class (through the shimmed `isaaclab.envs:ManagerBasedRLEnv` entry point
of packages/biped_tasks/.../config/h12_12dof/__init__.py:40-49, or through a real isaaclab install that points the
entry point here).  The step order is packages/biped_tasks/biped_tasks/utils/cat/cat_env.py:95-193; all of it runs
inside ONE kernel launch (h1v2_step), nothing on the host except this thin wrapper.

`flatten_cfg` turns the reference's resolved @configclass tree into the POD `H1v2Config` the C-ABI consumes.  It reads
term lists, weights, scales, noise, ranges, gains and limits from the tree -- nothing about the task is hard-coded here;
what the tree does not hold (the MJCF rigid-body model and MuJoCo solver parameters) comes from h1v2_default_config.
Anything in the tree the kernel cannot express raises NotImplementedError naming the offending entry.
"""
from __future__ import annotations

import math
import re
from typing import Any

from ._capi import CSTR_NAMES, H1v2Config, LOG_COUNT, LOG_ERR_XY, LOG_ERR_YAW, LOG_REW0, LOG_TERM_CONTACT, LOG_TERM_TIMEOUT, NJ, REW_NAMES, default_config

# MJCF (leg-major) joint order of h12_12dof.xml:71-132 == A/robots/h12.py:40-53
JOINT_NAMES = [f"{s}_{j}_joint" for s in ("left", "right") for j in ("hip_yaw", "hip_pitch", "hip_roll", "knee", "ankle_pitch", "ankle_roll")]
# PhysX breadth-first order used by joint_names=[".*"] without preserve_order
BREADTH_FIRST = [0, 6, 1, 7, 2, 8, 3, 9, 4, 10, 5, 11]
# contact-sensor slots of the kernel: bodies that own colliders in the PhysX asset (h12_12dof.urdf:116-121,168-191,387-392)
SLOT_BODIES = ["left_ankle_roll_link", "right_ankle_roll_link", "left_knee_link", "right_knee_link", "torso_link", "pelvis"]
# every rigid body of the articulation (names a body regex may legitimately match without owning a collider)
ALL_BODIES = SLOT_BODIES + [f"{s}_{j}_link" for s in ("left", "right") for j in (
    "hip_yaw", "hip_pitch", "hip_roll", "ankle_pitch", "shoulder_pitch", "shoulder_roll", "shoulder_yaw", "elbow", "wrist_roll", "wrist_pitch", "wrist_yaw")]

# reward function name -> kernel slot (include/h1v2_b200.h enum)
REW_FUNC_SLOT = {
    "is_terminated": 0, "track_lin_vel_xy_yaw_frame_exp": 1, "track_ang_vel_z_world_exp": 2, "feet_air_time_positive_biped": 3,
    "feet_slide": 4, "joint_pos_limits": 5, "joint_deviation_l1": 6, "ang_vel_xy_l2": 7, "joint_torques_l2": 8, "joint_acc_l2": 9,
    "action_rate_l2": 10, "flat_orientation_l2": 11, "lin_vel_z_l2": 12, "undesired_contacts": 13, "track_lin_vel_xy_exp": 14,
    "track_ang_vel_z_exp": 15, "feet_air_time": 16, "joint_vel_l2": 17, "base_height_l2": 18, "contact_forces": 19,
}
# functions a cfg may use twice (C12/rsl_env_cfg.py:343-386 has joint_deviation_l1 and joint_pos_limits on two joint sets each):
# the second term of the cfg goes to the kernel's second slot of that function
REW_FUNC_SLOT_B = {"joint_pos_limits": 20, "joint_deviation_l1": 21}
OBS_LAYOUT = ["base_ang_vel", "projected_gravity", "generated_commands", "joint_pos_rel", "joint_vel_rel", "last_action"]
# the Rough id's row (V/velocity_env_cfg.py:119-142): base_lin_vel first, height_scan last, no history
OBS_LAYOUT_ROUGH = ["base_lin_vel"] + OBS_LAYOUT + ["height_scan"]


def _get(obj, name, default=None):
    if obj is None:
        return default
    if isinstance(obj, dict):
        return obj.get(name, default)
    v = getattr(obj, name, default)
    return default if v is None else v


def _fname(func) -> str:
    return getattr(func, "__name__", str(func)).split(".")[-1]


def _match(patterns, names) -> list[int]:
    """isaaclab string_utils.resolve_matching_names semantics: full-match regexes, indices in `names` order."""
    if patterns is None:
        return list(range(len(names)))
    if isinstance(patterns, str):
        patterns = [patterns]
    return [i for i, n in enumerate(names) if any(re.fullmatch(p, n) for p in patterns)]


def _per_joint(value, default: float) -> list[float]:
    """float | {regex: float} -> 12 values in MJCF order."""
    out = [default] * NJ
    if value is None:
        return out
    if isinstance(value, dict):
        for pat, v in value.items():
            for i in _match(pat, JOINT_NAMES):
                out[i] = float(v)
        return out
    return [float(value)] * NJ


def _terms(group) -> list[tuple[str, Any]]:
    if group is None:
        return []
    names = group.__configclass_fields__() if hasattr(group, "__configclass_fields__") else [k for k in vars(group) if not k.startswith("_")]
    return [(n, getattr(group, n)) for n in names if getattr(group, n, None) is not None and hasattr(getattr(group, n), "func")]


def _joint_mask(params, key="asset_cfg") -> int:
    cfg = _get(params, key)
    ids = _match(_get(cfg, "joint_names"), JOINT_NAMES)
    return sum(1 << i for i in ids)


def _slot_mask(params, key="sensor_cfg") -> int:
    cfg = _get(params, key)
    pats = _get(cfg, "body_names")
    if pats in ("base", ["base"]):  # the upstream default names a body the H1-2 does not have; the H1-2 cfg overrides it
        return 0
    return sum(1 << s for s in _match(pats, SLOT_BODIES))


def reward_slots(cfg) -> dict[str, int]:
    """cfg reward term name -> kernel slot, in cfg order.  Zero-weight terms get a slot when one is free (a curriculum may raise
    them later); a non-zero-weight term without a slot raises."""
    used, out = set(), {}
    for n, t in _terms(cfg.rewards):
        f, w = _fname(t.func), float(t.weight)
        slot = REW_FUNC_SLOT.get(f)
        if slot in used:
            slot = REW_FUNC_SLOT_B.get(f)
        if slot is None or slot in used:
            if w == 0.0:
                continue  # RewardManager skips zero-weight terms
            if f not in REW_FUNC_SLOT:
                raise NotImplementedError(f"rewards.{n}: mdp.{f} is not implemented in the fused kernel")
            raise NotImplementedError(f"rewards.{n}: mdp.{f} appears more often than the kernel has slots for it")
        used.add(slot)
        out[n] = slot
    return out


def constraint_terms(cfg) -> dict[str, int]:
    """cfg constraint term name -> kernel constraint index (utils/cat/constraints.py function name -> _capi.CSTR_NAMES)."""
    out = {}
    for n, t in _terms(_get(cfg, "constraints")):
        f = _fname(t.func)
        if f not in CSTR_NAMES:
            raise NotImplementedError(f"constraints.{n}: constraints.{f} is not implemented in the CaT tail")
        if CSTR_NAMES.index(f) in out.values():
            raise NotImplementedError(f"constraints.{n}: constraints.{f} appears twice")
        out[n] = CSTR_NAMES.index(f)
    return out


def constraint_curriculum(cfg) -> list[tuple[int, int, float]]:
    """curriculums.modify_constraint_p terms (utils/cat/curriculums.py:20-42) -> [(constraint index, num_steps, init_max_p)]."""
    out, terms = [], constraint_terms(cfg)
    for n, t in _terms(_get(cfg, "curriculum")):
        if _fname(t.func) == "modify_constraint_p":
            p = t.params or {}
            if p["term_name"] not in terms:
                raise NotImplementedError(f"curriculum.{n}: constraint term {p['term_name']!r} does not exist")
            out.append((terms[p["term_name"]], int(p["num_steps"]), float(p["init_max_p"])))
    return out


def constraint_max_p(schedule, base, common_step_counter: int) -> list[float]:
    """max_p of every constraint at a step count: utils/cat/curriculums.py:27-34 for the scheduled terms, the cfg value otherwise."""
    p = list(base)
    for idx, num_steps, init in schedule:
        progress = min(common_step_counter / num_steps, 1.0)
        p[idx] = 1.0 / (20 + progress * (1.0 / init - 20))
    return p


def curriculum_schedule(cfg) -> list[tuple[int, float, int]]:
    """CurriculumManager terms -> [(reward slot, weight, num_steps)].  mdp.modify_reward_weight (C12/rsl_env_cfg.py:447-497): once
    common_step_counter > num_steps the term's weight becomes `weight`; curriculums.modify_constraint_p is handled by
    constraint_curriculum; anything else is refused."""
    out, slots = [], reward_slots(cfg)
    for n, t in _terms(_get(cfg, "curriculum")):
        f, p = _fname(t.func), (t.params or {})
        if f in ("modify_constraint_p", "terrain_levels_vel"):  # the terrain-level curriculum runs inside the kernel's reset (flatten_cfg)
            continue
        if f != "modify_reward_weight":
            raise NotImplementedError(f"curriculum.{n}: mdp.{f} is not implemented (only modify_reward_weight, modify_constraint_p and terrain_levels_vel are)")
        if p["term_name"] not in slots:
            raise NotImplementedError(f"curriculum.{n}: reward term {p['term_name']!r} has no kernel slot")
        out.append((slots[p["term_name"]], float(p["weight"]), int(p["num_steps"])))
    return out


def flatten_cfg(cfg) -> H1v2Config:
    """Resolved ManagerBasedRLEnvCfg tree (reference: config/h12_12dof/flat_env_cfg.py:13-48 and parents, or
    config/h12_12dof/rsl_env_cfg.py:44-540) -> H1v2Config."""
    c = default_config()  # rigid-body model, MuJoCo solver parameters; everything below is overwritten from the tree
    # managers the fused kernel does not have: refuse rather than silently drop them
    # ---- Constraints-as-Terminations manager (utils/cat/*; C12/cat_env_cfg.py:336-431) ----
    cterms = constraint_terms(cfg)
    constraint_curriculum(cfg)
    if cterms:
        c.cat_enable = 1
        for t in range(len(CSTR_NAMES)):
            c.cat_max_p[t] = 0.0  # a term the cfg does not hold never terminates
        all_joints = list(range(NJ))
        for n, idx in cterms.items():
            t = getattr(cfg.constraints, n)
            p, f = (t.params or {}), CSTR_NAMES[idx]
            c.cat_max_p[idx] = float(t.max_p)
            if f in ("joint_position_limits", "joint_velocity_limits", "joint_torque_limits", "no_move") and sorted(_match(_get(_get(p, "asset_cfg"), "joint_names"), JOINT_NAMES)) != all_joints:
                raise NotImplementedError(f"constraints.{n}: must cover all 12 leg joints")
            if f in ("foot_contact_force", "foot_contact") and _slot_mask(p, "asset_cfg") != 0b11:
                raise NotImplementedError(f"constraints.{n}: bodies must be the two ankle_roll links")
            if f == "contact":
                c.cat_contact_slots = _slot_mask(p, "asset_cfg")
            elif f == "foot_contact_force":
                c.cat_foot_force_limit = float(p["limit"])
            elif f == "no_move":
                c.cat_no_move_deadzone, c.cat_no_move_vel_limit = float(p["velocity_deadzone"]), float(p["joint_vel_limit"])
            elif f == "base_orientation":
                c.cat_orientation_limit = float(p["limit"])
            elif f == "base_height":
                c.cat_height, c.cat_height_std = float(p["height"]), float(p["std"])
            elif f == "foot_clearance":
                c.cat_clearance_min_height, c.cat_clearance_deadzone = float(p["min_height"]), float(p["velocity_deadzone"])
                if _slot_mask(p, "contact_asset_cfg") != 0b11:
                    raise NotImplementedError(f"constraints.{n}: bodies must be the two ankle_roll links")
    curriculum_schedule(cfg)  # raises on anything but modify_reward_weight; the schedule itself is applied by the env, host side
    c.sim_dt = float(cfg.sim.dt)
    c.decimation = int(cfg.decimation)
    c.episode_length_s = float(cfg.episode_length_s)
    g = _get(cfg.sim, "gravity", (0.0, 0.0, -9.81))
    c.gravity = float(-g[2])
    c.env_spacing = float(_get(cfg.scene, "env_spacing", 2.5))

    # ---- robot articulation (A/robots/h12.py:18-114) ----
    robot = cfg.scene.robot
    c.init_root_height = float(robot.init_state.pos[2])
    q0 = _per_joint(robot.init_state.joint_pos, 0.0)
    c.soft_limit_factor = float(_get(robot, "soft_joint_pos_limit_factor", 1.0))
    kp, kd, ef = [None] * NJ, [None] * NJ, [None] * NJ
    delays, vlims = set(), set()
    for name, act in robot.actuators.items():
        ids = _match(act.joint_names_expr, JOINT_NAMES)
        st, dm = _per_joint(act.stiffness, float("nan")), _per_joint(act.damping, float("nan"))
        el = _per_joint(_get(act, "effort_limit"), float("inf"))
        arm = _get(act, "armature")
        for i in ids:
            kp[i], kd[i], ef[i] = st[i], dm[i], el[i]
            if arm is not None:
                c.dof_armature[6 + i] = _per_joint(arm, 0.0)[i]
        delays.add((int(_get(act, "min_delay", 0)), int(_get(act, "max_delay", 0))))
        vl = _get(act, "velocity_limit")
        vlims.add(None if vl is None else float(vl))
    if any(v is None or (isinstance(v, float) and math.isnan(v)) for v in kp + kd):
        raise NotImplementedError("actuators: every one of the 12 leg joints needs stiffness and damping")
    if len(delays) != 1:
        raise NotImplementedError(f"actuators: all groups must share one (min_delay, max_delay); got {sorted(delays)}")
    c.min_delay, c.max_delay = delays.pop()
    if len(vlims) != 1:
        raise NotImplementedError(f"actuators: all groups must share one velocity_limit; got {sorted(map(str, vlims))}")
    vl = vlims.pop()
    c.joint_vel_limit = 0.0 if vl is None else vl
    for i in range(NJ):
        c.default_joint_pos[i], c.kp[i], c.kd[i] = q0[i], kp[i], kd[i]
        c.effort_limit[i] = ef[i] if math.isfinite(ef[i]) else 1e9

    # ---- actions (V/velocity_env_cfg.py:111) ----
    acts = [(n, getattr(cfg.actions, n)) for n in cfg.actions.__configclass_fields__() if getattr(cfg.actions, n) is not None]
    if len(acts) != 1 or not hasattr(acts[0][1], "joint_names"):
        raise NotImplementedError("actions: exactly one JointPositionAction term is supported")
    a = acts[0][1]
    if not _get(a, "use_default_offset", True):
        raise NotImplementedError("actions: use_default_offset=False is not supported")
    if isinstance(a.scale, dict):
        raise NotImplementedError("actions: per-joint scale dict is not supported")
    c.action_scale = float(a.scale)
    if _get(a, "preserve_order", False):
        order = []
        for pat in a.joint_names:
            order += [i for i in _match(pat, JOINT_NAMES) if i not in order]
    else:
        sel = set(_match(a.joint_names, JOINT_NAMES))
        order = [i for i in BREADTH_FIRST if i in sel]
    if sorted(order) != list(range(NJ)):
        raise NotImplementedError("actions: joint_names must select all 12 leg joints")
    for i in range(NJ):
        c.joint_perm[i] = order[i]

    # ---- observations (V/velocity_env_cfg.py:123-142; flat_env_cfg.py:22-27) ----
    pol = cfg.observations.policy
    terms = _terms(pol)
    got = [_fname(t.func) for _, t in terms]
    rough_obs = got != OBS_LAYOUT
    if rough_obs:
        lead, tail = got[:1] == ["base_lin_vel"], got[-1:] == ["height_scan"]
        if got[int(lead):len(got) - int(tail)] != OBS_LAYOUT:
            raise NotImplementedError(f"observations.policy: the fused kernel emits {OBS_LAYOUT}, optionally led by base_lin_vel and ended by height_scan; the cfg asks for {got}")
        c.obs_base_lin_vel, c.obs_height_scan = int(lead), int(tail)
    if not _get(pol, "concatenate_terms", True):
        raise NotImplementedError("observations.policy.concatenate_terms=False is not supported")
    c.history_length = int(_get(pol, "history_length", 0) or 1)
    if rough_obs and c.history_length != 1:
        raise NotImplementedError("observations.policy: base_lin_vel / height_scan cannot be combined with an observation history")
    if int(_get(pol, "history_step", 1) or 1) != 1:  # T/utils/history/observation_manager.py:441-451 (the CaT cfg sets 1)
        raise NotImplementedError("observations.policy.history_step != 1 is not supported")
    if _get(pol, "flatten_history_dim", True) is False:
        raise NotImplementedError("observations.policy.flatten_history_dim=False is not supported")
    for n, t in terms:
        th = _get(t, "history_length")
        if th not in (None, 0, c.history_length):
            raise NotImplementedError(f"observations.policy.{n}.history_length differs from the group's")
    c.enable_corruption = int(bool(_get(pol, "enable_corruption", False)))
    noise, scale = {}, {}
    for n, t in terms:
        f = _fname(t.func)
        nz = _get(t, "noise")
        if nz is None:
            noise[f] = 0.0
        else:
            lo, hi = float(nz.n_min), float(nz.n_max)
            if abs(lo + hi) > 1e-12:
                raise NotImplementedError(f"observations.policy.{n}: only symmetric additive uniform noise is supported")
            noise[f] = hi
        clip = _get(t, "clip")
        if f == "height_scan":
            c.scan_clip[0], c.scan_clip[1] = (float(clip[0]), float(clip[1])) if clip is not None else (-3.0e38, 3.0e38)
        elif clip is not None:
            raise NotImplementedError(f"observations.policy.{n}.clip is not supported")
        sc = _get(t, "scale")
        scale[f] = 1.0 if sc is None else float(sc)
    for n, t in terms:  # joint-indexed terms must list the joints in the action's order (one permutation in the kernel)
        ac = _get(_get(t, "params"), "asset_cfg")
        names = _get(ac, "joint_names")
        if names is None:
            jo = BREADTH_FIRST
        elif _get(ac, "preserve_order", False):
            jo = []
            for pat in ([names] if isinstance(names, str) else names):
                jo += [i for i in _match(pat, JOINT_NAMES) if i not in jo]
        else:
            sel = set(_match(names, JOINT_NAMES))
            jo = [i for i in BREADTH_FIRST if i in sel]
        if _fname(t.func) in ("joint_pos_rel", "joint_vel_rel") and list(jo) != list(order):
            raise NotImplementedError(f"observations.policy.{n}: joint order differs from the action term's")
    if noise["generated_commands"] or noise["last_action"]:
        raise NotImplementedError("observations.policy: noise on commands / last_action is not supported")
    c.noise_ang_vel, c.noise_gravity, c.noise_joint_pos, c.noise_joint_vel = noise["base_ang_vel"], noise["projected_gravity"], noise["joint_pos_rel"], noise["joint_vel_rel"]
    c.scale_ang_vel, c.scale_gravity, c.scale_cmd, c.scale_joint_pos, c.scale_joint_vel, c.scale_action = (scale[f] for f in OBS_LAYOUT)
    if c.obs_base_lin_vel:
        c.noise_lin_vel, c.scale_lin_vel = noise["base_lin_vel"], scale["base_lin_vel"]
    if c.obs_height_scan:  # V/velocity_env_cfg.py:61-68,133-138: RayCasterCfg on the torso (= pelvis frame), GridPattern, yaw-aligned
        hs = dict(terms)[[n for n, t in terms if _fname(t.func) == "height_scan"][0]]
        sensor = _get(cfg.scene, _get(_get(hs.params, "sensor_cfg"), "name", "height_scanner"))
        if sensor is None:
            raise NotImplementedError("observations.policy.height_scan: the scene has no such ray caster")
        body = str(_get(sensor, "prim_path", "")).rsplit("/", 1)[-1]
        if body not in ("torso_link", "pelvis", "base"):  # torso_link is welded to the pelvis at zero offset (h12_12dof.urdf:394-400)
            raise NotImplementedError(f"scene.height_scanner: attached to {body!r}; only the pelvis / torso_link frame is supported")
        if not _get(sensor, "attach_yaw_only", False):
            raise NotImplementedError("scene.height_scanner: attach_yaw_only=False is not supported")
        off = _get(_get(sensor, "offset"), "pos", (0.0, 0.0, 0.0))
        if abs(float(off[0])) > 1e-9 or abs(float(off[1])) > 1e-9:
            raise NotImplementedError("scene.height_scanner: a horizontal sensor offset is not supported")
        pat = _get(sensor, "pattern_cfg")
        if type(pat).__name__ != "GridPatternCfg" or _get(pat, "ordering", "xy") != "xy" or tuple(_get(pat, "direction", (0.0, 0.0, -1.0))) != (0.0, 0.0, -1.0):
            raise NotImplementedError("scene.height_scanner: only a downward GridPatternCfg with 'xy' ordering is supported")
        c.scan_size[0], c.scan_size[1], c.scan_resolution = float(pat.size[0]), float(pat.size[1]), float(pat.resolution)
        c.scan_offset = float(_get(hs.params, "offset", 0.5))  # mdp.height_scan(env, sensor_cfg, offset=0.5) [UPSTREAM default]
        c.noise_height_scan, c.scale_height_scan = noise["height_scan"], scale["height_scan"]

    # ---- terrain (V/velocity_env_cfg.py:40-58, utils/mdp/terrains.py:11-28) and its curriculum (V/mdp/curriculums.py:21-52) ----
    terrain = _get(cfg.scene, "terrain")
    ttype = _get(terrain, "terrain_type", "plane")
    if ttype == "generator":
        gen = _get(terrain, "terrain_generator")
        subs = dict(_get(gen, "sub_terrains") or {})
        if len(subs) != 1 or type(next(iter(subs.values()))).__name__ != "HfRandomUniformTerrainCfg":
            raise NotImplementedError("scene.terrain.terrain_generator: exactly one HfRandomUniformTerrainCfg sub-terrain is supported "
                                      f"(utils/mdp/terrains.py:11-28); got {[type(v).__name__ for v in subs.values()]}")
        sub = next(iter(subs.values()))
        if abs(float(gen.size[0]) - float(gen.size[1])) > 1e-9:
            raise NotImplementedError("scene.terrain.terrain_generator.size: tiles must be square")
        if _get(sub, "downsampled_scale") not in (None, float(gen.horizontal_scale)):
            raise NotImplementedError("scene.terrain.terrain_generator: downsampled_scale is not supported")
        c.terrain_enable = 1
        c.terrain_rows, c.terrain_cols = int(gen.num_rows), int(gen.num_cols)
        c.terrain_tile_size = float(gen.size[0])
        hs_, vs_ = float(gen.horizontal_scale), float(gen.vertical_scale)  # python doubles, as upstream's int() conversions see them
        c.terrain_hscale, c.terrain_vscale = hs_, vs_
        # hf_terrains.random_uniform_terrain [UPSTREAM]: int(noise / vertical_scale) on both ends and the step
        c.terrain_level_min, c.terrain_level_max = int(float(sub.noise_range[0]) / vs_), int(float(sub.noise_range[1]) / vs_)
        c.terrain_level_step = max(1, int(float(sub.noise_step) / vs_))
        bw = float(_get(sub, "border_width", 0.0))
        c.terrain_border_px = int(bw / hs_) + 1 if bw > 0 else 0  # height_field/utils.py height_field_to_mesh
        mil = _get(terrain, "max_init_terrain_level")
        c.terrain_max_init_level = -1 if mil is None else int(mil)
        cur_terms = [_fname(t.func) for _, t in _terms(_get(cfg, "curriculum"))]
        c.terrain_curriculum = int("terrain_levels_vel" in cur_terms and bool(_get(gen, "curriculum", False)))
    elif ttype != "plane":
        raise NotImplementedError(f"scene.terrain.terrain_type={ttype!r} is not supported (plane, or generator with a random-rough height field)")
    elif "terrain_levels_vel" in [_fname(t.func) for _, t in _terms(_get(cfg, "curriculum"))]:
        raise NotImplementedError("curriculum: terrain_levels_vel needs terrain_type='generator'")
    if c.terrain_enable and cterms:
        raise NotImplementedError("constraints: the Constraints-as-Terminations tail is not available on generated terrain")

    # ---- rewards (rough_env_cfg.py:18-62,112-120; flat_env_cfg.py:35-44; V/velocity_env_cfg.py:225-257) ----
    for i in range(len(REW_NAMES)):
        c.rew_weight[i] = 0.0
    std = None
    slots = reward_slots(cfg)
    for n, t in _terms(cfg.rewards):
        if n not in slots:
            continue
        w, f, slot, p = float(t.weight), _fname(t.func), slots[n], (t.params or {})
        c.rew_weight[slot] = w
        if "std" in p:
            if std is not None and abs(std - float(p["std"])) > 1e-9:
                raise NotImplementedError("rewards: tracking terms must share one std")
            std = float(p["std"])
        if f in ("feet_air_time", "feet_air_time_positive_biped"):
            c.feet_air_threshold = float(p["threshold"])
            if _slot_mask(p) != 0b11:
                raise NotImplementedError(f"rewards.{n}: sensor bodies must be the two ankle_roll links")
        elif f == "feet_slide" and _slot_mask(p) != 0b11:
            raise NotImplementedError(f"rewards.{n}: sensor bodies must be the two ankle_roll links")
        elif f == "joint_pos_limits":
            if slot == REW_FUNC_SLOT[f]:
                c.mask_pos_limits = _joint_mask(p)
            else:
                c.mask_pos_limits_b = _joint_mask(p)
        elif f == "joint_deviation_l1":
            if slot == REW_FUNC_SLOT[f]:
                c.mask_joint_dev = _joint_mask(p)
            else:
                c.mask_joint_dev_b = _joint_mask(p)
        elif f == "joint_torques_l2":
            c.mask_torques = _joint_mask(p)
        elif f == "undesired_contacts":
            c.mask_undesired_slots = _slot_mask(p)
            if abs(float(p.get("threshold", 1.0)) - 1.0) > 1e-9:
                raise NotImplementedError(f"rewards.{n}: threshold must equal the contact threshold 1.0")
        elif f == "contact_forces":
            c.mask_contact_forces_slots = _slot_mask(p)
            c.contact_forces_threshold = float(p["threshold"])
        elif f == "base_height_l2":
            c.base_height_target = float(p["target_height"])
    if std is not None:
        c.track_std = std

    # ---- terminations (V/velocity_env_cfg.py:264-268; rough_env_cfg.py:95-109) ----
    c.mask_illegal_slots = 0
    have_timeout = False
    for n, t in _terms(cfg.terminations):
        f = _fname(t.func)
        if f == "time_out":
            have_timeout = bool(_get(t, "time_out", False))
        elif f == "illegal_contact":
            c.mask_illegal_slots = _slot_mask(t.params)
            c.contact_threshold = float(t.params.get("threshold", 1.0))
        else:
            raise NotImplementedError(f"terminations.{n}: mdp.{f} is not implemented in the fused kernel")
    if not have_timeout:
        raise NotImplementedError("terminations: a time_out term (time_out=True) is required")

    # ---- commands (V/velocity_env_cfg.py:90-104; flat_env_cfg.py:46-48) ----
    cmd = cfg.commands.base_velocity
    ctype = getattr(_get(cmd, "class_type"), "__name__", "UniformVelocityCommand")
    if ctype == "UniformVelocityCommandWithDeadzone":  # T/utils/mdp/commands.py:19-96
        c.command_class = 1
        c.velocity_deadzone = float(_get(cmd, "velocity_deadzone", 0.1))
        if c.velocity_deadzone < 0.0:
            raise NotImplementedError("commands.base_velocity.velocity_deadzone must be >= 0")
        c.ang_vel_flip_prob = c.sim_dt / c.episode_length_s  # commands.py:37-38,86 (from the fp32 values the kernel holds)
    elif ctype == "UniformVelocityCommand":
        c.command_class = 0
    else:
        raise NotImplementedError(f"commands.base_velocity: command class {ctype} is not implemented in the fused kernel")
    r = cmd.ranges
    for dst, src in ((c.cmd_lin_x, r.lin_vel_x), (c.cmd_lin_y, r.lin_vel_y), (c.cmd_ang_z, r.ang_vel_z), (c.cmd_resample_time, cmd.resampling_time_range)):
        dst[0], dst[1] = float(src[0]), float(src[1])
    c.heading_command = int(bool(_get(cmd, "heading_command", False)))
    if c.heading_command:
        h = r.heading
        c.cmd_heading[0], c.cmd_heading[1] = float(h[0]), float(h[1])
    c.heading_stiffness = float(_get(cmd, "heading_control_stiffness", 1.0))
    c.rel_standing_envs = float(_get(cmd, "rel_standing_envs", 0.0))
    c.rel_heading_envs = float(_get(cmd, "rel_heading_envs", 1.0))

    # ---- events (V/velocity_env_cfg.py:153-217; rough_env_cfg.py:78-92) ----
    ground_mu = float(_get(_get(_get(cfg.scene, "terrain"), "physics_material"), "static_friction", 1.0))
    c.push_enable = 0
    c.mass_add_range[0] = c.mass_add_range[1] = 0.0
    for i in range(6):
        c.reset_pose_range[i][0] = c.reset_pose_range[i][1] = 0.0
        c.reset_vel_range[i][0] = c.reset_vel_range[i][1] = 0.0
    axes = ["x", "y", "z", "roll", "pitch", "yaw"]
    for n, t in _terms(cfg.events):
        f, p = _fname(t.func), (t.params or {})
        if f == "randomize_rigid_body_material":
            lo, hi = p["static_friction_range"]
            # "multiply" combine with the ground material (V/velocity_env_cfg.py:40-51)
            c.friction_range[0], c.friction_range[1] = float(lo) * ground_mu, float(hi) * ground_mu
            c.friction = 0.5 * (c.friction_range[0] + c.friction_range[1])
            if c.friction_range[0] < 0.3:  # stiff regime of the pyramidal regulariser (1/mu^2): give the fp32 Newton more iterations
                c.solver_iterations = max(c.solver_iterations, 30)
        elif f == "randomize_rigid_body_mass":
            if p.get("operation", "add") != "add":
                raise NotImplementedError(f"events.{n}: only operation='add' is supported")
            c.mass_recompute_inertia = int(bool(p.get("recompute_inertia", True)))
            c.mass_add_range[0], c.mass_add_range[1] = map(float, p["mass_distribution_params"])
        elif f == "apply_external_force_torque":
            if any(abs(float(v)) > 0 for v in tuple(p["force_range"]) + tuple(p["torque_range"])):
                raise NotImplementedError(f"events.{n}: non-zero external wrench is not supported")
        elif f == "reset_root_state_uniform":
            for i, ax in enumerate(axes):
                lo, hi = p.get("pose_range", {}).get(ax, (0.0, 0.0))
                c.reset_pose_range[i][0], c.reset_pose_range[i][1] = float(lo), float(hi)
                lo, hi = p.get("velocity_range", {}).get(ax, (0.0, 0.0))
                c.reset_vel_range[i][0], c.reset_vel_range[i][1] = float(lo), float(hi)
        elif f == "reset_joints_by_scale":
            c.reset_joint_pos_scale[0], c.reset_joint_pos_scale[1] = map(float, p["position_range"])
            c.reset_joint_vel_scale[0], c.reset_joint_vel_scale[1] = map(float, p["velocity_range"])
        elif f == "push_by_setting_velocity":
            vr = p["velocity_range"]
            if set(vr) - {"x", "y"} or tuple(vr.get("x", (0, 0))) != tuple(vr.get("y", (0, 0))):
                raise NotImplementedError(f"events.{n}: pushes must be one range on x and y")
            c.push_enable = 1
            c.push_interval_s[0], c.push_interval_s[1] = map(float, t.interval_range_s)
            c.push_vel_xy[0], c.push_vel_xy[1] = map(float, vr["x"])
        else:
            raise NotImplementedError(f"events.{n}: mdp.{f} is not implemented in the fused kernel")
    return c


def config_to_dict(c: H1v2Config) -> dict:
    """Plain-python view of an H1v2Config (golden fixtures, yaml dumps)."""
    def conv(v):
        if hasattr(v, "__len__"):
            return [conv(x) for x in v]
        return v
    return {name: conv(getattr(c, name)) for name, _ in c._fields_ if name != "reserved"}


class _ActionManagerView:
    def __init__(self, env):
        self._env = env
        self.total_action_dim = NJ
        self.active_terms = ["joint_pos"]

    @property
    def action(self):
        return self._env._last_action

    @property
    def prev_action(self):
        return self._env._prev_action


class _ObservationManagerView:
    def __init__(self, env):
        self._env = env
        self.group_obs_dim = {"policy": (env.sim.obs_dim,)}
        k = env.kernel_cfg
        self.active_terms = {"policy": (["base_lin_vel"] if k.obs_base_lin_vel else []) + ["base_ang_vel", "projected_gravity", "velocity_commands", "joint_pos",
                                        "joint_vel", "actions"] + (["height_scan"] if k.obs_height_scan else [])}

    def compute(self):
        """ObservationManager.compute(): appends to the history like upstream (observation_manager.py:318-355)."""
        return {"policy": self._env.sim.observe()}


class _NamesView:
    def __init__(self, names):
        self.active_terms = list(names)


class _CommandManagerView(_NamesView):
    def __init__(self, env):
        super().__init__(["base_velocity"])
        self._env = env

    def get_command(self, name: str):
        """CommandManager.get_command("base_velocity") -> Tensor[N,3] (v_x, v_y, w_z), as the mdp terms read it upstream."""
        if name != "base_velocity":
            raise KeyError(name)
        return self._env.sim.get_state(["command"])["command"]


class H1v2ManagerBasedRLEnv:
    """ManagerBasedRLEnv drop-in.  step(action) -> (obs_dict, rew, terminated, truncated, extras), all device tensors."""

    is_vector_env = True
    metadata = {"render_modes": [None, "human", "rgb_array"], "isaac_sim_version": "B200-native backend (no Isaac Sim)"}

    def __init__(self, cfg, render_mode: str | None = None, **kwargs):
        import torch

        from .backend import H1v2Sim
        if hasattr(cfg, "validate"):
            cfg.validate()
        self.cfg = cfg
        self.render_mode = render_mode
        self.kernel_cfg = flatten_cfg(cfg)
        seed = _get(cfg, "seed")
        self._seed = 42 if seed is None else int(seed)
        rank, world = self._dist_info()
        n = int(cfg.scene.num_envs)
        self.kernel_cfg.env_id_offset = rank * n  # envs shard by rank; the Philox key uses the global env id
        dev = torch.device(_get(cfg.sim, "device", "cuda:0"))
        self.sim = H1v2Sim(n, self.kernel_cfg, device=dev, seed=self._seed)
        self.num_envs, self.device = n, self.sim.device
        self.physics_dt = float(cfg.sim.dt)
        self.step_dt = self.physics_dt * int(cfg.decimation)
        self.max_episode_length_s = float(cfg.episode_length_s)
        self.max_episode_length = self.sim.max_episode_length
        self.cfg_is_finite_horizon = bool(_get(cfg, "is_finite_horizon", False))
        self.common_step_counter = 0
        self.extras: dict = {}
        self._log_keys = None
        self._last_action = torch.zeros((n, NJ), device=self.device)
        self._prev_action = torch.zeros((n, NJ), device=self.device)
        self.action_manager = _ActionManagerView(self)
        self.observation_manager = _ObservationManagerView(self)
        self.reward_manager = _NamesView([n for n, t in _terms(cfg.rewards) if float(t.weight) != 0.0])
        self.termination_manager = _NamesView(["time_out", "base_contact"])
        self.command_manager = _CommandManagerView(self)
        slots = reward_slots(cfg)
        self._rew_names = list(slots)  # every term that owns a slot is logged, like upstream's Episode_Reward/<term> (0 for weight 0)
        self._rew_slots = [slots[n] for n in self._rew_names]
        # CurriculumManager (modify_reward_weight only): pending (slot, weight, num_steps), applied once the counter passes num_steps
        self._curriculum = curriculum_schedule(cfg)
        self._configure_gym_env_spaces()
        self.obs_buf = {"policy": self.sim.observe()}
        print(f"[INFO]: B200-native environment: {n} envs on {self.device}, step_dt {self.step_dt:.3f} s, obs {self.sim.obs_dim}, seed {self._seed}"
              + (f", rank {rank}/{world}" if world > 1 else ""))

    @staticmethod
    def _dist_info():
        import os
        return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))

    # ---- gym surface (T/utils/cat/cat_env.py:250-275) ----
    def _configure_gym_env_spaces(self):
        import gymnasium as gym
        import numpy as np
        od = self.sim.obs_dim
        self.single_observation_space = gym.spaces.Dict({"policy": gym.spaces.Box(low=-np.inf, high=np.inf, shape=(od,))})
        self.single_action_space = gym.spaces.Box(low=-np.inf, high=np.inf, shape=(NJ,))
        self.observation_space = gym.vector.utils.batch_space(self.single_observation_space, self.num_envs)
        self.action_space = gym.vector.utils.batch_space(self.single_action_space, self.num_envs)

    @property
    def unwrapped(self):
        return self

    @property
    def episode_length_buf(self):
        return self.sim.episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        # assignable (scripts/rsl_rl/train.py:141 learn(init_at_random_ep_len=True)); the kernel keeps its bound buffer
        self.sim.episode_length_buf.copy_(value.to(self.sim.episode_length_buf.dtype))

    def seed(self, seed: int = -1) -> int:
        return self._seed

    # ---- MDP ----
    def reset(self, seed: int | None = None, options: dict | None = None):
        self.sim.reset(None)
        self._last_action.zero_(); self._prev_action.zero_()
        self.sim.episode_length_buf.zero_()
        self.obs_buf = {"policy": self.sim.observe()}
        self.extras = {}
        return self.obs_buf, self.extras

    def step(self, action):
        obs, rew, terminated, truncated = self.sim.step(action)
        self._prev_action, self._last_action = self._last_action, action
        self.common_step_counter += 1
        if self._curriculum:
            self._apply_curriculum()
        self.obs_buf = {"policy": obs}
        self.reward_buf, self.reset_terminated, self.reset_time_outs = rew, terminated, truncated
        self.reset_buf = terminated | truncated
        self.extras = {"log": self._log_dict()}
        return self.obs_buf, rew, terminated, truncated, self.extras

    def _apply_curriculum(self) -> None:
        """mdp.modify_reward_weight (C12/rsl_env_cfg.py:447-497): `if env.common_step_counter > num_steps: weight = w`.  Upstream
        evaluates it inside _reset_idx, i.e. in any step with a reset; at thousands of envs that is every step.  Host integers only."""
        due = [t for t in self._curriculum if self.common_step_counter > t[2]]
        if not due:
            return
        w = list(self.sim.cfg.rew_weight)
        for slot, weight, _ in due:
            w[slot] = weight
        if w != list(self.sim.cfg.rew_weight):
            self.sim.set_reward_weights(w)
        self._curriculum = [t for t in self._curriculum if t not in due]

    def _log_dict(self) -> dict:
        """extras["log"] (T/utils/cat/cat_env.py:217-245): 0-d device tensors, no host sync.  Values are those of the
        most recent step in which any env reset (upstream only writes the keys on such steps).  One gather per step: the
        entries are a snapshot, they do not change when the kernel updates its log vector later."""
        if self._log_keys is None:
            import torch
            self._log_keys = [f"Episode_Reward/{n}" for n in self._rew_names] + [
                "Episode_Termination/time_out", "Episode_Termination/base_contact",
                "Metrics/base_velocity/error_vel_xy", "Metrics/base_velocity/error_vel_yaw"]
            idx = [LOG_REW0 + s for s in self._rew_slots] + [LOG_TERM_TIMEOUT, LOG_TERM_CONTACT, LOG_ERR_XY, LOG_ERR_YAW]
            self._log_index = torch.tensor(idx, dtype=torch.long, device=self.device)
        log = dict(zip(self._log_keys, self.sim.log_buf[self._log_index].unbind(0)))
        if self.kernel_cfg.terrain_enable and self.kernel_cfg.terrain_curriculum:  # CurriculumManager: mdp.terrain_levels_vel returns the mean level
            log["Curriculum/terrain_levels"] = self.sim.terrain_log_buf[1].clone()
        return log

    def render(self, recompute: bool = False):
        return None

    def close(self):
        if getattr(self, "sim", None) is not None:
            self.sim.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class H1v2CaTEnv(H1v2ManagerBasedRLEnv):
    """CaTEnv drop-in (packages/biped_tasks/biped_tasks/utils/cat/cat_env.py:28-275): the same fused step followed by the
    Constraints-as-Terminations tail.  step(action) -> (obs_dict, reward * (1 - p), dones, time_outs, extras) where dones is a
    FLOAT tensor (the termination probability p, 1 for the envs that reset), exactly what cat_env.py:193 returns."""

    def __init__(self, cfg, render_mode: str | None = None, **kwargs):
        super().__init__(cfg, render_mode, **kwargs)
        if not self.kernel_cfg.cat_enable:
            raise ValueError("H1v2CaTEnv needs a cfg with a `constraints` group (use H1v2ManagerBasedRLEnv otherwise)")
        terms = constraint_terms(cfg)
        self.constraint_manager = _NamesView(list(terms))
        self._cstr_names, self._cstr_idx = list(terms), [terms[n] for n in terms]
        self._cstr_schedule = constraint_curriculum(cfg)
        self._cstr_base = list(self.kernel_cfg.cat_max_p)
        self._cstr_log = None
        if self._cstr_schedule:  # CurriculumManager.compute runs in the first _reset_idx, i.e. before the first step
            self.sim.set_constraint_max_p(constraint_max_p(self._cstr_schedule, self._cstr_base, 0))

    def step(self, action):
        import torch
        obs, rew, dones, truncated = self.sim.cat_step(action)
        self._prev_action, self._last_action = self._last_action, action
        self.common_step_counter += 1
        if self._curriculum:
            self._apply_curriculum()
        if self._cstr_schedule:
            self.sim.set_constraint_max_p(constraint_max_p(self._cstr_schedule, self._cstr_base, self.common_step_counter))
        self.obs_buf = {"policy": obs}
        self.reward_buf, self.reset_time_outs = rew, truncated
        self.reset_buf = dones >= 1.0
        self.reset_terminated = self.reset_buf & ~truncated
        log = self._log_dict()
        # Episode_Constraint_violation / _probability (constraint_manager.py:185-203): means over the envs that reset in this step,
        # the last such values otherwise; computed on the device from the tail's accumulators, no host sync
        acc = self.sim.cat_acc_buf
        cur = acc[:20] / acc[20].clamp(min=1.0)
        self._cstr_log = cur if self._cstr_log is None else torch.where(acc[20] > 0, cur, self._cstr_log)
        for i, n in zip(self._cstr_idx, self._cstr_names):
            log[f"Episode_Constraint_violation/{n}"] = self._cstr_log[i]
            log[f"Episode_Constraint_probability/{n}"] = self._cstr_log[10 + i]
        self.extras = {"log": log}
        return self.obs_buf, rew, dones, truncated, self.extras
