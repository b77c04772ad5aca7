// h1v2_config.cpp -- resolved configuration of Isaac-Velocity-Flat-H12_12dof-v0 as a POD (host only).
//
// Every number cites where the reference pins it (paths relative to the reference root):
//   V  = packages/biped_tasks/biped_tasks/tasks/locomotion/velocity
//   C12= V/config/h12_12dof ;  A = packages/biped_assets/biped_assets
#include <cmath>
#include <cstring>

#include "../../include/h1v2_b200.h"
#include "../../include/h1v2_model_h12.h"

extern "C" int h1v2_default_config(H1v2Config* c) {
  if (!c) return -1;
  std::memset(c, 0, sizeof(*c));
  // timing: V/velocity_env_cfg.py:302-305
  c->sim_dt = 0.005f;
  c->decimation = 4;
  c->episode_length_s = 20.0f;
  // action: V/velocity_env_cfg.py:111 ; default pose A/robots/h12.py:38-53 == h12_12dof.xml:362-368
  c->action_scale = 0.5f;
  for (int j = 0; j < H1V2_NJ; j++) c->default_joint_pos[j] = (float)h1v2_key_qpos[7 + j];
  // joint_names=[".*"] without preserve_order -> PhysX breadth-first order, L/R interleaved
  static const int perm[12] = {0, 6, 1, 7, 2, 8, 3, 9, 4, 10, 5, 11};
  for (int i = 0; i < H1V2_NJ; i++) c->joint_perm[i] = perm[i];
  // actuators: A/robots/h12.py:58-113 (MJCF order: yaw pitch roll knee ankle_pitch ankle_roll)
  static const float kp6[6] = {200, 200, 200, 300, 40, 40}, kd6[6] = {2.5f, 2.5f, 2.5f, 4, 2, 2};
  static const float ef6[6] = {220, 220, 220, 360, 45, 45};
  for (int j = 0; j < H1V2_NJ; j++) {
    c->kp[j] = kp6[j % 6];
    c->kd[j] = kd6[j % 6];
    c->effort_limit[j] = ef6[j % 6];
  }
  c->min_delay = 0;
  c->max_delay = 5;
  // physics: h12_12dof.xml:4-7 defaults, :73..134 ranges; MuJoCo defaults for solref/solimp;
  // contact solref = equal-weight mix of the robot geoms (0.005 1) and the floor (default 0.02 1)
  c->gravity = 9.81f;
  c->friction = 0.8f;  // robot static 0.8 x ground 1.0, "multiply": V/velocity_env_cfg.py:40-51,153-163
  for (int d = 0; d < 18; d++) {
    c->dof_damping[d] = (float)h1v2_dof_damping[d];
    c->dof_armature[d] = (float)h1v2_dof_armature[d];
    c->dof_frictionloss[d] = (float)h1v2_dof_frictionloss[d];
  }
  for (int j = 0; j < H1V2_NJ; j++) {
    c->act_frc_limit[j] = (float)h1v2_jnt_frcrange[j];
    c->joint_range[j][0] = (float)h1v2_jnt_range[j][0];
    c->joint_range[j][1] = (float)h1v2_jnt_range[j][1];
  }
  c->contact_solref[0] = 0.5f * (0.005f + 0.02f);
  c->contact_solref[1] = 1.0f;
  static const float solimp[5] = {0.9f, 0.95f, 0.001f, 0.5f, 2.0f};
  std::memcpy(c->contact_solimp, solimp, sizeof(solimp));
  std::memcpy(c->floss_solimp, solimp, sizeof(solimp));
  std::memcpy(c->limit_solimp, solimp, sizeof(solimp));
  c->floss_solref[0] = 0.02f; c->floss_solref[1] = 1.0f;
  c->limit_solref[0] = 0.02f; c->limit_solref[1] = 1.0f;
  c->solver_iterations = 12;  // Newton cap (MuJoCo: 100).  Enough at the task's mu = 0.8 (< 1e-5 of solves reach it); the kernel ends when its
                              // slowest warp does, so a larger cap costs ~9 % even when almost no solve uses it.  flatten_cfg raises it to
                              // 30 when the friction range reaches below 0.3 (10x fewer non-converged solves at mu ~ 0.1, tools/diag_rand3.py)
  c->solver_tolerance = 1e-5f;  // on the scaled gradient norm; the fp32 noise floor of that norm is ~3e-6 (profiles/r1_notes.md), below it iterations only chase rounding
  c->solver_step_tolerance = 1e-3f;
  c->solver_vel_tolerance = 2e-4f;  // rad/s per physics step: h * max_j |grad_j| * invweight0_j.  The gradient-norm test alone lets 2.3e-3 N m pass on a
                                    // swinging ankle (1/invweight0 = 0.011 kg m^2): 1e-3 rad/s, the whole single-step tolerance; free on B200 (the residual of a
                                    // converged Newton step is almost always far below both tests: 2.55 -> 2.55 iterations per substep, profiles/r2_notes.md)
  c->solver_ls_tolerance = 0.3f;  // MuJoCo default 0.01; 0.01..0.3 give the same Newton iteration histogram and parity on B200 (tools/diag_lstol.py), 0.3 is 5 % faster
  // observations: V/velocity_env_cfg.py:123-132 ; C12/flat_env_cfg.py:25-27
  c->history_length = 10;
  c->enable_corruption = 1;
  c->noise_ang_vel = 0.2f;
  c->noise_gravity = 0.05f;
  c->noise_joint_pos = 0.01f;
  c->noise_joint_vel = 1.5f;
  c->scale_ang_vel = c->scale_gravity = c->scale_cmd = c->scale_joint_pos = c->scale_joint_vel = c->scale_action = 1.0f;
  // rewards: C12/rough_env_cfg.py:18-62,112-120 ; C12/flat_env_cfg.py:35-44 ; V/velocity_env_cfg.py:225-257
  c->rew_weight[H1V2_REW_TERMINATION] = -200.0f;
  c->rew_weight[H1V2_REW_TRACK_LIN_XY_YAW] = 1.0f;
  c->rew_weight[H1V2_REW_TRACK_ANG_Z_WORLD] = 1.0f;
  c->rew_weight[H1V2_REW_FEET_AIR_BIPED] = 0.75f;
  c->rew_weight[H1V2_REW_FEET_SLIDE] = -0.25f;
  c->rew_weight[H1V2_REW_DOF_POS_LIMITS] = -1.0f;
  c->rew_weight[H1V2_REW_JOINT_DEV_HIP] = -0.2f;
  c->rew_weight[H1V2_REW_ANG_VEL_XY] = -0.05f;
  c->rew_weight[H1V2_REW_TORQUES] = -2.0e-6f;
  c->rew_weight[H1V2_REW_DOF_ACC] = -1.0e-7f;
  c->rew_weight[H1V2_REW_ACTION_RATE] = -0.005f;
  c->rew_weight[H1V2_REW_FLAT_ORI] = -1.0f;
  c->track_std = 0.5f;
  c->feet_air_threshold = 0.4f;
  c->contact_threshold = 1.0f;
  c->soft_limit_factor = 0.9f;  // A/robots/h12.py:56
  c->base_height_target = 0.98f;
  c->mask_pos_limits = (1u << 4) | (1u << 5) | (1u << 10) | (1u << 11);
  c->mask_joint_dev = (1u << 0) | (1u << 2) | (1u << 6) | (1u << 8);
  c->mask_torques = 0xFFFu;
  c->mask_undesired_slots = (1u << 2) | (1u << 3);
  // second-instance terms and contact_forces of the Rsl cfg (C12/rsl_env_cfg.py:353-401): present, weight 0 by default
  c->mask_pos_limits_b = (1u << 0) | (1u << 2) | (1u << 6) | (1u << 8);
  c->mask_joint_dev_b = (1u << 4) | (1u << 5) | (1u << 10) | (1u << 11);
  c->mask_contact_forces_slots = 0x3u;
  c->contact_forces_threshold = 800.0f;
  // terminations: C12/rough_env_cfg.py:95-109 (bodies with colliders: knee links, torso, pelvis)
  c->mask_illegal_slots = (1u << 2) | (1u << 3) | (1u << 4) | (1u << 5);
  // commands: V/velocity_env_cfg.py:90-104 ; C12/flat_env_cfg.py:46-48
  c->cmd_lin_x[0] = 0.0f; c->cmd_lin_x[1] = 1.0f;
  c->cmd_lin_y[0] = -0.5f; c->cmd_lin_y[1] = 0.5f;
  c->cmd_ang_z[0] = -1.0f; c->cmd_ang_z[1] = 1.0f;
  c->cmd_heading[0] = -(float)M_PI; c->cmd_heading[1] = (float)M_PI;
  c->cmd_resample_time[0] = c->cmd_resample_time[1] = 10.0f;
  c->rel_standing_envs = 0.02f;
  c->rel_heading_envs = 1.0f;
  c->heading_stiffness = 0.5f;
  c->heading_command = 1;
  c->command_class = 0;
  c->velocity_deadzone = 0.0f;
  c->ang_vel_flip_prob = 0.005f / 20.0f;  // T/utils/mdp/commands.py:86 physics_dt / max_episode_length_s
  // reset events: C12/rough_env_cfg.py:78-92 ; V/velocity_env_cfg.py:200-209
  c->reset_pose_range[0][0] = -0.5f; c->reset_pose_range[0][1] = 0.5f;
  c->reset_pose_range[1][0] = -0.5f; c->reset_pose_range[1][1] = 0.5f;
  c->reset_pose_range[5][0] = -3.14f; c->reset_pose_range[5][1] = 3.14f;
  c->reset_joint_pos_scale[0] = c->reset_joint_pos_scale[1] = 1.0f;
  c->init_root_height = 1.05f;  // A/robots/h12.py:38
  c->push_enable = 0;           // C12/rough_env_cfg.py:78
  c->push_interval_s[0] = 10.0f; c->push_interval_s[1] = 15.0f;  // V/velocity_env_cfg.py:212-217
  c->push_vel_xy[0] = -0.5f; c->push_vel_xy[1] = 0.5f;
  c->friction_range[0] = c->friction_range[1] = c->friction;
  c->env_id_offset = 0;
  c->env_spacing = 2.5f;
  c->joint_vel_limit = 100.0f;  // A/robots/h12.py:66,89,103 velocity_limit
  // isaaclab 2.1.0 reports link velocities at the link's COM: pelvis link inertial origin, A/models/h12/h12_12dof.urdf:16 (== h12_12dof.xml:67)
  c->root_link_com[0] = -0.0004f; c->root_link_com[1] = 3.7e-05f; c->root_link_com[2] = -0.046864f;
  c->body_vel_at_com = 1;
  // Constraints-as-Terminations tail: off; parameters of config/h12_12dof/cat_env_cfg.py:336-431, ConstraintManager defaults
  c->mass_recompute_inertia = 1;
  c->cat_enable = 0;
  c->cat_tau = 0.95f; c->cat_min_p = 0.0f;
  for (int t = 0; t < H1V2_NUM_CSTR; t++) c->cat_max_p[t] = 0.25f;
  c->cat_max_p[H1V2_CSTR_CONTACT] = 1.0f;
  c->cat_contact_slots = (1u << 2) | (1u << 3) | (1u << 4) | (1u << 5);
  c->cat_foot_force_limit = 750.0f;
  c->cat_no_move_deadzone = 0.2f; c->cat_no_move_vel_limit = 6.0f;
  c->cat_orientation_limit = 0.1f;
  c->cat_height = 1.0f; c->cat_height_std = 0.05f;
  c->cat_clearance_min_height = 0.1f; c->cat_clearance_deadzone = 0.2f;
  c->runaway_vel = 1000.0f;     // A/robots/h12.py:27-28 max_linear_velocity / max_angular_velocity
  return 0;
}

// Resolved cfg of Isaac-Velocity-Rsl-H12_12dof-v0 (C12/__init__.py:85-93 -> C12/rsl_env_cfg.py:503-540 H12_12dof_EnvCfg), the task
// the reference's deployed policies were trained on.  Same robot model and solver as the Flat id.
extern "C" int h1v2_rsl_config(H1v2Config* c) {
  if (h1v2_default_config(c) != 0) return -1;
  // actions: C12/rsl_env_cfg.py:104-127 (scale 0.25, preserve_order over the MJCF joint list)
  c->action_scale = 0.25f;
  for (int i = 0; i < H1V2_NJ; i++) c->joint_perm[i] = i;
  // robot: A/robots/h12.py:117-206 H12_12DOF_IDEAL -- IdealPDActuatorCfg, same gains and limits, no delay line
  c->min_delay = c->max_delay = 0;
  // observations: C12/rsl_env_cfg.py:133-203 (history 6, gyro x0.25, joint velocity x0.05)
  c->history_length = 6;
  c->scale_ang_vel = 0.25f;
  c->scale_joint_vel = 0.05f;
  // rewards: C12/rsl_env_cfg.py:278-407
  for (int t = 0; t < H1V2_NUM_REW; t++) c->rew_weight[t] = 0.0f;
  c->rew_weight[H1V2_REW_TRACK_LIN_XY_BASE] = 1.0f;      // :283-287 track_lin_vel_xy_exp, base frame
  c->rew_weight[H1V2_REW_TRACK_ANG_Z_BASE] = 0.5f;       // :288-292
  c->rew_weight[H1V2_REW_FEET_AIR_BIPED] = 0.75f;        // :295-305
  c->rew_weight[H1V2_REW_FEET_SLIDE] = -0.25f;           // :306-315
  c->rew_weight[H1V2_REW_FLAT_ORI] = -1.0f;              // :318-321
  c->rew_weight[H1V2_REW_BASE_HEIGHT] = -0.2f;           // :322-328
  c->base_height_target = 1.0f;
  c->rew_weight[H1V2_REW_TORQUES] = -1.0e-5f;            // :331-334 (all joints)
  c->rew_weight[H1V2_REW_JOINT_VEL] = -1.0e-3f;          // :335-338
  c->rew_weight[H1V2_REW_DOF_ACC] = -1.0e-7f;            // :339-342
  c->rew_weight[H1V2_REW_JOINT_DEV_HIP] = -0.2f;         // :343-357 hip yaw + roll
  c->mask_joint_dev = (1u << 0) | (1u << 2) | (1u << 6) | (1u << 8);
  c->rew_weight[H1V2_REW_JOINT_DEV_B] = -0.2f;           // :358-372 ankle roll + pitch
  c->mask_joint_dev_b = (1u << 4) | (1u << 5) | (1u << 10) | (1u << 11);
  c->rew_weight[H1V2_REW_DOF_POS_LIMITS] = -0.2f;        // :373-379 ankles
  c->mask_pos_limits = (1u << 4) | (1u << 5) | (1u << 10) | (1u << 11);
  c->rew_weight[H1V2_REW_DOF_POS_LIMITS_B] = -0.2f;      // :380-386 hip yaw + roll
  c->mask_pos_limits_b = (1u << 0) | (1u << 2) | (1u << 6) | (1u << 8);
  c->rew_weight[H1V2_REW_ACTION_RATE] = -0.01f;          // :389-392
  c->rew_weight[H1V2_REW_CONTACT_FORCES] = -1.0e-3f;     // :395-404 feet, 800 N
  c->mask_contact_forces_slots = 0x3u;
  c->contact_forces_threshold = 800.0f;
  c->rew_weight[H1V2_REW_TERMINATION] = -200.0f;         // :407
  // commands: C12/rsl_env_cfg.py:82-99 UniformVelocityCommandWithDeadzoneCfg, no heading control
  c->command_class = 1;
  c->velocity_deadzone = 0.0f;
  c->ang_vel_flip_prob = 0.005f / 20.0f;
  c->cmd_lin_x[0] = -1.0f; c->cmd_lin_x[1] = 1.0f;
  c->cmd_lin_y[0] = -1.0f; c->cmd_lin_y[1] = 1.0f;
  c->cmd_resample_time[0] = 5.0f; c->cmd_resample_time[1] = 8.0f;
  c->heading_command = 0;
  c->heading_stiffness = 1.0f;
  // events: C12/rsl_env_cfg.py:208-272 (friction 0.1..1.25 x ground 1.0, pushes every U(5,8) s of +-1 m/s)
  c->friction_range[0] = 0.1f; c->friction_range[1] = 1.25f;
  c->friction = 0.5f * (c->friction_range[0] + c->friction_range[1]);
  c->solver_iterations = 30;  // friction below 0.3: the stiff regime of the pyramidal regulariser
  c->reset_joint_vel_scale[0] = c->reset_joint_vel_scale[1] = 1.0f;
  c->push_enable = 1;
  c->push_interval_s[0] = 5.0f; c->push_interval_s[1] = 8.0f;
  c->push_vel_xy[0] = -1.0f; c->push_vel_xy[1] = 1.0f;
  return 0;
}

// Resolved cfg of Isaac-Velocity-CaT-Flat-H12_12dof-v0 (C12/__init__.py:63-71 -> C12/cat_env_cfg.py:528-565 H12_12dof_EnvCfg): the Rsl
// cfg's observation / action / command format on the delayed-PD robot, a 7-term reward, and the Constraints-as-Terminations tail.
extern "C" int h1v2_cat_config(H1v2Config* c) {
  H1v2Config d;
  if (h1v2_rsl_config(c) != 0 || h1v2_default_config(&d) != 0) return -1;
  c->min_delay = d.min_delay; c->max_delay = d.max_delay;  // C12/cat_env_cfg.py:44 H12_12DOF (DelayedPDActuatorCfg, A/robots/h12.py:18-114)
  for (int t = 0; t < H1V2_NUM_REW; t++) c->rew_weight[t] = 0.0f;
  c->rew_weight[H1V2_REW_TRACK_LIN_XY_BASE] = 1.0f;   // :304-308
  c->rew_weight[H1V2_REW_TRACK_ANG_Z_BASE] = 0.5f;    // :309-313
  c->rew_weight[H1V2_REW_TORQUES] = -1.0e-5f;         // :316
  c->rew_weight[H1V2_REW_DOF_ACC] = -2.5e-7f;         // :317
  c->rew_weight[H1V2_REW_JOINT_VEL] = -1.0e-3f;       // :318
  c->rew_weight[H1V2_REW_ACTION_RATE] = -0.01f;       // :319
  c->rew_weight[H1V2_REW_JOINT_DEV_HIP] = -0.1f;      // :320-331 hip yaw + roll, ankle pitch + roll of both legs
  c->mask_joint_dev = (1u << 0) | (1u << 2) | (1u << 4) | (1u << 5) | (1u << 6) | (1u << 8) | (1u << 10) | (1u << 11);
  c->base_height_target = d.base_height_target;        // no base_height_l2 term in this cfg
  c->mass_add_range[0] = 0.0f; c->mass_add_range[1] = 6.0f;  // :242-251 add_base_mass, recompute_inertia=False
  c->mass_recompute_inertia = 0;
  c->velocity_deadzone = 0.2f;                          // :48,114 VELOCITY_DEADZONE
  c->cat_enable = 1;                                    // :336-431; the cat_* parameters are h1v2_default_config's
  return 0;
}


// Resolved cfg of Isaac-Velocity-Rough-H12_12dof-v0 (C12/__init__.py:17-26 -> C12/rough_env_cfg.py:65-125 H12_12dof_RoughEnvCfg on
// V/velocity_env_cfg.py:280-324 LocomotionVelocityRoughEnvCfg).  Same robot, actuators and solver as the Flat id.
// Terrain generator: the reference's in-tree cfg packages/biped_tasks/biped_tasks/utils/mdp/terrains.py:11-28 (one HfRandomUniformTerrainCfg
// sub-terrain); upstream isaaclab's ROUGH_TERRAINS_CFG of the same name also holds mesh stairs / boxes / slopes, which this backend refuses.
extern "C" int h1v2_rough_config(H1v2Config* c) {
  if (h1v2_default_config(c) != 0) return -1;
  // rewards: C12/rough_env_cfg.py:18-62 (H12_12dof_Rewards) and :112-120 -- what C12/flat_env_cfg.py:35-44 overrides is undone here
  c->rew_weight[H1V2_REW_TORQUES] = -1.5e-7f;            // :114
  c->rew_weight[H1V2_REW_DOF_ACC] = -1.25e-7f;           // :116
  c->rew_weight[H1V2_REW_FEET_AIR_BIPED] = 0.25f;        // :34-42
  // commands: C12/rough_env_cfg.py:122-125
  c->cmd_lin_y[0] = c->cmd_lin_y[1] = 0.0f;
  // observations: V/velocity_env_cfg.py:119-142 -- base_lin_vel first, height_scan last, no history
  c->history_length = 1;
  c->obs_base_lin_vel = 1; c->noise_lin_vel = 0.1f; c->scale_lin_vel = 1.0f;                 // :123
  c->obs_height_scan = 1;                                                                     // :133-138, sensor :61-68
  c->scan_size[0] = 1.6f; c->scan_size[1] = 1.0f; c->scan_resolution = 0.1f;                  // GridPatternCfg(resolution=0.1, size=[1.6, 1.0])
  c->scan_offset = 0.5f;                                                                      // mdp.height_scan(offset=0.5) [upstream default]
  c->noise_height_scan = 0.1f; c->scale_height_scan = 1.0f; c->scan_clip[0] = -1.0f; c->scan_clip[1] = 1.0f;
  // terrain: V/velocity_env_cfg.py:40-58 (generator, max_init_terrain_level 5) with T/utils/mdp/terrains.py:11-28
  c->terrain_enable = 1;
  c->terrain_rows = 10; c->terrain_cols = 20;            // num_rows, num_cols
  c->terrain_tile_size = 8.0f;                           // size=(8.0, 8.0)
  c->terrain_hscale = 0.1f; c->terrain_vscale = 0.005f;  // horizontal_scale, vertical_scale
  c->terrain_level_min = 0; c->terrain_level_max = 4; c->terrain_level_step = 1;  // noise_range=(0.0, 0.02), noise_step=0.005 in units of 0.005
  c->terrain_border_px = 3;                              // border_width=0.25 -> int(0.25 / 0.1) + 1
  c->terrain_max_init_level = 5;
  c->terrain_curriculum = 1;                             // V/velocity_env_cfg.py:275 terrain_levels_vel
  return 0;
}
