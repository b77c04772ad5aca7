/*
 * h1v2_b200.h -- C-ABI of the B200-native batched simulation backend for the Unitree H1-2
 * velocity-tracking tasks  Isaac-Velocity-Flat-H12_12dof-v0,  Isaac-Velocity-Rsl-H12_12dof-v0,
 * Isaac-Velocity-CaT-Flat-H12_12dof-v0  and  Isaac-Velocity-Rough-H12_12dof-v0.
 *
 * The reference (olivier-stasse/h1v2-Isaac) has no FFI of its own: its seam is the Python
 * ManagerBasedRLEnv.step contract.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference root):
 *
 *   h1v2_create        <- gym.make(task, cfg=env_cfg)            scripts/rsl_rl/train.py:102
 *                         -> ManagerBasedRLEnv.__init__ / load_managers
 *                            packages/biped_tasks/biped_tasks/utils/cat/cat_env.py:31-93
 *   h1v2_default_config<- resolved cfg of the Flat id
 *                            .../velocity/config/h12_12dof/flat_env_cfg.py:13-48,
 *                            .../velocity/config/h12_12dof/rough_env_cfg.py:18-125,
 *                            .../velocity/velocity_env_cfg.py:86-324,
 *                            packages/biped_assets/biped_assets/robots/h12.py:18-114
 *   h1v2_rsl_config    <- resolved cfg of the Rsl id
 *                            .../velocity/config/h12_12dof/rsl_env_cfg.py:44-540, robots/h12.py:117-206
 *   h1v2_rough_config  <- resolved cfg of the Rough id
 *                            .../velocity/config/h12_12dof/rough_env_cfg.py:65-125, .../velocity/velocity_env_cfg.py:36-324,
 *                            packages/biped_tasks/biped_tasks/utils/mdp/terrains.py:11-28
 *   h1v2_get_terrain /
 *   h1v2_set_terrain   <- TerrainImporter / TerrainGenerator height field (upstream isaaclab; cfg utils/mdp/terrains.py:11-28)
 *   h1v2_get_terrain_log <- CurriculumManager: mdp.terrain_levels_vel   .../velocity/mdp/curriculums.py:21-52
 *   h1v2_set_reward_weights <- CurriculumManager: mdp.modify_reward_weight   rsl_env_cfg.py:447-497
 *   h1v2_step          <- ManagerBasedRLEnv.step(action)         utils/cat/cat_env.py:95-193
 *   h1v2_step_host     <- same call with HOST buffers (what a non-torch caller binds; used for e2e timing)
 *   h1v2_reset         <- ManagerBasedRLEnv.reset / _reset_idx   utils/cat/cat_env.py:195-248
 *   h1v2_get_state /
 *   h1v2_set_state     <- robot.data.* reads, write_root_*_to_sim / write_joint_state_to_sim (upstream
 *                         isaaclab; needed for identical-state parity tests)
 *   h1v2_get_log       <- extras["log"] filled in _reset_idx     utils/cat/cat_env.py:217-245
 *   h1v2_bind_episode_length <- env.episode_length_buf (assignable LongTensor; train.py:141
 *                         learn(init_at_random_ep_len=True))
 *
 * Conventions: plain pointers and sizes only; every device pointer is borrowed for the duration of the
 * call; persistent state is owned by the handle; calls enqueue work on the given CUDA stream and never
 * synchronise (except *_host and get_log_host, which must); return 0 on success, negative on error with a
 * thread-local message from h1v2_last_error().  One handle per process per GPU is the intended use; a process that holds
 * handles on several GPUs may call them with any current device (each call runs on its handle's device and restores the
 * caller's).  env_ids / n of h1v2_reset: a NULL list means "all envs" whatever n is.
 */
#ifndef H1V2_B200_H
#define H1V2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H1V2_NJ 12           /* actuated joints */
#define H1V2_NUM_REW 22      /* reward-term slots (union of the Flat / base / Rsl cfgs) */
#define H1V2_NUM_SLOT 6      /* contact-sensor bodies: 0,1 feet L/R; 2,3 knee_link L/R; 4 torso_link; 5 pelvis */
#define H1V2_OBS_TERM_DIM 45 /* ang_vel 3 | proj_g 3 | cmd 3 | q-q0 12 | qd 12 | last_action 12 */
#define H1V2_MAX_SCAN 256   /* most height-scan rays per env (the Rough id has 17 x 11 = 187) */
#define H1V2_MAX_HISTORY 10 /* the reference tasks use 10 (Flat), 6 (Rsl) and 1 (deploy) */
#define H1V2_LOG_DIM 32      /* see h1v2_get_log */
#define H1V2_NUM_CSTR 10     /* constraint terms of the CaT tail, order of H1V2_CSTR_* */
#define H1V2_CSTR_COLS 56    /* their columns: 1 + 12 + 12 + 12 + 2 + 12 + 1 + 1 + 1 + 2 */
#define H1V2_ACTION_ABS_MAX 1.0e6f /* an env whose action is non-finite or larger than this in magnitude is force-reset in that step
                                      (zero reward, terminated, counted in H1V2_LOG_NAN_RESETS), like a non-finite state */

/* reward term slots; weights[i]==0 means "term not in the cfg" (RewardManager skips zero weights) */
enum {
  H1V2_REW_TERMINATION = 0,        /* is_terminated                               rough_env_cfg.py:22 */
  H1V2_REW_TRACK_LIN_XY_YAW = 1,   /* track_lin_vel_xy_yaw_frame_exp              rough_env_cfg.py:24-28 */
  H1V2_REW_TRACK_ANG_Z_WORLD = 2,  /* track_ang_vel_z_world_exp                   rough_env_cfg.py:29-33 */
  H1V2_REW_FEET_AIR_BIPED = 3,     /* feet_air_time_positive_biped                mdp/rewards.py:38-62 */
  H1V2_REW_FEET_SLIDE = 4,         /* feet_slide                                  rough_env_cfg.py:43-50 */
  H1V2_REW_DOF_POS_LIMITS = 5,     /* joint_pos_limits (ankles)                   rough_env_cfg.py:52-56 */
  H1V2_REW_JOINT_DEV_HIP = 6,      /* joint_deviation_l1 (hip yaw+roll)           rough_env_cfg.py:58-62 */
  H1V2_REW_ANG_VEL_XY = 7,         /* ang_vel_xy_l2                               velocity_env_cfg.py:238 */
  H1V2_REW_TORQUES = 8,            /* joint_torques_l2                            velocity_env_cfg.py:239 */
  H1V2_REW_DOF_ACC = 9,            /* joint_acc_l2                                velocity_env_cfg.py:240 */
  H1V2_REW_ACTION_RATE = 10,       /* action_rate_l2                              velocity_env_cfg.py:241 */
  H1V2_REW_FLAT_ORI = 11,          /* flat_orientation_l2                         velocity_env_cfg.py:256 */
  H1V2_REW_LIN_VEL_Z = 12,         /* lin_vel_z_l2                                velocity_env_cfg.py:237 */
  H1V2_REW_UNDESIRED_CONTACTS = 13,/* undesired_contacts (knee links)             velocity_env_cfg.py:251-255 */
  H1V2_REW_TRACK_LIN_XY_BASE = 14, /* track_lin_vel_xy_exp (base frame)           velocity_env_cfg.py:226-230 */
  H1V2_REW_TRACK_ANG_Z_BASE = 15,  /* track_ang_vel_z_exp (base frame)            velocity_env_cfg.py:231-235 */
  H1V2_REW_FEET_AIR_L2 = 16,       /* feet_air_time                               mdp/rewards.py:13-35 */
  H1V2_REW_JOINT_VEL = 17,         /* joint_vel_l2                                rsl_env_cfg.py:335-338 */
  H1V2_REW_BASE_HEIGHT = 18,       /* base_height_l2                              rsl_env_cfg.py:322-328 */
  H1V2_REW_CONTACT_FORCES = 19,    /* contact_forces (own threshold and bodies)   rsl_env_cfg.py:395-404 */
  H1V2_REW_DOF_POS_LIMITS_B = 20,  /* second joint_pos_limits term (hips)         rsl_env_cfg.py:380-386 */
  H1V2_REW_JOINT_DEV_B = 21        /* second joint_deviation_l1 term (ankles)     rsl_env_cfg.py:358-372 */
};

/* constraint terms (utils/cat/constraints.py; parameters config/h12_12dof/cat_env_cfg.py:336-431) and their first column */
enum {
  H1V2_CSTR_CONTACT = 0,         /* contact                :86-99    col 0      */
  H1V2_CSTR_JOINT_POS = 1,       /* joint_position_limits  :22-31    col 1..12  (MJCF joint order) */
  H1V2_CSTR_JOINT_VEL = 2,       /* joint_velocity_limits  :34-43    col 13..24 */
  H1V2_CSTR_JOINT_TORQUE = 3,    /* joint_torque_limits    :46-55    col 25..36 */
  H1V2_CSTR_FOOT_FORCE = 4,      /* foot_contact_force     :161-168  col 37..38 */
  H1V2_CSTR_NO_MOVE = 5,         /* no_move                :197-235  col 39..50 */
  H1V2_CSTR_ORIENTATION = 6,     /* base_orientation       :102-108  col 51     */
  H1V2_CSTR_HEIGHT = 7,          /* base_height            :256-272  col 52     */
  H1V2_CSTR_FOOT_CONTACT = 8,    /* foot_contact           :171-194  col 53     */
  H1V2_CSTR_CLEARANCE = 9        /* foot_clearance         :275-308  col 54..55 */
};

/* layout of the log vector returned by h1v2_get_log (all float):
 *   [0]                 number of envs reset in the last step that reset anything (count)
 *   [1 .. NUM_REW]      Episode_Reward/<term>   = mean over reset envs of episode_sum / max_episode_length_s
 *   [23], [24]          Episode_Termination/time_out, /base_contact  (counts, as upstream's count_nonzero)
 *   [25], [26]          Metrics/base_velocity/error_vel_xy, error_vel_yaw (mean over reset envs)
 *   [27]                number of envs force-reset by the non-finite-state guard (cumulative)
 *   [28]                max Newton iterations used by any env in the last step
 *   [29]                number of (env,substep) solves that hit the iteration cap in the last step
 *   [30]                total Newton iterations over all (env,substep) solves of the last step
 *   [31]                (leg,substep) pairs of the last step whose contact list overflowed (points beyond 5 are dropped;
 *                       only reachable with shin / torso / pelvis on the ground, i.e. in the step that terminates the env)
 */
#define H1V2_LOG_COUNT 0
#define H1V2_LOG_REW0 1
#define H1V2_LOG_TERM_TIMEOUT 23
#define H1V2_LOG_TERM_CONTACT 24
#define H1V2_LOG_ERR_XY 25
#define H1V2_LOG_ERR_YAW 26
#define H1V2_LOG_NAN_RESETS 27
#define H1V2_LOG_MAX_ITERS 28
#define H1V2_LOG_CAP_HITS 29
#define H1V2_LOG_SUM_ITERS 30
#define H1V2_LOG_CONTACT_OVERFLOW 31 /* (leg,substep) pairs of the last step with more penetrating points than the 5-slot list holds */

typedef struct H1v2Config {
  /* ---- timing (velocity_env_cfg.py:302-305) ---- */
  float sim_dt;            /* 0.005 */
  int32_t decimation;      /* 4 */
  float episode_length_s;  /* 20 -> max_episode_length = ceil(20/0.02) = 1000 */
  /* ---- action: JointPositionAction (velocity_env_cfg.py:111) ---- */
  float action_scale;                  /* 0.5 */
  float default_joint_pos[H1V2_NJ];    /* MJCF (leg-major) order; h12.py:38-53 */
  int32_t joint_perm[H1V2_NJ];         /* external joint i  ->  MJCF index joint_perm[i] */
  /* ---- actuator: DelayedPDActuatorCfg (h12.py:58-113), MJCF order ---- */
  float kp[H1V2_NJ], kd[H1V2_NJ], effort_limit[H1V2_NJ];
  int32_t min_delay, max_delay;        /* physics steps, 0..5 */
  /* ---- physics model (MuJoCo semantics of h12_12dof.xml) ---- */
  float gravity;                       /* 9.81 */
  float friction;                      /* tangential Coulomb coefficient of every robot-floor contact */
  float dof_damping[18], dof_armature[18], dof_frictionloss[18];
  float act_frc_limit[H1V2_NJ];        /* joint actuatorfrcrange */
  float joint_range[H1V2_NJ][2];
  float contact_solref[2];             /* (timeconst, dampratio) after geom mixing */
  float contact_solimp[5];
  float floss_solref[2], floss_solimp[5];
  float limit_solref[2], limit_solimp[5];
  int32_t solver_iterations;           /* Newton iteration cap */
  float solver_tolerance;              /* on scaled gradient norm */
  float solver_step_tolerance;         /* stop when the Newton step max|dqacc| falls below this (rad/s^2) */
  float solver_ls_tolerance;           /* exact line search stops at |phi'(a)| <= tol*|phi'(0)| (MuJoCo ls_tolerance 0.01) */
  /* ---- observations (velocity_env_cfg.py:123-142, flat_env_cfg.py:25-27) ---- */
  int32_t history_length;              /* 10 */
  int32_t enable_corruption;           /* 1 */
  float noise_ang_vel, noise_gravity, noise_joint_pos, noise_joint_vel; /* half-widths of Unoise */
  float scale_ang_vel, scale_gravity, scale_cmd, scale_joint_pos, scale_joint_vel, scale_action; /* 1 */
  /* ---- rewards ---- */
  float rew_weight[H1V2_NUM_REW];
  float track_std;                     /* 0.5 */
  float feet_air_threshold;            /* 0.4 */
  float contact_threshold;             /* 1.0 N, ContactSensor force_threshold and term thresholds */
  float soft_limit_factor;             /* 0.9 */
  float base_height_target;
  uint32_t mask_pos_limits, mask_joint_dev, mask_torques; /* bit j = MJCF joint j */
  uint32_t mask_undesired_slots;       /* bit s = sensor slot s */
  /* ---- terminations (rough_env_cfg.py:95-109) ---- */
  uint32_t mask_illegal_slots;
  /* ---- commands (velocity_env_cfg.py:90-104, flat_env_cfg.py:46-48) ---- */
  float cmd_lin_x[2], cmd_lin_y[2], cmd_ang_z[2], cmd_heading[2];
  float cmd_resample_time[2];
  float rel_standing_envs, rel_heading_envs, heading_stiffness;
  int32_t heading_command;
  /* ---- events ---- */
  float reset_pose_range[6][2];        /* x y z roll pitch yaw (rough_env_cfg.py:80-92) */
  float reset_vel_range[6][2];
  float reset_joint_pos_scale[2], reset_joint_vel_scale[2];
  float init_root_height;              /* 1.05 */
  int32_t push_enable;                 /* interval event push_by_setting_velocity (velocity_env_cfg.py:212-217) */
  float push_interval_s[2], push_vel_xy[2];
  float mass_add_range[2];             /* startup randomize_rigid_body_mass on the base (kg) */
  float friction_range[2];             /* startup per-env friction; lo==hi -> constant */
  /* ---- sharding ---- */
  int64_t env_id_offset;               /* global id of local env 0 (rank * n_envs) for the Philox key */
  float env_spacing;                   /* 2.5 m grid (velocity_env_cfg.py:288) */
  float joint_vel_limit;               /* actuator velocity_limit (rad/s, robots/h12.py:66,89,103): joint velocities are clamped
                                          to it after every physics step, as PhysX does; <= 0 disables */
  /* ---- second instances of reward terms a cfg may hold twice, and contact_forces' own parameters (rsl_env_cfg.py:343-404) ---- */
  uint32_t mask_pos_limits_b, mask_joint_dev_b; /* bit j = MJCF joint j, for H1V2_REW_DOF_POS_LIMITS_B / H1V2_REW_JOINT_DEV_B */
  uint32_t mask_contact_forces_slots;  /* bit s = sensor slot s */
  float contact_forces_threshold;      /* N; sum over the bodies of max(0, max_h |F| - threshold) */
  /* ---- command class (utils/mdp/commands.py:19-96) ---- */
  int32_t command_class;               /* 0 = UniformVelocityCommand; 1 = UniformVelocityCommandWithDeadzone */
  float velocity_deadzone;             /* class 1 keeps about half of the envs inside |cmd_xy| < dead zone: per step every env outside
                                          (inside) moves across with the probability that balances the census of the previous step
                                          (the reference moves an exact randperm-picked number, commands.py:62-83).  0 (rsl_env_cfg.py:98):
                                          nobody is ever inside, every step n_envs/2 envs in expectation get their xy command zeroed */
  float ang_vel_flip_prob;             /* per-step probability of negating the yaw-rate command (commands.py:85-96): physics_dt / episode_length_s */
  /* ---- velocity reference points of the managers (isaaclab 2.1.0 ArticulationData: poses are of the link frame, velocities of the
   *      link's centre of mass; SURVEY App. A "State conventions").  The physics state stays MuJoCo's (pelvis ORIGIN velocity). ---- */
  float root_link_com[3];              /* COM of the root LINK (pelvis alone, h12_12dof.urdf:16 == h12_12dof.xml:67) in the pelvis frame:
                                          root_lin_vel_w/_b used by rewards and command metrics = v_origin + w x (R r) */
  int32_t body_vel_at_com;             /* 1: feet_slide reads the ankle_roll_link COM velocity (body_lin_vel_w), 0: link origin */
  int32_t mass_recompute_inertia;      /* randomize_rigid_body_mass(recompute_inertia=...): 1 (isaaclab default) rescales the base inertia
                                          with the mass, 0 (cat_env_cfg.py add_base_mass) adds the mass only */
  /* ---- Constraints-as-Terminations tail (utils/cat/constraint_manager.py, cat_env.py:147-153; cat_env_cfg.py:336-431) ---- */
  int32_t cat_enable;                  /* 1: h1v2_cat_step is available (keeps the per-env diagnostics buffer) */
  float cat_tau, cat_min_p;            /* ConstraintManager(tau=0.95, min_p=0.0) */
  float cat_max_p[H1V2_NUM_CSTR];      /* maximum termination probability per term (1.0 for contact, 0.25 for the rest) */
  uint32_t cat_contact_slots;          /* contact: bit s = sensor slot s */
  float cat_foot_force_limit;          /* 750 N */
  float cat_no_move_deadzone, cat_no_move_vel_limit; /* 0.2, 6.0 rad/s */
  float cat_orientation_limit;         /* 0.1 */
  float cat_height, cat_height_std;    /* 1.0, 0.05 m */
  float cat_clearance_min_height, cat_clearance_deadzone; /* 0.1 m, 0.2 */
  float runaway_vel;                   /* an env whose root speed (m/s, rad/s) or joint speed exceeds this is treated like a
                                          non-finite one: zero reward, terminated, reset (PhysX caps at 1000, h12.py:27-28) */
  float solver_vel_tolerance;          /* second convergence test of the Newton solver, rad/s (0 = off): h * max_j |grad_j| * invweight0_j, the joint-velocity
                                          error a residual gradient leaves after the implicit update.  MuJoCo's own test (solver_tolerance, on
                                          the gradient norm scaled by meaninertia * nv = 234) lets 2.3e-3 N m pass on an ankle of 0.0136 kg m^2:
                                          8.5e-4 rad/s, the whole single-step tolerance of the north star (profiles/r2_notes.md) */
  /* ---- rough terrain (velocity_env_cfg.py:40-68, utils/mdp/terrains.py:11-28, mdp/curriculums.py:21-52); all 0 on the flat ids ---- */
  int32_t terrain_enable;              /* 1: generated height field instead of the plane; positions are relative to the env's tile origin */
  int32_t terrain_rows, terrain_cols;  /* curriculum levels x terrain types: 10 x 20 tiles */
  float terrain_tile_size;             /* 8 m (square tiles) */
  float terrain_hscale, terrain_vscale;/* 0.1 m grid, 0.005 m height unit */
  int32_t terrain_level_min, terrain_level_max, terrain_level_step; /* HfRandomUniformTerrainCfg noise_range / noise_step in height
                                          units, as upstream's int() conversions give them: 0, 4, 1 (0 .. 0.02 m in 5 mm steps) */
  int32_t terrain_border_px;           /* flat rim of every tile in grid cells: int(border_width / hscale) + 1 = 3 */
  int32_t terrain_max_init_level;      /* TerrainImporterCfg.max_init_terrain_level = 5 (< 0: rows - 1) */
  int32_t terrain_curriculum;          /* 1: mdp.terrain_levels_vel moves an env's level on every reset */
  int32_t obs_base_lin_vel;            /* 1: base_lin_vel leads the observation (velocity_env_cfg.py:123) */
  float noise_lin_vel, scale_lin_vel;  /* 0.1, 1 */
  int32_t obs_height_scan;             /* 1: height_scan ends the observation (velocity_env_cfg.py:133-138): GridPattern rays about the
                                          scanner body (torso_link == pelvis frame, fixed joint), yaw-aligned */
  float scan_size[2], scan_resolution; /* 1.6 x 1.0 m at 0.1 m -> 17 x 11 = 187 rays, x fastest */
  float scan_offset;                   /* mdp.height_scan offset 0.5: value = sensor z - hit z - offset */
  float noise_height_scan, scale_height_scan, scan_clip[2]; /* 0.1, 1, (-1, 1): noise, then clip, then scale (ObservationManager order) */
  int32_t reserved[7];                 /* [0] != 0: keep per-env diagnostics of the last step (get_state's read-only fields);
                                          [1] > 0: line-search evaluations per Newton iteration (default 6);
                                          [2] in {1,2,4,8,16}: envs per warp (default: chosen from n_envs);
                                          [3] != 0: 8-env warps run the plain instantiation instead of the mirror-lane one; rest 0 */
} H1v2Config;

/* Natural-layout state exchange (row-major, env-major).  NULL members are skipped.
 * Joint-indexed arrays are in MJCF order.  Device pointers for the product, host pointers for the oracle. */
typedef struct H1v2State {
  float* root_pos;      /* [N,3] relative to the env origin */
  float* root_quat;     /* [N,4] w,x,y,z */
  float* root_lin_vel;  /* [N,3] world frame, of the pelvis origin */
  float* root_ang_vel;  /* [N,3] body frame */
  float* joint_pos;     /* [N,12] */
  float* joint_vel;     /* [N,12] */
  float* last_action;   /* [N,12] external order, as passed to step */
  float* target_hist;   /* [N,2,12] processed joint targets of the previous two control steps */
  int32_t* lag;         /* [N] actuator delay in physics steps */
  int32_t* fresh;       /* [N] 1 = no step since reset (delay line and obs history will back-fill) */
  float* command;       /* [N,3] */
  float* heading_target;/* [N] */
  float* time_left;     /* [N] */
  int32_t* is_standing; /* [N] */
  int32_t* is_heading;  /* [N] */
  float* cmd_metrics;   /* [N,2] */
  float* feet_timers;   /* [N,2,4] cur_air,last_air,cur_contact,last_contact */
  float* episode_sums;  /* [N,NUM_REW] */
  float* obs_history;   /* [N,H,45] oldest -> newest */
  float* friction;      /* [N] */
  float* mass_add;      /* [N] */
  float* push_time_left;/* [N] */
  /* read-only diagnostics of the last step (get only) */
  float* slot_force;    /* [N,NUM_SLOT,3] net contact force of the last substep, world */
  float* slot_force_hist; /* [N,NUM_SLOT,3] |F| of the last three substeps, oldest -> newest */
  float* applied_torque;/* [N,12] last substep, after the effort-limit clip */
  float* joint_acc;     /* [N,12] last substep */
  float* reward_terms;  /* [N,NUM_REW] weighted*dt value of every term in the last step */
  float* foot_vel;      /* [N,2,3] world linear velocity of the ankle_roll_link origins */
  float* solver_iters;  /* [N,3] max and sum of Newton iterations over the substeps of the last step; contact points dropped because a leg's
                           active list (5 entries) was full -- only possible with shin / torso / pelvis on the ground, i.e. in a terminating step */
  float* pre_reset_qpos;   /* [N,19] state after the physics substeps of the last step, BEFORE any reset */
  float* pre_reset_qvel;   /* [N,18] */
  float* pre_reset_timers; /* [N,2,4] */
  /* rough terrain (terrain_enable): the env's curriculum level (row of the tile grid; get and set) and its terrain type (column; get only) */
  int32_t* terrain_level;  /* [N] */
  int32_t* terrain_type;   /* [N] */
} H1v2State;

typedef struct H1v2Handle H1v2Handle;

int h1v2_default_config(H1v2Config* cfg); /* resolved cfg of Isaac-Velocity-Flat-H12_12dof-v0 */
int h1v2_rsl_config(H1v2Config* cfg);     /* resolved cfg of Isaac-Velocity-Rsl-H12_12dof-v0 (config/h12_12dof/rsl_env_cfg.py:503-540) */
int h1v2_cat_config(H1v2Config* cfg);     /* resolved cfg of Isaac-Velocity-CaT-Flat-H12_12dof-v0 (config/h12_12dof/cat_env_cfg.py:528-565) */
int h1v2_rough_config(H1v2Config* cfg);   /* resolved cfg of Isaac-Velocity-Rough-H12_12dof-v0 (config/h12_12dof/rough_env_cfg.py:65-125) with the
                                             reference's in-tree terrain generator cfg (utils/mdp/terrains.py:11-28) */
int h1v2_create(const H1v2Config* cfg, int32_t n_envs, int32_t device, uint64_t seed, H1v2Handle** out);
void h1v2_destroy(H1v2Handle* h);
const char* h1v2_last_error(void);
int h1v2_obs_dim(const H1v2Handle* h);
int h1v2_num_envs(const H1v2Handle* h);

/* episode_length_buf: int64[N] device buffer owned by the caller, read and written by every step.  Binding copies the current
 * counters into it; NULL un-binds (the counters move back into the handle).  Synchronises the device (rare call). */
int h1v2_bind_episode_length(H1v2Handle* h, int64_t* episode_length);
/* Debug aid: every device array of the handle lies between two 256-byte guard zones; returns how many guard bytes have been
 * overwritten (0 = no out-of-bounds store so far), -1 on error.  Synchronises the device. */
int64_t h1v2_check_guards(H1v2Handle* h);

/* Reset envs.  env_ids == NULL resets all.  No observation is produced (use h1v2_observe). */
int h1v2_reset(H1v2Handle* h, const int64_t* env_ids, int32_t n, void* cuda_stream);
/* Compute observations for the current state WITHOUT stepping (ManagerBasedRLEnv.reset's obs);
 * appends to the history exactly like observation_manager.compute(). */
int h1v2_observe(H1v2Handle* h, float* obs, void* cuda_stream);

/* One control step (decimation physics substeps + managers).  All pointers are device pointers. */
int h1v2_step(H1v2Handle* h, const float* actions /*[N,12]*/, float* obs /*[N,obs_dim]*/, float* rew /*[N]*/,
              uint8_t* terminated /*[N]*/, uint8_t* truncated /*[N]*/, void* cuda_stream);
/* Same with HOST buffers (pinned or pageable): H2D of actions and D2H of all outputs inside, synchronises. */
int h1v2_step_host(H1v2Handle* h, const float* actions, float* obs, float* rew, uint8_t* terminated,
                   uint8_t* truncated);
/* How h1v2_step_host moves the observations (decided at its first calls; H1V2_HOST_PATH=rows|assemble|hybrid, H1V2_HOST_ROWS_FRAC=<0..1> and
 * H1V2_HOST_THREADS=<n> override): mode 0 "rows" = the kernel writes whole [N, obs_dim] rows into the caller's buffer (zero-copy over PCIe when it
 * is pinned, staged otherwise); mode 1 "assemble" = only the step's new 45-float sample of an env crosses PCIe and `threads` host threads
 * assemble its row from a host mirror of the history ring while the kernel runs -- for ALL envs, or ("hybrid") for the envs from
 * h1v2_host_path_rows() upwards while the kernel writes the rows of the envs below it, so that PCIe DMA and the cores' memory bandwidth work
 * concurrently.  Bit-identical rows in every case (tests/test_gpu_parity.py::test_step_host_matches_device_path).  -1 = not decided yet.
 * Without an override a handle with a pinned observation buffer and at least four host threads times five candidates (assemble, rows, hybrid
 * at 1/4, 1/2, 3/4) over its first forty calls and keeps the fastest.  With pinned buffers the kernel raises one flag per warp in mapped host
 * memory once that warp's outputs are in the caller's buffers: the host threads take the envs over warp by warp while the rest of the grid runs.
 * The call orders itself after work queued earlier through the stream-taking entry points and synchronises before returning. */
int h1v2_host_path_info(const H1v2Handle* h, int32_t* mode, int32_t* threads);
/* envs whose observation rows the kernel itself writes into the caller's buffer: N in mode 0, 0 .. N in mode 1 (> 0: hybrid), -1 undecided */
int h1v2_host_path_rows(const H1v2Handle* h);

/* Replace the reward weights (CurriculumManager's modify_reward_weight, rsl_env_cfg.py:447-497).  Host array of
 * H1V2_NUM_REW floats; takes effect from the next step enqueued after the call.  Never synchronises.  (h1v2_step is CUDA-graph
 * capturable; a captured step keeps the parameter block of capture time, so re-capture after changing the weights.) */
int h1v2_set_reward_weights(H1v2Handle* h, const float* weights);

/* Constraints-as-Terminations step (CaTEnv.step, utils/cat/cat_env.py:95-193): h1v2_step followed by the constraint tail --
 * rew is scaled by 1 - p, dones[N] (float) = p, and 1 for envs that reset; truncated as in h1v2_step.  Needs cfg.cat_enable.
 * Two launches: the fused step kernel also computes the raw constraint columns, their maxima over all envs and the ordered
 * dead-zone list; one apply kernel turns them into probabilities (the maxima of ALL envs must be known first).  Every piece of
 * step-to-step state (running maxima, their parity, the first-step flag) lives on the device: CUDA-graph capturable like
 * h1v2_step (a captured step keeps the max_p of capture time). */
int h1v2_cat_step(H1v2Handle* h, const float* actions, float* obs, float* rew, float* dones, uint8_t* truncated, void* cuda_stream);
/* Same with HOST buffers (the CaT flavour of h1v2_step_host: same two observation paths, synchronises). */
int h1v2_cat_step_host(H1v2Handle* h, const float* actions, float* obs, float* rew, float* dones, uint8_t* truncated);
/* curriculums.modify_constraint_p (utils/cat/curriculums.py:20-42): new maximum probabilities, host array [H1V2_NUM_CSTR] */
int h1v2_set_constraint_max_p(H1v2Handle* h, const float* max_p);
/* host copies (synchronises): raw constraint columns and probabilities of the last cat step [H1V2_CSTR_COLS][N], the running
 * column maxima [H1V2_CSTR_COLS]; NULL members are skipped */
int h1v2_cat_debug(H1v2Handle* h, float* raw, float* probs, float* running_max);
/* Episode_Constraint_violation/<term> (percent) and Episode_Constraint_probability/<term> of the envs reset in the last cat step that
 * reset any (constraint_manager.py:185-203): out[0..9] violation, out[10..19] probability, out[20] number of such envs.  Synchronises. */
int h1v2_get_cat_log_host(H1v2Handle* h, float* out /*[2 * H1V2_NUM_CSTR + 1]*/);
/* device pointer to the SUMS behind that log (float[2 * H1V2_NUM_CSTR + 1]: violation sums x100, probability sums, count of the
 * envs reset in the last cat step), for callers that must not synchronise */
int h1v2_get_cat_log(H1v2Handle* h, const float** acc_dev);

/* Rough terrain (cfg.terrain_enable).  The height field is a grid of [rows * px + 1][cols * px + 1] vertices (px = tile_size / hscale,
 * second index fastest, metres) triangulated like isaaclab's convert_height_field_to_mesh (cell diagonal from (i, j) to (i+1, j+1));
 * vertex (i, j) sits at x = i * hscale - rows * tile / 2, y = j * hscale - cols * tile / 2; outside the grid the ground is flat at 0
 * (the 20 m border).  h1v2_create generates it from the seed (uniform levels on the interior vertices of every tile, flat tile rims:
 * HfRandomUniformTerrainCfg); h1v2_set_terrain replaces it with the caller's grid, e.g. the one a real isaaclab TerrainGenerator made.
 * Both take / fill HOST arrays and synchronise.  dims = {grid_x, grid_y}. */
int h1v2_terrain_dims(const H1v2Handle* h, int32_t dims[2]);
int h1v2_get_terrain(H1v2Handle* h, float* heights);
int h1v2_set_terrain(H1v2Handle* h, const float* heights);
/* device pointer to float[2]: [1] = Curriculum/terrain_levels, the mean level over all envs after the last step (curriculums.py:52) */
int h1v2_get_terrain_log(H1v2Handle* h, const float** log_dev);
int h1v2_get_state(H1v2Handle* h, const H1v2State* dst, void* cuda_stream);
int h1v2_set_state(H1v2Handle* h, const H1v2State* src, void* cuda_stream);
/* device pointer to the float[H1V2_LOG_DIM] log vector (valid for the handle's lifetime, updated by step) */
int h1v2_get_log(H1v2Handle* h, const float** log_dev);
int h1v2_get_log_host(H1v2Handle* h, float* log_host /*[H1V2_LOG_DIM]*/);
/* Envs per warp the step launch of an n_envs handle uses on a GPU with `sms` multiprocessors (host-only, no device needed): 4 (four mirror lanes per
 * lane, up to one warp per scheduler), 8 (two mirrors, up to ~4.75 warps per SM), 16 above; plain != 0: the rule of the plain instantiation
 * (cfg.reserved[3]).  cfg.reserved[2] overrides it per handle.  DESIGN.md section 3. */
int h1v2_envs_per_warp(int n_envs, int sms, int plain);
/* cumulative histogram of Newton iterations per (env,substep) solve since creation: hist32[k] = #solves with k iterations */
int h1v2_debug_iter_hist(H1v2Handle* h, float* hist32);
/* number of kernels launched by this handle since creation (for bench.py's gpu_launches) */
int64_t h1v2_launch_count(const H1v2Handle* h);
/* FP32 FMA throughput of the device (TFLOP/s, best of 6): the measured denominator of the FP32 roofline */
int h1v2_measure_fp32_peak(int32_t device, float* tflops);
/* fill actions[N,12] with N(0,1) draws: Philox(seed, stream 7), counter = step */
int h1v2_random_actions(H1v2Handle* h, float* actions, uint64_t step, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
