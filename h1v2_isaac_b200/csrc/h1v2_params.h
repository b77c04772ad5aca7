// h1v2_params.h -- kernel-side parameter block (passed as a __grid_constant__ kernel argument) and the
// internal HBM layout of the per-env state.  Host code (h1v2_capi.cu) fills both.
#pragma once
#include <stdint.h>

#include "../../include/h1v2_b200.h"

#define H1V2_OBS_MAXPASS ((H1V2_MAX_HISTORY * H1V2_OBS_TERM_DIM + 31) / 32)  // 32-column passes over one observation row
#define H1V2_EPSUM_F4 ((H1V2_NUM_REW + 3) / 4)  // float4 rows of the episode sums
#define H1V2_HIST_STRIDE 48  // floats per history slot (45 used; 192 B = 6 sectors)

// --- rigid-body constants of one leg (MJCF order: hip_yaw, hip_pitch, hip_roll, knee, ankle_pitch, ankle_roll) ---
struct KLeg {
  float pos[6][3];      // body offset in the parent frame        (h12_12dof.xml:71,76,81,86,91,96 / :107..)
  float ipos[6][3];     // COM in the body frame
  float inertia[6][6];  // xx yy zz xy xz yz about the COM, body frame
  float mass[6];
  float foot_pt[4][3];  // sole corners in the ankle_roll frame
  float shin_pt[2][3];  // shin capsule end centres in the knee_link frame
  float shin_rad;
};

struct KParams {
  KLeg leg[2];
  float root_mass, root_ipos[3], root_inertia[6];
  float root_pt[9][3];  // torso box corners 0..7, pelvis sphere centre 8
  float root_rad[9];
  // timing
  float h;              // physics dt
  int decimation;
  float step_dt;
  int64_t max_episode_length;
  float max_episode_length_s;
  // action / actuator
  float action_scale;
  float q0[12], kp[12], kd[12], effort[12], frc[12];
  int perm[12];      // external -> MJCF
  int inv_perm[12];  // MJCF -> external
  int min_delay, max_delay;
  // physics
  float gravity;
  int mass_scales_inertia;       // added base mass rescales the base inertia (randomize_rigid_body_mass recompute_inertia)
  float vel_limit, runaway_vel;  // joint velocity clamp (<= 0: off); runaway-state guard
  float damping[18], armature[18];
  float floss[18], floss_D[18], floss_lim[18], floss_B;  // lim = R*frictionloss
  float range_lo[12], range_hi[12], limit_invw[12], limit_K, limit_B, limit_imp[5];
  float contact_K, contact_B, contact_imp[5];
  float slot_tran[6];
  int max_iters;
  float tol;  // on |grad| * scale
  float step_tol;
  float vel_tol;  // second convergence test: h * max_j |grad_j| * invweight0_j below this (<= 0: off, stored as a huge number)
  int ls_max;    // line-search evaluations per Newton iteration (default 6, then the secant root of the bracket: measured on B200, iteration histogram and parity identical to 12; at <= 4 rare solves stall)
  float ls_tol;  // line-search tolerance on |phi'(alpha)| / |phi'(0)|  (MuJoCo opt.ls_tolerance = 0.01)
  float grad_scale;
  // observations
  int H, obs_dim, corrupt;
  float n_av, n_g, n_q, n_v;
  float s_av, s_g, s_cmd, s_q, s_v, s_a;
  // rewards
  float w[H1V2_NUM_REW];
  float inv_std2, air_thr, contact_thr, base_h;
  float root_com[3];  // COM of the root link in the pelvis frame: the managers read root_lin_vel at it (isaaclab ArticulationData)
  int foot_vel_com;   // feet_slide reads the ankle_roll_link COM velocity
  float soft_lo[12], soft_hi[12];
  uint32_t m_poslim, m_dev, m_tau, m_undesired, m_illegal, m_poslim_b, m_dev_b, m_cforce;
  float cforce_thr;
  // commands
  float c_lx[2], c_ly[2], c_wz[2], c_hd[2], c_rt[2];
  float rel_standing, rel_heading, k_heading;
  int heading_cmd;
  int cmd_class;          // 1: UniformVelocityCommandWithDeadzone (T/utils/mdp/commands.py:19-96), dead zone 0
  float deadzone, flip_prob;
  float max_command_step;
  // events
  float rp[6][2], rv[6][2], rjp[2], rjv[2], init_h;
  int push_enable;
  float push_int[2], push_v[2];
  // rng
  uint32_t key0;
  int64_t env_id_offset;
  int n;
  // Rough id (V/velocity_env_cfg.py:36-142, 270-276): observation layout and the height field
  int lut_dim;   // entries of the flatten table: obs_dim on the flat ids; 48 on the Rough id (base_lin_vel | the 45 regular terms)
  int rough;     // 1: the rough instantiation of the step kernel runs (terrain contacts, base_lin_vel, height scan)
  int lin_vel;   // base_lin_vel leads the row; its three values travel in floats 45..47 of the sample slot
  float n_lv, s_lv;
  int scan_nx, scan_ny, scan_col0;  // height-scan rays along x / y (x fastest), first column of the scan in the row (0 rays: no scan)
  float scan_res, scan_x0, scan_y0, scan_off, n_scan, s_scan, scan_lo, scan_hi;
  int t_rows, t_cols, t_npx, t_gx, t_gy, t_curriculum;  // tiles, grid cells per tile side, grid vertices, terrain_levels_vel on / off
  float t_inv_hs, t_half, t_tile;   // 1 / horizontal scale, half a tile, a tile
  int epw;  // envs per warp (1,2,4,8,16): lanes 2*epw..31 shadow the warp's first env (small-N mapping)
};

// --- Constraints-as-Terminations: what the step kernel leaves for the apply kernel (csrc/h1v2_cat.cuh).  raw == NULL: off ---
struct KCat {
  float* raw;      // [56][N] raw constraint columns of this step (the 12 no_move columns are filled in by the apply kernel's gather)
  float* qd;       // [12][N] pre-reset joint velocities, MJCF order: what no_move gathers from another env
  float* aux;      // [2][N]  pre-reset episode length | reset flag
  unsigned short* dz;  // [warps of the step launch] bit p: env warp*epw + p has its whole command inside the no_move dead zone
  int* cmax;       // [56]    this step's column maxima as float bits (candidates are positive: floor 1e-6)
  int* list;       // [warps] inclusive prefix of the dead-zone member counts per warp mask (the step kernel's last block); ctl[0] = total K
  int* ctl;        // [0] list length K  [1] apply launches done (running-maximum parity, first-step flag)  [2] apply ticket
  float* swing;    // [2][N]  swing_max_height of foot_clearance
  float* logacc;   // [21]    sums over the envs reset in this step: violation[10], probability[10], count
  uint32_t contact_slots;
  float foot_force_limit, no_move_deadzone, no_move_vel_limit, orientation_limit, height, height_std, clearance_min_height, clearance_deadzone, vel_limit;
};

// --- internal state, SoA of float4 so that every lane issues coalesced 128-bit accesses ---
//   lane index l = 2*env + side (side 0 = left leg, 1 = right leg)
struct KState {
  float4* root;    // [4][N]   (px,py,pz,qw) (qx,qy,qz,vx) (vy,vz,wx,wy) (wz,friction,mass_add,push_left)
  float4* leg;     // [3][2N]  (q0..q3) (q4,q5,qd0,qd1) (qd2..qd5)
  float4* act;     // [5][2N]  last_action[6] | T1[6] | T2[6] | pad2   (MJCF order within the leg)
  float4* cmd;     // [2][N]   (cx,cy,cz,heading_target) (time_left,err_xy,err_yaw,flags)
  float4* timers;  // [2N]     cur_air,last_air,cur_contact,last_contact of the lane's foot
  float4* warm;    // [3][2N]  solver warm start: qacc of the last substep (leg 6 | root 6)
  float4* epsum;   // [6][N]   episode sums of the 22 reward slots (2 pad)
  float* hist;     // [N][H][48]
  const int* lut;  // [obs_dim] (history index << 8) | offset in the 45-float sample, for the term-major flatten
  int64_t* ep_len; // [N] bound, owned by the caller
  float* diag;     // [N][H1V2_DIAG_DIM] or NULL
  KCat cat;        // Constraints-as-Terminations outputs (raw == NULL unless launched by h1v2_cat_step)
  const float* terrain_h;   // [t_gx][t_gy] height field in metres (Rough id), second index fastest
  const float* terrain_oz;  // [t_rows][t_cols] height of every tile's origin
  float* tlog;              // [2] sum of the envs' terrain levels of the step in flight | mean level after the last step
  int n_rows;         // host path (sample_out != NULL): envs [0, n_rows) get their observation ROWS written (into `obs`), the rest only their sample
  unsigned* host_flags;  // host path: [blocks of the launch] mapped host words; a warp stores host_seq there once its outputs are in the caller's memory
  unsigned host_seq;
  float* sample_out;  // [N][48] or NULL: this step's observation sample (45) | fresh flag at [45]; the host path of h1v2_step_host assembles the rows from it
  float* acc;      // [H1V2_LOG_DIM] log accumulators (atomics)
  float* log;      // [H1V2_LOG_DIM] published log vector
  unsigned long long* counters;  // [0] global step counter, [1] history head, [2] dead-zone census being taken, [3] census of the last step
  unsigned* done;                // blocks of the current launch that have finished (the last one publishes the log, advances the counters)
};
#define H1V2_DIAG_DIM 168
#define H1V2_DIAG_REW0 144
// diag layout: slot_force 0..17 | slot_hist 18..35 | applied_torque 36..47 | joint_acc 48..59 | (free 60..79)
//              | foot_vel 80..85 | newton iters 86 | cap hit 87 | iter sum 88 | contact-list overflow 89 | pre-reset qpos 96..114 | qvel 115..132 | timers 133..140
//              | command before this step's update 141..143 | reward_terms 144..165 | episode length (pre-reset) 166 | reset flag 167
#define FLAG_DELAY_FRESH 1
#define FLAG_HIST_FRESH 2
#define FLAG_LAG_SHIFT 2
#define FLAG_STANDING 32
#define FLAG_HEADING 64
// Rough id: the env's tile in the terrain grid, kept in the flag word (level = row, moved by the curriculum; type = column, fixed)
#define FLAG_LEVEL_SHIFT 8
#define FLAG_TYPE_SHIFT 16
#define FLAG_TERRAIN_MASK 0x00ffff00
