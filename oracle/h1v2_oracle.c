/*
 * h1v2_oracle.c -- CPU restatement (float64) of the hot path of olivier-stasse/h1v2-Isaac:
 * one ManagerBasedRLEnv.step of Isaac-Velocity-Flat-H12_12dof-v0 (and its Rsl / Rough siblings) with MuJoCo-semantics physics.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (h1v2_isaac_b200/) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" against the third-party engines that hold the arithmetic --
 * isaaclab 2.1.0, PhysX (isaacsim 4.5.0) and mujoco 3.3.6 are absent from /root/reference and not
 * installable here.  What IS pinned (tests/test_oracle_kat.py, tests/golden/):
 *   total mass 67.3675873 kg (D/utils/mj_logger.py:64), keyframe FK (h12_12dof.xml:362-368),
 *   projected_gravity closed form (D/controllers/rl.py:86-95), history fill/ordering
 *   (T/utils/history/circular_buffer.py:79-137, D/controllers/rl.py:60-81), feet_air_time* bodies
 *   (V/mdp/rewards.py:13-62, imported and run to make tests/golden/), soft limits (A/robots/h12.py:56),
 *   Philox4x32-10 known answers (Random123).
 *
 * Reference material each function follows (paths relative to the reference root; T/ V/ C12/ A/ D/ as
 * in SURVEY.md):
 *   step order ................ T/utils/cat/cat_env.py:95-193 (vendored ManagerBasedRLEnv.step), :195-248
 *   action / PD / decimation .. V/velocity_env_cfg.py:111 ; A/robots/h12.py:58-113 ;
 *                               D/robots/h12_mujoco.py:55-67
 *   physics ................... MuJoCo 3.3.6 semantics (published algorithm: CRBA/RNE, soft constraints,
 *                               pyramidal cones, implicitfast) as set up by D/simulator/sim_mujoco.py:39-41
 *                               on A/models/h12/scene/h12_12dof.xml ; SURVEY.md Appendix D
 *   contact sensor ............ V/velocity_env_cfg.py:69,314-315 ; SURVEY.md Appendix B
 *   observations .............. T/utils/history/observation_manager.py:318-355 ; V/velocity_env_cfg.py:123-142
 *   rewards ................... C12/rough_env_cfg.py:18-62 ; V/mdp/rewards.py ; SURVEY.md Appendix B
 *   terminations .............. C12/rough_env_cfg.py:95-109 ; T/utils/cat/constraints.py:86-99 (idiom)
 *   commands .................. V/velocity_env_cfg.py:90-104 ; T/utils/mdp/commands.py:47-59
 *   reset events .............. C12/rough_env_cfg.py:78-92 ; V/velocity_env_cfg.py:176-209
 *   rough terrain (Rough id) .. V/velocity_env_cfg.py:40-68,119-142,270-276 ; T/utils/mdp/terrains.py:11-28 (generator cfg) ;
 *                               V/mdp/curriculums.py:21-52 (terrain_levels_vel; pinned by tests/golden/terrain_curriculum.npz, produced by
 *                               the reference's own function) ; upstream isaaclab 2.1.0 terrain generator / height-field mesher /
 *                               RayCaster / mdp.height_scan [from memory: SURVEY App. B]; the physics of the rough scene has no
 *                               external target at all (the reference's MuJoCo model is flat)
 *
 * Deliberately written with dense, generic algorithms (13-body tree loops, dense 18x18 Cholesky, explicit
 * constraint-row lists) so that it shares no structure with the CUDA kernel it checks.
 */
#include "h1v2_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/h1v2_model_h12.h"

#define NB 13
#define NV 18
#define NJ 12
#define NSLOT 6
#define NREW H1V2_NUM_REW
#define MAXROW (18 + 24 + 4 * H1V2_NCOLL)
#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999
#define PI_F 3.14159265358979323846f

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011).  key=(seed mix, global env id), ctr=(step lo, step hi, stream, block) */
static void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void h1v2o_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }

static void rng4(uint64_t seed, int64_t env_gid, uint64_t step, uint32_t stream, uint32_t block, float u[4]) {
  uint32_t key[2] = {(uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u), (uint32_t)env_gid};
  uint32_t ctr[4] = {(uint32_t)step, (uint32_t)(step >> 32), stream, block};
  uint32_t o[4];
  philox4x32_10(ctr, key, o);
  for (int i = 0; i < 4; i++) u[i] = (float)(o[i] >> 8) * (1.0f / 16777216.0f);
}
void h1v2o_rng4(uint64_t seed, int64_t env_gid, uint64_t step, uint32_t stream, uint32_t block, float u[4]) {
  rng4(seed, env_gid, step, stream, block, u);
}
#define STREAM_OBS 0u
#define STREAM_RESET 1u
#define STREAM_CMD 2u
#define STREAM_EVENT 3u
#define STREAM_TERRAIN 5u
#define STREAM_ACTIONS 7u

static inline float uni(float u, float lo, float hi) { return fmaf(hi - lo, u, lo); } /* fused, like the kernel: draws are bit-identical */

/* ------------------------------------------------------------------------------------------ */
typedef struct {
  double qpos[19], qvel[18];
  float last_action[NJ];
  double T1[NJ], T2[NJ];
  int lag, fresh;
  float cmd[3], heading_target, time_left, metrics[2];
  int is_standing, is_heading;
  float timers[2][4];
  double ep_sums[NREW];
  float hist[H1V2_MAX_HISTORY][H1V2_OBS_TERM_DIM];
  int64_t ep_len;
  double friction, mass_add;
  float push_left;
  int level, type; /* rough terrain: row (curriculum level) and column (terrain type) of the env's tile */
  /* diagnostics */
  double slot_force[NSLOT][3], slot_hist[NSLOT][3], applied_tau[NJ], joint_acc[NJ], rew_terms[NREW], foot_vel[2][3];
  double slot_hist_diag[NSLOT][3]; /* slot_hist as the terminations / rewards of the last step saw it (before any reset cleared it) */
  int newton_iters;
  double newton_resid;
  double min_abs_dist; /* smallest |signed distance| of any contact candidate at a substep start during the last step */
  double min_limit_dist; /* same for joint-limit activation */
  double min_tri_margin; /* rough terrain: distance (cells) of a contact candidate near the ground to a triangle boundary of the height field */
} OEnv;

struct H1v2Oracle {
  H1v2Config cfg;
  int n;
  uint64_t seed;
  uint64_t step_counter;
  OEnv* env;
  float log[H1V2_LOG_DIM];
  int64_t max_episode_length;
  double soft_lo[NJ], soft_hi[NJ];
  int max_newton_iters;
  int nthreads;
  int dz_prev; /* envs inside the command dead zone at the end of the previous step (UniformVelocityCommandWithDeadzone balancing) */
  /* rough terrain (cfg.terrain_enable): height grid [gx][gy] in metres, tile-origin heights [rows][cols] */
  float* terrain;
  float* origin_z;
  int gx, gy, npx;
  float terrain_level_mean;
};

/* ------------------------------------------------------------------------------------------ */
/* small linear algebra */
static void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void quat2mat(const double q[4], double R[9]) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
}
static void matvec3(const double R[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; i++) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}
static void mattvec3(const double R[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; i++) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
static void matmul3(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void axis_angle_mat(const double a[3], double th, double R[9]) {
  double c = cos(th), s = sin(th), v = 1 - c;
  R[0] = c + a[0] * a[0] * v;        R[1] = a[0] * a[1] * v - a[2] * s; R[2] = a[0] * a[2] * v + a[1] * s;
  R[3] = a[1] * a[0] * v + a[2] * s; R[4] = c + a[1] * a[1] * v;        R[5] = a[1] * a[2] * v - a[0] * s;
  R[6] = a[2] * a[0] * v - a[1] * s; R[7] = a[2] * a[1] * v + a[0] * s; R[8] = c + a[2] * a[2] * v;
}
/* dense Cholesky A = L L^T in place (lower); returns 0 on success */
static int chol(double* A, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    if (d <= 0) return -1;
    d = sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return 0;
}
static void chol_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * b[k];
    b[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * b[k];
    b[i] = s / L[i * n + i];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* kinematics and dynamics, world-aligned spatial vectors [ang; lin] about the env-local origin */
typedef struct {
  double R[NB][9], x[NB][3];
  double S[NV][6];
  int dof_body[NV];
} Kin;

static const int body_of_dof[NV] = {0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12};

static int is_ancestor_or_self(int a, int b) { /* is body a an ancestor of (or equal to) body b */
  while (b >= 0) {
    if (a == b) return 1;
    b = h1v2_body_parent[b];
  }
  return 0;
}

static void kinematics(const double* qpos, Kin* k) {
  double q[4] = {qpos[3], qpos[4], qpos[5], qpos[6]};
  double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= nq;
  quat2mat(q, k->R[0]);
  for (int i = 0; i < 3; i++) k->x[0][i] = qpos[i];
  for (int b = 1; b < NB; b++) {
    int p = h1v2_body_parent[b];
    double Rj[9], off[3];
    axis_angle_mat(h1v2_jnt_axis[b - 1], qpos[7 + b - 1], Rj);
    matmul3(k->R[p], Rj, k->R[b]);
    matvec3(k->R[p], h1v2_body_pos[b], off);
    for (int i = 0; i < 3; i++) k->x[b][i] = k->x[p][i] + off[i];
  }
  memset(k->S, 0, sizeof(k->S));
  for (int d = 0; d < 3; d++) k->S[d][3 + d] = 1.0;
  for (int d = 0; d < 3; d++) {
    double w[3] = {k->R[0][d], k->R[0][3 + d], k->R[0][6 + d]}, u[3];
    cross3(k->x[0], w, u);
    for (int i = 0; i < 3; i++) { k->S[3 + d][i] = w[i]; k->S[3 + d][3 + i] = u[i]; }
  }
  for (int b = 1; b < NB; b++) {
    double w[3], u[3];
    matvec3(k->R[b], h1v2_jnt_axis[b - 1], w);
    cross3(k->x[b], w, u);
    for (int i = 0; i < 3; i++) { k->S[5 + b][i] = w[i]; k->S[5 + b][3 + i] = u[i]; }
  }
}

/* 6x6 spatial inertia of body b about the origin, world axes */
static void body_inertia6(const Kin* k, int b, double mass_add, int recompute_inertia, double I6[36]) {
  double m = h1v2_body_mass[b], scale = 1.0;
  /* randomize_rigid_body_mass(operation="add"): the inertia tensor is rescaled with the mass unless recompute_inertia=False
   * (C12/cat_env_cfg.py add_base_mass) */
  if (b == 0 && mass_add != 0.0) { scale = recompute_inertia ? (m + mass_add) / m : 1.0; m += mass_add; }
  double c[3], off[3];
  matvec3(k->R[b], h1v2_body_ipos[b], off);
  for (int i = 0; i < 3; i++) c[i] = k->x[b][i] + off[i];
  double T[9], Iw[9], Rt[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Rt[3 * i + j] = k->R[b][3 * j + i];
  matmul3(k->R[b], h1v2_body_inertia[b], T);
  matmul3(T, Rt, Iw);
  double cc = dot3(c, c);
  memset(I6, 0, 36 * sizeof(double));
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) I6[6 * i + j] = scale * Iw[3 * i + j] + m * ((i == j ? cc : 0.0) - c[i] * c[j]);
  /* m*[c]x in the upper-right, transpose in the lower-left */
  double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { I6[6 * i + 3 + j] = m * cx[3 * i + j]; I6[6 * (3 + i) + j] = m * cx[3 * j + i]; }
  for (int i = 0; i < 3; i++) I6[6 * (3 + i) + 3 + i] = m;
}

static void mass_matrix(const H1v2Config* cfg, const Kin* k, double mass_add, double M[NV * NV]) {
  double Ic[NB][36];
  for (int b = 0; b < NB; b++) body_inertia6(k, b, mass_add, cfg->mass_recompute_inertia, Ic[b]);
  for (int b = NB - 1; b >= 1; b--) {
    int p = h1v2_body_parent[b];
    for (int i = 0; i < 36; i++) Ic[p][i] += Ic[b][i];
  }
  memset(M, 0, NV * NV * sizeof(double));
  for (int j = 0; j < NV; j++) {
    double F[6];
    const double* I6 = Ic[body_of_dof[j]];
    for (int r = 0; r < 6; r++) {
      F[r] = 0;
      for (int c = 0; c < 6; c++) F[r] += I6[6 * r + c] * k->S[j][c];
    }
    for (int i = 0; i <= j; i++) {
      if (!is_ancestor_or_self(body_of_dof[i], body_of_dof[j])) continue;
      double s = 0;
      for (int r = 0; r < 6; r++) s += k->S[i][r] * F[r];
      M[i * NV + j] = M[j * NV + i] = s;
    }
  }
  for (int d = 0; d < NV; d++) M[d * NV + d] += cfg->dof_armature[d];
}

static void cross_motion(const double v[6], const double s[6], double o[6]) {
  double a[3], b[3], c[3];
  cross3(v, s, a);
  cross3(v, s + 3, b);
  cross3(v + 3, s, c);
  for (int i = 0; i < 3; i++) { o[i] = a[i]; o[3 + i] = b[i] + c[i]; }
}
static void cross_force(const double v[6], const double f[6], double o[6]) {
  double a[3], b[3], c[3];
  cross3(v, f, a);
  cross3(v + 3, f + 3, b);
  cross3(v, f + 3, c);
  for (int i = 0; i < 3; i++) { o[i] = a[i] + b[i]; o[3 + i] = c[i]; }
}

/* bias = C(q,v) + g(q): recursive Newton-Euler with zero joint acceleration (MuJoCo mj_rne, flg_acc=0) */
static void rne_bias(const H1v2Config* cfg, const Kin* k, const double* qvel, double mass_add, double bias[NV],
                     double cvel_out[NB][6]) {
  double cvel[NB][6], cacc[NB][6], f[NB][6];
  memset(cvel, 0, sizeof(cvel));
  memset(cacc, 0, sizeof(cacc));
  cacc[0][5] = cfg->gravity; /* -gravity vector */
  for (int d = 0; d < 3; d++)
    for (int i = 0; i < 6; i++) cvel[0][i] += k->S[d][i] * qvel[d];
  {
    double add_v[6] = {0}, add_a[6] = {0};
    for (int d = 3; d < 6; d++) {
      double sd[6];
      cross_motion(cvel[0], k->S[d], sd);
      for (int i = 0; i < 6; i++) { add_a[i] += sd[i] * qvel[d]; add_v[i] += k->S[d][i] * qvel[d]; }
    }
    for (int i = 0; i < 6; i++) { cacc[0][i] += add_a[i]; cvel[0][i] += add_v[i]; }
  }
  for (int b = 1; b < NB; b++) {
    int p = h1v2_body_parent[b], d = 5 + b;
    double sd[6];
    cross_motion(cvel[p], k->S[d], sd);
    for (int i = 0; i < 6; i++) {
      cacc[b][i] = cacc[p][i] + sd[i] * qvel[d];
      cvel[b][i] = cvel[p][i] + k->S[d][i] * qvel[d];
    }
  }
  for (int b = 0; b < NB; b++) {
    double I6[36], Ia[6], Iv[6], vf[6];
    body_inertia6(k, b, mass_add, cfg->mass_recompute_inertia, I6);
    for (int r = 0; r < 6; r++) {
      Ia[r] = Iv[r] = 0;
      for (int c = 0; c < 6; c++) { Ia[r] += I6[6 * r + c] * cacc[b][c]; Iv[r] += I6[6 * r + c] * cvel[b][c]; }
    }
    cross_force(cvel[b], Iv, vf);
    for (int i = 0; i < 6; i++) f[b][i] = Ia[i] + vf[i];
  }
  for (int b = NB - 1; b >= 1; b--)
    for (int i = 0; i < 6; i++) f[h1v2_body_parent[b]][i] += f[b][i];
  for (int d = 0; d < NV; d++) {
    double s = 0;
    for (int i = 0; i < 6; i++) s += k->S[d][i] * f[body_of_dof[d]][i];
    bias[d] = s;
  }
  if (cvel_out) memcpy(cvel_out, cvel, sizeof(cvel));
}

/* ------------------------------------------------------------------------------------------ */
/* soft-constraint parameters (MuJoCo mj_makeImpedance / getimpedance semantics) */
static double impedance(const float* solimp, double pos) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (dmin < MINIMP) dmin = MINIMP; if (dmin > MAXIMP) dmin = MAXIMP;
  if (dmax < MINIMP) dmax = MINIMP; if (dmax > MAXIMP) dmax = MAXIMP;
  if (mid < MINIMP) mid = MINIMP;   if (mid > MAXIMP) mid = MAXIMP;
  if (power < 1) power = 1;
  if (dmin == dmax || width <= MINVAL) return 0.5 * (dmin + dmax);
  double x = fabs(pos) / width;
  if (x >= 1) return dmax;
  if (x == 0) return dmin;
  double y;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}
static void kb_from_solref(const float* solref, const float* solimp, double dt, double* K, double* B) {
  double tc = solref[0], dr = solref[1], dmax = solimp[1];
  if (dmax < MINIMP) dmax = MINIMP; if (dmax > MAXIMP) dmax = MAXIMP;
  if (tc < 2 * dt) tc = 2 * dt; /* refsafe */
  double kd = dmax * dmax * tc * tc * dr * dr, bd = dmax * tc;
  *K = 1.0 / (kd > MINVAL ? kd : MINVAL);
  *B = 2.0 / (bd > MINVAL ? bd : MINVAL);
}


/* ------------------------------------------------------------------------------------------ */
/* Rough terrain.  Restates what upstream isaaclab builds from the reference's generator cfg (T/utils/mdp/terrains.py:11-28 through
 * V/velocity_env_cfg.py:40-58): a grid of rows x cols square tiles (terrain_generator.py curriculum layout: row = difficulty level,
 * column = terrain type), every tile a height field of tile/hscale + 1 vertices per side whose rim of border_px vertices is flat and
 * whose interior vertices hold independent uniform draws from {level_min, level_min + step, .. level_max} * vscale
 * (height_field/hf_terrains.py random_uniform_terrain with downsampled_scale == horizontal_scale; the @height_field_to_mesh decorator
 * adds the rim), meshed by convert_height_field_to_mesh: every cell is split along the diagonal (i,j) -> (i+1,j+1).  The mesh is
 * centred on the world origin; around it lies the flat border at z = 0 (terrain_generator.py _add_terrain_border).  The tile origin
 * is the tile centre at the highest vertex of the central 2 m x 2 m patch (height_field/utils.py).  [UPSTREAM, from memory of
 * isaaclab 2.1.0; the draws themselves come from this backend's Philox stream, not numpy's generator.] */
static int terrain_npx(const H1v2Config* c) { return (int)lroundf(c->terrain_tile_size / c->terrain_hscale); }

static void terrain_origin_heights(H1v2Oracle* o) {
  const H1v2Config* c = &o->cfg;
  const int npx = o->npx;
  const int a1 = (int)((c->terrain_tile_size * 0.5f - 1.0f) / c->terrain_hscale), a2 = (int)((c->terrain_tile_size * 0.5f + 1.0f) / c->terrain_hscale);
  for (int r = 0; r < c->terrain_rows; r++)
    for (int q = 0; q < c->terrain_cols; q++) {
      float m = -1e30f;
      for (int a = a1; a < a2; a++)
        for (int b = a1; b < a2; b++) {
          float h = o->terrain[(size_t)(r * npx + a) * o->gy + (q * npx + b)];
          if (h > m) m = h;
        }
      o->origin_z[r * c->terrain_cols + q] = m;
    }
}

static void terrain_generate(H1v2Oracle* o) {
  const H1v2Config* c = &o->cfg;
  const int npx = o->npx, bp = c->terrain_border_px;
  const int nlev = (c->terrain_level_max - c->terrain_level_min) / (c->terrain_level_step > 0 ? c->terrain_level_step : 1) + 1;
  for (int i = 0; i < o->gx; i++)
    for (int j0 = 0; j0 < o->gy; j0 += 4) {
      float u[4];
      rng4(o->seed, (int64_t)i, (uint64_t)(j0 / 4), STREAM_TERRAIN, 0, u);
      for (int k = 0; k < 4 && j0 + k < o->gy; k++) {
        const int j = j0 + k, a = i % npx, b = j % npx;
        /* rim vertices of a tile (shared with its neighbour) are flat; the last grid line belongs to the last tile's rim */
        const int rim = a < bp || a > npx - bp || b < bp || b > npx - bp || i == o->gx - 1 || j == o->gy - 1;
        int lev = (int)(u[k] * (float)nlev);
        if (lev > nlev - 1) lev = nlev - 1;
        o->terrain[(size_t)i * o->gy + j] = rim ? 0.0f : (float)(c->terrain_level_min + lev * c->terrain_level_step) * c->terrain_vscale;
      }
    }
  terrain_origin_heights(o);
}

/* height (relative to the tile origin's z) and unit normal of the terrain under the point (lx, ly) given relative to the origin of tile
 * (level, type).  margin: distance (in cells) of the point from the nearest cell edge or diagonal, where the triangle changes */
static void terrain_query(const H1v2Oracle* o, int level, int type, double lx, double ly, double* h, double n[3], double* margin) {
  const H1v2Config* c = &o->cfg;
  const double half = (double)(0.5f * c->terrain_tile_size), inv = (double)(1.0f / c->terrain_hscale); /* the fp32 constants the kernel holds */
  const double oz = o->origin_z[level * c->terrain_cols + type];
  double a = (lx + half) * inv, b = (ly + half) * inv;
  double fa = floor(a), fb = floor(b);
  long i = (long)level * o->npx + (long)fa, j = (long)type * o->npx + (long)fb;
  double u = a - fa, v = b - fb;
  n[0] = 0; n[1] = 0; n[2] = 1;
  if (margin) *margin = 1.0;
  if (i < 0 || j < 0 || i >= o->gx - 1 || j >= o->gy - 1) { *h = -oz; return; } /* the flat border around the tiles */
  const float* H = o->terrain;
  double h00 = H[(size_t)i * o->gy + j], h10 = H[(size_t)(i + 1) * o->gy + j], h01 = H[(size_t)i * o->gy + j + 1], h11 = H[(size_t)(i + 1) * o->gy + j + 1];
  double dx, dy;
  if (v >= u) { dx = h11 - h01; dy = h01 - h00; } else { dx = h10 - h00; dy = h11 - h10; }
  *h = h00 + u * dx + v * dy - oz;
  double gx = dx * inv, gy = dy * inv, s = 1.0 / sqrt(1.0 + gx * gx + gy * gy);
  n[0] = -gx * s; n[1] = -gy * s; n[2] = s;
  if (margin) {
    double m = fmin(fmin(u, 1 - u), fmin(v, 1 - v));
    *margin = fmin(m, fabs(u - v));
  }
}

/* contact frame from the normal, MuJoCo mju_makeFrame: t1 = the y axis (z if the normal is within 30 deg of y) made orthogonal to n,
 * t2 = n x t1.  On the plane: t1 = y, t2 = -x. */
static void contact_frame(const double n[3], double t1[3], double t2[3]) {
  double y[3] = {0, 0, 0};
  if (n[1] < 0.5 && n[1] > -0.5) y[1] = 1; else y[2] = 1;
  double d = dot3(n, y), nn = 0;
  for (int i = 0; i < 3; i++) { t1[i] = y[i] - n[i] * d; nn += t1[i] * t1[i]; }
  nn = sqrt(nn);
  for (int i = 0; i < 3; i++) t1[i] /= nn;
  cross3(n, t1, t2);
}

typedef struct {
  int n;
  double dirv[MAXROW][3]; /* world direction of a contact row's force: n +- mu t */
  double min_tri_margin;  /* rough terrain: smallest distance (cells) of an active or nearly active contact point to a triangle boundary */
  double J[MAXROW][NV];
  double aref[MAXROW], R[MAXROW], D[MAXROW], floss[MAXROW];
  int type[MAXROW];  /* 0 friction-loss (two-sided box), 1 unilateral (limit / pyramid edge) */
  double min_abs_dist, min_limit_dist;
  int coll[MAXROW];  /* collider index for contact rows, -1 otherwise */
  int edge[MAXROW];  /* pyramid edge 0..3 */
  double force[MAXROW];
} Rows;

/* row cost s(jar), force = -ds/djar, active = inside the quadratic zone */
static double row_eval(const Rows* r, int i, double jar, double* force, int* active) {
  if (r->type[i] == 0) {
    double f = r->floss[i], lim = r->R[i] * f;
    if (jar <= -lim) { *force = f; *active = 0; return f * (-0.5 * lim - jar); }
    if (jar >= lim) { *force = -f; *active = 0; return f * (-0.5 * lim + jar); }
    *force = -r->D[i] * jar; *active = 1; return 0.5 * r->D[i] * jar * jar;
  }
  if (jar < 0) { *force = -r->D[i] * jar; *active = 1; return 0.5 * r->D[i] * jar * jar; }
  *force = 0; *active = 0; return 0;
}

static void build_rows(const H1v2Oracle* o, const Kin* k, const OEnv* e, Rows* r) {
  const H1v2Config* cfg = &o->cfg;
  const double dt = cfg->sim_dt;
  const double* qvel = e->qvel;
  r->n = 0;
  r->min_abs_dist = 1e30; r->min_limit_dist = 1e30; r->min_tri_margin = 1.0;
  double K, B;
  /* (a) dof friction loss */
  kb_from_solref(cfg->floss_solref, cfg->floss_solimp, dt, &K, &B);
  for (int d = 0; d < NV; d++) {
    if (cfg->dof_frictionloss[d] <= 0) continue;
    int i = r->n++;
    memset(r->J[i], 0, sizeof(r->J[i]));
    r->J[i][d] = 1.0;
    double imp = impedance(cfg->floss_solimp, 0.0);
    r->aref[i] = -B * qvel[d];
    r->R[i] = fmax(MINVAL, (1 - imp) * h1v2_dof_invweight0[d] / imp);
    r->D[i] = 1.0 / r->R[i];
    r->floss[i] = cfg->dof_frictionloss[d];
    r->type[i] = 0; r->coll[i] = -1; r->edge[i] = 0;
  }
  /* (b) joint limits, margin 0 */
  kb_from_solref(cfg->limit_solref, cfg->limit_solimp, dt, &K, &B);
  for (int j = 0; j < NJ; j++) {
    double q = e->qpos[7 + j];
    for (int side = 0; side < 2; side++) {
      double dist = side == 0 ? q - cfg->joint_range[j][0] : cfg->joint_range[j][1] - q;
      if (fabs(dist) < r->min_limit_dist) r->min_limit_dist = fabs(dist);
      if (dist >= 0) continue;
      int i = r->n++;
      memset(r->J[i], 0, sizeof(r->J[i]));
      r->J[i][6 + j] = side == 0 ? 1.0 : -1.0;
      double imp = impedance(cfg->limit_solimp, dist);
      double vel = r->J[i][6 + j] * qvel[6 + j];
      r->aref[i] = -B * vel - K * imp * dist;
      r->R[i] = fmax(MINVAL, (1 - imp) * h1v2_dof_invweight0[6 + j] / imp);
      r->D[i] = 1.0 / r->R[i];
      r->floss[i] = 0; r->type[i] = 1; r->coll[i] = -1; r->edge[i] = 0;
    }
  }
  /* (c) sphere/point-vs-plane contacts, condim 3, pyramidal cone */
  kb_from_solref(cfg->contact_solref, cfg->contact_solimp, dt, &K, &B);
  const double mu = e->friction;
  for (int c = 0; c < H1V2_NCOLL; c++) {
    int b = (int)h1v2_coll[c][0];
    double rad = h1v2_coll[c][4];
    int slot = (int)h1v2_coll[c][5];
    double off[3], ctr[3];
    matvec3(k->R[b], &h1v2_coll[c][1], off);
    for (int i = 0; i < 3; i++) ctr[i] = k->x[b][i] + off[i];
    double nrm[3] = {0, 0, 1}, hterr = 0, tri = 1.0;
    if (cfg->terrain_enable) terrain_query(o, e->level, e->type, ctr[0], ctr[1], &hterr, nrm, &tri);
    /* point / sphere against the plane of the triangle under it: distance along the plane normal */
    double dist = cfg->terrain_enable ? (ctr[2] - hterr) * nrm[2] - rad : ctr[2] - rad;
    if (fabs(dist) < r->min_abs_dist) r->min_abs_dist = fabs(dist);
    if (dist < 0.02 && tri < r->min_tri_margin) r->min_tri_margin = tri;
    if (dist >= 0) continue;
    double p[3] = {ctr[0], ctr[1], 0.5 * dist}; /* midway between the surfaces: centre - n (rad + dist / 2) */
    if (cfg->terrain_enable)
      for (int i = 0; i < 3; i++) p[i] = ctr[i] - nrm[i] * (rad + 0.5 * dist);
    /* translational jacobian of the body-fixed point at p: v = v_O + w x p */
    double Jp[3][NV];
    for (int d = 0; d < NV; d++) {
      if (!is_ancestor_or_self(body_of_dof[d], b)) { Jp[0][d] = Jp[1][d] = Jp[2][d] = 0; continue; }
      double wxp[3];
      cross3(k->S[d], p, wxp);
      for (int i = 0; i < 3; i++) Jp[i][d] = k->S[d][3 + i] + wxp[i];
    }
    double imp = impedance(cfg->contact_solimp, dist);
    double tran = h1v2_slot_invweight_tran[slot];
    double dA = tran + mu * mu * tran;
    double R0 = fmax(MINVAL, (1 - imp) * dA / imp);
    double Rpy = fmax(MINVAL, 2 * mu * mu * R0);
    /* edges: n + mu*t1, n - mu*t1, n + mu*t2, n - mu*t2 (mju_makeFrame; on the plane n = z, t1 = y, t2 = -x) */
    if (!cfg->terrain_enable) { /* the plane: the arithmetic of the flat ids, operation for operation (profiles/roofline.json counts it) */
      static const double ex[4] = {0, 0, -1, 1}, ey[4] = {1, -1, 0, 0};
      for (int ed = 0; ed < 4; ed++) {
        int i = r->n++;
        double vel = 0;
        for (int d = 0; d < NV; d++) {
          r->J[i][d] = Jp[2][d] + mu * (ex[ed] * Jp[0][d] + ey[ed] * Jp[1][d]);
          vel += r->J[i][d] * qvel[d];
        }
        r->aref[i] = -B * vel - K * imp * dist;
        r->R[i] = Rpy; r->D[i] = 1.0 / Rpy; r->floss[i] = 0; r->type[i] = 1; r->coll[i] = c; r->edge[i] = ed;
      }
      continue;
    }
    double t1[3], t2[3];
    contact_frame(nrm, t1, t2);
    for (int ed = 0; ed < 4; ed++) {
      int i = r->n++;
      const double* t = ed < 2 ? t1 : t2;
      const double sg = (ed & 1) ? -1.0 : 1.0;
      for (int a = 0; a < 3; a++) r->dirv[i][a] = nrm[a] + mu * sg * t[a];
      double vel = 0;
      for (int d = 0; d < NV; d++) {
        r->J[i][d] = r->dirv[i][0] * Jp[0][d] + r->dirv[i][1] * Jp[1][d] + r->dirv[i][2] * Jp[2][d];
        vel += r->J[i][d] * qvel[d];
      }
      r->aref[i] = -B * vel - K * imp * dist;
      r->R[i] = Rpy; r->D[i] = 1.0 / Rpy; r->floss[i] = 0; r->type[i] = 1; r->coll[i] = c; r->edge[i] = ed;
    }
  }
}

/* primal Newton with exact line search on the convex cost (MuJoCo mj_solNewton semantics, cold start) */
static void solve_constraints(const H1v2Oracle* o, const double M[NV * NV], const double qacc_smooth[NV], Rows* r,
                              double qacc[NV], int* iters_out, double* resid_out) {
  const int n = r->n;
  const double scale = 1.0 / (H1V2_MEANINERTIA * NV);
  double jar[MAXROW], Ma[NV], grad[NV], search[NV], Jv[MAXROW];
  memcpy(qacc, qacc_smooth, NV * sizeof(double));
  int it = 0;
  double gnorm = 0;
  for (it = 0; it < o->max_newton_iters; it++) {
    /* gradient at qacc */
    for (int i = 0; i < n; i++) {
      double s = -r->aref[i];
      for (int d = 0; d < NV; d++) s += r->J[i][d] * qacc[d];
      jar[i] = s;
    }
    for (int d = 0; d < NV; d++) {
      double s = 0;
      for (int c = 0; c < NV; c++) s += M[d * NV + c] * (qacc[c] - qacc_smooth[c]);
      Ma[d] = s; grad[d] = s;
    }
    double H[NV * NV];
    memcpy(H, M, sizeof(H));
    for (int i = 0; i < n; i++) {
      double f; int act;
      row_eval(r, i, jar[i], &f, &act);
      r->force[i] = f;
      for (int d = 0; d < NV; d++) grad[d] -= r->J[i][d] * f;
      if (act)
        for (int a = 0; a < NV; a++) {
          if (r->J[i][a] == 0) continue;
          for (int b = 0; b < NV; b++) H[a * NV + b] += r->D[i] * r->J[i][a] * r->J[i][b];
        }
    }
    gnorm = 0;
    for (int d = 0; d < NV; d++) gnorm += grad[d] * grad[d];
    gnorm = sqrt(gnorm);
    if (gnorm * scale < 1e-12) break;
    if (chol(H, NV) != 0) { fprintf(stderr, "oracle: Hessian not SPD\n"); break; }
    for (int d = 0; d < NV; d++) search[d] = -grad[d];
    chol_solve(H, NV, search);
    /* exact line search: root of the increasing piecewise-linear phi'(alpha) */
    for (int i = 0; i < n; i++) {
      double s = 0;
      for (int d = 0; d < NV; d++) s += r->J[i][d] * search[d];
      Jv[i] = s;
    }
    double sMs = 0, sMa = 0;
    for (int d = 0; d < NV; d++) {
      double s = 0;
      for (int c = 0; c < NV; c++) s += M[d * NV + c] * search[c];
      sMs += search[d] * s; sMa += search[d] * Ma[d];
    }
#define DPHI(alpha, d1, d2)                                      \
  do {                                                           \
    d1 = sMa + (alpha) * sMs; d2 = sMs;                          \
    for (int i_ = 0; i_ < n; i_++) {                             \
      double f_; int a_;                                         \
      row_eval(r, i_, jar[i_] + (alpha) * Jv[i_], &f_, &a_);     \
      d1 -= f_ * Jv[i_];                                         \
      if (a_) d2 += r->D[i_] * Jv[i_] * Jv[i_];                  \
    }                                                            \
  } while (0)
    double lo = 0, hi = 1, d1, d2, alpha;
    DPHI(hi, d1, d2);
    int guard = 0;
    while (d1 < 0 && guard++ < 60) { lo = hi; hi *= 2; DPHI(hi, d1, d2); }
    alpha = hi;
    for (int ls = 0; ls < 200; ls++) {
      DPHI(alpha, d1, d2);
      if (fabs(d1) < 1e-15 * (1 + fabs(sMa))) break;
      if (d1 > 0) hi = alpha; else lo = alpha;
      double nxt = alpha - d1 / d2;
      if (!(nxt > lo && nxt < hi)) nxt = 0.5 * (lo + hi);
      if (nxt == alpha) break;
      alpha = nxt;
    }
#undef DPHI
    for (int d = 0; d < NV; d++) qacc[d] += alpha * search[d];
  }
  /* final forces at the solution */
  for (int i = 0; i < n; i++) {
    double s = -r->aref[i], f; int act;
    for (int d = 0; d < NV; d++) s += r->J[i][d] * qacc[d];
    row_eval(r, i, s, &f, &act);
    r->force[i] = f;
  }
  *iters_out = it;
  *resid_out = gnorm * scale;
}

/* ------------------------------------------------------------------------------------------ */
static void quat_integrate(double q[4], const double w[3], double h) {
  double ang = sqrt(dot3(w, w)) * h;
  double dq[4] = {1, 0, 0, 0};
  if (ang > 0) {
    double s = sin(0.5 * ang) / (ang / h);
    dq[0] = cos(0.5 * ang); dq[1] = w[0] * s; dq[2] = w[1] * s; dq[3] = w[2] * s;
  }
  double r[4] = {q[0] * dq[0] - q[1] * dq[1] - q[2] * dq[2] - q[3] * dq[3],
                 q[0] * dq[1] + q[1] * dq[0] + q[2] * dq[3] - q[3] * dq[2],
                 q[0] * dq[2] - q[1] * dq[3] + q[2] * dq[0] + q[3] * dq[1],
                 q[0] * dq[3] + q[1] * dq[2] - q[2] * dq[1] + q[3] * dq[0]};
  double n = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
  for (int i = 0; i < 4; i++) q[i] = r[i] / n;
}

/* one physics substep with joint torques ctrl (already effort-clipped by the actuator model) */
static void physics_substep(const H1v2Oracle* o, OEnv* e, const double ctrl[NJ]) {
  const H1v2Config* cfg = &o->cfg;
  const double h = cfg->sim_dt;
  Kin k;
  kinematics(e->qpos, &k);
  double M[NV * NV], bias[NV], fs[NV], qacc_s[NV], L[NV * NV];
  mass_matrix(cfg, &k, e->mass_add, M);
  rne_bias(cfg, &k, e->qvel, e->mass_add, bias, NULL);
  for (int d = 0; d < NV; d++) fs[d] = -bias[d] - cfg->dof_damping[d] * e->qvel[d];
  for (int j = 0; j < NJ; j++) {
    double u = ctrl[j], lim = cfg->act_frc_limit[j];
    if (lim > 0) { if (u > lim) u = lim; if (u < -lim) u = -lim; }
    fs[6 + j] += u;
  }
  memcpy(L, M, sizeof(L));
  chol(L, NV);
  memcpy(qacc_s, fs, sizeof(fs));
  chol_solve(L, NV, qacc_s);
  static _Thread_local Rows rows;
  build_rows(o, &k, e, &rows);
  if (rows.min_abs_dist < e->min_abs_dist) e->min_abs_dist = rows.min_abs_dist;
  if (rows.min_tri_margin < e->min_tri_margin) e->min_tri_margin = rows.min_tri_margin;
  if (rows.min_limit_dist < e->min_limit_dist) e->min_limit_dist = rows.min_limit_dist;
  double qacc[NV];
  solve_constraints(o, M, qacc_s, &rows, qacc, &e->newton_iters, &e->newton_resid);
  /* constraint force and per-slot net contact force */
  double fc[NV] = {0};
  memset(e->slot_force, 0, sizeof(e->slot_force));
  for (int i = 0; i < rows.n; i++) {
    for (int d = 0; d < NV; d++) fc[d] += rows.J[i][d] * rows.force[i];
    if (rows.coll[i] >= 0) { /* pyramid edge force f along n +- mu t */
      int slot = (int)h1v2_coll[rows.coll[i]][5];
      if (cfg->terrain_enable) {
        for (int a = 0; a < 3; a++) e->slot_force[slot][a] += rows.force[i] * rows.dirv[i][a];
      } else { /* the plane: n = z, t1 = y, t2 = -x */
        static const double ex[4] = {0, 0, -1, 1}, ey[4] = {1, -1, 0, 0};
        double f = rows.force[i], mu = e->friction;
        e->slot_force[slot][0] += f * mu * ex[rows.edge[i]];
        e->slot_force[slot][1] += f * mu * ey[rows.edge[i]];
        e->slot_force[slot][2] += f;
      }
    }
  }
  /* implicitfast: (M + h*diag(damping)) qacc = qfrc_smooth + qfrc_constraint */
  double A[NV * NV], rhs[NV];
  memcpy(A, M, sizeof(A));
  for (int d = 0; d < NV; d++) { A[d * NV + d] += h * cfg->dof_damping[d]; rhs[d] = fs[d] + fc[d]; }
  chol(A, NV);
  chol_solve(A, NV, rhs);
  for (int j = 0; j < NJ; j++) e->joint_acc[j] = rhs[6 + j];
  for (int d = 0; d < NV; d++) e->qvel[d] += h * rhs[d];
  if (cfg->joint_vel_limit > 0.f) /* actuator velocity_limit (A/robots/h12.py:66,89,103), PhysX joint velocity clamp */
    for (int j = 0; j < NJ; j++) e->qvel[6 + j] = fmin(fmax(e->qvel[6 + j], -(double)cfg->joint_vel_limit), (double)cfg->joint_vel_limit);
  for (int i = 0; i < 3; i++) e->qpos[i] += h * e->qvel[i];
  quat_integrate(e->qpos + 3, e->qvel + 3, h);
  for (int j = 0; j < NJ; j++) e->qpos[7 + j] += h * e->qvel[6 + j];
}

/* world velocity of the ankle_roll_links (body_lin_vel_w of the feet): isaaclab 2.1.0 ArticulationData reports link velocities at
 * the link's centre of mass (at_com), the MuJoCo-side convention would be the link origin */
static void foot_velocities(const OEnv* e, int at_com, double fv[2][3]) {
  Kin k;
  kinematics(e->qpos, &k);
  static const int foot_body[2] = {6, 12};
  for (int f = 0; f < 2; f++) {
    int b = foot_body[f];
    double v[6] = {0};
    for (int d = 0; d < NV; d++)
      if (is_ancestor_or_self(body_of_dof[d], b))
        for (int i = 0; i < 6; i++) v[i] += k.S[d][i] * e->qvel[d];
    double wxp[3], p[3] = {k.x[b][0], k.x[b][1], k.x[b][2]};
    if (at_com) {
      double c[3];
      matvec3(k.R[b], h1v2_body_ipos[b], c);
      for (int i = 0; i < 3; i++) p[i] += c[i];
    }
    cross3(v, p, wxp);
    for (int i = 0; i < 3; i++) fv[f][i] = v[3 + i] + wxp[i];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* managers */
static float wrap_to_pi_f(float a) {
  /* isaaclab.utils.math.wrap_to_pi (2.1.0): ((a+pi) mod 2pi) - pi, with +pi for positive odd multiples */
  float two_pi = 2.0f * PI_F;
  float w = fmodf(a + PI_F, two_pi);
  if (w < 0) w += two_pi;
  if (w == 0.0f && a > 0.0f) return PI_F;
  return w - PI_F;
}

static void resample_command(H1v2Oracle* o, int ei, uint32_t block0) {
  OEnv* e = &o->env[ei];
  const H1v2Config* c = &o->cfg;
  float u[4], v[4];
  rng4(o->seed, c->env_id_offset + ei, o->step_counter, STREAM_CMD, block0, u);
  rng4(o->seed, c->env_id_offset + ei, o->step_counter, STREAM_CMD, block0 + 1, v);
  e->time_left = uni(v[2], c->cmd_resample_time[0], c->cmd_resample_time[1]);
  e->cmd[0] = uni(u[0], c->cmd_lin_x[0], c->cmd_lin_x[1]);
  e->cmd[1] = uni(u[1], c->cmd_lin_y[0], c->cmd_lin_y[1]);
  e->cmd[2] = uni(u[2], c->cmd_ang_z[0], c->cmd_ang_z[1]);
  if (c->heading_command) {
    e->heading_target = uni(u[3], c->cmd_heading[0], c->cmd_heading[1]);
    e->is_heading = v[0] <= c->rel_heading_envs;
  }
  e->is_standing = v[1] <= c->rel_standing_envs;
}

/* mdp.terrain_levels_vel (V/mdp/curriculums.py:21-52) + TerrainImporter.update_env_origins [UPSTREAM]: evaluated in _reset_idx on the
 * state the episode ended in.  fp32 like the torch tensors it runs on; the position is relative to the env origin already. */
static void terrain_curriculum(H1v2Oracle* o, int ei) {
  OEnv* e = &o->env[ei];
  const H1v2Config* c = &o->cfg;
  if (!c->terrain_enable || !c->terrain_curriculum) return;
  const float x = (float)e->qpos[0], y = (float)e->qpos[1];
  const float dist = sqrtf(fmaf(y, y, x * x));
  const float cn = sqrtf(fmaf(e->cmd[1], e->cmd[1], e->cmd[0] * e->cmd[0]));
  const int up = dist > c->terrain_tile_size * 0.5f;
  const int down = (dist < cn * c->episode_length_s * 0.5f) && !up;
  int lev = e->level + up - down;
  if (lev >= c->terrain_rows) { /* robots that solve the last level are sent to a random one */
    float u[4];
    rng4(o->seed, c->env_id_offset + ei, o->step_counter, STREAM_RESET, 10, u);
    lev = (int)(u[0] * (float)c->terrain_rows);
    if (lev > c->terrain_rows - 1) lev = c->terrain_rows - 1;
  } else if (lev < 0) lev = 0;
  e->level = lev;
}

static void reset_env(H1v2Oracle* o, int ei) {
  OEnv* e = &o->env[ei];
  const H1v2Config* c = &o->cfg;
  const int64_t gid = c->env_id_offset + ei;
  terrain_curriculum(o, ei); /* CurriculumManager.compute(env_ids) runs first in _reset_idx (cat_env.py:197-200) */
  float u0[4], u1[4], u2[4], u3[4];
  rng4(o->seed, gid, o->step_counter, STREAM_RESET, 0, u0);
  rng4(o->seed, gid, o->step_counter, STREAM_RESET, 1, u1);
  rng4(o->seed, gid, o->step_counter, STREAM_RESET, 2, u2);
  rng4(o->seed, gid, o->step_counter, STREAM_RESET, 3, u3);
  /* scene.reset: actuator delay line + lag, contact sensor */
  e->lag = c->min_delay + (int)(u1[2] * (float)(c->max_delay - c->min_delay + 1));
  if (e->lag > c->max_delay) e->lag = c->max_delay;
  e->fresh = 3; /* bit0: delay line empty, bit1: observation history empty */
  memset(e->timers, 0, sizeof(e->timers));
  memset(e->slot_force, 0, sizeof(e->slot_force));
  memset(e->slot_hist, 0, sizeof(e->slot_hist));
  /* reset_root_state_uniform */
  float px = uni(u0[0], c->reset_pose_range[0][0], c->reset_pose_range[0][1]);
  float py = uni(u0[1], c->reset_pose_range[1][0], c->reset_pose_range[1][1]);
  float pz = uni(u0[3], c->reset_pose_range[2][0], c->reset_pose_range[2][1]);
  float roll = uni(u1[0], c->reset_pose_range[3][0], c->reset_pose_range[3][1]);
  float pitch = uni(u1[1], c->reset_pose_range[4][0], c->reset_pose_range[4][1]);
  float yaw = uni(u0[2], c->reset_pose_range[5][0], c->reset_pose_range[5][1]);
  e->qpos[0] = px; e->qpos[1] = py; e->qpos[2] = (double)c->init_root_height + pz;
  double cr = cos(0.5 * roll), sr = sin(0.5 * roll), cp = cos(0.5 * pitch), sp = sin(0.5 * pitch);
  double cy = cos(0.5 * yaw), sy = sin(0.5 * yaw);
  e->qpos[3] = cy * cr * cp + sy * sr * sp;
  e->qpos[4] = cy * sr * cp - sy * cr * sp;
  e->qpos[5] = cy * cr * sp + sy * sr * cp;
  e->qpos[6] = sy * cr * cp - cy * sr * sp;
  double vw[3], ww[3], R[9];
  for (int i = 0; i < 3; i++) {
    vw[i] = uni(u2[i], c->reset_vel_range[i][0], c->reset_vel_range[i][1]);
    ww[i] = uni(u3[i], c->reset_vel_range[3 + i][0], c->reset_vel_range[3 + i][1]);
  }
  quat2mat(e->qpos + 3, R);
  for (int i = 0; i < 3; i++) e->qvel[i] = vw[i];
  mattvec3(R, ww, e->qvel + 3);
  /* reset_joints_by_scale */
  for (int j = 0; j < NJ; j++) {
    float uj[4];
    rng4(o->seed, gid, o->step_counter, STREAM_RESET, 4 + (uint32_t)(j / 2), uj);
    float sp_ = uni(uj[2 * (j % 2)], c->reset_joint_pos_scale[0], c->reset_joint_pos_scale[1]);
    float sv_ = uni(uj[2 * (j % 2) + 1], c->reset_joint_vel_scale[0], c->reset_joint_vel_scale[1]);
    double q = (double)c->default_joint_pos[j] * sp_;
    if (q < o->soft_lo[j]) q = o->soft_lo[j];
    if (q > o->soft_hi[j]) q = o->soft_hi[j];
    e->qpos[7 + j] = q;
    e->qvel[6 + j] = 0.0 * sv_;
  }
  /* managers */
  memset(e->hist, 0, sizeof(e->hist));
  memset(e->last_action, 0, sizeof(e->last_action));
  memset(e->T1, 0, sizeof(e->T1));
  memset(e->T2, 0, sizeof(e->T2));
  memset(e->ep_sums, 0, sizeof(e->ep_sums));
  e->metrics[0] = e->metrics[1] = 0;
  resample_command(o, ei, 0);
  if (c->push_enable) {
    float up[4];
    rng4(o->seed, gid, o->step_counter, STREAM_EVENT, 1, up);
    e->push_left = uni(up[0], c->push_interval_s[0], c->push_interval_s[1]);
  }
  e->ep_len = 0;
}

/* vb / vw: velocity of the root LINK's centre of mass in the base / world frame (isaaclab 2.1.0 root_lin_vel_b / root_lin_vel_w,
 * SURVEY App. A); the state itself holds MuJoCo's pelvis-origin velocity: v_com = v_origin + w x (R r_com) */
static void root_derived(const H1v2Config* c, const OEnv* e, double R[9], double vb[3], double wb[3], double g[3], double* heading,
                         double ww[3], double vw[3]) {
  double q[4] = {e->qpos[3], e->qpos[4], e->qpos[5], e->qpos[6]};
  quat2mat(q, R);
  mattvec3(R, e->qvel, vb);
  for (int i = 0; i < 3; i++) wb[i] = e->qvel[3 + i];
  double rc[3] = {c->root_link_com[0], c->root_link_com[1], c->root_link_com[2]}, wxr[3];
  cross3(wb, rc, wxr);
  for (int i = 0; i < 3; i++) vb[i] += wxr[i];
  matvec3(R, vb, vw);
  matvec3(R, wb, ww);
  g[0] = -R[6]; g[1] = -R[7]; g[2] = -R[8]; /* R^T (0,0,-1) */
  *heading = atan2(R[3], R[0]);
}

/* norm(cmd_xy) < dead zone (commands.py:62), compared on the squares with one multiply and one fused multiply-add like the kernel */
static int cmd_in_deadzone(const float* cmd, float deadzone) { return fmaf(cmd[1], cmd[1], cmd[0] * cmd[0]) < deadzone * deadzone; }

static void update_command(H1v2Oracle* o, int ei) {
  OEnv* e = &o->env[ei];
  const H1v2Config* c = &o->cfg;
  double R[9], vb[3], wb[3], g[3], heading, ww[3], vw[3];
  root_derived(c, e, R, vb, wb, g, &heading, ww, vw);
  /* _update_metrics */
  float max_command_step = c->cmd_resample_time[1] / (c->sim_dt * (float)c->decimation);
  float ex = e->cmd[0] - (float)vb[0], ey = e->cmd[1] - (float)vb[1];
  e->metrics[0] += sqrtf(ex * ex + ey * ey) / max_command_step;
  e->metrics[1] += fabsf(e->cmd[2] - (float)wb[2]) / max_command_step;
  /* time_left -= dt ; resample */
  e->time_left -= c->sim_dt * (float)c->decimation;
  if (e->time_left <= 0.0f) resample_command(o, ei, 2);
  /* _update_command */
  if (c->heading_command && e->is_heading) {
    float err = wrap_to_pi_f(e->heading_target - (float)heading);
    float w = c->heading_stiffness * err;
    if (w < c->cmd_ang_z[0]) w = c->cmd_ang_z[0];
    if (w > c->cmd_ang_z[1]) w = c->cmd_ang_z[1];
    e->cmd[2] = w;
  }
  if (c->command_class == 0) {
    if (e->is_standing) e->cmd[0] = e->cmd[1] = e->cmd[2] = 0;
  } else {
    /* UniformVelocityCommandWithDeadzone._update_command (T/utils/mdp/commands.py:41-96): standing envs are not zeroed by the
     * override.  Balancing (:62-83): the reference counts the envs whose norm(cmd_xy) is below the dead zone and moves exactly
     * |n // 2 - count| envs across it, picked by torch.randperm: active ones get cmd_xy = 0, dead-zone ones a fresh command through
     * _resample (new command and new time_left).  Restated per env: an independent draw with the probability that moves the same
     * number in expectation, computed from the count at the end of the PREVIOUS step (the product runs one launch per step and
     * has no process-wide count inside it).  With velocity_deadzone == 0 (C12/rsl_env_cfg.py:98) `norm < 0` never holds, the
     * count stays 0 and every env loses its xy command with probability (n // 2) / n on EVERY step.  Then the yaw-rate command
     * changes sign with probability physics_dt / max_episode_length_s (:85-96). */
    float u[4];
    rng4(o->seed, c->env_id_offset + ei, o->step_counter, STREAM_CMD, 4, u);
    const int target = o->n / 2, cur = o->dz_prev < o->n ? o->dz_prev : o->n;
    int in_dz = cmd_in_deadzone(e->cmd, c->velocity_deadzone);
    if (cur < target) {
      if (!in_dz && u[0] < (float)(target - cur) / (float)(o->n - cur)) e->cmd[0] = e->cmd[1] = 0;
    } else if (cur > target) {
      if (in_dz && u[0] < (float)(cur - target) / (float)cur) resample_command(o, ei, 6);
    }
    if (u[1] < c->ang_vel_flip_prob) e->cmd[2] = -e->cmd[2];
  }
}

/* rays along one axis of a GridPatternCfg: len(torch.arange(-size/2, size/2 + 1e-9, resolution)); the 1e-4 absorbs the fp32 rounding of the
 * config values (0.1f > 0.1) */
static int scan_count(float size, float res) { return (int)floor((double)size / (double)res + 1e-4) + 1; }

static void compute_obs(H1v2Oracle* o, int ei, float* obs_out) {
  OEnv* e = &o->env[ei];
  const H1v2Config* c = &o->cfg;
  const int H = c->history_length;
  double R[9], vb[3], wb[3], g[3], heading, ww[3], vw[3];
  root_derived(c, e, R, vb, wb, g, &heading, ww, vw);
  float s[H1V2_OBS_TERM_DIM];
  float n_av[3] = {0}, n_g[3] = {0}, n_q[NJ] = {0}, n_v[NJ] = {0};
  if (c->enable_corruption) {
    const int64_t gid = c->env_id_offset + ei;
    float b0[4], b1[4];
    rng4(o->seed, gid, o->step_counter, STREAM_OBS, 0, b0);
    rng4(o->seed, gid, o->step_counter, STREAM_OBS, 1, b1);
    for (int i = 0; i < 3; i++) {
      n_av[i] = uni(b0[i], -c->noise_ang_vel, c->noise_ang_vel);
      n_g[i] = uni(b1[i], -c->noise_gravity, c->noise_gravity);
    }
    for (int side = 0; side < 2; side++) {
      float a[4], b[4], d[4];
      rng4(o->seed, gid, o->step_counter, STREAM_OBS, 2 + 4 * side, a);
      rng4(o->seed, gid, o->step_counter, STREAM_OBS, 3 + 4 * side, b);
      rng4(o->seed, gid, o->step_counter, STREAM_OBS, 4 + 4 * side, d);
      float up[6] = {a[0], a[1], a[2], a[3], b[0], b[1]};
      float uv[6] = {b[2], b[3], d[0], d[1], d[2], d[3]};
      for (int k = 0; k < 6; k++) {
        n_q[6 * side + k] = uni(up[k], -c->noise_joint_pos, c->noise_joint_pos);
        n_v[6 * side + k] = uni(uv[k], -c->noise_joint_vel, c->noise_joint_vel);
      }
    }
  }
  for (int i = 0; i < 3; i++) {
    s[i] = (float)((wb[i] + n_av[i]) * c->scale_ang_vel);
    s[3 + i] = (float)((g[i] + n_g[i]) * c->scale_gravity);
    s[6 + i] = e->cmd[i] * c->scale_cmd;
  }
  for (int i = 0; i < NJ; i++) { /* external order */
    int j = c->joint_perm[i];
    s[9 + i] = (float)((e->qpos[7 + j] - c->default_joint_pos[j] + n_q[j]) * c->scale_joint_pos);
    s[21 + i] = (float)((e->qvel[6 + j] + n_v[j]) * c->scale_joint_vel);
    s[33 + i] = e->last_action[i] * c->scale_action;
  }
  if (c->obs_base_lin_vel || c->obs_height_scan) {
    /* Rough id (V/velocity_env_cfg.py:119-142): base_lin_vel | the 45 terms above | height_scan, no history */
    e->fresh &= ~2;
    memcpy(e->hist[0], s, sizeof(s));
    if (!obs_out) return;
    int w = 0;
    if (c->obs_base_lin_vel) {
      float n_lv[3] = {0, 0, 0};
      if (c->enable_corruption) {
        float b[4];
        rng4(o->seed, c->env_id_offset + ei, o->step_counter, STREAM_OBS, 10, b);
        for (int i = 0; i < 3; i++) n_lv[i] = uni(b[i], -c->noise_lin_vel, c->noise_lin_vel);
      }
      for (int i = 0; i < 3; i++) obs_out[w++] = (float)((vb[i] + n_lv[i]) * c->scale_lin_vel);
    }
    for (int k = 0; k < H1V2_OBS_TERM_DIM; k++) obs_out[w++] = s[k];
    if (c->obs_height_scan) {
      /* mdp.height_scan [UPSTREAM]: sensor.data.pos_w.z - ray_hits_w.z - offset over a GridPattern (x fastest, "xy" indexing) cast
       * straight down from the scanner body (torso_link: the pelvis frame, fixed joint h12_12dof.urdf:394-400), attach_yaw_only */
      const int nx = scan_count(c->scan_size[0], c->scan_resolution), ny = scan_count(c->scan_size[1], c->scan_resolution);
      const double cy = cos(heading), sy = sin(heading);
      for (int r = 0; r < nx * ny; r++) {
        const int ix = r % nx, iy = r / nx;
        const double gx = (double)((float)ix * c->scan_resolution - 0.5f * c->scan_size[0]), gy = (double)((float)iy * c->scan_resolution - 0.5f * c->scan_size[1]);
        const double lx = e->qpos[0] + cy * gx - sy * gy, ly = e->qpos[1] + sy * gx + cy * gy;
        double h = 0, nrm[3];
        if (c->terrain_enable) terrain_query(o, e->level, e->type, lx, ly, &h, nrm, NULL);
        float v = (float)(e->qpos[2] - h - c->scan_offset);
        if (c->enable_corruption) {
          float b[4];
          rng4(o->seed, c->env_id_offset + ei, o->step_counter, STREAM_OBS, 16 + (uint32_t)(r / 4), b);
          v += uni(b[r % 4], -c->noise_height_scan, c->noise_height_scan);
        }
        if (v < c->scan_clip[0]) v = c->scan_clip[0];
        if (v > c->scan_clip[1]) v = c->scan_clip[1];
        obs_out[w++] = v * c->scale_height_scan;
      }
    }
    return;
  }
  /* history append: first push after a reset fills every slot (circular_buffer.py:131-135) */
  if (e->fresh & 2) {
    for (int h = 0; h < H; h++) memcpy(e->hist[h], s, sizeof(s));
    e->fresh &= ~2;
  } else {
    for (int h = 0; h + 1 < H; h++) memcpy(e->hist[h], e->hist[h + 1], sizeof(s));
    memcpy(e->hist[H - 1], s, sizeof(s));
  }
  /* term-major flatten, oldest -> newest inside each term block (observation_manager.py:335-355) */
  if (obs_out) {
    static const int off[7] = {0, 3, 6, 9, 21, 33, 45};
    int w = 0;
    for (int t = 0; t < 6; t++)
      for (int h = 0; h < H; h++)
        for (int k = off[t]; k < off[t + 1]; k++) obs_out[w++] = e->hist[h][k];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* values the tail consumes that normally come out of the physics loop; tests may inject the CUDA kernel's own
 * post-physics values here so that rewards / masks / observations are compared on IDENTICAL states */
typedef struct H1v2Inject {
  const float* qpos;      /* [N,19] */
  const float* qvel;      /* [N,18] */
  const float* timers;    /* [N,8] */
  const float* slot_hist; /* [N,6,3] */
  const float* tau;       /* [N,12] */
  const float* qacc;      /* [N,12] */
  const float* foot_vel;  /* [N,6] */
} H1v2Inject;

static void step_env(H1v2Oracle* o, int ei, const float* action, float* rew_out, uint8_t* term_out, uint8_t* trunc_out,
                     int* reset_flag, int* nan_flag, const H1v2Inject* inj) {
  OEnv* e = &o->env[ei];
  const H1v2Config* c = &o->cfg;
  const float step_dt = c->sim_dt * (float)c->decimation;
  e->min_abs_dist = 1e30; e->min_limit_dist = 1e30; e->min_tri_margin = 1.0;
  /* -- action manager: process_action -- */
  float prev_action[NJ];
  memcpy(prev_action, e->last_action, sizeof(prev_action));
  double T0[NJ];
  for (int i = 0; i < NJ; i++) {
    e->last_action[i] = action[i];
    T0[c->joint_perm[i]] = (double)c->action_scale * action[i] + c->default_joint_pos[c->joint_perm[i]];
  }
  if (e->fresh & 1) { memcpy(e->T1, T0, sizeof(T0)); memcpy(e->T2, T0, sizeof(T0)); }
  /* -- physics loop -- */
  double qd_prev[NJ];
  for (int k = 0; k < (inj ? 0 : c->decimation); k++) {
    int age = e->lag - k; /* physics steps between the delayed sample and the current control step's first push */
    const double* T = age <= 0 ? T0 : (age <= c->decimation ? e->T1 : e->T2);
    double tau[NJ];
    for (int j = 0; j < NJ; j++) {
      double t = c->kp[j] * (T[j] - e->qpos[7 + j]) + c->kd[j] * (0.0 - e->qvel[6 + j]);
      if (t > c->effort_limit[j]) t = c->effort_limit[j];
      if (t < -c->effort_limit[j]) t = -c->effort_limit[j];
      tau[j] = t;
      e->applied_tau[j] = t;
      qd_prev[j] = e->qvel[6 + j];
    }
    physics_substep(o, e, tau);
    (void)qd_prev;
    /* contact sensor at sim dt: history of |F| and air/contact timers */
    const float el = c->sim_dt;
    for (int s = 0; s < NSLOT; s++) {
      double nf = sqrt(dot3(e->slot_force[s], e->slot_force[s]));
      e->slot_hist[s][0] = e->slot_hist[s][1]; e->slot_hist[s][1] = e->slot_hist[s][2]; e->slot_hist[s][2] = nf;
    }
    for (int f = 0; f < 2; f++) {
      float nf = (float)sqrt(dot3(e->slot_force[f], e->slot_force[f]));
      int is_c = nf > c->contact_threshold;
      float* t = e->timers[f]; /* cur_air,last_air,cur_contact,last_contact */
      int first_c = (t[0] > 0) && is_c, first_d = (t[2] > 0) && !is_c;
      t[1] = first_c ? t[0] + el : t[1];
      t[0] = is_c ? 0.0f : t[0] + el;
      t[3] = first_d ? t[2] + el : t[3];
      t[2] = is_c ? t[2] + el : 0.0f;
    }
  }
  memcpy(e->T2, e->T1, sizeof(T0));
  memcpy(e->T1, T0, sizeof(T0));
  e->fresh &= ~1;
  if (inj) {
    for (int i = 0; i < 19; i++) e->qpos[i] = inj->qpos[(size_t)ei * 19 + i];
    for (int i = 0; i < 18; i++) e->qvel[i] = inj->qvel[(size_t)ei * 18 + i];
    for (int i = 0; i < 8; i++) e->timers[i / 4][i % 4] = inj->timers[(size_t)ei * 8 + i];
    for (int i = 0; i < 18; i++) e->slot_hist[i / 3][i % 3] = inj->slot_hist[(size_t)ei * 18 + i];
    for (int i = 0; i < NJ; i++) { e->applied_tau[i] = inj->tau[(size_t)ei * NJ + i]; e->joint_acc[i] = inj->qacc[(size_t)ei * NJ + i]; }
    if (inj->foot_vel)
      for (int i = 0; i < 6; i++) e->foot_vel[i / 3][i % 3] = inj->foot_vel[(size_t)ei * 6 + i];
    else /* not injected: the oracle's own kinematics on the injected state (checks the kernel's foot velocity) */
      foot_velocities(e, o->cfg.body_vel_at_com, e->foot_vel);
  } else {
    foot_velocities(e, o->cfg.body_vel_at_com, e->foot_vel);
  }
  /* non-finite guard: force a reset, zero reward */
  int bad = 0;
  for (int i = 0; i < 19; i++) if (!isfinite(e->qpos[i])) bad = 1;
  const double runaway = c->runaway_vel > 0.f ? (double)c->runaway_vel : 3.0e38;
  for (int i = 0; i < 18; i++) if (!(fabs((double)(float)e->qvel[i]) <= runaway)) bad = 1;
  /* an action that is non-finite or absurdly large (|a| > H1V2_ACTION_ABS_MAX) is contained the same way: the effort clip keeps the
   * physics finite, but last_action (an observation term) and action_rate_l2 would carry it on */
  for (int j = 0; j < NJ; j++) if (!(fabsf(e->last_action[j]) <= H1V2_ACTION_ABS_MAX)) bad = 1;
  *nan_flag = bad;
  /* -- counters, terminations -- */
  e->ep_len += 1;
  int time_out = e->ep_len >= o->max_episode_length;
  int contact = 0;
  double Cmax[NSLOT];
  for (int s = 0; s < NSLOT; s++) {
    Cmax[s] = fmax(e->slot_hist[s][0], fmax(e->slot_hist[s][1], e->slot_hist[s][2]));
    for (int h = 0; h < 3; h++) e->slot_hist_diag[s][h] = e->slot_hist[s][h];
    if (((c->mask_illegal_slots >> s) & 1u) && (float)Cmax[s] > c->contact_threshold) contact = 1;
  }
  if (bad) contact = 1;
  *term_out = (uint8_t)contact;
  *trunc_out = (uint8_t)time_out;
  int reset = contact || time_out;
  /* -- rewards on the pre-reset state -- */
  double R[9], vb[3], wb[3], g[3], heading, ww[3], vw[3], r[NREW];
  root_derived(c, e, R, vb, wb, g, &heading, ww, vw);
  memset(r, 0, sizeof(r));
  double std2 = (double)c->track_std * c->track_std;
  double cy = cos(heading), sy = sin(heading);
  /* yaw-frame velocity: yaw_quat(root_quat)^-1 * v_w */
  double vyaw[2] = {cy * vw[0] + sy * vw[1], -sy * vw[0] + cy * vw[1]};
  double cmdn = sqrt((double)e->cmd[0] * e->cmd[0] + (double)e->cmd[1] * e->cmd[1]);
  int moving = (float)sqrtf(e->cmd[0] * e->cmd[0] + e->cmd[1] * e->cmd[1]) > 0.1f;
  (void)cmdn;
  r[H1V2_REW_TERMINATION] = contact ? 1.0 : 0.0;
  {
    double ex = e->cmd[0] - vyaw[0], ey = e->cmd[1] - vyaw[1];
    r[H1V2_REW_TRACK_LIN_XY_YAW] = exp(-(ex * ex + ey * ey) / std2);
    double ez = e->cmd[2] - ww[2];
    r[H1V2_REW_TRACK_ANG_Z_WORLD] = exp(-(ez * ez) / std2);
    ex = e->cmd[0] - vb[0]; ey = e->cmd[1] - vb[1];
    r[H1V2_REW_TRACK_LIN_XY_BASE] = exp(-(ex * ex + ey * ey) / std2);
    ez = e->cmd[2] - wb[2];
    r[H1V2_REW_TRACK_ANG_Z_BASE] = exp(-(ez * ez) / std2);
  }
  { /* feet_air_time_positive_biped, V/mdp/rewards.py:38-62 */
    int inc[2] = {e->timers[0][2] > 0.0f, e->timers[1][2] > 0.0f};
    float mode[2] = {inc[0] ? e->timers[0][2] : e->timers[0][0], inc[1] ? e->timers[1][2] : e->timers[1][0]};
    int single = (inc[0] + inc[1]) == 1;
    float v0 = single ? mode[0] : 0.0f, v1 = single ? mode[1] : 0.0f;
    float m = v0 < v1 ? v0 : v1;
    if (m > c->feet_air_threshold) m = c->feet_air_threshold;
    r[H1V2_REW_FEET_AIR_BIPED] = moving ? m : 0.0;
    /* feet_air_time, V/mdp/rewards.py:13-35 */
    double s = 0;
    for (int f = 0; f < 2; f++) {
      int first = (e->timers[f][2] > 0.0f) && (e->timers[f][2] < step_dt + 1e-8f);
      s += ((double)e->timers[f][1] - c->feet_air_threshold) * first;
    }
    r[H1V2_REW_FEET_AIR_L2] = moving ? s : 0.0;
  }
  for (int f = 0; f < 2; f++) {
    int in_c = (float)Cmax[f] > c->contact_threshold;
    r[H1V2_REW_FEET_SLIDE] += sqrt(e->foot_vel[f][0] * e->foot_vel[f][0] + e->foot_vel[f][1] * e->foot_vel[f][1]) * in_c;
  }
  for (int j = 0; j < NJ; j++) {
    double q = e->qpos[7 + j], qd = e->qvel[6 + j];
    if ((c->mask_pos_limits >> j) & 1u) {
      double lo = q - o->soft_lo[j], hi = q - o->soft_hi[j];
      r[H1V2_REW_DOF_POS_LIMITS] += -(lo < 0 ? lo : 0) + (hi > 0 ? hi : 0);
    }
    if ((c->mask_joint_dev >> j) & 1u) r[H1V2_REW_JOINT_DEV_HIP] += fabs(q - c->default_joint_pos[j]);
    if ((c->mask_pos_limits_b >> j) & 1u) { /* second joint_pos_limits term of a cfg (C12/rsl_env_cfg.py:380-386) */
      double lo = q - o->soft_lo[j], hi = q - o->soft_hi[j];
      r[H1V2_REW_DOF_POS_LIMITS_B] += -(lo < 0 ? lo : 0) + (hi > 0 ? hi : 0);
    }
    if ((c->mask_joint_dev_b >> j) & 1u) r[H1V2_REW_JOINT_DEV_B] += fabs(q - c->default_joint_pos[j]); /* rsl_env_cfg.py:358-372 */
    if ((c->mask_torques >> j) & 1u) r[H1V2_REW_TORQUES] += e->applied_tau[j] * e->applied_tau[j];
    r[H1V2_REW_DOF_ACC] += e->joint_acc[j] * e->joint_acc[j];
    r[H1V2_REW_JOINT_VEL] += qd * qd;
    double da = (double)e->last_action[j] - prev_action[j];
    r[H1V2_REW_ACTION_RATE] += da * da;
  }
  r[H1V2_REW_ANG_VEL_XY] = wb[0] * wb[0] + wb[1] * wb[1];
  r[H1V2_REW_FLAT_ORI] = g[0] * g[0] + g[1] * g[1];
  r[H1V2_REW_LIN_VEL_Z] = vb[2] * vb[2];
  r[H1V2_REW_BASE_HEIGHT] = (e->qpos[2] - c->base_height_target) * (e->qpos[2] - c->base_height_target);
  for (int s = 0; s < NSLOT; s++) {
    if ((c->mask_undesired_slots >> s) & 1u) r[H1V2_REW_UNDESIRED_CONTACTS] += (float)Cmax[s] > c->contact_threshold;
    if ((c->mask_contact_forces_slots >> s) & 1u) { /* contact_forces: sum_b clip(max_h |F_b| - threshold, min 0) (rsl_env_cfg.py:395-404) */
      double ex = Cmax[s] - c->contact_forces_threshold;
      r[H1V2_REW_CONTACT_FORCES] += ex > 0 ? ex : 0;
    }
  }
  double total = 0;
  for (int t = 0; t < NREW; t++) {
    double v = c->rew_weight[t] == 0.0f ? 0.0 : (double)c->rew_weight[t] * r[t] * step_dt;
    if (bad) v = 0;
    e->rew_terms[t] = v;
    e->ep_sums[t] += v;
    total += v;
  }
  *rew_out = (float)total;
  *reset_flag = reset;
}

/* ------------------------------------------------------------------------------------------ */
int h1v2o_create(const H1v2Config* cfg, int32_t n_envs, uint64_t seed, H1v2Oracle** out) {
  H1v2Oracle* o = (H1v2Oracle*)calloc(1, sizeof(H1v2Oracle));
  o->cfg = *cfg;
  o->n = n_envs;
  o->seed = seed;
  o->env = (OEnv*)calloc((size_t)n_envs, sizeof(OEnv));
  o->max_newton_iters = 100;
  o->nthreads = 1;
  double step_dt = (double)cfg->sim_dt * cfg->decimation;
  o->max_episode_length = (int64_t)ceil((double)cfg->episode_length_s / step_dt * (1.0 - 1e-6)); /* float dt: 0.005f*4 is not 0.02 */
  for (int j = 0; j < NJ; j++) {
    double lo = cfg->joint_range[j][0], hi = cfg->joint_range[j][1];
    double mid = 0.5 * (lo + hi), half = 0.5 * (hi - lo) * cfg->soft_limit_factor;
    o->soft_lo[j] = mid - half; o->soft_hi[j] = mid + half;
  }
  /* startup events: per-env friction and base mass */
  for (int i = 0; i < n_envs; i++) {
    float u[4];
    rng4(seed, cfg->env_id_offset + i, 0, STREAM_EVENT, 0, u);
    o->env[i].friction = (double)uni(u[0], cfg->friction_range[0], cfg->friction_range[1]);
    o->env[i].mass_add = (double)uni(u[1], cfg->mass_add_range[0], cfg->mass_add_range[1]);
  }
  if (cfg->terrain_enable) {
    o->npx = terrain_npx(cfg);
    o->gx = cfg->terrain_rows * o->npx + 1; o->gy = cfg->terrain_cols * o->npx + 1;
    o->terrain = (float*)calloc((size_t)o->gx * o->gy, sizeof(float));
    o->origin_z = (float*)calloc((size_t)cfg->terrain_rows * cfg->terrain_cols, sizeof(float));
    terrain_generate(o);
    /* TerrainImporter._compute_env_origins_curriculum [UPSTREAM]: level = randint(0, max_init_level + 1), type = floor(i / (n / cols)) */
    const int max_init = cfg->terrain_max_init_level < 0 || cfg->terrain_max_init_level > cfg->terrain_rows - 1 ? cfg->terrain_rows - 1 : cfg->terrain_max_init_level;
    for (int i = 0; i < n_envs; i++) {
      float u[4];
      rng4(seed, cfg->env_id_offset + i, 0, STREAM_EVENT, 3, u);
      int lev = (int)(u[0] * (float)(max_init + 1));
      o->env[i].level = lev > max_init ? max_init : lev;
      /* torch.div(arange(n), n / cols, rounding_mode="floor"): the divisor is the fp32 value of the python double, the floor is exact
       * (c10 div_floor_floating works through fmod): 1024 / fp32(204.8) is 4.99999993 -> 4, where a rounded fp32 division says 5 */
      int ty = (int)floor((double)i / (double)(float)((double)n_envs / (double)cfg->terrain_cols));
      o->env[i].type = ty > cfg->terrain_cols - 1 ? cfg->terrain_cols - 1 : ty;
    }
  }
  o->step_counter = 0;
  for (int i = 0; i < n_envs; i++) reset_env(o, i);
  *out = o;
  return 0;
}
void h1v2o_destroy(H1v2Oracle* o) { if (o) { free(o->terrain); free(o->origin_z); free(o->env); free(o); } }
void h1v2o_set_threads(H1v2Oracle* o, int n) { o->nthreads = n > 0 ? n : 1; }
int h1v2o_obs_dim(const H1v2Oracle* o) {
  const H1v2Config* c = &o->cfg;
  if (c->obs_base_lin_vel || c->obs_height_scan)
    return (c->obs_base_lin_vel ? 3 : 0) + H1V2_OBS_TERM_DIM + (c->obs_height_scan ? scan_count(c->scan_size[0], c->scan_resolution) * scan_count(c->scan_size[1], c->scan_resolution) : 0);
  return c->history_length * H1V2_OBS_TERM_DIM;
}
int h1v2o_terrain_dims(const H1v2Oracle* o, int32_t dims[2]) { dims[0] = o->gx; dims[1] = o->gy; return o->terrain ? 0 : -1; }
int h1v2o_get_terrain(const H1v2Oracle* o, float* out) { if (!o->terrain) return -1; memcpy(out, o->terrain, sizeof(float) * (size_t)o->gx * o->gy); return 0; }
int h1v2o_set_terrain(H1v2Oracle* o, const float* in) {
  if (!o->terrain) return -1;
  memcpy(o->terrain, in, sizeof(float) * (size_t)o->gx * o->gy);
  terrain_origin_heights(o);
  return 0;
}
float h1v2o_terrain_level_mean(const H1v2Oracle* o) { return o->terrain_level_mean; }
/* height (relative to the origin of tile (level, type)) and normal under a point given relative to that origin */
int h1v2o_terrain_query(const H1v2Oracle* o, int level, int type, double lx, double ly, double* h, double* n) {
  if (!o->terrain) return -1;
  terrain_query(o, level, type, lx, ly, h, n, NULL);
  return 0;
}
int h1v2o_tri_margin(H1v2Oracle* o, double* out) { for (int i = 0; i < o->n; i++) out[i] = o->env[i].min_tri_margin; return 0; }
int64_t h1v2o_max_episode_length(const H1v2Oracle* o) { return o->max_episode_length; }

int h1v2o_reset(H1v2Oracle* o, const int64_t* env_ids, int32_t n) {
  if (!env_ids) { for (int i = 0; i < o->n; i++) reset_env(o, i); return 0; }
  for (int i = 0; i < n; i++) reset_env(o, (int)env_ids[i]);
  return 0;
}
int h1v2o_observe(H1v2Oracle* o, float* obs) {
  int od = h1v2o_obs_dim(o);
  for (int i = 0; i < o->n; i++) compute_obs(o, i, obs + (size_t)i * od);
  return 0;
}

/* minimal pthread parallel-for (libgomp is not in the image): interleaved static partition over envs */
typedef struct {
  H1v2Oracle* o; const float* actions; float* obs; float* rew; uint8_t* term; uint8_t* trunc;
  int* reset; int* bad; int phase, tid, nthreads; const H1v2Inject* inj;
} Job;
static void post_env(H1v2Oracle* o, int i, const int* reset, float* obs);
static void* job_main(void* p) {
  Job* j = (Job*)p;
  for (int i = j->tid; i < j->o->n; i += j->nthreads) {
    if (j->phase == 0)
      step_env(j->o, i, j->actions + (size_t)i * NJ, j->rew + i, j->term + i, j->trunc + i, j->reset + i, j->bad + i, j->inj);
    else
      post_env(j->o, i, j->reset, j->obs);
  }
  return NULL;
}
static void run_phase(Job proto, int phase) {
  int nt = proto.o->nthreads;
  if (nt > proto.o->n) nt = proto.o->n;
  if (nt < 1) nt = 1;
  pthread_t th[256];
  Job jobs[256];
  if (nt > 256) nt = 256;
  for (int t = 0; t < nt; t++) {
    jobs[t] = proto; jobs[t].phase = phase; jobs[t].tid = t; jobs[t].nthreads = nt;
    if (t > 0) pthread_create(&th[t], NULL, job_main, &jobs[t]);
  }
  job_main(&jobs[0]);
  for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
}

static void post_env(H1v2Oracle* o, int i, const int* reset, float* obs) {
  const int od = h1v2o_obs_dim(o);
  {

    if (reset[i]) reset_env(o, i);
    update_command(o, i);
    /* interval event: push_by_setting_velocity (velocity_env_cfg.py:212-217) */
    if (o->cfg.push_enable) {
      OEnv* e = &o->env[i];
      e->push_left -= o->cfg.sim_dt * (float)o->cfg.decimation;
      if (e->push_left < 1e-6f) {
        float u[4];
        rng4(o->seed, o->cfg.env_id_offset + i, o->step_counter, STREAM_EVENT, 2, u);
        e->push_left = uni(u[2], o->cfg.push_interval_s[0], o->cfg.push_interval_s[1]);
        e->qvel[0] += uni(u[0], o->cfg.push_vel_xy[0], o->cfg.push_vel_xy[1]);
        e->qvel[1] += uni(u[1], o->cfg.push_vel_xy[0], o->cfg.push_vel_xy[1]);
      }
    }
    compute_obs(o, i, obs ? obs + (size_t)i * od : NULL);
    }
}

static int step_impl(H1v2Oracle* o, const float* actions, float* obs, float* rew, uint8_t* terminated, uint8_t* truncated,
                     const H1v2Inject* inj);
int h1v2o_step(H1v2Oracle* o, const float* actions, float* obs, float* rew, uint8_t* terminated, uint8_t* truncated) {
  return step_impl(o, actions, obs, rew, terminated, truncated, NULL);
}
/* one control step whose physics loop is replaced by injected post-physics values (see H1v2Inject) */
int h1v2o_step_injected(H1v2Oracle* o, const float* actions, const float* qpos, const float* qvel, const float* timers,
                        const float* slot_hist, const float* tau, const float* qacc, const float* foot_vel, float* obs, float* rew,
                        uint8_t* terminated, uint8_t* truncated) {
  H1v2Inject inj = {qpos, qvel, timers, slot_hist, tau, qacc, foot_vel};
  return step_impl(o, actions, obs, rew, terminated, truncated, &inj);
}
static int step_impl(H1v2Oracle* o, const float* actions, float* obs, float* rew, uint8_t* terminated, uint8_t* truncated,
                     const H1v2Inject* inj) {
  const int n = o->n;
  o->step_counter += 1;
  int* reset = (int*)calloc((size_t)n, sizeof(int));
  int* bad = (int*)calloc((size_t)n, sizeof(int));
  Job proto = {o, actions, obs, rew, terminated, truncated, reset, bad, 0, 0, 1, inj};
  run_phase(proto, 0);
  /* logging of the envs about to reset (cat_env.py:217-245) */
  double sums[NREW] = {0}, mxy = 0, myaw = 0;
  int cnt = 0, c_to = 0, c_bc = 0, nbad = 0, max_it = 0;
  const double max_len_s = o->cfg.episode_length_s;
  for (int i = 0; i < n; i++) {
    if (o->env[i].newton_iters > max_it) max_it = o->env[i].newton_iters;
    if (!reset[i]) continue;
    cnt++;
    for (int t = 0; t < NREW; t++) sums[t] += o->env[i].ep_sums[t] / max_len_s;
    c_to += truncated[i]; c_bc += terminated[i]; nbad += bad[i];
    mxy += o->env[i].metrics[0]; myaw += o->env[i].metrics[1];
  }
  if (cnt > 0) {
    o->log[H1V2_LOG_COUNT] = (float)cnt;
    for (int t = 0; t < NREW; t++) o->log[H1V2_LOG_REW0 + t] = (float)(sums[t] / cnt);
    o->log[H1V2_LOG_TERM_TIMEOUT] = (float)c_to;
    o->log[H1V2_LOG_TERM_CONTACT] = (float)c_bc;
    o->log[H1V2_LOG_ERR_XY] = (float)(mxy / cnt);
    o->log[H1V2_LOG_ERR_YAW] = (float)(myaw / cnt);
  } else {
    o->log[H1V2_LOG_COUNT] = 0;
  }
  o->log[H1V2_LOG_NAN_RESETS] += (float)nbad;
  o->log[H1V2_LOG_MAX_ITERS] = (float)max_it;
  run_phase(proto, 1);
  if (o->cfg.terrain_enable) { /* Curriculum/terrain_levels: mean level over all envs (curriculums.py:52) */
    double sum = 0;
    for (int i = 0; i < n; i++) sum += o->env[i].level;
    o->terrain_level_mean = (float)(sum / n);
  }
  if (o->cfg.command_class == 1) { /* census for the next step's balancing */
    int cnt = 0;
    for (int i = 0; i < n; i++) cnt += cmd_in_deadzone(o->env[i].cmd, o->cfg.velocity_deadzone);
    o->dz_prev = cnt;
  }
  free(reset);
  free(bad);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
#define CPY_OUT(field, expr, count, type)                                                 \
  if (s->field)                                                                           \
    for (int i = 0; i < o->n; i++)                                                        \
      for (int k = 0; k < (count); k++) s->field[(size_t)i * (count) + k] = (type)(expr);

int h1v2o_get_state(H1v2Oracle* o, const H1v2State* s) {
  const int H = o->cfg.history_length;
  CPY_OUT(root_pos, o->env[i].qpos[k], 3, float)
  CPY_OUT(root_quat, o->env[i].qpos[3 + k], 4, float)
  CPY_OUT(root_lin_vel, o->env[i].qvel[k], 3, float)
  CPY_OUT(root_ang_vel, o->env[i].qvel[3 + k], 3, float)
  CPY_OUT(joint_pos, o->env[i].qpos[7 + k], NJ, float)
  CPY_OUT(joint_vel, o->env[i].qvel[6 + k], NJ, float)
  CPY_OUT(last_action, o->env[i].last_action[k], NJ, float)
  CPY_OUT(target_hist, (k < NJ ? o->env[i].T1[k] : o->env[i].T2[k - NJ]), 2 * NJ, float)
  CPY_OUT(lag, o->env[i].lag, 1, int32_t)
  CPY_OUT(fresh, o->env[i].fresh, 1, int32_t)
  CPY_OUT(command, o->env[i].cmd[k], 3, float)
  CPY_OUT(heading_target, o->env[i].heading_target, 1, float)
  CPY_OUT(time_left, o->env[i].time_left, 1, float)
  CPY_OUT(is_standing, o->env[i].is_standing, 1, int32_t)
  CPY_OUT(is_heading, o->env[i].is_heading, 1, int32_t)
  CPY_OUT(cmd_metrics, o->env[i].metrics[k], 2, float)
  CPY_OUT(feet_timers, o->env[i].timers[k / 4][k % 4], 8, float)
  CPY_OUT(episode_sums, o->env[i].ep_sums[k], NREW, float)
  CPY_OUT(obs_history, o->env[i].hist[k / H1V2_OBS_TERM_DIM][k % H1V2_OBS_TERM_DIM], H * H1V2_OBS_TERM_DIM, float)
  CPY_OUT(friction, o->env[i].friction, 1, float)
  CPY_OUT(mass_add, o->env[i].mass_add, 1, float)
  CPY_OUT(push_time_left, o->env[i].push_left, 1, float)
  CPY_OUT(slot_force, o->env[i].slot_force[k / 3][k % 3], NSLOT * 3, float)
  CPY_OUT(slot_force_hist, o->env[i].slot_hist_diag[k / 3][k % 3], NSLOT * 3, float)
  CPY_OUT(applied_torque, o->env[i].applied_tau[k], NJ, float)
  CPY_OUT(joint_acc, o->env[i].joint_acc[k], NJ, float)
  CPY_OUT(reward_terms, o->env[i].rew_terms[k], NREW, float)
  CPY_OUT(foot_vel, o->env[i].foot_vel[k / 3][k % 3], 6, float)
  CPY_OUT(solver_iters, (k < 2 ? o->env[i].newton_iters : 0), 3, float) /* [2]: contact-list overflow, a kernel-only notion (the oracle's row list is unbounded) */
  CPY_OUT(pre_reset_qpos, 0.0f, 19, float)
  CPY_OUT(pre_reset_qvel, 0.0f, 18, float)
  CPY_OUT(pre_reset_timers, 0.0f, 8, float)
  CPY_OUT(terrain_level, o->env[i].level, 1, int32_t)
  CPY_OUT(terrain_type, o->env[i].type, 1, int32_t)
  return 0;
}

#define CPY_IN(field, lhs, count)                      \
  if (s->field)                                        \
    for (int i = 0; i < o->n; i++)                     \
      for (int k = 0; k < (count); k++) lhs = s->field[(size_t)i * (count) + k];

int h1v2o_set_state(H1v2Oracle* o, const H1v2State* s) {
  const int H = o->cfg.history_length;
  CPY_IN(root_pos, o->env[i].qpos[k], 3)
  CPY_IN(root_quat, o->env[i].qpos[3 + k], 4)
  CPY_IN(root_lin_vel, o->env[i].qvel[k], 3)
  CPY_IN(root_ang_vel, o->env[i].qvel[3 + k], 3)
  CPY_IN(joint_pos, o->env[i].qpos[7 + k], NJ)
  CPY_IN(joint_vel, o->env[i].qvel[6 + k], NJ)
  CPY_IN(last_action, o->env[i].last_action[k], NJ)
  CPY_IN(target_hist, *(k < NJ ? &o->env[i].T1[k] : &o->env[i].T2[k - NJ]), 2 * NJ)
  CPY_IN(lag, o->env[i].lag, 1)
  CPY_IN(fresh, o->env[i].fresh, 1)
  CPY_IN(command, o->env[i].cmd[k], 3)
  CPY_IN(heading_target, o->env[i].heading_target, 1)
  CPY_IN(time_left, o->env[i].time_left, 1)
  CPY_IN(is_standing, o->env[i].is_standing, 1)
  CPY_IN(is_heading, o->env[i].is_heading, 1)
  CPY_IN(cmd_metrics, o->env[i].metrics[k], 2)
  CPY_IN(feet_timers, o->env[i].timers[k / 4][k % 4], 8)
  CPY_IN(episode_sums, o->env[i].ep_sums[k], NREW)
  CPY_IN(obs_history, o->env[i].hist[k / H1V2_OBS_TERM_DIM][k % H1V2_OBS_TERM_DIM], H * H1V2_OBS_TERM_DIM)
  CPY_IN(friction, o->env[i].friction, 1)
  CPY_IN(mass_add, o->env[i].mass_add, 1)
  CPY_IN(push_time_left, o->env[i].push_left, 1)
  CPY_IN(terrain_level, o->env[i].level, 1)
  return 0;
}
int h1v2o_get_episode_length(H1v2Oracle* o, int64_t* out) { for (int i = 0; i < o->n; i++) out[i] = o->env[i].ep_len; return 0; }
int h1v2o_set_episode_length(H1v2Oracle* o, const int64_t* in) { for (int i = 0; i < o->n; i++) o->env[i].ep_len = in[i]; return 0; }
int h1v2o_set_reward_weights(H1v2Oracle* o, const float* w) { for (int t = 0; t < NREW; t++) o->cfg.rew_weight[t] = w[t]; return 0; }
int h1v2o_get_log(H1v2Oracle* o, float* out) { memcpy(out, o->log, sizeof(o->log)); return 0; }
/* smallest distance to a contact / joint-limit activation boundary seen at a substep start during the last step:
 * envs within rounding error of a boundary are excluded from strict float-vs-double comparisons (tests) */
int h1v2o_activation_margin(H1v2Oracle* o, double* contact, double* limit) {
  for (int i = 0; i < o->n; i++) { contact[i] = o->env[i].min_abs_dist; limit[i] = o->env[i].min_limit_dist; }
  return 0;
}
int h1v2o_solver_stats(H1v2Oracle* o, int32_t* iters, double* resid) {
  for (int i = 0; i < o->n; i++) { iters[i] = o->env[i].newton_iters; resid[i] = o->env[i].newton_resid; }
  return 0;
}

/* ---- building blocks exposed for the known-answer tests (tests/test_oracle_*.py) ---- */
void h1v2o_fk(const double* qpos, double* R_out /*[13,9]*/, double* x_out /*[13,3]*/) {
  Kin k;
  kinematics(qpos, &k);
  memcpy(R_out, k.R, sizeof(k.R));
  memcpy(x_out, k.x, sizeof(k.x));
}
void h1v2o_mass_matrix(const H1v2Config* cfg, const double* qpos, double* M) {
  Kin k;
  kinematics(qpos, &k);
  mass_matrix(cfg, &k, 0.0, M);
}
void h1v2o_bias(const H1v2Config* cfg, const double* qpos, const double* qvel, double* bias) {
  Kin k;
  kinematics(qpos, &k);
  rne_bias(cfg, &k, qvel, 0.0, bias, NULL);
}
/* raw physics substep on (qpos,qvel) with joint torques; returns per-slot force too */
void h1v2o_physics_step(const H1v2Config* cfg, double* qpos, double* qvel, const double* ctrl, double friction,
                        double* slot_force /*[6,3] or NULL*/, int32_t* iters, double* resid) {
  H1v2Oracle o;
  memset(&o, 0, sizeof(o));
  o.cfg = *cfg;
  o.cfg.terrain_enable = 0; /* raw plane physics: this helper owns no height field */
  o.max_newton_iters = 100;
  OEnv e;
  memset(&e, 0, sizeof(e));
  memcpy(e.qpos, qpos, sizeof(e.qpos));
  memcpy(e.qvel, qvel, sizeof(e.qvel));
  e.friction = friction;
  physics_substep(&o, &e, ctrl);
  memcpy(qpos, e.qpos, sizeof(e.qpos));
  memcpy(qvel, e.qvel, sizeof(e.qvel));
  if (slot_force) memcpy(slot_force, e.slot_force, sizeof(e.slot_force));
  if (iters) *iters = e.newton_iters;
  if (resid) *resid = e.newton_resid;
}
double h1v2o_total_energy(const H1v2Config* cfg, const double* qpos, const double* qvel) {
  Kin k;
  kinematics(qpos, &k);
  double M[NV * NV];
  H1v2Config c0 = *cfg;
  mass_matrix(&c0, &k, 0.0, M);
  double ke = 0;
  for (int i = 0; i < NV; i++)
    for (int j = 0; j < NV; j++) ke += 0.5 * qvel[i] * M[i * NV + j] * qvel[j];
  double pe = 0;
  for (int b = 0; b < NB; b++) {
    double off[3];
    matvec3(k.R[b], h1v2_body_ipos[b], off);
    pe += h1v2_body_mass[b] * cfg->gravity * (k.x[b][2] + off[2]);
  }
  return ke + pe;
}
float h1v2o_wrap_to_pi(float a) { return wrap_to_pi_f(a); }
