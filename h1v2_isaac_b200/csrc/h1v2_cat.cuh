// h1v2_cat.cuh -- the Constraints-as-Terminations tail, run after the fused step kernel.
// Reference: packages/biped_tasks/biped_tasks/utils/cat/constraints.py (the ten constraint bodies, parameters of
// config/h12_12dof/cat_env_cfg.py:336-431), utils/cat/constraint_manager.py:23-86 (CaT: running column maxima, probabilities),
// :185-229 (per-term episode statistics), utils/cat/cat_env.py:147-169 (reward *= 1 - p ; dones = p, 1 on reset).
//
// Unlike the rest of the step this tail couples the envs of one process: every constraint column is normalised by a Polyak
// average of its maximum over ALL envs of the current step, and no_move judges env i on the joints of another env (the
// reference gathers the rows whose command is inside the dead zone and tiles that block over the batch).  So it is four small
// memory-bound launches over the values the step kernel left in its per-env diagnostics rows (csrc/h1v2_params.h):
//   cat_count_kernel, cat_scan_kernel   ordered list of the envs whose whole command is inside the dead zone (two parallel passes
//                     over chunks of 1024 envs); clears the column maxima and the log sums
//   cat_raw_kernel    thread per env: the 56 raw constraint columns, their maxima (atomicMax), the swing-height tracker
//   cat_apply_kernel  thread per env: running maxima, probabilities, p = max, reward *= 1 - p, dones, episode statistics, log
#pragma once
#include "h1v2_step.cuh"

namespace h1v2 {

struct CatParams {
  int n, first;                 // first: no running maximum yet (constraint_manager.py:59-62)
  float tau, min_p, max_p[H1V2_NUM_CSTR];
  uint32_t contact_slots;
  float foot_force_limit, no_move_deadzone, no_move_vel_limit, orientation_limit, height, height_std, clearance_min_height, clearance_deadzone;
  float step_dt, vel_limit;
};
struct CatState {
  float* raw;     // [56][N] raw constraint columns of the last step
  float* probs;   // [56][N]
  float* rmax;    // [2][56]  running maxima, double-buffered by step parity
  int* cmax;      // [56]     this step's column maxima as float bits (all candidates are positive: clamp at 1e-6)
  int* list;      // [N]      envs whose command is inside the dead zone, ascending
  int* count;     // [1]
  int* chunk_count;  // [ceil(N / 1024)] dead-zone members per chunk of 1024 envs
  float* swing;   // [2][N]   swing_max_height of foot_clearance
  float* sums;    // [2][10][N] per-term episode sums: violation count, probability
  float* logacc;  // [21]     sums over the envs reset in this step: violation[10], probability[10], count
};
__device__ __constant__ const int kCstrCol0[H1V2_NUM_CSTR + 1] = {0, 1, 13, 25, 37, 39, 51, 52, 53, 54, 56};

__device__ __forceinline__ bool cmd_all_inside(const float* dg, float dz) { return fabsf(dg[141]) < dz && fabsf(dg[142]) < dz && fabsf(dg[143]) < dz; }

// Ordered compaction in two parallel passes over chunks of 1024 envs: per-chunk counts, then every chunk places its own members
// after the sum of the counts before it (the reference's boolean-mask gather keeps ascending env order, constraints.py:216-222).
__global__ void __launch_bounds__(1024) cat_count_kernel(const float* __restrict__ diag, const CatParams C, const CatState T) {
  const int tid = threadIdx.x, i = blockIdx.x * 1024 + tid;
  if (blockIdx.x == 0) {
    if (tid < H1V2_CSTR_COLS) T.cmax[tid] = __float_as_int(1e-6f);  // constraint.max(0).clamp(min=1e-6), constraint_manager.py:56
    if (tid < 2 * H1V2_NUM_CSTR + 1) T.logacc[tid] = 0.f;
  }
  const bool f = i < C.n && cmd_all_inside(diag + (size_t)i * H1V2_DIAG_DIM, C.no_move_deadzone);
  const int cnt = __syncthreads_count(f);
  if (tid == 0) T.chunk_count[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(1024) cat_scan_kernel(const float* __restrict__ diag, const CatParams C, const CatState T) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, i = blockIdx.x * 1024 + tid;
  if (w == 0) {  // sum of the counts of the chunks before this one (and, in the last chunk, the total)
    int b = 0;
    for (int k = lane; k < (int)blockIdx.x; k += 32) b += T.chunk_count[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (lane == 0) {
      base_s = b;
      if (blockIdx.x == gridDim.x - 1) *T.count = b + T.chunk_count[blockIdx.x];
    }
  }
  const bool f = i < C.n && cmd_all_inside(diag + (size_t)i * H1V2_DIAG_DIM, C.no_move_deadzone);
  const unsigned m = __ballot_sync(0xffffffffu, f);
  if (lane == 0) warp_tot[w] = __popc(m);
  __syncthreads();
  int off = 0;
  for (int k = 0; k < w; k++) off += warp_tot[k];
  if (f) T.list[base_s + off + __popc(m & ((1u << lane) - 1u))] = i;
}

// height of the ankle_roll_link origin above the ground (body_link_pos_w z, constraints.py:283)
__device__ __forceinline__ float foot_height(const KLeg& LG, const M3& R0, float root_z, const float* q) {
  M3 R = R0;
  V3 x = mk3(0.f, 0.f, 0.f);
#pragma unroll 1
  for (int i = 0; i < 6; i++) {
    x = x + mulv(R, ld3(LG.pos[i]));
    real s_, c_;
    sincos_lim(q[i], s_, c_);
    rotate_rt(R, joint_axis(i), s_, c_);
  }
  return root_z + x.z;
}

__global__ void __launch_bounds__(64) cat_raw_kernel(const __grid_constant__ KParams P, const float* __restrict__ diag, const CatParams C, const CatState T) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = env < C.n;
  const float* dg = diag + (size_t)(valid ? env : 0) * H1V2_DIAG_DIM;
  const int N = C.n;
  float col[H1V2_CSTR_COLS];
  // slot histories: dg[18 + 3 s + h]
  float cm[6];
#pragma unroll
  for (int s = 0; s < 6; s++) cm[s] = fmaxf(dg[18 + 3 * s], fmaxf(dg[19 + 3 * s], dg[20 + 3 * s]));
  bool any = false;
#pragma unroll
  for (int s = 0; s < 6; s++) any |= ((C.contact_slots >> s) & 1u) && cm[s] > 1.0f;
  col[0] = any ? 1.f : 0.f;
  const int K = *T.count;
  const float* src = K > 0 ? diag + (size_t)T.list[(valid ? env : 0) % K] * H1V2_DIAG_DIM : dg;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    const float q = dg[103 + j], qd = dg[121 + j], tau = dg[36 + j];
    col[1 + j] = fmaxf(P.soft_lo[j] - q, q - P.soft_hi[j]);
    col[13 + j] = fabsf(qd) - C.vel_limit;
    col[25 + j] = fabsf(tau) - P.effort[j];
    col[39 + j] = K > 0 ? fabsf(src[121 + j]) - C.no_move_vel_limit : 0.f;
  }
  col[37] = cm[0] - C.foot_force_limit;
  col[38] = cm[1] - C.foot_force_limit;
  const M3 R0 = quat2mat(dg[99], dg[100], dg[101], dg[102]);
  const float gx = -R0.cx.z, gy = -R0.cy.z;  // projected gravity R^T (0,0,-1)
  col[51] = sqrtf(gx * gx + gy * gy) - C.orientation_limit;
  const float z = dg[98];
  col[52] = (z < C.height - C.height_std || z > C.height + C.height_std) ? 1.f : 0.f;
  const int nfeet = (cm[0] > 1.0f) + (cm[1] > 1.0f);
  col[53] = (nfeet < 1 || nfeet > 2) ? 1.f : 0.f;
  const float dz = C.clearance_deadzone;
  const float active = (fabsf(dg[141]) > dz || fabsf(dg[142]) > dz || fabsf(dg[143]) > dz) ? 1.f : 0.f;
#pragma unroll
  for (int f = 0; f < 2; f++) {
    const float cct = dg[133 + 4 * f + 2];
    const bool touchdown = cct > 0.f && cct < C.step_dt + 1.0e-8f;
    const float sw = valid ? T.swing[(size_t)f * N + env] : 0.f;
    col[54 + f] = (C.clearance_min_height - sw) * (touchdown ? 1.f : 0.f) * active;
    const float fz = foot_height(P.leg[f], R0, z, dg + 103 + 6 * f);
    if (valid) T.swing[(size_t)f * N + env] = touchdown ? 0.f : fmaxf(sw, fz);
  }
#pragma unroll
  for (int c = 0; c < H1V2_CSTR_COLS; c++) {
    if (valid) T.raw[(size_t)c * N + env] = col[c];
    float v = valid ? col[c] : -3.0e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 1e-6f) atomicMax(T.cmax + c, __float_as_int(v));
  }
}

// term of every column (kCstrCol0 expanded), so that the column loop unrolls with compile-time term indices
__host__ __device__ constexpr int cstr_term_of_col(int c) {
  return c < 1 ? 0 : c < 13 ? 1 : c < 25 ? 2 : c < 37 ? 3 : c < 39 ? 4 : c < 51 ? 5 : c < 52 ? 6 : c < 53 ? 7 : c < 54 ? 8 : 9;
}

__global__ void __launch_bounds__(64) cat_apply_kernel(const float* __restrict__ diag, const CatParams C, const CatState T, int parity,
                                                       float* __restrict__ rew, float* __restrict__ dones) {
  __shared__ float rm_s[H1V2_CSTR_COLS];
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = C.n;
  if (threadIdx.x < H1V2_CSTR_COLS) {  // running maxima of this step, once per block (constraint_manager.py:59-62)
    const int c = threadIdx.x;
    const float cm = __int_as_float(T.cmax[c]);
    const float rm = C.first ? cm : __fadd_rn(__fmul_rn(T.rmax[parity * H1V2_CSTR_COLS + c], C.tau), __fmul_rn(1.f - C.tau, cm));
    rm_s[c] = rm;
    if (blockIdx.x == 0) T.rmax[(parity ^ 1) * H1V2_CSTR_COLS + c] = rm;
  }
  __syncthreads();
  if (env >= N) return;
  const float* dg = diag + (size_t)env * H1V2_DIAG_DIM;
  const bool reset = dg[167] != 0.f;
  const float inv_len = 1.f / fmaxf(dg[166], 1.f);
  float tmax[H1V2_NUM_CSTR];
#pragma unroll
  for (int t = 0; t < H1V2_NUM_CSTR; t++) tmax[t] = 0.f;
#pragma unroll
  for (int c = 0; c < H1V2_CSTR_COLS; c++) {  // independent loads: the compiler keeps them all in flight
    const int t = cstr_term_of_col(c);
    const float v = T.raw[(size_t)c * N + env];
    float pc = 0.f;
    if (v > 0.f) pc = C.min_p + fminf(fmaxf(__fdiv_rn(v, rm_s[c]), 0.f), 1.f) * (C.max_p[t] - C.min_p);  // :70-77
    T.probs[(size_t)c * N + env] = pc;
    tmax[t] = fmaxf(tmax[t], pc);
  }
  float p = 0.f;
#pragma unroll
  for (int t = 0; t < H1V2_NUM_CSTR; t++) {
    p = fmaxf(p, tmax[t]);
    // per-term episode statistics (constraint_manager.py:221-227) and their log on reset (:185-203)
    float sv = T.sums[(size_t)t * N + env] + (tmax[t] > 0.f ? 1.f : 0.f);
    float sp = T.sums[(size_t)(H1V2_NUM_CSTR + t) * N + env] + tmax[t];
    if (reset) {
      atomicAdd(T.logacc + t, sv * inv_len * 100.f);
      atomicAdd(T.logacc + H1V2_NUM_CSTR + t, sp * inv_len);
      sv = sp = 0.f;
    }
    T.sums[(size_t)t * N + env] = sv;
    T.sums[(size_t)(H1V2_NUM_CSTR + t) * N + env] = sp;
  }
  if (reset) atomicAdd(T.logacc + 2 * H1V2_NUM_CSTR, 1.f);
  rew[env] *= 1.f - p;             // cat_env.py:152
  dones[env] = reset ? 1.f : p;    // :153,167
}

}  // namespace h1v2
