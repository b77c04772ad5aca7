// h1v2_cat.cuh -- the Constraints-as-Terminations tail, run after the fused step kernel.
// Reference: packages/biped_tasks/biped_tasks/utils/cat/constraints.py (the ten constraint bodies, parameters of
// config/h12_12dof/cat_env_cfg.py:336-431), utils/cat/constraint_manager.py:23-86 (CaT: running column maxima, probabilities),
// :185-229 (per-term episode statistics), utils/cat/cat_env.py:147-169 (reward *= 1 - p ; dones = p, 1 on reset).
//
// Unlike the rest of the step this tail couples the envs of one process: every constraint column is normalised by a Polyak
// average of its maximum over ALL envs of the current step, and no_move judges env i on the joints of another env (the
// reference gathers the rows whose command is inside the dead zone and tiles that block over the batch).  So it takes TWO
// launches per control step:
//   step_kernel<true> (csrc/h1v2_step.cuh, with KState::cat set)  computes the raw constraint columns on the pre-reset state it
//                     holds in registers, stores them coalesced [56][N], reduces their maxima (warp shuffle + one atomicMax per
//                     column and warp), keeps the swing-height tracker, flags the dead-zone members (one bit mask per warp); its last block adds the
//                     prefix sums that address them by rank
//   cat_apply_kernel  thread per env: the no_move gather, running maxima, probabilities, p = max, reward *= 1 - p, dones,
//                     per-term episode statistics, log; its last block advances the step parity and clears the maxima
// All step-to-step state (running maxima, their parity, the first-step flag) lives on the device, so a captured CUDA graph of
// the two launches replays correctly.
#pragma once
#include "h1v2_step.cuh"

namespace h1v2 {

struct CatParams {
  int n, epw, nmask;  // envs, envs per warp of the step launch, number of member masks (= its warps)
  float tau, min_p, max_p[H1V2_NUM_CSTR];
};
struct CatState {
  KCat k;         // what the step kernel fills (csrc/h1v2_params.h)
  float* probs;   // [56][N]
  float* rmax;    // [2][56]  running maxima, double-buffered by the parity of ctl[1]
  float* sums;    // [2][10][N] per-term episode sums: violation count, probability
};
__device__ __constant__ const int kCstrCol0[H1V2_NUM_CSTR + 1] = {0, 1, 13, 25, 37, 39, 51, 52, 53, 54, 56};

// term of every column (kCstrCol0 expanded), so that the column loop unrolls with compile-time term indices
__host__ __device__ constexpr int cstr_term_of_col(int c) {
  return c < 1 ? 0 : c < 13 ? 1 : c < 25 ? 2 : c < 37 ? 3 : c < 39 ? 4 : c < 51 ? 5 : c < 52 ? 6 : c < 53 ? 7 : c < 54 ? 8 : 9;
}

__global__ void __launch_bounds__(64) cat_apply_kernel(const CatParams C, const CatState T, float* __restrict__ rew, float* __restrict__ dones) {
  __shared__ float rm_s[H1V2_CSTR_COLS];
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = C.n;
  const KCat& Kc = T.k;
  const int done = __ldcg(Kc.ctl + 1);  // apply launches before this one: parity of the double-buffered running maxima; 0 = none yet
  const int parity = done & 1;
  if (threadIdx.x < H1V2_CSTR_COLS) {  // running maxima of this step, once per block (constraint_manager.py:59-62)
    const int c = threadIdx.x;
    const float cm = __int_as_float(__ldcg(Kc.cmax + c));
    const float rm = done == 0 ? cm : __fadd_rn(__fmul_rn(T.rmax[parity * H1V2_CSTR_COLS + c], C.tau), __fmul_rn(1.f - C.tau, cm));
    rm_s[c] = rm;
    if (blockIdx.x == 0) T.rmax[(parity ^ 1) * H1V2_CSTR_COLS + c] = rm;
  }
  __syncthreads();
  if (env < N) {
    const bool reset = Kc.aux[(size_t)N + env] != 0.f;
    const float inv_len = 1.f / fmaxf(Kc.aux[env], 1.f);
    // Two phases -- every load first, then the arithmetic and the stores: raw / probs / sums may alias as far as the compiler knows,
    // so a load placed after a store waits for it; interleaved, the 56 columns were 56 serialised L2 round trips (25 us at 4096 envs).
    float v[H1V2_CSTR_COLS], sv_[H1V2_NUM_CSTR], sp_[H1V2_NUM_CSTR];
    // no_move (constraints.py:209-231): env i is judged on the joint velocities of dead-zone member i mod K
    const int K = __ldcg(Kc.ctl);
    int src = env;
    if (K > 0) {  // member number r of the ascending dead-zone list: first mask whose inclusive prefix exceeds r, then the bit inside it
      const int r = env % K;
      int lo = 0, hi = C.nmask - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldcg(Kc.list + mid) > r) hi = mid; else lo = mid + 1;
      }
      const unsigned m = (unsigned)__ldcg(Kc.dz + lo);
      const int within = r - (__ldcg(Kc.list + lo) - __popc(m));
      src = lo * C.epw + (int)__fns(m, 0, within + 1);
    }
#pragma unroll
    for (int c = 0; c < H1V2_CSTR_COLS; c++)
      v[c] = (c >= 39 && c < 51) ? (K > 0 ? fabsf(__ldcg(Kc.qd + (size_t)(c - 39) * N + src)) - Kc.no_move_vel_limit : 0.f) : __ldcg(Kc.raw + (size_t)c * N + env);
#pragma unroll
    for (int t = 0; t < H1V2_NUM_CSTR; t++) { sv_[t] = T.sums[(size_t)t * N + env]; sp_[t] = T.sums[(size_t)(H1V2_NUM_CSTR + t) * N + env]; }
    const float rew_in = rew[env];
    float tmax[H1V2_NUM_CSTR];
#pragma unroll
    for (int t = 0; t < H1V2_NUM_CSTR; t++) tmax[t] = 0.f;
#pragma unroll
    for (int c = 0; c < H1V2_CSTR_COLS; c++) {
      const int t = cstr_term_of_col(c);
      float pc = 0.f;
      if (v[c] > 0.f) pc = C.min_p + fminf(fmaxf(__fdiv_rn(v[c], rm_s[c]), 0.f), 1.f) * (C.max_p[t] - C.min_p);  // :70-77
      T.probs[(size_t)c * N + env] = pc;
      if (c >= 39 && c < 51) Kc.raw[(size_t)c * N + env] = v[c];  // kept for h1v2_cat_debug
      tmax[t] = fmaxf(tmax[t], pc);
    }
    float p = 0.f;
#pragma unroll
    for (int t = 0; t < H1V2_NUM_CSTR; t++) {
      p = fmaxf(p, tmax[t]);
      // per-term episode statistics (constraint_manager.py:221-227) and their log on reset (:185-203)
      float sv = sv_[t] + (tmax[t] > 0.f ? 1.f : 0.f);
      float sp = sp_[t] + tmax[t];
      if (reset) {
        atomicAdd(Kc.logacc + t, sv * inv_len * 100.f);
        atomicAdd(Kc.logacc + H1V2_NUM_CSTR + t, sp * inv_len);
        sv = sp = 0.f;
      }
      T.sums[(size_t)t * N + env] = sv;
      T.sums[(size_t)(H1V2_NUM_CSTR + t) * N + env] = sp;
    }
    if (reset) atomicAdd(Kc.logacc + 2 * H1V2_NUM_CSTR, 1.f);
    rew[env] = rew_in * (1.f - p);   // cat_env.py:152
    dones[env] = reset ? 1.f : p;    // :153,167
  }
  // the last block to finish closes the step: parity / first-step flag advance, column maxima back to their floor
  __shared__ unsigned ticket_s;
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); ticket_s = atomicAdd((unsigned*)(Kc.ctl + 2), 1u); }
  __syncthreads();
  if (ticket_s == gridDim.x - 1) {
    if (threadIdx.x < H1V2_CSTR_COLS) Kc.cmax[threadIdx.x] = __float_as_int(1e-6f);  // constraint.max(0).clamp(min=1e-6), constraint_manager.py:56
    if (threadIdx.x == 0) { Kc.ctl[1] = done + 1; Kc.ctl[2] = 0; }
  }
}

}  // namespace h1v2
