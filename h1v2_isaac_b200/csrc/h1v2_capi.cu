// h1v2_capi.cu -- host side of the C-ABI declared in include/h1v2_b200.h.
// Owns the device state, builds the kernel parameter block from H1v2Config + the compiled model tables,
// launches the fused step kernel.  No torch types, no synchronisation except where the header says so.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#define CPU_PAUSE() _mm_pause()
#else
#define CPU_PAUSE() ((void)0)
#endif

#include "../../include/h1v2_model_h12.h"
#include "h1v2_step.cuh"
#include "h1v2_cat.cuh"

using namespace h1v2;

static thread_local std::string g_err;
static int fail(const std::string& m) {
  g_err = m;
  return -1;
}
#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));           \
  } while (0)
// Every entry point runs on the handle's device whatever the caller's current device is, and leaves the caller's current
// device as it found it (a process may hold handles on several GPUs, e.g. one trainer thread per GPU).
struct DeviceGuard {
  int prev = -1;
  bool changed = false, ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev) { ok = cudaSetDevice(dev) == cudaSuccess; changed = ok; }
  }
  ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};
// inside h1v2_create, once the handle exists: release it on failure
#define CKH(call)                                                                                     \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      fail(std::string(#call) + ": " + cudaGetErrorString(e_));                                       \
      h1v2_destroy(h);                                                                                \
      return -1;                                                                                      \
    }                                                                                                 \
  } while (0)

struct H1v2Handle {
  H1v2Config cfg;
  KParams P;
  KState S;
  int n = 0, device = 0;
  uint64_t seed = 0;
  int64_t launches = 0;
  bool attr_set = false;
  bool no_quad = false;             // cfg.reserved[3] != 0: 8-env warps run the plain instantiation (A/B switch of the mirror-lane split)
  std::vector<void*> allocs;       // base pointers (guard zone first)
  std::vector<size_t> alloc_bytes;  // payload bytes between the guard zones
  int64_t* own_ep_len = nullptr;
  // staging for h1v2_step_host
  float *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr;
  uint8_t *d_term = nullptr, *d_trunc = nullptr;
  cudaStream_t host_stream = nullptr;
  // ordering of the host path after work queued through the stream-taking entry points
  cudaStream_t last_stream = nullptr;
  bool last_stream_set = false;
  cudaEvent_t order_ev = nullptr;
  // host-assembly path of h1v2_step_host (see HostPool below)
  float* h_sample = nullptr;      // pinned + mapped [N][48]: the step's new sample, written by the kernel (zero-copy)
  float* h_sample_dev = nullptr;  // its device alias
  float* h_ring = nullptr;        // host mirror of the history ring [N][H][48]
  bool ring_valid = false;        // the mirror equals the device ring
  uint64_t hist_launches = 0;     // step / observe launches so far == device counters[1] (the history head)
  struct HostPool* pool = nullptr;
  int host_mode = -1;             // -1 undecided, 0 full rows over PCIe (zero-copy / staged), 1 samples + host assembly (of the envs >= host_rows)
  int host_rows = 0;              // mode 1: envs [0, host_rows) still get their rows from the kernel over PCIe (a multiple of epw) -- the hybrid
  cudaEvent_t host_ev = nullptr;   // end of the last streaming host step: the stream-taking entry points wait for it (order_after_host)
  bool host_pending = false;
  unsigned* h_flags = nullptr;     // pinned + mapped [blocks]: per-warp completion words of the step in flight (written by the kernel)
  unsigned* h_flags_dev = nullptr;
  unsigned host_seq = 0;
  int host_calib = -1;            // >= 0: calls made so far while the candidates are being timed on this host (see step_host_impl)
  double host_calib_t[5] = {1e30, 1e30, 1e30, 1e30, 1e30};
  double host_calib_s[5][5] = {};  // the five timed calls of every candidate
  double host_chosen_t = 0.0, host_ema = 0.0;  // watchdog: calibrated time of the chosen candidate, running mean of the calls since
  int host_since = 0, host_recal = 0;          // calls since the calibration ended; re-calibrations so far (at most two)
  // Constraints-as-Terminations tail (cfg.cat_enable)
  CatState cat = {};            // all step-to-step CaT state lives on the device (graph-replayable)
  uint8_t* cat_term = nullptr;  // scratch for the step kernel's terminated flags (h1v2_cat_step reports dones instead)
  float* d_dones = nullptr;     // staging of h1v2_cat_step_host
  float cat_log[2 * H1V2_NUM_CSTR + 1] = {};
  // Rough id: height field and tile origins (device copies in S.terrain_h / S.terrain_oz)
  bool rough = false;
  float* d_terrain = nullptr;
  float* d_origin_z = nullptr;
};

// Every stream-taking entry point notes its stream: h1v2_step_host runs on a private non-blocking stream and orders itself
// after that work at entry (a C caller may do h1v2_reset(h, ids, n, NULL) and then h1v2_step_host).
static inline void note_stream(H1v2Handle* h, cudaStream_t st) { h->last_stream = st; h->last_stream_set = true; }
// ... and the other way round: a host step in streaming mode returns when all its outputs are in the caller's memory and the last block
// has published the launch's bookkeeping; what is left on the private stream is the kernel's exit.  Work queued on another stream
// afterwards waits for the event recorded behind that launch.  The caller's stream may be capturing a CUDA graph: there the wait
// has to be an external-event node (cudaEventWaitExternal) -- a plain wait on an event recorded outside the capture, or any host-side
// synchronisation, invalidates the capture.
static inline void order_after_host(H1v2Handle* h, cudaStream_t st) {
  if (!h->host_pending || st == h->host_stream) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  cudaStreamWaitEvent(st, h->host_ev, cs == cudaStreamCaptureStatusActive ? cudaEventWaitExternal : cudaEventWaitDefault);
}

// ------------------------------------------------------------------------------------------------------
// auxiliary kernels (not on the step path)
// ------------------------------------------------------------------------------------------------------
__global__ void reset_kernel(const __grid_constant__ KParams P, const KState S, const int64_t* __restrict__ ids, int n_ids,
                             float* cat_sums) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = gtid >> 1, side = gtid & 1;
  if (i >= n_ids) return;
  const int env = ids ? (int)ids[i] : i;
  if (env < 0 || env >= P.n) return;
  const int N = P.n, N2 = 2 * P.n, lidx = 2 * env + side;
  const int64_t gid = P.env_id_offset + env;
  const unsigned long long step = S.counters[0];
  float4 r3 = S.root[3 * N + env];
  real mu = r3.y, mass_add = r3.z, push_left = r3.w;
  real rp[3], rq[4], rv[3], rw[3], q[6], qd[6], la[6], T1[6], T2[6];
  float4 tm;
  CmdState cmd;
  cmd.flags = 0; cmd.heading_target = 0.f;
  if (P.rough && side == 0) {  // (side 1's command state is discarded) the env keeps its terrain tile; _reset_idx runs the terrain-level curriculum on the state being left (curriculums.py:21-52)
    const float4 r0 = S.root[env], c0 = S.cmd[env];
    cmd.flags = __float_as_int(S.cmd[N + env].w) & FLAG_TERRAIN_MASK;
    cmd.flags = terrain_curriculum(P, cmd.flags, r0.x, r0.y, c0.x, c0.y, gid, step);
  }
  reset_env(P, side, gid, step, rp, rq, rv, rw, q, qd, la, T1, T2, tm, cmd, push_left);
  if (side == 0) {
    S.root[env] = make_float4(rp[0], rp[1], rp[2], rq[0]);
    S.root[N + env] = make_float4(rq[1], rq[2], rq[3], rv[0]);
    S.root[2 * N + env] = make_float4(rv[1], rv[2], rw[0], rw[1]);
    S.root[3 * N + env] = make_float4(rw[2], mu, mass_add, push_left);
    S.cmd[env] = make_float4(cmd.c[0], cmd.c[1], cmd.c[2], cmd.heading_target);
    S.cmd[N + env] = make_float4(cmd.time_left, 0.f, 0.f, __int_as_float(cmd.flags));
    S.ep_len[env] = 0;
    for (int k = 0; k < H1V2_EPSUM_F4; k++) S.epsum[(size_t)k * N + env] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cat_sums)  // _reset_idx -> constraint_manager.reset(env_ids) (constraint_manager.py:213-214): the per-term episode sums restart
      for (int k = 0; k < 2 * H1V2_NUM_CSTR; k++) cat_sums[(size_t)k * N + env] = 0.f;
  }
  S.leg[lidx] = make_float4(q[0], q[1], q[2], q[3]);
  S.leg[N2 + lidx] = make_float4(q[4], q[5], qd[0], qd[1]);
  S.leg[2 * N2 + lidx] = make_float4(qd[2], qd[3], qd[4], qd[5]);
  for (int k = 0; k < 5; k++) S.act[(size_t)k * N2 + lidx] = make_float4(0.f, 0.f, 0.f, 0.f);
  S.timers[lidx] = tm;
  for (int k = 0; k < 3; k++) S.warm[(size_t)k * N2 + lidx] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void startup_kernel(const __grid_constant__ KParams P, const KState S, float fr_lo, float fr_hi, float ma_lo, float ma_hi, int max_init) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= P.n) return;
  float u[4];
  rng4(P.key0, P.env_id_offset + env, 0ull, STREAM_EVENT, 0, u);
  S.root[3 * P.n + env] = make_float4(0.f, uni(u[0], fr_lo, fr_hi), uni(u[1], ma_lo, ma_hi), 0.f);
  if (P.rough) {
    // TerrainImporter._compute_env_origins_curriculum [UPSTREAM]: level = randint(0, max_init_level + 1), type = floor(i / (n / cols))
    float v[4];
    rng4(P.key0, P.env_id_offset + env, 0ull, STREAM_EVENT, 3, v);
    const int level = min((int)__fmul_rn(v[0], (float)(max_init + 1)), max_init);
    // torch.div(arange(n), n / cols, rounding_mode="floor"): fp32 divisor, exact floor (1024 / fp32(204.8) = 4.99999993 -> 4)
    const int type = min((int)floor((double)env / (double)(float)((double)P.n / (double)P.t_cols)), P.t_cols - 1);
    S.cmd[P.n + env] = make_float4(0.f, 0.f, 0.f, __int_as_float((level << FLAG_LEVEL_SHIFT) | (type << FLAG_TYPE_SHIFT)));
  }
}

__global__ void random_actions_kernel(const __grid_constant__ KParams P, float* __restrict__ actions, unsigned long long step) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= P.n) return;
  const int64_t gid = P.env_id_offset + env;
#pragma unroll
  for (int b = 0; b < 3; b++) {
    float u[4], z[4];
    rng4(P.key0, gid, step, STREAM_ACTIONS, b, u);
    // Box-Muller on (u0,u1) and (u2,u3); 1-u keeps the log argument in (0,1]
    float r0 = sqrtf(-2.f * logf(1.f - u[0])), r1 = sqrtf(-2.f * logf(1.f - u[2]));
    float s0, c0, s1, c1;
    sincospif(2.f * u[1], &s0, &c0);
    sincospif(2.f * u[3], &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
#pragma unroll
    for (int k = 0; k < 4; k++) actions[(size_t)env * 12 + 4 * b + k] = z[k];
  }
}

// FP32 FMA peak micro-benchmark (denominator of the FP32 roofline; MEASURED_PEAKS.json has no FP32 entry)
__global__ void fma_peak_kernel(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 0.999f, c = 1e-3f;
  for (int i = 0; i < iters; i++) {
    a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// natural-layout state exchange; one thread per env
__global__ void state_io_kernel(const __grid_constant__ KParams P, const KState S, const H1v2State st, int set) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= P.n) return;
  const int N = P.n, N2 = 2 * P.n, H = P.H;
  float4 r[4], c[2];
  for (int k = 0; k < 4; k++) r[k] = S.root[(size_t)k * N + env];
  for (int k = 0; k < 2; k++) c[k] = S.cmd[(size_t)k * N + env];
  int flags = __float_as_int(c[1].w);
  float root[16] = {r[0].x, r[0].y, r[0].z, r[0].w, r[1].x, r[1].y, r[1].z, r[1].w, r[2].x, r[2].y, r[2].z, r[2].w, r[3].x, r[3].y, r[3].z, r[3].w};
  float leg[2][12], act[2][20], tmr[2][4], es[4 * H1V2_EPSUM_F4];
  for (int s = 0; s < 2; s++) {
    const int l = 2 * env + s;
    for (int k = 0; k < 3; k++) { float4 v = S.leg[(size_t)k * N2 + l]; leg[s][4 * k] = v.x; leg[s][4 * k + 1] = v.y; leg[s][4 * k + 2] = v.z; leg[s][4 * k + 3] = v.w; }
    for (int k = 0; k < 5; k++) { float4 v = S.act[(size_t)k * N2 + l]; act[s][4 * k] = v.x; act[s][4 * k + 1] = v.y; act[s][4 * k + 2] = v.z; act[s][4 * k + 3] = v.w; }
    float4 t = S.timers[l]; tmr[s][0] = t.x; tmr[s][1] = t.y; tmr[s][2] = t.z; tmr[s][3] = t.w;
  }
  for (int k = 0; k < H1V2_EPSUM_F4; k++) { float4 v = S.epsum[(size_t)k * N + env]; es[4 * k] = v.x; es[4 * k + 1] = v.y; es[4 * k + 2] = v.z; es[4 * k + 3] = v.w; }
  const int head = (int)(S.counters[1] % (unsigned long long)H);
  if (!set) {
    if (st.root_pos) for (int k = 0; k < 3; k++) st.root_pos[env * 3 + k] = root[k];
    if (st.root_quat) for (int k = 0; k < 4; k++) st.root_quat[env * 4 + k] = root[3 + k];
    if (st.root_lin_vel) for (int k = 0; k < 3; k++) st.root_lin_vel[env * 3 + k] = root[7 + k];
    if (st.root_ang_vel) for (int k = 0; k < 3; k++) st.root_ang_vel[env * 3 + k] = root[10 + k];
    for (int s = 0; s < 2; s++)
      for (int k = 0; k < 6; k++) {
        const int j = 6 * s + k;
        if (st.joint_pos) st.joint_pos[env * 12 + j] = leg[s][k];
        if (st.joint_vel) st.joint_vel[env * 12 + j] = leg[s][6 + k];
        if (st.last_action) st.last_action[env * 12 + P.inv_perm[j]] = act[s][k];
        if (st.target_hist) { st.target_hist[env * 24 + j] = act[s][6 + k]; st.target_hist[env * 24 + 12 + j] = act[s][12 + k]; }
      }
    if (st.lag) st.lag[env] = (flags >> FLAG_LAG_SHIFT) & 7;
    if (st.fresh) st.fresh[env] = flags & 3;
    if (st.command) { st.command[env * 3] = c[0].x; st.command[env * 3 + 1] = c[0].y; st.command[env * 3 + 2] = c[0].z; }
    if (st.heading_target) st.heading_target[env] = c[0].w;
    if (st.time_left) st.time_left[env] = c[1].x;
    if (st.is_standing) st.is_standing[env] = (flags & FLAG_STANDING) != 0;
    if (st.is_heading) st.is_heading[env] = (flags & FLAG_HEADING) != 0;
    if (st.cmd_metrics) { st.cmd_metrics[env * 2] = c[1].y; st.cmd_metrics[env * 2 + 1] = c[1].z; }
    if (st.feet_timers) for (int s = 0; s < 2; s++) for (int k = 0; k < 4; k++) st.feet_timers[env * 8 + 4 * s + k] = tmr[s][k];
    if (st.episode_sums) for (int k = 0; k < H1V2_NUM_REW; k++) st.episode_sums[env * H1V2_NUM_REW + k] = es[k];
    if (st.obs_history)
      for (int hh = 0; hh < H; hh++) {
        int sl = (head + 1 + hh) % H;
        for (int k = 0; k < 45; k++) st.obs_history[((size_t)env * H + hh) * 45 + k] = S.hist[((size_t)env * H + sl) * H1V2_HIST_STRIDE + k];
      }
    if (st.friction) st.friction[env] = root[13];
    if (st.mass_add) st.mass_add[env] = root[14];
    if (st.push_time_left) st.push_time_left[env] = root[15];
    if (st.terrain_level) st.terrain_level[env] = (flags >> FLAG_LEVEL_SHIFT) & 255;
    if (st.terrain_type) st.terrain_type[env] = (flags >> FLAG_TYPE_SHIFT) & 255;
    if (S.diag) {
      const float* dg = S.diag + (size_t)env * H1V2_DIAG_DIM;
      if (st.slot_force) for (int k = 0; k < 18; k++) st.slot_force[env * 18 + k] = dg[k];
      if (st.slot_force_hist) for (int k = 0; k < 18; k++) st.slot_force_hist[env * 18 + k] = dg[18 + k];
      if (st.applied_torque) for (int k = 0; k < 12; k++) st.applied_torque[env * 12 + k] = dg[36 + k];
      if (st.joint_acc) for (int k = 0; k < 12; k++) st.joint_acc[env * 12 + k] = dg[48 + k];
      if (st.reward_terms) for (int k = 0; k < H1V2_NUM_REW; k++) st.reward_terms[env * H1V2_NUM_REW + k] = dg[H1V2_DIAG_REW0 + k];
      if (st.foot_vel) for (int k = 0; k < 6; k++) st.foot_vel[env * 6 + k] = dg[80 + k];
      if (st.pre_reset_qpos) for (int k = 0; k < 19; k++) st.pre_reset_qpos[env * 19 + k] = dg[96 + k];
      if (st.pre_reset_qvel) for (int k = 0; k < 18; k++) st.pre_reset_qvel[env * 18 + k] = dg[115 + k];
      if (st.pre_reset_timers) for (int k = 0; k < 8; k++) st.pre_reset_timers[env * 8 + k] = dg[133 + k];
      if (st.solver_iters) { st.solver_iters[env * 3] = dg[86]; st.solver_iters[env * 3 + 1] = dg[88]; st.solver_iters[env * 3 + 2] = dg[89]; }
    }
    return;
  }
  // ---- set ----
  if (st.root_pos) for (int k = 0; k < 3; k++) root[k] = st.root_pos[env * 3 + k];
  if (st.root_quat) for (int k = 0; k < 4; k++) root[3 + k] = st.root_quat[env * 4 + k];
  if (st.root_lin_vel) for (int k = 0; k < 3; k++) root[7 + k] = st.root_lin_vel[env * 3 + k];
  if (st.root_ang_vel) for (int k = 0; k < 3; k++) root[10 + k] = st.root_ang_vel[env * 3 + k];
  if (st.friction) root[13] = st.friction[env];
  if (st.mass_add) root[14] = st.mass_add[env];
  if (st.push_time_left) root[15] = st.push_time_left[env];
  for (int s = 0; s < 2; s++)
    for (int k = 0; k < 6; k++) {
      const int j = 6 * s + k;
      if (st.joint_pos) leg[s][k] = st.joint_pos[env * 12 + j];
      if (st.joint_vel) leg[s][6 + k] = st.joint_vel[env * 12 + j];
      if (st.last_action) act[s][k] = st.last_action[env * 12 + P.inv_perm[j]];
      if (st.target_hist) { act[s][6 + k] = st.target_hist[env * 24 + j]; act[s][12 + k] = st.target_hist[env * 24 + 12 + j]; }
    }
  if (st.lag) flags = (flags & ~(7 << FLAG_LAG_SHIFT)) | ((st.lag[env] & 7) << FLAG_LAG_SHIFT);
  if (st.fresh) flags = (flags & ~3) | (st.fresh[env] & 3);
  if (st.is_standing) flags = (flags & ~FLAG_STANDING) | (st.is_standing[env] ? FLAG_STANDING : 0);
  if (st.is_heading) flags = (flags & ~FLAG_HEADING) | (st.is_heading[env] ? FLAG_HEADING : 0);
  if (st.terrain_level && P.rough) flags = (flags & ~(255 << FLAG_LEVEL_SHIFT)) | (min(max(st.terrain_level[env], 0), P.t_rows - 1) << FLAG_LEVEL_SHIFT);
  if (st.command) { c[0].x = st.command[env * 3]; c[0].y = st.command[env * 3 + 1]; c[0].z = st.command[env * 3 + 2]; }
  if (st.heading_target) c[0].w = st.heading_target[env];
  if (st.time_left) c[1].x = st.time_left[env];
  if (st.cmd_metrics) { c[1].y = st.cmd_metrics[env * 2]; c[1].z = st.cmd_metrics[env * 2 + 1]; }
  c[1].w = __int_as_float(flags);
  if (st.feet_timers) for (int s = 0; s < 2; s++) for (int k = 0; k < 4; k++) tmr[s][k] = st.feet_timers[env * 8 + 4 * s + k];
  if (st.episode_sums) for (int k = 0; k < H1V2_NUM_REW; k++) es[k] = st.episode_sums[env * H1V2_NUM_REW + k];
  if (st.obs_history)
    for (int hh = 0; hh < H; hh++) {
      int sl = (head + 1 + hh) % H;
      for (int k = 0; k < 45; k++) S.hist[((size_t)env * H + sl) * H1V2_HIST_STRIDE + k] = st.obs_history[((size_t)env * H + hh) * 45 + k];
    }
  for (int k = 0; k < 4; k++) S.root[(size_t)k * N + env] = make_float4(root[4 * k], root[4 * k + 1], root[4 * k + 2], root[4 * k + 3]);
  for (int k = 0; k < 2; k++) S.cmd[(size_t)k * N + env] = c[k];
  for (int s = 0; s < 2; s++) {
    const int l = 2 * env + s;
    for (int k = 0; k < 3; k++) S.leg[(size_t)k * N2 + l] = make_float4(leg[s][4 * k], leg[s][4 * k + 1], leg[s][4 * k + 2], leg[s][4 * k + 3]);
    for (int k = 0; k < 5; k++) S.act[(size_t)k * N2 + l] = make_float4(act[s][4 * k], act[s][4 * k + 1], act[s][4 * k + 2], act[s][4 * k + 3]);
    S.timers[l] = make_float4(tmr[s][0], tmr[s][1], tmr[s][2], tmr[s][3]);
  }
  for (int k = 0; k < H1V2_EPSUM_F4; k++) S.epsum[(size_t)k * N + env] = make_float4(es[4 * k], es[4 * k + 1], es[4 * k + 2], es[4 * k + 3]);
}

// ------------------------------------------------------------------------------------------------------
// parameter block
// ------------------------------------------------------------------------------------------------------
static double impedance_h(const float* solimp, double pos) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  auto cl = [](double v) { return v < 1e-4 ? 1e-4 : (v > 0.9999 ? 0.9999 : v); };
  dmin = cl(dmin); dmax = cl(dmax); mid = cl(mid);
  if (power < 1) power = 1;
  if (dmin == dmax || width <= 1e-15) return 0.5 * (dmin + dmax);
  double x = std::fabs(pos) / width;
  if (x >= 1) return dmax;
  if (x == 0) return dmin;
  double y = power == 1 ? x : (x <= mid ? std::pow(x, power) / std::pow(mid, power - 1) : 1 - std::pow(1 - x, power) / std::pow(1 - mid, power - 1));
  return dmin + y * (dmax - dmin);
}
static void kb_h(const float* solref, const float* solimp, double dt, float* K, float* B) {
  double tc = solref[0], dr = solref[1], dmax = solimp[1];
  dmax = dmax < 1e-4 ? 1e-4 : (dmax > 0.9999 ? 0.9999 : dmax);
  if (tc < 2 * dt) tc = 2 * dt;  // refsafe
  *K = (float)(1.0 / std::fmax(1e-15, dmax * dmax * tc * tc * dr * dr));
  *B = (float)(2.0 / std::fmax(1e-15, dmax * tc));
}

// envs per warp: 16 fills the lanes; with few envs, fewer per warp shorten every warp's Newton loop (it runs as long as
// its slowest env).  Measured on B200 (tools/diag_epw.py): the plain kernel is best at the smallest group that keeps the grid within
// ~3.5 warps per SM (round 1: 4096 envs: 8 per warp, 0.309 ms vs 0.330; 2048: 4; 1024: 2; >= 8192: 16).  The mirror-lane
// instantiations (lanes that would idle mirror the working ones and share the independent loops of the Newton trip) move that:
// below one warp per scheduler at 4 per warp (2368 envs on 148 SMs) 4 per warp with FOUR mirrors per lane is ahead of everything else
// (profiles/r4_notes.md: 2048 envs 0.1784 ms against 0.1911 at 8 and 0.2053 plain; 1024: 0.1699 / 0.1847 / 0.1973; 256: 0.1610
// against 0.1703 at one env per warp), then 8 per warp with two mirrors up to ~4.75 warps per SM (5624 envs; 5120 envs: 0.2255
// against 0.2357 ms at 16 per warp, 6144 envs: 0.2389 against 0.2379), 16 per warp above.  cfg.reserved[2] overrides the result.
extern "C" int h1v2_envs_per_warp(int n_envs, int sms, int plain) {
  const int n = n_envs < 1 ? 1 : n_envs;
  if (sms < 1) sms = 148;
  int epw = 1;
  while (epw < 16 && 2 * ((n + epw - 1) / epw) > 7 * sms) epw *= 2;
  if (plain) return epw;
  if ((n + 3) / 4 <= 4 * sms) return 4;
  return 4 * ((n + 7) / 8) <= 19 * sms ? 8 : 16;
}

static int build_params(const H1v2Config& c, int n, uint64_t seed, KParams& P) {
  std::memset(&P, 0, sizeof(P));
  static const int axis_expect[6] = {2, 1, 0, 1, 1, 0};
  for (int s = 0; s < 2; s++)
    for (int i = 0; i < 6; i++) {
      const int b = 1 + 6 * s + i;
      if (h1v2_body_parent[b] != (i == 0 ? 0 : b - 1)) return fail("model: unexpected kinematic tree");
      for (int k = 0; k < 3; k++)
        if (h1v2_jnt_axis[b - 1][k] != (k == axis_expect[i] ? 1.0 : 0.0)) return fail("model: unexpected joint axis pattern");
      KLeg& L = P.leg[s];
      for (int k = 0; k < 3; k++) { L.pos[i][k] = (float)h1v2_body_pos[b][k]; L.ipos[i][k] = (float)h1v2_body_ipos[b][k]; }
      const double* I = h1v2_body_inertia[b];
      const double six[6] = {I[0], I[4], I[8], I[1], I[2], I[5]};
      for (int k = 0; k < 6; k++) L.inertia[i][k] = (float)six[k];
      L.mass[i] = (float)h1v2_body_mass[b];
    }
  {
    int nf[2] = {0, 0}, ns[2] = {0, 0}, nr = 0;
    for (int r = 0; r < H1V2_NCOLL; r++) {
      const double* row = h1v2_coll[r];
      const int b = (int)row[0], slot = (int)row[5];
      if (slot < 2) { float* d = P.leg[slot].foot_pt[nf[slot]++]; d[0] = (float)row[1]; d[1] = (float)row[2]; d[2] = (float)row[3]; }
      else if (slot < 4) { float* d = P.leg[slot - 2].shin_pt[ns[slot - 2]++]; d[0] = (float)row[1]; d[1] = (float)row[2]; d[2] = (float)row[3]; P.leg[slot - 2].shin_rad = (float)row[4]; }
      else { float* d = P.root_pt[nr]; d[0] = (float)row[1]; d[1] = (float)row[2]; d[2] = (float)row[3]; P.root_rad[nr++] = (float)row[4]; }
      (void)b;
    }
    if (nf[0] != 4 || nf[1] != 4 || ns[0] != 2 || ns[1] != 2 || nr != 9) return fail("model: unexpected collider table");
  }
  P.root_mass = (float)h1v2_body_mass[0];
  for (int k = 0; k < 3; k++) P.root_ipos[k] = (float)h1v2_body_ipos[0][k];
  {
    const double* I = h1v2_body_inertia[0];
    const double six[6] = {I[0], I[4], I[8], I[1], I[2], I[5]};
    for (int k = 0; k < 6; k++) P.root_inertia[k] = (float)six[k];
  }
  P.h = c.sim_dt; P.decimation = c.decimation; P.step_dt = c.sim_dt * (float)c.decimation;
  P.max_episode_length = (int64_t)std::ceil((double)c.episode_length_s / ((double)c.sim_dt * c.decimation) * (1.0 - 1e-6));
  P.max_episode_length_s = c.episode_length_s;
  P.action_scale = c.action_scale;
  bool seen[12] = {false};
  for (int i = 0; i < 12; i++) {
    P.q0[i] = c.default_joint_pos[i]; P.kp[i] = c.kp[i]; P.kd[i] = c.kd[i]; P.effort[i] = c.effort_limit[i]; P.frc[i] = c.act_frc_limit[i];
    const int j = c.joint_perm[i];
    if (j < 0 || j >= 12 || seen[j]) return fail("config: joint_perm is not a permutation");
    seen[j] = true;
    P.perm[i] = j; P.inv_perm[j] = i;
    P.range_lo[i] = c.joint_range[i][0]; P.range_hi[i] = c.joint_range[i][1];
    const double mid = 0.5 * ((double)c.joint_range[i][0] + c.joint_range[i][1]), half = 0.5 * ((double)c.joint_range[i][1] - c.joint_range[i][0]) * c.soft_limit_factor;
    P.soft_lo[i] = (float)(mid - half); P.soft_hi[i] = (float)(mid + half);
    P.limit_invw[i] = (float)h1v2_dof_invweight0[6 + i];
  }
  if (c.min_delay < 0 || c.max_delay > 7 || c.max_delay < c.min_delay) return fail("config: delays must satisfy 0 <= min <= max <= 7");
  if (c.max_delay > 2 * c.decimation) return fail("config: max_delay exceeds two control steps");
  P.min_delay = c.min_delay; P.max_delay = c.max_delay;
  P.gravity = c.gravity;
  P.mass_scales_inertia = c.mass_recompute_inertia;
  P.vel_limit = c.joint_vel_limit; P.runaway_vel = c.runaway_vel > 0.f ? c.runaway_vel : 3.0e38f;
  float Kf, Bf;
  kb_h(c.floss_solref, c.floss_solimp, c.sim_dt, &Kf, &Bf);
  P.floss_B = Bf;
  const double imp0 = impedance_h(c.floss_solimp, 0.0);
  for (int d = 0; d < 18; d++) {
    P.damping[d] = c.dof_damping[d]; P.armature[d] = c.dof_armature[d]; P.floss[d] = c.dof_frictionloss[d];
    const double R = std::fmax(1e-15, (1 - imp0) * h1v2_dof_invweight0[d] / imp0);
    P.floss_D[d] = c.dof_frictionloss[d] > 0 ? (float)(1.0 / R) : 0.f;
    P.floss_lim[d] = (float)(R * c.dof_frictionloss[d]);
  }
  kb_h(c.limit_solref, c.limit_solimp, c.sim_dt, &P.limit_K, &P.limit_B);
  kb_h(c.contact_solref, c.contact_solimp, c.sim_dt, &P.contact_K, &P.contact_B);
  for (int k = 0; k < 5; k++) { P.limit_imp[k] = c.limit_solimp[k]; P.contact_imp[k] = c.contact_solimp[k]; }
  for (int s = 0; s < 6; s++) P.slot_tran[s] = (float)h1v2_slot_invweight_tran[s];
  P.max_iters = c.solver_iterations; P.tol = c.solver_tolerance; P.step_tol = c.solver_step_tolerance; P.vel_tol = c.solver_vel_tolerance > 0.f ? c.solver_vel_tolerance / c.sim_dt : 3.0e38f; P.ls_tol = c.solver_ls_tolerance > 0.f ? c.solver_ls_tolerance : 0.01f;
  P.ls_max = c.reserved[1] > 0 ? c.reserved[1] : 6;  // reserved[1]: tuning knob for the line-search trip cap
  P.grad_scale = (float)(1.0 / (H1V2_MEANINERTIA * 18.0));
  if (c.history_length < 1 || c.history_length > H1V2_MAX_HISTORY) return fail("config: history_length out of range");
  P.H = c.history_length; P.obs_dim = c.history_length * H1V2_OBS_TERM_DIM; P.corrupt = c.enable_corruption;
  P.lut_dim = P.obs_dim;
  P.rough = (c.terrain_enable || c.obs_base_lin_vel || c.obs_height_scan) ? 1 : 0;
  if (P.rough) {
    // Rough id (V/velocity_env_cfg.py:119-142): [base_lin_vel 3] | the 45 regular terms | [height_scan], no history
    if (c.history_length != 1) return fail("config: base_lin_vel / height_scan / terrain need history_length 1 (the Rough id has no observation history)");
    if (c.cat_enable) return fail("config: the Constraints-as-Terminations tail is not available on the rough-terrain instantiation");
    P.lin_vel = c.obs_base_lin_vel ? 1 : 0; P.n_lv = c.noise_lin_vel; P.s_lv = c.scale_lin_vel;
    P.lut_dim = (P.lin_vel ? 3 : 0) + H1V2_OBS_TERM_DIM;
    P.obs_dim = P.lut_dim;
    if (c.obs_height_scan) {
      if (!(c.scan_resolution > 0.f) || !(c.scan_size[0] >= 0.f) || !(c.scan_size[1] >= 0.f)) return fail("config: bad height-scan pattern");
      // len(torch.arange(-size/2, size/2 + 1e-9, res)); the 1e-4 absorbs the fp32 rounding of the config values (0.1f > 0.1)
      P.scan_nx = (int)std::floor((double)c.scan_size[0] / (double)c.scan_resolution + 1e-4) + 1;
      P.scan_ny = (int)std::floor((double)c.scan_size[1] / (double)c.scan_resolution + 1e-4) + 1;
      if (P.scan_nx * P.scan_ny > H1V2_MAX_SCAN) return fail("config: too many height-scan rays");
      P.scan_res = c.scan_resolution; P.scan_x0 = -0.5f * c.scan_size[0]; P.scan_y0 = -0.5f * c.scan_size[1];
      P.scan_off = c.scan_offset; P.n_scan = c.noise_height_scan; P.s_scan = c.scale_height_scan; P.scan_lo = c.scan_clip[0]; P.scan_hi = c.scan_clip[1];
      P.scan_col0 = P.lut_dim;
      P.obs_dim += P.scan_nx * P.scan_ny;
    }
    if (c.terrain_enable) {
      if (c.terrain_rows < 1 || c.terrain_rows > 255 || c.terrain_cols < 1 || c.terrain_cols > 255) return fail("config: terrain_rows / terrain_cols must be in 1..255");
      if (!(c.terrain_tile_size > 0.f) || !(c.terrain_hscale > 0.f)) return fail("config: bad terrain tile size / horizontal scale");
      P.t_rows = c.terrain_rows; P.t_cols = c.terrain_cols; P.t_curriculum = c.terrain_curriculum ? 1 : 0;
      P.t_npx = (int)std::lround(c.terrain_tile_size / c.terrain_hscale);
      P.t_tile = c.terrain_tile_size; P.t_half = 0.5f * c.terrain_tile_size; P.t_inv_hs = 1.0f / c.terrain_hscale;
    } else {  // a plane under the rough instantiation: one flat tile
      P.t_rows = P.t_cols = 1; P.t_curriculum = 0; P.t_npx = 2; P.t_tile = 8.f; P.t_half = 4.f; P.t_inv_hs = 0.25f;
    }
    P.t_gx = P.t_rows * P.t_npx + 1; P.t_gy = P.t_cols * P.t_npx + 1;
    if ((size_t)P.t_gx * P.t_gy > ((size_t)1 << 28)) return fail("config: terrain grid too large");
  }
  P.n_av = c.noise_ang_vel; P.n_g = c.noise_gravity; P.n_q = c.noise_joint_pos; P.n_v = c.noise_joint_vel;
  P.s_av = c.scale_ang_vel; P.s_g = c.scale_gravity; P.s_cmd = c.scale_cmd; P.s_q = c.scale_joint_pos; P.s_v = c.scale_joint_vel; P.s_a = c.scale_action;
  for (int t = 0; t < H1V2_NUM_REW; t++) P.w[t] = c.rew_weight[t];
  P.inv_std2 = 1.f / (c.track_std * c.track_std); P.air_thr = c.feet_air_threshold; P.contact_thr = c.contact_threshold; P.base_h = c.base_height_target;
  P.m_poslim = c.mask_pos_limits; P.m_dev = c.mask_joint_dev; P.m_poslim_b = c.mask_pos_limits_b; P.m_dev_b = c.mask_joint_dev_b;
  P.m_cforce = c.mask_contact_forces_slots; P.cforce_thr = c.contact_forces_threshold;
  for (int k = 0; k < 3; k++) P.root_com[k] = c.root_link_com[k];
  P.foot_vel_com = c.body_vel_at_com;
  P.m_tau = c.mask_torques; P.m_undesired = c.mask_undesired_slots; P.m_illegal = c.mask_illegal_slots;
  for (int k = 0; k < 2; k++) {
    P.c_lx[k] = c.cmd_lin_x[k]; P.c_ly[k] = c.cmd_lin_y[k]; P.c_wz[k] = c.cmd_ang_z[k]; P.c_hd[k] = c.cmd_heading[k]; P.c_rt[k] = c.cmd_resample_time[k];
    P.rjp[k] = c.reset_joint_pos_scale[k]; P.rjv[k] = c.reset_joint_vel_scale[k]; P.push_int[k] = c.push_interval_s[k]; P.push_v[k] = c.push_vel_xy[k];
    for (int a = 0; a < 6; a++) { P.rp[a][k] = c.reset_pose_range[a][k]; P.rv[a][k] = c.reset_vel_range[a][k]; }
  }
  P.rel_standing = c.rel_standing_envs; P.rel_heading = c.rel_heading_envs; P.k_heading = c.heading_stiffness; P.heading_cmd = c.heading_command;
  P.max_command_step = c.cmd_resample_time[1] / P.step_dt;
  if (c.command_class != 0 && c.command_class != 1) return fail("config: command_class must be 0 (UniformVelocityCommand) or 1 (UniformVelocityCommandWithDeadzone)");
  if (c.command_class == 1 && !(c.velocity_deadzone >= 0.f)) return fail("config: velocity_deadzone must be >= 0");
  P.cmd_class = c.command_class;
  P.deadzone = c.velocity_deadzone;
  P.flip_prob = c.ang_vel_flip_prob;
  P.init_h = c.init_root_height; P.push_enable = c.push_enable;
  P.key0 = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u);
  P.env_id_offset = c.env_id_offset;
  P.n = n;
  {
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    P.epw = h1v2_envs_per_warp(n, sms, c.reserved[3] != 0);
    if (c.reserved[2] == 1 || c.reserved[2] == 2 || c.reserved[2] == 4 || c.reserved[2] == 8 || c.reserved[2] == 16) P.epw = c.reserved[2];
  }
  return 0;
}

// Every device array of a handle sits between two 256-byte guard zones filled with a pattern; h1v2_check_guards counts the guard
// bytes that no longer hold it.  (compute-sanitizer is closed on the GPU pool this was developed on -- profiles/
// r2_sanitizer_unavailable.txt -- so out-of-bounds stores are caught this way: tests/test_gpu_env.py.)
static const size_t kGuard = 256;
static const unsigned char kGuardByte = 0xA5;
template <typename T>
static int dalloc(H1v2Handle* h, T** p, size_t count) {
  const size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
  unsigned char* q = nullptr;
  cudaError_t e = cudaMalloc(&q, bytes + 2 * kGuard);
  if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
  e = cudaMemset(q, kGuardByte, bytes + 2 * kGuard);
  if (e == cudaSuccess) e = cudaMemset(q + kGuard, 0, bytes);
  if (e != cudaSuccess) { cudaFree(q); return fail(std::string("cudaMemset: ") + cudaGetErrorString(e)); }
  h->allocs.push_back(q);
  h->alloc_bytes.push_back(bytes);
  *p = (T*)(q + kGuard);
  return 0;
}

// device-visible alias of a host pointer when it is pinned (cudaHostAlloc / cudaHostRegister / torch pin_memory), else NULL
template <typename T>
static T* mapped_alias(T* host) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
  return (T*)a.devicePointer;
}


// ------------------------------------------------------------------------------------------------------
// Host side of h1v2_step_host, mode 1: the kernel ships only the NEW 45-float sample of every env (+ its first-push flag),
// the term-major [N, 45 H] rows are assembled on the host from a host mirror of the history ring -- bit-identical to the rows
// the kernel emits (observation_manager.py:335-355, circular_buffer.py:79-87,131-135).  9/10 of every row is history the
// host already has; over PCIe go 192 B per env instead of 1.8 KB.  The old part of the rows (history slots head+1 .. head-1)
// does not depend on the step in flight, so the pool writes it WHILE the kernel runs (phase 1); after the stream has
// synchronised only the newest slot is scattered into the rows and the mirror (phase 2).
// ------------------------------------------------------------------------------------------------------
struct HostPool {
  int nthreads = 1;
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable cv;
  std::atomic<uint64_t> gen{0};   // job generation; workers spin on it for a short while after a job (a caller in a step loop
  std::atomic<int> sleepers{0};   // comes back within microseconds), then sleep on cv: no futex wake-up on the hot path
  bool stop = false;
  std::atomic<uint64_t> phase2{0};
  std::atomic<int> done{0};
  // job (written by the caller before gen is advanced)
  const float* sample = nullptr;
  float* ring = nullptr;
  float* obs = nullptr;
  int n = 0, H = 0, head = 0;
  int e0 = 0;  // first env the pool assembles (the envs below it get their rows from the kernel)
  // streaming hand-over (flags != NULL): the kernel raises flags[w] = seq once warp w's outputs are in host memory; the workers take the new
  // slot of an env as soon as its warp has reported instead of waiting for the whole launch
  const volatile unsigned* flags = nullptr;
  unsigned seq = 0;
  int epw = 16;
  std::atomic<int> failed{0};
};
static const int kTermOff[7] = {0, 3, 6, 9, 21, 33, 45};
static inline void env_range(const HostPool* p, int t, int& a, int& b) {
  const int cnt = p->n - p->e0, per = (cnt + p->nthreads - 1) / p->nthreads;
  a = p->e0 + std::min(cnt, t * per); b = std::min(p->n, a + per);
}
// phase 1: history slots older than the step in flight, slot by slot.  (Measured on the B200 box's 16-core host and not kept: writing
// term-major doubles the time of this phase -- 42 vs 20 us at 4096 envs -- and non-temporal row stores make a 32768-env step 20 %
// slower, 1.23 vs 1.01 ms.)
static void assemble_old(const HostPool* p, int t) {
  int a, b; env_range(p, t, a, b);
  const int H = p->H, od = 45 * H;
  for (int e = a; e < b; e++) {
    const float* ring = p->ring + (size_t)e * H * H1V2_HIST_STRIDE;
    float* row = p->obs + (size_t)e * od;
    for (int hh = 0; hh + 1 < H; hh++) {
      int sl = p->head + 1 + hh; sl = sl >= H ? sl - H : sl;
      const float* src = ring + sl * H1V2_HIST_STRIDE;
      for (int k = 0; k < 9; k++) row[(k / 3) * 3 * H + hh * 3 + (k % 3)] = src[k];
      std::memcpy(row + 9 * H + hh * 12, src + 9, 48);
      std::memcpy(row + 21 * H + hh * 12, src + 21, 48);
      std::memcpy(row + 33 * H + hh * 12, src + 33, 48);
    }
  }
}
// spin until warp w of the launch in flight has reported; false after ~2 s (a faulted kernel never reports)
static bool wait_flag(HostPool* p, int w) {
  if (p->flags[w] == p->seq) return true;
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spin = 1;; spin++) {
    if (p->flags[w] == p->seq) return true;
    if (p->failed.load(std::memory_order_relaxed)) return false;
    CPU_PAUSE();
    if ((spin & 0xffff) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 2.0) {
      p->failed.store(1, std::memory_order_relaxed);
      return false;
    }
  }
}
static void assemble_new_range(const HostPool* p, int a, int b);
static void assemble_new_streaming(HostPool* p, int t) {  // the new slot of every env of the range, warp by warp as the kernel reports them
  int a, b; env_range(p, t, a, b);
  for (int e = a; e < b;) {
    const int w = e / p->epw, e1 = std::min(b, (w + 1) * p->epw);
    if (!wait_flag(p, w)) return;
    std::atomic_thread_fence(std::memory_order_acquire);
    assemble_new_range(p, e, e1);
    e = e1;
  }
}
static void assemble_new(const HostPool* p, int t) {  // phase 2: the new sample; a first push fills the whole ring and row
  int a, b; env_range(p, t, a, b);
  assemble_new_range(p, a, b);
}
static void assemble_new_range(const HostPool* p, int a, int b) {
  const int H = p->H, od = 45 * H;
  for (int e = a; e < b; e++) {
    const float* s = p->sample + (size_t)e * H1V2_HIST_STRIDE;
    float* ring = p->ring + (size_t)e * H * H1V2_HIST_STRIDE;
    float* row = p->obs + (size_t)e * od;
    const bool fresh = s[45] != 0.f;
    for (int sl = fresh ? 0 : p->head; sl < (fresh ? H : p->head + 1); sl++) std::memcpy(ring + sl * H1V2_HIST_STRIDE, s, 45 * sizeof(float));
    for (int hh = fresh ? 0 : H - 1; hh < H; hh++)
      for (int tm = 0; tm < 6; tm++) {
        const int d = kTermOff[tm + 1] - kTermOff[tm];
        std::memcpy(row + kTermOff[tm] * H + hh * d, s + kTermOff[tm], d * sizeof(float));
      }
  }
}
static void pool_worker(HostPool* p, int t) {
  uint64_t seen = 0;
  for (;;) {
    bool got = false;
    for (int spin = 0; spin < 20000 && !got; spin++) {  // ~100-200 us
      got = p->gen.load(std::memory_order_acquire) != seen;
      if (!got) CPU_PAUSE();
    }
    if (!got) {
      std::unique_lock<std::mutex> lk(p->m);
      p->sleepers.fetch_add(1, std::memory_order_relaxed);
      p->cv.wait(lk, [&] { return p->stop || p->gen.load(std::memory_order_acquire) != seen; });
      p->sleepers.fetch_sub(1, std::memory_order_relaxed);
      if (p->stop) return;
    }
    seen = p->gen.load(std::memory_order_acquire);
    assemble_old(p, t);
    if (p->flags) assemble_new_streaming(p, t);
    else {
      while (p->phase2.load(std::memory_order_acquire) != seen) CPU_PAUSE();  // the kernel is in flight: < 1 ms
      assemble_new(p, t);
    }
    p->done.fetch_add(1, std::memory_order_release);
  }
}
static HostPool* pool_create() {
  HostPool* p = new HostPool();
  int nt = 0;
  if (const char* e = std::getenv("H1V2_HOST_THREADS")) nt = std::atoi(e);
  if (nt <= 0) {  // the box's cores shared by the ranks of this node (torchrun exports LOCAL_WORLD_SIZE); at most 16 per handle
    int ranks = 1;
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, std::atoi(e));
    nt = (int)std::thread::hardware_concurrency() / ranks;
  }
  p->nthreads = std::max(1, std::min(nt, 16));
  for (int t = 1; t < p->nthreads; t++) p->threads.emplace_back(pool_worker, p, t);
  return p;
}
static void pool_destroy(HostPool* p) {
  if (!p) return;
  { std::lock_guard<std::mutex> lk(p->m); p->stop = true; }
  p->cv.notify_all();
  for (auto& t : p->threads) t.join();
  delete p;
}


// ------------------------------------------------------------------------------------------------------
// Rough id: the height field.  Host-side generation of what upstream isaaclab builds from the reference's generator cfg
// (packages/biped_tasks/biped_tasks/utils/mdp/terrains.py:11-28): every tile a height field whose rim of border_px vertices is flat
// and whose interior vertices hold independent uniform draws from {level_min .. level_max} * vscale (height_field/hf_terrains.py
// random_uniform_terrain; the draws come from this backend's Philox stream 5, keyed by the grid row).
// ------------------------------------------------------------------------------------------------------
static void philox_host(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&o)[4]) {
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
static void terrain_generate_host(const H1v2Config& c, const KParams& P, std::vector<float>& H) {
  H.assign((size_t)P.t_gx * P.t_gy, 0.f);
  if (!c.terrain_enable) return;
  const int npx = P.t_npx, bp = c.terrain_border_px;
  const int nlev = (c.terrain_level_max - c.terrain_level_min) / (c.terrain_level_step > 0 ? c.terrain_level_step : 1) + 1;
  for (int i = 0; i < P.t_gx; i++)
    for (int j0 = 0; j0 < P.t_gy; j0 += 4) {
      uint32_t o[4];
      philox_host((uint32_t)(j0 / 4), 0u, 5u /* STREAM_TERRAIN */, 0u, P.key0, (uint32_t)i, o);
      for (int k = 0; k < 4 && j0 + k < P.t_gy; k++) {
        const int j = j0 + k, a = i % npx, b = j % npx;
        const bool rim = a < bp || a > npx - bp || b < bp || b > npx - bp || i == P.t_gx - 1 || j == P.t_gy - 1;
        const float u = (float)(o[k] >> 8) * (1.0f / 16777216.0f);
        int lev = (int)(u * (float)nlev);
        if (lev > nlev - 1) lev = nlev - 1;
        H[(size_t)i * P.t_gy + j] = rim ? 0.f : (float)(c.terrain_level_min + lev * c.terrain_level_step) * c.terrain_vscale;
      }
    }
}
// tile origin = tile centre at the highest vertex of the central 2 m x 2 m patch (height_field/utils.py height_field_to_mesh)
static void terrain_origins_host(const H1v2Config& c, const KParams& P, const float* H, std::vector<float>& oz) {
  oz.assign((size_t)P.t_rows * P.t_cols, 0.f);
  if (!c.terrain_enable) return;
  const int a1 = (int)((c.terrain_tile_size * 0.5f - 1.0f) / c.terrain_hscale), a2 = (int)((c.terrain_tile_size * 0.5f + 1.0f) / c.terrain_hscale);
  for (int r = 0; r < P.t_rows; r++)
    for (int q = 0; q < P.t_cols; q++) {
      float m = -1e30f;
      for (int a = a1; a < a2; a++)
        for (int b = a1; b < a2; b++) m = std::fmax(m, H[(size_t)(r * P.t_npx + a) * P.t_gy + (q * P.t_npx + b)]);
      oz[(size_t)r * P.t_cols + q] = m;
    }
}
static int terrain_upload(H1v2Handle* h, const float* H) {
  std::vector<float> oz;
  terrain_origins_host(h->cfg, h->P, H, oz);
  if (cudaMemcpy(h->d_terrain, H, sizeof(float) * (size_t)h->P.t_gx * h->P.t_gy, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->d_origin_z, oz.data(), sizeof(float) * oz.size(), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail("terrain upload: cudaMemcpy failed");
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------------
extern "C" {

const char* h1v2_last_error(void) { return g_err.c_str(); }

int h1v2_create(const H1v2Config* cfg, int32_t n_envs, int32_t device, uint64_t seed, H1v2Handle** out) {
  if (!cfg || !out || n_envs <= 0) return fail("h1v2_create: bad arguments");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail("h1v2_create: no CUDA device (the CUDA path is the product; there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail("h1v2_create: no such CUDA device");
  DeviceGuard guard(device);
  if (!guard.ok) return fail("h1v2_create: cudaSetDevice failed");
  H1v2Handle* h = new H1v2Handle();
  h->cfg = *cfg; h->n = n_envs; h->device = device; h->seed = seed;
  if (build_params(*cfg, n_envs, seed, h->P) != 0) { delete h; return -1; }
  h->no_quad = cfg->reserved[3] != 0;
  const size_t N = (size_t)n_envs;
  KState& S = h->S;
  int rc = 0;
  rc |= dalloc(h, &S.root, 4 * N);
  rc |= dalloc(h, &S.leg, 3 * 2 * N);
  rc |= dalloc(h, &S.act, 5 * 2 * N);
  rc |= dalloc(h, &S.cmd, 2 * N);
  rc |= dalloc(h, &S.timers, 2 * N);
  rc |= dalloc(h, &S.warm, 3 * 2 * N);
  rc |= dalloc(h, &S.epsum, H1V2_EPSUM_F4 * N);
  rc |= dalloc(h, &S.hist, N * (size_t)h->P.H * H1V2_HIST_STRIDE);
  rc |= dalloc(h, &S.acc, (size_t)H1V2_LOG_DIM + 32);  // + cumulative histogram of Newton iterations per solve
  rc |= dalloc(h, &S.log, (size_t)H1V2_LOG_DIM);
  rc |= dalloc(h, &S.counters, (size_t)4);
  rc |= dalloc(h, &S.done, (size_t)1);
  rc |= dalloc(h, &h->own_ep_len, N);
  int* lut_d = nullptr;
  rc |= dalloc(h, &lut_d, (size_t)h->P.obs_dim);
  if (cfg->reserved[0]) rc |= dalloc(h, &S.diag, N * H1V2_DIAG_DIM);
  h->rough = h->P.rough != 0;
  if (h->rough) {
    rc |= dalloc(h, &h->d_terrain, (size_t)h->P.t_gx * h->P.t_gy);
    rc |= dalloc(h, &h->d_origin_z, (size_t)h->P.t_rows * h->P.t_cols);
    rc |= dalloc(h, &S.tlog, (size_t)2);
    S.terrain_h = h->d_terrain; S.terrain_oz = h->d_origin_z;
  }
  if (cfg->cat_enable) {
    CatState& T = h->cat;
    rc |= dalloc(h, &T.k.raw, (size_t)H1V2_CSTR_COLS * N);
    rc |= dalloc(h, &T.k.qd, (size_t)12 * N);
    rc |= dalloc(h, &T.k.aux, (size_t)2 * N);
    rc |= dalloc(h, &T.k.dz, N);
    rc |= dalloc(h, &T.k.cmax, (size_t)H1V2_CSTR_COLS);
    rc |= dalloc(h, &T.k.list, N);
    rc |= dalloc(h, &T.k.ctl, (size_t)4);
    rc |= dalloc(h, &T.k.swing, 2 * N);
    rc |= dalloc(h, &T.k.logacc, (size_t)2 * H1V2_NUM_CSTR + 1);
    rc |= dalloc(h, &T.probs, (size_t)H1V2_CSTR_COLS * N);
    rc |= dalloc(h, &T.rmax, (size_t)2 * H1V2_CSTR_COLS);
    rc |= dalloc(h, &T.sums, (size_t)2 * H1V2_NUM_CSTR * N);
    rc |= dalloc(h, &h->cat_term, N);
    if (rc == 0) {  // column maxima start at their floor (constraint.max(0).clamp(min=1e-6)); the apply kernel restores it after every step
      std::vector<float> floor_v(H1V2_CSTR_COLS, 1e-6f);
      if (cudaMemcpy(T.k.cmax, floor_v.data(), sizeof(float) * H1V2_CSTR_COLS, cudaMemcpyHostToDevice) != cudaSuccess) rc = fail("cudaMemcpy: cat maxima");
    }
    T.k.contact_slots = cfg->cat_contact_slots;
    T.k.foot_force_limit = cfg->cat_foot_force_limit; T.k.no_move_deadzone = cfg->cat_no_move_deadzone; T.k.no_move_vel_limit = cfg->cat_no_move_vel_limit;
    T.k.orientation_limit = cfg->cat_orientation_limit; T.k.height = cfg->cat_height; T.k.height_std = cfg->cat_height_std;
    T.k.clearance_min_height = cfg->cat_clearance_min_height; T.k.clearance_deadzone = cfg->cat_clearance_deadzone;
    T.k.vel_limit = cfg->joint_vel_limit;
  }
  if (rc != 0) { h1v2_destroy(h); return -1; }
  S.ep_len = h->own_ep_len;
  {
    std::vector<int> lut(h->P.obs_dim);
    static const int off[7] = {0, 3, 6, 9, 21, 33, 45};
    int w = 0;
    if (h->P.lin_vel)  // base_lin_vel leads the Rough id's row; it travels in floats 45..47 of the slot
      for (int k = 45; k < 48; k++) lut[w++] = k;
    for (int t = 0; t < 6; t++)
      for (int hh = 0; hh < h->P.H; hh++)
        for (int k = off[t]; k < off[t + 1]; k++) lut[w++] = (hh << 8) | k;
    CKH(cudaMemcpy(lut_d, lut.data(), lut.size() * sizeof(int), cudaMemcpyHostToDevice));
    S.lut = lut_d;
  }
  if (h->rough) {
    std::vector<float> H;
    terrain_generate_host(*cfg, h->P, H);
    if (terrain_upload(h, H.data()) != 0) { h1v2_destroy(h); return -1; }
  }
  CKH(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
  CKH(cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
  CKH(cudaEventCreateWithFlags(&h->host_ev, cudaEventDisableTiming));
  {
    const int rows = h->P.t_rows, mi = cfg->terrain_max_init_level;
    startup_kernel<<<(n_envs + 127) / 128, 128>>>(h->P, h->S, cfg->friction_range[0], cfg->friction_range[1], cfg->mass_add_range[0], cfg->mass_add_range[1],
                                                  (mi < 0 || mi > rows - 1) ? rows - 1 : mi);
  }
  reset_kernel<<<(2 * n_envs + 127) / 128, 128>>>(h->P, h->S, nullptr, n_envs, h->cat.sums);
  h->launches += 2;
  CKH(cudaGetLastError());
  CKH(cudaDeviceSynchronize());
  *out = h;
  return 0;
}

void h1v2_destroy(H1v2Handle* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  for (void* p : {(void*)h->d_act, (void*)h->d_obs, (void*)h->d_rew, (void*)h->d_term, (void*)h->d_trunc, (void*)h->d_dones})
    if (p) cudaFree(p);
  if (h->host_stream) cudaStreamDestroy(h->host_stream);
  if (h->order_ev) cudaEventDestroy(h->order_ev);
  if (h->host_ev) cudaEventDestroy(h->host_ev);
  pool_destroy(h->pool);
  if (h->h_sample) cudaFreeHost(h->h_sample);
  if (h->h_flags) cudaFreeHost(h->h_flags);
  std::free(h->h_ring);
  delete h;
}

int64_t h1v2_check_guards(H1v2Handle* h) {
  if (!h) return -1;
  DeviceGuard guard(h->device);
  if (cudaDeviceSynchronize() != cudaSuccess) { fail("h1v2_check_guards: device error before the check"); return -1; }
  int64_t bad = 0;
  unsigned char buf[2 * 256];
  for (size_t i = 0; i < h->allocs.size(); i++) {
    const unsigned char* q = (const unsigned char*)h->allocs[i];
    if (cudaMemcpy(buf, q, kGuard, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(buf + kGuard, q + kGuard + h->alloc_bytes[i], kGuard, cudaMemcpyDeviceToHost) != cudaSuccess) {
      fail("h1v2_check_guards: cudaMemcpy failed");
      return -1;
    }
    for (size_t k = 0; k < 2 * kGuard; k++) bad += buf[k] != kGuardByte;
  }
  return bad;
}

int h1v2_obs_dim(const H1v2Handle* h) { return h ? h->P.obs_dim : -1; }
int h1v2_num_envs(const H1v2Handle* h) { return h ? h->n : -1; }
int64_t h1v2_launch_count(const H1v2Handle* h) { return h ? h->launches : -1; }

int h1v2_bind_episode_length(H1v2Handle* h, int64_t* episode_length) {
  if (!h) return fail("null handle");
  DeviceGuard guard(h->device);
  // rare call: wait for every step in flight (on whatever stream), then move the counters to their new home
  int64_t* to = episode_length ? episode_length : h->own_ep_len;
  if (to == h->S.ep_len) return 0;
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(to, h->S.ep_len, sizeof(int64_t) * h->n, cudaMemcpyDeviceToDevice));
  h->S.ep_len = to;
  return 0;
}

int h1v2_reset(H1v2Handle* h, const int64_t* env_ids, int32_t n, void* cuda_stream) {
  if (!h) return fail("null handle");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int cnt = env_ids ? n : h->n;
  if (cnt <= 0) return 0;
  order_after_host(h, st);
  reset_kernel<<<(2 * cnt + 127) / 128, 128, 0, st>>>(h->P, h->S, env_ids, cnt, h->cat.sums);
  note_stream(h, st);
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

static int launch_step(H1v2Handle* h, bool do_step, const float* actions, float* obs, float* rew, uint8_t* term, uint8_t* trunc, cudaStream_t st,
                       float* sample_out = nullptr, bool cat = false, int n_rows = 0, unsigned* host_flags = nullptr, unsigned host_seq = 0) {
  DeviceGuard guard(h->device);
  KState S = h->S;
  S.sample_out = sample_out;
  S.n_rows = n_rows;
  S.host_flags = host_flags; S.host_seq = host_seq;
  order_after_host(h, st);
  if (cat) S.cat = h->cat.k;  // the step kernel also leaves the raw constraint columns and their maxima (h1v2_cat_step)
  h->hist_launches += 1;  // every launch advances the history head (device counters[1])
  if (!sample_out) { h->ring_valid = false; note_stream(h, st); }  // the host mirror of the ring misses this launch's sample
  const int threads = H1V2_BLOCK;
  const int blocks = (h->n + h->P.epw - 1) / h->P.epw;  // one warp per block, epw envs per warp
  const size_t smem = (size_t)(h->rough ? SMEM_FLOATS_ROUGH : SMEM_FLOATS) * H1V2_BLOCK * sizeof(real);
  if (!h->attr_set && h->rough) {
    CK(cudaFuncSetAttribute((step_kernel<true, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((step_kernel<true, false, true>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute((step_kernel<false, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((step_kernel<true, false, true, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((step_kernel<true, false, true, 2>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute((step_kernel<true, false, true, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((step_kernel<true, false, true, 4>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    h->attr_set = true;
  }
  if (!h->attr_set) {  // function attributes are per device: keep the flag with the handle, not with the process
    CK(cudaFuncSetAttribute(step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(step_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((step_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((step_kernel<true, true>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute((step_kernel<true, false, false, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((step_kernel<true, false, false, 2>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute((step_kernel<true, false, false, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((step_kernel<true, false, false, 4>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute((step_kernel<true, true, false, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((step_kernel<true, true, false, 2>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute((step_kernel<true, true, false, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((step_kernel<true, true, false, 4>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    h->attr_set = true;
  }
  const bool quad = h->P.epw == 8 && !h->no_quad;  // 8 envs per warp: the mirror-lane instantiations (two mirrors per lane)
  const bool octo = h->P.epw == 4 && !h->no_quad;  // 4 envs per warp: four mirrors per lane
  if (h->rough && do_step && quad)
    step_kernel<true, false, true, 2><<<blocks, threads, smem, st>>>(h->P, S, actions, obs, rew, term, trunc);
  else if (h->rough && do_step && octo)
    step_kernel<true, false, true, 4><<<blocks, threads, smem, st>>>(h->P, S, actions, obs, rew, term, trunc);
  else if (h->rough && do_step)
    step_kernel<true, false, true><<<blocks, threads, smem, st>>>(h->P, S, actions, obs, rew, term, trunc);
  else if (h->rough)
    step_kernel<false, false, true><<<blocks, threads, smem, st>>>(h->P, S, nullptr, obs, nullptr, nullptr, nullptr);
  else if (do_step && cat && quad)
    step_kernel<true, true, false, 2><<<blocks, threads, smem, st>>>(h->P, S, actions, obs, rew, term, trunc);
  else if (do_step && cat && octo)
    step_kernel<true, true, false, 4><<<blocks, threads, smem, st>>>(h->P, S, actions, obs, rew, term, trunc);
  else if (do_step && cat)
    step_kernel<true, true><<<blocks, threads, smem, st>>>(h->P, S, actions, obs, rew, term, trunc);
  else if (do_step && quad)
    step_kernel<true, false, false, 2><<<blocks, threads, smem, st>>>(h->P, S, actions, obs, rew, term, trunc);
  else if (do_step && octo)
    step_kernel<true, false, false, 4><<<blocks, threads, smem, st>>>(h->P, S, actions, obs, rew, term, trunc);
  else if (do_step)
    step_kernel<true><<<blocks, threads, smem, st>>>(h->P, S, actions, obs, rew, term, trunc);
  else  // the observe-only launch stages the history rings in the same shared-memory window
    step_kernel<false><<<blocks, threads, smem, st>>>(h->P, S, nullptr, obs, nullptr, nullptr, nullptr);
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

int h1v2_observe(H1v2Handle* h, float* obs, void* cuda_stream) {
  if (!h || !obs) return fail("h1v2_observe: bad arguments");
  return launch_step(h, false, nullptr, obs, nullptr, nullptr, nullptr, (cudaStream_t)cuda_stream);
}

int h1v2_step(H1v2Handle* h, const float* actions, float* obs, float* rew, uint8_t* terminated, uint8_t* truncated, void* cuda_stream) {
  if (!h || !actions || !obs || !rew || !terminated || !truncated) return fail("h1v2_step: bad arguments");
  return launch_step(h, true, actions, obs, rew, terminated, truncated, (cudaStream_t)cuda_stream);
}

}  // extern "C"

static CatParams cat_params(const H1v2Handle* h) {
  const H1v2Config& c = h->cfg;
  CatParams C;
  C.n = h->n; C.epw = h->P.epw; C.nmask = (h->n + h->P.epw - 1) / h->P.epw;
  C.tau = c.cat_tau; C.min_p = c.cat_min_p;
  for (int t = 0; t < H1V2_NUM_CSTR; t++) C.max_p[t] = c.cat_max_p[t];
  return C;
}
static int launch_cat(H1v2Handle* h, const float* actions, float* obs, float* rew, float* dones, uint8_t* trunc, cudaStream_t st, float* sample_out, int n_rows = 0) {
  if (launch_step(h, true, actions, obs, rew, h->cat_term, trunc, st, sample_out, true, n_rows) != 0) return -1;
  DeviceGuard guard(h->device);
  cat_apply_kernel<<<(h->n + 63) / 64, 64, 0, st>>>(cat_params(h), h->cat, rew, dones);  // small blocks: at 4096 envs the tail is latency-bound
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// dones != NULL: the Constraints-as-Terminations step (h1v2_cat_step_host); reward and dones are staged through device buffers
// (the apply kernel rewrites the reward in place), `terminated` is unused
static int step_host_impl(H1v2Handle* h, const float* actions, float* obs, float* rew, uint8_t* terminated, float* dones, uint8_t* truncated) {
  const bool cat = dones != nullptr;
  DeviceGuard guard(h->device);
  const size_t N = (size_t)h->n, od = (size_t)h->P.obs_dim;
  cudaStream_t st = h->host_stream;
  if (h->last_stream_set) {  // order this step after whatever the stream-taking entry points queued on the caller's streams
    CK(cudaEventRecord(h->order_ev, h->last_stream));
    CK(cudaStreamWaitEvent(st, h->order_ev, 0));
    h->last_stream_set = false;
  }
  float* obs_dev = mapped_alias(obs);  // device alias of a pinned caller buffer, else NULL
  // candidates of the calibration: {mode, fraction of the envs whose rows the kernel writes over PCIe}
  static const struct { int mode; float frac; } kCand[5] = {{1, 0.f}, {0, 1.f}, {1, 0.25f}, {1, 0.5f}, {1, 0.75f}};
  const int epw = h->P.epw;
  auto rows_for = [&](float frac) { return std::min(h->n, (int)std::lround(frac * h->n / epw) * epw); };
  auto set_rows = [&](int r) { if (r != h->host_rows) { h->host_rows = r; h->ring_valid = false; } };  // the mirror is only kept for the assembled envs
  if (h->host_mode < 0) {
    // Three ways to bring the [N, 45 H] rows into the caller's buffer.  "rows": the kernel writes them (zero-copy over PCIe when the
    // buffer is pinned) -- PCIe-bound at large N (59 MB per step at 32768 envs).  "assemble": only the new 45-float sample of every
    // env crosses PCIe, host threads assemble the rows from a host mirror of the ring -- bound by the cores' memory bandwidth.
    // "hybrid": the kernel writes the rows of the first envs while the host threads assemble those of the rest: PCIe DMA and the
    // cores work concurrently.  Which wins depends on the host (cores per rank, cache, memory bandwidth, PCIe) as much as on the env
    // count, so a handle with a pinned buffer and at least four host threads measures: five candidates (assemble, rows, hybrid at
    // 1/4, 1/2, 3/4) x eight calls each (three warm-up: ring fetch, page faults, thread wake-up), from call 40 onwards the fastest.
    // All of them return bit-identical results, so the caller sees nothing of it.  H1V2_HOST_PATH=rows|assemble|hybrid
    // (H1V2_HOST_ROWS_FRAC, default 0.5) overrides.
    h->pool = pool_create();
    const char* e = std::getenv("H1V2_HOST_PATH");
    if (e && !std::strcmp(e, "rows")) h->host_mode = 0;
    else if (e && !std::strcmp(e, "assemble") && !h->rough) h->host_mode = 1;
    else if (e && !std::strcmp(e, "hybrid") && !h->rough && obs_dev) {
      const char* f = std::getenv("H1V2_HOST_ROWS_FRAC");
      h->host_mode = 1;
      set_rows(rows_for(f ? std::min(1.f, std::max(0.f, (float)std::atof(f))) : 0.5f));
    }
    else if (h->rough || h->pool->nthreads < 4) h->host_mode = 0;  // the Rough id has no history to assemble: rows
    else { h->host_mode = 1; h->host_calib = 0; }
  }
  if (h->host_mode == 1 && h->host_rows > 0 && !obs_dev) set_rows(0);  // a pageable buffer cannot take the kernel's rows directly
  const auto t_call0 = std::chrono::steady_clock::now();
  struct CalibGuard {  // times this call and advances the calibration when it returns
    H1v2Handle* h; std::chrono::steady_clock::time_point t0; bool hybrid_ok; int r25, r50, r75;
    ~CalibGuard() {
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (h->host_calib < 0) {
        // watchdog: the choice was made from 5 calls per candidate; if the steady state turns out 40 % slower than what was measured (other
        // ranks' host threads, a busier host), measure again -- at most twice, results are bit-identical whatever is picked
        if (h->host_chosen_t > 0.0 && h->host_recal < 2) {
          h->host_ema = h->host_since == 0 ? dt : 0.95 * h->host_ema + 0.05 * dt;
          if (++h->host_since >= 64 && h->host_ema > 1.4 * h->host_chosen_t) {
            h->host_recal++; h->host_since = 0; h->host_chosen_t = 0.0;
            for (int i = 0; i < 5; i++) h->host_calib_t[i] = 1e30;
            h->host_mode = 1; h->host_calib = 0;
            if (h->host_rows != 0) { h->host_rows = 0; h->ring_valid = false; }
          }
        }
        return;
      }
      const int per = 8, warm = 3, ncand = hybrid_ok ? 5 : 2;
      const int k = h->host_calib++, c = k / per;
      if (k % per >= warm) {
        h->host_calib_s[c][k % per - warm] = dt;
        if (k % per == per - 1) {
          // the FASTEST of the five decides.  (The median was tried: the first candidate is timed right after the handle's first calls -- first
          // touch of the caller's rows by the host threads, thread wake-up -- and lost to a worse one on an otherwise idle host:
          // profiles/r4_notes.md.  What the fastest call cannot see -- other ranks' host threads in the steady state -- is the watchdog's job.)
          double v[5];
          for (int i = 0; i < 5; i++) v[i] = h->host_calib_s[c][i];
          std::sort(v, v + 5);
          h->host_calib_t[c] = v[0];
        }
      }
      const int rows_of[5] = {0, 0, r25, r50, r75};
      auto apply = [&](int cand) {
        h->host_mode = kCand[cand].mode;
        const int r = kCand[cand].mode == 1 ? rows_of[cand] : 0;
        if (r != h->host_rows) { h->host_rows = r; h->ring_valid = false; }
      };
      if ((k + 1) % per == 0) {
        if (c + 1 < ncand) apply(c + 1);
        else {
          // the plain assemble path is the default; another candidate has to beat it by a margin the timing noise of five calls does not
          // reach (rows: 5 %, a hybrid split: 2 %)
          int best = 0;
          double tb = h->host_calib_t[0];
          for (int i = 1; i < ncand; i++) {
            const double t = h->host_calib_t[i] * (i == 1 ? 1.05 : 1.02);
            if (t < tb) { tb = t; best = i; }
          }
          if (std::getenv("H1V2_HOST_DEBUG"))
            std::fprintf(stderr, "[h1v2 host path] calibration (ms): assemble %.4f rows %.4f hybrid 1/4 %.4f 1/2 %.4f 3/4 %.4f -> candidate %d\n", h->host_calib_t[0] * 1e3,
                         h->host_calib_t[1] * 1e3, h->host_calib_t[2] * 1e3, h->host_calib_t[3] * 1e3, h->host_calib_t[4] * 1e3, best);
          apply(best);
          h->host_calib = -1;
          h->host_chosen_t = h->host_calib_t[best]; h->host_since = 0;
        }
      }
    }
  } calib_guard{h, t_call0, obs_dev != nullptr, rows_for(0.25f), rows_for(0.5f), rows_for(0.75f)};
  if (!h->d_act) {
    CK(cudaMalloc(&h->d_act, N * 12 * sizeof(float)));
    CK(cudaMalloc(&h->d_rew, N * sizeof(float)));
    CK(cudaMalloc(&h->d_term, N));
    CK(cudaMalloc(&h->d_trunc, N));
  }
  const float* act_dev_ = mapped_alias(const_cast<float*>(actions));
  if (!act_dev_) { CK(cudaMemcpyAsync(h->d_act, actions, N * 12 * sizeof(float), cudaMemcpyHostToDevice, st)); act_dev_ = h->d_act; }
  if (cat && !h->d_dones) CK(cudaMalloc(&h->d_dones, N * sizeof(float)));
  float* rew_dev = cat ? nullptr : mapped_alias(rew);
  uint8_t* term_dev = cat ? h->cat_term : mapped_alias(terminated);
  uint8_t* trunc_dev = mapped_alias(truncated);
  // one launch sequence for both flavours: the plain step, or the step + the constraint apply kernel
  // streaming hand-over: possible when every output of the step kernel lands in mapped host memory (pinned buffers, no constraint tail)
  static const bool no_streaming = std::getenv("H1V2_HOST_NO_STREAMING") != nullptr;  // diagnostics: the round-2 hand-over (synchronise, then the new slots)
  const bool stream_out = !cat && rew_dev && term_dev && trunc_dev && h->host_mode == 1 && !no_streaming;
  auto launch = [&](float* obs_arg, float* sample_arg, int n_rows) -> int {
    if (cat) return launch_cat(h, act_dev_, obs_arg, h->d_rew, h->d_dones, trunc_dev ? trunc_dev : h->d_trunc, st, sample_arg, n_rows);
    return launch_step(h, true, act_dev_, obs_arg, rew_dev ? rew_dev : h->d_rew, term_dev ? term_dev : h->d_term, trunc_dev ? trunc_dev : h->d_trunc, st, sample_arg, false, n_rows,
                       stream_out ? h->h_flags_dev : nullptr, h->host_seq);
  };
  auto copy_back = [&]() -> int {
    if (!rew_dev) CK(cudaMemcpyAsync(rew, h->d_rew, N * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (cat) CK(cudaMemcpyAsync(dones, h->d_dones, N * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (!term_dev) CK(cudaMemcpyAsync(terminated, h->d_term, N, cudaMemcpyDeviceToHost, st));
    if (!trunc_dev) CK(cudaMemcpyAsync(truncated, h->d_trunc, N, cudaMemcpyDeviceToHost, st));
    return 0;
  };
  if (h->host_mode == 1) {
    const int H = h->P.H;
    const size_t ring_bytes = N * H * H1V2_HIST_STRIDE * sizeof(float);
    if (!h->h_sample) {
      CK(cudaHostAlloc(&h->h_sample, N * H1V2_HIST_STRIDE * sizeof(float), cudaHostAllocMapped));
      CK(cudaHostGetDevicePointer(&h->h_sample_dev, h->h_sample, 0));
      h->h_ring = (float*)std::aligned_alloc(64, (ring_bytes + 63) / 64 * 64);
      if (!h->h_ring) return fail("h1v2_step_host: out of host memory");
      const size_t nblk = (N + h->P.epw - 1) / h->P.epw;
      CK(cudaHostAlloc(&h->h_flags, (nblk + 1) * sizeof(unsigned), cudaHostAllocMapped));  // one word per warp + one for the launch's bookkeeping
      CK(cudaHostGetDevicePointer(&h->h_flags_dev, h->h_flags, 0));
      std::memset(h->h_flags, 0, (nblk + 1) * sizeof(unsigned));
    }
    if (!h->ring_valid) {  // first call, or the device path ran in between: fetch the ring once, and the head from the device
      unsigned long long head_counter = 0;  // (a CUDA graph may have replayed steps this library never saw on the host)
      CK(cudaMemcpyAsync(h->h_ring, h->S.hist, ring_bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(&head_counter, h->S.counters + 1, sizeof(head_counter), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      h->hist_launches = head_counter;
    }
    HostPool* p = h->pool;
    const int head = (int)((h->hist_launches + 1) % (uint64_t)H);  // the slot this launch writes (step_kernel: counters[1] + 1)
    // hybrid: the envs below host_rows get their rows from the kernel (zero-copy into the pinned caller buffer), the pool takes the rest
    if (stream_out && ++h->host_seq == 0) h->host_seq = 1;  // 0 is what the flag words start with
    if (launch(h->host_rows > 0 ? obs_dev : nullptr, h->h_sample_dev, h->host_rows) != 0 || copy_back() != 0) return -1;
    p->sample = h->h_sample; p->ring = h->h_ring; p->obs = obs; p->n = h->n; p->H = H; p->head = head; p->e0 = h->host_rows;
    p->flags = stream_out ? h->h_flags : nullptr; p->seq = h->host_seq; p->epw = h->P.epw; p->failed.store(0, std::memory_order_relaxed);
    p->done.store(0, std::memory_order_relaxed);
    const uint64_t gen = p->gen.load(std::memory_order_relaxed) + 1;
    p->gen.store(gen, std::memory_order_release);  // publishes the job to the spinning workers
    {
      std::lock_guard<std::mutex> lk(p->m);  // a worker that went to sleep checks gen under this mutex: no lost wake-up
      if (p->sleepers.load(std::memory_order_relaxed) > 0) p->cv.notify_all();
    }
    static const bool dbg = std::getenv("H1V2_HOST_DEBUG") != nullptr;  // breakdown of a call: tools/diag_e2e.py
    static double acc[5] = {0, 0, 0, 0, 0};
    static int nacc = 0;
    const auto t1 = std::chrono::steady_clock::now();
    assemble_old(p, 0);  // the caller's thread is worker 0
    const auto t2 = std::chrono::steady_clock::now();
    cudaError_t e = cudaSuccess;
    auto t3 = t2, t4 = t2;
    if (stream_out) {
      // warp by warp as the kernel reports them; the rows the kernel wrote itself (hybrid) are complete once their warps have reported
      assemble_new_streaming(p, 0);
      t3 = std::chrono::steady_clock::now();
      for (int w = 0; w < (h->host_rows + h->P.epw - 1) / h->P.epw && wait_flag(p, w); w++) {}
      t4 = std::chrono::steady_clock::now();
      while (p->done.load(std::memory_order_acquire) != p->nthreads - 1) CPU_PAUSE();
      wait_flag(p, (h->n + h->P.epw - 1) / h->P.epw);  // the last block's bookkeeping (log vector, counters) has been published too
      // Every warp and the last block have reported: all outputs are in the caller's memory, all device-side state of the step is
      // written; what is left of the launch is the kernel's exit.  Waiting for the stream here would cost the driver's completion
      // latency (~20 us) for nothing anyone can see; the stream-taking entry points still order themselves after this launch
      // (order_after_host).  A kernel that faulted never reports: the timeout above ends the wait, the synchronise then names the error.
      if (p->failed.load(std::memory_order_relaxed)) {
        e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaErrorLaunchTimeout;
      } else {
        e = cudaEventRecord(h->host_ev, st);
        h->host_pending = e == cudaSuccess;
      }
    } else {
      e = cudaStreamSynchronize(st);
      t3 = std::chrono::steady_clock::now();
      p->phase2.store(gen, std::memory_order_release);  // release the workers even on failure
      if (e == cudaSuccess) assemble_new(p, 0);
      t4 = std::chrono::steady_clock::now();
      while (p->done.load(std::memory_order_acquire) != p->nthreads - 1) CPU_PAUSE();
    }
    if (dbg) {
      const auto t5 = std::chrono::steady_clock::now();
      auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
      acc[0] += us(t_call0, t1); acc[1] += us(t1, t2); acc[2] += us(t2, t3); acc[3] += us(t3, t4); acc[4] += us(t4, t5);
      if (++nacc == 100) {
        std::fprintf(stderr, "[h1v2 host path] per call: enqueue %.1f us | old slots (worker 0) %.1f | wait for the stream (streaming: new slots as the warps report) %.1f | new slot (streaming: row warps) %.1f | pool + final sync %.1f\n",
                     acc[0] / 100, acc[1] / 100, acc[2] / 100, acc[3] / 100, acc[4] / 100);
        nacc = 0; acc[0] = acc[1] = acc[2] = acc[3] = acc[4] = 0;
      }
    }
    if (e != cudaSuccess) { h->ring_valid = false; return fail(std::string("h1v2_step_host: ") + cudaGetErrorString(e)); }
    h->ring_valid = true;
    return 0;
  }
  // mode 0.  Pinned caller buffers are written by the kernel itself (zero-copy over PCIe): the 1.8 KB observation row of an env
  // leaves the GPU as soon as its warp has emitted it, overlapping the transfer with the rest of the step instead of
  // serialising a 450-float-per-env D2H copy behind the kernel.  Pageable buffers take the staged path.
  if (!obs_dev && !h->d_obs) CK(cudaMalloc(&h->d_obs, N * od * sizeof(float)));
  if (launch(obs_dev ? obs_dev : h->d_obs, nullptr, 0) != 0) return -1;
  h->last_stream_set = false;  // nothing of this launch is left in flight after the synchronise below
  if (!obs_dev) CK(cudaMemcpyAsync(obs, h->d_obs, N * od * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (copy_back() != 0) return -1;
  CK(cudaStreamSynchronize(st));
  return 0;
}

extern "C" {

int h1v2_step_host(H1v2Handle* h, const float* actions, float* obs, float* rew, uint8_t* terminated, uint8_t* truncated) {
  if (!h || !actions || !obs || !rew || !terminated || !truncated) return fail("h1v2_step_host: bad arguments");
  return step_host_impl(h, actions, obs, rew, terminated, nullptr, truncated);
}
int h1v2_cat_step_host(H1v2Handle* h, const float* actions, float* obs, float* rew, float* dones, uint8_t* truncated) {
  if (!h || !actions || !obs || !rew || !dones || !truncated) return fail("h1v2_cat_step_host: bad arguments");
  if (!h->cfg.cat_enable) return fail("h1v2_cat_step_host: the handle was created without cfg.cat_enable");
  return step_host_impl(h, actions, obs, rew, nullptr, dones, truncated);
}

int h1v2_host_path_info(const H1v2Handle* h, int32_t* mode, int32_t* threads) {
  if (!h || !mode || !threads) return fail("h1v2_host_path_info: bad arguments");
  *mode = h->host_mode;
  *threads = h->pool ? h->pool->nthreads : 0;
  return 0;
}
int h1v2_host_path_rows(const H1v2Handle* h) { return h ? (h->host_mode == 0 ? h->n : (h->host_mode == 1 ? h->host_rows : -1)) : -1; }

int h1v2_cat_step(H1v2Handle* h, const float* actions, float* obs, float* rew, float* dones, uint8_t* truncated, void* cuda_stream) {
  if (!h || !actions || !obs || !rew || !dones || !truncated) return fail("h1v2_cat_step: bad arguments");
  if (h->rough) return fail("h1v2_cat_step: not available on the rough-terrain instantiation");
  if (!h->cfg.cat_enable) return fail("h1v2_cat_step: the handle was created without cfg.cat_enable");
  return launch_cat(h, actions, obs, rew, dones, truncated, (cudaStream_t)cuda_stream, nullptr);
}
int h1v2_set_constraint_max_p(H1v2Handle* h, const float* max_p) {
  if (!h || !max_p) return fail("h1v2_set_constraint_max_p: bad arguments");
  for (int t = 0; t < H1V2_NUM_CSTR; t++) {
    if (!(max_p[t] >= 0.f && max_p[t] <= 1.f)) return fail("h1v2_set_constraint_max_p: probabilities must be in [0, 1]");
    h->cfg.cat_max_p[t] = max_p[t];
  }
  return 0;
}
int h1v2_cat_debug(H1v2Handle* h, float* raw, float* probs, float* running_max) {
  if (!h || !h->cfg.cat_enable) return fail("h1v2_cat_debug: bad arguments");
  DeviceGuard guard(h->device);
  CK(cudaDeviceSynchronize());
  const size_t cols = (size_t)H1V2_CSTR_COLS * h->n * sizeof(float);
  if (raw) CK(cudaMemcpy(raw, h->cat.k.raw, cols, cudaMemcpyDeviceToHost));
  if (probs) CK(cudaMemcpy(probs, h->cat.probs, cols, cudaMemcpyDeviceToHost));
  if (running_max) {
    int ctl[4];
    CK(cudaMemcpy(ctl, h->cat.k.ctl, sizeof(ctl), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(running_max, h->cat.rmax + (ctl[1] & 1) * H1V2_CSTR_COLS, sizeof(float) * H1V2_CSTR_COLS, cudaMemcpyDeviceToHost));
  }
  return 0;
}
int h1v2_get_cat_log(H1v2Handle* h, const float** acc_dev) {
  if (!h || !acc_dev || !h->cfg.cat_enable) return fail("h1v2_get_cat_log: bad arguments");
  *acc_dev = h->cat.k.logacc;
  return 0;
}
int h1v2_get_cat_log_host(H1v2Handle* h, float* out) {
  if (!h || !out || !h->cfg.cat_enable) return fail("h1v2_get_cat_log_host: bad arguments");
  DeviceGuard guard(h->device);
  CK(cudaDeviceSynchronize());
  float acc[2 * H1V2_NUM_CSTR + 1];
  CK(cudaMemcpy(acc, h->cat.k.logacc, sizeof(acc), cudaMemcpyDeviceToHost));
  const float cnt = acc[2 * H1V2_NUM_CSTR];
  if (cnt > 0.f) {  // upstream writes the keys only in steps with a reset; keep the last such values otherwise
    for (int t = 0; t < 2 * H1V2_NUM_CSTR; t++) h->cat_log[t] = acc[t] / cnt;
    h->cat_log[2 * H1V2_NUM_CSTR] = cnt;
  }
  for (int t = 0; t < 2 * H1V2_NUM_CSTR + 1; t++) out[t] = h->cat_log[t];
  return 0;
}

int h1v2_set_reward_weights(H1v2Handle* h, const float* weights) {
  if (!h || !weights) return fail("h1v2_set_reward_weights: bad arguments");
  for (int t = 0; t < H1V2_NUM_REW; t++) {
    if (!std::isfinite(weights[t])) return fail("h1v2_set_reward_weights: non-finite weight");
    h->cfg.rew_weight[t] = weights[t];
    h->P.w[t] = weights[t];  // the parameter block travels by value with every launch
  }
  return 0;
}
int h1v2_terrain_dims(const H1v2Handle* h, int32_t dims[2]) {
  if (!h || !dims || !h->rough) return fail("h1v2_terrain_dims: the handle has no terrain (cfg.terrain_enable)");
  dims[0] = h->P.t_gx; dims[1] = h->P.t_gy;
  return 0;
}
int h1v2_get_terrain(H1v2Handle* h, float* heights) {
  if (!h || !heights || !h->rough) return fail("h1v2_get_terrain: the handle has no terrain (cfg.terrain_enable)");
  DeviceGuard guard(h->device);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(heights, h->d_terrain, sizeof(float) * (size_t)h->P.t_gx * h->P.t_gy, cudaMemcpyDeviceToHost));
  return 0;
}
int h1v2_set_terrain(H1v2Handle* h, const float* heights) {
  if (!h || !heights || !h->rough) return fail("h1v2_set_terrain: the handle has no terrain (cfg.terrain_enable)");
  DeviceGuard guard(h->device);
  for (size_t i = 0, n = (size_t)h->P.t_gx * h->P.t_gy; i < n; i++)
    if (!std::isfinite(heights[i])) return fail("h1v2_set_terrain: non-finite height");
  CK(cudaDeviceSynchronize());
  return terrain_upload(h, heights);
}
int h1v2_get_terrain_log(H1v2Handle* h, const float** log_dev) {
  if (!h || !log_dev || !h->rough) return fail("h1v2_get_terrain_log: the handle has no terrain (cfg.terrain_enable)");
  *log_dev = h->S.tlog;
  return 0;
}
int h1v2_get_state(H1v2Handle* h, const H1v2State* dst, void* cuda_stream) {
  if (!h || !dst) return fail("h1v2_get_state: bad arguments");
  DeviceGuard guard(h->device);
  order_after_host(h, (cudaStream_t)cuda_stream);
  state_io_kernel<<<(h->n + 63) / 64, 64, 0, (cudaStream_t)cuda_stream>>>(h->P, h->S, *dst, 0);
  note_stream(h, (cudaStream_t)cuda_stream);
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}
int h1v2_set_state(H1v2Handle* h, const H1v2State* src, void* cuda_stream) {
  if (!h || !src) return fail("h1v2_set_state: bad arguments");
  DeviceGuard guard(h->device);
  order_after_host(h, (cudaStream_t)cuda_stream);
  state_io_kernel<<<(h->n + 63) / 64, 64, 0, (cudaStream_t)cuda_stream>>>(h->P, h->S, *src, 1);
  note_stream(h, (cudaStream_t)cuda_stream);
  h->ring_valid = false;  // may have replaced the observation history
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}
int h1v2_get_log(H1v2Handle* h, const float** log_dev) {
  if (!h || !log_dev) return fail("h1v2_get_log: bad arguments");
  *log_dev = h->S.log;
  return 0;
}
#ifdef H1V2_WARPCLOCK
// diagnostic variant only (not declared in the header, not in the product library): tools/diag_warpclock.py
extern "C" int h1v2_debug_warpclock(unsigned long long* out, int nwarps) {
  return cudaMemcpyFromSymbol(out, h1v2::g_warpclock, sizeof(unsigned long long) * 4 * (size_t)nwarps) == cudaSuccess ? 0 : -1;
}
#endif
int h1v2_debug_iter_hist(H1v2Handle* h, float* hist32) {
  if (!h || !hist32) return fail("h1v2_debug_iter_hist: bad arguments");
  DeviceGuard guard(h->device);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hist32, h->S.acc + H1V2_LOG_DIM, sizeof(float) * 32, cudaMemcpyDeviceToHost));
  return 0;
}
int h1v2_get_log_host(H1v2Handle* h, float* log_host) {
  if (!h || !log_host) return fail("h1v2_get_log_host: bad arguments");
  DeviceGuard guard(h->device);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(log_host, h->S.log, sizeof(float) * H1V2_LOG_DIM, cudaMemcpyDeviceToHost));
  return 0;
}
int h1v2_measure_fp32_peak(int32_t device, float* tflops) {
  if (!tflops) return fail("h1v2_measure_fp32_peak: bad arguments");
  DeviceGuard guard(device);
  if (!guard.ok) return fail("h1v2_measure_fp32_peak: no such CUDA device");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  float* buf = nullptr;
  CK(cudaMalloc(&buf, sizeof(float) * blocks * threads));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 0.f;
  for (int rep = 0; rep < 6; rep++) {
    CK(cudaEventRecord(e0));
    fma_peak_kernel<<<blocks, threads>>>(buf, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    best = fmaxf(best, (float)(fl / (ms * 1e-3) / 1e12));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
  *tflops = best;
  return 0;
}

int h1v2_random_actions(H1v2Handle* h, float* actions, uint64_t step, void* cuda_stream) {
  if (!h || !actions) return fail("h1v2_random_actions: bad arguments");
  DeviceGuard guard(h->device);
  order_after_host(h, (cudaStream_t)cuda_stream);
  random_actions_kernel<<<(h->n + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(h->P, actions, step);
  note_stream(h, (cudaStream_t)cuda_stream);
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"
