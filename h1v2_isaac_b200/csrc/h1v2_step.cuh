// h1v2_step.cuh -- the fused control-step kernel: action -> 4 x (PD actuator + physics substep + contact sensor)
// -> terminations -> rewards -> reset -> command -> observation, one launch, no host sync.
// Step order follows packages/biped_tasks/biped_tasks/utils/cat/cat_env.py:95-193 (vendored ManagerBasedRLEnv.step).
#pragma once
#include "h1v2_physics.cuh"

namespace h1v2 {

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

struct RootDerived {
  M3 R;
  V3 vb, vw, wb, ww, g;  // vb / vw: velocity of the root link's COM (base / world frame), as isaaclab's root_lin_vel_b / _w
  real hx, hy;  // forward axis of the base projected on the ground, unnormalised: heading = atan2(hy, hx)
};
__device__ __forceinline__ RootDerived root_derived(const KParams& P, const real (&rq)[4], const real (&rv)[3], const real (&rw)[3]) {
  RootDerived d;
  d.R = quat2mat(rq[0], rq[1], rq[2], rq[3]);
  d.wb = mk3(rw[0], rw[1], rw[2]);
  // the state holds the pelvis ORIGIN velocity (MuJoCo qvel); the managers read the root link's COM velocity: v + w x r
  d.vb = mulTv(d.R, mk3(rv[0], rv[1], rv[2])) + cross(d.wb, mk3(P.root_com[0], P.root_com[1], P.root_com[2]));
  d.vw = mulv(d.R, d.vb);
  d.ww = mulv(d.R, d.wb);
  d.g = mk3(-d.R.cx.z, -d.R.cy.z, -d.R.cz.z);  // R^T (0,0,-1)
  d.hx = d.R.cx.x; d.hy = d.R.cx.y;
  return d;
}

// world linear velocity of the ankle_roll_link origin of this lane's leg (body_lin_vel_w of the foot, feet_slide).
// Rolled over the joints with q / qd staged in the lane's shared-memory column: call it while the column is free
// (after the physics loop, before the history prefetch).
__device__ __forceinline__ V3 foot_velocity(const KLeg& LG, unsigned tid, const M3& R0, V3 om0, V3 v0, const real (&q)[6], const real (&qd)[6], bool at_com, real& ankle_z) {
  extern __shared__ __align__(16) real smem_raw[];
  const Smem sm{smem_raw + tid};
#pragma unroll
  for (int j = 0; j < 6; j++) { sm.jf(j, F_XQ) = q[j]; sm.jf(j, F_FLC) = qd[j]; }
  M3 R = R0;
  V3 x = mk3(0.f, 0.f, 0.f), om = om0, vo = v0;
#pragma unroll 1
  for (int i = 0; i < 6; i++) {
    const int ax = joint_axis(i);
    const real qdi = sm.jf(i, F_FLC);
    x = x + mulv(R, ld3(LG.pos[i]));
    const V3 wi = axis_rt(R, ax);
    om = fma3(wi, qdi, om);
    vo = fma3(cross(x, wi), qdi, vo);
    real s_, c_;
    sincos_lim(sm.jf(i, F_XQ), s_, c_);
    rotate_rt(R, ax, s_, c_);
  }
  ankle_z = x.z;  // ankle_roll_link origin relative to the pelvis origin, world axes (foot_clearance's body_link_pos_w)
  if (at_com) x = x + mulv(R, ld3(LG.ipos[5]));  // body_lin_vel_w is the velocity of the link's COM (isaaclab ArticulationData)
  return vo + cross(om, x);
}

struct CmdState {
  float c[3], heading_target, time_left, m_xy, m_yaw;
  int flags;
};
__device__ __forceinline__ void resample_command(const KParams& P, CmdState& c, int64_t gid, unsigned long long step, uint32_t block0) {
  float u[4], v[4];
  rng4(P.key0, gid, step, STREAM_CMD, block0, u);
  rng4(P.key0, gid, step, STREAM_CMD, block0 + 1, v);
  c.time_left = uni(v[2], P.c_rt[0], P.c_rt[1]);
  c.c[0] = uni(u[0], P.c_lx[0], P.c_lx[1]);
  c.c[1] = uni(u[1], P.c_ly[0], P.c_ly[1]);
  c.c[2] = uni(u[2], P.c_wz[0], P.c_wz[1]);
  if (P.heading_cmd) {
    c.heading_target = uni(u[3], P.c_hd[0], P.c_hd[1]);
    c.flags = (c.flags & ~FLAG_HEADING) | (v[0] <= P.rel_heading ? FLAG_HEADING : 0);
  }
  c.flags = (c.flags & ~FLAG_STANDING) | (v[1] <= P.rel_standing ? FLAG_STANDING : 0);
}

// reset of one env (reset_root_state_uniform, reset_joints_by_scale, manager resets); both lanes compute the root
__device__ __forceinline__ void reset_env(const KParams& P, int side, int64_t gid, unsigned long long step, real (&rp)[3],
                                          real (&rq)[4], real (&rv)[3], real (&rw)[3], real (&q)[6], real (&qd)[6],
                                          real (&la)[6], real (&T1)[6], real (&T2)[6], float4& timers, CmdState& cmd,
                                          real& push_left) {
  float u0[4], u1[4], u2[4], u3[4];
  rng4(P.key0, gid, step, STREAM_RESET, 0, u0);
  rng4(P.key0, gid, step, STREAM_RESET, 1, u1);
  rng4(P.key0, gid, step, STREAM_RESET, 2, u2);
  rng4(P.key0, gid, step, STREAM_RESET, 3, u3);
  int lag = P.min_delay + (int)(u1[2] * (real)(P.max_delay - P.min_delay + 1));
  lag = min(lag, P.max_delay);
  cmd.flags = (cmd.flags & FLAG_TERRAIN_MASK) | FLAG_DELAY_FRESH | FLAG_HIST_FRESH | (lag << FLAG_LAG_SHIFT);  // the env keeps its terrain tile
  timers = make_float4(0.f, 0.f, 0.f, 0.f);
  rp[0] = uni(u0[0], P.rp[0][0], P.rp[0][1]);
  rp[1] = uni(u0[1], P.rp[1][0], P.rp[1][1]);
  rp[2] = P.init_h + uni(u0[3], P.rp[2][0], P.rp[2][1]);
  real roll = uni(u1[0], P.rp[3][0], P.rp[3][1]), pitch = uni(u1[1], P.rp[4][0], P.rp[4][1]);
  real yaw = uni(u0[2], P.rp[5][0], P.rp[5][1]);
  real sr, cr, sp, cp, sy, cy;
  sincos_lim(0.5f * roll, sr, cr); sincos_lim(0.5f * pitch, sp, cp); sincos_lim(0.5f * yaw, sy, cy);
  rq[0] = cy * cr * cp + sy * sr * sp;
  rq[1] = cy * sr * cp - sy * cr * sp;
  rq[2] = cy * cr * sp + sy * sr * cp;
  rq[3] = sy * cr * cp - cy * sr * sp;
  V3 vw = mk3(uni(u2[0], P.rv[0][0], P.rv[0][1]), uni(u2[1], P.rv[1][0], P.rv[1][1]), uni(u2[2], P.rv[2][0], P.rv[2][1]));
  V3 ww = mk3(uni(u3[0], P.rv[3][0], P.rv[3][1]), uni(u3[1], P.rv[4][0], P.rv[4][1]), uni(u3[2], P.rv[5][0], P.rv[5][1]));
  M3 R = quat2mat(rq[0], rq[1], rq[2], rq[3]);
  V3 wb = mulTv(R, ww);
  rv[0] = vw.x; rv[1] = vw.y; rv[2] = vw.z;
  rw[0] = wb.x; rw[1] = wb.y; rw[2] = wb.z;
#pragma unroll
  for (int b = 0; b < 3; b++) {
    float uj[4];
    rng4(P.key0, gid, step, STREAM_RESET, 4 + 3 * side + b, uj);
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int k = 2 * b + e, j = 6 * side + k;
      real spos = uni(uj[2 * e], P.rjp[0], P.rjp[1]);
      real qq = P.q0[j] * spos;
      q[k] = r_min(r_max(qq, P.soft_lo[j]), P.soft_hi[j]);
      qd[k] = 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; k++) la[k] = T1[k] = T2[k] = 0.f;
  cmd.m_xy = cmd.m_yaw = 0.f;
  resample_command(P, cmd, gid, step, 0);
  if (P.push_enable) {
    float up[4];
    rng4(P.key0, gid, step, STREAM_EVENT, 1, up);
    push_left = uni(up[0], P.push_int[0], P.push_int[1]);
  }
}

// CommandTerm.compute(dt) of UniformVelocityCommand (V/velocity_env_cfg.py:90-104; T/utils/mdp/commands.py:47-59)
// Returns true when the env ends the step inside the command dead zone (class 1 only; counted for the next step's balancing).
__device__ __forceinline__ bool update_command(const KParams& P, CmdState& c, const RootDerived& rd, int64_t gid,
                                               unsigned long long step, unsigned dz_prev) {
  real ex = c.c[0] - rd.vb.x, ey = c.c[1] - rd.vb.y;
  c.m_xy += r_sqrt(ex * ex + ey * ey) / P.max_command_step;
  c.m_yaw += r_abs(c.c[2] - rd.wb.z) / P.max_command_step;
  c.time_left -= P.step_dt;
  if (c.time_left <= 0.0f) resample_command(P, c, gid, step, 2);
  if (P.heading_cmd && (c.flags & FLAG_HEADING)) {
    real err = wrap_to_pi(c.heading_target - r_atan2(rd.hy, rd.hx));
    c.c[2] = r_min(r_max(P.k_heading * err, P.c_wz[0]), P.c_wz[1]);
  }
  if (P.cmd_class == 0) {
    if (c.flags & FLAG_STANDING) c.c[0] = c.c[1] = c.c[2] = 0.f;
    return false;
  }
  // UniformVelocityCommandWithDeadzone._update_command (T/utils/mdp/commands.py:41-96).  The override never zeroes standing
  // envs.  Balancing (:62-83): the reference counts the envs whose |cmd_xy| is inside the dead zone and moves exactly
  // |n // 2 - count| envs, picked by randperm, across it: active ones get cmd_xy = 0, dead-zone ones a fresh command
  // (_resample: new command AND new time_left).  Here every env draws on its own with the probability that moves the same
  // number in expectation, from the count published at the end of the PREVIOUS step (one launch per step: no grid-wide count
  // inside it).  With velocity_deadzone == 0 (C12/rsl_env_cfg.py:98) no env is ever inside, the count is 0 and every env loses
  // its xy command with probability (n // 2) / n on every step.  Then the yaw-rate command flips sign with probability
  // physics_dt / max_episode_length_s (:85-96).
  float u[4];
  rng4(P.key0, gid, step, STREAM_CMD, 4, u);
  const int target = P.n / 2, cur = (int)min(dz_prev, (unsigned)P.n);
  // |cmd_xy| < dead zone, as one multiply + one fused multiply-add against the squared threshold: the oracle does the same, bit for bit
  const float dz2 = __fmul_rn(P.deadzone, P.deadzone);
  bool in_dz = __fmaf_rn(c.c[1], c.c[1], __fmul_rn(c.c[0], c.c[0])) < dz2;
  if (cur < target) {
    if (!in_dz && u[0] < __fdiv_rn((real)(target - cur), (real)(P.n - cur))) c.c[0] = c.c[1] = 0.f;
  } else if (cur > target) {
    if (in_dz && u[0] < __fdiv_rn((real)(cur - target), (real)cur)) resample_command(P, c, gid, step, 6);
  }
  if (u[1] < P.flip_prob) c.c[2] = -c.c[2];
  return __fmaf_rn(c.c[1], c.c[1], __fmul_rn(c.c[0], c.c[0])) < dz2;
}

// ---- history rings of the warp's 16 envs: global -> shared memory, asynchronously ----
// The rings of consecutive envs are contiguous in HBM ([N][H][48] floats), so the warp's block is ONE contiguous, 16-byte
// aligned range: a single bulk copy (cp.async.bulk, the 1-D TMA path: UBLKCP in SASS) issued by one lane moves up to 26.9 KB
// with no registers and no per-lane instructions, and signals an mbarrier with the byte count when it has landed.  (Round 1
// issued ~60 16-byte cp.async per lane, ~1 900 warp-instructions per step in a kernel bound by instruction issue.)  The copy
// is started right after the physics (the per-thread columns are dead by then) and awaited only when the observation is
// emitted, so the HBM latency hides behind the reward / termination / reset / command code (profiles/r1k: with register-staged
// loads the emission was 27 % of the step, 76 % of it long-scoreboard stalls).
// The mbarrier lives in the last 8 bytes of the (float-build) column window: a chunk of whole rings never reaches them
// (7040 floats are not a multiple of the 48-float slot), and physics only uses them as lane 30/31's fifth contact point.
#define H1V2_MBAR_OFFSET (SMEM_FLOATS * H1V2_BLOCK * 4 - 8)
__device__ __forceinline__ int hist_envs_per_chunk(int H, int epw) { return min(epw, (SMEM_FLOATS * H1V2_BLOCK) / (H * H1V2_HIST_STRIDE)); }
// chunk: 0 for the first copy of a launch (initialises the barrier), 1 for the second (H = 10 at 16 envs per warp: 14 + 2 rings)
__device__ __forceinline__ void hist_prefetch(const KParams& P, const KState& S, unsigned tid, unsigned bid, int e0, int chunk) {
  extern __shared__ __align__(16) real smem_raw[];
  const int lane = tid & 31;
  const int warp_env0 = (int)bid * P.epw;
  const int ne = min(hist_envs_per_chunk(P.H, P.epw), min(P.epw, P.n - warp_env0) - e0);
  const unsigned bytes = (unsigned)(ne * P.H * H1V2_HIST_STRIDE * 4);
  const float* src = S.hist + (size_t)(warp_env0 + e0) * P.H * H1V2_HIST_STRIDE;
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_raw), mbar = dst + H1V2_MBAR_OFFSET;
  // every lane's earlier (generic-proxy) accesses to the window are ordered before the async-proxy writes of the copy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    if (chunk == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
  }
}
// wait until the bytes of copy `chunk` have landed (phase parity = chunk & 1); every lane waits for itself
__device__ __forceinline__ void hist_wait(int chunk) {
  extern __shared__ __align__(16) real smem_raw[];
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem_raw) + H1V2_MBAR_OFFSET;
  unsigned ok = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(mbar), "r"((unsigned)(chunk & 1)) : "memory");
  } while (!ok);
}

// observation sample of this step -> history ring slot `head`; then the warp cooperatively emits the flattened rows
__device__ __forceinline__ void emit_observation(const KParams& P, const KState& S, unsigned tid, unsigned bid, int env, int side, bool valid, int64_t gid,
                                                 unsigned long long step, int head, const RootDerived& rd, const CmdState& cmd,
                                                 const real (&q)[6], const real (&qd)[6], const real (&la)[6], float* obs) {
  extern __shared__ __align__(16) real smem_raw[];
  const int H = P.H;
  // ---- the sample: 18 values of this lane's leg, 9 root values on the env's first lane (noise: V/velocity_env_cfg.py:124-131) ----
  real sv[18], s9[9];
  {
    real nq[6] = {0, 0, 0, 0, 0, 0}, nv[6] = {0, 0, 0, 0, 0, 0};
    if (P.corrupt) {
      float a[4], b[4], d[4];
      rng4(P.key0, gid, step, STREAM_OBS, 2 + 4 * side, a);
      rng4(P.key0, gid, step, STREAM_OBS, 3 + 4 * side, b);
      rng4(P.key0, gid, step, STREAM_OBS, 4 + 4 * side, d);
      nq[0] = a[0]; nq[1] = a[1]; nq[2] = a[2]; nq[3] = a[3]; nq[4] = b[0]; nq[5] = b[1];
      nv[0] = b[2]; nv[1] = b[3]; nv[2] = d[0]; nv[3] = d[1]; nv[4] = d[2]; nv[5] = d[3];
#pragma unroll
      for (int k = 0; k < 6; k++) { nq[k] = uni(nq[k], -P.n_q, P.n_q); nv[k] = uni(nv[k], -P.n_v, P.n_v); }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int j = 6 * side + k;
      sv[k] = (q[k] - P.q0[j] + nq[k]) * P.s_q;
      sv[6 + k] = (qd[k] + nv[k]) * P.s_v;
      sv[12 + k] = la[k] * P.s_a;
    }
    real n0[3] = {0, 0, 0}, n1[3] = {0, 0, 0};
    if (P.corrupt && side == 0) {
      float b0[4], b1[4];
      rng4(P.key0, gid, step, STREAM_OBS, 0, b0);
      rng4(P.key0, gid, step, STREAM_OBS, 1, b1);
#pragma unroll
      for (int i = 0; i < 3; i++) { n0[i] = uni(b0[i], -P.n_av, P.n_av); n1[i] = uni(b1[i], -P.n_g, P.n_g); }
    }
    s9[0] = (rd.wb.x + n0[0]) * P.s_av; s9[1] = (rd.wb.y + n0[1]) * P.s_av; s9[2] = (rd.wb.z + n0[2]) * P.s_av;
    s9[3] = (rd.g.x + n1[0]) * P.s_g; s9[4] = (rd.g.y + n1[1]) * P.s_g; s9[5] = (rd.g.z + n1[2]) * P.s_g;
    s9[6] = cmd.c[0] * P.s_cmd; s9[7] = cmd.c[1] * P.s_cmd; s9[8] = cmd.c[2] * P.s_cmd;
  }
  // Rough id: base_lin_vel (V/velocity_env_cfg.py:123) travels in floats 45..47 of the slot; the flatten table puts it first
  real slv[3] = {0.f, 0.f, 0.f};
  if (P.lin_vel) {
    real nl[3] = {0.f, 0.f, 0.f};
    if (P.corrupt && side == 0) {
      float b[4];
      rng4(P.key0, gid, step, STREAM_OBS, 10, b);
#pragma unroll
      for (int i = 0; i < 3; i++) nl[i] = uni(b[i], -P.n_lv, P.n_lv);
    }
    slv[0] = (rd.vb.x + nl[0]) * P.s_lv; slv[1] = (rd.vb.y + nl[1]) * P.s_lv; slv[2] = (rd.vb.z + nl[2]) * P.s_lv;
  }
  if (valid) {  // ring slot `head` in HBM
    float* slot = S.hist + ((size_t)env * H + head) * H1V2_HIST_STRIDE;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int i = P.inv_perm[6 * side + k];
      slot[9 + i] = sv[k]; slot[21 + i] = sv[6 + k]; slot[33 + i] = sv[12 + k];
    }
    if (side == 0) {
#pragma unroll
      for (int k = 0; k < 9; k++) slot[k] = s9[k];
      if (P.lin_vel) { slot[45] = slv[0]; slot[46] = slv[1]; slot[47] = slv[2]; }
    }
  }
  // ---- cooperative flatten from the shared-memory copy: term-major, oldest -> newest inside each term block
  // (packages/biped_tasks/biped_tasks/utils/history/observation_manager.py:335-355, circular_buffer.py:79-87,131-135).
  // Each lane owns output columns lane, lane+32, ...; their source offset inside an env's ring is the same for every
  // env, so it is decoded once (packed: regular offset | offset in the newest slot << 16; -1 = beyond obs_dim). ----
  const int lane = tid & 31;
  const int warp_env0 = (int)bid * P.epw;
  const int npass = (P.lut_dim + 31) >> 5;
  int off[H1V2_OBS_MAXPASS];
#pragma unroll
  for (int i = 0; i < H1V2_OBS_MAXPASS; i++) {
    const int idx = lane + 32 * i;
    off[i] = -1;
    if (i < npass && idx < P.lut_dim) {
      const int hk = __ldg(S.lut + idx);
      const int hh = hk >> 8, k = hk & 255;
      int sl = head + 1 + hh;
      sl = sl >= H ? sl - H : sl;
      off[i] = (sl * H1V2_HIST_STRIDE + k) | ((head * H1V2_HIST_STRIDE + k) << 16);
    }
  }
  const int nenv = min(P.epw, P.n - warp_env0), epc = hist_envs_per_chunk(H, P.epw), ring = H * H1V2_HIST_STRIDE;
  const int my_e = (int)(lane >> 1);
  // host path (h1v2_step_host): the warps of envs [0, n_rows) write whole rows (PCIe DMA into the caller's pinned buffer), the others
  // only their new sample (host threads assemble those rows meanwhile): the two halves of the host's work run concurrently
  const bool warp_samples = S.sample_out != nullptr && warp_env0 >= S.n_rows;
  if (warp_samples) obs = nullptr;
#pragma unroll 1
  for (int e0 = 0; e0 < nenv; e0 += epc) {
    if (e0 > 0) { __syncwarp(); hist_prefetch(P, S, tid, bid, e0, 1); }  // H = 10 at 16 envs per warp: the 16 rings do not fit at once
    __syncwarp();  // the barrier's initialisation by lane 0 is visible to every lane
    hist_wait(e0 > 0 ? 1 : 0);
    __syncwarp();
    if (my_e >= e0 && my_e < e0 + epc) {  // the new sample replaces the stale slot `head` of this env's copy
      float* sl = reinterpret_cast<float*>(smem_raw) + (my_e - e0) * ring + head * H1V2_HIST_STRIDE;
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const int i = P.inv_perm[6 * side + k];
        sl[9 + i] = sv[k]; sl[21 + i] = sv[6 + k]; sl[33 + i] = sv[12 + k];
      }
      if (side == 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) sl[k] = s9[k];
        if (P.lin_vel) { sl[45] = slv[0]; sl[46] = slv[1]; sl[47] = slv[2]; }
      }
    }
    __syncwarp();
    const int e1 = min(nenv, e0 + epc);
    if (warp_samples) {
      // host path (h1v2_step_host): only what is NEW leaves the GPU -- the 45-float sample of every env of the chunk (+ the
      // first-push flag), 192 B per env instead of the 1.8 KB row the host can assemble from its own copy of the ring.
      // The warp's samples are one contiguous 16-byte-aligned range: coalesced float4 stores (zero-copy over PCIe when pinned).
      const int nq = (e1 - e0) * (H1V2_HIST_STRIDE / 4);
      float4* dst4 = reinterpret_cast<float4*>(S.sample_out + (size_t)(warp_env0 + e0) * H1V2_HIST_STRIDE);
      for (int q0 = 0; q0 < nq; q0 += 32) {
        const int qi = q0 + lane;
        const int e = min(qi / (H1V2_HIST_STRIDE / 4), e1 - e0 - 1), k4 = qi % (H1V2_HIST_STRIDE / 4);
        const int fl = __shfl_sync(0xffffffffu, cmd.flags, 2 * (e0 + e));
        if (qi < nq) {
          float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(smem_raw) + e * ring + head * H1V2_HIST_STRIDE + 4 * k4);
          if (k4 == H1V2_HIST_STRIDE / 4 - 1) v.y = (fl & FLAG_HIST_FRESH) ? 1.f : 0.f;  // float 45 of the slot (45..47 are padding)
          dst4[qi] = v;
        }
      }
    }
#pragma unroll 1
    for (int e = e0; e < e1; e++) {
      const int env_e = warp_env0 + e;
      const bool fresh = (__shfl_sync(0xffffffffu, cmd.flags, 2 * e) & FLAG_HIST_FRESH) != 0;  // warp-uniform
      const float* hsm = reinterpret_cast<const float*>(smem_raw) + (e - e0) * ring;
      float* orow = obs + (size_t)env_e * P.obs_dim + lane;
      const int sh = fresh ? 16 : 0;
#pragma unroll
      for (int i = 0; i < H1V2_OBS_MAXPASS; i++)
        if (off[i] >= 0) {
          const float v = hsm[(off[i] >> sh) & 0xffff];
          if (obs) orow[32 * i] = v;
        }
      if (fresh) {  // back-fill: the whole ring holds the first sample (circular_buffer.py:131-135)
        float* hbase = S.hist + (size_t)env_e * ring;
#pragma unroll
        for (int i = 0; i < H1V2_OBS_MAXPASS; i++)
          if (off[i] >= 0) hbase[off[i] & 0xffff] = hsm[(off[i] >> 16) & 0xffff];
      }
    }
  }
}

// Rough id: height_scan (V/velocity_env_cfg.py:61-68,133-138; upstream mdp.height_scan = sensor z - hit z - offset): GridPattern rays --
// x fastest, "xy" indexing -- about the scanner body (torso_link = the pelvis frame: fixed joint), rotated by the base yaw only, cast
// straight down onto the height field: one terrain lookup per ray.  Noise, then clip, then scale.
// The envs of the warp leave their scan frame (position, yaw, tile) in shared memory (the per-thread columns are dead by now); the
// (env, group of four consecutive rays) pairs are then dealt round-robin to the 32 lanes, so that no lane idles and the lookups of
// several groups are in flight together (the first version walked env by env: 11 % of the instructions but 24 % of the step's
// stall samples, all of it exposed L2 latency; profiles/r3_regions_rough_32768.txt).  Four rays share one Philox draw.
#define SCAN_FRAME 8  // px py pz cos sin | i0 j0 (int bits) | oz
template <bool ROUGH>
__device__ __forceinline__ void emit_height_scan(const KParams& P, const KState& S, unsigned tid, unsigned bid, unsigned long long step,
                                                 const real (&rp)[3], const RootDerived& rd, int flags, float* obs) {
  extern __shared__ __align__(16) real smem_raw[];
  float* fr = reinterpret_cast<float*>(smem_raw);
  const int lane = tid & 31;
  const int warp_env0 = (int)bid * P.epw, nenv = min(P.epw, P.n - warp_env0);
  const int nrays = P.scan_nx * P.scan_ny, ngrp = (nrays + 3) >> 2;
  __syncwarp();  // every lane is done with the history window
  if ((lane & 1) == 0 && (lane >> 1) < nenv) {
    const real hn = r_rsqrt(r_max(rd.hx * rd.hx + rd.hy * rd.hy, 1e-30f));
    float* f = fr + (lane >> 1) * SCAN_FRAME;
    f[0] = (float)rp[0]; f[1] = (float)rp[1]; f[2] = (float)rp[2]; f[3] = (float)(rd.hx * hn); f[4] = (float)(rd.hy * hn);
    TerrainEnv te;
    te.i0 = te.j0 = 0; te.oz = 0.f;
    if (ROUGH) te = terrain_env(P, S.terrain_oz, flags);
    f[5] = __int_as_float(te.i0); f[6] = __int_as_float(te.j0); f[7] = (float)te.oz;
  }
  __syncwarp();
  const int nitem = nenv * ngrp;
  const float inv_ngrp = 1.0f / (float)ngrp, inv_nx = 1.0f / (float)P.scan_nx;
#pragma unroll 1
  for (int it = lane; it < nitem; it += 32) {
    const int e = (int)(((float)it + 0.5f) * inv_ngrp), g = it - e * ngrp;  // exact for these small integers
    const float* f = fr + e * SCAN_FRAME;
    const real px = f[0], py = f[1], pz = f[2], cy = f[3], sy = f[4];
    TerrainEnv te;
    te.i0 = __float_as_int(f[5]); te.j0 = __float_as_int(f[6]); te.oz = f[7];
    // all sixteen loads of the group's four lookups are issued before anything consumes them, and the Philox rounds (inlined here: a
    // call would fence the loads) run while they are in flight
    real hv[4];
    if (ROUGH) {
      float c00[4], c01[4], c10[4], c11[4];
      real uu[4], vv[4];
      bool in[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int r = 4 * g + k;
        const int iy = (int)(((float)r + 0.5f) * inv_nx), ix = r - iy * P.scan_nx;
        const real gx = r_fma((real)ix, P.scan_res, P.scan_x0), gy = r_fma((real)iy, P.scan_res, P.scan_y0);
        const real a = (px + cy * gx - sy * gy + P.t_half) * P.t_inv_hs, b = (py + sy * gx + cy * gy + P.t_half) * P.t_inv_hs;
        const real fa = r_floor(a), fb = r_floor(b);
        const int i = te.i0 + (int)fa, j = te.j0 + (int)fb;
        uu[k] = a - fa; vv[k] = b - fb;
        in[k] = i >= 0 && j >= 0 && i < P.t_gx - 1 && j < P.t_gy - 1;
        const float* p = S.terrain_h + (in[k] ? (size_t)i * P.t_gy + j : 0);
        c00[k] = __ldg(p); c01[k] = __ldg(p + 1); c10[k] = __ldg(p + P.t_gy); c11[k] = __ldg(p + P.t_gy + 1);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {  // terrain_sample's arithmetic (h1v2_terrain.cuh), on the values loaded above
        const bool upper = vv[k] >= uu[k];
        const real dx = upper ? c11[k] - c01[k] : c10[k] - c00[k], dy = upper ? c01[k] - c00[k] : c11[k] - c10[k];
        hv[k] = in[k] ? r_fma(uu[k], dx, r_fma(vv[k], dy, (real)c00[k])) - te.oz : -te.oz;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) hv[k] = 0.f;
    }
    float nz[4] = {0.f, 0.f, 0.f, 0.f};
    if (P.corrupt) {
      uint32_t o[4];
      philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), STREAM_OBS, (uint32_t)(16 + g), P.key0, (uint32_t)(P.env_id_offset + warp_env0 + e), o);
#pragma unroll
      for (int k = 0; k < 4; k++) nz[k] = uni((float)(o[k] >> 8) * (1.0f / 16777216.0f), -P.n_scan, P.n_scan);
    }
    float* orow = obs + (size_t)(warp_env0 + e) * P.obs_dim + P.scan_col0 + 4 * g;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float v = (float)(pz - hv[k] - P.scan_off) + nz[k];
      v = fminf(fmaxf(v, P.scan_lo), P.scan_hi) * P.s_scan;
      if (obs && 4 * g + k < nrays) orow[k] = v;
    }
  }
}

// Constraints-as-Terminations column maximum (constraint.max(0).clamp(min=1e-6), constraint_manager.py:56): reduce over the lanes
// of the same side (xor 2..16 keeps the lane parity), then lanes 0 / 1 publish; positive floats order like their bit patterns.
// Constraints-as-Terminations column maxima (constraint.max(0).clamp(min=1e-6), constraint_manager.py:56).  A candidate only
// matters when it is positive, and positive floats order like their bit patterns: redux.sync (one instruction) over the lanes
// of the same side; the result is then PUBLISHED BY DIFFERENT LANES FOR DIFFERENT COLUMNS (lane 2 i + side takes the i-th
// column of its side), so a warp's 56 filtered atomics go out in two or three parallel rounds instead of 56 dependent round
// trips of lanes 0 / 1 (24 us of a 0.84 ms step at 32768 envs).  All warps of a launch aim at the same 56 words: publish only what
// beats the maximum seen so far (a stale read only costs a redundant atomic).
struct CatMax {
  unsigned v[3];   // this lane's column values of rounds 0..2 (bit patterns of max(value, 0))
  int col[3];
  int n;           // columns seen so far on this lane's side
};
__device__ __forceinline__ void cat_max_init(CatMax& M) { M.n = 0; M.v[0] = M.v[1] = M.v[2] = 0u; M.col[0] = M.col[1] = M.col[2] = 0; }
__device__ __forceinline__ void cat_col_max(CatMax& M, int col, real v, bool valid, unsigned tid) {
  const unsigned bits = (valid && v > 0.f) ? (unsigned)__float_as_int((float)v) : 0u;
  const unsigned m = __reduce_max_sync((tid & 1) ? 0xaaaaaaaau : 0x55555555u, bits);
  const int slot = M.n & 15, round = M.n >> 4;  // compile-time after unrolling
  if ((int)((tid & 31) >> 1) == slot) { M.v[round] = m; M.col[round] = col; }
  M.n++;
}
__device__ __forceinline__ void cat_max_publish(const CatMax& M, int* cmax) {
#pragma unroll
  for (int r = 0; r < 3; r++)
    if (M.v[r] > (unsigned)__float_as_int(1e-6f) && (int)M.v[r] > __ldcg(cmax + M.col[r])) atomicMax(cmax + M.col[r], (int)M.v[r]);
}

// publishes the log vector, clears the accumulators, advances the step counter and the history head.
// Run by the 32 lanes of the LAST block of a step launch to finish (ticket counter S.done): one launch per control step.
__device__ __forceinline__ void finalize_step(const KState& S, bool do_step, bool cat, unsigned t, int n, int epw) {
  if (S.tlog && do_step && t == 0) {  // Curriculum/terrain_levels: mean level over all envs (V/mdp/curriculums.py:52)
    S.tlog[1] = __ldcg(S.tlog) / (float)n;
    S.tlog[0] = 0.f;
  }
  if (do_step && cat) {
    // Constraints-as-Terminations: the envs whose command is inside the no_move dead zone, in ascending env order (the reference's
    // boolean-mask gather, constraints.py:216-222), are addressed by rank: the blocks left one member mask per warp, this block adds
    // the inclusive prefix of their populations; the apply kernel finds member r by a binary search over the prefix and a bit select.
    // Clears the log sums of the previous step (the apply kernel of THIS step accumulates into them next).
    const KCat& T = S.cat;
    if (t < 2 * H1V2_NUM_CSTR + 1) T.logacc[t] = 0.f;
    const int nw = (n + epw - 1) / epw;  // warps of the step launch, one 16-bit member mask each
    // every lane takes a contiguous range of masks; loads are issued eight at a time (a dependent walk through global memory
    // was a 30 us tail of this single warp at 32768 envs, profiles/r2_notes.md)
    const int per = (nw + 31) / 32, a = min(nw, (int)t * per), b = min(nw, a + per);
    int cnt = 0;
    for (int i0 = a; i0 < b; i0 += 8) {
      unsigned m[8];
#pragma unroll
      for (int k = 0; k < 8; k++) m[k] = (i0 + k < b) ? (unsigned)__ldcg(T.dz + i0 + k) : 0u;
#pragma unroll
      for (int k = 0; k < 8; k++) cnt += __popc(m[k]);
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL_MASK, incl, o);
      if ((int)t >= o) incl += v;
    }
    int run = incl - cnt;
    for (int i0 = a; i0 < b; i0 += 8) {
      unsigned m[8];
#pragma unroll
      for (int k = 0; k < 8; k++) m[k] = (i0 + k < b) ? (unsigned)__ldcg(T.dz + i0 + k) : 0u;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        run += __popc(m[k]);
        if (i0 + k < b) T.list[i0 + k] = run;  // members in the masks 0..i (inclusive prefix)
      }
    }
    if (t == 31) T.ctl[0] = incl;
  }
  if (do_step) {
    const real cnt = __ldcg(S.acc + H1V2_LOG_COUNT);
    const real a = __ldcg(S.acc + t);  // t < 32 == H1V2_LOG_DIM
    if (t == H1V2_LOG_COUNT) S.log[t] = a;
    else if (t == H1V2_LOG_NAN_RESETS) S.log[t] += a;
    else if (t == H1V2_LOG_MAX_ITERS) S.log[t] = (real)__float_as_int(a);
    else if (t == H1V2_LOG_CAP_HITS || t == H1V2_LOG_SUM_ITERS || t == H1V2_LOG_CONTACT_OVERFLOW) S.log[t] = a;
    else if (cnt > 0.f) {
      const bool mean = (t >= H1V2_LOG_REW0 && t < H1V2_LOG_REW0 + H1V2_NUM_REW) || t == H1V2_LOG_ERR_XY || t == H1V2_LOG_ERR_YAW;
      S.log[t] = mean ? a / cnt : a;
    }
    __syncwarp();
    S.acc[t] = 0.f;
  }
  if (t == 0) {
    if (do_step) {
      S.counters[0] += 1ull;
      S.counters[3] = __ldcg(S.counters + 2);  // dead-zone census of this step, read by the next one
      S.counters[2] = 0ull;
    }
    S.counters[1] += 1ull;
    *S.done = 0u;
  }
}

// CAT: the Constraints-as-Terminations flavour (h1v2_cat_step) is its own instantiation, so the plain step carries none of its code
// ROUGH: the Rough id's flavour (height-field contacts, base_lin_vel, height scan, terrain-level curriculum), its own instantiation too
#ifdef H1V2_WARPCLOCK
// diagnostic variant only: per warp {globaltimer at entry, clock64 cycles of the physics loop, of the whole kernel, trips | lstrips << 32}
__device__ unsigned long long g_warpclock[4 * 16384];
#endif
// NM > 1: the mirror-lane instantiations -- NM = 2 for 8 envs per warp (2369..5624 envs on a B200, BASELINE configs[1]): lanes 16..31 mirror lanes 0..15;
// NM = 4 for 4 envs per warp (<= 2368 envs): lanes l, l+8, l+16, l+24 --
// instead of shadowing the warp's first env, and share the independent loops of the Newton trip with them (h1v2_physics.cuh)
template <bool DO_STEP, bool CAT = false, bool ROUGH = false, int NM = 1>
__global__ void __launch_bounds__(H1V2_BLOCK) step_kernel(const __grid_constant__ KParams P, const KState S, const float* __restrict__ actions,
                                                  float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ term,
                                                  uint8_t* __restrict__ trunc) {
  // thread / block index through volatile asm: under register pressure ptxas otherwise re-reads the special registers
  // (S2R, ~20 cycles each) all over the Newton loop instead of keeping one value alive (profiles/r1e: 7 % of all instructions)
  unsigned tid, bid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bid));
  // lane pair p of warp w owns env w*epw + p for p < epw; the remaining lanes shadow the warp's first env (their work is
  // bit-identical to it, so they never lengthen the warp's Newton loop) and store nothing
#ifdef H1V2_WARPCLOCK
  unsigned long long wc_g0, wc_c1 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(wc_g0));
  const long long wc_c0 = clock64();
  int wc_trips = 0, wc_ls = 0;
#endif
  const int side = tid & 1;
  constexpr bool QUAD = NM > 1;  // NM mirrors per lane: 2 at 8 envs per warp, 4 at 4 envs per warp
  const int slot = QUAD ? (int)((tid >> 1) & (16u / NM - 1u)) : (int)(tid >> 1);
  const int warp_env0 = (int)bid * P.epw;
  const bool in_range = (QUAD || slot < P.epw) && warp_env0 + slot < P.n;
  const bool valid = in_range && (!QUAD || tid < 32u / NM);  // a mirror lane computes everything and stores nothing
  const int env = in_range ? warp_env0 + slot : min(warp_env0, P.n - 1);
  const int lidx = 2 * env + side;
  const int N = P.n, N2 = 2 * P.n;
  const int64_t gid = P.env_id_offset + env;
  const unsigned long long step = S.counters[0] + (DO_STEP ? 1ull : 0ull);
  const int head = (int)((S.counters[1] + 1ull) % (unsigned long long)P.H);

  // ---- load the physics state (coalesced 128-bit).  The actuator line and the command state are read AFTER the
  //      physics loop: whatever is live across substep() costs registers inside the Newton iteration. ----
  real rp[3], rq[4], rv[3], rw[3], q[6], qd[6];
  real mu, mass_add, push_left;
  int flags0;
  {
    float4 r0 = S.root[env], r1 = S.root[N + env], r2 = S.root[2 * N + env], r3 = S.root[3 * N + env];
    rp[0] = r0.x; rp[1] = r0.y; rp[2] = r0.z; rq[0] = r0.w; rq[1] = r1.x; rq[2] = r1.y; rq[3] = r1.z;
    rv[0] = r1.w; rv[1] = r2.x; rv[2] = r2.y; rw[0] = r2.z; rw[1] = r2.w; rw[2] = r3.x;
    mu = r3.y; mass_add = r3.z; push_left = r3.w;
    float4 l0 = S.leg[lidx], l1 = S.leg[N2 + lidx], l2 = S.leg[2 * N2 + lidx];
    q[0] = l0.x; q[1] = l0.y; q[2] = l0.z; q[3] = l0.w; q[4] = l1.x; q[5] = l1.y;
    qd[0] = l1.z; qd[1] = l1.w; qd[2] = l2.x; qd[3] = l2.y; qd[4] = l2.z; qd[5] = l2.w;
    flags0 = __float_as_int(S.cmd[N + env].w);
  }
  float4 tm = S.timers[lidx];
  if (DO_STEP) {  // warm the L1 with the actuator line and this env's action row: the PD loop re-reads them every substep
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.act + N2 + lidx));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.act + 2 * N2 + lidx));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.act + 3 * N2 + lidx));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.act + 4 * N2 + lidx));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(actions + (size_t)env * 12 + 6 * side));
  }
  real wl[6], wr[6];
  {
    float4 w0 = S.warm[lidx], w1 = S.warm[N2 + lidx], w2 = S.warm[2 * N2 + lidx];
    wl[0] = w0.x; wl[1] = w0.y; wl[2] = w0.z; wl[3] = w0.w; wl[4] = w1.x; wl[5] = w1.y;
    wr[0] = w1.z; wr[1] = w1.w; wr[2] = w2.x; wr[3] = w2.y; wr[4] = w2.z; wr[5] = w2.w;
  }
  TerrainEnv te;
  te.i0 = te.j0 = 0; te.oz = 0.f;
  if (ROUGH) te = terrain_env(P, S.terrain_oz, flags0);
  real la[6], T1[6], T2[6];
  CmdState cmd;
  real h_foot[3] = {0, 0, 0}, h_shin[3] = {0, 0, 0}, h_torso[3] = {0, 0, 0}, h_pelvis[3] = {0, 0, 0};
  real s_tau = 0.f, s_acc = 0.f;
  SubOut so;
  int max_it = 0, ncap = 0, sum_it = 0, novf = 0;

  if (DO_STEP) {
    // ---- physics loop: DelayedPD actuator (A/robots/h12.py:58-113) + substep + ContactSensor at sim dt ----
    // process_action (JointPositionAction, V/velocity_env_cfg.py:111): target = scale * a + q0; the delay line holds the
    // targets of the previous two control steps.  Targets are re-read per substep (L1/L2 hits) instead of held in registers.
    const int lag = (flags0 >> FLAG_LAG_SHIFT) & 7;
    const bool line_fresh = (flags0 & FLAG_DELAY_FRESH) != 0;
    bool use_warm = !line_fresh;
#pragma unroll 1
    for (int k = 0; k < P.decimation; k++) {
      const int age = lag - k;
      real tau[6];
      {
        real T[6];
        if (age <= 0 || line_fresh) {
#pragma unroll
          for (int i = 0; i < 6; i++) T[i] = r_fma(P.action_scale, __ldg(actions + (size_t)env * 12 + P.inv_perm[6 * side + i]), P.q0[6 * side + i]);
        } else if (age <= P.decimation) {
          const float4 a1 = S.act[N2 + lidx], a2 = S.act[2 * N2 + lidx];
          T[0] = a1.z; T[1] = a1.w; T[2] = a2.x; T[3] = a2.y; T[4] = a2.z; T[5] = a2.w;
        } else {
          const float4 a3 = S.act[3 * N2 + lidx], a4 = S.act[4 * N2 + lidx];
          T[0] = a3.x; T[1] = a3.y; T[2] = a3.z; T[3] = a3.w; T[4] = a4.x; T[5] = a4.y;
        }
        s_tau = 0.f;
#pragma unroll
        for (int i = 0; i < 6; i++) {
          const int j = 6 * side + i;
          const real t = P.kp[j] * (T[i] - q[i]) + P.kd[j] * (0.f - qd[i]);
          tau[i] = r_min(r_max(t, -P.effort[j]), P.effort[j]);
          if ((P.m_tau >> j) & 1u) s_tau = r_fma(tau[i], tau[i], s_tau);
        }
        if (S.diag && valid) {
          float* dg = S.diag + (size_t)env * H1V2_DIAG_DIM;
#pragma unroll
          for (int i = 0; i < 6; i++) dg[36 + 6 * side + i] = tau[i];
        }
        if (CAT && k == P.decimation - 1) {  // joint_torque_limits (constraints.py:55-66) on the applied torque of the last substep
          CatMax CM;
          cat_max_init(CM);
#pragma unroll
          for (int i = 0; i < 6; i++) {  // unrolled: tau[] must stay in registers (no dynamically indexed locals in this kernel)
            const int j = 6 * side + i;
            const real v = r_abs(tau[i]) - P.effort[j];
            if (valid) S.cat.raw[(size_t)(25 + j) * N + env] = v;
            cat_col_max(CM, 25 + j, v, valid, tid);
          }
          cat_max_publish(CM, S.cat.cmax);
        }
      }
      substep<ROUGH, NM>(P, tid, side, rp, rq, rv, rw, q, qd, tau, mu, mass_add, wl, wr, use_warm, so, S.terrain_h, te);
      use_warm = true;
#ifdef H1V2_WARPCLOCK
      wc_trips += so.trips; wc_ls += so.lstrips;
#endif
      max_it = max(max_it, so.iters); ncap += so.capped; sum_it += so.iters; novf += so.overflow;
      if (S.diag && valid && side == 0) atomicAdd(S.acc + H1V2_LOG_DIM + min(so.iters, 31), 1.f);  // iteration histogram: diagnostics handles only
      s_acc = 0.f;
#pragma unroll
      for (int i = 0; i < 6; i++) s_acc = r_fma(so.qacc[i], so.qacc[i], s_acc);
      if (S.diag && valid) {
        float* dg = S.diag + (size_t)env * H1V2_DIAG_DIM;
#pragma unroll
        for (int i = 0; i < 6; i++) dg[48 + 6 * side + i] = so.qacc[i];
      }
      real nf = r_sqrt(dot(so.F_foot, so.F_foot));
      h_foot[0] = h_foot[1]; h_foot[1] = h_foot[2]; h_foot[2] = nf;
      h_shin[0] = h_shin[1]; h_shin[1] = h_shin[2]; h_shin[2] = r_sqrt(dot(so.F_shin, so.F_shin));
      h_torso[0] = h_torso[1]; h_torso[1] = h_torso[2]; h_torso[2] = r_sqrt(dot(so.F_torso, so.F_torso));
      h_pelvis[0] = h_pelvis[1]; h_pelvis[1] = h_pelvis[2]; h_pelvis[2] = r_sqrt(dot(so.F_pelvis, so.F_pelvis));
      {  // air / contact timers of the lane's foot (ContactSensor)
        const real el = P.h;
        const bool is_c = nf > P.contact_thr;
        const bool first_c = (tm.x > 0.f) && is_c, first_d = (tm.z > 0.f) && !is_c;
        tm.y = first_c ? tm.x + el : tm.y;
        tm.x = is_c ? 0.f : tm.x + el;
        tm.w = first_d ? tm.z + el : tm.w;
        tm.z = is_c ? tm.z + el : 0.f;
      }
    }
  }
#ifdef H1V2_WARPCLOCK
  wc_c1 = clock64() - wc_c0;
#endif
  V3 fv = mk3(0.f, 0.f, 0.f);
  real ankle_z = 0.f;
  if (DO_STEP) {
    const M3 Rn = quat2mat(rq[0], rq[1], rq[2], rq[3]);
    fv = foot_velocity(P.leg[side], tid, Rn, mulv(Rn, mk3(rw[0], rw[1], rw[2])), mk3(rv[0], rv[1], rv[2]), q, qd, P.foot_vel_com != 0, ankle_z);
    __syncwarp();  // every lane is done with its column before the asynchronous copy lands in it
  }
  hist_prefetch(P, S, tid, bid, 0, 0);
  // ---- actuator line and command state ----
  {
    float4 a0 = S.act[lidx], a1 = S.act[N2 + lidx], a2 = S.act[2 * N2 + lidx], a3 = S.act[3 * N2 + lidx], a4 = S.act[4 * N2 + lidx];
    la[0] = a0.x; la[1] = a0.y; la[2] = a0.z; la[3] = a0.w; la[4] = a1.x; la[5] = a1.y;
    T1[0] = a1.z; T1[1] = a1.w; T1[2] = a2.x; T1[3] = a2.y; T1[4] = a2.z; T1[5] = a2.w;
    T2[0] = a3.x; T2[1] = a3.y; T2[2] = a3.z; T2[3] = a3.w; T2[4] = a4.x; T2[5] = a4.y;
    float4 c0 = S.cmd[env], c1 = S.cmd[N + env];
    cmd.c[0] = c0.x; cmd.c[1] = c0.y; cmd.c[2] = c0.z; cmd.heading_target = c0.w;
    cmd.time_left = c1.x; cmd.m_xy = c1.y; cmd.m_yaw = c1.z; cmd.flags = __float_as_int(c1.w);
  }
  if (DO_STEP) {
    // action manager bookkeeping: prev_action <- action ; action <- a ; shift the delay line
    real prev[6];
    {
      const bool line_fresh = (cmd.flags & FLAG_DELAY_FRESH) != 0;
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const int j = 6 * side + k;
        prev[k] = la[k];
        la[k] = __ldg(actions + (size_t)env * 12 + P.inv_perm[j]);
        const real T0 = r_fma(P.action_scale, la[k], P.q0[j]);
        T2[k] = line_fresh ? T0 : T1[k];
        T1[k] = T0;
      }
      cmd.flags &= ~FLAG_DELAY_FRESH;
    }

    // ---- non-finite / runaway guard ----
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 6; k++) bad |= !isfinite(q[k]) || !(r_abs(qd[k]) <= P.runaway_vel) || !(r_abs(la[k]) <= H1V2_ACTION_ABS_MAX);
#pragma unroll
    for (int k = 0; k < 3; k++) bad |= !isfinite(rp[k]) || !(r_abs(rv[k]) <= P.runaway_vel) || !(r_abs(rw[k]) <= P.runaway_vel);
    bad |= !isfinite(rq[0]) || !isfinite(rq[1]) || !isfinite(rq[2]) || !isfinite(rq[3]);
    bad = (__shfl_xor_sync(FULL_MASK, (int)bad, 1) | (int)bad) != 0;

    // ---- counters and terminations (C12/rough_env_cfg.py:95-109) ----
    int64_t ep_len = S.ep_len[env] + 1;
    const bool time_out = ep_len >= P.max_episode_length;
    const real C_foot = r_max(h_foot[0], r_max(h_foot[1], h_foot[2]));
    const real C_shin = r_max(h_shin[0], r_max(h_shin[1], h_shin[2]));
    const real C_torso = r_max(h_torso[0], r_max(h_torso[1], h_torso[2]));
    const real C_pelvis = r_max(h_pelvis[0], r_max(h_pelvis[1], h_pelvis[2]));
    int contact = (((P.m_illegal >> side) & 1u) && C_foot > P.contact_thr) || (((P.m_illegal >> (2 + side)) & 1u) && C_shin > P.contact_thr) ||
                  (((P.m_illegal >> 4) & 1u) && C_torso > P.contact_thr) || (((P.m_illegal >> 5) & 1u) && C_pelvis > P.contact_thr);
    contact = (__shfl_xor_sync(FULL_MASK, contact, 1) | contact) | (int)bad;
    const bool reset = contact || time_out;

    // ---- rewards on the pre-reset state ----
    const RootDerived rd = root_derived(P, rq, rv, rw);
    real r[H1V2_NUM_REW];
#pragma unroll
    for (int t = 0; t < H1V2_NUM_REW; t++) r[t] = 0.f;
    {
      real s_lim = 0.f, s_dev = 0.f, s_lim_b = 0.f, s_dev_b = 0.f, s_vel = 0.f, s_da = 0.f;
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const int j = 6 * side + k;
        const real lim = -r_min(q[k] - P.soft_lo[j], 0.f) + r_max(q[k] - P.soft_hi[j], 0.f), dev = r_abs(q[k] - P.q0[j]);
        if ((P.m_poslim >> j) & 1u) s_lim += lim;
        if ((P.m_dev >> j) & 1u) s_dev += dev;
        if ((P.m_poslim_b >> j) & 1u) s_lim_b += lim;
        if ((P.m_dev_b >> j) & 1u) s_dev_b += dev;
        s_vel = r_fma(qd[k], qd[k], s_vel);
        real da = la[k] - prev[k];
        s_da = r_fma(da, da, s_da);
      }
      r[H1V2_REW_DOF_POS_LIMITS] = pair_sum(s_lim);
      r[H1V2_REW_JOINT_DEV_HIP] = pair_sum(s_dev);
      r[H1V2_REW_DOF_POS_LIMITS_B] = pair_sum(s_lim_b);
      r[H1V2_REW_JOINT_DEV_B] = pair_sum(s_dev_b);
      r[H1V2_REW_TORQUES] = pair_sum(s_tau);
      r[H1V2_REW_DOF_ACC] = pair_sum(s_acc);
      r[H1V2_REW_JOINT_VEL] = pair_sum(s_vel);
      r[H1V2_REW_ACTION_RATE] = pair_sum(s_da);
      r[H1V2_REW_TERMINATION] = contact ? 1.f : 0.f;
      const real hn = r_rsqrt(r_max(rd.hx * rd.hx + rd.hy * rd.hy, 1e-30f));  // cos / sin of the heading without the angle itself
      const real ch = rd.hx * hn, sh = rd.hy * hn;
      real ex = cmd.c[0] - (ch * rd.vw.x + sh * rd.vw.y), ey = cmd.c[1] - (-sh * rd.vw.x + ch * rd.vw.y);
      r[H1V2_REW_TRACK_LIN_XY_YAW] = r_exp(-(ex * ex + ey * ey) * P.inv_std2);
      real ez = cmd.c[2] - rd.ww.z;
      r[H1V2_REW_TRACK_ANG_Z_WORLD] = r_exp(-(ez * ez) * P.inv_std2);
      ex = cmd.c[0] - rd.vb.x; ey = cmd.c[1] - rd.vb.y;
      r[H1V2_REW_TRACK_LIN_XY_BASE] = r_exp(-(ex * ex + ey * ey) * P.inv_std2);
      ez = cmd.c[2] - rd.wb.z;
      r[H1V2_REW_TRACK_ANG_Z_BASE] = r_exp(-(ez * ez) * P.inv_std2);
      const bool moving = r_sqrt(cmd.c[0] * cmd.c[0] + cmd.c[1] * cmd.c[1]) > 0.1f;
      // feet_air_time_positive_biped (V/mdp/rewards.py:38-62) and feet_air_time (:13-35)
      const int inc = tm.z > 0.f;
      const real mode_t = inc ? tm.z : tm.x;
      const int inc_p = __shfl_xor_sync(FULL_MASK, inc, 1);
      const real mode_p = __shfl_xor_sync(FULL_MASK, mode_t, 1);
      const bool single = (inc + inc_p) == 1;
      real mn = r_min(single ? mode_t : 0.f, single ? mode_p : 0.f);
      mn = r_min(mn, P.air_thr);
      r[H1V2_REW_FEET_AIR_BIPED] = moving ? mn : 0.f;
      const bool first = (tm.z > 0.f) && (tm.z < P.step_dt + 1e-8f);
      real l2 = first ? (tm.y - P.air_thr) : 0.f;
      l2 = pair_sum(l2);
      r[H1V2_REW_FEET_AIR_L2] = moving ? l2 : 0.f;
      real slide = (C_foot > P.contact_thr) ? r_sqrt(fv.x * fv.x + fv.y * fv.y) : 0.f;
      r[H1V2_REW_FEET_SLIDE] = pair_sum(slide);
      r[H1V2_REW_ANG_VEL_XY] = rd.wb.x * rd.wb.x + rd.wb.y * rd.wb.y;
      r[H1V2_REW_FLAT_ORI] = rd.g.x * rd.g.x + rd.g.y * rd.g.y;
      r[H1V2_REW_LIN_VEL_Z] = rd.vb.z * rd.vb.z;
      r[H1V2_REW_BASE_HEIGHT] = (rp[2] - P.base_h) * (rp[2] - P.base_h);
      real und = 0.f, cf = 0.f;
      // undesired_contacts counts bodies above the sensor threshold; contact_forces (C12/rsl_env_cfg.py:395-404) sums the
      // excess of max_h |F| over its own threshold on its own bodies
      if ((P.m_undesired >> side) & 1u) und += C_foot > P.contact_thr;
      if ((P.m_undesired >> (2 + side)) & 1u) und += C_shin > P.contact_thr;
      if ((P.m_cforce >> side) & 1u) cf += r_max(C_foot - P.cforce_thr, 0.f);
      if ((P.m_cforce >> (2 + side)) & 1u) cf += r_max(C_shin - P.cforce_thr, 0.f);
      if (side == 0) {
        if ((P.m_undesired >> 4) & 1u) und += C_torso > P.contact_thr;
        if ((P.m_undesired >> 5) & 1u) und += C_pelvis > P.contact_thr;
        if ((P.m_cforce >> 4) & 1u) cf += r_max(C_torso - P.cforce_thr, 0.f);
        if ((P.m_cforce >> 5) & 1u) cf += r_max(C_pelvis - P.cforce_thr, 0.f);
      }
      r[H1V2_REW_UNDESIRED_CONTACTS] = pair_sum(und);
      r[H1V2_REW_CONTACT_FORCES] = pair_sum(cf);
    }
    real total = 0.f;
#pragma unroll
    for (int t = 0; t < H1V2_NUM_REW; t++) {
      real v = (P.w[t] == 0.f || bad) ? 0.f : P.w[t] * r[t] * P.step_dt;
      r[t] = v;
      total += v;
    }
    // episode sums: lane 0 owns float4 0..2 (terms 0..11), lane 1 owns float4 3..5 (terms 12..21 and two pads)
    real es[12];
    const int ef0 = side == 0 ? 0 : 3, enf4 = 3;
    {
      real rsel[12];
#pragma unroll
      for (int i = 0; i < 12; i++) rsel[i] = side == 0 ? r[i] : (12 + i < H1V2_NUM_REW ? r[12 + i < H1V2_NUM_REW ? 12 + i : 0] : 0.f);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < enf4) e4 = S.epsum[(size_t)(ef0 + i) * N + env];
        es[4 * i + 0] = e4.x + rsel[4 * i + 0]; es[4 * i + 1] = e4.y + rsel[4 * i + 1];
        es[4 * i + 2] = e4.z + rsel[4 * i + 2]; es[4 * i + 3] = e4.w + rsel[4 * i + 3];
      }
    }
    if (valid && side == 0) {
      rew[env] = total;
      term[env] = (uint8_t)(contact != 0);
      trunc[env] = (uint8_t)time_out;
    }
    if (CAT) {
      // ---- Constraints-as-Terminations: the raw constraint columns of this step on the pre-reset state, coalesced [56][N], and their
      //      maxima (T/utils/cat/constraints.py:22-308, parameters C12/cat_env_cfg.py:336-431).  The apply kernel turns them into
      //      probabilities once the maxima of ALL envs are known; the 12 no_move columns need another env's joint velocities and are
      //      gathered there -- their maxima are known here: the gather tiles the dead-zone members over the batch, so every member
      //      appears (constraints.py:209-231). ----
      const KCat& T = S.cat;
      CatMax CM;
      cat_max_init(CM);
      const real dzn = T.no_move_deadzone;
      const bool in_dz = r_abs(cmd.c[0]) < dzn && r_abs(cmd.c[1]) < dzn && r_abs(cmd.c[2]) < dzn;
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const int j = 6 * side + k;
        const real vp = r_max(P.soft_lo[j] - q[k], q[k] - P.soft_hi[j]);   // joint_position_limits :22-31
        const real vv = r_abs(qd[k]) - T.vel_limit;                         // joint_velocity_limits :40-52
        if (valid) { T.raw[(size_t)(1 + j) * N + env] = vp; T.raw[(size_t)(13 + j) * N + env] = vv; T.qd[(size_t)j * N + env] = qd[k]; }
        cat_col_max(CM, 1 + j, vp, valid, tid);
        cat_col_max(CM, 13 + j, vv, valid, tid);
        cat_col_max(CM, 39 + j, r_abs(qd[k]) - T.no_move_vel_limit, valid && in_dz, tid);  // no_move :209-231
      }
      const real vf = C_foot - T.foot_force_limit;                          // foot_contact_force :161-168
      if (valid) T.raw[(size_t)(37 + side) * N + env] = vf;
      cat_col_max(CM, 37 + side, vf, valid, tid);
      {                                                                      // foot_clearance :268-308
        const real dzc = T.clearance_deadzone;
        const real active = (r_abs(cmd.c[0]) > dzc || r_abs(cmd.c[1]) > dzc || r_abs(cmd.c[2]) > dzc) ? 1.f : 0.f;
        const bool touchdown = tm.z > 0.f && tm.z < P.step_dt + 1.0e-8f;
        const real sw = valid ? T.swing[(size_t)side * N + env] : 0.f;
        const real vc = (T.clearance_min_height - sw) * (touchdown ? 1.f : 0.f) * active;
        if (valid) { T.raw[(size_t)(54 + side) * N + env] = vc; T.swing[(size_t)side * N + env] = touchdown ? 0.f : r_max(sw, rp[2] + ankle_z); }
        cat_col_max(CM, 54 + side, vc, valid, tid);
      }
      int anyc = (((T.contact_slots >> side) & 1u) && C_foot > 1.0f) || (((T.contact_slots >> (2 + side)) & 1u) && C_shin > 1.0f) ||
                 (((T.contact_slots >> 4) & 1u) && C_torso > 1.0f) || (((T.contact_slots >> 5) & 1u) && C_pelvis > 1.0f);
      anyc |= __shfl_xor_sync(FULL_MASK, anyc, 1);
      const int nfeet = (C_foot > 1.0f) + __shfl_xor_sync(FULL_MASK, (int)(C_foot > 1.0f), 1);
      const real v0 = anyc ? 1.f : 0.f;                                                         // contact :86-99
      const real v51 = r_sqrt(rd.g.x * rd.g.x + rd.g.y * rd.g.y) - T.orientation_limit;        // base_orientation :234-246
      const real v52 = (rp[2] < T.height - T.height_std || rp[2] > T.height + T.height_std) ? 1.f : 0.f;  // base_height :249-266
      const real v53 = (nfeet < 1 || nfeet > 2) ? 1.f : 0.f;                                    // foot_contact :171-193
      if (valid && side == 0) {
        T.raw[env] = v0; T.raw[(size_t)51 * N + env] = v51; T.raw[(size_t)52 * N + env] = v52; T.raw[(size_t)53 * N + env] = v53;
        T.aux[env] = (float)ep_len; T.aux[(size_t)N + env] = reset ? 1.f : 0.f;
      }
      {  // dead-zone members of this warp as a bit mask (bit p = env warp_env0 + p): the even lanes' ballot bits, packed
        unsigned m = __ballot_sync(FULL_MASK, in_dz && valid && side == 0) & 0x55555555u;
        m = (m | (m >> 1)) & 0x33333333u; m = (m | (m >> 2)) & 0x0f0f0f0fu; m = (m | (m >> 4)) & 0x00ff00ffu; m = (m | (m >> 8)) & 0x0000ffffu;
        if ((tid & 31) == 0) T.dz[bid] = (unsigned short)m;
      }
      const bool v0side = valid && side == 0;
      cat_col_max(CM, 0, v0, v0side, tid); cat_col_max(CM, 51, v51, v0side, tid);
      cat_col_max(CM, 52, v52, v0side, tid); cat_col_max(CM, 53, v53, v0side, tid);
      cat_max_publish(CM, T.cmax);
    }
    const int novf_pair = novf + __shfl_xor_sync(FULL_MASK, novf, 1);
    if (S.diag && valid) {
      float* dg = S.diag + (size_t)env * H1V2_DIAG_DIM;
      dg[3 * side + 0] = so.F_foot.x; dg[3 * side + 1] = so.F_foot.y; dg[3 * side + 2] = so.F_foot.z;
      dg[6 + 3 * side + 0] = so.F_shin.x; dg[6 + 3 * side + 1] = so.F_shin.y; dg[6 + 3 * side + 2] = so.F_shin.z;
#pragma unroll
      for (int i = 0; i < 3; i++) { dg[18 + 3 * side + i] = h_foot[i]; dg[24 + 3 * side + i] = h_shin[i]; }
      dg[80 + 3 * side + 0] = fv.x; dg[80 + 3 * side + 1] = fv.y; dg[80 + 3 * side + 2] = fv.z;
#pragma unroll
      for (int k = 0; k < 6; k++) { dg[96 + 7 + 6 * side + k] = q[k]; dg[115 + 6 + 6 * side + k] = qd[k]; }
      dg[133 + 4 * side + 0] = tm.x; dg[133 + 4 * side + 1] = tm.y; dg[133 + 4 * side + 2] = tm.z; dg[133 + 4 * side + 3] = tm.w;
      if (side == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { dg[96 + k] = rp[k]; dg[115 + k] = rv[k]; dg[118 + k] = rw[k]; }
#pragma unroll
        for (int k = 0; k < 4; k++) dg[99 + k] = rq[k];
        dg[12] = so.F_torso.x; dg[13] = so.F_torso.y; dg[14] = so.F_torso.z;
        dg[15] = so.F_pelvis.x; dg[16] = so.F_pelvis.y; dg[17] = so.F_pelvis.z;
#pragma unroll
        for (int i = 0; i < 3; i++) { dg[30 + i] = h_torso[i]; dg[33 + i] = h_pelvis[i]; }
#pragma unroll
        for (int t = 0; t < H1V2_NUM_REW; t++) dg[H1V2_DIAG_REW0 + t] = r[t];
        dg[86] = (real)max_it; dg[87] = (real)ncap; dg[88] = (real)sum_it;
        dg[89] = (real)novf_pair;  // contact points dropped by either leg's full list
        // for the Constraints-as-Terminations tail: the command the mdp terms of this step read (before its update), the episode length
        dg[141] = cmd.c[0]; dg[142] = cmd.c[1]; dg[143] = cmd.c[2];
        dg[166] = (real)ep_len; dg[167] = reset ? 1.f : 0.f;
      }
    }
    // solver statistics
    {
      int wm = __reduce_max_sync(0xffffffffu, valid ? max_it : 0);
      int wc = __reduce_add_sync(0xffffffffu, (valid && side == 0) ? ncap : 0);
      int ws = __reduce_add_sync(0xffffffffu, (valid && side == 0) ? sum_it : 0);
      int wo = __reduce_add_sync(0xffffffffu, valid ? novf : 0);  // per leg: each lane owns its own contact list
      if ((tid & 31) == 0) {
        atomicMax((int*)(S.acc + H1V2_LOG_MAX_ITERS), wm);
        if (wc) atomicAdd(S.acc + H1V2_LOG_CAP_HITS, (real)wc);
        atomicAdd(S.acc + H1V2_LOG_SUM_ITERS, (real)ws);
        if (wo) atomicAdd(S.acc + H1V2_LOG_CONTACT_OVERFLOW, (real)wo);
      }
    }
    // ---- reset (T/utils/cat/cat_env.py:195-248): log, then new state ----
    if (reset) {
      if (valid) {
        const real inv = 1.f / P.max_episode_length_s;
        const int f0 = side == 0 ? 0 : 12, cnt = side == 0 ? 12 : H1V2_NUM_REW - 12;
#pragma unroll
        for (int i = 0; i < 12; i++)
          if (i < cnt) atomicAdd(S.acc + H1V2_LOG_REW0 + f0 + i, es[i] * inv);
        if (side == 0) {
          atomicAdd(S.acc + H1V2_LOG_COUNT, 1.f);
          if (time_out) atomicAdd(S.acc + H1V2_LOG_TERM_TIMEOUT, 1.f);
          if (contact) atomicAdd(S.acc + H1V2_LOG_TERM_CONTACT, 1.f);
          atomicAdd(S.acc + H1V2_LOG_ERR_XY, cmd.m_xy);
          atomicAdd(S.acc + H1V2_LOG_ERR_YAW, cmd.m_yaw);
          if (bad) atomicAdd(S.acc + H1V2_LOG_NAN_RESETS, 1.f);
        }
      }
      // CurriculumManager.compute(env_ids) comes first in _reset_idx (cat_env.py:197-200): the terrain level moves on the state and the
      // command the episode ended with
      if (ROUGH) cmd.flags = terrain_curriculum(P, cmd.flags, (float)rp[0], (float)rp[1], cmd.c[0], cmd.c[1], gid, step);
      reset_env(P, side, gid, step, rp, rq, rv, rw, q, qd, la, T1, T2, tm, cmd, push_left);
#pragma unroll
      for (int i = 0; i < 12; i++) es[i] = 0.f;
      ep_len = 0;
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (i < enf4) S.epsum[(size_t)(ef0 + i) * N + env] = make_float4(es[4 * i], es[4 * i + 1], es[4 * i + 2], es[4 * i + 3]);
      if (side == 0) S.ep_len[env] = ep_len;
    }
  }

  // ---- command manager, interval events, observation (on the post-reset state) ----
  RootDerived rd = root_derived(P, rq, rv, rw);
  if (DO_STEP) {
    const bool in_dz = update_command(P, cmd, rd, gid, step, (unsigned)S.counters[3]);
    if (P.cmd_class == 1) {  // envs inside the dead zone at the end of this step, for the next step's balancing
      const unsigned m = __ballot_sync(FULL_MASK, in_dz && valid && side == 0);
      if ((tid & 31) == 0 && m) atomicAdd(S.counters + 2, (unsigned long long)__popc(m));
    }
    if (P.push_enable) {  // push_by_setting_velocity (V/velocity_env_cfg.py:212-217)
      push_left -= P.step_dt;
      if (push_left < 1e-6f) {
        float u[4];
        rng4(P.key0, gid, step, STREAM_EVENT, 2, u);
        push_left = uni(u[2], P.push_int[0], P.push_int[1]);
        rv[0] += uni(u[0], P.push_v[0], P.push_v[1]);
        rv[1] += uni(u[1], P.push_v[0], P.push_v[1]);
        rd = root_derived(P, rq, rv, rw);
      }
    }
  }
  emit_observation(P, S, tid, bid, env, side, valid, gid, step, head, rd, cmd, q, qd, la, obs);
  if (ROUGH) {
    if (P.scan_nx > 0) emit_height_scan<true>(P, S, tid, bid, step, rp, rd, cmd.flags, obs);
    if (DO_STEP) {  // terrain level census of the step (after the resets)
      const int lv = __reduce_add_sync(FULL_MASK, (valid && side == 0) ? ((cmd.flags >> FLAG_LEVEL_SHIFT) & 255) : 0);
      if ((tid & 31) == 0) atomicAdd(S.tlog, (float)lv);
    }
  }
  cmd.flags &= ~FLAG_HIST_FRESH;
  if (S.host_flags) {
    // host path (h1v2_step_host): everything this warp owes the caller -- reward, flags, sample or row -- has been stored into mapped
    // host memory; make it visible system-wide, then raise the warp's flag: the host threads pick the envs up warp by warp while the
    // rest of the grid is still running, instead of waiting for the whole launch
    __threadfence_system();
    __syncwarp();
    if ((tid & 31) == 0) *((volatile unsigned*)S.host_flags + bid) = S.host_seq;
  }

  // ---- store state ----
  if (valid) {
    if (side == 0) {
      S.root[env] = make_float4(rp[0], rp[1], rp[2], rq[0]);
      S.root[N + env] = make_float4(rq[1], rq[2], rq[3], rv[0]);
      S.root[2 * N + env] = make_float4(rv[1], rv[2], rw[0], rw[1]);
      S.root[3 * N + env] = make_float4(rw[2], mu, mass_add, push_left);
      S.cmd[env] = make_float4(cmd.c[0], cmd.c[1], cmd.c[2], cmd.heading_target);
      S.cmd[N + env] = make_float4(cmd.time_left, cmd.m_xy, cmd.m_yaw, __int_as_float(cmd.flags));
    }
    S.leg[lidx] = make_float4(q[0], q[1], q[2], q[3]);
    S.leg[N2 + lidx] = make_float4(q[4], q[5], qd[0], qd[1]);
    S.leg[2 * N2 + lidx] = make_float4(qd[2], qd[3], qd[4], qd[5]);
    S.act[lidx] = make_float4(la[0], la[1], la[2], la[3]);
    S.act[N2 + lidx] = make_float4(la[4], la[5], T1[0], T1[1]);
    S.act[2 * N2 + lidx] = make_float4(T1[2], T1[3], T1[4], T1[5]);
    S.act[3 * N2 + lidx] = make_float4(T2[0], T2[1], T2[2], T2[3]);
    S.act[4 * N2 + lidx] = make_float4(T2[4], T2[5], 0.f, 0.f);
    S.timers[lidx] = tm;
    S.warm[lidx] = make_float4(wl[0], wl[1], wl[2], wl[3]);
    S.warm[N2 + lidx] = make_float4(wl[4], wl[5], wr[0], wr[1]);
    S.warm[2 * N2 + lidx] = make_float4(wr[2], wr[3], wr[4], wr[5]);
  }
#ifdef H1V2_WARPCLOCK
  if (DO_STEP && tid == 0 && bid < 16384) {
    g_warpclock[4 * bid] = wc_g0; g_warpclock[4 * bid + 1] = wc_c1; g_warpclock[4 * bid + 2] = clock64() - wc_c0;
    g_warpclock[4 * bid + 3] = (unsigned long long)wc_trips | ((unsigned long long)wc_ls << 32);
  }
#endif
  // ---- the last block to get here finalizes the step ----
  __threadfence();
  unsigned ticket = 0;
  if (tid == 0) ticket = atomicAdd(S.done, 1u);
  ticket = __shfl_sync(FULL_MASK, ticket, 0);
  if (ticket == gridDim.x - 1) {
    __threadfence();
    finalize_step(S, DO_STEP, CAT, tid, P.n, P.epw);
    if (S.host_flags) {  // host path: the launch's bookkeeping (log vector, counters) is done too -- the word after the warps' flags
      __threadfence_system();
      __syncwarp();
      if ((tid & 31) == 0) *((volatile unsigned*)S.host_flags + gridDim.x) = S.host_seq;
    }
  }
}

}  // namespace h1v2
