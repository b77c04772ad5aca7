// h1v2_physics.cuh -- one physics substep of the H1-2 (floating base + 2 x 6-dof legs) for ONE LANE PAIR.
//
// Work decomposition: two adjacent lanes own one environment; lane `side` owns one leg.
// The mass matrix of a floating base with two serial branches is block-arrow:
//        [ A    B_L   B_R ]      A   6x6 root block (replicated bit-identically on both lanes)
//    M = [ B_L' C_L   0   ]      B_s 6x6 root-leg coupling of this lane's leg
//        [ B_R' 0     C_R ]      C_s 6x6 leg block
// so every factorisation is: Cholesky(C_s) per lane, Y_s = L_s^-1 B_s', one pair-exchange of Y_s'Y_s (21 floats),
// Cholesky of the 6x6 Schur complement on both lanes.  Contacts attach a symmetric 6x6 "stiffness" K to the foot /
// shin / root bodies, so the Newton Hessian M + J'DJ keeps exactly the same block-arrow shape.
//
// Semantics restated from the reference's physics target (MuJoCo 3.3.6 as configured by
// packages/biped_deploy/biped_deploy/simulator/sim_mujoco.py:39-41 on
// packages/biped_assets/biped_assets/models/h12/scene/h12_12dof.xml):
// CRBA + RNE bias, soft constraints (friction loss, joint limits, pyramidal contacts with solref/solimp),
// primal Newton with exact line search, implicitfast velocity update, semi-implicit Euler.
#pragma once
#include "h1v2_math.cuh"
#include "h1v2_params.h"

namespace h1v2 {

__device__ __forceinline__ float pair_sum(float v, unsigned pm) { return v + __shfl_xor_sync(pm, v, 1); }
__device__ __forceinline__ V3 pair_sum(V3 v, unsigned pm) { return mk3(pair_sum(v.x, pm), pair_sum(v.y, pm), pair_sum(v.z, pm)); }

struct Blk {  // one lane's share of a block-arrow symmetric matrix
  float C[21];
  float B[36];  // B[k*6+j]: root dof k, leg joint j
  float A[21];
};
struct Fac {
  float L[21], invd[6];
  float Y[36];  // Y[i*6+k] = (L^-1 B')[i][k]
  float La[21], inva[6];
};

__device__ __forceinline__ void factor(const Blk& Mx, Fac& F, unsigned pm) {
#pragma unroll
  for (int i = 0; i < 21; i++) F.L[i] = Mx.C[i];
  chol6(F.L, F.invd);
#pragma unroll
  for (int k = 0; k < 6; k++) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      float s = Mx.B[k * 6 + i];
#pragma unroll
      for (int m = 0; m < i; m++) s = fmaf(-F.L[TI(i, m)], F.Y[m * 6 + k], s);
      F.Y[i * 6 + k] = s * F.invd[i];
    }
  }
#pragma unroll
  for (int a = 0; a < 6; a++) {
#pragma unroll
    for (int b = 0; b <= a; b++) {
      float g = 0.f;
#pragma unroll
      for (int i = 0; i < 6; i++) g = fmaf(F.Y[i * 6 + a], F.Y[i * 6 + b], g);
      F.La[TI(a, b)] = Mx.A[TI(a, b)] - pair_sum(g, pm);
    }
  }
  chol6(F.La, F.inva);
}

// solve M x = b ; b_leg/x_leg own 6, b_root/x_root replicated
__device__ __forceinline__ void solve(const Fac& F, float (&leg)[6], float (&root)[6], unsigned pm) {
  fwd6(F.L, F.invd, leg);
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 6; i++) t = fmaf(F.Y[i * 6 + k], leg[i], t);
    root[k] -= pair_sum(t, pm);
  }
  fwd6(F.La, F.inva, root);
  bwd6(F.La, F.inva, root);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float s = leg[i];
#pragma unroll
    for (int k = 0; k < 6; k++) s = fmaf(-F.Y[i * 6 + k], root[k], s);
    leg[i] = s;
  }
  bwd6(F.L, F.invd, leg);
}

// y = M x (leg part own, root part replicated)
__device__ __forceinline__ void matvec(const Blk& Mx, const float (&xl)[6], const float (&xr)[6], float (&yl)[6], float (&yr)[6],
                                       unsigned pm) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 6; j++) s = fmaf(i >= j ? Mx.C[TI(i, j)] : Mx.C[TI(j, i)], xl[j], s);
#pragma unroll
    for (int k = 0; k < 6; k++) s = fmaf(Mx.B[k * 6 + i], xr[k], s);
    yl[i] = s;
  }
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 6; j++) t = fmaf(Mx.B[k * 6 + j], xl[j], t);
    float s = pair_sum(t, pm);
#pragma unroll
    for (int m = 0; m < 6; m++) s = fmaf(k >= m ? Mx.A[TI(k, m)] : Mx.A[TI(m, k)], xr[m], s);
    yr[k] = s;
  }
}

// MuJoCo getimpedance (solimp = d0, dmax, width, midpoint, power), margin 0
__device__ __forceinline__ float impedance(const float* si, float pos) {
  float dmin = fminf(fmaxf(si[0], 1e-4f), 0.9999f), dmax = fminf(fmaxf(si[1], 1e-4f), 0.9999f);
  float width = si[2], mid = fminf(fmaxf(si[3], 1e-4f), 0.9999f), power = fmaxf(si[4], 1.f);
  if (dmin == dmax || width <= 1e-15f) return 0.5f * (dmin + dmax);
  float x = fabsf(pos) / width;
  if (x >= 1.f) return dmax;
  if (x == 0.f) return dmin;
  float y;
  if (power == 1.f) y = x;
  else if (power == 2.f) y = (x <= mid) ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);
  else y = (x <= mid) ? powf(x, power) / powf(mid, power - 1.f) : 1.f - powf(1.f - x, power) / powf(1.f - mid, power - 1.f);
  return dmin + y * (dmax - dmin);
}

// general symmetric 6x6 about O in [ang;lin] order: aa (6: xx yy zz xy xz yz), al (9: row=ang, col=lin), ll (6)
struct K6 {
  float aa[6], al[9], ll[6];
};
__device__ __forceinline__ void k6_zero(K6& K) {
#pragma unroll
  for (int i = 0; i < 6; i++) K.aa[i] = K.ll[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 9; i++) K.al[i] = 0.f;
}
__device__ __forceinline__ void k6_add(K6& K, const K6& o) {
#pragma unroll
  for (int i = 0; i < 6; i++) { K.aa[i] += o.aa[i]; K.ll[i] += o.ll[i]; }
#pragma unroll
  for (int i = 0; i < 9; i++) K.al[i] += o.al[i];
}
__device__ __forceinline__ V3 sym3_mul(const float* s, V3 v) {
  return mk3(fmaf(s[0], v.x, fmaf(s[3], v.y, s[4] * v.z)), fmaf(s[3], v.x, fmaf(s[1], v.y, s[5] * v.z)),
             fmaf(s[4], v.x, fmaf(s[5], v.y, s[2] * v.z)));
}
// (n,l) = K (w,u)
__device__ __forceinline__ void k6_apply(const K6& K, V3 w, V3 u, V3& n, V3& l) {
  n = sym3_mul(K.aa, w) + mk3(fmaf(K.al[0], u.x, fmaf(K.al[1], u.y, K.al[2] * u.z)), fmaf(K.al[3], u.x, fmaf(K.al[4], u.y, K.al[5] * u.z)),
                              fmaf(K.al[6], u.x, fmaf(K.al[7], u.y, K.al[8] * u.z)));
  l = sym3_mul(K.ll, u) + mk3(fmaf(K.al[0], w.x, fmaf(K.al[3], w.y, K.al[6] * w.z)), fmaf(K.al[1], w.x, fmaf(K.al[4], w.y, K.al[7] * w.z)),
                              fmaf(K.al[2], w.x, fmaf(K.al[5], w.y, K.al[8] * w.z)));
}

#define NPT 11  // candidate contact points of a lane: 0..3 sole corners, 4..5 shin ends, 6..9 torso corners, 10 pelvis (lane 0)

struct Contacts {
  unsigned mask;     // active candidates (dist < 0)
  float r[NPT][3];   // contact point relative to O
  float ub[NPT][3];  // B * point velocity
  float kap[NPT];    // K * imp * dist
  float D[NPT];      // 1/R of the four pyramid edges
  float e[NPT][3];   // J_p x + ub at the current iterate
  float us[NPT][3];  // J_p search
};

// one contact point at the iterate: edge residuals -> force vector and (optionally) the 3x3 weight W = D * sum_active a a'
// edges a_k = (0,mu,1) (0,-mu,1) (-mu,0,1) (mu,0,1)
__device__ __forceinline__ void point_eval(V3 e, float kap, float D, float mu, V3& F, float (&W)[5]) {
  float j0 = fmaf(mu, e.y, e.z) + kap, j1 = fmaf(-mu, e.y, e.z) + kap;
  float j2 = fmaf(-mu, e.x, e.z) + kap, j3 = fmaf(mu, e.x, e.z) + kap;
  float a0 = j0 < 0.f ? D : 0.f, a1 = j1 < 0.f ? D : 0.f, a2 = j2 < 0.f ? D : 0.f, a3 = j3 < 0.f ? D : 0.f;
  float f0 = -a0 * j0, f1 = -a1 * j1, f2 = -a2 * j2, f3 = -a3 * j3;
  F = mk3(mu * (f3 - f2), mu * (f0 - f1), (f0 + f1) + (f2 + f3));
  W[0] = mu * mu * (a2 + a3);         // xx
  W[1] = mu * mu * (a0 + a1);         // yy
  W[2] = (a0 + a1) + (a2 + a3);       // zz
  W[3] = mu * (a3 - a2);              // xz
  W[4] = mu * (a0 - a1);              // yz
}
// K += X' W X with X = [G | 1], G = -[r]x   (point velocity u = v_O + w x r)
__device__ __forceinline__ void k6_add_point(K6& K, V3 r, const float (&W)[5]) {
  // columns of G: g_j = e_j x r
  V3 g0 = mk3(0.f, -r.z, r.y), g1 = mk3(r.z, 0.f, -r.x), g2 = mk3(-r.y, r.x, 0.f);
  // W g_j  (W symmetric with xy = 0)
  V3 w0 = mk3(W[0] * g0.x + W[3] * g0.z, W[1] * g0.y + W[4] * g0.z, W[3] * g0.x + W[4] * g0.y + W[2] * g0.z);
  V3 w1 = mk3(W[0] * g1.x + W[3] * g1.z, W[1] * g1.y + W[4] * g1.z, W[3] * g1.x + W[4] * g1.y + W[2] * g1.z);
  V3 w2 = mk3(W[0] * g2.x + W[3] * g2.z, W[1] * g2.y + W[4] * g2.z, W[3] * g2.x + W[4] * g2.y + W[2] * g2.z);
  K.aa[0] += dot(g0, w0); K.aa[1] += dot(g1, w1); K.aa[2] += dot(g2, w2);
  K.aa[3] += dot(g0, w1); K.aa[4] += dot(g0, w2); K.aa[5] += dot(g1, w2);
  // al[i][c] = (G' W)[i][c] = (W g_i)[c]
  K.al[0] += w0.x; K.al[1] += w0.y; K.al[2] += w0.z;
  K.al[3] += w1.x; K.al[4] += w1.y; K.al[5] += w1.z;
  K.al[6] += w2.x; K.al[7] += w2.y; K.al[8] += w2.z;
  K.ll[0] += W[0]; K.ll[1] += W[1]; K.ll[2] += W[2]; K.ll[4] += W[3]; K.ll[5] += W[4];
}

// line-search contribution of one contact point at step alpha: d1 += D*jar*jv, d2 += D*jv^2 over active edges
__device__ __forceinline__ void point_ls(V3 e, V3 us, float a, float kap, float D, float mu, float& d1, float& d2) {
  V3 ea = fma3(us, a, e);
  float j0 = fmaf(mu, ea.y, ea.z) + kap, j1 = fmaf(-mu, ea.y, ea.z) + kap;
  float j2 = fmaf(-mu, ea.x, ea.z) + kap, j3 = fmaf(mu, ea.x, ea.z) + kap;
  float v0 = fmaf(mu, us.y, us.z), v1 = fmaf(-mu, us.y, us.z), v2 = fmaf(-mu, us.x, us.z), v3 = fmaf(mu, us.x, us.z);
  float s1 = 0.f, s2 = 0.f;
  if (j0 < 0.f) { s1 = fmaf(j0, v0, s1); s2 = fmaf(v0, v0, s2); }
  if (j1 < 0.f) { s1 = fmaf(j1, v1, s1); s2 = fmaf(v1, v1, s2); }
  if (j2 < 0.f) { s1 = fmaf(j2, v2, s1); s2 = fmaf(v2, v2, s2); }
  if (j3 < 0.f) { s1 = fmaf(j3, v3, s1); s2 = fmaf(v3, v3, s2); }
  d1 = fmaf(D, s1, d1);
  d2 = fmaf(D, s2, d2);
}

// friction-loss row: force and activity at residual jar
__device__ __forceinline__ float floss_force(float jar, float D, float lim, float fl, float& act) {
  if (jar <= -lim) { act = 0.f; return fl; }
  if (jar >= lim) { act = 0.f; return -fl; }
  act = D;
  return -D * jar;
}

struct SubOut {
  V3 F_foot, F_shin, F_torso, F_pelvis;  // net contact forces (torso pair-summed, pelvis valid on both lanes)
  float qacc[6];                         // joint accelerations of this leg after the implicit update
  int iters, capped;
};

// ----------------------------------------------------------------------------------------------------------
// One physics substep.  State is updated in place.  tau = joint torques after the actuator's effort clip.
// ----------------------------------------------------------------------------------------------------------
__device__ __noinline__ void substep(const KParams& P, const int side, const unsigned pm, float (&rp)[3], float (&rq)[4],
                                     float (&rv)[3], float (&rw)[3], float (&q)[6], float (&qd)[6], const float (&tau)[6],
                                     const float mu, const float mass_add, float (&wl)[6], float (&wr)[6], const bool use_warm,
                                     SubOut& out) {
  const KLeg& LG = P.leg[side];
  const float h = P.h;
  // ---- root frame ----
  {
    float n = rsqrtf(rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3]);
    rq[0] *= n; rq[1] *= n; rq[2] *= n; rq[3] *= n;
  }
  const M3 R0 = quat2mat(rq[0], rq[1], rq[2], rq[3]);
  const V3 v0 = mk3(rv[0], rv[1], rv[2]);
  const V3 om0 = mulv(R0, mk3(rw[0], rw[1], rw[2]));
  const V3 acc0 = mk3(0.f, 0.f, P.gravity) + cross(v0, om0);  // cacc of the root: -gravity + free-joint cdof_dot
  const float m0 = P.root_mass + mass_add;
  const RI I0 = body_inertia(R0, mulv(R0, ld3(P.root_ipos)), m0, P.root_inertia, m0 / P.root_mass);
  V3 f0n, f0l;
  {
    V3 an, al, vn, vl;
    ri_apply(I0, mk3(0.f, 0.f, 0.f), acc0, an, al);
    ri_apply(I0, om0, v0, vn, vl);
    f0n = an + cross(om0, vn) + cross(v0, vl);
    f0l = al + cross(om0, vl);
  }
  // ---- leg forward pass: kinematics, velocities, bias accelerations, per-body inertia and force ----
  V3 w[6], u[6];
  RI Ib[6];
  V3 fn[6], fl[6];
  M3 Rshin, Rfoot;
  V3 xshin, xfoot, om_shin, vo_shin, om_foot, vo_foot;
  {
    M3 R = R0;
    V3 x = mk3(0.f, 0.f, 0.f), om = om0, vo = v0, al = mk3(0.f, 0.f, 0.f), ao = acc0;
#define LEG_JOINT(i, AX)                                                              \
  {                                                                                   \
    x = x + mulv(R, ld3(LG.pos[i]));                                                  \
    V3 wi = axis_col<AX>(R);                                                          \
    V3 ui = cross(x, wi);                                                             \
    V3 sdw = cross(om, wi), sdu = cross(om, ui) + cross(vo, wi);                      \
    al = fma3(sdw, qd[i], al); ao = fma3(sdu, qd[i], ao);                             \
    om = fma3(wi, qd[i], om); vo = fma3(ui, qd[i], vo);                               \
    rotate<AX>(R, q[i]);                                                              \
    w[i] = wi; u[i] = ui;                                                             \
    Ib[i] = body_inertia(R, x + mulv(R, ld3(LG.ipos[i])), LG.mass[i], LG.inertia[i], 1.f); \
    V3 an, aL, vn, vL;                                                                \
    ri_apply(Ib[i], al, ao, an, aL);                                                  \
    ri_apply(Ib[i], om, vo, vn, vL);                                                  \
    fn[i] = an + cross(om, vn) + cross(vo, vL);                                       \
    fl[i] = aL + cross(om, vL);                                                       \
  }
    LEG_JOINT(0, 2)
    LEG_JOINT(1, 1)
    LEG_JOINT(2, 0)
    LEG_JOINT(3, 1)
    Rshin = R; xshin = x; om_shin = om; vo_shin = vo;
    LEG_JOINT(4, 1)
    LEG_JOINT(5, 0)
    Rfoot = R; xfoot = x; om_foot = om; vo_foot = vo;
#undef LEG_JOINT
  }
  // ---- backward pass: composite inertia -> M blocks, composite force -> bias ----
  Blk M;
  float bias_leg[6];
  RI Ic;
  V3 fcn, fcl;
#pragma unroll
  for (int j = 5; j >= 0; j--) {
    if (j == 5) { Ic = Ib[5]; fcn = fn[5]; fcl = fl[5]; }
    else { Ic = Ic + Ib[j]; fcn = fcn + fn[j]; fcl = fcl + fl[j]; }
    V3 n, l;
    ri_apply(Ic, w[j], u[j], n, l);
    bias_leg[j] = dot(w[j], fcn) + dot(u[j], fcl);
#pragma unroll
    for (int i = 0; i <= j; i++) M.C[TI(j, i)] = dot(w[i], n) + dot(u[i], l);
    M.C[TI(j, j)] += P.armature[6 + 6 * side + j];
    M.B[0 * 6 + j] = l.x; M.B[1 * 6 + j] = l.y; M.B[2 * 6 + j] = l.z;
    M.B[3 * 6 + j] = dot(R0.cx, n); M.B[4 * 6 + j] = dot(R0.cy, n); M.B[5 * 6 + j] = dot(R0.cz, n);
  }
  // ---- root block and root bias (replicated: partner contributions are added as self+partner) ----
  float bias_root[6];
  {
    RI It;
    It.m = I0.m + pair_sum(Ic.m, pm);
    It.mc = I0.mc + pair_sum(Ic.mc, pm);
    It.xx = I0.xx + pair_sum(Ic.xx, pm); It.yy = I0.yy + pair_sum(Ic.yy, pm); It.zz = I0.zz + pair_sum(Ic.zz, pm);
    It.xy = I0.xy + pair_sum(Ic.xy, pm); It.xz = I0.xz + pair_sum(Ic.xz, pm); It.yz = I0.yz + pair_sum(Ic.yz, pm);
    V3 ftn = f0n + pair_sum(fcn, pm), ftl = f0l + pair_sum(fcl, pm);
    V3 c[3] = {R0.cx, R0.cy, R0.cz};
#pragma unroll
    for (int i = 0; i < 21; i++) M.A[i] = 0.f;
    M.A[TI(0, 0)] = M.A[TI(1, 1)] = M.A[TI(2, 2)] = It.m;
#pragma unroll
    for (int b = 0; b < 3; b++) {
      V3 n = rot_inertia_mul(It, c[b]);
      V3 l = cross(c[b], It.mc);
      M.A[TI(3 + b, 0)] = l.x; M.A[TI(3 + b, 1)] = l.y; M.A[TI(3 + b, 2)] = l.z;
#pragma unroll
      for (int a = 0; a <= b; a++) M.A[TI(3 + b, 3 + a)] = dot(c[a], n);
    }
#pragma unroll
    for (int k = 0; k < 6; k++) M.A[TI(k, k)] += P.armature[k];
    bias_root[0] = ftl.x; bias_root[1] = ftl.y; bias_root[2] = ftl.z;
    bias_root[3] = dot(c[0], ftn); bias_root[4] = dot(c[1], ftn); bias_root[5] = dot(c[2], ftn);
  }
  // ---- smooth forces ----
  float fs_leg[6], fs_root[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const int d = 6 + 6 * side + j;
    float lim = P.frc[6 * side + j];
    float t = lim > 0.f ? fminf(fmaxf(tau[j], -lim), lim) : tau[j];
    fs_leg[j] = t - bias_leg[j] - P.damping[d] * qd[j];
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    fs_root[k] = -bias_root[k] - P.damping[k] * rv[k];
    fs_root[3 + k] = -bias_root[3 + k] - P.damping[3 + k] * rw[k];
  }
  // ---- constraint rows owned by this lane ----
  // dof rows: friction loss on own leg dofs and on root dofs 3*side..3*side+2; joint limits on own leg
  float fl_c[6], rfl_c[3];             // B_f * velocity (the -aref of the row)
  float lim_sig[6], lim_c[6], lim_D[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    fl_c[j] = P.floss_B * qd[j];
    const int jj = 6 * side + j;
    float dlo = q[j] - P.range_lo[jj], dhi = P.range_hi[jj] - q[j];
    float sig = dlo < 0.f ? 1.f : (dhi < 0.f ? -1.f : 0.f);
    float dist = dlo < 0.f ? dlo : dhi;
    float imp = impedance(P.limit_imp, dist);
    lim_sig[j] = sig;
    lim_c[j] = sig * P.limit_B * qd[j] + P.limit_K * imp * dist;
    lim_D[j] = sig != 0.f ? 1.f / fmaxf(1e-15f, (1.f - imp) * P.limit_invw[jj] / imp) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) rfl_c[k] = P.floss_B * (side == 0 ? rv[k] : rw[k]);
  // contact candidates
  Contacts CT;
  CT.mask = 0u;
  {
    const float pz = rp[2];
    const float mu2 = mu * mu;
#define CAND(p, Rb, xb, omb, vob, lp, rad, slot)                                                           \
  {                                                                                                        \
    V3 c = xb + mulv(Rb, lp);                                                                              \
    float dist = pz + c.z - (rad);                                                                         \
    if (dist < 0.f) {                                                                                      \
      CT.mask |= 1u << (p);                                                                                \
      V3 rc = mk3(c.x, c.y, 0.5f * dist - pz);                                                             \
      V3 vel = vob + cross(omb, rc);                                                                       \
      float imp = impedance(P.contact_imp, dist);                                                          \
      float tr = P.slot_tran[slot];                                                                        \
      float R0_ = fmaxf(1e-15f, (1.f - imp) * (tr + mu2 * tr) / imp);                                      \
      CT.r[p][0] = rc.x; CT.r[p][1] = rc.y; CT.r[p][2] = rc.z;                                             \
      CT.ub[p][0] = P.contact_B * vel.x; CT.ub[p][1] = P.contact_B * vel.y; CT.ub[p][2] = P.contact_B * vel.z; \
      CT.kap[p] = P.contact_K * imp * dist;                                                                \
      CT.D[p] = 1.f / fmaxf(1e-15f, 2.f * mu2 * R0_);                                                      \
    }                                                                                                      \
  }
#pragma unroll
    for (int p = 0; p < 4; p++) CAND(p, Rfoot, xfoot, om_foot, vo_foot, ld3(LG.foot_pt[p]), 0.f, side)
#pragma unroll
    for (int p = 0; p < 2; p++) CAND(4 + p, Rshin, xshin, om_shin, vo_shin, ld3(LG.shin_pt[p]), LG.shin_rad, 2 + side)
    const V3 zero = mk3(0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < 4; p++) CAND(6 + p, R0, zero, om0, v0, ld3(P.root_pt[4 * side + p]), P.root_rad[4 * side + p], 4)
    if (side == 0) CAND(10, R0, zero, om0, v0, ld3(P.root_pt[8]), P.root_rad[8], 5)
#undef CAND
  }
  const unsigned m_foot = CT.mask & 0xFu, m_shin = CT.mask & 0x30u, m_root = CT.mask & 0x7C0u;

  // ---- solve: phase 0 smooth acceleration, phases 1.. Newton, last phase implicit update ----
  float xl[6], xr[6];        // iterate qacc (leg own, root replicated)
  float Mal[6], Mar[6];      // M (x - x_smooth)
  float jl[6], jr[6];        // J' f at the last evaluation (leg own; root: sum over both lanes)
  V3 F_foot = mk3(0, 0, 0), F_shin = F_foot, F_torso = F_foot, F_pelvis = F_foot;
  int mode = 0, it = 0, capped = 0;
  Fac F;
  for (;;) {
    Blk Hm;
    float rl[6], rr[6];
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 21; i++) { Hm.C[i] = M.C[i]; Hm.A[i] = M.A[i]; }
#pragma unroll
      for (int i = 0; i < 36; i++) Hm.B[i] = M.B[i];
#pragma unroll
      for (int i = 0; i < 6; i++) { rl[i] = fs_leg[i]; rr[i] = fs_root[i]; }
    } else if (mode == 1) {
      // ---- evaluate all rows at x: forces, gradient, Hessian increments ----
      float dC[6], dAo[3];  // diagonal increments from dof rows
      float gl[6], gr_own[6];
#pragma unroll
      for (int j = 0; j < 6; j++) {
        const int d = 6 + 6 * side + j;
        float act;
        float f = floss_force(xl[j] + fl_c[j], P.floss_D[d], P.floss_lim[d], P.floss[d], act);
        float jar = fmaf(lim_sig[j], xl[j], lim_c[j]);
        float la = (lim_sig[j] != 0.f && jar < 0.f) ? lim_D[j] : 0.f;
        f += lim_sig[j] * (-la * jar);
        gl[j] = f;
        dC[j] = act + la;
      }
#pragma unroll
      for (int k = 0; k < 6; k++) gr_own[k] = 0.f;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int d = 3 * side + k;
        float act;
        float f = floss_force(xr[d] + rfl_c[k], P.floss_D[d], P.floss_lim[d], P.floss[d], act);
        gr_own[d] = f;
        dAo[k] = act;
      }
      // body "velocities" of the iterate: V_root, V_shin, V_foot
      V3 Va_root = fma3(R0.cx, xr[3], fma3(R0.cy, xr[4], R0.cz * xr[5])), Vl_root = mk3(xr[0], xr[1], xr[2]);
      V3 Va_shin = Va_root, Vl_shin = Vl_root;
#pragma unroll
      for (int j = 0; j < 4; j++) { Va_shin = fma3(w[j], xl[j], Va_shin); Vl_shin = fma3(u[j], xl[j], Vl_shin); }
      V3 Va_foot = fma3(w[4], xl[4], fma3(w[5], xl[5], Va_shin)), Vl_foot = fma3(u[4], xl[4], fma3(u[5], xl[5], Vl_shin));
      K6 Kf, Ks, Kr;
      k6_zero(Kf); k6_zero(Ks); k6_zero(Kr);
      V3 Wn_f = mk3(0, 0, 0), Wl_f = Wn_f, Wn_s = Wn_f, Wl_s = Wn_f, Wn_r = Wn_f, Wl_r = Wn_f;
      F_foot = F_shin = F_torso = F_pelvis = mk3(0, 0, 0);
#define POINT(p, Va, Vl, Kacc, Wn, Wl, Facc)                                  \
  if (CT.mask & (1u << (p))) {                                                \
    V3 r = ld3(CT.r[p]);                                                      \
    V3 e = Vl + cross(Va, r) + ld3(CT.ub[p]);                                 \
    CT.e[p][0] = e.x; CT.e[p][1] = e.y; CT.e[p][2] = e.z;                     \
    V3 Fp; float Wp[5];                                                       \
    point_eval(e, CT.kap[p], CT.D[p], mu, Fp, Wp);                            \
    k6_add_point(Kacc, r, Wp);                                                \
    Wn = Wn + cross(r, Fp); Wl = Wl + Fp; Facc = Facc + Fp;                   \
  }
      if (m_foot) {
#pragma unroll
        for (int p = 0; p < 4; p++) POINT(p, Va_foot, Vl_foot, Kf, Wn_f, Wl_f, F_foot)
      }
      if (m_shin) {
#pragma unroll
        for (int p = 4; p < 6; p++) POINT(p, Va_shin, Vl_shin, Ks, Wn_s, Wl_s, F_shin)
      }
      if (m_root) {
#pragma unroll
        for (int p = 6; p < 10; p++) POINT(p, Va_root, Vl_root, Kr, Wn_r, Wl_r, F_torso)
        POINT(10, Va_root, Vl_root, Kr, Wn_r, Wl_r, F_pelvis)
      }
#undef POINT
      // J' f: composite wrenches (foot -> joints 4,5 ; foot+shin -> joints 0..3 ; all -> root)
      V3 Wn_fs = Wn_f + Wn_s, Wl_fs = Wl_f + Wl_s;
#pragma unroll
      for (int j = 0; j < 6; j++) gl[j] += (j >= 4) ? dot(w[j], Wn_f) + dot(u[j], Wl_f) : dot(w[j], Wn_fs) + dot(u[j], Wl_fs);
      V3 Wn_t = Wn_fs + Wn_r, Wl_t = Wl_fs + Wl_r;
      gr_own[0] += Wl_t.x; gr_own[1] += Wl_t.y; gr_own[2] += Wl_t.z;
      gr_own[3] += dot(R0.cx, Wn_t); gr_own[4] += dot(R0.cy, Wn_t); gr_own[5] += dot(R0.cz, Wn_t);
      float gn2 = 0.f;
#pragma unroll
      for (int j = 0; j < 6; j++) { jl[j] = gl[j]; rl[j] = gl[j] - Mal[j]; gn2 = fmaf(rl[j], rl[j], gn2); }
      gn2 = pair_sum(gn2, pm);
#pragma unroll
      for (int k = 0; k < 6; k++) { jr[k] = pair_sum(gr_own[k], pm); rr[k] = jr[k] - Mar[k]; gn2 = fmaf(rr[k], rr[k], gn2); }
      F_torso = pair_sum(F_torso, pm);
      F_pelvis = pair_sum(F_pelvis, pm);
      const bool conv = sqrtf(gn2) * P.grad_scale < P.tol;
      if (conv || it >= P.max_iters) {
        capped = !conv;
        mode = 2;
      } else {
        it++;
        // ---- Hessian = M + J' D J ----
#pragma unroll
        for (int i = 0; i < 21; i++) Hm.C[i] = M.C[i];
#pragma unroll
        for (int i = 0; i < 36; i++) Hm.B[i] = M.B[i];
#pragma unroll
        for (int j = 0; j < 6; j++) Hm.C[TI(j, j)] += dC[j];
        float dA[21];
#pragma unroll
        for (int i = 0; i < 21; i++) dA[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) dA[TI(3 * side + k, 3 * side + k)] = dAo[k];
        if (CT.mask) {
          K6 Kc = Kf;  // composite stiffness seen by joints 4,5
#pragma unroll
          for (int j = 5; j >= 0; j--) {
            if (j == 3) k6_add(Kc, Ks);
            V3 n, l;
            k6_apply(Kc, w[j], u[j], n, l);
#pragma unroll
            for (int i = 0; i <= j; i++) Hm.C[TI(j, i)] += dot(w[i], n) + dot(u[i], l);
            Hm.B[0 * 6 + j] += l.x; Hm.B[1 * 6 + j] += l.y; Hm.B[2 * 6 + j] += l.z;
            Hm.B[3 * 6 + j] += dot(R0.cx, n); Hm.B[4 * 6 + j] += dot(R0.cy, n); Hm.B[5 * 6 + j] += dot(R0.cz, n);
          }
          k6_add(Kc, Kr);
          // root block: trans-trans = ll, rot_b-trans = (al' c_b), rot-rot = c_a' aa c_b
          dA[TI(0, 0)] += Kc.ll[0]; dA[TI(1, 1)] += Kc.ll[1]; dA[TI(2, 2)] += Kc.ll[2];
          dA[TI(1, 0)] += Kc.ll[3]; dA[TI(2, 0)] += Kc.ll[4]; dA[TI(2, 1)] += Kc.ll[5];
          V3 c[3] = {R0.cx, R0.cy, R0.cz};
#pragma unroll
          for (int b = 0; b < 3; b++) {
            V3 n, l;
            k6_apply(Kc, c[b], mk3(0.f, 0.f, 0.f), n, l);
            dA[TI(3 + b, 0)] += l.x; dA[TI(3 + b, 1)] += l.y; dA[TI(3 + b, 2)] += l.z;
#pragma unroll
            for (int a = 0; a <= b; a++) dA[TI(3 + b, 3 + a)] += dot(c[a], n);
          }
        }
#pragma unroll
        for (int i = 0; i < 21; i++) Hm.A[i] = M.A[i] + pair_sum(dA[i], pm);
      }
    }
    if (mode == 2) {
      // ---- implicitfast: (M + h*diag(damping)) qacc = f_smooth + J'f ----
#pragma unroll
      for (int i = 0; i < 21; i++) { Hm.C[i] = M.C[i]; Hm.A[i] = M.A[i]; }
#pragma unroll
      for (int i = 0; i < 36; i++) Hm.B[i] = M.B[i];
#pragma unroll
      for (int j = 0; j < 6; j++) {
        Hm.C[TI(j, j)] += h * P.damping[6 + 6 * side + j];
        Hm.A[TI(j, j)] += h * P.damping[j];
        rl[j] = fs_leg[j] + jl[j];
        rr[j] = fs_root[j] + jr[j];
      }
    }
    factor(Hm, F, pm);
    solve(F, rl, rr, pm);
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 6; i++) { xl[i] = rl[i]; xr[i] = rr[i]; Mal[i] = 0.f; Mar[i] = 0.f; jl[i] = 0.f; jr[i] = 0.f; }
      if (use_warm) {  // warm start from the previous substep's acceleration: Ma = M (x_w - x_smooth)
        float dl[6], dr[6];
#pragma unroll
        for (int i = 0; i < 6; i++) { dl[i] = wl[i] - xl[i]; dr[i] = wr[i] - xr[i]; xl[i] = wl[i]; xr[i] = wr[i]; }
        matvec(M, dl, dr, Mal, Mar, pm);
      }
      mode = 1;
      continue;
    }
    if (mode == 2) {
#pragma unroll
      for (int i = 0; i < 6; i++) { xl[i] = rl[i]; xr[i] = rr[i]; }
      break;
    }
    {  // the Newton step no longer moves the acceleration: accept the iterate (J'f of this evaluation is current)
      float smax = 0.f;
#pragma unroll
      for (int i = 0; i < 6; i++) smax = fmaxf(smax, fabsf(rl[i]));
      smax = fmaxf(smax, __shfl_xor_sync(pm, smax, 1));
#pragma unroll
      for (int i = 0; i < 6; i++) smax = fmaxf(smax, fabsf(rr[i]));
      if (smax < P.step_tol) { mode = 2; continue; }
    }
    // ---- exact line search along (rl, rr) ----
    float Msl[6], Msr[6];
    matvec(M, rl, rr, Msl, Msr, pm);
    float sMs = 0.f, sMa = 0.f, gs = 0.f;
#pragma unroll
    for (int i = 0; i < 6; i++) { sMs = fmaf(rl[i], Msl[i], sMs); sMa = fmaf(rl[i], Mal[i], sMa); gs = fmaf(rl[i], jl[i], gs); }
    sMs = pair_sum(sMs, pm); sMa = pair_sum(sMa, pm); gs = pair_sum(gs, pm);
#pragma unroll
    for (int k = 0; k < 6; k++) { sMs = fmaf(rr[k], Msr[k], sMs); sMa = fmaf(rr[k], Mar[k], sMa); gs = fmaf(rr[k], jr[k], gs); }
    const float d10 = sMa - gs;  // phi'(0) = grad . search  (< 0)
    // J_p search for the active points
    {
      V3 Sa_root = fma3(R0.cx, rr[3], fma3(R0.cy, rr[4], R0.cz * rr[5])), Sl_root = mk3(rr[0], rr[1], rr[2]);
      V3 Sa_shin = Sa_root, Sl_shin = Sl_root;
#pragma unroll
      for (int j = 0; j < 4; j++) { Sa_shin = fma3(w[j], rl[j], Sa_shin); Sl_shin = fma3(u[j], rl[j], Sl_shin); }
      V3 Sa_foot = fma3(w[4], rl[4], fma3(w[5], rl[5], Sa_shin)), Sl_foot = fma3(u[4], rl[4], fma3(u[5], rl[5], Sl_shin));
#pragma unroll
      for (int p = 0; p < NPT; p++)
        if (CT.mask & (1u << p)) {
          V3 r = ld3(CT.r[p]);
          V3 s = p < 4 ? Sl_foot + cross(Sa_foot, r) : (p < 6 ? Sl_shin + cross(Sa_shin, r) : Sl_root + cross(Sa_root, r));
          CT.us[p][0] = s.x; CT.us[p][1] = s.y; CT.us[p][2] = s.z;
        }
    }
    float alpha = 1.f, lo = 0.f, hi = 1e30f;
#pragma unroll 1
    for (int ls = 0; ls < 12; ls++) {
      float d1 = 0.f, d2 = 0.f;
#pragma unroll
      for (int j = 0; j < 6; j++) {
        const int d = 6 + 6 * side + j;
        float act;
        float f = floss_force(fmaf(alpha, rl[j], xl[j]) + fl_c[j], P.floss_D[d], P.floss_lim[d], P.floss[d], act);
        d1 = fmaf(-f, rl[j], d1); d2 = fmaf(act * rl[j], rl[j], d2);
        float jv = lim_sig[j] * rl[j];
        float jar = fmaf(lim_sig[j], fmaf(alpha, rl[j], xl[j]), lim_c[j]);
        if (lim_sig[j] != 0.f && jar < 0.f) { d1 = fmaf(lim_D[j] * jar, jv, d1); d2 = fmaf(lim_D[j] * jv, jv, d2); }
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int d = 3 * side + k;
        float act;
        float f = floss_force(fmaf(alpha, rr[d], xr[d]) + rfl_c[k], P.floss_D[d], P.floss_lim[d], P.floss[d], act);
        d1 = fmaf(-f, rr[d], d1); d2 = fmaf(act * rr[d], rr[d], d2);
      }
#pragma unroll
      for (int p = 0; p < NPT; p++)
        if (CT.mask & (1u << p)) point_ls(ld3(CT.e[p]), ld3(CT.us[p]), alpha, CT.kap[p], CT.D[p], mu, d1, d2);
      d1 = pair_sum(d1, pm) + fmaf(alpha, sMs, sMa);
      d2 = pair_sum(d2, pm) + sMs;
      if (fabsf(d1) <= 1e-5f * fabsf(d10) || !(d2 > 0.f)) break;
      if (d1 < 0.f) lo = alpha; else hi = alpha;
      float nx = alpha - d1 / d2;
      if (!(nx > lo && nx < hi)) nx = hi < 1e29f ? 0.5f * (lo + hi) : 2.f * alpha;
      const bool tiny = fabsf(nx - alpha) <= 1e-4f * alpha;
      alpha = nx;
      if (tiny) break;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
      xl[i] = fmaf(alpha, rl[i], xl[i]); xr[i] = fmaf(alpha, rr[i], xr[i]);
      Mal[i] = fmaf(alpha, Msl[i], Mal[i]); Mar[i] = fmaf(alpha, Msr[i], Mar[i]);
    }
  }
  // ---- integrate (semi-implicit Euler; quaternion on SO(3) with the body-frame angular velocity) ----
#pragma unroll
  for (int j = 0; j < 6; j++) { out.qacc[j] = xl[j]; wl[j] = xl[j]; wr[j] = xr[j]; qd[j] = fmaf(h, xl[j], qd[j]); q[j] = fmaf(h, qd[j], q[j]); }
#pragma unroll
  for (int k = 0; k < 3; k++) { rv[k] = fmaf(h, xr[k], rv[k]); rw[k] = fmaf(h, xr[3 + k], rw[k]); rp[k] = fmaf(h, rv[k], rp[k]); }
  {
    float wn = sqrtf(rw[0] * rw[0] + rw[1] * rw[1] + rw[2] * rw[2]);
    float ang = wn * h, dw = 1.f, dx = 0.f, dy = 0.f, dz = 0.f;
    if (ang > 0.f) {
      float s, c;
      sincosf(0.5f * ang, &s, &c);
      s /= wn;
      dw = c; dx = rw[0] * s; dy = rw[1] * s; dz = rw[2] * s;
    }
    float a = rq[0], b = rq[1], c = rq[2], d = rq[3];
    float nw = a * dw - b * dx - c * dy - d * dz, nx = a * dx + b * dw + c * dz - d * dy;
    float ny = a * dy - b * dz + c * dw + d * dx, nz = a * dz + b * dy - c * dx + d * dw;
    float n = rsqrtf(nw * nw + nx * nx + ny * ny + nz * nz);
    rq[0] = nw * n; rq[1] = nx * n; rq[2] = ny * n; rq[3] = nz * n;
  }
  out.F_foot = F_foot; out.F_shin = F_shin; out.F_torso = F_torso; out.F_pelvis = F_pelvis;
  out.iters = it; out.capped = capped;
}

}  // namespace h1v2
