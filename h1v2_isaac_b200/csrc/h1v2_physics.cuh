// h1v2_physics.cuh -- one physics substep of the H1-2 (floating base + 2 x 6-dof legs) for ONE LANE PAIR.
//
// Work decomposition: two adjacent lanes own one environment; lane `side` owns one leg and
// both lanes carry the 6 root dofs bit-identically.  Nothing is ever formed as an 18x18 matrix:
//
//   * every linear system of the step --  M a = f (smooth),  (M + J'DJ + diag) s = -grad (Newton),
//     (M + h*damping) a = f + J'f (implicitfast) -- is the forward dynamics of an articulated body whose link
//     inertias are augmented by the contact "stiffness" K of the links in contact, so it is solved by the
//     articulated-body algorithm: one tip->root sweep per leg (6x6 articulated inertia, rank-1 downdate per joint),
//     a pair exchange of the two legs' hand-offs in root-joint space (27 floats), a 6x6 Cholesky on both lanes and
//     one root->tip sweep.  M x products are one RNE-like sweep.
//   * leg quantities are expressed in world axes about the ANKLE (foot-link origin): the light distal links and the
//     large contact forces sit next to the reference point, which removes the m*r^2 cancellation that limits fp32
//     when everything is referenced to the pelvis (profiles/r1_notes.md).  Root-body quantities are about the pelvis
//     origin; the two meet in root-joint space, which is reference-point independent.
//   * bulk per-joint data (axis, moment arm, link inertia, ABA hand-off) and the active contact list live in a
//     per-thread column of shared memory ([field][thread], conflict free); small per-joint vectors stay in registers.
//
// Semantics restated from the reference's physics target (MuJoCo 3.3.6 as configured by
// packages/biped_deploy/biped_deploy/simulator/sim_mujoco.py:39-41 on
// packages/biped_assets/biped_assets/models/h12/scene/h12_12dof.xml):
// RNE bias, soft constraints (friction loss, joint limits, pyramidal contacts with solref/solimp),
// primal Newton with exact line search, implicitfast velocity update, semi-implicit Euler.
#pragma once
#include "h1v2_math.cuh"
#include "h1v2_params.h"
#include "h1v2_terrain.cuh"

namespace h1v2 {

#define H1V2_BLOCK 32
// unroll factors of the joint loops inside the Newton trip (tuned on B200, tools/diag_unroll.sh; 1 = rolled)
#define H1V2_PRAGMA(x) _Pragma(#x)
#define UNROLL(n) H1V2_PRAGMA(unroll n)
#ifndef U_SWEEP1
#define U_SWEEP1 1
#endif
#ifndef U_SWEEP2
#define U_SWEEP2 1
#endif
#ifndef U_MPROD
#define U_MPROD 1
#endif
#ifndef U_LSJ
#define U_LSJ 1
#endif
#ifndef U_EVJ
#define U_EVJ 1
#endif
#ifndef U_EVP
#define U_EVP 1  // contact-point loops of the Newton trip: evaluate | line search | step (tools/build_variants.sh experiments)
#endif
#ifndef U_LSP
#define U_LSP 1
#endif
#ifndef U_STP
#define U_STP 1
#endif
#ifndef H1V2_FINAL_MX
#define H1V2_FINAL_MX 1  // rhs of the implicit update = M x (1) or f_smooth + J'f re-evaluated (0, round 1)
#endif
#ifndef U_PRO
#define U_PRO 1  // once-per-substep joint loops (kinematics passes, smooth forces)
#endif
// shared-memory column of one thread: 6 joints x JSTRIDE floats, then MAXC contact points x PSTRIDE floats.
// 220 floats per thread = 28160 B per warp: 8 warps per SM ((28160 + 1024 reserved) * 8 = 228 KB exactly).  Measured on
// B200: throughput still grows linearly with resident warps at 7 per SM (tools: H1V2_SMEM_PAD experiment, profiles/r1_notes.md),
// so every real here is worth keeping out: the link mass comes from the parameter block, the ABA 1/D shares a slot
// with the extra-diagonal / M-product scratch, and the active contact list holds 5 points.
#define JSTRIDE 30
#define F_W 0      // 3  joint axis (world)
#define F_U 3      // 3  (joint position - O_s) x axis
#define F_I 6      // 9  link inertia about O_s: m*c (3) xx yy zz xy xz yz   (the mass itself is KLeg::mass)
#define F_X 15     // 6  scratch: RNE link force | ABA U = IA*S (n,l) | M-product link force
#define F_DG 21    // 1  extra diagonal of the current solve -> ABA 1/(S'IA S + armature + diag) -> (M s) after the M-product
#define F_DINV F_DG
#define F_XQ 22    // 1  acceleration iterate
#define F_R 23     // 1  rhs of the current solve -> ABA reduced rhs -> solution / search direction
#define F_MA 24    // 1  (M x - f_smooth)
#define F_G 25     // 1  J'f at the last evaluation
#define F_FLC 26   // 1  friction-loss row offset  B*qd
#define F_LIMC 27  // 1  limit row offset
#define F_LIMD 28  // 1  limit row sign * D (0 = no limit row)
#define F_FS 29    // 1  smooth force
#define MAXC 5     // 4 sole corners + 1: any further point means shin / torso / pelvis on the ground, i.e. the env terminates this step
#define PSTRIDE 8  // r (3) | row residual e (3) | K*imp*dist | 1/R   (owner link from the list position)
#define PSTRIDE_ROUGH 11  // rough instantiation: + the contact normal (3); e is then held in the contact frame
#define PT_BASE (6 * JSTRIDE)
#define SMEM_FLOATS (PT_BASE + MAXC * PSTRIDE)
#define SMEM_FLOATS_ROUGH (PT_BASE + MAXC * PSTRIDE_ROUGH)  // 235 floats per thread: 7 warps per SM

template <int PS>
struct SmemT {
  real* base;  // this thread's column
  __device__ __forceinline__ real& jf(int j, int f) const { return base[(j * JSTRIDE + f) * H1V2_BLOCK]; }
  __device__ __forceinline__ real& pf(int p, int f) const { return base[(PT_BASE + p * PS + f) * H1V2_BLOCK]; }
  __device__ __forceinline__ V3 jv(int j, int f) const { return mk3(jf(j, f), jf(j, f + 1), jf(j, f + 2)); }
  __device__ __forceinline__ void sjv(int j, int f, V3 v) const { jf(j, f) = v.x; jf(j, f + 1) = v.y; jf(j, f + 2) = v.z; }
  __device__ __forceinline__ V3 pv(int p, int f) const { return mk3(pf(p, f), pf(p, f + 1), pf(p, f + 2)); }
  __device__ __forceinline__ RI ji(int j, real m) const {
    RI I;
    I.m = m; I.mc = jv(j, F_I);
    I.xx = jf(j, F_I + 3); I.yy = jf(j, F_I + 4); I.zz = jf(j, F_I + 5); I.xy = jf(j, F_I + 6); I.xz = jf(j, F_I + 7); I.yz = jf(j, F_I + 8);
    return I;
  }
  __device__ __forceinline__ void sji(int j, const RI& I) const {
    sjv(j, F_I, I.mc);
    jf(j, F_I + 3) = I.xx; jf(j, F_I + 4) = I.yy; jf(j, F_I + 5) = I.zz; jf(j, F_I + 6) = I.xy; jf(j, F_I + 7) = I.xz; jf(j, F_I + 8) = I.yz;
  }
};
typedef SmemT<PSTRIDE> Smem;

// Sum over the two lanes of an env.  ALWAYS executed by the whole, converged warp with the constant full mask: a shuffle
// with a per-pair register mask compiles to a BSSY / WARPSYNC.COLLECTIVE / SHFL / ENDCOLLECTIVE / BSYNC sequence through
// fixed registers (~13 instructions and spills, profiles/r1g), the constant-mask one to a single SHFL.  The Newton loop
// below is therefore warp-synchronous: every lane runs every trip, lanes whose solve is finished are masked by selects.
#define FULL_MASK 0xffffffffu
__device__ __forceinline__ real pair_sum(real v) { return v + __shfl_xor_sync(FULL_MASK, v, 1); }
__device__ __forceinline__ V3 pair_sum(V3 v) { return mk3(pair_sum(v.x), pair_sum(v.y), pair_sum(v.z)); }
// Mirror-lane instantiations (fewer than 16 envs per warp): the lanes that would idle MIRROR the working ones bit for bit (same
// env, same leg, same shared-memory column), and the loops whose iterations are independent -- joint rows of the gradient
// evaluation, sole points, rows and contact points of the line search, the step -- are dealt to the mirrors round-robin (same
// instruction stream, different index: no divergence); this sum over the mirrors closes them.  Addition commutes, so all
// mirrors hold the same bits afterwards and stay mirrors.
// NM = mirrors per lane: 1 (plain, 16 envs per warp), 2 (8 envs per warp, lanes l and l+16), 4 (4 envs per warp: l, l+8, l+16, l+24 -- the
// two butterfly stages add symmetric pairs, so all four end with the same bits too).
template <int NM>
__device__ __forceinline__ real mirror_sum(real v) {
  if (NM == 4) v = v + __shfl_xor_sync(FULL_MASK, v, 8);
  return NM > 1 ? v + __shfl_xor_sync(FULL_MASK, v, 16) : v;
}
template <int NM>
__device__ __forceinline__ real mirror_max(real v) {
  if (NM == 4) v = r_max(v, __shfl_xor_sync(FULL_MASK, v, 8));
  return NM > 1 ? r_max(v, __shfl_xor_sync(FULL_MASK, v, 16)) : v;
}

// MuJoCo getimpedance (solimp = d0, dmax, width, midpoint, power), margin 0
__device__ __forceinline__ real impedance(const float* si, real pos) {
  real dmin = r_min(r_max(si[0], 1e-4f), 0.9999f), dmax = r_min(r_max(si[1], 1e-4f), 0.9999f);
  real width = si[2], mid = r_min(r_max(si[3], 1e-4f), 0.9999f), power = r_max(si[4], 1.f);
  if (dmin == dmax || width <= 1e-15f) return 0.5f * (dmin + dmax);
  real x = r_abs(pos) / width;
  if (x >= 1.f) return dmax;
  if (x == 0.f) return dmin;
  real y;
  if (power == 1.f) y = x;
  else if (power == 2.f) y = (x <= mid) ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);
  else y = (x <= mid) ? r_pow(x, power) / r_pow(mid, power - 1.f) : 1.f - r_pow(1.f - x, power) / r_pow(1.f - mid, power - 1.f);
  return dmin + y * (dmax - dmin);
}

// general symmetric 6x6 in [ang;lin] order: aa (xx yy zz xy xz yz), al (row = ang, col = lin), ll (xx yy zz xy xz yz)
struct K6 {
  real aa[6], al[9], ll[6];
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
// K += rigid inertia: n = I w + mc x u  ->  al = [mc]x ;  l = m u + w x mc
__device__ __forceinline__ void k6_add_rigid(K6& K, const RI& I) {
  K.aa[0] += I.xx; K.aa[1] += I.yy; K.aa[2] += I.zz; K.aa[3] += I.xy; K.aa[4] += I.xz; K.aa[5] += I.yz;
  K.ll[0] += I.m; K.ll[1] += I.m; K.ll[2] += I.m;
  K.al[1] -= I.mc.z; K.al[2] += I.mc.y; K.al[3] += I.mc.z; K.al[5] -= I.mc.x; K.al[6] -= I.mc.y; K.al[7] += I.mc.x;
}
__device__ __forceinline__ V3 sym3_mul(const real* s, V3 v) {
  return mk3(r_fma(s[0], v.x, r_fma(s[3], v.y, s[4] * v.z)), r_fma(s[3], v.x, r_fma(s[1], v.y, s[5] * v.z)),
             r_fma(s[4], v.x, r_fma(s[5], v.y, s[2] * v.z)));
}
// (n,l) = K (w,u)
__device__ __forceinline__ void k6_apply(const K6& K, V3 w, V3 u, V3& n, V3& l) {
  n = sym3_mul(K.aa, w) + mk3(r_fma(K.al[0], u.x, r_fma(K.al[1], u.y, K.al[2] * u.z)), r_fma(K.al[3], u.x, r_fma(K.al[4], u.y, K.al[5] * u.z)),
                              r_fma(K.al[6], u.x, r_fma(K.al[7], u.y, K.al[8] * u.z)));
  l = sym3_mul(K.ll, u) + mk3(r_fma(K.al[0], w.x, r_fma(K.al[3], w.y, K.al[6] * w.z)), r_fma(K.al[1], w.x, r_fma(K.al[4], w.y, K.al[7] * w.z)),
                              r_fma(K.al[2], w.x, r_fma(K.al[5], w.y, K.al[8] * w.z)));
}
// K -= (n;l)(n;l)' * s
__device__ __forceinline__ void k6_rank1_sub(K6& K, V3 n, V3 l, real s) {
  V3 ns = n * s, ls = l * s;
  K.aa[0] = r_fma(-ns.x, n.x, K.aa[0]); K.aa[1] = r_fma(-ns.y, n.y, K.aa[1]); K.aa[2] = r_fma(-ns.z, n.z, K.aa[2]);
  K.aa[3] = r_fma(-ns.x, n.y, K.aa[3]); K.aa[4] = r_fma(-ns.x, n.z, K.aa[4]); K.aa[5] = r_fma(-ns.y, n.z, K.aa[5]);
  K.ll[0] = r_fma(-ls.x, l.x, K.ll[0]); K.ll[1] = r_fma(-ls.y, l.y, K.ll[1]); K.ll[2] = r_fma(-ls.z, l.z, K.ll[2]);
  K.ll[3] = r_fma(-ls.x, l.y, K.ll[3]); K.ll[4] = r_fma(-ls.x, l.z, K.ll[4]); K.ll[5] = r_fma(-ls.y, l.z, K.ll[5]);
  K.al[0] = r_fma(-ns.x, l.x, K.al[0]); K.al[1] = r_fma(-ns.x, l.y, K.al[1]); K.al[2] = r_fma(-ns.x, l.z, K.al[2]);
  K.al[3] = r_fma(-ns.y, l.x, K.al[3]); K.al[4] = r_fma(-ns.y, l.y, K.al[4]); K.al[5] = r_fma(-ns.y, l.z, K.al[5]);
  K.al[6] = r_fma(-ns.z, l.x, K.al[6]); K.al[7] = r_fma(-ns.z, l.y, K.al[7]); K.al[8] = r_fma(-ns.z, l.z, K.al[8]);
}

// K about O_s  ->  K about O_r, with O_s = O_r + d  (motion at O_s: u_s = u_r + w x d):  K_r = X' K_s X, X = [[1,0],[-[d]x,1]]
//   ll_r = ll ;  al_r = al + [d]x ll ;  aa_r = aa - al_r [d]x + [d]x al'
__device__ __forceinline__ void k6_shift(K6& K, V3 d) {
  const V3 c0 = cross(d, mk3(K.ll[0], K.ll[3], K.ll[4])), c1 = cross(d, mk3(K.ll[3], K.ll[1], K.ll[5])), c2 = cross(d, mk3(K.ll[4], K.ll[5], K.ll[2]));
  const V3 b0 = mk3(K.al[0], K.al[1], K.al[2]), b1 = mk3(K.al[3], K.al[4], K.al[5]), b2 = mk3(K.al[6], K.al[7], K.al[8]);
  const V3 r0 = b0 + mk3(c0.x, c1.x, c2.x), r1 = b1 + mk3(c0.y, c1.y, c2.y), r2 = b2 + mk3(c0.z, c1.z, c2.z);
  const V3 p0 = cross(d, r0), p1 = cross(d, r1), p2 = cross(d, r2);
  const V3 q0 = cross(d, b0), q1 = cross(d, b1), q2 = cross(d, b2);
  K.aa[0] += p0.x + q0.x; K.aa[1] += p1.y + q1.y; K.aa[2] += p2.z + q2.z;
  K.aa[3] += p0.y + q1.x; K.aa[4] += p0.z + q2.x; K.aa[5] += p1.z + q2.y;
  K.al[0] = r0.x; K.al[1] = r0.y; K.al[2] = r0.z; K.al[3] = r1.x; K.al[4] = r1.y; K.al[5] = r1.z; K.al[6] = r2.x; K.al[7] = r2.y; K.al[8] = r2.z;
}

// one contact point at the iterate.  Edge residuals j_k = a_k . e + kappa over the four pyramid edges
// a_k = (0,mu,1) (0,-mu,1) (-mu,0,1) (mu,0,1)   (normal z, tangents y and -x: mju_makeFrame on the plane normal);
// an edge is active where j_k < 0, with force -D*j_k along a_k.
struct Edges {
  real j0, j1, j2, j3;
};
__device__ __forceinline__ Edges point_edges(V3 e, real kap, real mu) {
  Edges E;
  E.j0 = r_fma(mu, e.y, e.z) + kap; E.j1 = r_fma(-mu, e.y, e.z) + kap;
  E.j2 = r_fma(-mu, e.x, e.z) + kap; E.j3 = r_fma(mu, e.x, e.z) + kap;
  return E;
}
// contact force of the point (world axes)
__device__ __forceinline__ V3 point_force(V3 e, real kap, real D, real mu) {
  const Edges E = point_edges(e, kap, mu);
  const real f0 = E.j0 < 0.f ? -D * E.j0 : 0.f, f1 = E.j1 < 0.f ? -D * E.j1 : 0.f;
  const real f2 = E.j2 < 0.f ? -D * E.j2 : 0.f, f3 = E.j3 < 0.f ? -D * E.j3 : 0.f;
  return mk3(mu * (f3 - f2), mu * (f0 - f1), (f0 + f1) + (f2 + f3));
}
// 3x3 weight W = D * sum_active a a'  as  xx yy zz xz yz
__device__ __forceinline__ void point_weight(V3 e, real kap, real D, real mu, real (&W)[5]) {
  const Edges E = point_edges(e, kap, mu);
  const real a0 = E.j0 < 0.f ? D : 0.f, a1 = E.j1 < 0.f ? D : 0.f, a2 = E.j2 < 0.f ? D : 0.f, a3 = E.j3 < 0.f ? D : 0.f;
  W[0] = mu * mu * (a2 + a3);
  W[1] = mu * mu * (a0 + a1);
  W[2] = (a0 + a1) + (a2 + a3);
  W[3] = mu * (a3 - a2);
  W[4] = mu * (a0 - a1);
}
__device__ __forceinline__ V3 w5_mul(const real (&W)[5], V3 g) {
  return mk3(W[0] * g.x + W[3] * g.z, W[1] * g.y + W[4] * g.z, W[3] * g.x + W[4] * g.y + W[2] * g.z);
}
// K += X' W X with X = [G | 1], G = -[r]x   (point velocity u = v_O + w x r)
__device__ __forceinline__ void k6_add_point(K6& K, V3 r, const real (&W)[5]) {
  V3 g0 = mk3(0.f, -r.z, r.y), g1 = mk3(r.z, 0.f, -r.x), g2 = mk3(-r.y, r.x, 0.f);  // columns of G: e_j x r
  V3 w0 = w5_mul(W, g0), w1 = w5_mul(W, g1), w2 = w5_mul(W, g2);
  K.aa[0] += dot(g0, w0); K.aa[1] += dot(g1, w1); K.aa[2] += dot(g2, w2);
  K.aa[3] += dot(g0, w1); K.aa[4] += dot(g0, w2); K.aa[5] += dot(g1, w2);
  K.al[0] += w0.x; K.al[1] += w0.y; K.al[2] += w0.z;
  K.al[3] += w1.x; K.al[4] += w1.y; K.al[5] += w1.z;
  K.al[6] += w2.x; K.al[7] += w2.y; K.al[8] += w2.z;
  K.ll[0] += W[0]; K.ll[1] += W[1]; K.ll[2] += W[2]; K.ll[4] += W[3]; K.ll[5] += W[4];
}
// the same for a point whose weight W is given in its contact frame (axes ax, ay, n): K += X' (L W L') X
__device__ __forceinline__ void k6_add_point_rot(K6& K, V3 r, const real (&W)[5], V3 n) {
  const CFrame cf = cframe(n);
  V3 ax, ay;
  contact_axes(cf, ax, ay);
  // columns of Ww = L W L':  Ww e_k = L (W (L' e_k)),  L' e_k = (ax_k, ay_k, n_k)
  const V3 w0 = from_contact(cf, w5_mul(W, mk3(ax.x, ay.x, n.x)));
  const V3 w1 = from_contact(cf, w5_mul(W, mk3(ax.y, ay.y, n.y)));
  const V3 w2 = from_contact(cf, w5_mul(W, mk3(ax.z, ay.z, n.z)));
  // Ww g_j over the columns g_j = e_j x r of G
  const V3 g0 = mk3(0.f, -r.z, r.y), g1 = mk3(r.z, 0.f, -r.x), g2 = mk3(-r.y, r.x, 0.f);
  const V3 v0 = fma3(w1, g0.y, w2 * g0.z), v1 = fma3(w0, g1.x, w2 * g1.z), v2 = fma3(w0, g2.x, w1 * g2.y);
  K.aa[0] += dot(g0, v0); K.aa[1] += dot(g1, v1); K.aa[2] += dot(g2, v2);
  K.aa[3] += dot(g0, v1); K.aa[4] += dot(g0, v2); K.aa[5] += dot(g1, v2);
  K.al[0] += v0.x; K.al[1] += v0.y; K.al[2] += v0.z;
  K.al[3] += v1.x; K.al[4] += v1.y; K.al[5] += v1.z;
  K.al[6] += v2.x; K.al[7] += v2.y; K.al[8] += v2.z;
  K.ll[0] += w0.x; K.ll[1] += w1.y; K.ll[2] += w2.z; K.ll[3] += w0.y; K.ll[4] += w0.z; K.ll[5] += w1.z;
}
// line-search contribution of one contact point: d1 += D*jar*jv, d2 += D*jv^2 over the edges active at ea
__device__ __forceinline__ void point_ls(V3 ea, V3 us, real kap, real D, real mu, real& d1, real& d2) {
  const Edges E = point_edges(ea, kap, mu);
  const real v0 = r_fma(mu, us.y, us.z), v1 = r_fma(-mu, us.y, us.z), v2 = r_fma(-mu, us.x, us.z), v3 = r_fma(mu, us.x, us.z);
  // branch-free: an inactive edge contributes min(j,0) = 0 to the first sum and a zeroed velocity to the second
  const real w0 = E.j0 < 0.f ? v0 : 0.f, w1 = E.j1 < 0.f ? v1 : 0.f, w2 = E.j2 < 0.f ? v2 : 0.f, w3 = E.j3 < 0.f ? v3 : 0.f;
  const real s1 = r_fma(E.j0, w0, r_fma(E.j1, w1, r_fma(E.j2, w2, E.j3 * w3)));
  const real s2 = r_fma(w0, w0, r_fma(w1, w1, r_fma(w2, w2, w3 * w3)));
  d1 = r_fma(D, s1, d1);
  d2 = r_fma(D, s2, d2);
}
// friction-loss row: force and activity (= D inside the quadratic zone) at residual jar.  The zone boundary R*fl is where
// the linear force -D*jar reaches +-fl (D = 1/R), so the row is a clamp: no branches and no third constant.
__device__ __forceinline__ real floss_force(real jar, real D, real fl, real& act) {
  const real f = -D * jar;
  act = r_abs(f) < fl ? D : 0.f;
  return r_min(r_max(f, -fl), fl);
}

struct SubOut {
  V3 F_foot, F_shin, F_torso, F_pelvis;  // net contact forces (torso/pelvis pair-summed, valid on both lanes)
  real qacc[6];                         // joint accelerations of this leg after the implicit update
  int iters, capped, overflow;
#ifdef H1V2_WARPCLOCK
  int trips, lstrips;  // diagnostic variant only (tools/diag_warpclock.py): warp-synchronous trips of the Newton loop / line search
#endif
};

// root-joint-space projection helpers.  Root dofs seen from reference point O_s = O_r + d:
//   translation k: (0, e_k) ; rotation k: (c_k, c_k x d)   with c_k = columns of the root rotation
struct RootBasis {
  V3 c[3], rd[3];
};
__device__ __forceinline__ void root_project_force(const RootBasis& B, V3 n, V3 l, real (&out)[6]) {
  out[0] = l.x; out[1] = l.y; out[2] = l.z;
#pragma unroll
  for (int k = 0; k < 3; k++) out[3 + k] = dot(B.c[k], n) + dot(B.rd[k], l);
}
// A (packed lower 6x6) += S_r' K S_r
__device__ __forceinline__ void root_project_k6(const RootBasis& B, const K6& K, real (&A)[21]) {
  A[TI(0, 0)] += K.ll[0]; A[TI(1, 1)] += K.ll[1]; A[TI(2, 2)] += K.ll[2];
  A[TI(1, 0)] += K.ll[3]; A[TI(2, 0)] += K.ll[4]; A[TI(2, 1)] += K.ll[5];
#pragma unroll
  for (int b = 0; b < 3; b++) {
    V3 n, l;
    k6_apply(K, B.c[b], B.rd[b], n, l);
    A[TI(3 + b, 0)] += l.x; A[TI(3 + b, 1)] += l.y; A[TI(3 + b, 2)] += l.z;
#pragma unroll
    for (int a = 0; a <= b; a++) A[TI(3 + b, 3 + a)] += dot(B.c[a], n) + dot(B.rd[a], l);
  }
}
// spatial "velocity" of the root body for generalized vector xr, about O_r + d
__device__ __forceinline__ void root_motion(const RootBasis& B, const real (&xr)[6], V3 d, V3& a, V3& l) {
  a = fma3(B.c[0], xr[3], fma3(B.c[1], xr[4], B.c[2] * xr[5]));
  l = mk3(xr[0], xr[1], xr[2]) + cross(a, d);
}

// same about the pelvis origin itself (d = 0): rotation k is (c_k, 0)
__device__ __forceinline__ void root_project_force0(const V3 (&c)[3], V3 n, V3 l, real (&out)[6]) {
  out[0] = l.x; out[1] = l.y; out[2] = l.z;
#pragma unroll
  for (int k = 0; k < 3; k++) out[3 + k] = dot(c[k], n);
}
__device__ __forceinline__ void root_project_k60(const V3 (&c)[3], const K6& K, real (&A)[21]) {
  A[TI(0, 0)] += K.ll[0]; A[TI(1, 1)] += K.ll[1]; A[TI(2, 2)] += K.ll[2];
  A[TI(1, 0)] += K.ll[3]; A[TI(2, 0)] += K.ll[4]; A[TI(2, 1)] += K.ll[5];
#pragma unroll
  for (int b = 0; b < 3; b++) {
    const V3 n = sym3_mul(K.aa, c[b]);
    const V3 l = mk3(r_fma(K.al[0], c[b].x, r_fma(K.al[3], c[b].y, K.al[6] * c[b].z)), r_fma(K.al[1], c[b].x, r_fma(K.al[4], c[b].y, K.al[7] * c[b].z)),
                     r_fma(K.al[2], c[b].x, r_fma(K.al[5], c[b].y, K.al[8] * c[b].z)));
    A[TI(3 + b, 0)] += l.x; A[TI(3 + b, 1)] += l.y; A[TI(3 + b, 2)] += l.z;
#pragma unroll
    for (int a = 0; a <= b; a++) A[TI(3 + b, 3 + a)] += dot(c[a], n);
  }
}

// joint axis in the link frame: hip_yaw z, hip_pitch y, hip_roll x, knee y, ankle_pitch y, ankle_roll x (h12_12dof.xml:71-96);
// checked against the compiled model by build_params.  The index is warp-uniform, so the branches below do not diverge.
__device__ __forceinline__ int joint_axis(int i) { return (0x146 >> (2 * i)) & 3; }
__device__ __forceinline__ void rotate_rt(M3& R, int ax, real s, real c) {
  if (ax == 2) rotate_sc<2>(R, s, c);
  else if (ax == 1) rotate_sc<1>(R, s, c);
  else rotate_sc<0>(R, s, c);
}
__device__ __forceinline__ V3 axis_rt(const M3& R, int ax) { return ax == 0 ? R.cx : (ax == 1 ? R.cy : R.cz); }
__device__ __noinline__ real impedance_general(const float* si, real pos) { return impedance(si, pos); }
// MuJoCo's default solimp (power 2, every geom and joint of h12_12dof.xml) inline; anything else out of line (powf paths)
__device__ __forceinline__ real impedance_call(const float* si, real pos) {
  if (si[4] != 2.f || si[2] <= 1e-15f) return impedance_general(si, pos);
  const real dmin = r_min(r_max(si[0], 1e-4f), 0.9999f), dmax = r_min(r_max(si[1], 1e-4f), 0.9999f);
  const real mid = r_min(r_max(si[3], 1e-4f), 0.9999f);
  if (dmin == dmax) return 0.5f * (dmin + dmax);
  const real x = r_abs(pos) / si[2];
  if (x >= 1.f) return dmax;
  if (x == 0.f) return dmin;
  const real y = (x <= mid) ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);
  return dmin + y * (dmax - dmin);
}

// ----------------------------------------------------------------------------------------------------------
// One physics substep.  State is updated in place.  tau = joint torques after the actuator's effort clip.
// wl/wr: warm start in (if use_warm), solution out.
//
// Code-size discipline (profiles/r1d: the first version was instruction-fetch bound -- 242 KB of SASS, 65 % hit rate
// in the SM instruction cache, the GPC instruction cache at 74 % of its request rate): every loop over joints or
// contact points is rolled, per-joint inputs are staged in the shared-memory column so that rolled loops can index
// them, each helper has one call site inside the Newton loop, and nothing is kept in registers between the
// "evaluate" and "solve" halves of an iteration that can be re-derived from the contact list (the contact
// stiffness is accumulated straight into the articulated inertia during the tip->root sweep).
// ----------------------------------------------------------------------------------------------------------
// ROUGH: contacts against the height field of the Rough id (terrain height and triangle normal under every candidate, residuals
// held in the contact frame) -- its own instantiation, the plane kernel carries none of that code.
template <bool ROUGH, int NM>
__device__ __forceinline__ void substep(const KParams& P, const unsigned tid, const int side, real (&rp)[3],
                                     real (&rq)[4], real (&rv)[3], real (&rw)[3], real (&q)[6], real (&qd)[6],
                                     const real (&tau)[6], const real mu, const real mass_add, real (&wl)[6], real (&wr)[6],
                                     const bool use_warm, SubOut& out, const float* __restrict__ terrain_h, const TerrainEnv& te) {
  extern __shared__ __align__(16) real smem_raw[];
  constexpr bool QUAD = NM > 1;  // mirror-lane instantiation
  const SmemT<ROUGH ? PSTRIDE_ROUGH : PSTRIDE> sm{smem_raw + (NM == 4 ? (tid & 7u) : NM == 2 ? (tid & 15u) : tid)};
  const int q0 = NM == 4 ? (int)(tid >> 3) : NM == 2 ? (int)(tid >> 4) : 0;  // first iteration of a dealt loop; its stride is QS
  constexpr int QS = NM;
  const KLeg& LG = P.leg[side];
  const real h = P.h;
  const int j0 = 6 * side;
  const V3 zero3 = mk3(0.f, 0.f, 0.f);
  // ---- stage the per-joint inputs in the column: q -> F_XQ, qd -> F_FLC, tau -> F_FS, warm start -> F_R ----
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const real lim = P.frc[j0 + j];
    sm.jf(j, F_XQ) = q[j]; sm.jf(j, F_FLC) = qd[j];
    sm.jf(j, F_FS) = lim > 0.f ? r_min(r_max(tau[j], -lim), lim) : tau[j];
    sm.jf(j, F_R) = use_warm ? wl[j] : 0.f;
  }
  // ---- root frame (about O_r = pelvis origin) ----
  {
    real n = r_rsqrt(rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3]);
    rq[0] *= n; rq[1] *= n; rq[2] *= n; rq[3] *= n;
  }
  const M3 R0 = quat2mat(rq[0], rq[1], rq[2], rq[3]);
  const V3 v0 = mk3(rv[0], rv[1], rv[2]);
  const V3 om0 = mulv(R0, mk3(rw[0], rw[1], rw[2]));
  const V3 acc0 = mk3(0.f, 0.f, P.gravity) + cross(v0, om0);  // root cacc: -gravity + free-joint cdof_dot * qvel
  const real m0 = P.root_mass + mass_add;
  const RI I0 = body_inertia(R0, mulv(R0, ld3(P.root_ipos)), m0, P.root_inertia, P.mass_scales_inertia ? m0 / P.root_mass : 1.f);
  real fs_root[6];  // first the bias force of the root link about O_r, then the smooth force of the root dofs
  {
    V3 an, al, vn, vl;
    ri_apply(I0, zero3, acc0, an, al);
    ri_apply(I0, om0, v0, vn, vl);
    const V3 c3[3] = {R0.cx, R0.cy, R0.cz};
    root_project_force0(c3, an + cross(om0, vn) + cross(v0, vl), al + cross(om0, vl), fs_root);
  }
  // ---- pass 1: joint sines/cosines (-> F_G, F_DG) and the ankle position d (O_s = O_r + d) ----
  V3 d;
  {
    M3 R = R0;
    V3 x = zero3;
UNROLL(U_PRO)
    for (int i = 0; i < 6; i++) {
      x = x + mulv(R, ld3(LG.pos[i]));
      real s, c;
      sincos_lim(sm.jf(i, F_XQ), s, c);
      sm.jf(i, F_G) = s; sm.jf(i, F_DG) = c;
      rotate_rt(R, joint_axis(i), s, c);
    }
    d = x;
  }
  RootBasis RB;  // root dofs seen from O_s
  RB.c[0] = R0.cx; RB.c[1] = R0.cy; RB.c[2] = R0.cz;
#pragma unroll
  for (int k = 0; k < 3; k++) RB.rd[k] = cross(RB.c[k], d);
  // ---- pass 2: kinematics about O_s, velocities, bias accelerations, link inertia and RNE link force ----
  M3 Rshin, Rfoot;
  V3 xshin, om_shin, vo_shin, om_foot, vo_foot;
  {
    M3 R = R0;
    V3 x = zero3 - d, om = om0, vo = v0 + cross(om0, d), al = zero3, ao = acc0;
UNROLL(U_PRO)
    for (int i = 0; i < 6; i++) {
      const int ax = joint_axis(i);
      const real qdi = sm.jf(i, F_FLC);
      x = x + mulv(R, ld3(LG.pos[i]));
      const V3 wi = axis_rt(R, ax);
      const V3 ui = cross(x, wi);
      const V3 sdw = cross(om, wi), sdu = cross(om, ui) + cross(vo, wi);
      al = fma3(sdw, qdi, al); ao = fma3(sdu, qdi, ao);
      om = fma3(wi, qdi, om); vo = fma3(ui, qdi, vo);
      rotate_rt(R, ax, sm.jf(i, F_G), sm.jf(i, F_DG));
      sm.sjv(i, F_W, wi); sm.sjv(i, F_U, ui);
      const RI Ii = body_inertia(R, x + mulv(R, ld3(LG.ipos[i])), LG.mass[i], LG.inertia[i], 1.f);
      sm.sji(i, Ii);
      V3 an, aL, vn, vL;
      ri_apply(Ii, al, ao, an, aL);
      ri_apply(Ii, om, vo, vn, vL);
      sm.sjv(i, F_X, an + cross(om, vn) + cross(vo, vL));
      sm.sjv(i, F_X + 3, aL + cross(om, vL));
      if (i == 3) { Rshin = R; xshin = x; om_shin = om; vo_shin = vo; }
    }
    Rfoot = R; om_foot = om; vo_foot = vo;
  }
  // ---- smooth forces: tau - bias - damping*v ; dof-row constants (friction loss, joint limits) ----
  {
    V3 fcn = zero3, fcl = zero3;
UNROLL(U_PRO)
    for (int j = 5; j >= 0; j--) {
      fcn = fcn + sm.jv(j, F_X); fcl = fcl + sm.jv(j, F_X + 3);
      const int jj = j0 + j;
      const real qj = sm.jf(j, F_XQ), qdj = sm.jf(j, F_FLC);
      const real fs = sm.jf(j, F_FS) - (dot(sm.jv(j, F_W), fcn) + dot(sm.jv(j, F_U), fcl)) - P.damping[6 + jj] * qdj;
      sm.jf(j, F_FS) = fs;
      sm.jf(j, F_XQ) = 0.f; sm.jf(j, F_MA) = -fs; sm.jf(j, F_G) = 0.f;
      sm.jf(j, F_FLC) = P.floss_B * qdj;
      const real dlo = qj - P.range_lo[jj], dhi = P.range_hi[jj] - qj;
      const real sig = dlo < 0.f ? 1.f : (dhi < 0.f ? -1.f : 0.f);
      real lc = 0.f, lD = 0.f;
      if (sig != 0.f) {
        const real dist = dlo < 0.f ? dlo : dhi;
        const real imp = impedance_call(P.limit_imp, dist);
        lc = sig * P.limit_B * qdj + P.limit_K * imp * dist;
        lD = sig / r_max(1e-15f, (1.f - imp) * P.limit_invw[jj] / imp);
      }
      sm.jf(j, F_LIMC) = lc; sm.jf(j, F_LIMD) = lD;
    }
    real bl[6];
    root_project_force(RB, fcn, fcl, bl);
#pragma unroll
    for (int k = 0; k < 6; k++) fs_root[k] = -(fs_root[k] + pair_sum(bl[k])) - P.damping[k] * (k < 3 ? rv[k] : rw[k - 3]);
  }
  real rfl_c[3];  // friction-loss row offsets of the root dofs 3*side..3*side+2 owned by this lane
#pragma unroll
  for (int k = 0; k < 3; k++) rfl_c[k] = P.floss_B * (side == 0 ? rv[k] : rw[k]);
  // ---- contact candidates -> compact active list, ordered foot | shin | torso | pelvis (feet, shins about O_s; root about O_r).
  //      Slot 3..5 of a point holds the row residual e = J x + B*velocity at the current iterate x (x = 0 here). ----
  int n_foot = 0, e_shin = 0, e_torso = 0, nact = 0, overflow = 0;
  {
    const real pz = rp[2];
    const real mu2 = mu * mu;
    const int ncand = side == 0 ? 11 : 10;
#pragma unroll 1
    for (int p = 0; p < ncand; p++) {
      V3 lp, xb, omb, vob;
      M3 Rb;
      real rad, href;
      int slot;
      if (p < 4) { lp = ld3(LG.foot_pt[p]); Rb = Rfoot; xb = zero3; omb = om_foot; vob = vo_foot; rad = 0.f; slot = side; href = pz + d.z; }
      else if (p < 6) { lp = ld3(LG.shin_pt[p - 4]); Rb = Rshin; xb = xshin; omb = om_shin; vob = vo_shin; rad = LG.shin_rad; slot = 2 + side; href = pz + d.z; }
      else {
        const int rpi = p < 10 ? 4 * side + (p - 6) : 8;
        lp = ld3(P.root_pt[rpi]); Rb = R0; xb = zero3; omb = om0; vob = v0; rad = P.root_rad[rpi]; slot = p < 10 ? 4 : 5; href = pz;
      }
      const V3 c = xb + mulv(Rb, lp);
      real dist = href + c.z - rad;
      V3 nrm = mk3(0.f, 0.f, 1.f);
      if (ROUGH) {  // point / sphere against the plane of the triangle under it
        real ht, gx, gy;
        terrain_sample(P, terrain_h, te, rp[0] + c.x + (p < 6 ? d.x : 0.f), rp[1] + c.y + (p < 6 ? d.y : 0.f), ht, gx, gy);
        const real s = r_rsqrt(r_fma(gx, gx, r_fma(gy, gy, 1.f)));
        nrm = mk3(-gx * s, -gy * s, s);
        dist = (href + c.z - ht) * s - rad;
      }
      if (dist < 0.f) {
        if (nact < MAXC) {
          const V3 rc = ROUGH ? c - nrm * (rad + 0.5f * dist) : mk3(c.x, c.y, 0.5f * dist - href);  // midway between the surfaces
          V3 vel = vob + cross(omb, rc);
          if (ROUGH) {
            vel = to_contact(cframe(nrm), vel);
            sm.pf(nact, 8) = nrm.x; sm.pf(nact, 9) = nrm.y; sm.pf(nact, 10) = nrm.z;
          }
          const real imp = impedance_call(P.contact_imp, dist);
          const real tr = P.slot_tran[slot];
          const real Rn = r_max(1e-15f, (1.f - imp) * (tr + mu2 * tr) / imp);
          sm.pf(nact, 0) = rc.x; sm.pf(nact, 1) = rc.y; sm.pf(nact, 2) = rc.z;
          sm.pf(nact, 3) = P.contact_B * vel.x; sm.pf(nact, 4) = P.contact_B * vel.y; sm.pf(nact, 5) = P.contact_B * vel.z;
          sm.pf(nact, 6) = P.contact_K * imp * dist;
          sm.pf(nact, 7) = 1.f / r_max(1e-15f, 2.f * mu2 * Rn);
          nact++;
          n_foot += p < 4; e_shin += p < 6; e_torso += p < 10;
        } else {
          overflow = 1;
        }
      }
    }
  }

  // ---- iterate state: per-joint scalars live in the shared-memory column, root 6-vectors in registers ----
  real xr[6], Mar[6], jr[6], rr[6], Msr[6];
  V3 Wl_f = zero3, Wl_s = zero3, F_torso = zero3, F_pelvis = zero3;  // net contact forces per body at the last evaluation
  int it = 0, capped = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) { xr[i] = 0.f; Mar[i] = -fs_root[i]; jr[i] = 0.f; rr[i] = use_warm ? wr[i] : 0.f; }
  const float* arm = P.armature + 6 + j0;
  const float* flD = P.floss_D + 6 + j0;
  const float* flF = P.floss + 6 + j0;
  const V3 c3[3] = {R0.cx, R0.cy, R0.cz};

  // Warp-synchronous loop.  Trip 0 injects the warm start (x = 0 -> previous acceleration, unit step, no solve; a fresh
  // env injects zero).  Every later trip evaluates the rows at x, then solves: a lane in MODE_NEWTON takes a Newton step
  // with exact line search, a lane whose iterate has converged takes the implicitfast update (MODE_FINAL) in the same
  // trip and is MODE_DONE afterwards; done lanes keep running on frozen state (all their writes are selects).
  enum { MODE_NEWTON = 1, MODE_FINAL = 2, MODE_DONE = 3 };
  int mode = MODE_NEWTON;
#ifdef H1V2_WARPCLOCK
  out.trips = 0; out.lstrips = 0;
#endif
#pragma unroll 1
  for (int trip = 0;; trip++) {
#ifdef H1V2_WARPCLOCK
    out.trips++;
#endif
    const bool first = trip == 0;
    real dg_own[3] = {0.f, 0.f, 0.f};  // extra Hessian diagonal of this lane's three root rows
    bool final_trip = false;            // this trip's solve is the lane's implicitfast update
    if (!first) {
      // ---- evaluate all rows at x: forces and gradient ----
      real gr_own[6];
#pragma unroll
      for (int k = 0; k < 6; k++) gr_own[k] = 0.f;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int dk = 3 * side + k;
        const real f = floss_force((side == 0 ? xr[k] : xr[3 + k]) + rfl_c[k], P.floss_D[dk], P.floss[dk], dg_own[k]);
        if (side == 0) gr_own[k] = f; else gr_own[3 + k] = f;
      }
      V3 Wn_f = zero3, Wn_s = zero3, Wn_r = zero3, Wl_r = zero3;
      Wl_f = Wl_s = F_torso = zero3;
      if (QUAD) {  // the sole points (the common case) dealt to the mirrors; shin / torso / pelvis points below by both
#pragma unroll 1
        for (int p = q0; p < n_foot; p += QS) {
          const V3 r = sm.pv(p, 0);
          V3 Fp = point_force(sm.pv(p, 3), sm.pf(p, 6), sm.pf(p, 7), mu);
          if (ROUGH) Fp = from_contact(cframe(sm.pv(p, 8)), Fp);
          Wn_f = Wn_f + cross(r, Fp); Wl_f = Wl_f + Fp;
        }
        Wn_f = mk3(mirror_sum<NM>(Wn_f.x), mirror_sum<NM>(Wn_f.y), mirror_sum<NM>(Wn_f.z));
        Wl_f = mk3(mirror_sum<NM>(Wl_f.x), mirror_sum<NM>(Wl_f.y), mirror_sum<NM>(Wl_f.z));
      }
UNROLL(U_EVP)
      for (int p = (QUAD) ? n_foot : 0; p < nact; p++) {
        const V3 r = sm.pv(p, 0);
        V3 Fp = point_force(sm.pv(p, 3), sm.pf(p, 6), sm.pf(p, 7), mu);
        if (ROUGH) Fp = from_contact(cframe(sm.pv(p, 8)), Fp);  // contact frame -> world
        const V3 mom = cross(r, Fp);
        const real isf = p < n_foot ? 1.f : 0.f, iss = (p >= n_foot && p < e_shin) ? 1.f : 0.f, isr = p >= e_shin ? 1.f : 0.f;
        const real ist = (p >= e_shin && p < e_torso) ? 1.f : 0.f;
        Wn_f = fma3(mom, isf, Wn_f); Wl_f = fma3(Fp, isf, Wl_f);
        Wn_s = fma3(mom, iss, Wn_s); Wl_s = fma3(Fp, iss, Wl_s);
        Wn_r = fma3(mom, isr, Wn_r); Wl_r = fma3(Fp, isr, Wl_r);
        F_torso = fma3(Fp, ist, F_torso);
      }
      F_pelvis = Wl_r - F_torso;
      const V3 Wn_leg = Wn_f + Wn_s, Wl_leg = Wl_f + Wl_s;  // wrench of the leg contacts about O_s
      {
        real t6[6];
        root_project_force(RB, Wn_leg, Wl_leg, t6);
#pragma unroll
        for (int k = 0; k < 6; k++) gr_own[k] += t6[k];
        root_project_force0(c3, Wn_r, Wl_r, t6);
#pragma unroll
        for (int k = 0; k < 6; k++) gr_own[k] += t6[k];
      }
      real gn2 = 0.f, gv = 0.f;
UNROLL(U_EVJ)
      for (int j = q0; j < 6; j += QS) {  // joint rows (friction loss, limit) + J'f of the contacts: gradient g - (M x - f)
        const real x = sm.jf(j, F_XQ);
        real act;
        real f = floss_force(x + sm.jf(j, F_FLC), flD[j], flF[j], act);
        const real lD = sm.jf(j, F_LIMD);
        const real sig = lD > 0.f ? 1.f : (lD < 0.f ? -1.f : 0.f);
        const real jar = r_fma(sig, x, sm.jf(j, F_LIMC));
        const real la = (sig != 0.f && jar < 0.f) ? r_abs(lD) : 0.f;
        f += sig * (-la * jar);
        const V3 wj = sm.jv(j, F_W), uj = sm.jv(j, F_U);
        const real g = f + ((j >= 4) ? dot(wj, Wn_f) + dot(uj, Wl_f) : dot(wj, Wn_leg) + dot(uj, Wl_leg));
        const real r = g - sm.jf(j, F_MA);
        sm.jf(j, F_G) = g; sm.jf(j, F_R) = r; sm.jf(j, F_DG) = act + la;
        gn2 = r_fma(r, r, gn2);
        gv = r_max(gv, r_abs(r) * P.limit_invw[j0 + j]);  // ~ |(M^-1 grad)_j|: what this residual does to the joint's acceleration
      }
      if (QUAD) __syncwarp();  // the mirror's joints are in the shared column
      F_torso = pair_sum(F_torso);
      F_pelvis = pair_sum(F_pelvis);
      gn2 = mirror_sum<NM>(pair_sum(gn2));
      gv = r_max(gv, __shfl_xor_sync(FULL_MASK, gv, 1));
      gv = mirror_max<NM>(gv);
#pragma unroll
      for (int k = 0; k < 6; k++) { jr[k] = pair_sum(gr_own[k]); rr[k] = jr[k] - Mar[k]; gn2 = r_fma(rr[k], rr[k], gn2); }
      if (mode == MODE_NEWTON) {
        const bool conv = r_sqrt(gn2) * P.grad_scale < P.tol && gv < P.vel_tol;
        if (conv || it >= P.max_iters) { capped = !conv; mode = MODE_FINAL; }
        else it++;
      }
      const bool newton = mode == MODE_NEWTON;
      final_trip = mode == MODE_FINAL;
      if (!newton) {
        // implicitfast: (M + h*diag(damping)) qacc = f_smooth + J'f   (a done lane repeats this harmlessly)
#if H1V2_FINAL_MX
        // ... with f_smooth + J'f taken as M x, what it equals at the minimiser (grad = M x - f_smooth - J'f = 0), instead of
        // re-evaluated from the rows.  The two differ by the residual gradient g: the re-evaluated form returns x - M^-1 g, i.e.
        // it hands g (in stiff contact directions mostly fp32 noise of D * J x, or what a dropped negligible Newton step left)
        // to the acceleration through M^-1 -- 1e-3 rad/s per step on a loaded ankle (0.0136 kg m^2) -- while the iterate itself
        // is off by H^-1 g only, H = M + J'DJ >> M exactly where g is large (profiles/r2_notes.md, tools/diag_fp64.py)
#pragma unroll 1
        for (int j = 0; j < 6; j++) { sm.jf(j, F_DG) = h * P.damping[6 + j0 + j]; sm.jf(j, F_R) = sm.jf(j, F_FS) + sm.jf(j, F_MA); }
#pragma unroll
        for (int k = 0; k < 6; k++) rr[k] = fs_root[k] + Mar[k];
#else
#pragma unroll 1
        for (int j = 0; j < 6; j++) { sm.jf(j, F_DG) = h * P.damping[6 + j0 + j]; sm.jf(j, F_R) = sm.jf(j, F_FS) + sm.jf(j, F_G); }
#pragma unroll
        for (int k = 0; k < 6; k++) rr[k] = fs_root[k] + jr[k];
#endif
      }
      // ---- ABA sweep 1 (tip -> root): articulated inertia and reduced rhs ----
      K6 IA;
      k6_zero(IA);
      V3 pn = zero3, pl = zero3;
      const real Dk = newton ? 1.f : 0.f;  // contact stiffness enters the Newton Hessian only
UNROLL(U_SWEEP1)
      for (int j = 5; j >= 0; j--) {
        k6_add_rigid(IA, sm.ji(j, LG.mass[j]));
        if (j == 5 || j == 3) {  // contact stiffness of the foot / shin link, straight into the articulated inertia
          const int p1 = j == 5 ? n_foot : e_shin;
#pragma unroll 1
          for (int p = j == 5 ? 0 : n_foot; p < p1; p++) {
            real Wp[5];
            point_weight(sm.pv(p, 3), sm.pf(p, 6), Dk * sm.pf(p, 7), mu, Wp);
            if (ROUGH) k6_add_point_rot(IA, sm.pv(p, 0), Wp, sm.pv(p, 8));
            else k6_add_point(IA, sm.pv(p, 0), Wp);
          }
        }
        const V3 wj = sm.jv(j, F_W), uj = sm.jv(j, F_U);
        V3 n, l;
        k6_apply(IA, wj, uj, n, l);
        const real dinv = 1.f / (dot(wj, n) + dot(uj, l) + arm[j] + sm.jf(j, F_DG));
        const real t = sm.jf(j, F_R) - (dot(wj, pn) + dot(uj, pl));
        sm.sjv(j, F_X, n); sm.sjv(j, F_X + 3, l);
        sm.jf(j, F_DINV) = dinv; sm.jf(j, F_R) = t;
        k6_rank1_sub(IA, n, l, dinv);
        const real td = t * dinv;
        pn = fma3(n, td, pn); pl = fma3(l, td, pl);
      }
      // ---- root block: both legs' hand-offs moved to the pelvis origin + own link + root contact points; 6x6 Cholesky on both lanes ----
      {
        k6_shift(IA, d);
        pn = pn + cross(d, pl);
        RI Ih = I0;  // half of the root link on each lane: the pair sum below restores it exactly
        Ih.m *= 0.5f; Ih.mc = Ih.mc * 0.5f;
        Ih.xx *= 0.5f; Ih.yy *= 0.5f; Ih.zz *= 0.5f; Ih.xy *= 0.5f; Ih.xz *= 0.5f; Ih.yz *= 0.5f;
        k6_add_rigid(IA, Ih);
#pragma unroll 1
        for (int p = e_shin; p < nact; p++) {
          real Wp[5];
          point_weight(sm.pv(p, 3), sm.pf(p, 6), Dk * sm.pf(p, 7), mu, Wp);
          if (ROUGH) k6_add_point_rot(IA, sm.pv(p, 0), Wp, sm.pv(p, 8));
          else k6_add_point(IA, sm.pv(p, 0), Wp);
        }
#pragma unroll
        for (int i = 0; i < 6; i++) { IA.aa[i] = pair_sum(IA.aa[i]); IA.ll[i] = pair_sum(IA.ll[i]); }
#pragma unroll
        for (int i = 0; i < 9; i++) IA.al[i] = pair_sum(IA.al[i]);
        pn = pair_sum(pn); pl = pair_sum(pl);
        real A[21], g[6];
#pragma unroll
        for (int i = 0; i < 21; i++) A[i] = 0.f;
        root_project_k60(c3, IA, A);
        root_project_force0(c3, pn, pl, g);
        // diagonal: armature, friction-loss curvature of the root rows (Newton) or h*damping (implicit update)
        const real o0 = newton ? dg_own[0] : 0.f, o1 = newton ? dg_own[1] : 0.f, o2 = newton ? dg_own[2] : 0.f;
        const real e0 = pair_sum(side == 0 ? o0 : 0.f), e1 = pair_sum(side == 0 ? o1 : 0.f), e2 = pair_sum(side == 0 ? o2 : 0.f);
        const real e3 = pair_sum(side == 0 ? 0.f : o0), e4 = pair_sum(side == 0 ? 0.f : o1), e5 = pair_sum(side == 0 ? 0.f : o2);
        const real ex[6] = {e0, e1, e2, e3, e4, e5};
#pragma unroll
        for (int k = 0; k < 6; k++) {
          A[TI(k, k)] += P.armature[k] + (newton ? ex[k] : h * P.damping[k]);
          rr[k] -= g[k];
        }
        real inva[6];
        chol6(A, inva);
        fwd6(A, inva, rr);
        bwd6(A, inva, rr);
      }
    }
    // ---- sweep 2 (root -> tip): joint accelerations; the same sweep starts the M-product of the direction and
    //      leaves the direction's body accelerations (root about O_r; shin, foot about O_s) for the line search ----
    V3 Sa_r, Sl_r, Sa_s, Sl_s, Sa_f, Sl_f;
    Sa_r = fma3(c3[0], rr[3], fma3(c3[1], rr[4], c3[2] * rr[5]));
    Sl_r = mk3(rr[0], rr[1], rr[2]);
    Sa_f = Sa_r; Sl_f = Sl_r + cross(Sa_r, d);
    Sa_s = Sa_f; Sl_s = Sl_f;
    real smax = 0.f;
UNROLL(U_SWEEP2)
    for (int j = 0; j < 6; j++) {
      real s = sm.jf(j, F_R);
      if (!first) {
        s = (s - (dot(sm.jv(j, F_X), Sa_f) + dot(sm.jv(j, F_X + 3), Sl_f))) * sm.jf(j, F_DINV);
        sm.jf(j, F_R) = s;
      }
      smax = r_max(smax, r_abs(s));
      Sa_f = fma3(sm.jv(j, F_W), s, Sa_f); Sl_f = fma3(sm.jv(j, F_U), s, Sl_f);
      if (j == 3) { Sa_s = Sa_f; Sl_s = Sl_f; }
      V3 n, l;
      ri_apply(sm.ji(j, LG.mass[j]), Sa_f, Sl_f, n, l);
      sm.sjv(j, F_X, n); sm.sjv(j, F_X + 3, l);
    }
    smax = r_max(smax, __shfl_xor_sync(FULL_MASK, smax, 1));
#pragma unroll
    for (int i = 0; i < 6; i++) smax = r_max(smax, r_abs(rr[i]));
    // what this lane does with the direction: Newton step with line search | unit step (trip 0) | nothing.
    // A Newton step that no longer moves the acceleration is dropped and the iterate accepted (J'f is current).
    bool search = !first && mode == MODE_NEWTON;
    if (search && smax < P.step_tol) { search = false; mode = MODE_FINAL; }
    // ---- M-product, tip -> root half:  Ms = M s  (stored in the F_DG slot) ; line-search scalars ----
    real sMs = 0.f, sMa = 0.f, gs = 0.f;
    {
      V3 fcn = zero3, fcl = zero3;
UNROLL(U_MPROD)
      for (int j = 5; j >= 0; j--) {
        fcn = fcn + sm.jv(j, F_X); fcl = fcl + sm.jv(j, F_X + 3);
        const real s = sm.jf(j, F_R);
        const real ms = dot(sm.jv(j, F_W), fcn) + dot(sm.jv(j, F_U), fcl) + arm[j] * s;
        sm.jf(j, F_DG) = ms;
        sMs = r_fma(s, ms, sMs); sMa = r_fma(s, sm.jf(j, F_MA), sMa); gs = r_fma(s, sm.jf(j, F_G), gs);
      }
      real bl[6], b0[6];
      root_project_force(RB, fcn, fcl, bl);
      V3 n0, f0;
      ri_apply(I0, Sa_r, Sl_r, n0, f0);
      root_project_force0(c3, n0, f0, b0);
#pragma unroll
      for (int k = 0; k < 6; k++) Msr[k] = b0[k] + pair_sum(bl[k]) + P.armature[k] * rr[k];
    }
    real alpha = first ? 1.f : 0.f;
    {
      // ---- exact line search along the direction (all lanes run the trips; only searching lanes move alpha) ----
      sMs = pair_sum(sMs); sMa = pair_sum(sMa); gs = pair_sum(gs);
#pragma unroll
      for (int k = 0; k < 6; k++) { sMs = r_fma(rr[k], Msr[k], sMs); sMa = r_fma(rr[k], Mar[k], sMa); gs = r_fma(rr[k], jr[k], gs); }
      const real d10 = sMa - gs;  // phi'(0) = grad . search  (< 0)
      real lo = 0.f, hi = 1e30f, dlo = d10, dhi = 0.f;  // bracket of the minimiser and phi' at its ends
      if (search) alpha = 1.f;
#pragma unroll 1
      for (int ls = 0; __any_sync(FULL_MASK, search); ls++) {
#ifdef H1V2_WARPCLOCK
        out.lstrips++;
#endif
        real d1 = 0.f, d2 = 0.f;
UNROLL(U_LSJ)
        for (int j = q0; j < 6; j += QS) {
          const real s = sm.jf(j, F_R);
          const real xa = r_fma(alpha, s, sm.jf(j, F_XQ));
          real act;
          const real f = floss_force(xa + sm.jf(j, F_FLC), flD[j], flF[j], act);
          d1 = r_fma(-f, s, d1); d2 = r_fma(act * s, s, d2);
          const real lD = sm.jf(j, F_LIMD);  // limit row, branch-free: lD = 0 (no row) or an inactive row weigh nothing
          const real sig = lD > 0.f ? 1.f : (lD < 0.f ? -1.f : 0.f);
          const real jar = r_fma(sig, xa, sm.jf(j, F_LIMC));
          const real lw = jar < 0.f ? r_abs(lD) : 0.f;
          d1 = r_fma(lw * jar, sig * s, d1); d2 = r_fma(lw * s, s, d2);
        }
        if (!QUAD || q0 == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const int dk = 3 * side + k;
          const real sk = side == 0 ? rr[k] : rr[3 + k], xk = side == 0 ? xr[k] : xr[3 + k];
          real act;
          const real f = floss_force(r_fma(alpha, sk, xk) + rfl_c[k], P.floss_D[dk], P.floss[dk], act);
          d1 = r_fma(-f, sk, d1); d2 = r_fma(act * sk, sk, d2);
        }
        }
UNROLL(U_LSP)
        for (int p = q0; p < nact; p += QS) {
          const V3 r = sm.pv(p, 0);
          const bool isf = p < n_foot, iss = p < e_shin;
          const V3 Sa = isf ? Sa_f : (iss ? Sa_s : Sa_r), Sl = isf ? Sl_f : (iss ? Sl_s : Sl_r);
          V3 us = Sl + cross(Sa, r);
          if (ROUGH) us = to_contact(cframe(sm.pv(p, 8)), us);
          point_ls(fma3(us, alpha, sm.pv(p, 3)), us, sm.pf(p, 6), sm.pf(p, 7), mu, d1, d2);
        }
        d1 = mirror_sum<NM>(pair_sum(d1)) + r_fma(alpha, sMs, sMa);
        d2 = mirror_sum<NM>(pair_sum(d2)) + sMs;
        if (search) {
          if (r_abs(d1) <= P.ls_tol * r_abs(d10) || !(d2 > 0.f)) search = false;
          else {
            if (d1 < 0.f) { lo = alpha; dlo = d1; } else { hi = alpha; dhi = d1; }
            if (ls + 1 >= P.ls_max) {  // out of trips: phi' is monotone, take the secant root inside the bracket (regula falsi)
              if (hi < 1e29f) alpha = lo + (hi - lo) * (-dlo / (dhi - dlo));
              search = false;
            } else {
              real nx = alpha - d1 / d2;
              if (!(nx > lo && nx < hi)) nx = hi < 1e29f ? 0.5f * (lo + hi) : 2.f * alpha;
              if (r_abs(nx - alpha) <= 1e-4f * alpha) search = false;
              alpha = nx;
            }
          }
        }
      }
    }
    // ---- take the step: x += alpha s ; M x - f += alpha M s ; row residuals of the contact points += alpha J s.
    //      A lane that just took its implicit update keeps the result: qacc -> F_FS, root part -> wr. ----
    const bool move = alpha != 0.f;
#pragma unroll 1
    for (int j = q0; j < 6; j += QS) {
      const real s = sm.jf(j, F_R);
      if (move) {
        sm.jf(j, F_XQ) = r_fma(alpha, s, sm.jf(j, F_XQ));
        sm.jf(j, F_MA) = r_fma(alpha, sm.jf(j, F_DG), sm.jf(j, F_MA));
      }
      if (final_trip) sm.jf(j, F_FS) = s;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
      if (move) { xr[i] = r_fma(alpha, rr[i], xr[i]); Mar[i] = r_fma(alpha, Msr[i], Mar[i]); }
      if (final_trip) wr[i] = rr[i];
    }
    if (move) {
UNROLL(U_STP)
      for (int p = q0; p < nact; p += QS) {
        const V3 r = sm.pv(p, 0);
        const bool isf = p < n_foot, iss = p < e_shin;
        const V3 Sa = isf ? Sa_f : (iss ? Sa_s : Sa_r), Sl = isf ? Sl_f : (iss ? Sl_s : Sl_r);
        V3 us = Sl + cross(Sa, r);
        if (ROUGH) us = to_contact(cframe(sm.pv(p, 8)), us);
        const V3 e = fma3(us, alpha, sm.pv(p, 3));
        sm.pf(p, 3) = e.x; sm.pf(p, 4) = e.y; sm.pf(p, 5) = e.z;
      }
    }
    if (final_trip) mode = MODE_DONE;
    if (QUAD) __syncwarp();  // both mirrors' halves of the step are in the shared column
    if (__all_sync(FULL_MASK, mode == MODE_DONE)) break;
  }
  // ---- integrate (semi-implicit Euler; quaternion on SO(3) with the body-frame angular velocity) ----
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const real a = sm.jf(j, F_FS);
    out.qacc[j] = a; wl[j] = a;
    qd[j] = r_fma(h, a, qd[j]);
    if (P.vel_limit > 0.f) qd[j] = r_min(r_max(qd[j], -P.vel_limit), P.vel_limit);  // actuator velocity_limit (A/robots/h12.py:66)
    q[j] = r_fma(h, qd[j], q[j]);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) { rv[k] = r_fma(h, wr[k], rv[k]); rw[k] = r_fma(h, wr[3 + k], rw[k]); rp[k] = r_fma(h, rv[k], rp[k]); }
  {
    real wn = r_sqrt(rw[0] * rw[0] + rw[1] * rw[1] + rw[2] * rw[2]);
    real ang = wn * h, dw = 1.f, dx = 0.f, dy = 0.f, dz = 0.f;
    if (ang > 0.f) {
      real s, c;
      sincos_lim(0.5f * ang, s, c);
      s /= wn;
      dw = c; dx = rw[0] * s; dy = rw[1] * s; dz = rw[2] * s;
    }
    real a = rq[0], b = rq[1], c = rq[2], e = rq[3];
    real nw = a * dw - b * dx - c * dy - e * dz, nx = a * dx + b * dw + c * dz - e * dy;
    real ny = a * dy - b * dz + c * dw + e * dx, nz = a * dz + b * dy - c * dx + e * dw;
    real n = r_rsqrt(nw * nw + nx * nx + ny * ny + nz * nz);
    rq[0] = nw * n; rq[1] = nx * n; rq[2] = ny * n; rq[3] = nz * n;
  }
  out.F_foot = Wl_f; out.F_shin = Wl_s; out.F_torso = F_torso; out.F_pelvis = F_pelvis;
  out.iters = it; out.capped = capped; out.overflow = overflow;
}

}  // namespace h1v2
