// h1v2_math.cuh -- small fixed-size math used by the step kernel (device only, everything inlined/unrolled).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace h1v2 {

// ---- arithmetic type of the step kernel.  float is the product; -DH1V2_FP64=1 builds the SAME kernel source in double
//      (libh1v2_b200_f64.so, test infrastructure: tests/test_gpu_fp64.py) to separate rounding from algorithm when the
//      fp32 kernel is compared with the float64 oracle.  HBM state, parameters, RNG and observations stay float in both.
#ifndef H1V2_FP64
#define H1V2_FP64 0
#endif
#if H1V2_FP64
typedef double real;
__device__ __forceinline__ real r_fma(real a, real b, real c) { return fma(a, b, c); }
__device__ __forceinline__ real r_min(real a, real b) { return fmin(a, b); }
__device__ __forceinline__ real r_max(real a, real b) { return fmax(a, b); }
__device__ __forceinline__ real r_abs(real a) { return fabs(a); }
__device__ __forceinline__ real r_sqrt(real a) { return sqrt(a); }
__device__ __forceinline__ real r_rsqrt(real a) { return 1.0 / sqrt(a); }
__device__ __forceinline__ real r_pow(real a, real b) { return pow(a, b); }
__device__ __forceinline__ real r_rint(real a) { return rint(a); }
__device__ __forceinline__ real r_floor(real a) { return floor(a); }
__device__ __forceinline__ real r_exp(real a) { return exp(a); }
__device__ __forceinline__ real r_atan2(real a, real b) { return atan2(a, b); }
#else
typedef float real;
__device__ __forceinline__ real r_fma(real a, real b, real c) { return fmaf(a, b, c); }
__device__ __forceinline__ real r_min(real a, real b) { return fminf(a, b); }
__device__ __forceinline__ real r_max(real a, real b) { return fmaxf(a, b); }
__device__ __forceinline__ real r_abs(real a) { return fabsf(a); }
__device__ __forceinline__ real r_sqrt(real a) { return sqrtf(a); }
__device__ __forceinline__ real r_rsqrt(real a) { return rsqrtf(a); }
__device__ __forceinline__ real r_pow(real a, real b) { return powf(a, b); }
__device__ __forceinline__ real r_rint(real a) { return rintf(a); }
__device__ __forceinline__ real r_floor(real a) { return floorf(a); }
__device__ __forceinline__ real r_exp(real a) { return expf(a); }
__device__ __forceinline__ real r_atan2(real a, real b) { return atan2f(a, b); }
#endif

struct V3 {
  real x, y, z;
};
__device__ __forceinline__ V3 mk3(real x, real y, real z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, real s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(real s, V3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ real dot(V3 a, V3 b) { return r_fma(a.x, b.x, r_fma(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk3(r_fma(a.y, b.z, -a.z * b.y), r_fma(a.z, b.x, -a.x * b.z), r_fma(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ V3 fma3(V3 a, real s, V3 b) { return mk3(r_fma(a.x, s, b.x), r_fma(a.y, s, b.y), r_fma(a.z, s, b.z)); }
__device__ __forceinline__ V3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ real comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

struct M3 {  // columns
  V3 cx, cy, cz;
};
__device__ __forceinline__ V3 mulv(const M3& R, V3 v) { return fma3(R.cx, v.x, fma3(R.cy, v.y, R.cz * v.z)); }
__device__ __forceinline__ V3 mulTv(const M3& R, V3 v) { return mk3(dot(R.cx, v), dot(R.cy, v), dot(R.cz, v)); }
__device__ __forceinline__ M3 quat2mat(real w, real x, real y, real z) {
  M3 R;
  R.cx = mk3(1.f - 2.f * (y * y + z * z), 2.f * (x * y + w * z), 2.f * (x * z - w * y));
  R.cy = mk3(2.f * (x * y - w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + w * x));
  R.cz = mk3(2.f * (x * z + w * y), 2.f * (y * z - w * x), 1.f - 2.f * (x * x + y * y));
  return R;
}
// sine/cosine for bounded arguments (joint angles, half rotation angles: |x| < ~100): Cody-Waite reduction by pi/2
// and degree-7/8 minimax polynomials on [-pi/4, pi/4]; ~1 ulp, no slow path (keeps the kernel's code small)
__device__ __forceinline__ void sincos_lim(real x, real& s, real& c) {
#if H1V2_FP64
  sincos(x, &s, &c);
  return;
#else
  const real kf = r_rint(x * 0.636619772367581343f);
  const int k = (int)kf;
  real r = r_fma(kf, -1.57079601287841796875f, x);
  r = r_fma(kf, -3.1391647326017846353e-7f, r);
  r = r_fma(kf, -5.3903025299577647655e-15f, r);
  const real r2 = r * r;
  real sp = r_fma(r2, -1.95152959e-4f, 8.33216087e-3f);
  sp = r_fma(sp, r2, -1.66666546e-1f);
  sp = r_fma(sp * r2, r, r);
  real cp = r_fma(r2, 2.44331571e-5f, -1.38873163e-3f);
  cp = r_fma(cp, r2, 4.16666457e-2f);
  cp = r_fma(cp, r2, -0.5f);
  cp = r_fma(cp, r2, 1.0f);
  const real ss = (k & 1) ? cp : sp, cc = (k & 1) ? sp : cp;
  s = (k & 2) ? -ss : ss;
  c = ((k + 1) & 2) ? -cc : cc;
#endif
}
// R <- R * Rot(axis, th), axis in {0:x, 1:y, 2:z}
template <int AXIS>
__device__ __forceinline__ void rotate(M3& R, real th) {
  real s, c;
  sincos_lim(th, s, c);
  if (AXIS == 2) {
    V3 X = fma3(R.cx, c, R.cy * s), Y = fma3(R.cy, c, R.cx * (-s));
    R.cx = X; R.cy = Y;
  } else if (AXIS == 1) {
    V3 X = fma3(R.cx, c, R.cz * (-s)), Z = fma3(R.cz, c, R.cx * s);
    R.cx = X; R.cz = Z;
  } else {
    V3 Y = fma3(R.cy, c, R.cz * s), Z = fma3(R.cz, c, R.cy * (-s));
    R.cy = Y; R.cz = Z;
  }
}
// same with the sine/cosine already known
template <int AXIS>
__device__ __forceinline__ void rotate_sc(M3& R, real s, real c) {
  if (AXIS == 2) {
    V3 X = fma3(R.cx, c, R.cy * s), Y = fma3(R.cy, c, R.cx * (-s));
    R.cx = X; R.cy = Y;
  } else if (AXIS == 1) {
    V3 X = fma3(R.cx, c, R.cz * (-s)), Z = fma3(R.cz, c, R.cx * s);
    R.cx = X; R.cz = Z;
  } else {
    V3 Y = fma3(R.cy, c, R.cz * s), Z = fma3(R.cz, c, R.cy * (-s));
    R.cy = Y; R.cz = Z;
  }
}
template <int AXIS>
__device__ __forceinline__ V3 axis_col(const M3& R) { return AXIS == 0 ? R.cx : (AXIS == 1 ? R.cy : R.cz); }

// rigid spatial inertia about the reference point O, world axes: mass, mass*com, rotational inertia about O
struct RI {
  real m;
  V3 mc;
  real xx, yy, zz, xy, xz, yz;
};
__device__ __forceinline__ RI operator+(const RI& a, const RI& b) {
  RI r;
  r.m = a.m + b.m; r.mc = a.mc + b.mc;
  r.xx = a.xx + b.xx; r.yy = a.yy + b.yy; r.zz = a.zz + b.zz;
  r.xy = a.xy + b.xy; r.xz = a.xz + b.xz; r.yz = a.yz + b.yz;
  return r;
}
__device__ __forceinline__ V3 rot_inertia_mul(const RI& I, V3 w) {
  return mk3(r_fma(I.xx, w.x, r_fma(I.xy, w.y, I.xz * w.z)), r_fma(I.xy, w.x, r_fma(I.yy, w.y, I.yz * w.z)),
             r_fma(I.xz, w.x, r_fma(I.yz, w.y, I.zz * w.z)));
}
// (n,l) = I * (w,u)
__device__ __forceinline__ void ri_apply(const RI& I, V3 w, V3 u, V3& n, V3& l) {
  n = rot_inertia_mul(I, w) + cross(I.mc, u);
  l = fma3(u, I.m, cross(w, I.mc));
}
// world inertia of a body: rotation R, COM c (relative to O), mass m, body-frame inertia ib = xx yy zz xy xz yz
__device__ __forceinline__ RI body_inertia(const M3& R, V3 c, real m, const float* ib, real iscale) {
  V3 t0 = mulv(R, mk3(ib[0], ib[3], ib[4])) * iscale;
  V3 t1 = mulv(R, mk3(ib[3], ib[1], ib[5])) * iscale;
  V3 t2 = mulv(R, mk3(ib[4], ib[5], ib[2])) * iscale;
  RI I;
  I.m = m;
  I.mc = c * m;
  real cc = dot(c, c);
  // Iw(i,j) = sum_k t_k[i] * Rcol_k[j]
  I.xx = r_fma(t0.x, R.cx.x, r_fma(t1.x, R.cy.x, t2.x * R.cz.x)) + m * (cc - c.x * c.x);
  I.yy = r_fma(t0.y, R.cx.y, r_fma(t1.y, R.cy.y, t2.y * R.cz.y)) + m * (cc - c.y * c.y);
  I.zz = r_fma(t0.z, R.cx.z, r_fma(t1.z, R.cy.z, t2.z * R.cz.z)) + m * (cc - c.z * c.z);
  I.xy = r_fma(t0.x, R.cx.y, r_fma(t1.x, R.cy.y, t2.x * R.cz.y)) - m * c.x * c.y;
  I.xz = r_fma(t0.x, R.cx.z, r_fma(t1.x, R.cy.z, t2.x * R.cz.z)) - m * c.x * c.z;
  I.yz = r_fma(t0.y, R.cx.z, r_fma(t1.y, R.cy.z, t2.y * R.cz.z)) - m * c.y * c.z;
  return I;
}

// packed lower-triangular index
#define TI(i, j) ((i) * ((i) + 1) / 2 + (j))

// in-place Cholesky of a packed symmetric 6x6 (lower); invd = 1/diag(L)
__device__ __forceinline__ void chol6(real (&A)[21], real (&invd)[6]) {
#pragma unroll
  for (int j = 0; j < 6; j++) {
    real d = A[TI(j, j)];
#pragma unroll
    for (int k = 0; k < j; k++) d = r_fma(-A[TI(j, k)], A[TI(j, k)], d);
    d = r_max(d, 1e-12f);
    real r = r_rsqrt(d);
    invd[j] = r;
    A[TI(j, j)] = d * r;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      real s = A[TI(i, j)];
#pragma unroll
      for (int k = 0; k < j; k++) s = r_fma(-A[TI(i, k)], A[TI(j, k)], s);
      A[TI(i, j)] = s * r;
    }
  }
}
__device__ __forceinline__ void fwd6(const real (&L)[21], const real (&invd)[6], real (&b)[6]) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    real s = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) s = r_fma(-L[TI(i, k)], b[k], s);
    b[i] = s * invd[i];
  }
}
__device__ __forceinline__ void bwd6(const real (&L)[21], const real (&invd)[6], real (&b)[6]) {
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    real s = b[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s = r_fma(-L[TI(k, i)], b[k], s);
    b[i] = s * invd[i];
  }
}

// ---------------- Philox4x32-10, same stream contract as the oracle (SURVEY 8(a) RNG note) ----------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
#define STREAM_OBS 0u
#define STREAM_RESET 1u
#define STREAM_CMD 2u
#define STREAM_EVENT 3u
#define STREAM_ACTIONS 7u
// out-of-line (one copy of the 10 Philox rounds), result returned by value so that no caller array has its address taken
__device__ __noinline__ float4 rng4v(uint32_t key0, uint32_t gid, uint32_t step_lo, uint32_t step_hi, uint32_t stream, uint32_t block) {
  uint32_t o[4];
  philox4x32_10(step_lo, step_hi, stream, block, key0, gid, o);
  const float k = 1.0f / 16777216.0f;
  return make_float4((float)(o[0] >> 8) * k, (float)(o[1] >> 8) * k, (float)(o[2] >> 8) * k, (float)(o[3] >> 8) * k);
}
__device__ __forceinline__ void rng4(uint32_t key0, int64_t gid, unsigned long long step, uint32_t stream, uint32_t block,
                                     float (&u)[4]) {
  const float4 r = rng4v(key0, (uint32_t)gid, (uint32_t)step, (uint32_t)(step >> 32), stream, block);
  u[0] = r.x; u[1] = r.y; u[2] = r.z; u[3] = r.w;
}
// one rounding (fused multiply-add) on both sides: the oracle calls fmaf too, so every uniform draw is bit-identical to it
__device__ __forceinline__ float uni(float u, float lo, float hi) { return fmaf(hi - lo, u, lo); }

__device__ __forceinline__ real wrap_to_pi(real a) {
  const real PI = 3.14159265358979323846f, TWO_PI = 2.0f * 3.14159265358979323846f;
  real w = a + PI;
  w = w - TWO_PI * r_floor(w / TWO_PI);
  if (w < 0.f) w += TWO_PI;
  if (w >= TWO_PI) w -= TWO_PI;
  if (w == 0.0f && a > 0.0f) return PI;
  return w - PI;
}

}  // namespace h1v2
