// h1v2_terrain.cuh -- height-field terrain of the Rough id: lookup, contact frame, terrain-level curriculum.
//
// What it restates (upstream isaaclab 2.1.0 as driven by the reference's cfgs): a grid of rows x cols square tiles
// (terrain_generator.py, curriculum layout: row = level, column = terrain type), every tile a height field meshed by
// convert_height_field_to_mesh -- each grid cell split along the diagonal (i, j) -> (i + 1, j + 1) -- the generator cfg is
// packages/biped_tasks/biped_tasks/utils/mdp/terrains.py:11-28, the importer V/velocity_env_cfg.py:40-58.  Positions are kept
// RELATIVE TO THE ENV'S TILE ORIGIN (tile centre, height of the highest vertex of its central 2 m patch), which is what
// env.scene.env_origins is upstream: a lookup works in tile-local coordinates (|x| < ~10 m, full fp32 resolution) and
// the curriculum's walked distance is |root_pos_xy|.
#pragma once
#include "h1v2_math.cuh"
#include "h1v2_params.h"

namespace h1v2 {

struct TerrainEnv {
  int i0, j0;  // grid vertex of the tile's corner
  real oz;     // height of the tile origin
};
__device__ __forceinline__ TerrainEnv terrain_env(const KParams& P, const float* __restrict__ origin_z, int flags) {
  const int level = (flags >> FLAG_LEVEL_SHIFT) & 255, type = (flags >> FLAG_TYPE_SHIFT) & 255;
  TerrainEnv t;
  t.i0 = level * P.t_npx; t.j0 = type * P.t_npx;
  t.oz = __ldg(origin_z + level * P.t_cols + type);
  return t;
}
// height (relative to the tile origin) and gradient of the triangle under the point (lx, ly), given relative to the tile origin.
// Outside the grid lies the flat border at world height 0 (terrain_generator.py _add_terrain_border).
__device__ __forceinline__ void terrain_sample(const KParams& P, const float* __restrict__ H, const TerrainEnv& te, real lx, real ly, real& h, real& gx, real& gy) {
  const real a = (lx + P.t_half) * P.t_inv_hs, b = (ly + P.t_half) * P.t_inv_hs;
  const real fa = r_floor(a), fb = r_floor(b);
  const int i = te.i0 + (int)fa, j = te.j0 + (int)fb;
  h = -te.oz; gx = 0.f; gy = 0.f;
  if (i >= 0 && j >= 0 && i < P.t_gx - 1 && j < P.t_gy - 1) {
    const float* p = H + (size_t)i * P.t_gy + j;
    const real h00 = __ldg(p), h01 = __ldg(p + 1), h10 = __ldg(p + P.t_gy), h11 = __ldg(p + P.t_gy + 1);
    const real u = a - fa, v = b - fb;
    const bool upper = v >= u;  // triangle (i,j) (i+1,j+1) (i,j+1) | (i,j) (i+1,j) (i+1,j+1)
    const real dx = upper ? h11 - h01 : h10 - h00, dy = upper ? h01 - h00 : h11 - h10;
    h = r_fma(u, dx, r_fma(v, dy, h00)) - te.oz;
    gx = dx * P.t_inv_hs; gy = dy * P.t_inv_hs;
  }
}
// Contact frame of MuJoCo's mju_makeFrame in the axis order the plane code uses (local x, y, z = -t2, t1, n; on the plane x, y, z):
// t1 = (e - n d) / sqrt(1 - d^2) with e the world y axis (z when the normal is within 30 degrees of y) and d = n . e, t2 = n x t1, so
// -t2 = t1 x n = (e x n) / sqrt(1 - d^2).  Only the normal is stored per contact point; the closed forms below rebuild what each
// use needs (a dot product with n, one component of n x v, two scalings) instead of the axes themselves.
struct CFrame {
  V3 n;
  real d, inv;
  bool usey;
};
__device__ __forceinline__ CFrame cframe(V3 n) {
  CFrame f;
  f.n = n; f.usey = r_abs(n.y) < 0.5f;
  f.d = f.usey ? n.y : n.z;
  f.inv = r_rsqrt(r_fma(-f.d, f.d, 1.f));
  return f;
}
__device__ __forceinline__ V3 to_contact(const CFrame& f, V3 v) {  // (ax . v, ay . v, n . v)
  const real nv = dot(f.n, v);
  const real ev = f.usey ? v.y : v.z;
  const real cx = f.usey ? r_fma(f.n.z, v.x, -f.n.x * v.z) : r_fma(f.n.x, v.y, -f.n.y * v.x);  // (e x n) . v
  return mk3(cx * f.inv, r_fma(-f.d, nv, ev) * f.inv, nv);
}
__device__ __forceinline__ V3 from_contact(const CFrame& f, V3 w) {  // ax w.x + ay w.y + n w.z
  const real a = w.x * f.inv, b = w.y * f.inv;
  V3 r = f.n * r_fma(-b, f.d, w.z);
  r.x += f.usey ? a * f.n.z : -a * f.n.y;
  r.y += f.usey ? b : a * f.n.x;
  r.z += f.usey ? -a * f.n.x : b;
  return r;
}
__device__ __forceinline__ void contact_axes(const CFrame& f, V3& ax, V3& ay) {
  ax = f.usey ? mk3(f.n.z * f.inv, 0.f, -f.n.x * f.inv) : mk3(-f.n.y * f.inv, f.n.x * f.inv, 0.f);
  const V3 t = mk3(0.f, f.usey ? 1.f : 0.f, f.usey ? 0.f : 1.f) - f.n * f.d;
  ay = t * f.inv;
}

// mdp.terrain_levels_vel (V/mdp/curriculums.py:21-52) + TerrainImporter.update_env_origins: on reset, an env that walked further than
// half a tile moves one level up, one that covered less than half the distance its command asked for moves one down; past the last
// level it is sent to a random one.  IEEE fp32 operations in the order torch evaluates them (the oracle does the same, bit for bit).
__device__ __forceinline__ int terrain_curriculum(const KParams& P, int flags, float x, float y, float cx, float cy, int64_t gid, unsigned long long step) {
  if (!P.t_curriculum) return flags;
  const float dist = __fsqrt_rn(__fmaf_rn(y, y, __fmul_rn(x, x)));
  const float cn = __fsqrt_rn(__fmaf_rn(cy, cy, __fmul_rn(cx, cx)));
  const bool up = dist > __fmul_rn(P.t_tile, 0.5f);
  const bool down = dist < __fmul_rn(__fmul_rn(cn, P.max_episode_length_s), 0.5f) && !up;
  int lev = ((flags >> FLAG_LEVEL_SHIFT) & 255) + (up ? 1 : 0) - (down ? 1 : 0);
  if (lev >= P.t_rows) {
    float u[4];
    rng4(P.key0, gid, step, STREAM_RESET, 10, u);
    lev = min((int)__fmul_rn(u[0], (float)P.t_rows), P.t_rows - 1);
  } else if (lev < 0) lev = 0;
  return (flags & ~(255 << FLAG_LEVEL_SHIFT)) | (lev << FLAG_LEVEL_SHIFT);
}

}  // namespace h1v2
