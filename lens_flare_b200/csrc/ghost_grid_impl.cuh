// ghost_grid_impl.cuh -- per-ray machinery of the ray-grid ghost path, templated on the
// arithmetic type.  Included by exactly two translation units:
//   ghost_grid_f32.cu  (R = float,  FMA contraction ON : the throughput kernels)
//   ghost_grid_f64.cu  (R = double, --fmad=false       : the parity kernels, whose operation
//                       order is that of oracle/lf_oracle.c so that sums agree bit for bit)
//
//   trace_splat_kernel   per ray: PARAXIAL_GRID (the reference's ABCD matrices,
//                        pathtracer.cpp:588-689, applied per ray and per axis) or EXACT_GRID
//                        (sphere/plane hit, vector Snell, Fresnel / quarter-wave coating),
//                        aperture-mask lookup at each stop crossing, fixed-point splat
//   trace_dump_kernel    same trace, writes lfb_ray_hit records (parity instrument)
//
// Work layout: one CTA = a 16x16 patch of one ghost's N x N ray grid, so every lane of a CTA
// walks the same surface sequence (no divergence in the surface loop) and neighbouring lanes
// land on neighbouring sensor pixels.  The prescription sits in __constant__ memory (all
// indices are warp-uniform -> constant-cache broadcast); the 1 MB aperture mask stays
// L2-resident.  The work is scalar FP32 ALU/MUFU math: tensor cores / TMA do not apply.
#pragma once
#include <math_constants.h>

#include "lfb_internal.h"

namespace lfb {
namespace LFB_TU {

__constant__ DevLens c_lens;

// ---------------------------------------------------------------------------
// per-ray machinery, templated on the arithmetic type
// ---------------------------------------------------------------------------
template <typename R> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float cos(float x) { return cosf(x); }
  static __device__ __forceinline__ float floor(float x) { return floorf(x); }
  static __device__ __forceinline__ float zv(int k) { return c_lens.zv[k]; }
  static __device__ __forceinline__ float nan() { return CUDART_NAN_F; }
  static __device__ __forceinline__ long long to_fixed(float v, double scale) {
    return __float2ll_rn(v * (float)scale);
  }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double cos(double x) { return ::cos(x); }
  static __device__ __forceinline__ double floor(double x) { return ::floor(x); }
  static __device__ __forceinline__ double zv(int k) { return c_lens.zv_d[k]; }
  static __device__ __forceinline__ double nan() { return CUDART_NAN; }
  static __device__ __forceinline__ long long to_fixed(double v, double scale) {
    return __double2ll_rn(v * scale);
  }
};

template <typename R>
struct Ray {
  R ox, oy, oz, dx, dy, dz, w, xa, ya;
  unsigned flags;
};

template <typename R>
__device__ __forceinline__ R mask_lookup(const float* __restrict__ tex, int tw, int th, R xa, R ya) {
  const R h = (R)c_lens.h_stop;
  R u = (xa / h * (R)0.5 + (R)0.5) * (R)tw;
  R v = ((R)0.5 - ya / h * (R)0.5) * (R)th;
  R fu = Num<R>::floor(u), fv = Num<R>::floor(v);
  if (!(fu >= (R)0 && fu < (R)tw && fv >= (R)0 && fv < (R)th)) return (R)0;
  return (R)__ldg(tex + (int)fv * tw + (int)fu);
}

// Single-surface reflectance: exact single-layer (Airy) film of index max(sqrt(n0 n2), 1.38)
// and quarter-wave thickness at lambda0, or bare Fresnel when lambda0 == 0.
template <typename R>
__device__ __forceinline__ R reflectance(R n0, R n2, R cos0, R cos2, R lambda0, R lambda) {
  if (lambda0 <= (R)0) {
    R rs = (n0 * cos0 - n2 * cos2) / (n0 * cos0 + n2 * cos2);
    R rp = (n2 * cos0 - n0 * cos2) / (n2 * cos0 + n0 * cos2);
    return (R)0.5 * (rs * rs + rp * rp);
  }
  R n1 = Num<R>::sqrt(n0 * n2);
  if (n1 < (R)1.38) n1 = (R)1.38;
  R d1 = lambda0 / ((R)4 * n1);
  R e1 = n0 / n1, k1 = (R)1 - e1 * e1 * ((R)1 - cos0 * cos0);
  if (k1 < (R)0) return (R)1;
  R cos1 = Num<R>::sqrt(k1);
  R cd = Num<R>::cos((R)(4.0 * 3.14159265358979323846) * n1 * d1 * cos1 / lambda);
  R r01s = (n0 * cos0 - n1 * cos1) / (n0 * cos0 + n1 * cos1);
  R r12s = (n1 * cos1 - n2 * cos2) / (n1 * cos1 + n2 * cos2);
  R r01p = (n1 * cos0 - n0 * cos1) / (n1 * cos0 + n0 * cos1);
  R r12p = (n2 * cos1 - n1 * cos2) / (n2 * cos1 + n1 * cos2);
  R ps = r01s * r12s, pp = r01p * r12p;
  R Rs = (r01s * r01s + r12s * r12s + (R)2 * ps * cd) / ((R)1 + ps * ps + (R)2 * ps * cd);
  R Rp = (r01p * r01p + r12p * r12p + (R)2 * pp * cd) / ((R)1 + pp * pp + (R)2 * pp * cd);
  return (R)0.5 * (Rs + Rp);
}

// Intersect surface k (sphere of curvature c through the vertex, or the vertex plane).
// On success the ray origin is the hit point and (nx,ny,nz) is the unit normal facing the ray.
template <typename R>
__device__ __forceinline__ bool hit_surface(int k, Ray<R>& r, R& nx, R& ny, R& nz, R& cos0) {
  const R c = (R)c_lens.c[k];
  const R zk = Num<R>::zv(k);
  R px = r.ox, py = r.oy, pz = r.oz - zk;
  R pd = px * r.dx + py * r.dy + pz * r.dz;
  R pp = px * px + py * py + pz * pz;
  R B = c * pd - r.dz;
  R Cq = c * pp - (R)2 * pz;
  R disc = B * B - c * Cq;
  if (disc < (R)0) { r.flags |= LFB_RAY_MISSED; return false; }
  R sq = Num<R>::sqrt(disc);
  R t = -Cq / (B + (B < (R)0 ? -sq : sq));
  R hx = px + t * r.dx, hy = py + t * r.dy, hz = pz + t * r.dz;
  r.ox = hx; r.oy = hy; r.oz = hz + zk;
  R sa = (R)c_lens.semi[k];
  if (hx * hx + hy * hy > sa * sa) { r.flags |= LFB_RAY_VIGNETTED; return false; }
  nx = -c * hx; ny = -c * hy; nz = (R)1 - c * hz;
  R nd = nx * r.dx + ny * r.dy + nz * r.dz;
  if (nd > (R)0) { nx = -nx; ny = -ny; nz = -nz; nd = -nd; }
  cos0 = -nd;
  return true;
}

template <typename R>
__device__ __forceinline__ bool refract_at(int lam, int k, bool forward, Ray<R>& r) {
  R nx, ny, nz, cos0;
  if (!hit_surface<R>(k, r, nx, ny, nz, cos0)) return false;
  R na = (R)(k == 0 ? 1.0f : c_lens.ior[lam][k - 1]), nb = (R)c_lens.ior[lam][k];
  R n0 = forward ? na : nb, n2 = forward ? nb : na;
  if (n0 == n2) return true;
  R eta = n0 / n2, k2 = (R)1 - eta * eta * ((R)1 - cos0 * cos0);
  if (k2 < (R)0) { r.flags |= LFB_RAY_TIR; return false; }
  R cos2 = Num<R>::sqrt(k2), f = eta * cos0 - cos2;
  r.dx = eta * r.dx + f * nx; r.dy = eta * r.dy + f * ny; r.dz = eta * r.dz + f * nz;
  r.w *= (R)1 - reflectance<R>(n0, n2, cos0, cos2, (R)c_lens.coat[k], (R)c_lens.lambda_nm[lam]);
  return true;
}

template <typename R>
__device__ __forceinline__ bool reflect_at(int lam, int k, bool forward, Ray<R>& r) {
  R nx, ny, nz, cos0;
  if (!hit_surface<R>(k, r, nx, ny, nz, cos0)) return false;
  R na = (R)(k == 0 ? 1.0f : c_lens.ior[lam][k - 1]), nb = (R)c_lens.ior[lam][k];
  R n0 = forward ? na : nb, n2 = forward ? nb : na;
  R two_c = (R)2 * cos0;
  r.dx += two_c * nx; r.dy += two_c * ny; r.dz += two_c * nz;
  R refl = (R)0;
  if (n0 != n2) {
    R eta = n0 / n2, k2 = (R)1 - eta * eta * ((R)1 - cos0 * cos0);
    refl = k2 < (R)0 ? (R)1
                     : reflectance<R>(n0, n2, cos0, Num<R>::sqrt(k2), (R)c_lens.coat[k], (R)c_lens.lambda_nm[lam]);
  }
  r.w *= refl;
  return true;
}

template <typename R>
__device__ __forceinline__ void to_plane(Ray<R>& r, R z) {
  R t = (z - r.oz) / r.dz;
  r.ox += t * r.dx; r.oy += t * r.dy; r.oz = z;
}

template <typename R>
__device__ __forceinline__ void cross_stop(const float* __restrict__ tex, int tw, int th, Ray<R>& r) {
  to_plane<R>(r, Num<R>::zv(c_lens.stop));
  r.xa = r.ox; r.ya = r.oy;
  R m = mask_lookup<R>(tex, tw, th, r.xa, r.ya);
  if (m == (R)0) r.flags |= LFB_RAY_STOPPED;
  r.w *= m;
}

// EXACT_GRID: trace one ray along ghost (i,j) (i<0: direct path).  Returns false when the
// ray died (missed / vignetted / TIR); STOPPED rays keep going with weight 0 so that their
// positions stay comparable.  EARLY_OUT lets the throughput kernel drop them at once.
template <typename R, bool EARLY_OUT>
__device__ __forceinline__ bool trace_exact(const Job& J, const float* __restrict__ tex, int tw, int th, R x, R y,
                                            Ray<R>& r) {
  const int n = c_lens.n_surfaces, stop = c_lens.stop, lam = J.lambda, i = J.i, j = J.j;
  r.ox = x; r.oy = y; r.oz = (R)0;
  r.dx = (R)J.sin_t; r.dy = (R)0; r.dz = (R)J.cos_t;
  r.w = (R)1; r.flags = 0; r.xa = r.ya = Num<R>::nan();
  if (J.inv_dist != 0.0) {
    // point light at -D (sin t, 0, cos t): the ray through (x, y, 0) runs along v = (x/D + sin t, y/D, cos t); its weight is
    // the irradiance at the entrance point relative to the vertex, (D/|E-P|)^2 * (cos of incidence / cos t) = |v|^-3
    const R inv_d = (R)J.inv_dist;
    const R vx = x * inv_d + r.dx, vy = y * inv_d, vz = r.dz;
    const R q = vx * vx + vy * vy + vz * vz;
    const R len = sqrt(q);
    r.dx = vx / len; r.dy = vy / len; r.dz = vz / len;
    r.w = (R)1 / (q * len);
  }
#define LFB_STEP(expr) do { if (!(expr)) return false; if (EARLY_OUT && r.w == (R)0) return false; } while (0)
#define LFB_FWD(k) do { if ((k) == stop) { cross_stop<R>(tex, tw, th, r); if (EARLY_OUT && r.w == (R)0) return false; } \
                        else LFB_STEP(refract_at<R>(lam, (k), true, r)); } while (0)
  if (i < 0) {
    for (int k = 0; k < n; k++) LFB_FWD(k);
  } else {
    for (int k = 0; k < j; k++) LFB_FWD(k);
    LFB_STEP(reflect_at<R>(lam, j, true, r));
    for (int k = j - 1; k > i; k--) {
      if (k == stop) { cross_stop<R>(tex, tw, th, r); if (EARLY_OUT && r.w == (R)0) return false; }
      else LFB_STEP(refract_at<R>(lam, k, false, r));
    }
    LFB_STEP(reflect_at<R>(lam, i, false, r));
    for (int k = i + 1; k < n; k++) LFB_FWD(k);
  }
#undef LFB_FWD
#undef LFB_STEP
  to_plane<R>(r, Num<R>::zv(n));
  return true;
}

template <typename R>
struct Hit {
  R xs, ys, xa, ya, w;
  unsigned flags;
  bool alive;
};

template <typename R, int MODE, bool EARLY_OUT>
__device__ __forceinline__ Hit<R> trace_ray(const Job& J, const FrameGeom& g, const float* __restrict__ tex, int a, int b) {
  const R P = (R)c_lens.P;
  const R cell = (R)2 * P / (R)g.N;
  const R x = -P + ((R)a + (R)0.5) * cell;
  const R y = -P + ((R)b + (R)0.5) * cell;
  Hit<R> h;
  if (MODE == LFB_MODE_PARAXIAL_GRID) {
    R th = (R)J.theta, th_y = (R)0;
    const bool point = J.inv_dist != 0.0;  // paraxial point light: the entrance angle grows linearly across the pupil
    if (point) { th = x * (R)J.inv_dist + th; th_y = y * (R)J.inv_dist; }
    R w = (R)1;
    h.flags = 0; h.xa = h.ya = Num<R>::nan();
    for (int c = 0; c < J.n_cross; c++) {
      R xa = x * (R)J.cross[c][0] + th * (R)J.cross[c][1];
      R ya = y * (R)J.cross[c][0];
      if (point) ya = ya + th_y * (R)J.cross[c][1];
      R m = mask_lookup<R>(tex, g.tex_w, g.tex_h, xa, ya);
      if (m == (R)0) h.flags |= LFB_RAY_STOPPED;
      w *= m;
      h.xa = xa; h.ya = ya;
    }
    h.xs = x * (R)J.full[0] + th * (R)J.full[1];
    h.ys = y * (R)J.full[0];
    if (point) h.ys = h.ys + th_y * (R)J.full[1];
    h.w = w; h.alive = true;
  } else {
    Ray<R> r;
    h.alive = trace_exact<R, EARLY_OUT>(J, tex, g.tex_w, g.tex_h, x, y, r);
    h.xs = r.ox; h.ys = r.oy; h.xa = r.xa; h.ya = r.ya; h.flags = r.flags;
    h.w = h.alive ? r.w : (R)0;
    if (!h.alive) h.xs = h.ys = Num<R>::nan();
  }
  return h;
}

template <typename R>
__device__ __forceinline__ void to_pixel(const Job& J, R xs, R ys, R& px, R& py) {
  R X = -(R)J.ppu * xs, Y = (R)J.ppu * ys;
  px = (R)J.sx + (X * (R)J.cs - Y * (R)J.sn);
  py = (R)J.sy + (X * (R)J.sn + Y * (R)J.cs);
}

template <typename R>
__device__ __forceinline__ void deposit(unsigned long long* __restrict__ accum, const FrameGeom& g, int ix, int iy,
                                        R wgt, const R* chan) {
  if (ix < 0 || ix >= g.W || iy < 0 || iy >= g.H) return;
  grow_bbox(g.bbox, ix, iy, ix, iy);
  if (g.tile_bits) mark_tile(g.tile_bits, g.tiles_w, ix >> kTilePxLog2, iy >> kTilePxLog2);
  unsigned long long* p = accum + 3 * ((size_t)ix + (size_t)iy * g.W);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    long long q = Num<R>::to_fixed(wgt * chan[c], g.fp_scale);
    if (q != 0) atomicAdd(p + c, (unsigned long long)q);
  }
}

template <typename R, int MODE>
__global__ void __launch_bounds__(256) trace_splat_kernel(const Job* __restrict__ jobs, FrameGeom g,
                                                          const float* __restrict__ tex,
                                                          unsigned long long* __restrict__ accum) {
  const int job_id = blockIdx.x / g.tiles_per_job;
  const int tile = blockIdx.x - job_id * g.tiles_per_job;
  const Job& J = jobs[job_id];
  const int a = (tile % g.tiles_x) * 16 + (threadIdx.x & 15);
  const int b = (tile / g.tiles_x) * 16 + (threadIdx.x >> 4);
  if (a >= g.N || b >= g.N) return;
  Hit<R> h = trace_ray<R, MODE, true>(J, g, tex, a, b);
  if (!h.alive || !(h.w > (R)0)) return;
  R px, py;
  to_pixel<R>(J, h.xs, h.ys, px, py);
  const R chan[3] = {(R)J.chan[0], (R)J.chan[1], (R)J.chan[2]};
  if (g.splat == LFB_SPLAT_NEAREST) {
    deposit<R>(accum, g, (int)Num<R>::floor(px), (int)Num<R>::floor(py), h.w, chan);
  } else {
    R qx = px - (R)0.5, qy = py - (R)0.5;
    R fx0 = Num<R>::floor(qx), fy0 = Num<R>::floor(qy);
    R fx = qx - fx0, fy = qy - fy0;
    int ix = (int)fx0, iy = (int)fy0;
    deposit<R>(accum, g, ix, iy, h.w * (((R)1 - fx) * ((R)1 - fy)), chan);
    deposit<R>(accum, g, ix + 1, iy, h.w * (fx * ((R)1 - fy)), chan);
    deposit<R>(accum, g, ix, iy + 1, h.w * (((R)1 - fx) * fy), chan);
    deposit<R>(accum, g, ix + 1, iy + 1, h.w * (fx * fy), chan);
  }
}

template <typename R, int MODE>
__global__ void __launch_bounds__(256) trace_dump_kernel(const Job* __restrict__ job, FrameGeom g,
                                                         const float* __restrict__ tex,
                                                         lfb_ray_hit* __restrict__ out) {
  const Job& J = *job;
  const int tile = blockIdx.x;
  const int a = (tile % g.tiles_x) * 16 + (threadIdx.x & 15);
  const int b = (tile / g.tiles_x) * 16 + (threadIdx.x >> 4);
  if (a >= g.N || b >= g.N) return;
  Hit<R> h = trace_ray<R, MODE, false>(J, g, tex, a, b);
  lfb_ray_hit rec;
  rec.x_s = h.xs; rec.y_s = h.ys; rec.x_ap = h.xa; rec.y_ap = h.ya;
  rec.weight = h.w; rec.flags = h.flags; rec.pad = 0;
  if (h.xs == h.xs) {
    R px, py;
    to_pixel<R>(J, h.xs, h.ys, px, py);
    rec.px = px; rec.py = py;
    R fx = Num<R>::floor(px), fy = Num<R>::floor(py);
    if (!(fx >= (R)0 && fx < (R)g.W && fy >= (R)0 && fy < (R)g.H)) rec.flags |= LFB_RAY_OFF_SENSOR;
  } else {
    rec.px = rec.py = CUDART_NAN;
  }
  out[(size_t)b * g.N + a] = rec;
}

template <typename R>
cudaError_t launch_trace_splat_t(const Job* jobs, int n_jobs, const FrameGeom& g, int mode, const float* tex,
                                 unsigned long long* accum, cudaStream_t s) {
  if (n_jobs <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)n_jobs * (unsigned)g.tiles_per_job;
  if (mode == LFB_MODE_PARAXIAL_GRID) trace_splat_kernel<R, LFB_MODE_PARAXIAL_GRID><<<blocks, 256, 0, s>>>(jobs, g, tex, accum);
  else trace_splat_kernel<R, LFB_MODE_EXACT_GRID><<<blocks, 256, 0, s>>>(jobs, g, tex, accum);
  return cudaGetLastError();
}

template <typename R>
cudaError_t launch_trace_dump_t(const Job* job, const FrameGeom& g, int mode, const float* tex, lfb_ray_hit* out,
                                cudaStream_t s) {
  const unsigned blocks = (unsigned)g.tiles_per_job;
  if (mode == LFB_MODE_PARAXIAL_GRID) trace_dump_kernel<R, LFB_MODE_PARAXIAL_GRID><<<blocks, 256, 0, s>>>(job, g, tex, out);
  else trace_dump_kernel<R, LFB_MODE_EXACT_GRID><<<blocks, 256, 0, s>>>(job, g, tex, out);
  return cudaGetLastError();
}

}  // namespace LFB_TU
}  // namespace lfb
