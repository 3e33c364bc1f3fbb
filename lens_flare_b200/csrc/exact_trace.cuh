// exact_trace.cuh -- the throughput kernels of the ghost path: EXACT_GRID.
//
// Templated on the geometry type T:
//   T = float   LFB_FP32    the FP32 kernels north_star asks for (exact_f32.cu)
//   T = double  LFB_STRICT  the same kernels with FP64 positions / directions (exact_f64.cu): per-ray sensor hits within
//                           1e-5 lens units of the double oracle (measured ~1e-9), which FP32 cannot give at |x| ~ 500
// Weights (Fresnel / coating products, mask products) are FP32 in both.
//
// There is no memory stream to speak of (rays are generated from their grid index, a ghost's whole description is ~1.5 KB,
// the 1 MB aperture mask is L2/L1 resident), so the design minimises instructions and stalls per ray (ncu, profiles/):
//
//  * STEP PROGRAM.  The host flattens each ghost (i, j, lambda) into a straight list of steps (refract / reflect / stop
//    plane / sensor plane) with every ray-independent quantity already computed (lfb_internal.h: StepT).  A CTA stages
//    its ghost's program in shared memory once; the per-ray loop has no "which surface / which direction / which glass"
//    logic and no divisions by constants.
//  * UNIFIED STEP.  Planes are c = 0 surfaces with an infinite clear radius; a missed surface or a total internal
//    reflection turns the state into NaNs that the next clear-radius test catches: ONE death test per step.
//  * FAST SCALAR MATH.  float: rcp.approx / sqrt.approx (one MUFU each).  double: the 20-bit MUFU seeds
//    (rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64) refined by two Newton steps, no IEEE division or square-root sequences.
//  * POLYNOMIAL WEIGHTS.  The factor a surface multiplies the ray's weight by (R at a reflection, 1 - R at a refraction;
//    bare Fresnel or the quarter-wave film) is a degree-7 polynomial in the cosine in the rarer medium, fitted per
//    (wavelength, surface, direction) on the host in double: 9 FMAs in their own dependency chain and NO memory access in
//    the surface loop (round 1's table lookups were 40 % of the kernel's stall cycles).  Rays steeper than acos(0.6)
//    fall back to the 1024-interval table.
//  * MIRROR SYMMETRY.  A light's bundle and the lens are symmetric about the meridional plane: ray (x, -y) is the mirror
//    image of ray (x, y).  One trace serves both; only the (asymmetric) aperture mask is looked up twice.
//  * PREFIX CACHE.  The forward sweep 0 .. j-1 shared by all ghosts of a (light, wavelength) is traced once
//    (prefix_kernel) and the ray states on every surface cached in L2; ghost jobs start ON their first-reflection surface.
//  * WARP-AUTONOMOUS SPLAT.  A warp owns its 32 ray pairs from trace to flush: survivors stay in registers, deposits go
//    to a 64-pixel u64 fixed-point tile per warp in shared memory and each touched pixel is flushed with one global atomic
//    (direct global atomics when the footprint exceeds the tile).  The only CTA barrier is the one after staging the
//    program.  Integer accumulation keeps the frame bit-stable for any schedule, kernel or GPU count.
//  * DIRTY TILES.  Every flush marks the 16 x 16 sensor tiles it touches in a byte map (one plain store per tile: no read,
//    no atomic), so that finalize / read-back / the cross-GPU reduce visit the ~1 % of the frame that is not zero.
//
// Kernels: prefix_kernel (forward sweeps + the direct path), ghost_kernel (one job per ghost pair), family_kernel (one
// thread follows every ghost sharing a first reflection; forks at each second reflection), dump_kernel (per-ray records).
// Parity: per-ray and image tolerances against the double-precision oracle are in tests/test_gpu_parity.py; the FP64
// kernels in ghost_grid_impl.cuh remain the oracle-order, bit-exact instruments.
#pragma once
#include <math_constants.h>
#include <string.h>

#include "lfb_internal.h"

namespace lfb {
namespace xt {

// ---------------------------------------------------------------------------------------------------------------
// scalar math per geometry type
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct M;
template <> struct M<float> {
  static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
  static __device__ __forceinline__ float rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
  }
  static __device__ __forceinline__ float sqrt(float x) {  // NaN for x < 0
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
  }
  static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
  static __device__ __forceinline__ float copysign(float m, float s) { return copysignf(m, s); }
  static __device__ __forceinline__ float floor(float x) { return floorf(x); }
  static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float nan() { return CUDART_NAN_F; }
  // g0 when nd carries a sign bit (the normal faces the ray), -g0 otherwise
  static __device__ __forceinline__ float flip_unless_neg(float g0, float nd) {
    return __int_as_float(__float_as_int(g0) ^ (~__float_as_int(nd) & 0x80000000));
  }
  static __device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
};
template <> struct M<double> {
  static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
  static __device__ __forceinline__ double rcp(double x) {  // 20-bit seed, two Newton steps: ~1e-16 relative
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = ::fma(-x, y, 1.0);
    y = ::fma(y, e, y);
    e = ::fma(-x, y, 1.0);
    return ::fma(y, e, y);
  }
  static __device__ __forceinline__ double sqrt(double x) {  // NaN for x < 0; 0 for 0
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y;
    double r = ::fma(-g, y, 1.0);
    y = ::fma(0.5 * y, r, y);
    g = x * y;
    r = ::fma(-g, y, 1.0);
    y = ::fma(0.5 * y, r, y);
    g = x * y;
    g = ::fma(::fma(-g, g, x), 0.5 * y, g);
    return x == 0.0 ? 0.0 : g;
  }
  static __device__ __forceinline__ double abs(double x) { return fabs(x); }
  static __device__ __forceinline__ double copysign(double m, double s) { return ::copysign(m, s); }
  static __device__ __forceinline__ double floor(double x) { return ::floor(x); }
  static __device__ __forceinline__ double max(double a, double b) { return fmax(a, b); }
  static __device__ __forceinline__ double nan() { return CUDART_NAN; }
  static __device__ __forceinline__ double flip_unless_neg(double g0, double nd) { return signbit(nd) ? g0 : -g0; }
  static __device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
};

// The per-job constants in the geometry type (the host fills both copies, lfb_internal.h: Job).
template <typename T> struct JobC;
template <> struct JobC<float> {
  float sin_t, cos_t, inv_d, sx, sy, cs, sn, ppu;
  __device__ __forceinline__ explicit JobC(const Job& J)
      : sin_t(J.f_sin_t), cos_t(J.f_cos_t), inv_d(J.f_inv_dist), sx(J.f_sx), sy(J.f_sy), cs(J.f_cs), sn(J.f_sn), ppu(J.f_ppu) {}
};
template <> struct JobC<double> {
  double sin_t, cos_t, inv_d, sx, sy, cs, sn, ppu;
  __device__ __forceinline__ explicit JobC(const Job& J)
      : sin_t(J.sin_t), cos_t(J.cos_t), inv_d(J.inv_dist), sx(J.sx), sy(J.sy), cs(J.cs), sn(J.sn), ppu(J.ppu) {}
};

struct MaskGeom {
  const float* tex;
  int tw, th;
  float su, sv, ou, ov;  // u = xa*su + ou ; v = ya*sv + ov
  __device__ __forceinline__ explicit MaskGeom(const FrameGeom& g, const float* t)
      : tex(t), tw(g.tex_w), th(g.tex_h), su(g.mask_su), sv(g.mask_sv), ou(g.mask_ou), ov(g.mask_ov) {}
};

template <typename T>
__device__ __forceinline__ float mask_lookup(const MaskGeom& K, T xa, T ya) {
  const T fu = M<T>::floor(M<T>::fma(xa, (T)K.su, (T)K.ou)), fv = M<T>::floor(M<T>::fma(ya, (T)K.sv, (T)K.ov));
  if (!(fu >= (T)0 && fu < (T)K.tw && fv >= (T)0 && fv < (T)K.th)) return 0.f;
  return __ldg(K.tex + ((unsigned)(int)fv * (unsigned)K.tw + (unsigned)(int)fu));  // in range, checked above
}

// Reflectance from the interface's table (rays steeper than acos(kPolyV0) only): R over v = the cosine of the ray's angle
// in the RARER of the two media (in which R is analytic: R -> 1 linearly as v -> 0, i.e. at grazing incidence or at the
// critical angle), 1024 intervals, linear interpolation of (R_i, R_{i+1} - R_i) pairs built on the host in double.
__device__ __forceinline__ float reflectance_lut(const float2* __restrict__ lut, int table, float v) {
  const float t = fminf(fmaxf(v, 0.f), 1.f) * (float)kLutSize;  // fmaxf(NaN, 0) = 0: beyond the critical angle R = 1
  const int i = min((int)t, kLutSize - 1);
  const float2 e = __ldg(lut + ((unsigned)table * (unsigned)kLutSize + (unsigned)i));
  return fmaf(t - (float)i, e.y, e.x);
}

// The factor a step multiplies the weight by: the step's polynomial (Estrin: depth 4) on v >= kPolyV0, else the table.
template <typename T>
__device__ __forceinline__ float weight_factor(const StepT<T>& S, bool refl, const float2* __restrict__ lut, float v0, float v) {
  if (v >= v0) {
    const float x = fmaf(-v, kPolyScale, kPolyScale);  // (1 - v) / (1 - v0) in [0, 1]
    const float x2 = x * x, x4 = x2 * x2;
    const float a = fmaf(S.p[1], x, S.p[0]), b = fmaf(S.p[3], x, S.p[2]), c = fmaf(S.p[5], x, S.p[4]), d = fmaf(S.p[7], x, S.p[6]);
    return fmaf(x4, fmaf(d, x2, c), fmaf(b, x2, a));
  }
  const float R = reflectance_lut(lut, S.lut, v);  // v NaN (total internal reflection) lands here: R = 1
  return refl ? R : 1.f - R;
}

// A ray on (or heading for) a surface.  Position is relative to the current surface's vertex.
template <typename T>
struct RayState {
  T ox, oy, oz, dx, dy, dz;
  float w, ma, mb;  // Fresnel product; mask products of the ray and of its mirror image
};

struct RayDiag {  // parity-instrument variant only
  unsigned flags;
  double xa, ya;
};

// First half of a step: move the ray onto the step's surface (sphere or plane through its vertex) and test the clear
// aperture.  False = the ray is dead (missed, vignetted, or NaN from a total internal reflection upstream).
template <typename T, bool FLAGS>
__device__ __forceinline__ bool propagate(const StepT<T>& S, RayState<T>& r, RayDiag& o) {
  typedef M<T> m;
  const T c = S.c;
  const T pz = r.oz + S.dz;
  const T pd = m::fma(r.ox, r.dx, m::fma(r.oy, r.dy, pz * r.dz));
  const T pp = m::fma(r.ox, r.ox, m::fma(r.oy, r.oy, pz * pz));
  const T B = m::fma(c, pd, -r.dz);
  const T Cq = m::fma(c, pp, (T)-2 * pz);
  const T disc = m::fma(B, B, -c * Cq);
  if (FLAGS && disc < (T)0) { o.flags |= LFB_RAY_MISSED; return false; }
  const T t = -Cq * m::rcp(B + m::copysign(m::sqrt(disc), B));
  r.ox = m::fma(t, r.dx, r.ox); r.oy = m::fma(t, r.dy, r.oy); r.oz = m::fma(t, r.dz, pz);
  if (!(m::fma(r.ox, r.ox, r.oy * r.oy) <= (T)S.semi2)) {  // outside the clear aperture, or NaN from a miss / TIR upstream
    if (FLAGS) o.flags |= LFB_RAY_VIGNETTED;
    return false;
  }
  return true;
}

// Second half: what the surface does to the ray (mask lookup at the stop; refraction or reflection + weight elsewhere).
//   MIRROR  also look the mask up at (xa, -ya) for the mirror-image ray
//   FLAGS   parity-instrument variant: classify the death, and let rays the mask stopped continue with weight 0 so that
//           their positions stay comparable with the oracle
template <typename T, bool MIRROR, bool FLAGS>
__device__ __forceinline__ bool interact(const StepT<T>& S, const MaskGeom& K, const FrameGeom& g, RayState<T>& r, RayDiag& o) {
  typedef M<T> m;
  const int op = S.op;
  if (op >= STEP_PASS) {  // no change of direction: identical media, the stop, the sensor
    if (op == STEP_STOP) {
      const float mk = mask_lookup<T>(K, r.ox, r.oy);
      r.ma *= mk;
      if (MIRROR) r.mb *= mask_lookup<T>(K, r.ox, -r.oy);
      if (FLAGS) { o.xa = (double)r.ox; o.ya = (double)r.oy; if (mk == 0.f) o.flags |= LFB_RAY_STOPPED; }
      else if (MIRROR ? (r.ma == 0.f && r.mb == 0.f) : (r.ma == 0.f)) return false;
    }
    return true;
  }
  const T c = S.c, eta = S.eta;
  const T nx = -c * r.ox, ny = -c * r.oy, nz = m::fma(-c, r.oz, (T)1);
  const T nd = m::fma(nx, r.dx, m::fma(ny, r.dy, nz * r.dz));
  const T c0 = m::abs(nd);
  const T s2 = m::fma(-c0, c0, (T)1);
  const T k2 = m::fma(-S.eta2, s2, (T)1);
  const T c2 = m::sqrt(k2);  // NaN beyond the critical angle
  const bool refl = op == STEP_REFLECT;
  if (FLAGS && !refl && k2 < (T)0) { o.flags |= LFB_RAY_TIR; return false; }
  // refract: d' = eta d + (eta c0 - c2) N with N = -sign(nd) n the normal facing the ray;  reflect: d' = d - 2 nd n
  const T gg = m::flip_unless_neg(m::fma(eta, c0, -c2), nd);
  const T alpha = refl ? (T)1 : eta, beta = refl ? (T)-2 * nd : gg;
  r.dx = m::fma(alpha, r.dx, beta * nx); r.dy = m::fma(alpha, r.dy, beta * ny); r.dz = m::fma(alpha, r.dz, beta * nz);
  r.w *= weight_factor<T>(S, refl, g.lut, g.poly_v0, (float)(eta > (T)1 ? c2 : c0));
  return true;
}

// Entrance ray of grid point (x, y).  Directional light (inv_d = 0): the bundle's common direction.  Point light at
// -D (sin t, 0, cos t): along v = (x/D + sin t, y/D, cos t), carrying the irradiance at the entrance point relative to the
// vertex, |v|^-3 (lfb_light.distance).  The light lies in the plane y = 0, so the mirror pair (x, -y) stays a mirror pair.
template <typename T>
__device__ __forceinline__ void start_ray(RayState<T>& r, T x, T y, const JobC<T>& J) {
  typedef M<T> m;
  r.ox = x; r.oy = y; r.oz = (T)0; r.dx = J.sin_t; r.dy = (T)0; r.dz = J.cos_t; r.w = 1.f; r.ma = 1.f; r.mb = 1.f;
  if (J.inv_d != (T)0) {  // uniform per job
    const T vx = m::fma(x, J.inv_d, J.sin_t), vy = m::mul_rn(y, J.inv_d);
    const T q = m::fma(vx, vx, m::fma(vy, vy, m::mul_rn(J.cos_t, J.cos_t)));
    const T rl = m::rcp(m::sqrt(q));
    r.dx = m::mul_rn(vx, rl); r.dy = m::mul_rn(vy, rl); r.dz = m::mul_rn(J.cos_t, rl);
    r.w = (float)m::mul_rn(m::mul_rn(rl, rl), rl);
  }
}

template <typename T>
__device__ __forceinline__ void grid_point(const FrameGeom& g, int a, int b, T& x, T& y) {
  x = M<T>::fma((T)a + (T)0.5, (T)g.cell_d, -(T)g.P_d);
  y = M<T>::fma((T)b + (T)0.5, (T)g.cell_d, -(T)g.P_d);
}
template <>
__device__ __forceinline__ void grid_point<float>(const FrameGeom& g, int a, int b, float& x, float& y) {
  x = fmaf((float)a + 0.5f, g.cell, -g.P);
  y = fmaf((float)b + 0.5f, g.cell, -g.P);
}

// ---------------------------------------------------------------------------------------------------------------
// PREFIX CACHE.  Every ghost (i, j) of a light and wavelength begins with the same forward sweep through surfaces
// 0 .. j-1.  prefix_kernel traces that sweep ONCE per (light, wavelength) "slot" and caches each ray's state as it
// arrives ON every surface (NaN = dead); a ghost job loads the state at its first-reflection surface j and starts with
// the reflection.  Layout: prefix[((slot * n_surf + k) * PARTS + part) * half_rays + ray], 16-byte parts:
//   float  (PARTS = 2)  (ox, oy, oz, w) (dx, dy, ma, mb); dz = +sqrt(1 - dx^2 - dy^2) on load (the sweep travels forward)
//   double (PARTS = 4)  (ox, oy) (oz, dx) (dy, dz) (w, ma, mb, -)
// canon(): what a load does to a state that was never stored -- applied by jobs that trace their own forward sweep (no
// cache, e.g. over the memory budget), so that a frame has the same bits with or without the cache.
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct PrefixIO;
template <> struct PrefixIO<float> {
  static constexpr int kParts = 2;
  struct Raw { float4 a, b; };  // a cached state as loaded: nothing is decoded (no instruction waits on the loads) until take()
  static __device__ __forceinline__ void store(float4* dst, size_t hr, const RayState<float>& r) {
    dst[0] = make_float4(r.ox, r.oy, r.oz, r.w);
    dst[hr] = make_float4(r.dx, r.dy, r.ma, r.mb);
  }
  static __device__ __forceinline__ void store_dead(float4* dst, size_t hr) {
    dst[0] = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f);
  }
  static __device__ __forceinline__ void canon(RayState<float>& r) {
    r.dz = M<float>::sqrt(fmaxf(fmaf(-r.dx, r.dx, fmaf(-r.dy, r.dy, 1.f)), 0.f));
  }
  static __device__ __forceinline__ Raw issue(const float4* __restrict__ src, size_t hr) {
    Raw q;
    q.a = __ldg(src);
    q.b = __ldg(src + hr);  // unconditionally: a dead ray's second part is stale, never used
    return q;
  }
  static __device__ __forceinline__ bool take(const Raw& q, RayState<float>& r) {
    if (!(q.a.x == q.a.x)) return false;  // NaN: the ray died in the forward sweep before reaching this surface
    r.ox = q.a.x; r.oy = q.a.y; r.oz = q.a.z; r.w = q.a.w;
    r.dx = q.b.x; r.dy = q.b.y; r.ma = q.b.z; r.mb = q.b.w;
    canon(r);
    return true;
  }
};
template <> struct PrefixIO<double> {
  static constexpr int kParts = 4;
  struct Raw { double2 a, b, c; float4 w; };
  static __device__ __forceinline__ void store(float4* dst, size_t hr, const RayState<double>& r) {
    double2* d = reinterpret_cast<double2*>(dst);
    d[0] = make_double2(r.ox, r.oy);
    d[hr] = make_double2(r.oz, r.dx);
    d[2 * hr] = make_double2(r.dy, r.dz);
    dst[3 * hr] = make_float4(r.w, r.ma, r.mb, 0.f);
  }
  static __device__ __forceinline__ void store_dead(float4* dst, size_t) {
    reinterpret_cast<double2*>(dst)[0] = make_double2(CUDART_NAN, 0.0);
  }
  static __device__ __forceinline__ void canon(RayState<double>&) {}
  static __device__ __forceinline__ Raw issue(const float4* __restrict__ src, size_t hr) {
    const double2* s = reinterpret_cast<const double2*>(src);
    Raw q;
    q.a = __ldg(s); q.b = __ldg(s + hr); q.c = __ldg(s + 2 * hr);
    q.w = __ldg(src + 3 * hr);
    return q;
  }
  static __device__ __forceinline__ bool take(const Raw& q, RayState<double>& r) {
    if (!(q.a.x == q.a.x)) return false;
    r.ox = q.a.x; r.oy = q.a.y; r.oz = q.b.x; r.dx = q.b.y; r.dy = q.c.x; r.dz = q.c.y;
    r.w = q.w.x; r.ma = q.w.y; r.mb = q.w.z;
    return true;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// landing: sensor point -> pixel taps -> per-warp shared-memory tile -> global atomics + dirty-tile marks
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWarpTilePx = 64;  // per-warp shared-memory sensor tile (pixels)

struct Tap {  // one image of a ray on the sensor: top-left pixel of its footprint and the bilinear fractions
  int ix, iy;
  float fx, fy;
};

template <typename T>
__device__ __forceinline__ Tap to_tap(const JobC<T>& J, bool bilinear, T xs, T ys) {
  typedef M<T> m;
  const T X = -J.ppu * xs, Y = J.ppu * ys;
  const T px = J.sx + m::fma(X, J.cs, -m::mul_rn(Y, J.sn));
  const T py = J.sy + m::fma(X, J.sn, m::mul_rn(Y, J.cs));
  Tap t;
  if (bilinear) {
    const T qx = m::sub_rn(px, (T)0.5), qy = m::sub_rn(py, (T)0.5);
    const T fx0 = m::floor(qx), fy0 = m::floor(qy);
    t.fx = (float)m::sub_rn(qx, fx0); t.fy = (float)m::sub_rn(qy, fy0);
    // NaN or far outside: park the tap where the footprint test rejects it (int conversion of huge values is undefined)
    const bool sane = fx0 >= (T)-2 && fx0 < (T)65536 && fy0 >= (T)-2 && fy0 < (T)65536;
    t.ix = sane ? (int)fx0 : -2; t.iy = sane ? (int)fy0 : -2;
  } else {
    const T fx0 = m::floor(px), fy0 = m::floor(py);
    const bool sane = fx0 >= (T)-1 && fx0 < (T)65536 && fy0 >= (T)-1 && fy0 < (T)65536;
    t.fx = t.fy = 0.f;
    t.ix = sane ? (int)fx0 : -1; t.iy = sane ? (int)fy0 : -1;
  }
  return t;
}

// Footprint of one tap clipped to the sensor (bilinear: 2x2 pixels from (ix, iy); nearest: the pixel itself).
__device__ __forceinline__ bool footprint(bool bilinear, const Tap& t, int W, int H, int& x0, int& y0, int& x1, int& y1) {
  if (bilinear) {
    if (!(t.ix >= -1 && t.ix < W && t.iy >= -1 && t.iy < H)) return false;
    x0 = max(t.ix, 0); y0 = max(t.iy, 0); x1 = min(t.ix + 1, W - 1); y1 = min(t.iy + 1, H - 1);
  } else {
    if (!(t.ix >= 0 && t.ix < W && t.iy >= 0 && t.iy < H)) return false;
    x0 = x1 = t.ix; y0 = y1 = t.iy;
  }
  return true;
}

struct WarpSplat {  // what a warp needs to deposit its lanes' results
  unsigned long long* tile;   // this warp's shared-memory tile
  unsigned long long* accum;  // global sensor accumulators
  int* bbox;
  unsigned* tile_bits;
  int tiles_w;
  int W, H;
  float ch0, ch1, ch2;        // radiance * rgb_weight * ray area * 2^bits
  bool bilinear;
  bool peek_l1;               // look at the tile bytes through L1 (default; lfb_options.experiment bit 0 = in L2)
  __device__ __forceinline__ WarpSplat(const FrameGeom& g, const Job& J, unsigned long long* tile_, unsigned long long* accum_)
      : tile(tile_), accum(accum_), bbox(g.bbox), tile_bits(g.tile_bits), tiles_w(g.tiles_w), W(g.W), H(g.H), ch0(J.f_chan[0]),
        ch1(J.f_chan[1]), ch2(J.f_chan[2]), bilinear(g.splat == LFB_SPLAT_BILINEAR), peek_l1((g.pad2 & 1) == 0) {}
};

// Shared-memory accesses of the per-warp tile by 32-bit shared address: the tile pointer travels through structs and
// selects, where the compiler loses the address space and would emit generic loads / atomics (L1TEX path, long scoreboard).
// volatile keeps them in program order among themselves and with the __syncwarp()s that separate the tile's phases (zero |
// deposit | flush); NO "memory" clobber: that would also pin every ordinary load between them, and the job constants a
// landing needs would each be fetched from L2 at their first use, one round trip after the other (measured: +60 %).
__device__ __forceinline__ void sts_u64(unsigned addr, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v)); }
__device__ __forceinline__ unsigned long long lds_u64(unsigned addr) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
// 64-bit add into the tile with NATIVE 32-bit shared atomics (a 64-bit shared-memory add compiles to a compare-and-swap spin
// loop, ATOMS.CAST.SPIN): add the low word, carry into the high word.  The carries commute with everything else, so the
// pair holds the exact sum mod 2^64 in any order; deposits are < 2^32 almost always, so this is one atomic.
__device__ __forceinline__ void tile_add_u64(unsigned addr, unsigned long long v) {
  const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  unsigned old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(lo));
  const unsigned up = hi + ((old + lo) < old ? 1u : 0u);
  if (up) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr + 4u), "r"(up));
}
__device__ __forceinline__ void redg_add_u64(unsigned long long* p, unsigned long long v) { asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v)); }
// Pull the cache line with a job's float constants (pixel mapping, channel weights: lfb_internal.h Job::f_*) into L1 at kernel
// start: they are first needed when a ray lands, and must not cost an L2 round trip there.
__device__ __forceinline__ void prefetch_job_constants(const Job& J) { asm volatile("prefetch.global.L1 [%0];" ::"l"(&J.f_sin_t)); }

// The pixels of one tap into the warp's shared-memory tile (TILE: origin (tx0, ty0), row pitch tw) or straight into the
// accumulators.  Explicit roundings: no FMA contraction, so every instantiation of every kernel produces the same bits.
template <bool TILE>
__device__ __forceinline__ void splat_tap(const WarpSplat& S, unsigned tile, int tx0, int ty0, int tw, const Tap& t, float w) {
  float wt[4];
  if (S.bilinear) {
    const float gx = __fsub_rn(1.f, t.fx), gy = __fsub_rn(1.f, t.fy);
    wt[0] = __fmul_rn(w, __fmul_rn(gx, gy)); wt[1] = __fmul_rn(w, __fmul_rn(t.fx, gy));
    wt[2] = __fmul_rn(w, __fmul_rn(gx, t.fy)); wt[3] = __fmul_rn(w, __fmul_rn(t.fx, t.fy));
  } else {
    wt[0] = w; wt[1] = wt[2] = wt[3] = 0.f;
  }
  unsigned char* mark[4];
  unsigned seen[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {  // global path: LOOK at the four pixels' tile bytes first, so that the looks overlap
    mark[q] = nullptr; seen[q] = 1;
    if (!TILE && S.tile_bits) {
      const int jx = t.ix + (q & 1), jy = t.iy + (q >> 1);
      if (!(wt[q] == 0.f || jx < 0 || jx >= S.W || jy < 0 || jy >= S.H)) {
        mark[q] = tile_byte(S.tile_bits, S.tiles_w, jx >> kTilePxLog2, jy >> kTilePxLog2);
        seen[q] = tile_peek(mark[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int jx = t.ix + (q & 1), jy = t.iy + (q >> 1);
    if (wt[q] == 0.f || jx < 0 || jx >= S.W || jy < 0 || jy >= S.H) continue;
    // channels a wavelength does not feed (RGB lenses: two of three) are skipped without converting anything
    const long long v0 = S.ch0 != 0.f ? __float2ll_rn(__fmul_rn(wt[q], S.ch0)) : 0ll;
    const long long v1 = S.ch1 != 0.f ? __float2ll_rn(__fmul_rn(wt[q], S.ch1)) : 0ll;
    const long long v2 = S.ch2 != 0.f ? __float2ll_rn(__fmul_rn(wt[q], S.ch2)) : 0ll;
    if (TILE) {
      const unsigned dst = tile + 24u * (unsigned)((jy - ty0) * tw + (jx - tx0));
      if (v0) tile_add_u64(dst, (unsigned long long)v0);
      if (v1) tile_add_u64(dst + 8, (unsigned long long)v1);
      if (v2) tile_add_u64(dst + 16, (unsigned long long)v2);
    } else {
      unsigned long long* dst = S.accum + 3 * ((size_t)jx + (size_t)jy * S.W);
      if (v0) redg_add_u64(dst, (unsigned long long)v0);
      if (v1) redg_add_u64(dst + 1, (unsigned long long)v1);
      if (v2) redg_add_u64(dst + 2, (unsigned long long)v2);
    }
  }
  if (!TILE) {
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (!seen[q]) *mark[q] = 1;
  }
}

// Deposit one image per lane (`lands`: this lane has one, at tap t with weight w); warp-collective (all 32 lanes call it).
// The footprint of the warp's landing lanes decides: <= 64 pixels -> the warp's shared-memory tile, flushed with one global
// atomic per touched (pixel, channel); larger -> straight to the accumulators.
__device__ __forceinline__ void land_image(const WarpSplat& S, bool lands, const Tap& t, float w, int lane) {
  const unsigned landing = __ballot_sync(0xffffffffu, lands);
  if (!landing) return;
  int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
  if (lands) {
    footprint(S.bilinear, t, S.W, S.H, bx0, by0, bx1, by1);  // true: `lands` was decided by it
    bx0 = __reduce_min_sync(landing, bx0); by0 = __reduce_min_sync(landing, by0);
    bx1 = __reduce_max_sync(landing, bx1); by1 = __reduce_max_sync(landing, by1);
  }
  const int leader = __ffs(landing) - 1;
  bx0 = __shfl_sync(0xffffffffu, bx0, leader); by0 = __shfl_sync(0xffffffffu, by0, leader);
  bx1 = __shfl_sync(0xffffffffu, bx1, leader); by1 = __shfl_sync(0xffffffffu, by1, leader);
  if (lane == 0) grow_bbox(S.bbox, bx0, by0, bx1, by1);
  const int tw = bx1 - bx0 + 1;
  const int area = tw * (by1 - by0 + 1);
  if (area > kWarpTilePx) {  // (uniform) the footprint exceeds the tile: straight to the accumulators
    if (lands) splat_tap<false>(S, 0u, bx0, by0, tw, t, w);
    return;
  }
  const unsigned tile = (unsigned)__cvta_generic_to_shared(S.tile);
  // the dirty-tile bytes under the footprint (a box of <= 64 pixels spans at most 5 x 1 ... 2 x 2 ... 1 x 5 tiles: lanes 0 .. 29 take
  // a 6 x 5 patch): LOOK now, store after the splat, so that nothing waits for the look
  unsigned char* mark = nullptr;
  unsigned seen = 1;
  if (S.tile_bits) {
    const int tx = (bx0 >> kTilePxLog2) + lane % 6, ty = (by0 >> kTilePxLog2) + lane / 6;
    if (tx <= (bx1 >> kTilePxLog2) && ty <= (by1 >> kTilePxLog2)) {
      mark = tile_byte(S.tile_bits, S.tiles_w, tx, ty);
      seen = S.peek_l1 ? (unsigned)__ldca(mark) : tile_peek(mark);
    }
  }
  for (int q = lane; q < 3 * area; q += 32) sts_u64(tile + 8u * (unsigned)q, 0ull);
  __syncwarp();
  if (lands) splat_tap<true>(S, tile, bx0, by0, tw, t, w);
  __syncwarp();
  const float inv_tw = M<float>::rcp((float)tw);
  for (int q = lane; q < area; q += 32) {
    const int jy = (int)(((float)q + 0.5f) * inv_tw);  // q / tw, exact for these small integers
    const int jx = q - jy * tw;
    unsigned long long* dst = S.accum + 3 * ((size_t)(bx0 + jx) + (size_t)(by0 + jy) * S.W);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const unsigned long long v = lds_u64(tile + 8u * (unsigned)(3 * q + c));
      if (v) redg_add_u64(dst + c, v);
    }
  }
  if (!seen) *mark = 1;
  __syncwarp();  // the tile is reused by the warp's next landing
}

static __device__ __noinline__ void land_image_call(const WarpSplat& S, bool lands, const Tap& t, float w, int lane) { land_image(S, lands, t, w, lane); }

// Deposit one (ray, mirror image) result per lane; warp-collective.  The two images land on opposite sides of the meridional
// axis, far apart for most ghosts, so each gets its own footprint / tile pass (one box around both would exceed the tile and
// send every tap to the accumulators one atomic at a time -- what round 1 did).  Returns whether this lane landed (statistics).
template <typename T, bool CALL = false>
__device__ __forceinline__ bool warp_land(const WarpSplat& S, const JobC<T>& J, bool alive, T xs, T ys, float wa, float wb, bool has_mirror,
                                          int lane) {
  Tap ta, tb;
  ta.ix = ta.iy = tb.ix = tb.iy = 0; ta.fx = ta.fy = tb.fx = tb.fy = 0.f;
  bool la = false, lb = false;
  int x0, y0, x1, y1;
  if (alive && wa > 0.f) {
    ta = to_tap<T>(J, S.bilinear, xs, ys);
    la = footprint(S.bilinear, ta, S.W, S.H, x0, y0, x1, y1);
  }
  if (alive && wb > 0.f && has_mirror) {
    tb = to_tap<T>(J, S.bilinear, xs, -ys);
    lb = footprint(S.bilinear, tb, S.W, S.H, x0, y0, x1, y1);
  }
  if (CALL) {  // experiment: one out-of-line copy of the landing code instead of two inlined ones
    land_image_call(S, la, ta, wa, lane);
    land_image_call(S, lb, tb, wb, lane);
  } else {
    land_image(S, la, ta, wa, lane);
    land_image(S, lb, tb, wb, lane);
  }
  return la || lb;
}

// statistics (STATS instantiations): executed surface steps / ray pairs started / ray pairs landed, one atomic each per warp
__device__ __forceinline__ void flush_stats(unsigned long long* stats, unsigned steps, unsigned started, unsigned landed) {
  steps = __reduce_add_sync(0xffffffffu, steps);
  started = __reduce_add_sync(0xffffffffu, started);
  landed = __reduce_add_sync(0xffffffffu, landed);
  if ((threadIdx.x & 31) == 0 && stats) {
    if (steps) atomicAdd(stats + 0, (unsigned long long)steps);
    if (started) atomicAdd(stats + 1, (unsigned long long)started);
    if (landed) atomicAdd(stats + 2, (unsigned long long)landed);
  }
}

template <typename T>
__device__ __forceinline__ void stage_program(StepT<T>* dst, const StepT<T>* __restrict__ src, int n_steps, int tid, int nthreads) {
  constexpr int W16 = sizeof(StepT<T>) / 16;
  const float4* s = reinterpret_cast<const float4*>(src);
  float4* d = reinterpret_cast<float4*>(dst);
  for (int q = tid; q < n_steps * W16; q += nthreads) d[q] = __ldg(s + q);
}

// Two-phase staging for the ghost / family kernels: the program's 16-byte words are requested into registers first (at most
// two per thread: LFB_MAX_STEPS * sizeof(StepD) / 16 = 250 <= 2 * 128), other loads are issued, and only then are they
// committed to shared memory -- so the wait for the program overlaps the wait for the job header and the ray state.
struct Staged { float4 w0, w1; };
template <typename T, int BT>
__device__ __forceinline__ Staged stage_issue(const StepT<T>* __restrict__ src, int n_steps, int tid) {
  static_assert(LFB_MAX_STEPS * (sizeof(StepT<T>) / 16) <= 2 * BT, "two staged words per thread");
  const int nq = n_steps * (int)(sizeof(StepT<T>) / 16);
  const float4* s = reinterpret_cast<const float4*>(src);
  Staged q;
  if (tid < nq) q.w0 = __ldg(s + tid);
  if (tid + BT < nq) q.w1 = __ldg(s + tid + BT);
  return q;
}
template <typename T, int BT>
__device__ __forceinline__ void stage_commit(StepT<T>* dst, const Staged& q, int n_steps, int tid) {
  const int nq = n_steps * (int)(sizeof(StepT<T>) / 16);
  float4* d = reinterpret_cast<float4*>(dst);
  if (tid < nq) d[tid] = q.w0;
  if (tid + BT < nq) d[tid + BT] = q.w1;
}

constexpr int kPrefixThreads = 256;

template <typename T, bool STATS>
__global__ void __launch_bounds__(kPrefixThreads) prefix_kernel(const Job* __restrict__ slots, const StepT<T>* __restrict__ progs, FrameGeom g,
                                                                const float* __restrict__ tex, float4* __restrict__ prefix,
                                                                unsigned long long* __restrict__ accum) {
  typedef PrefixIO<T> io;
  __shared__ __align__(16) StepT<T> s_prog[LFB_MAX_SURFACES + 2];
  __shared__ unsigned long long s_tile[(kPrefixThreads / 32) * kWarpTilePx * 3];
  const int half_rows = (g.N + 1) / 2;
  const int slot = blockIdx.z;  // 3-D grid (patch column, patch row, slot)
  const Job& J = slots[slot];
  const int n_steps = J.n_steps;  // forward refractions 0 .. n_surf-1 (the stop included); step n_surf is the sensor
  const bool do_direct = accum != nullptr && J.i != 0;  // this shard owns the slot's direct (unreflected) path: splat it too
  const int tid = threadIdx.x;
  stage_program<T>(s_prog, progs + (size_t)slot * LFB_MAX_STEPS, n_steps + 1, tid, kPrefixThreads);
  __syncthreads();
  const int a = blockIdx.x * 16 + (tid & 15), bp = blockIdx.y * 16 + (tid >> 4);
  const bool in_grid = a < g.N && bp < half_rows;
  const int b = g.N - 1 - bp;
  const size_t ray = (size_t)bp * g.N + a;
  const size_t hr = (size_t)g.half_rays;
  const MaskGeom K(g, tex);
  const JobC<T> JC(J);
  RayState<T> r;
  {
    T x, y;
    grid_point<T>(g, a, b, x, y);
    start_ray<T>(r, x, y, JC);
  }
  RayDiag o;
  bool alive = in_grid;
  unsigned n_exec = 0;
  float4* base = prefix + (size_t)slot * g.n_surf * io::kParts * hr + ray;
#pragma unroll 1
  for (int s = 0; s < n_steps; s++) {
    const StepT<T>& S = s_prog[s];
    if (alive) { alive = propagate<T, false>(S, r, o); if (STATS) n_exec++; }
    if (in_grid) {
      float4* dst = base + (size_t)s * io::kParts * hr;
      if (alive) io::store(dst, hr, r);
      else io::store_dead(dst, hr);
    }
    if (alive) alive = interact<T, true, false>(S, K, g, r, o);
  }
  bool landed = false;
  if (do_direct) {  // the direct path: on to the sensor plane and splat, one warp at a time
    if (alive) { alive = propagate<T, false>(s_prog[n_steps], r, o); if (STATS) n_exec++; }
    const WarpSplat WS(g, J, s_tile + (tid >> 5) * (kWarpTilePx * 3), accum);
    landed = warp_land<T>(WS, JC, alive, r.ox, r.oy, r.w * r.ma, r.w * r.mb, b != bp, tid & 31);
  }
  if (STATS) flush_stats(g.stats, n_exec, in_grid ? 1u : 0u, landed ? 1u : 0u);
}

// ---------------------------------------------------------------------------------------------------------------
// GHOST KERNEL: one job per ghost pair (i, j).  BT threads = 16 x BT/16 ray pairs of the upper half grid.
// ---------------------------------------------------------------------------------------------------------------
// STAGED = false (an experiment, lfb_options.ctas_per_sm = 2): the program is not staged in shared memory -- every warp reads
// its steps straight from global memory (warp-uniform 16-byte loads, L1-resident: a program is ~1.5 KB and the CTAs of an SM
// work on a handful of jobs), so the kernel has NO CTA barrier and no staging wait in its prologue.
template <typename T, bool STAGED>
struct StepFetch {
  const StepT<T>* base;
  __device__ __forceinline__ const StepT<T>& operator[](int s) const { return base[s]; }
};
template <typename T>
struct StepFetch<T, false> {
  const StepT<T>* base;
  __device__ __forceinline__ StepT<T> operator[](int s) const {
    StepT<T> S;
    const float4* src = reinterpret_cast<const float4*>(base + s);
    float4* dst = reinterpret_cast<float4*>(&S);
#pragma unroll
    for (int q = 0; q < (int)(sizeof(StepT<T>) / 16); q++) dst[q] = __ldg(src + q);
    return S;
  }
};

template <typename T, int MINB, int BT, bool STATS, bool STAGED = true, bool CALL = false>
__global__ void __launch_bounds__(BT, MINB) ghost_kernel(const Job* __restrict__ jobs, const StepT<T>* __restrict__ progs,
                                                         const __grid_constant__ JobHeads heads, FrameGeom g, const float* __restrict__ tex,
                                                         unsigned long long* __restrict__ accum) {
  typedef PrefixIO<T> io;
  constexpr int PH = BT / 16;
  __shared__ __align__(16) StepT<T> s_prog_mem[STAGED ? LFB_MAX_STEPS : 1];
  __shared__ unsigned long long s_tile[(BT / 32) * kWarpTilePx * 3];
  // 3-D grid (patch column, patch row, job): no integer divisions in the prologue every warp pays
  const Job& J = jobs[blockIdx.z];
  const int tid = threadIdx.x, lane = tid & 31;
  const int half_rows = (g.N + 1) / 2;
  const int a = blockIdx.x * 16 + (tid & 15), bp = blockIdx.y * PH + (tid >> 4);
  const int b = g.N - 1 - bp;
  const bool in_grid = a < g.N && bp < half_rows;
  // The job's header comes from the kernel parameters (constant bank), so the program words and the ray's cached state on
  // the first-reflection surface are requested at once, before the CTA's only barrier; the job's float constants (pixel
  // mapping, channel weights) are not needed until a ray lands.
  const unsigned head = heads.h[blockIdx.z];
  const int slot = (int)(head & 0xffffu) - 1, j_first = (int)((head >> 16) & 31u), n_steps = (int)(head >> 21);
  prefetch_job_constants(J);
  Staged staged;
  if (STAGED) staged = stage_issue<T, BT>(progs + (size_t)blockIdx.z * LFB_MAX_STEPS, n_steps, tid);
  typename io::Raw raw;
  const bool cached = slot >= 0 && in_grid;  // (slot: uniform per CTA) the state ON the first-reflection surface comes from the prefix cache
  if (cached) {
    const size_t hr = (size_t)g.half_rays;
    raw = io::issue(g.prefix + ((size_t)(slot * g.n_surf + j_first) * io::kParts) * hr + ((size_t)bp * g.N + a), hr);
  }
  if (STAGED) {
    stage_commit<T, BT>(s_prog_mem, staged, n_steps, tid);
    __syncthreads();  // the only CTA-wide barrier; nothing above waits for the state loads
  }
  const StepFetch<T, STAGED> s_prog = {STAGED ? s_prog_mem : progs + (size_t)blockIdx.z * LFB_MAX_STEPS};
  const MaskGeom K(g, tex);
  const JobC<T> JC(J);
  RayState<T> r;
  RayDiag o;
  bool alive = false;
  unsigned n_exec = 0;
  int s = 0;
  if (slot >= 0) {
    if (cached) alive = io::take(raw, r);
    if (alive) alive = interact<T, true, false>(s_prog[0], K, g, r, o);
    s = 1;
  } else if (in_grid) {  // no cache (the direct path, or the cache is over its memory budget): from the entrance grid
    T x, y;
    grid_point<T>(g, a, b, x, y);
    start_ray<T>(r, x, y, JC);
    alive = true;
    const int jc = J.i < 0 ? -1 : j_first;  // the step at which a cached state would have been loaded
#pragma unroll 1
    for (; alive && s <= jc; s++) {
      const StepT<T>& S = s_prog[s];
      alive = propagate<T, false>(S, r, o);
      if (STATS) n_exec++;
      if (alive && s == jc) io::canon(r);
      if (alive) alive = interact<T, true, false>(S, K, g, r, o);
    }
  }
#pragma unroll 1
  for (; alive && s < n_steps; s++) {
    const StepT<T>& S = s_prog[s];
    alive = propagate<T, false>(S, r, o);
    if (STATS) n_exec++;
    if (alive) alive = interact<T, true, false>(S, K, g, r, o);
  }
  const WarpSplat WS(g, J, s_tile + (tid >> 5) * (kWarpTilePx * 3), accum);
  const bool landed = warp_land<T, CALL>(WS, JC, alive, r.ox, r.oy, r.w * r.ma, r.w * r.mb, b != bp, lane);
  if (STATS) flush_stats(g.stats, n_exec, in_grid ? 1u : 0u, landed ? 1u : 0u);
}

// ---------------------------------------------------------------------------------------------------------------
// GHOST FAMILIES.  After the forward sweep, the ghosts (i, j) of one (light, wavelength) with the same first reflection j
// also share the BACKWARD sweep j-1, j-2, ... : ghost (i, j) leaves it at surface i.  One thread follows the whole
// family: it starts ON surface j (prefix cache), reflects, and walks backward; at every candidate surface k it forks --
// a copy of the ray reflects at k and runs the forward program k+1 .. n-1 to the sensor, where the warp splats it -- and
// then continues backward through k.  28 ghost jobs per (light, wavelength) become 7 family jobs: the per-warp fixed
// costs (launch prologue, state load, program staging) are paid once per family and the backward sweeps are not
// repeated.  The arithmetic along every ghost path is unchanged, so the frame is bit-identical to the ghost kernel's.
//
// Family program (shared memory): [0] reflect at j; then for k = j-1 .. 0 two steps: (backward step at k: refraction or
// the stop plane, fork step: reflection at k seen from behind, op < 0 when ghost (k, j) is not wanted).  The forward
// program of the slot (refractions 0 .. n-1 + sensor) supplies the suffix k+1 .. n of every fork.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int MINB, int BT, bool STATS>
__global__ void __launch_bounds__(BT, MINB) family_kernel(const Job* __restrict__ fams, const StepT<T>* __restrict__ fam_progs,
                                                          const __grid_constant__ JobHeads heads, const Job* __restrict__ slots,
                                                          const StepT<T>* __restrict__ slot_progs, FrameGeom g, const float* __restrict__ tex,
                                                          unsigned long long* __restrict__ accum) {
  typedef PrefixIO<T> io;
  constexpr int PH = BT / 16;
  __shared__ __align__(16) StepT<T> s_fam[2 * LFB_MAX_SURFACES + 2];
  __shared__ __align__(16) StepT<T> s_fwd[LFB_MAX_SURFACES + 2];
  __shared__ unsigned long long s_tile[(BT / 32) * kWarpTilePx * 3];

  const int half_rows = (g.N + 1) / 2;
  const Job& J = fams[blockIdx.z];
  const unsigned head = heads.h[blockIdx.z];  // from the kernel parameters: no global load before the prologue's loads
  const int slot = (int)(head & 0xffffu) - 1, j = (int)((head >> 16) & 31u), n_fam = (int)(head >> 21);
  prefetch_job_constants(J);
  const int n_fwd = g.n_surf + 1;  // forward refractions 0 .. n-1 and the sensor
  const int tid = threadIdx.x, lane = tid & 31;
  const int a = blockIdx.x * 16 + (tid & 15), bp = blockIdx.y * PH + (tid >> 4);
  const int b = g.N - 1 - bp;
  const bool in_grid = a < g.N && bp < half_rows;
  // both programs and the ray's cached state are requested at once, before the barrier
  const int n_stage = n_fam;
  const Staged st_fam = stage_issue<T, BT>(fam_progs + (size_t)blockIdx.z * LFB_MAX_STEPS, n_stage, tid);
  const Staged st_fwd = stage_issue<T, BT>(slot_progs + (size_t)slot * LFB_MAX_STEPS, n_fwd, tid);
  typename io::Raw raw;
  if (in_grid) {
    const size_t hr = (size_t)g.half_rays;
    raw = io::issue(g.prefix + ((size_t)(slot * g.n_surf + j) * io::kParts) * hr + ((size_t)bp * g.N + a), hr);
  }
  stage_commit<T, BT>(s_fam, st_fam, n_stage, tid);
  stage_commit<T, BT>(s_fwd, st_fwd, n_fwd, tid);
  __syncthreads();  // the only CTA-wide barrier; nothing above waits for the state loads

  const MaskGeom K(g, tex);
  const JobC<T> JC(J);
  const WarpSplat WS(g, J, s_tile + (tid >> 5) * (kWarpTilePx * 3), accum);
  RayState<T> r;
  RayDiag o;
  bool alive = false;
  unsigned n_exec = 0, n_landed = 0;
  if (in_grid) alive = io::take(raw, r);
  if (alive) alive = interact<T, true, false>(s_fam[0], K, g, r, o);  // the first reflection, at j
  const int n_back = (n_fam - 1) >> 1;  // backward surfaces in this job's program: j-1 .. down to its lowest fork
#pragma unroll 1
  for (int e = 0; e < n_back && __any_sync(0xffffffffu, alive); e++) {
    const int k = j - 1 - e;
    const StepT<T>& Sb = s_fam[1 + 2 * e];
    const StepT<T>& Sf = s_fam[2 + 2 * e];
    if (alive) { alive = propagate<T, false>(Sb, r, o); if (STATS) n_exec++; }  // onto surface k, travelling backward
    if (Sf.op >= 0) {  // ghost (k, j): fork a copy that reflects here and runs forward to the sensor
      RayState<T> q = r;
      bool a2 = alive;
      if (a2) a2 = interact<T, true, false>(Sf, K, g, q, o);
#pragma unroll 1
      for (int s = k + 1; s < n_fwd; s++) {
        if (a2) { a2 = propagate<T, false>(s_fwd[s], q, o); if (STATS) n_exec++; }
        if (a2) a2 = interact<T, true, false>(s_fwd[s], K, g, q, o);
      }
      if (warp_land<T>(WS, JC, a2, q.ox, q.oy, q.w * q.ma, q.w * q.mb, b != bp, lane)) n_landed++;
    }
    if (alive) alive = interact<T, true, false>(Sb, K, g, r, o);  // on through surface k (or the stop's mask)
  }
  if (STATS) flush_stats(g.stats, n_exec, in_grid ? 1u : 0u, n_landed);
}

// Parity instrument for the same trace code: one record per ray (flags, positions, weight).
template <typename T>
__global__ void __launch_bounds__(256) dump_kernel(const Job* __restrict__ job, const StepT<T>* __restrict__ prog, FrameGeom g,
                                                   const float* __restrict__ tex, lfb_ray_hit* __restrict__ out) {
  __shared__ __align__(16) StepT<T> s_prog[LFB_MAX_STEPS];
  const Job& J = *job;
  const int n_steps = J.n_steps;
  stage_program<T>(s_prog, prog, n_steps, threadIdx.x, 256);
  __syncthreads();
  const int tile = blockIdx.x;
  const int a = (tile % g.tiles_x) * 16 + (threadIdx.x & 15);
  const int b = (tile / g.tiles_x) * 16 + (threadIdx.x >> 4);
  if (a >= g.N || b >= g.N) return;
  const MaskGeom K(g, tex);
  const JobC<T> JC(J);
  RayState<T> r;
  {
    T x, y;
    grid_point<T>(g, a, b, x, y);
    start_ray<T>(r, x, y, JC);
  }
  RayDiag o;
  o.flags = 0; o.xa = o.ya = CUDART_NAN;
  bool alive = true;
  for (int s = 0; alive && s < n_steps; s++) {
    alive = propagate<T, true>(s_prog[s], r, o);
    if (alive) alive = interact<T, false, true>(s_prog[s], K, g, r, o);
  }
  lfb_ray_hit rec;
  rec.x_ap = o.xa; rec.y_ap = o.ya; rec.flags = o.flags; rec.pad = 0;
  if (alive) {
    const T X = -JC.ppu * r.ox, Y = JC.ppu * r.oy;
    const T px = JC.sx + M<T>::fma(X, JC.cs, -M<T>::mul_rn(Y, JC.sn));
    const T py = JC.sy + M<T>::fma(X, JC.sn, M<T>::mul_rn(Y, JC.cs));
    rec.x_s = (double)r.ox; rec.y_s = (double)r.oy; rec.weight = (double)(r.w * r.ma); rec.px = (double)px; rec.py = (double)py;
    const T fx = M<T>::floor(px), fy = M<T>::floor(py);
    if (!(fx >= (T)0 && fx < (T)g.W && fy >= (T)0 && fy < (T)g.H)) rec.flags |= LFB_RAY_OFF_SENSOR;
  } else {
    rec.x_s = rec.y_s = rec.px = rec.py = CUDART_NAN;
    rec.weight = 0.0;
  }
  out[(size_t)b * g.N + a] = rec;
}

// ---------------------------------------------------------------------------------------------------------------
// host launchers (instantiated per geometry type by exact_f32.cu / exact_f64.cu)
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct Tune;  // CTAs per SM the register allocation targets: default / alternative
// (measured on B200, tools/kernel_ab.py: A is the faster of each pair at cfg2 / cfg3)
template <> struct Tune<float> { static constexpr int kGhostA = 12, kGhostB = 10, kFamilyA = 8, kFamilyB = 10; };
template <> struct Tune<double> { static constexpr int kGhostA = 8, kGhostB = 6, kFamilyA = 6, kFamilyB = 5; };

template <typename T>
cudaError_t launch_prefix_t(const Job* slots, const StepT<T>* progs, int n_slots, const FrameGeom& g, const float* tex, float4* prefix,
                            unsigned long long* accum_for_direct, bool stats, cudaStream_t s) {
  if (n_slots <= 0) return cudaSuccess;
  const dim3 nb((g.N + 15) / 16, ((g.N + 1) / 2 + 15) / 16, n_slots);  // (patch column, patch row, slot)
  if (nb.y > 65535u || nb.z > 65535u) return cudaErrorInvalidConfiguration;
  if (stats) prefix_kernel<T, true><<<nb, kPrefixThreads, 0, s>>>(slots, progs, g, tex, prefix, accum_for_direct);
  else prefix_kernel<T, false><<<nb, kPrefixThreads, 0, s>>>(slots, progs, g, tex, prefix, accum_for_direct);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_ghosts_t(const Job* jobs, const StepT<T>* progs, const unsigned* heads, int n_jobs, const FrameGeom& g, const float* tex,
                            unsigned long long* accum, int ctas_per_sm, bool stats, cudaStream_t s) {
  constexpr int BT = 128;
  JobHeads H;
  for (int z0 = 0; z0 < n_jobs; z0 += kHeadsPerLaunch) {  // kHeadsPerLaunch jobs per launch: their headers travel as kernel parameters
    const int nz = n_jobs - z0 < kHeadsPerLaunch ? n_jobs - z0 : kHeadsPerLaunch;
    memcpy(H.h, heads + z0, sizeof(unsigned) * (size_t)nz);
    const dim3 nb((g.N + 15) / 16, ((g.N + 1) / 2 + BT / 16 - 1) / (BT / 16), nz);  // (patch column, patch row, job)
    if (nb.y > 65535u) return cudaErrorInvalidConfiguration;
    const Job* J = jobs + z0;
    const StepT<T>* P = progs + (size_t)z0 * LFB_MAX_STEPS;
    if (stats) ghost_kernel<T, Tune<T>::kGhostA, BT, true><<<nb, BT, 0, s>>>(J, P, H, g, tex, accum);
    else if (ctas_per_sm == 1) ghost_kernel<T, Tune<T>::kGhostB, BT, false><<<nb, BT, 0, s>>>(J, P, H, g, tex, accum);
    else if (ctas_per_sm == 2) ghost_kernel<T, Tune<T>::kGhostA, BT, false, false><<<nb, BT, 0, s>>>(J, P, H, g, tex, accum);
    else if (ctas_per_sm == 3) ghost_kernel<T, Tune<T>::kGhostA, BT, false, true, true><<<nb, BT, 0, s>>>(J, P, H, g, tex, accum);
    else ghost_kernel<T, Tune<T>::kGhostA, BT, false><<<nb, BT, 0, s>>>(J, P, H, g, tex, accum);
  }
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_families_t(const Job* fams, const StepT<T>* fam_progs, const unsigned* heads, int n_fams, const Job* slots,
                              const StepT<T>* slot_progs, const FrameGeom& g, const float* tex, unsigned long long* accum, int ctas_per_sm,
                              bool stats, cudaStream_t s) {
  constexpr int BT = 128;
  JobHeads H;
  for (int z0 = 0; z0 < n_fams; z0 += kHeadsPerLaunch) {
    const int nz = n_fams - z0 < kHeadsPerLaunch ? n_fams - z0 : kHeadsPerLaunch;
    memcpy(H.h, heads + z0, sizeof(unsigned) * (size_t)nz);
    const dim3 nb((g.N + 15) / 16, ((g.N + 1) / 2 + BT / 16 - 1) / (BT / 16), nz);  // (patch column, patch row, family)
    if (nb.y > 65535u) return cudaErrorInvalidConfiguration;
    const Job* J = fams + z0;
    const StepT<T>* P = fam_progs + (size_t)z0 * LFB_MAX_STEPS;
    if (stats) family_kernel<T, Tune<T>::kFamilyA, BT, true><<<nb, BT, 0, s>>>(J, P, H, slots, slot_progs, g, tex, accum);
    else if (ctas_per_sm == 1) family_kernel<T, Tune<T>::kFamilyB, BT, false><<<nb, BT, 0, s>>>(J, P, H, slots, slot_progs, g, tex, accum);
    else family_kernel<T, Tune<T>::kFamilyA, BT, false><<<nb, BT, 0, s>>>(J, P, H, slots, slot_progs, g, tex, accum);
  }
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_dump_t(const Job* job, const StepT<T>* prog, const FrameGeom& g, const float* tex, lfb_ray_hit* out, cudaStream_t s) {
  dump_kernel<T><<<(unsigned)g.tiles_per_job, 256, 0, s>>>(job, prog, g, tex, out);
  return cudaGetLastError();
}

}  // namespace xt
}  // namespace lfb
