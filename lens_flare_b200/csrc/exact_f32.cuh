// exact_f32.cuh -- the throughput kernels of the ghost path: EXACT_GRID in FP32.
//
// There is no memory stream to speak of (rays are generated from their grid index, a ghost's whole description is ~1 KB,
// the 1 MB aperture mask is L2/L1 resident), so the design minimises instructions and stalls per ray (ncu, profiles/):
//
//  * STEP PROGRAM.  The host flattens each ghost (i, j, lambda) into a straight list of steps (refract / reflect / stop
//    plane / sensor plane) with every ray-independent quantity already computed: curvature, vertex shift, clear radius^2,
//    n0/n2 and its square, the reflectance table of the interface.  A CTA stages its ghost's program in shared memory
//    once; the per-ray loop has no "which surface / which direction / which glass" logic and no divisions by constants.
//  * UNIFIED STEP.  Planes are c = 0 surfaces with an infinite clear radius; a missed surface or a total internal
//    reflection turns the state into NaNs that the next clear-radius test catches: ONE death test per step.
//  * FAST SCALAR MATH.  rcp.approx / sqrt.approx (one MUFU each) instead of the IEEE division and square root sequences
//    (MUFU + FCHK + slow-path branch + Newton FFMAs).
//  * TABULATED REFLECTANCES (v4).  Fresnel / quarter-wave-film reflectance of every (wavelength, surface, direction) is
//    tabulated on the host in double over the cosine in the rarer medium (where it is analytic): ~10 instructions and one
//    8-byte load per surface instead of ~60, cheap enough to carry the weight along in a single trace.
//  * MIRROR SYMMETRY (v3).  A distant light's bundle and the lens are symmetric about the meridional plane: ray (x, -y) is
//    the mirror image of ray (x, y).  One trace serves both; only the (asymmetric) aperture mask is looked up twice.
//  * PREFIX CACHE (v5).  The forward sweep 0 .. j-1 shared by all ghosts of a (light, wavelength) is traced once
//    (prefix_kernel) and the ray states on every surface cached in L2; ghost jobs start ON their first-reflection surface.
//  * WARP-AUTONOMOUS SPLAT (v6).  A warp owns its 32 ray pairs from trace to flush: survivors stay in registers, deposits
//    go to a 64-pixel u64 fixed-point tile per warp in shared memory and each touched pixel is flushed with one global
//    atomic (direct global atomics when the footprint exceeds the tile).  The only CTA barrier is the one after staging the
//    program.  Integer accumulation keeps the frame bit-stable for any schedule, kernel generation or GPU count.
//
// Kernel generations, all selectable and tested against each other and the oracle (tests/test_gpu_parity.py):
//   exact_splat_kernel   v3  two passes (geometry, then survivors with closed-form Fresnel)     LFB_EXACT_WEIGHTS=closed
//   exact_splat1_kernel  v4  one pass with tabulated reflectances, CTA-level queue and tile      LFB_EXACT_PREFIX=0
//   exact_splat2_kernel  v5  + prefix cache                                                      LFB_EXACT_WARP=0
//   exact_splat3_kernel  v6  + warp-autonomous splat (default for small frames)                  LFB_EXACT_FAMILY=0
//   exact_splat4_kernel  v6i v6 with two ray pairs per thread in lockstep (opt-in: +1 % only)     LFB_EXACT_ILP=2
//   exact_family_kernel  v7  ghost families: one thread per (ray, first reflection), forks at every second reflection
//                            (default from 16 384 family CTAs up; LFB_EXACT_FAMILY=1 forces it)
// Parity: per-ray and image tolerances against the double-precision oracle are in tests/test_gpu_parity.py; the FP64
// kernels in ghost_grid_impl.cuh remain the bit-exact instruments.
#pragma once
#include <math_constants.h>

#include "lfb_internal.h"

namespace lfb {
namespace xf32 {

__device__ __forceinline__ float frcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fsqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Reflectance of one interface (n0 -> n2, incidence cosine c0, transmission cosine c2): bare Fresnel,
// or the exact single-layer (Airy) film with index n1 and quarter-wave thickness at lambda0
// (phase = pi * lambda0 / lambda), polarisation-averaged.  One reciprocal.
__device__ __forceinline__ float reflectance(const Step& S, float c0, float c2) {
  const float n0 = S.n0, n2 = S.n2;
  if (S.n1 == 0.f) {  // bare
    const float a = n0 * c0, b = n2 * c2, c = n2 * c0, d = n0 * c2;
    const float u = (a - b) * (c + d), v = (c - d) * (a + b), q = (a + b) * (c + d);
    return 0.5f * (u * u + v * v) * frcp(q * q);
  }
  const float n1 = S.n1;
  const float k1 = fmaf(-S.e1sq, fmaf(-c0, c0, 1.f), 1.f);  // 1 - (n0/n1)^2 sin^2
  if (k1 < 0.f) return 1.f;
  const float c1 = fsqrt(k1);
  const float cd2 = 2.f * __cosf(S.phase * c1);
  // s: r01 = (A-B)/(A+B), r12 = (B-C)/(B+C) with A = n0 c0, B = n1 c1, C = n2 c2
  const float A = n0 * c0, B = n1 * c1, C = n2 * c2;
  const float a1 = A - B, b1 = A + B, a2 = B - C, b2 = B + C;
  const float Ps = a1 * b2, Qs = a2 * b1, Ss = b1 * b2, Ts = a1 * a2;
  const float ms = Ps * Qs * cd2;  // = Ss*Ts*cd2
  const float Ns = fmaf(Ps, Ps, fmaf(Qs, Qs, ms)), Ds = fmaf(Ss, Ss, fmaf(Ts, Ts, ms));
  // p: r01 = (n1 c0 - n0 c1)/(n1 c0 + n0 c1), r12 = (n2 c1 - n1 c2)/(n2 c1 + n1 c2)
  const float E = n1 * c0, F = n0 * c1, G = n2 * c1, H = n1 * c2;
  const float e1 = E - F, f1 = E + F, e2 = G - H, f2 = G + H;
  const float Pp = e1 * f2, Qp = e2 * f1, Sp = f1 * f2, Tp = e1 * e2;
  const float mp = Pp * Qp * cd2;
  const float Np = fmaf(Pp, Pp, fmaf(Qp, Qp, mp)), Dp = fmaf(Sp, Sp, fmaf(Tp, Tp, mp));
  return 0.5f * fmaf(Ns, Dp, Np * Ds) * frcp(Ds * Dp);
}

struct MaskGeom {
  const float* tex;
  int tw, th;
  float su, sv, ou, ov;  // u = xa*su + ou ; v = ya*sv + ov
};

__device__ __forceinline__ float mask_lookup(const MaskGeom& M, float xa, float ya) {
  const float fu = floorf(fmaf(xa, M.su, M.ou)), fv = floorf(fmaf(ya, M.sv, M.ov));
  if (!(fu >= 0.f && fu < (float)M.tw && fv >= 0.f && fv < (float)M.th)) return 0.f;
  return __ldg(M.tex + ((unsigned)(int)fv * (unsigned)M.tw + (unsigned)(int)fu));  // in range, checked above
}

// What one traced ray yields.  The bundle of a distant light is mirror-symmetric about the meridional
// (x-z) plane and so is the lens: ray (x, -y) is the mirror image of ray (x, y) -- same path, same
// incidence angles, same Fresnel weights, sensor point (xs, -ys).  Only the aperture mask is not
// symmetric, so ONE trace serves both rays and carries two mask weights: wa for (x, y), wb for (x, -y).
struct RayOut {
  float xs, ys, xa, ya;  // sensor point and last stop crossing of the traced ray (x, y)
  float wa, wb;          // weights of the ray and of its mirror image (Fresnel product x mask products)
  unsigned flags;        // FLAGS variant only
};

// Reflectance from the interface's table: R as a function of v = the cosine of the ray's angle in the RARER of the two
// media (in which R is analytic: R -> 1 linearly as v -> 0, i.e. at grazing incidence or at the critical angle), 1024
// intervals, linear interpolation of (R_i, R_{i+1} - R_i) pairs built on the host in double from the exact formulas
// (one 8-byte load, ~10 instructions, instead of ~60 for the closed-form coated interface).
__device__ __forceinline__ float reflectance_lut(const float2* __restrict__ lut, int table, float v) {
  const float t = fminf(fmaxf(v, 0.f), 1.f) * (float)kLutSize;  // fmaxf(NaN, 0) = 0: beyond the critical angle R = 1
  const int i = min((int)t, kLutSize - 1);
  const float2 e = __ldg(lut + ((unsigned)table * (unsigned)kLutSize + (unsigned)i));  // unsigned: one IMAD.WIDE.U32, no sign extension
  return fmaf(t - (float)i, e.y, e.x);
}

// Trace one ray through a step program.
//   WEIGHTS    0: geometry + mask only; 1: closed-form Fresnel / coating weights; 2: tabulated weights
//   MIRROR     also look the mask up at (xa, -ya) for the mirror-image ray
//   FLAGS      parity-instrument variant: classify the death (missed / vignetted / TIR), and let rays the
//              mask stopped continue with weight 0 so that their positions stay comparable with the oracle
// Every step runs the same straight-line code (planes are c = 0 surfaces with an infinite clear radius); a
// missed surface or a total internal reflection turns the state into NaNs, which the next clear-radius
// test catches, so the throughput variants carry ONE death test per step.
struct RayState {
  float ox, oy, oz, dx, dy, dz;  // position relative to the current surface's vertex; unit direction
  float w, ma, mb;               // Fresnel product; mask products of the ray and of its mirror image
};

// First half of a step: move the ray onto the step's surface (sphere or plane through its vertex) and test the clear
// aperture.  False = the ray is dead (missed, vignetted, or NaN from a total internal reflection upstream).
template <bool FLAGS>
__device__ __forceinline__ bool propagate(const Step& S, RayState& r, RayOut& o) {
  const float c = S.c;
  const float pz = r.oz + S.dz;
  const float pd = fmaf(r.ox, r.dx, fmaf(r.oy, r.dy, pz * r.dz));
  const float pp = fmaf(r.ox, r.ox, fmaf(r.oy, r.oy, pz * pz));
  const float B = fmaf(c, pd, -r.dz);
  const float Cq = fmaf(c, pp, -2.f * pz);
  const float disc = fmaf(B, B, -c * Cq);
  if (FLAGS && disc < 0.f) { o.flags |= LFB_RAY_MISSED; return false; }
  const float t = -Cq * frcp(B + copysignf(fsqrt(disc), B));
  r.ox = fmaf(t, r.dx, r.ox); r.oy = fmaf(t, r.dy, r.oy); r.oz = fmaf(t, r.dz, pz);
  if (!(fmaf(r.ox, r.ox, r.oy * r.oy) <= S.semi2)) {  // outside the clear aperture, or NaN from a miss / TIR upstream
    if (FLAGS) o.flags |= LFB_RAY_VIGNETTED;
    return false;
  }
  return true;
}

// Second half: what the surface does to the ray (mask lookup at the stop; refraction or reflection + weight elsewhere).
template <int WEIGHTS, bool MIRROR, bool FLAGS>
__device__ __forceinline__ bool interact(const Step& S, const MaskGeom& M, const float2* __restrict__ lut, RayState& r, RayOut& o) {
  const int op = S.op;
  if (op >= STEP_PASS) {  // no change of direction: identical media, the stop, the sensor
    if (op == STEP_STOP) {
      const float m = mask_lookup(M, r.ox, r.oy);
      r.ma *= m;
      if (MIRROR) r.mb *= mask_lookup(M, r.ox, -r.oy);
      if (FLAGS) { o.xa = r.ox; o.ya = r.oy; if (m == 0.f) o.flags |= LFB_RAY_STOPPED; }
      else if (MIRROR ? (r.ma == 0.f && r.mb == 0.f) : (r.ma == 0.f)) return false;
    }
    return true;
  }
  const float c = S.c, eta = S.eta;
  const float nx = -c * r.ox, ny = -c * r.oy, nz = fmaf(-c, r.oz, 1.f);
  const float nd = fmaf(nx, r.dx, fmaf(ny, r.dy, nz * r.dz));
  const float c0 = fabsf(nd);
  const float s2 = fmaf(-c0, c0, 1.f);
  const float k2 = fmaf(-S.eta2, s2, 1.f);
  const float c2 = fsqrt(k2);  // NaN beyond the critical angle
  const bool refl = op == STEP_REFLECT;
  if (FLAGS && !refl && k2 < 0.f) { o.flags |= LFB_RAY_TIR; return false; }
  // refract: d' = eta d + (eta c0 - c2) N with N = -sign(nd) n the normal facing the ray;  reflect: d' = d - 2 nd n
  const float g = __int_as_float(__float_as_int(fmaf(eta, c0, -c2)) ^ (~__float_as_int(nd) & 0x80000000));
  const float alpha = refl ? 1.f : eta, beta = refl ? -2.f * nd : g;
  r.dx = fmaf(alpha, r.dx, beta * nx); r.dy = fmaf(alpha, r.dy, beta * ny); r.dz = fmaf(alpha, r.dz, beta * nz);
  if (WEIGHTS == 1) {
    const float R = (k2 < 0.f) ? 1.f : reflectance(S, c0, c2);
    r.w *= refl ? R : 1.f - R;
  } else if (WEIGHTS == 2) {
    const float R = reflectance_lut(lut, S.lut, eta > 1.f ? c2 : c0);
    r.w *= refl ? R : 1.f - R;
  }
  return true;
}

// Run a step program on a ray state.  ON_SURFACE: the state already sits on the first step's surface (it was loaded
// from the prefix buffer), so step 0 only interacts.
template <int WEIGHTS, bool MIRROR, bool FLAGS, bool ON_SURFACE>
__device__ __forceinline__ bool run_program(const Step* __restrict__ prog, int n_steps, const MaskGeom& M, const float2* __restrict__ lut,
                                            RayState& r, RayOut& o) {
  int s = 0;
  if (ON_SURFACE) {
    if (!interact<WEIGHTS, MIRROR, FLAGS>(prog[0], M, lut, r, o)) return false;
    s = 1;
  }
#pragma unroll 1
  for (; s < n_steps; s++) {
    const Step& S = prog[s];
    if (!propagate<FLAGS>(S, r, o)) return false;
    if (!interact<WEIGHTS, MIRROR, FLAGS>(S, M, lut, r, o)) return false;
  }
  o.xs = r.ox; o.ys = r.oy; o.wa = r.w * r.ma; o.wb = r.w * r.mb;
  return true;
}

// Entrance ray of grid point (x, y).  Directional light (inv_d = 0): the bundle's common direction.  Point light at
// -D (sin t, 0, cos t): along v = (x/D + sin t, y/D, cos t), carrying the irradiance at the entrance point relative to the
// vertex, |v|^-3 (lfb_light.distance).  The light lies in the plane y = 0, so the mirror pair (x, -y) stays a mirror pair.
__device__ __forceinline__ void start_ray(RayState& r, float x, float y, float sin_t, float cos_t, float inv_d) {
  r.ox = x; r.oy = y; r.oz = 0.f; r.dx = sin_t; r.dy = 0.f; r.dz = cos_t; r.w = 1.f; r.ma = 1.f; r.mb = 1.f;
  if (inv_d != 0.f) {  // uniform per job
    const float vx = fmaf(x, inv_d, sin_t), vy = __fmul_rn(y, inv_d);
    const float q = fmaf(vx, vx, fmaf(vy, vy, __fmul_rn(cos_t, cos_t)));
    const float rl = rsqrtf(q);
    r.dx = __fmul_rn(vx, rl); r.dy = __fmul_rn(vy, rl); r.dz = __fmul_rn(cos_t, rl);
    r.w = __fmul_rn(__fmul_rn(rl, rl), rl);
  }
}

template <int WEIGHTS, bool MIRROR, bool FLAGS>
__device__ __forceinline__ bool trace(const Step* __restrict__ prog, int n_steps, const MaskGeom& M, const float2* __restrict__ lut,
                                      float x, float y, float sin_t, float cos_t, float inv_d, RayOut& o) {
  RayState r;
  start_ray(r, x, y, sin_t, cos_t, inv_d);
  if (FLAGS) { o.flags = 0; o.xa = o.ya = CUDART_NAN_F; }
  return run_program<WEIGHTS, MIRROR, FLAGS, false>(prog, n_steps, M, lut, r, o);
}

struct PixMap {
  float sx, sy, cs, sn, ppu;
};
__device__ __forceinline__ void to_pixel(const PixMap& P, float xs, float ys, float& px, float& py) {
  const float X = -P.ppu * xs, Y = P.ppu * ys;
  px = P.sx + fmaf(X, P.cs, -__fmul_rn(Y, P.sn));
  py = P.sy + fmaf(X, P.sn, __fmul_rn(Y, P.cs));
}

constexpr int kThreads = 256;
constexpr int kTilePx = 256;  // shared-memory sensor tile capacity (pixels)

// Footprint of one splat (bilinear: 2x2 taps around (px, py) - 0.5; nearest: the pixel holding (px, py)),
// clipped to the sensor.  False when nothing lands (or px/py is NaN).
__device__ __forceinline__ bool footprint(bool bilinear, float px, float py, int W, int H, int& x0, int& y0, int& x1, int& y1) {
  if (bilinear) {
    const float fx = floorf(px - 0.5f), fy = floorf(py - 0.5f);
    if (!(fx >= -1.f && fx < (float)W && fy >= -1.f && fy < (float)H)) return false;
    x0 = max((int)fx, 0); y0 = max((int)fy, 0); x1 = min((int)fx + 1, W - 1); y1 = min((int)fy + 1, H - 1);
  } else {
    const float fx = floorf(px), fy = floorf(py);
    if (!(fx >= 0.f && fx < (float)W && fy >= 0.f && fy < (float)H)) return false;
    x0 = x1 = (int)fx; y0 = y1 = (int)fy;
  }
  return true;
}

struct SplatCtx {
  unsigned long long* tile;   // shared-memory tile (or nullptr: straight to global)
  unsigned long long* accum;  // global sensor accumulators
  int tx0, ty0, tw, W, H;
  float ch0, ch1, ch2;        // radiance * rgb_weight * ray area * 2^bits
  bool bilinear;
};

__device__ __forceinline__ void splat(const SplatCtx& C, float px, float py, float w) {
  int ix, iy, ntap;
  float wt[4];
  if (C.bilinear) {
    const float qx = __fsub_rn(px, 0.5f), qy = __fsub_rn(py, 0.5f);
    const float fx0 = floorf(qx), fy0 = floorf(qy);
    const float fx = __fsub_rn(qx, fx0), fy = __fsub_rn(qy, fy0);
    ix = (int)fx0; iy = (int)fy0; ntap = 4;
    // explicit roundings: no FMA contraction, so every instantiation of the kernel produces the same bits
    const float gx = __fsub_rn(1.f, fx), gy = __fsub_rn(1.f, fy);
    wt[0] = __fmul_rn(w, __fmul_rn(gx, gy)); wt[1] = __fmul_rn(w, __fmul_rn(fx, gy));
    wt[2] = __fmul_rn(w, __fmul_rn(gx, fy)); wt[3] = __fmul_rn(w, __fmul_rn(fx, fy));
  } else {
    ix = (int)floorf(px); iy = (int)floorf(py); ntap = 1;
    wt[0] = w;
  }
#pragma unroll
  for (int t = 0; t < 4; t++) {
    if (t >= ntap) break;
    const int jx = ix + (t & 1), jy = iy + (t >> 1);
    if (jx < 0 || jx >= C.W || jy < 0 || jy >= C.H) continue;
    const long long q0 = __float2ll_rn(wt[t] * C.ch0), q1 = __float2ll_rn(wt[t] * C.ch1), q2 = __float2ll_rn(wt[t] * C.ch2);
    unsigned long long* dst = C.tile ? C.tile + 3 * ((jy - C.ty0) * C.tw + (jx - C.tx0)) : C.accum + 3 * ((size_t)jx + (size_t)jy * C.W);
    if (q0) atomicAdd(dst + 0, (unsigned long long)q0);
    if (q1) atomicAdd(dst + 1, (unsigned long long)q1);
    if (q2) atomicAdd(dst + 2, (unsigned long long)q2);
  }
}

// One CTA = a (16*RX) x (16*RY) patch of the UPPER HALF of one ghost's ray grid (rows with y >= 0; each
// traced ray also stands for its mirror image in the lower half); RPT = RX*RY rays per thread in pass 1.
template <int RX, int RY, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) exact_splat_kernel(const Job* __restrict__ jobs, const Step* __restrict__ progs,
                                                               FrameGeom g, const float* __restrict__ tex,
                                                               unsigned long long* __restrict__ accum) {
  constexpr int RPT = RX * RY, PATCH = RPT * kThreads, PW = 16 * RX, PH = 16 * RY;
  __shared__ Step s_prog[LFB_MAX_STEPS];
  __shared__ unsigned long long s_tile[kTilePx * 3];
  __shared__ float4 s_qp[PATCH];         // survivor queue: (pxA, pyA, pxB, pyB)
  __shared__ unsigned short s_qid[PATCH];  // ray id within the patch | mirror-alive bits
  __shared__ int s_count, s_bbox[4];

  const int half_rows = (g.N + 1) / 2;  // rows b' = 0 .. half_rows-1 stand for b = N-1-b' (y > 0) and b' itself (y < 0)
  const int patches_x = (g.N + PW - 1) / PW;
  const int patches_per_job = patches_x * ((half_rows + PH - 1) / PH);
  const int job_id = blockIdx.x / patches_per_job;
  const int patch = blockIdx.x - job_id * patches_per_job;
  const int a0 = (patch % patches_x) * PW, b0 = (patch / patches_x) * PH;
  const Job& J = jobs[job_id];
  const int n_steps = J.n_steps;
  const int tid = threadIdx.x;

  {  // stage the ghost's program (n_steps * 48 B) and reset the CTA state
    const float4* src = reinterpret_cast<const float4*>(progs + (size_t)job_id * LFB_MAX_STEPS);
    float4* dst = reinterpret_cast<float4*>(s_prog);
    for (int q = tid; q < n_steps * 3; q += kThreads) dst[q] = __ldg(src + q);
    if (tid == 0) { s_count = 0; s_bbox[0] = s_bbox[1] = 0x7fffffff; s_bbox[2] = s_bbox[3] = -0x7fffffff; }
  }
  __syncthreads();

  const float P = g.P;
  const float cell = 2.f * P / (float)g.N;
  const float sin_t = (float)J.sin_t, cos_t = (float)J.cos_t, inv_d = (float)J.inv_dist;
  MaskGeom M;
  M.tex = tex; M.tw = g.tex_w; M.th = g.tex_h;
  {
    const float h = g.h_stop;
    M.su = 0.5f * (float)g.tex_w / h; M.ou = 0.5f * (float)g.tex_w;
    M.sv = -0.5f * (float)g.tex_h / h; M.ov = 0.5f * (float)g.tex_h;
  }
  PixMap PM;
  PM.sx = (float)J.sx; PM.sy = (float)J.sy; PM.cs = (float)J.cs; PM.sn = (float)J.sn; PM.ppu = (float)J.ppu;
  const bool bilinear = g.splat == LFB_SPLAT_BILINEAR;

  // ---- pass 1: geometry + mask only; queue the survivors with their sensor pixels -----------------
  int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
#pragma unroll 1
  for (int r = 0; r < RPT; r++) {
    const int la = (tid & 15) + 16 * (r % RX), lb = (tid >> 4) + 16 * (r / RX);
    const int a = a0 + la, bp = b0 + lb;  // bp: row of the half grid
    const int b = g.N - 1 - bp;
    unsigned bits = 0;  // 1: the traced ray lands, 2: its mirror image lands
    float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a < g.N && bp < half_rows) {
      RayOut o;
      if (trace<0, true, false>(s_prog, n_steps, M, g.lut, fmaf((float)a + 0.5f, cell, -P), fmaf((float)b + 0.5f, cell, -P), sin_t, cos_t, inv_d, o)) {
        int x0, y0, x1, y1;
        if (o.wa > 0.f) {
          to_pixel(PM, o.xs, o.ys, pp.x, pp.y);
          if (footprint(bilinear, pp.x, pp.y, g.W, g.H, x0, y0, x1, y1)) {
            bits |= 1; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
          }
        }
        if (o.wb > 0.f && b != bp) {  // b == bp: the middle row of an odd grid is its own mirror image
          to_pixel(PM, o.xs, -o.ys, pp.z, pp.w);
          if (footprint(bilinear, pp.z, pp.w, g.W, g.H, x0, y0, x1, y1)) {
            bits |= 2; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
          }
        }
      }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, bits != 0);
    if (ballot) {
      const int lane = tid & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_count, __popc(ballot));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (bits) {
        const int slot = base + __popc(ballot & ((1u << lane) - 1u));
        s_qid[slot] = (unsigned short)((lb * PW + la) | (bits << 14));
        s_qp[slot] = pp;
      }
    }
  }
  bx0 = __reduce_min_sync(0xffffffffu, bx0); by0 = __reduce_min_sync(0xffffffffu, by0);
  bx1 = __reduce_max_sync(0xffffffffu, bx1); by1 = __reduce_max_sync(0xffffffffu, by1);
  if ((tid & 31) == 0 && bx1 >= bx0) {
    atomicMin(&s_bbox[0], bx0); atomicMin(&s_bbox[1], by0);
    atomicMax(&s_bbox[2], bx1); atomicMax(&s_bbox[3], by1);
  }
  __syncthreads();
  const int count = s_count;
  if (count == 0) return;
  if (tid == 0) grow_bbox(g.bbox, s_bbox[0], s_bbox[1], s_bbox[2], s_bbox[3]);
  SplatCtx C;
  C.tx0 = s_bbox[0]; C.ty0 = s_bbox[1];
  C.tw = s_bbox[2] - C.tx0 + 1;
  const int th = s_bbox[3] - C.ty0 + 1;
  const bool use_tile = C.tw * th <= kTilePx;
  C.tile = use_tile ? s_tile : nullptr;
  C.accum = accum; C.W = g.W; C.H = g.H; C.bilinear = bilinear;
  if (use_tile) {
    for (int q = tid; q < C.tw * th * 3; q += kThreads) s_tile[q] = 0ull;
    __syncthreads();
  }

  // ---- pass 2: survivors only, dense warps: full trace with Fresnel / coating weights ---------------
  const float scale = (float)g.fp_scale;
  C.ch0 = (float)J.chan[0] * scale; C.ch1 = (float)J.chan[1] * scale; C.ch2 = (float)J.chan[2] * scale;
  for (int q = tid; q < count; q += kThreads) {
    const unsigned id = s_qid[q];
    const int la = (id & 0x3fff) % PW, lb = (id & 0x3fff) / PW;
    const int a = a0 + la, b = g.N - 1 - (b0 + lb);
    RayOut o;
    if (!trace<1, true, false>(s_prog, n_steps, M, g.lut, fmaf((float)a + 0.5f, cell, -P), fmaf((float)b + 0.5f, cell, -P), sin_t, cos_t, inv_d, o)) continue;
    const float4 pp = s_qp[q];
    if ((id & (1u << 14)) && o.wa > 0.f) splat(C, pp.x, pp.y, o.wa);
    if ((id & (2u << 14)) && o.wb > 0.f) splat(C, pp.z, pp.w, o.wb);
  }
  if (!use_tile) return;
  __syncthreads();
  // ---- flush the tile: one global atomic per touched (pixel, channel) ------------------------------
  for (int q = tid; q < C.tw * th * 3; q += kThreads) {
    const unsigned long long v = s_tile[q];
    if (v) {
      const int p = q / 3, c = q - 3 * p;
      const int jy = C.ty0 + p / C.tw, jx = C.tx0 + (p - (p / C.tw) * C.tw);
      atomicAdd(accum + 3 * ((size_t)jx + (size_t)jy * g.W) + c, v);
    }
  }
}

// v4: ONE pass.  With tabulated reflectances the weight costs ~10 instructions per surface, so it is carried along in
// the single trace instead of re-tracing the survivors: the queue holds each surviving ray pair's sensor pixels AND
// weights, and the second phase only splats (dense warps) into the shared-memory tile.  Everything ray-independent
// (float copies of the job constants, cell size, mask mapping) is precomputed on the host: the per-CTA prologue is a few
// loads, which matters because a CTA lives for only ~10 surface steps per ray.
__device__ __forceinline__ void splat1(const SplatCtx& C, float px, float py, float w) {
  int ix, iy;
  float wt[4];
  if (C.bilinear) {
    const float qx = __fsub_rn(px, 0.5f), qy = __fsub_rn(py, 0.5f);
    const float fx0 = floorf(qx), fy0 = floorf(qy);
    const float fx = __fsub_rn(qx, fx0), fy = __fsub_rn(qy, fy0);
    ix = (int)fx0; iy = (int)fy0;
    // explicit roundings: no FMA contraction, so every instantiation of the kernel produces the same bits
    const float gx = __fsub_rn(1.f, fx), gy = __fsub_rn(1.f, fy);
    wt[0] = __fmul_rn(w, __fmul_rn(gx, gy)); wt[1] = __fmul_rn(w, __fmul_rn(fx, gy));
    wt[2] = __fmul_rn(w, __fmul_rn(gx, fy)); wt[3] = __fmul_rn(w, __fmul_rn(fx, fy));
  } else {
    ix = (int)floorf(px); iy = (int)floorf(py);
    wt[0] = w; wt[1] = wt[2] = wt[3] = 0.f;
  }
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const int jx = ix + (t & 1), jy = iy + (t >> 1);
    if (wt[t] == 0.f || jx < 0 || jx >= C.W || jy < 0 || jy >= C.H) continue;
    unsigned long long* dst = C.tile ? C.tile + 3 * ((jy - C.ty0) * C.tw + (jx - C.tx0)) : C.accum + 3 * ((size_t)jx + (size_t)jy * C.W);
    // channels a wavelength does not feed (RGB lenses: two of three) are skipped without converting anything
    if (C.ch0 != 0.f) { const long long q = __float2ll_rn(__fmul_rn(wt[t], C.ch0)); if (q) atomicAdd(dst + 0, (unsigned long long)q); }
    if (C.ch1 != 0.f) { const long long q = __float2ll_rn(__fmul_rn(wt[t], C.ch1)); if (q) atomicAdd(dst + 1, (unsigned long long)q); }
    if (C.ch2 != 0.f) { const long long q = __float2ll_rn(__fmul_rn(wt[t], C.ch2)); if (q) atomicAdd(dst + 2, (unsigned long long)q); }
  }
}

template <int RX, int RY, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) exact_splat1_kernel(const Job* __restrict__ jobs, const Step* __restrict__ progs,
                                                                      FrameGeom g, const float* __restrict__ tex,
                                                                      unsigned long long* __restrict__ accum) {
  constexpr int RPT = RX * RY, PATCH = RPT * kThreads, PW = 16 * RX, PH = 16 * RY;
  static_assert(kTilePx <= kThreads, "one thread per tile pixel in the zero / flush loops");
  __shared__ Step s_prog[LFB_MAX_STEPS];
  __shared__ unsigned long long s_tile[kTilePx * 3];
  __shared__ float4 s_qp[PATCH];  // survivor queue: (pxA, pyA, pxB, pyB)
  __shared__ float2 s_qw[PATCH];  //                 (wA, wB); 0 = that image of the pair does not land
  __shared__ int s_count, s_bbox[4];

  const int half_rows = (g.N + 1) / 2;
  const int patches_x = (g.N + PW - 1) / PW;
  const int patches_per_job = patches_x * ((half_rows + PH - 1) / PH);
  const int job_id = blockIdx.x / patches_per_job;
  const int patch = blockIdx.x - job_id * patches_per_job;
  const int a0 = (patch % patches_x) * PW, b0 = (patch / patches_x) * PH;
  const Job& J = jobs[job_id];
  const int n_steps = J.n_steps;
  const int tid = threadIdx.x;
  {
    const float4* src = reinterpret_cast<const float4*>(progs + (size_t)job_id * LFB_MAX_STEPS);
    float4* dst = reinterpret_cast<float4*>(s_prog);
    for (int q = tid; q < n_steps * 3; q += kThreads) dst[q] = __ldg(src + q);
    if (tid == 0) { s_count = 0; s_bbox[0] = s_bbox[1] = 0x7fffffff; s_bbox[2] = s_bbox[3] = -0x7fffffff; }
  }
  __syncthreads();

  const float P = g.P, cell = g.cell;
  const float sin_t = J.f_sin_t, cos_t = J.f_cos_t, inv_d = J.f_inv_dist;
  MaskGeom M;
  M.tex = tex; M.tw = g.tex_w; M.th = g.tex_h;
  M.su = g.mask_su; M.ou = g.mask_ou; M.sv = g.mask_sv; M.ov = g.mask_ov;
  PixMap PM;
  PM.sx = J.f_sx; PM.sy = J.f_sy; PM.cs = J.f_cs; PM.sn = J.f_sn; PM.ppu = J.f_ppu;
  const bool bilinear = g.splat == LFB_SPLAT_BILINEAR;

  // ---- phase 1: trace (geometry, mask, tabulated weights); queue what lands ------------------------
  int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
#pragma unroll 1
  for (int r = 0; r < RPT; r++) {
    const int la = (tid & 15) + 16 * (r % RX), lb = (tid >> 4) + 16 * (r / RX);
    const int a = a0 + la, bp = b0 + lb;
    const int b = g.N - 1 - bp;
    float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 ww = make_float2(0.f, 0.f);
    if (a < g.N && bp < half_rows) {
      RayOut o;
      if (trace<2, true, false>(s_prog, n_steps, M, g.lut, fmaf((float)a + 0.5f, cell, -P), fmaf((float)b + 0.5f, cell, -P), sin_t, cos_t, inv_d, o)) {
        int x0, y0, x1, y1;
        if (o.wa > 0.f) {
          to_pixel(PM, o.xs, o.ys, pp.x, pp.y);
          if (footprint(bilinear, pp.x, pp.y, g.W, g.H, x0, y0, x1, y1)) {
            ww.x = o.wa; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
          }
        }
        if (o.wb > 0.f && b != bp) {
          to_pixel(PM, o.xs, -o.ys, pp.z, pp.w);
          if (footprint(bilinear, pp.z, pp.w, g.W, g.H, x0, y0, x1, y1)) {
            ww.y = o.wb; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
          }
        }
      }
    }
    const bool lands = ww.x > 0.f || ww.y > 0.f;
    const unsigned ballot = __ballot_sync(0xffffffffu, lands);
    if (ballot) {
      const int lane = tid & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_count, __popc(ballot));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (lands) {
        const int slot = base + __popc(ballot & ((1u << lane) - 1u));
        s_qp[slot] = pp; s_qw[slot] = ww;
      }
    }
  }
  if (__any_sync(0xffffffffu, bx1 >= bx0)) {
    bx0 = __reduce_min_sync(0xffffffffu, bx0); by0 = __reduce_min_sync(0xffffffffu, by0);
    bx1 = __reduce_max_sync(0xffffffffu, bx1); by1 = __reduce_max_sync(0xffffffffu, by1);
    if ((tid & 31) == 0) {
      atomicMin(&s_bbox[0], bx0); atomicMin(&s_bbox[1], by0);
      atomicMax(&s_bbox[2], bx1); atomicMax(&s_bbox[3], by1);
    }
  }
  __syncthreads();
  const int count = s_count;
  if (count == 0) return;
  if (tid == 0) grow_bbox(g.bbox, s_bbox[0], s_bbox[1], s_bbox[2], s_bbox[3]);
  SplatCtx C;
  C.tx0 = s_bbox[0]; C.ty0 = s_bbox[1];
  C.tw = s_bbox[2] - C.tx0 + 1;
  const int area = C.tw * (s_bbox[3] - C.ty0 + 1);
  const bool use_tile = area <= kTilePx;
  C.tile = use_tile ? s_tile : nullptr;
  C.accum = accum; C.W = g.W; C.H = g.H; C.bilinear = bilinear;
  if (use_tile) {  // one thread per tile pixel
    if (tid < area) { s_tile[3 * tid] = 0ull; s_tile[3 * tid + 1] = 0ull; s_tile[3 * tid + 2] = 0ull; }
    __syncthreads();
  }
  // ---- phase 2: splat the queue (dense warps) -----------------------------------------------------------
  C.ch0 = J.f_chan[0]; C.ch1 = J.f_chan[1]; C.ch2 = J.f_chan[2];
  for (int q = tid; q < count; q += kThreads) {
    const float4 pp = s_qp[q];
    const float2 ww = s_qw[q];
    if (ww.x > 0.f) splat1(C, pp.x, pp.y, ww.x);
    if (ww.y > 0.f) splat1(C, pp.z, pp.w, ww.y);
  }
  if (!use_tile) return;
  __syncthreads();
  if (tid < area) {  // flush: one global atomic per touched (pixel, channel)
    const int jy = (int)(((float)tid + 0.5f) * frcp((float)C.tw));  // tid / tw, exact for these small integers
    const int jx = tid - jy * C.tw;
    unsigned long long* dst = accum + 3 * ((size_t)(C.tx0 + jx) + (size_t)(C.ty0 + jy) * g.W);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const unsigned long long v = s_tile[3 * tid + c];
      if (v) atomicAdd(dst + c, v);
    }
  }
}

constexpr int kWarpTilePx = 64;  // per-warp shared-memory sensor tile (pixels)

struct WarpSplat {  // what a warp needs to deposit its lanes' results
  unsigned long long* tile;   // this warp's shared-memory tile
  unsigned long long* accum;
  int* bbox;
  int W, H;
  float ch0, ch1, ch2;
  bool bilinear;
};

// Deposit one (ray, mirror image) result per lane; warp-collective (all 32 lanes call it).
__device__ __forceinline__ void warp_splat(const WarpSplat& S, const PixMap& PM, bool alive, const RayOut& o, bool has_mirror, int lane) {
  int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
  float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 ww = make_float2(0.f, 0.f);
  if (alive) {
    int x0, y0, x1, y1;
    if (o.wa > 0.f) {
      to_pixel(PM, o.xs, o.ys, pp.x, pp.y);
      if (footprint(S.bilinear, pp.x, pp.y, S.W, S.H, x0, y0, x1, y1)) {
        ww.x = o.wa; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
      }
    }
    if (o.wb > 0.f && has_mirror) {
      to_pixel(PM, o.xs, -o.ys, pp.z, pp.w);
      if (footprint(S.bilinear, pp.z, pp.w, S.W, S.H, x0, y0, x1, y1)) {
        ww.y = o.wb; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
      }
    }
  }
  const bool lands = ww.x > 0.f || ww.y > 0.f;
  if (!__any_sync(0xffffffffu, lands)) return;
  bx0 = __reduce_min_sync(0xffffffffu, bx0); by0 = __reduce_min_sync(0xffffffffu, by0);
  bx1 = __reduce_max_sync(0xffffffffu, bx1); by1 = __reduce_max_sync(0xffffffffu, by1);
  if (lane == 0) grow_bbox(S.bbox, bx0, by0, bx1, by1);
  SplatCtx C;
  C.tx0 = bx0; C.ty0 = by0;
  C.tw = bx1 - bx0 + 1;
  const int area = C.tw * (by1 - by0 + 1);
  const bool use_tile = area <= kWarpTilePx;
  C.tile = use_tile ? S.tile : nullptr;
  C.accum = S.accum; C.W = S.W; C.H = S.H; C.bilinear = S.bilinear;
  C.ch0 = S.ch0; C.ch1 = S.ch1; C.ch2 = S.ch2;
  if (use_tile) {
    for (int q = lane; q < 3 * area; q += 32) S.tile[q] = 0ull;
    __syncwarp();
  }
  if (ww.x > 0.f) splat1(C, pp.x, pp.y, ww.x);
  if (ww.y > 0.f) splat1(C, pp.z, pp.w, ww.y);
  if (!use_tile) return;
  __syncwarp();
  const float inv_tw = frcp((float)C.tw);
  for (int t = lane; t < area; t += 32) {
    const int jy = (int)(((float)t + 0.5f) * inv_tw);  // t / tw, exact for these small integers
    const int jx = t - jy * C.tw;
    unsigned long long* dst = S.accum + 3 * ((size_t)(bx0 + jx) + (size_t)(by0 + jy) * S.W);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const unsigned long long v = S.tile[3 * t + c];
      if (v) atomicAdd(dst + c, v);
    }
  }
  __syncwarp();  // the tile is reused by the warp's next fork
}

// ---------------------------------------------------------------------------------------------------------------
// v5: PREFIX SHARING.  Every ghost (i, j) of a light and wavelength begins with the same forward sweep through surfaces
// 0 .. j-1 -- 28 ghosts re-trace it 28 times, and more than half of all executed steps are in it (most rays die late in
// the sweep or right after the first reflection).  prefix_kernel traces that sweep ONCE per (light, wavelength) "slot"
// and caches each ray's state as it arrives ON every surface (32 bytes: position, direction x/y, Fresnel product, the
// two mask products; NaN = dead).  A ghost job then loads the state at its first-reflection surface j and starts with
// the reflection.  The cache is written once and read ~3 times per ray per slot out of L2.
// Layout: prefix[((slot * n_surf + k) * 2 + part) * half_rays + ray], part 0 = (ox, oy, oz, w), part 1 = (dx, dy, ma, mb);
// dz = +sqrt(1 - dx^2 - dy^2) (the sweep only travels forward).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) prefix_kernel(const Job* __restrict__ slots, const Step* __restrict__ progs, FrameGeom g,
                                                          const float* __restrict__ tex, float4* __restrict__ prefix,
                                                          unsigned long long* __restrict__ accum) {
  __shared__ Step s_prog[LFB_MAX_STEPS];
  __shared__ unsigned long long s_tile[(kThreads / 32) * kWarpTilePx * 3];
  const int half_rows = (g.N + 1) / 2;
  const int slot = blockIdx.z;  // 3-D grid (patch column, patch row, slot)
  const Job& J = slots[slot];
  const int n_steps = J.n_steps;  // forward refractions 0 .. n_surf-1 (the stop included); step n_surf is the sensor
  const bool do_direct = accum != nullptr && J.i != 0;  // this shard owns the slot's direct (unreflected) path: splat it too
  const int tid = threadIdx.x;
  {
    const float4* src = reinterpret_cast<const float4*>(progs + (size_t)slot * LFB_MAX_STEPS);
    float4* dst = reinterpret_cast<float4*>(s_prog);
    for (int q = tid; q < (n_steps + 1) * 3; q += kThreads) dst[q] = __ldg(src + q);
  }
  __syncthreads();
  const int a = blockIdx.x * 16 + (tid & 15), bp = blockIdx.y * 16 + (tid >> 4);
  const bool in_grid = a < g.N && bp < half_rows;
  const int b = g.N - 1 - bp;
  const size_t ray = (size_t)bp * g.N + a;
  MaskGeom M;
  M.tex = tex; M.tw = g.tex_w; M.th = g.tex_h;
  M.su = g.mask_su; M.ou = g.mask_ou; M.sv = g.mask_sv; M.ov = g.mask_ov;
  RayState r;
  start_ray(r, fmaf((float)a + 0.5f, g.cell, -g.P), fmaf((float)b + 0.5f, g.cell, -g.P), J.f_sin_t, J.f_cos_t, J.f_inv_dist);
  RayOut o;
  bool alive = in_grid;
  float4* base = prefix + (size_t)slot * g.n_surf * 2 * g.half_rays + ray;
#pragma unroll 1
  for (int s = 0; s < n_steps; s++) {
    const Step& S = s_prog[s];
    if (alive) alive = propagate<false>(S, r, o);
    if (in_grid) {
      float4* dst = base + (size_t)s * 2 * g.half_rays;
      if (alive) {
        dst[0] = make_float4(r.ox, r.oy, r.oz, r.w);
        dst[g.half_rays] = make_float4(r.dx, r.dy, r.ma, r.mb);
      } else {
        dst[0] = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f);
      }
    }
    if (alive) alive = interact<2, true, false>(S, M, g.lut, r, o);
  }
  if (!do_direct) return;
  // the direct path: on to the sensor plane and splat, one warp at a time
  if (alive) alive = propagate<false>(s_prog[n_steps], r, o);
  if (alive) { o.xs = r.ox; o.ys = r.oy; o.wa = r.w * r.ma; o.wb = r.w * r.mb; }
  PixMap PM;
  PM.sx = J.f_sx; PM.sy = J.f_sy; PM.cs = J.f_cs; PM.sn = J.f_sn; PM.ppu = J.f_ppu;
  WarpSplat WS;
  WS.tile = s_tile + (tid >> 5) * (kWarpTilePx * 3);
  WS.accum = accum; WS.bbox = g.bbox; WS.W = g.W; WS.H = g.H;
  WS.ch0 = J.f_chan[0]; WS.ch1 = J.f_chan[1]; WS.ch2 = J.f_chan[2];
  WS.bilinear = g.splat == LFB_SPLAT_BILINEAR;
  warp_splat(WS, PM, alive, o, b != bp, tid & 31);
}

// The ghost kernel of v5: exact_splat1_kernel with the ray states loaded from the prefix cache (jobs whose slot is >= 0);
// jobs without a slot (the direct path) trace from the entrance as before.
template <int RX, int RY, int MINB, int BT>
__global__ void __launch_bounds__(BT, MINB) exact_splat2_kernel(const Job* __restrict__ jobs, const Step* __restrict__ progs,
                                                                      FrameGeom g, const float* __restrict__ tex,
                                                                      unsigned long long* __restrict__ accum) {
  constexpr int RPT = RX * RY, PATCH = RPT * BT, PW = 16 * RX, PH = (BT / 16) * RY;  // BT threads = 16 x BT/16 rays per round
  __shared__ Step s_prog[LFB_MAX_STEPS];
  __shared__ unsigned long long s_tile[kTilePx * 3];
  __shared__ float4 s_qp[PATCH];
  __shared__ float2 s_qw[PATCH];
  __shared__ int s_count, s_bbox[4];

  const int half_rows = (g.N + 1) / 2;
  const int patches_x = (g.N + PW - 1) / PW;
  const int patches_per_job = patches_x * ((half_rows + PH - 1) / PH);
  const int job_id = blockIdx.x / patches_per_job;
  const int patch = blockIdx.x - job_id * patches_per_job;
  const int a0 = (patch % patches_x) * PW, b0 = (patch / patches_x) * PH;
  const Job& J = jobs[job_id];
  const int n_steps = J.n_steps;
  const int tid = threadIdx.x;
  {
    const float4* src = reinterpret_cast<const float4*>(progs + (size_t)job_id * LFB_MAX_STEPS);
    float4* dst = reinterpret_cast<float4*>(s_prog);
    for (int q = tid; q < n_steps * 3; q += BT) dst[q] = __ldg(src + q);
    if (tid == 0) { s_count = 0; s_bbox[0] = s_bbox[1] = 0x7fffffff; s_bbox[2] = s_bbox[3] = -0x7fffffff; }
  }
  __syncthreads();

  MaskGeom M;
  M.tex = tex; M.tw = g.tex_w; M.th = g.tex_h;
  M.su = g.mask_su; M.ou = g.mask_ou; M.sv = g.mask_sv; M.ov = g.mask_ov;
  PixMap PM;
  PM.sx = J.f_sx; PM.sy = J.f_sy; PM.cs = J.f_cs; PM.sn = J.f_sn; PM.ppu = J.f_ppu;
  const bool bilinear = g.splat == LFB_SPLAT_BILINEAR;
  const int slot = J.slot;
  const float4* pre = slot >= 0 ? g.prefix + ((size_t)(slot * g.n_surf + J.j_first) * 2) * g.half_rays : nullptr;

  int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
#pragma unroll 1
  for (int rr = 0; rr < RPT; rr++) {
    const int a = a0 + (tid & 15) + 16 * (rr % RX), bp = b0 + (tid >> 4) + (BT / 16) * (rr / RX);
    const int b = g.N - 1 - bp;
    float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 ww = make_float2(0.f, 0.f);
    if (a < g.N && bp < half_rows) {
      RayState r;
      RayOut o;
      bool alive;
      if (pre) {
        const float4* src = pre + ((size_t)bp * g.N + a);
        const float4 s0 = __ldg(src);
        alive = s0.x == s0.x;  // NaN: the ray died in the forward sweep before reaching surface j
        if (alive) {
          const float4 s1 = __ldg(src + g.half_rays);
          r.ox = s0.x; r.oy = s0.y; r.oz = s0.z; r.w = s0.w;
          r.dx = s1.x; r.dy = s1.y; r.ma = s1.z; r.mb = s1.w;
          r.dz = fsqrt(fmaxf(fmaf(-r.dx, r.dx, fmaf(-r.dy, r.dy, 1.f)), 0.f));
          alive = run_program<2, true, false, true>(s_prog, n_steps, M, g.lut, r, o);
        }
      } else {
        alive = trace<2, true, false>(s_prog, n_steps, M, g.lut, fmaf((float)a + 0.5f, g.cell, -g.P), fmaf((float)b + 0.5f, g.cell, -g.P),
                                      J.f_sin_t, J.f_cos_t, J.f_inv_dist, o);
      }
      if (alive) {
        int x0, y0, x1, y1;
        if (o.wa > 0.f) {
          to_pixel(PM, o.xs, o.ys, pp.x, pp.y);
          if (footprint(bilinear, pp.x, pp.y, g.W, g.H, x0, y0, x1, y1)) {
            ww.x = o.wa; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
          }
        }
        if (o.wb > 0.f && b != bp) {
          to_pixel(PM, o.xs, -o.ys, pp.z, pp.w);
          if (footprint(bilinear, pp.z, pp.w, g.W, g.H, x0, y0, x1, y1)) {
            ww.y = o.wb; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
          }
        }
      }
    }
    const bool lands = ww.x > 0.f || ww.y > 0.f;
    const unsigned ballot = __ballot_sync(0xffffffffu, lands);
    if (ballot) {
      const int lane = tid & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_count, __popc(ballot));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (lands) {
        const int q = base + __popc(ballot & ((1u << lane) - 1u));
        s_qp[q] = pp; s_qw[q] = ww;
      }
    }
  }
  if (__any_sync(0xffffffffu, bx1 >= bx0)) {
    bx0 = __reduce_min_sync(0xffffffffu, bx0); by0 = __reduce_min_sync(0xffffffffu, by0);
    bx1 = __reduce_max_sync(0xffffffffu, bx1); by1 = __reduce_max_sync(0xffffffffu, by1);
    if ((tid & 31) == 0) {
      atomicMin(&s_bbox[0], bx0); atomicMin(&s_bbox[1], by0);
      atomicMax(&s_bbox[2], bx1); atomicMax(&s_bbox[3], by1);
    }
  }
  __syncthreads();
  const int count = s_count;
  if (count == 0) return;
  if (tid == 0) grow_bbox(g.bbox, s_bbox[0], s_bbox[1], s_bbox[2], s_bbox[3]);
  SplatCtx C;
  C.tx0 = s_bbox[0]; C.ty0 = s_bbox[1];
  C.tw = s_bbox[2] - C.tx0 + 1;
  const int area = C.tw * (s_bbox[3] - C.ty0 + 1);
  const bool use_tile = area <= kTilePx;
  C.tile = use_tile ? s_tile : nullptr;
  C.accum = accum; C.W = g.W; C.H = g.H; C.bilinear = bilinear;
  if (use_tile) {
    for (int q = tid; q < 3 * area; q += BT) s_tile[q] = 0ull;
    __syncthreads();
  }
  C.ch0 = J.f_chan[0]; C.ch1 = J.f_chan[1]; C.ch2 = J.f_chan[2];
  for (int q = tid; q < count; q += BT) {
    const float4 qp = s_qp[q];
    const float2 qw = s_qw[q];
    if (qw.x > 0.f) splat1(C, qp.x, qp.y, qw.x);
    if (qw.y > 0.f) splat1(C, qp.z, qp.w, qw.y);
  }
  if (!use_tile) return;
  __syncthreads();
  const float inv_tw = frcp((float)C.tw);
  for (int t = tid; t < area; t += BT) {
    const int jy = (int)(((float)t + 0.5f) * inv_tw);  // t / tw, exact for these small integers
    const int jx = t - jy * C.tw;
    unsigned long long* dst = accum + 3 * ((size_t)(C.tx0 + jx) + (size_t)(C.ty0 + jy) * g.W);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const unsigned long long v = s_tile[3 * t + c];
      if (v) atomicAdd(dst + c, v);
    }
  }
}

// v6: WARP-AUTONOMOUS splat.  ncu on v5: 20 % of the stall cycles are CTA barriers -- warps whose rays died early wait for
// the one warp still tracing.  Here a warp owns its 32 ray pairs from trace to flush: survivors stay in registers (no
// queue), the sensor tile is per warp (64 pixels), and the only CTA-wide barrier is the one after staging the program.
// Flushing per warp instead of per CTA costs more global atomics (a few per warp), which the L2 absorbs.

template <int MINB, int BT>
__global__ void __launch_bounds__(BT, MINB) exact_splat3_kernel(const Job* __restrict__ jobs, const Step* __restrict__ progs,
                                                                FrameGeom g, const float* __restrict__ tex,
                                                                unsigned long long* __restrict__ accum) {
  constexpr int PH = BT / 16;  // BT threads = 16 x BT/16 ray pairs
  __shared__ Step s_prog[LFB_MAX_STEPS];
  __shared__ unsigned long long s_tile[(BT / 32) * kWarpTilePx * 3];

  // 3-D grid (patch column, patch row, job): no integer divisions in the prologue every warp pays
  const int half_rows = (g.N + 1) / 2;
  const int job_id = blockIdx.z;
  const Job& J = jobs[job_id];
  const int n_steps = J.n_steps;
  const int tid = threadIdx.x, lane = tid & 31;
  {
    const float4* src = reinterpret_cast<const float4*>(progs + (size_t)job_id * LFB_MAX_STEPS);
    float4* dst = reinterpret_cast<float4*>(s_prog);
    for (int q = tid; q < n_steps * 3; q += BT) dst[q] = __ldg(src + q);
  }
  __syncthreads();  // the only CTA-wide barrier

  MaskGeom M;
  M.tex = tex; M.tw = g.tex_w; M.th = g.tex_h;
  M.su = g.mask_su; M.ou = g.mask_ou; M.sv = g.mask_sv; M.ov = g.mask_ov;
  PixMap PM;
  PM.sx = J.f_sx; PM.sy = J.f_sy; PM.cs = J.f_cs; PM.sn = J.f_sn; PM.ppu = J.f_ppu;
  const bool bilinear = g.splat == LFB_SPLAT_BILINEAR;

  const int a = blockIdx.x * 16 + (tid & 15), bp = blockIdx.y * PH + (tid >> 4);
  const int b = g.N - 1 - bp;
  // Only lanes whose ray (or its mirror image) lands carry a sensor point, weights and a footprint: nothing below reads
  // them on the other lanes, so the trace loop's many death exits need no sentinel values kept in registers.
  int bx0, by0, bx1, by1;
  float4 pp;
  float2 ww;
  bool lands = false;
  if (a < g.N && bp < half_rows) {
    RayState r;
    RayOut o;
    bool alive;
    if (J.slot >= 0) {
      const float4* src = g.prefix + ((size_t)(J.slot * g.n_surf + J.j_first) * 2) * g.half_rays + ((size_t)bp * g.N + a);
      const float4 s0 = __ldg(src);
      alive = s0.x == s0.x;  // NaN: the ray died in the forward sweep before reaching surface j
      if (alive) {
        const float4 s1 = __ldg(src + g.half_rays);
        r.ox = s0.x; r.oy = s0.y; r.oz = s0.z; r.w = s0.w;
        r.dx = s1.x; r.dy = s1.y; r.ma = s1.z; r.mb = s1.w;
        r.dz = fsqrt(fmaxf(fmaf(-r.dx, r.dx, fmaf(-r.dy, r.dy, 1.f)), 0.f));
        alive = run_program<2, true, false, true>(s_prog, n_steps, M, g.lut, r, o);
      }
    } else {
      alive = trace<2, true, false>(s_prog, n_steps, M, g.lut, fmaf((float)a + 0.5f, g.cell, -g.P), fmaf((float)b + 0.5f, g.cell, -g.P),
                                    J.f_sin_t, J.f_cos_t, J.f_inv_dist, o);
    }
    if (alive) {
      int x0, y0, x1, y1;
      bx0 = by0 = 0x7fffffff; bx1 = by1 = -0x7fffffff;
      pp = make_float4(0.f, 0.f, 0.f, 0.f);
      ww = make_float2(0.f, 0.f);
      if (o.wa > 0.f) {
        to_pixel(PM, o.xs, o.ys, pp.x, pp.y);
        if (footprint(bilinear, pp.x, pp.y, g.W, g.H, x0, y0, x1, y1)) {
          ww.x = o.wa; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
        }
      }
      if (o.wb > 0.f && b != bp) {
        to_pixel(PM, o.xs, -o.ys, pp.z, pp.w);
        if (footprint(bilinear, pp.z, pp.w, g.W, g.H, x0, y0, x1, y1)) {
          ww.y = o.wb; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
        }
      }
      lands = ww.x > 0.f || ww.y > 0.f;
    }
  }
  const unsigned landing = __ballot_sync(0xffffffffu, lands);
  if (!landing) return;  // the whole warp is done: nothing below involves other warps
  if (lands) {  // the footprint of the warp's landing rays, reduced among those lanes only
    bx0 = __reduce_min_sync(landing, bx0); by0 = __reduce_min_sync(landing, by0);
    bx1 = __reduce_max_sync(landing, bx1); by1 = __reduce_max_sync(landing, by1);
  }
  const int leader = __ffs(landing) - 1;
  bx0 = __shfl_sync(0xffffffffu, bx0, leader); by0 = __shfl_sync(0xffffffffu, by0, leader);
  bx1 = __shfl_sync(0xffffffffu, bx1, leader); by1 = __shfl_sync(0xffffffffu, by1, leader);
  if (lane == 0) grow_bbox(g.bbox, bx0, by0, bx1, by1);
  SplatCtx C;
  C.tx0 = bx0; C.ty0 = by0;
  C.tw = bx1 - bx0 + 1;
  const int area = C.tw * (by1 - by0 + 1);
  const bool use_tile = area <= kWarpTilePx;
  unsigned long long* tile = s_tile + (tid >> 5) * (kWarpTilePx * 3);
  C.tile = use_tile ? tile : nullptr;
  C.accum = accum; C.W = g.W; C.H = g.H; C.bilinear = bilinear;
  C.ch0 = J.f_chan[0]; C.ch1 = J.f_chan[1]; C.ch2 = J.f_chan[2];
  if (use_tile) {
    for (int q = lane; q < 3 * area; q += 32) tile[q] = 0ull;
    __syncwarp();
  }
  if (lands) {
    if (ww.x > 0.f) splat1(C, pp.x, pp.y, ww.x);
    if (ww.y > 0.f) splat1(C, pp.z, pp.w, ww.y);
  }
  if (!use_tile) return;
  __syncwarp();
  const float inv_tw = frcp((float)C.tw);
  for (int t = lane; t < area; t += 32) {
    const int jy = (int)(((float)t + 0.5f) * inv_tw);  // t / tw, exact for these small integers
    const int jx = t - jy * C.tw;
    unsigned long long* dst = accum + 3 * ((size_t)(bx0 + jx) + (size_t)(by0 + jy) * g.W);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const unsigned long long v = tile[3 * t + c];
      if (v) atomicAdd(dst + c, v);
    }
  }
}

// v6i: TWO RAY PAIRS PER THREAD IN LOCKSTEP (opt-in, LFB_EXACT_ILP=2).  The final v6 capture is latency bound (1.3 eligible
// of 9.9 active warps per scheduler): here a thread carries the ray pairs of rows bp and bp + PH through the SAME step
// program with straight-line code -- two independent dependency chains per warp instruction stream, the step constants
// read from shared memory once for both.  Death needs no branch: a dead ray is a NaN state that stays NaN (the table and
// mask lookups clamp NaN indices), and the loop only exits when both are dead.  Same arithmetic per ray => same bits.
__device__ __forceinline__ void propagate2(const Step& S, RayState& r, bool& alive) {  // propagate<> without the branch
  const float c = S.c;
  const float pz = r.oz + S.dz;
  const float pd = fmaf(r.ox, r.dx, fmaf(r.oy, r.dy, pz * r.dz));
  const float pp = fmaf(r.ox, r.ox, fmaf(r.oy, r.oy, pz * pz));
  const float B = fmaf(c, pd, -r.dz);
  const float Cq = fmaf(c, pp, -2.f * pz);
  const float disc = fmaf(B, B, -c * Cq);
  const float t = -Cq * frcp(B + copysignf(fsqrt(disc), B));
  r.ox = fmaf(t, r.dx, r.ox); r.oy = fmaf(t, r.dy, r.oy); r.oz = fmaf(t, r.dz, pz);
  alive = alive && (fmaf(r.ox, r.ox, r.oy * r.oy) <= S.semi2);
}
__device__ __forceinline__ void bend2(const Step& S, bool refl, const float2* __restrict__ lut, RayState& r) {  // interact<2,...>
  const float c = S.c, eta = S.eta;
  const float nx = -c * r.ox, ny = -c * r.oy, nz = fmaf(-c, r.oz, 1.f);
  const float nd = fmaf(nx, r.dx, fmaf(ny, r.dy, nz * r.dz));
  const float c0 = fabsf(nd);
  const float s2 = fmaf(-c0, c0, 1.f);
  const float k2 = fmaf(-S.eta2, s2, 1.f);
  const float c2 = fsqrt(k2);
  const float g = __int_as_float(__float_as_int(fmaf(eta, c0, -c2)) ^ (~__float_as_int(nd) & 0x80000000));
  const float alpha = refl ? 1.f : eta, beta = refl ? -2.f * nd : g;
  r.dx = fmaf(alpha, r.dx, beta * nx); r.dy = fmaf(alpha, r.dy, beta * ny); r.dz = fmaf(alpha, r.dz, beta * nz);
  const float R = reflectance_lut(lut, S.lut, eta > 1.f ? c2 : c0);
  r.w *= refl ? R : 1.f - R;
}
// one step for the lane's two rays: every (uniform) branch is taken once for both, so the two chains share basic blocks
__device__ __forceinline__ void step_pair(const Step& S, const MaskGeom& M, const float2* __restrict__ lut, RayState& r0, RayState& r1,
                                          bool& al0, bool& al1) {
  propagate2(S, r0, al0);
  propagate2(S, r1, al1);
  const int op = S.op;
  if (op >= STEP_PASS) {
    if (op == STEP_STOP) {
      r0.ma *= mask_lookup(M, r0.ox, r0.oy);
      r1.ma *= mask_lookup(M, r1.ox, r1.oy);
      r0.mb *= mask_lookup(M, r0.ox, -r0.oy);
      r1.mb *= mask_lookup(M, r1.ox, -r1.oy);
      al0 = al0 && !(r0.ma == 0.f && r0.mb == 0.f);
      al1 = al1 && !(r1.ma == 0.f && r1.mb == 0.f);
    }
    return;
  }
  const bool refl = op == STEP_REFLECT;
  bend2(S, refl, lut, r0);
  bend2(S, refl, lut, r1);
}

// the landing / tile / flush phase of one ray pair per lane (what exact_splat3_kernel does after its trace)
__device__ __forceinline__ void warp_land(const FrameGeom& g, const Job& J, const PixMap& PM, bool bilinear, bool alive, const RayOut& o,
                                          bool has_mirror, unsigned long long* tile, unsigned long long* __restrict__ accum, int lane) {
  int bx0, by0, bx1, by1;
  float4 pp;
  float2 ww;
  bool lands = false;
  if (alive) {
    int x0, y0, x1, y1;
    bx0 = by0 = 0x7fffffff; bx1 = by1 = -0x7fffffff;
    pp = make_float4(0.f, 0.f, 0.f, 0.f);
    ww = make_float2(0.f, 0.f);
    if (o.wa > 0.f) {
      to_pixel(PM, o.xs, o.ys, pp.x, pp.y);
      if (footprint(bilinear, pp.x, pp.y, g.W, g.H, x0, y0, x1, y1)) {
        ww.x = o.wa; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
      }
    }
    if (o.wb > 0.f && has_mirror) {
      to_pixel(PM, o.xs, -o.ys, pp.z, pp.w);
      if (footprint(bilinear, pp.z, pp.w, g.W, g.H, x0, y0, x1, y1)) {
        ww.y = o.wb; bx0 = min(bx0, x0); by0 = min(by0, y0); bx1 = max(bx1, x1); by1 = max(by1, y1);
      }
    }
    lands = ww.x > 0.f || ww.y > 0.f;
  }
  const unsigned landing = __ballot_sync(0xffffffffu, lands);
  if (!landing) return;
  if (lands) {
    bx0 = __reduce_min_sync(landing, bx0); by0 = __reduce_min_sync(landing, by0);
    bx1 = __reduce_max_sync(landing, bx1); by1 = __reduce_max_sync(landing, by1);
  }
  const int leader = __ffs(landing) - 1;
  bx0 = __shfl_sync(0xffffffffu, bx0, leader); by0 = __shfl_sync(0xffffffffu, by0, leader);
  bx1 = __shfl_sync(0xffffffffu, bx1, leader); by1 = __shfl_sync(0xffffffffu, by1, leader);
  if (lane == 0) grow_bbox(g.bbox, bx0, by0, bx1, by1);
  SplatCtx C;
  C.tx0 = bx0; C.ty0 = by0;
  C.tw = bx1 - bx0 + 1;
  const int area = C.tw * (by1 - by0 + 1);
  const bool use_tile = area <= kWarpTilePx;
  C.tile = use_tile ? tile : nullptr;
  C.accum = accum; C.W = g.W; C.H = g.H; C.bilinear = bilinear;
  C.ch0 = J.f_chan[0]; C.ch1 = J.f_chan[1]; C.ch2 = J.f_chan[2];
  if (use_tile) {
    for (int q = lane; q < 3 * area; q += 32) tile[q] = 0ull;
    __syncwarp();
  }
  if (lands) {
    if (ww.x > 0.f) splat1(C, pp.x, pp.y, ww.x);
    if (ww.y > 0.f) splat1(C, pp.z, pp.w, ww.y);
  }
  if (!use_tile) return;
  __syncwarp();
  const float inv_tw = frcp((float)C.tw);
  for (int t = lane; t < area; t += 32) {
    const int jy = (int)(((float)t + 0.5f) * inv_tw);
    const int jx = t - jy * C.tw;
    unsigned long long* dst = accum + 3 * ((size_t)(bx0 + jx) + (size_t)(by0 + jy) * g.W);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const unsigned long long v = tile[3 * t + c];
      if (v) atomicAdd(dst + c, v);
    }
  }
  __syncwarp();  // the tile is reused by the lane's second ray pair
}

template <int MINB, int BT>
__global__ void __launch_bounds__(BT, MINB) exact_splat4_kernel(const Job* __restrict__ jobs, const Step* __restrict__ progs,
                                                                FrameGeom g, const float* __restrict__ tex,
                                                                unsigned long long* __restrict__ accum) {
  constexpr int PH = BT / 16;  // BT threads = 16 x PH lanes, each with the ray pairs of rows bp and bp + PH
  __shared__ Step s_prog[LFB_MAX_STEPS];
  __shared__ unsigned long long s_tile[(BT / 32) * kWarpTilePx * 3];
  const int half_rows = (g.N + 1) / 2;
  const int job_id = blockIdx.z;
  const Job& J = jobs[job_id];
  const int n_steps = J.n_steps;
  const int tid = threadIdx.x, lane = tid & 31;
  {
    const float4* src = reinterpret_cast<const float4*>(progs + (size_t)job_id * LFB_MAX_STEPS);
    float4* dst = reinterpret_cast<float4*>(s_prog);
    for (int q = tid; q < n_steps * 3; q += BT) dst[q] = __ldg(src + q);
  }
  __syncthreads();  // the only CTA-wide barrier
  MaskGeom M;
  M.tex = tex; M.tw = g.tex_w; M.th = g.tex_h;
  M.su = g.mask_su; M.ou = g.mask_ou; M.sv = g.mask_sv; M.ov = g.mask_ov;
  PixMap PM;
  PM.sx = J.f_sx; PM.sy = J.f_sy; PM.cs = J.f_cs; PM.sn = J.f_sn; PM.ppu = J.f_ppu;
  const bool bilinear = g.splat == LFB_SPLAT_BILINEAR;
  const int a = blockIdx.x * 16 + (tid & 15);
  const int bp0 = blockIdx.y * (2 * PH) + (tid >> 4), bp1 = bp0 + PH;
  RayState r0, r1;
  bool al0 = a < g.N && bp0 < half_rows, al1 = a < g.N && bp1 < half_rows;
  int s_begin = 0;
  if (J.slot >= 0) {  // (uniform) both states start ON the first-reflection surface, from the prefix cache
    const float4* base = g.prefix + ((size_t)(J.slot * g.n_surf + J.j_first) * 2) * g.half_rays;
    const float4 nanv = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, 0.f);
    const float4 p0 = al0 ? __ldg(base + ((size_t)bp0 * g.N + a)) : nanv;
    const float4 p1 = al1 ? __ldg(base + ((size_t)bp1 * g.N + a)) : nanv;
    al0 = al0 && p0.x == p0.x;
    al1 = al1 && p1.x == p1.x;
    const float4 q0 = al0 ? __ldg(base + g.half_rays + ((size_t)bp0 * g.N + a)) : nanv;
    const float4 q1 = al1 ? __ldg(base + g.half_rays + ((size_t)bp1 * g.N + a)) : nanv;
    r0.ox = p0.x; r0.oy = p0.y; r0.oz = p0.z; r0.w = p0.w; r0.dx = q0.x; r0.dy = q0.y; r0.ma = q0.z; r0.mb = q0.w;
    r1.ox = p1.x; r1.oy = p1.y; r1.oz = p1.z; r1.w = p1.w; r1.dx = q1.x; r1.dy = q1.y; r1.ma = q1.z; r1.mb = q1.w;
    r0.dz = fsqrt(fmaxf(fmaf(-r0.dx, r0.dx, fmaf(-r0.dy, r0.dy, 1.f)), 0.f));
    r1.dz = fsqrt(fmaxf(fmaf(-r1.dx, r1.dx, fmaf(-r1.dy, r1.dy, 1.f)), 0.f));
    RayOut o;  // interact<> wants one; unused in the throughput variant
    if (al0) al0 = interact<2, true, false>(s_prog[0], M, g.lut, r0, o);  // step 0: interaction only
    if (al1) al1 = interact<2, true, false>(s_prog[0], M, g.lut, r1, o);
    s_begin = 1;
  } else {  // the direct path (or a job without a cached sweep): from the entrance grid
    start_ray(r0, fmaf((float)a + 0.5f, g.cell, -g.P), fmaf((float)(g.N - 1 - bp0) + 0.5f, g.cell, -g.P), J.f_sin_t, J.f_cos_t, J.f_inv_dist);
    start_ray(r1, fmaf((float)a + 0.5f, g.cell, -g.P), fmaf((float)(g.N - 1 - bp1) + 0.5f, g.cell, -g.P), J.f_sin_t, J.f_cos_t, J.f_inv_dist);
  }
#pragma unroll 1
  for (int s = s_begin; s < n_steps && (al0 || al1); s++) {
    const Step& S = s_prog[s];
    step_pair(S, M, g.lut, r0, r1, al0, al1);
  }
  unsigned long long* tile = s_tile + (tid >> 5) * (kWarpTilePx * 3);
  RayOut o0, o1;
  o0.xs = r0.ox; o0.ys = r0.oy; o0.wa = r0.w * r0.ma; o0.wb = r0.w * r0.mb;
  o1.xs = r1.ox; o1.ys = r1.oy; o1.wa = r1.w * r1.ma; o1.wb = r1.w * r1.mb;
  warp_land(g, J, PM, bilinear, al0, o0, (g.N - 1 - bp0) != bp0, tile, accum, lane);
  warp_land(g, J, PM, bilinear, al1, o1, (g.N - 1 - bp1) != bp1, tile, accum, lane);
}

// ---------------------------------------------------------------------------------------------------------------
// v7: GHOST FAMILIES.  After the forward sweep, the ghosts (i, j) of one (light, wavelength) with the same first
// reflection j also share the BACKWARD sweep j-1, j-2, ... : ghost (i, j) leaves it at surface i.  One thread now follows
// the whole family: it starts ON surface j (prefix cache), reflects, and walks backward; at every candidate surface k
// it forks -- a copy of the ray reflects at k and runs the forward program k+1 .. n-1 to the sensor, where the warp
// splats it -- and then continues backward through k.  28 ghost jobs per (light, wavelength) become 7 family jobs: the
// per-warp fixed costs (launch prologue, state load, program staging), which dominated v6, are paid once per family, and
// the backward sweeps are not repeated.  The arithmetic along every ghost path is unchanged, so the frame is bit-identical.
//
// Family program (shared memory): [0] reflect at j; then for k = j-1 .. 0 two steps: (backward step at k: refraction or
// the stop plane, fork step: reflection at k seen from behind, op < 0 when ghost (k, j) is not wanted).  The forward
// program of the slot (refractions 0 .. n-1 + sensor) supplies the suffix k+1 .. n of every fork.
// ---------------------------------------------------------------------------------------------------------------
template <int MINB, int BT>
__global__ void __launch_bounds__(BT, MINB) exact_family_kernel(const Job* __restrict__ fams, const Step* __restrict__ fam_progs,
                                                                const Job* __restrict__ slots, const Step* __restrict__ slot_progs,
                                                                FrameGeom g, const float* __restrict__ tex,
                                                                unsigned long long* __restrict__ accum) {
  constexpr int PH = BT / 16;
  __shared__ Step s_fam[2 * LFB_MAX_SURFACES + 2];
  __shared__ Step s_fwd[LFB_MAX_SURFACES + 2];
  __shared__ unsigned long long s_tile[(BT / 32) * kWarpTilePx * 3];

  const int half_rows = (g.N + 1) / 2;
  const Job& J = fams[blockIdx.z];
  const int slot = J.slot, j = J.j_first, n_fam = J.n_steps;
  const int n_fwd = g.n_surf + 1;  // forward refractions 0 .. n-1 and the sensor
  const int tid = threadIdx.x, lane = tid & 31;
  const int a = blockIdx.x * 16 + (tid & 15), bp = blockIdx.y * PH + (tid >> 4);
  const int b = g.N - 1 - bp;
  const bool in_grid = a < g.N && bp < half_rows;
  {
    const float4* src = reinterpret_cast<const float4*>(fam_progs + (size_t)blockIdx.z * LFB_MAX_STEPS);
    float4* dst = reinterpret_cast<float4*>(s_fam);
    for (int q = tid; q < n_fam * 3; q += BT) dst[q] = __ldg(src + q);
    const float4* src2 = reinterpret_cast<const float4*>(slot_progs + (size_t)slot * LFB_MAX_STEPS);
    float4* dst2 = reinterpret_cast<float4*>(s_fwd);
    for (int q = tid; q < n_fwd * 3; q += BT) dst2[q] = __ldg(src2 + q);
  }
  __syncthreads();  // the only CTA-wide barrier

  MaskGeom M;
  M.tex = tex; M.tw = g.tex_w; M.th = g.tex_h;
  M.su = g.mask_su; M.ou = g.mask_ou; M.sv = g.mask_sv; M.ov = g.mask_ov;
  PixMap PM;
  PM.sx = J.f_sx; PM.sy = J.f_sy; PM.cs = J.f_cs; PM.sn = J.f_sn; PM.ppu = J.f_ppu;
  WarpSplat WS;
  WS.tile = s_tile + (tid >> 5) * (kWarpTilePx * 3);
  WS.accum = accum; WS.bbox = g.bbox; WS.W = g.W; WS.H = g.H;
  WS.ch0 = J.f_chan[0]; WS.ch1 = J.f_chan[1]; WS.ch2 = J.f_chan[2];
  WS.bilinear = g.splat == LFB_SPLAT_BILINEAR;

  RayState r;
  RayOut o;
  bool alive = false;
  if (in_grid) {
    const float4* src = g.prefix + ((size_t)(slot * g.n_surf + j) * 2) * g.half_rays + ((size_t)bp * g.N + a);
    const float4 s0 = __ldg(src);
    alive = s0.x == s0.x;  // NaN: the ray died in the forward sweep before reaching surface j
    if (alive) {
      const float4 s1 = __ldg(src + g.half_rays);
      r.ox = s0.x; r.oy = s0.y; r.oz = s0.z; r.w = s0.w;
      r.dx = s1.x; r.dy = s1.y; r.ma = s1.z; r.mb = s1.w;
      r.dz = fsqrt(fmaxf(fmaf(-r.dx, r.dx, fmaf(-r.dy, r.dy, 1.f)), 0.f));
      alive = interact<2, true, false>(s_fam[0], M, g.lut, r, o);  // the first reflection, at j
    }
  }
  if (!__any_sync(0xffffffffu, alive)) return;  // nothing below involves other warps
  const int n_back = (n_fam - 1) >> 1;  // backward surfaces in this job's program: j-1 .. down to its lowest fork
#pragma unroll 1
  for (int e = 0; e < n_back; e++) {
    const int k = j - 1 - e;
    const Step& Sb = s_fam[1 + 2 * e];
    const Step& Sf = s_fam[2 + 2 * e];
    if (alive) alive = propagate<false>(Sb, r, o);  // onto surface k, travelling backward
    if (Sf.op >= 0) {  // ghost (k, j): fork a copy that reflects here and runs forward to the sensor
      RayState q = r;
      RayOut oq;
      bool a2 = alive;
      if (a2) a2 = interact<2, true, false>(Sf, M, g.lut, q, oq);
#pragma unroll 1
      for (int s = k + 1; s < n_fwd; s++) {
        if (a2) a2 = propagate<false>(s_fwd[s], q, oq);
        if (a2) a2 = interact<2, true, false>(s_fwd[s], M, g.lut, q, oq);
      }
      if (a2) { oq.xs = q.ox; oq.ys = q.oy; oq.wa = q.w * q.ma; oq.wb = q.w * q.mb; }
      warp_splat(WS, PM, a2, oq, b != bp, lane);
    }
    if (alive) alive = interact<2, true, false>(Sb, M, g.lut, r, o);  // on through surface k (or the stop's mask)
    if (!__any_sync(0xffffffffu, alive)) return;
  }
}

// Parity instrument for the same trace code: one record per ray (flags, positions, weight).
__global__ void __launch_bounds__(kThreads) exact_dump_kernel(const Job* __restrict__ job, const Step* __restrict__ prog, FrameGeom g,
                                                              const float* __restrict__ tex, lfb_ray_hit* __restrict__ out) {
  __shared__ Step s_prog[LFB_MAX_STEPS];
  const Job& J = *job;
  const int n_steps = J.n_steps;
  {
    const float4* src = reinterpret_cast<const float4*>(prog);
    float4* dst = reinterpret_cast<float4*>(s_prog);
    for (int q = threadIdx.x; q < n_steps * 3; q += kThreads) dst[q] = __ldg(src + q);
  }
  __syncthreads();
  const int tile = blockIdx.x;
  const int a = (tile % g.tiles_x) * 16 + (threadIdx.x & 15);
  const int b = (tile / g.tiles_x) * 16 + (threadIdx.x >> 4);
  if (a >= g.N || b >= g.N) return;
  const float P = g.P;
  const float cell = 2.f * P / (float)g.N;
  MaskGeom M;
  M.tex = tex; M.tw = g.tex_w; M.th = g.tex_h;
  const float h = g.h_stop;
  M.su = 0.5f * (float)g.tex_w / h; M.ou = 0.5f * (float)g.tex_w;
  M.sv = -0.5f * (float)g.tex_h / h; M.ov = 0.5f * (float)g.tex_h;
  PixMap PM;
  PM.sx = (float)J.sx; PM.sy = (float)J.sy; PM.cs = (float)J.cs; PM.sn = (float)J.sn; PM.ppu = (float)J.ppu;
  RayOut o;
  const bool alive = g.lut ? trace<2, false, true>(s_prog, n_steps, M, g.lut, fmaf((float)a + 0.5f, cell, -P), fmaf((float)b + 0.5f, cell, -P),
                                                   (float)J.sin_t, (float)J.cos_t, (float)J.inv_dist, o)
                           : trace<1, false, true>(s_prog, n_steps, M, g.lut, fmaf((float)a + 0.5f, cell, -P), fmaf((float)b + 0.5f, cell, -P),
                                                   (float)J.sin_t, (float)J.cos_t, (float)J.inv_dist, o);
  lfb_ray_hit rec;
  rec.x_ap = o.xa; rec.y_ap = o.ya; rec.flags = o.flags; rec.pad = 0;
  if (alive) {
    float px, py;
    to_pixel(PM, o.xs, o.ys, px, py);
    rec.x_s = o.xs; rec.y_s = o.ys; rec.weight = o.wa; rec.px = px; rec.py = py;
    const float fx = floorf(px), fy = floorf(py);
    if (!(fx >= 0.f && fx < (float)g.W && fy >= 0.f && fy < (float)g.H)) rec.flags |= LFB_RAY_OFF_SENSOR;
  } else {
    rec.x_s = rec.y_s = rec.px = rec.py = CUDART_NAN;
    rec.weight = 0.0;
  }
  out[(size_t)b * g.N + a] = rec;
}

}  // namespace xf32
}  // namespace lfb
