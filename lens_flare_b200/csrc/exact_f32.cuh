// exact_f32.cuh -- the throughput kernel of the ghost path: EXACT_GRID in FP32.
//
// The kernel is bound by instruction issue on the scalar FP32 pipes (ncu, profiles/): there is no
// memory stream to speak of (rays are generated from their grid index, a ghost's whole description is
// ~1 KB, the 1 MB aperture mask is L2/L1 resident), so the design minimises instructions per ray:
//
//  * STEP PROGRAM.  The host flattens each ghost (i, j, lambda) into a straight list of steps
//    (refract / reflect / stop plane / sensor plane) with every ray-independent quantity already
//    computed: curvature, vertex shift, clear radius^2, n0/n2 and its square, the coating's film index,
//    (n0/n1)^2 and phase factor.  A CTA stages its ghost's program in shared memory once; the per-ray
//    loop has no "which surface / which direction / which glass" logic and no divisions by constants.
//  * FAST SCALAR MATH.  rcp.approx / sqrt.approx / cos.approx (one MUFU each) instead of the IEEE
//    division and square root sequences (MUFU + FCHK + slow-path branch + Newton FFMAs), and Fresnel /
//    thin-film reflectances restructured to ONE reciprocal per surface.
//  * TWO PASSES WITH SURVIVOR COMPACTION.  Most rays of a bundle die (vignetted, total internal
//    reflection, blocked by the aperture mask) and carry no energy.  Pass 1 traces geometry only and
//    queues the survivors of the CTA's ray patch (with their sensor pixel) in shared memory; pass 2
//    re-traces only the survivors, in dense warps, with the Fresnel / coating weights.
//  * SHARED-MEMORY SENSOR TILE.  Neighbouring rays land on neighbouring (often the same) pixels: the
//    CTA accumulates its deposits in a small u64 fixed-point tile in shared memory and flushes each
//    touched pixel with one global atomic (falls back to direct global atomics when the patch's footprint
//    exceeds the tile).  Integer accumulation keeps the frame bit-stable for any schedule or GPU count.
//
// Parity: per-ray and image tolerances against the double-precision oracle are in
// tests/test_gpu_parity.py; the FP64 kernels in ghost_grid_impl.cuh remain the bit-exact instruments.
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

struct RayOut {
  float xs, ys, xa, ya, w;
  unsigned flags;
};

struct MaskGeom {
  const float* tex;
  int tw, th;
  float su, sv, ou, ov;  // u = xa*su + ou ; v = ya*sv + ov
};

__device__ __forceinline__ float mask_lookup(const MaskGeom& M, float xa, float ya) {
  const float fu = floorf(fmaf(xa, M.su, M.ou)), fv = floorf(fmaf(ya, M.sv, M.ov));
  if (!(fu >= 0.f && fu < (float)M.tw && fv >= 0.f && fv < (float)M.th)) return 0.f;
  return __ldg(M.tex + (int)fv * M.tw + (int)fu);
}

// Trace one ray through a step program.  WEIGHTS = false: geometry only (w = 1 if the ray survives
// every step with a non-zero mask, else 0).  KEEP_GOING = true (dump instrument): rays stopped by the
// mask continue with weight 0 so that their positions stay comparable with the oracle.
template <bool WEIGHTS, bool KEEP_GOING>
__device__ __forceinline__ bool trace(const Step* __restrict__ prog, int n_steps, const MaskGeom& M, float x, float y,
                                      float sin_t, float cos_t, RayOut& o) {
  float ox = x, oy = y, oz = 0.f, dx = sin_t, dy = 0.f, dz = cos_t, w = 1.f;
  o.flags = 0;
  o.xa = o.ya = CUDART_NAN_F;
  for (int s = 0; s < n_steps; s++) {
    const Step S = prog[s];
    const float px = ox, py = oy, pz = oz + S.dz;
    if (S.op >= STEP_STOP) {  // planes perpendicular to the axis: the stop and the sensor
      const float t = -pz * frcp(dz);
      ox = fmaf(t, dx, px); oy = fmaf(t, dy, py); oz = 0.f;
      if (S.op == STEP_STOP) {
        o.xa = ox; o.ya = oy;
        const float m = mask_lookup(M, ox, oy);
        w *= m;
        if (m == 0.f) {
          o.flags |= LFB_RAY_STOPPED;
          if (!KEEP_GOING) { o.w = 0.f; return false; }
        }
      }
      continue;
    }
    const float c = S.c;
    const float pd = fmaf(px, dx, fmaf(py, dy, pz * dz));
    const float pp = fmaf(px, px, fmaf(py, py, pz * pz));
    const float B = fmaf(c, pd, -dz);
    const float Cq = fmaf(c, pp, -2.f * pz);
    const float disc = fmaf(B, B, -c * Cq);
    if (disc < 0.f) { o.flags |= LFB_RAY_MISSED; o.w = 0.f; return false; }
    const float t = -Cq * frcp(B + copysignf(fsqrt(disc), B));
    const float hx = fmaf(t, dx, px), hy = fmaf(t, dy, py), hz = fmaf(t, dz, pz);
    ox = hx; oy = hy; oz = hz;
    if (fmaf(hx, hx, hy * hy) > S.semi2) { o.flags |= LFB_RAY_VIGNETTED; o.w = 0.f; return false; }
    if (S.op == STEP_PASS) continue;
    const float nx = -c * hx, ny = -c * hy, nz = fmaf(-c, hz, 1.f);
    const float nd = fmaf(nx, dx, fmaf(ny, dy, nz * dz));
    const float c0 = fabsf(nd);
    const float k2 = fmaf(-S.eta2, fmaf(-c0, c0, 1.f), 1.f);
    if (S.op == STEP_REFLECT) {
      const float m2 = -2.f * nd;
      dx = fmaf(m2, nx, dx); dy = fmaf(m2, ny, dy); dz = fmaf(m2, nz, dz);
      if (WEIGHTS) w *= (k2 < 0.f) ? 1.f : reflectance(S, c0, fsqrt(k2));
    } else {
      if (k2 < 0.f) { o.flags |= LFB_RAY_TIR; o.w = 0.f; return false; }
      const float c2 = fsqrt(k2);
      // d' = eta d + (eta c0 - c2) N with N = -sign(nd) n the normal facing the ray: flip the factor's sign when nd > 0
      const float g = __int_as_float(__float_as_int(fmaf(S.eta, c0, -c2)) ^ (~__float_as_int(nd) & 0x80000000));
      dx = fmaf(S.eta, dx, g * nx); dy = fmaf(S.eta, dy, g * ny); dz = fmaf(S.eta, dz, g * nz);
      if (WEIGHTS) w *= 1.f - reflectance(S, c0, c2);
    }
  }
  o.xs = ox; o.ys = oy; o.w = w;
  return true;
}

struct PixMap {
  float sx, sy, cs, sn, ppu;
};
__device__ __forceinline__ void to_pixel(const PixMap& P, float xs, float ys, float& px, float& py) {
  const float X = -P.ppu * xs, Y = P.ppu * ys;
  px = P.sx + (X * P.cs - Y * P.sn);
  py = P.sy + (X * P.sn + Y * P.cs);
}

constexpr int kThreads = 256;
constexpr int kTilePx = 256;  // shared-memory sensor tile capacity (pixels)

// One CTA = a (16*RX) x (16*RY) patch of one ghost's ray grid; RPT = RX*RY rays per thread in pass 1.
template <int RX, int RY>
__global__ void __launch_bounds__(kThreads) exact_splat_kernel(const Job* __restrict__ jobs, const Step* __restrict__ progs,
                                                               FrameGeom g, const float* __restrict__ tex,
                                                               unsigned long long* __restrict__ accum) {
  constexpr int RPT = RX * RY, PATCH = RPT * kThreads;
  __shared__ Step s_prog[LFB_MAX_STEPS];
  __shared__ unsigned long long s_tile[kTilePx * 3];
  __shared__ float s_qx[PATCH], s_qy[PATCH];
  __shared__ unsigned short s_qid[PATCH];
  __shared__ int s_count, s_bbox[4];

  const int patches_x = (g.N + 16 * RX - 1) / (16 * RX);
  const int patches_per_job = patches_x * ((g.N + 16 * RY - 1) / (16 * RY));
  const int job_id = blockIdx.x / patches_per_job;
  const int patch = blockIdx.x - job_id * patches_per_job;
  const int a0 = (patch % patches_x) * (16 * RX), b0 = (patch / patches_x) * (16 * RY);
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
  const float sin_t = (float)J.sin_t, cos_t = (float)J.cos_t;
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

  // ---- pass 1: geometry only; queue the survivors with their sensor pixel ---------------------
  int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
#pragma unroll 1
  for (int r = 0; r < RPT; r++) {
    const int la = (tid & 15) + 16 * (r % RX), lb = (tid >> 4) + 16 * (r / RX);
    const int a = a0 + la, b = b0 + lb;
    bool live = a < g.N && b < g.N;
    float px = 0.f, py = 0.f;
    int ix0 = 0, iy0 = 0, ix1 = 0, iy1 = 0;
    if (live) {
      RayOut o;
      live = trace<false, false>(s_prog, n_steps, M, fmaf((float)a + 0.5f, cell, -P), fmaf((float)b + 0.5f, cell, -P), sin_t, cos_t, o);
      if (live) {
        to_pixel(PM, o.xs, o.ys, px, py);
        if (bilinear) {
          const float fx = floorf(px - 0.5f), fy = floorf(py - 0.5f);
          live = fx >= -1.f && fx < (float)g.W && fy >= -1.f && fy < (float)g.H;  // false for NaN
          ix0 = max((int)fx, 0); iy0 = max((int)fy, 0);
          ix1 = min((int)fx + 1, g.W - 1); iy1 = min((int)fy + 1, g.H - 1);
        } else {
          const float fx = floorf(px), fy = floorf(py);
          live = fx >= 0.f && fx < (float)g.W && fy >= 0.f && fy < (float)g.H;
          ix0 = ix1 = (int)fx; iy0 = iy1 = (int)fy;
        }
      }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, live);
    if (ballot) {
      const int lane = tid & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_count, __popc(ballot));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (live) {
        const int slot = base + __popc(ballot & ((1u << lane) - 1u));
        s_qid[slot] = (unsigned short)(lb * (16 * RX) + la);
        s_qx[slot] = px; s_qy[slot] = py;
        bx0 = min(bx0, ix0); by0 = min(by0, iy0); bx1 = max(bx1, ix1); by1 = max(by1, iy1);
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
  const int tx0 = s_bbox[0], ty0 = s_bbox[1];
  const int tw = s_bbox[2] - tx0 + 1, th = s_bbox[3] - ty0 + 1;
  const bool use_tile = tw * th <= kTilePx;
  if (use_tile) {
    for (int q = tid; q < tw * th * 3; q += kThreads) s_tile[q] = 0ull;
    __syncthreads();
  }

  // ---- pass 2: survivors only, dense warps: full trace with Fresnel / coating weights -----------
  const float scale = (float)g.fp_scale;
  const float ch0 = (float)J.chan[0] * scale, ch1 = (float)J.chan[1] * scale, ch2 = (float)J.chan[2] * scale;
  for (int q = tid; q < count; q += kThreads) {
    const int id = s_qid[q];
    const int la = id % (16 * RX), lb = id / (16 * RX);
    RayOut o;
    if (!trace<true, false>(s_prog, n_steps, M, fmaf((float)(a0 + la) + 0.5f, cell, -P), fmaf((float)(b0 + lb) + 0.5f, cell, -P), sin_t, cos_t, o)) continue;
    if (!(o.w > 0.f)) continue;
    const float px = s_qx[q], py = s_qy[q];
    int ix, iy, ntap;
    float wt[4];
    if (bilinear) {
      const float qx = px - 0.5f, qy = py - 0.5f;
      const float fx0 = floorf(qx), fy0 = floorf(qy);
      const float fx = qx - fx0, fy = qy - fy0;
      ix = (int)fx0; iy = (int)fy0; ntap = 4;
      wt[0] = o.w * ((1.f - fx) * (1.f - fy)); wt[1] = o.w * (fx * (1.f - fy));
      wt[2] = o.w * ((1.f - fx) * fy); wt[3] = o.w * (fx * fy);
    } else {
      ix = (int)floorf(px); iy = (int)floorf(py); ntap = 1;
      wt[0] = o.w;
    }
#pragma unroll
    for (int t = 0; t < 4; t++) {
      if (t >= ntap) break;
      const int jx = ix + (t & 1), jy = iy + (t >> 1);
      if (jx < 0 || jx >= g.W || jy < 0 || jy >= g.H) continue;
      const long long q0 = __float2ll_rn(wt[t] * ch0), q1 = __float2ll_rn(wt[t] * ch1), q2 = __float2ll_rn(wt[t] * ch2);
      unsigned long long* dst = use_tile ? s_tile + 3 * ((jy - ty0) * tw + (jx - tx0)) : accum + 3 * ((size_t)jx + (size_t)jy * g.W);
      if (q0) atomicAdd(dst + 0, (unsigned long long)q0);
      if (q1) atomicAdd(dst + 1, (unsigned long long)q1);
      if (q2) atomicAdd(dst + 2, (unsigned long long)q2);
    }
  }
  if (!use_tile) return;
  __syncthreads();
  // ---- flush the tile: one global atomic per touched (pixel, channel) --------------------------
  for (int q = tid; q < tw * th * 3; q += kThreads) {
    const unsigned long long v = s_tile[q];
    if (v) {
      const int p = q / 3, c = q - 3 * p;
      const int jy = ty0 + p / tw, jx = tx0 + (p - (p / tw) * tw);
      atomicAdd(accum + 3 * ((size_t)jx + (size_t)jy * g.W) + c, v);
    }
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
  const bool alive = trace<true, true>(s_prog, n_steps, M, fmaf((float)a + 0.5f, cell, -P), fmaf((float)b + 0.5f, cell, -P),
                                       (float)J.sin_t, (float)J.cos_t, o);
  lfb_ray_hit rec;
  rec.x_ap = o.xa; rec.y_ap = o.ya; rec.flags = o.flags; rec.pad = 0;
  if (alive) {
    float px, py;
    to_pixel(PM, o.xs, o.ys, px, py);
    rec.x_s = o.xs; rec.y_s = o.ys; rec.weight = o.w; rec.px = px; rec.py = py;
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
