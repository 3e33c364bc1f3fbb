// lfb_internal.h -- shared between the host C-ABI layer and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lfb200.h"

namespace lfb {

// Lens tables as the kernels see them (one copy in __constant__ memory per device).
// Everything the per-ray loops index is warp-uniform, so constant-cache broadcasts apply.
struct DevLens {
  int n_surfaces, stop, n_lambda, pad;
  float c[LFB_MAX_SURFACES];        // curvature
  float d[LFB_MAX_SURFACES];        // thickness after surface k
  float zv[LFB_MAX_SURFACES + 1];   // vertex z of surface k; zv[n] = sensor plane
  float semi[LFB_MAX_SURFACES];     // clear semi-aperture (radius)
  float coat[LFB_MAX_SURFACES];     // quarter-wave design wavelength (nm), 0 = bare
  double zv_d[LFB_MAX_SURFACES + 1];
  float ior[LFB_MAX_LAMBDA][LFB_MAX_SURFACES];
  float lambda_nm[LFB_MAX_LAMBDA];
  double P;           // entrance half height
  double h_stop;      // stop half height
  double h_stop_neg;  // reference's asymmetric re-aim height
};

// One (light, ghost pair, wavelength) unit of work.  Filled on the host except the
// paraxial matrices, which the device set-up kernel writes (paraxial_setup_kernel).
struct Job {
  int light, i, j, lambda;  // i = j = -1: direct path
  int n_cross, n_steps;     // n_steps: length of the job's FP32 step program
  int slot, j_first;        // FP32 prefix path: (light, lambda) slot of the cached forward sweep and the surface of the first
                            // reflection the job's program starts ON (slot < 0: the program starts at the entrance)
  float theta;
  float pad3;
  double chan[3];           // radiance * rgb_weight * ray area * px_per_unit^2
  double sx, sy, cs, sn, ppu;  // sensor mapping (sun pixel, rotation, pixels per lens unit)
  double sin_t, cos_t;         // EXACT_GRID: direction of the light's parallel bundle
  double inv_dist;             // point light: 1 / distance to the first vertex (0: directional)
  double cross[3][4];       // entrance -> stop plane, per crossing (m00,m01,m10,m11)
  double full[4];           // entrance -> sensor
  // the same constants as floats, for the FP32 throughput kernel (no per-CTA double -> float conversions)
  float f_sin_t, f_cos_t, f_sx, f_sy, f_cs, f_sn, f_ppu, f_inv_dist;
  float f_chan[4];          // chan * 2^fixed_point_bits
};

struct FrameGeom {
  int W, H, N, splat;
  int tiles_x, tiles_per_job;
  int tex_w, tex_h;
  double fp_scale;  // 2^fixed_point_bits
  float P, h_stop;  // entrance half height, stop half height
  float cell;       // 2P / N
  float mask_su, mask_sv, mask_ou, mask_ov;  // aperture texel of a stop-plane point: u = x*su + ou, v = y*sv + ov
  const float4* prefix;  // cached forward sweeps: [slot][surface][part 0|1][ray of the half grid], see exact_f32.cuh
  int half_rays, n_surf;
  const float2* lut;  // reflectance tables R(sin^2 theta0), one per (wavelength, surface, direction): (R_i, R_{i+1} - R_i)
  int* bbox;        // device int[4] = {min_x, min_y, max_x, max_y} of every pixel the frame deposits into (or nullptr)
  int patch, pad;   // FP32 EXACT_GRID tuning: rays per thread in pass 1 (1, 2 or 4); resident CTAs/SM target (0 -> 4)
};

// FP32 EXACT_GRID step program (exact_f32.cuh): a ghost flattened into straight-line steps with every
// ray-independent quantity precomputed on the host.  48 bytes = 3 x float4.
enum StepOp { STEP_REFRACT = 0, STEP_REFLECT = 1, STEP_PASS = 2, STEP_STOP = 3, STEP_SENSOR = 4 };
#define LFB_MAX_STEPS (3 * LFB_MAX_SURFACES + 2)
constexpr int kLutSize = 1024;  // intervals of the table variable (cosine in the rarer medium) on [0, 1]
struct Step {
  float c, dz, semi2, eta;   // curvature (0: plane); z of the previous vertex minus z of this one; clear radius^2; n0/n2
  float eta2, phase;         // (n0/n2)^2; coating phase factor pi * lambda0 / lambda
  int op, lut;               // StepOp; index of the interface's reflectance table (kLutSize float2 entries each)
  float n0, n2, n1, e1sq;    // indices before/after along the ray; film index (0 = bare); (n0/n1)^2
};

// REF_QUADS: one rasterisable triangle of a ghost quad (pathtracer.cpp:346-410 state
// after the y-sort and the -0.5 shift), in draw order.
struct RefTri {
  float x0, y0, u0, v0, x1, y1, u1, v1, x2, y2, u2, v2;
  int min_x, max_x, min_y, max_y;  // loop bounds, upper exclusive
  double col[3];
};

// Per-frame constants of REF_QUADS computed once on the host with the host libm (the
// reference calls atan/cosf/sinf per vertex with frame-constant arguments, :414).
struct RefFrame {
  int W, H, tex_w, tex_h;
  int n_pairs, n_lambda;
  float angle_to_sun;  // ray angle theta
  float cs, sn;        // cosf/sinf of float(atan((ay-.5)/(ax-.5)))
  double gb_mid_w, gb_mid_h;
  int has_sun, pad;
};

// one __constant__ copy of the lens per translation unit
cudaError_t upload_lens_f32(const DevLens& h, cudaStream_t s);
cudaError_t upload_lens_f64(const DevLens& h, cudaStream_t s);
cudaError_t upload_lens_ref(const DevLens& h, cudaStream_t s);

// kernels' host launchers (definitions in the .cu files)
cudaError_t launch_paraxial_setup(Job* jobs, int n_jobs, int physical_backward, cudaStream_t s);
cudaError_t launch_trace_splat_f32(const Job* jobs, const Step* progs, int n_jobs, const FrameGeom& g, int mode,
                                   const float* tex, unsigned long long* accum, cudaStream_t s);
cudaError_t launch_trace_splat_f64(const Job* jobs, int n_jobs, const FrameGeom& g, int mode, const float* tex,
                                   unsigned long long* accum, cudaStream_t s);
// FP32 EXACT_GRID prefix pass: trace the forward sweep of every (light, lambda) slot once and cache the ray states
cudaError_t launch_prefix_f32(const Job* slots, const Step* progs, int n_slots, const FrameGeom& g, const float* tex, float4* prefix,
                              unsigned long long* accum_for_direct, cudaStream_t s);
// FP32 EXACT_GRID ghost families (v7): one job per (light, lambda, first reflection j)
cudaError_t launch_family_f32(const Job* fams, const Step* fam_progs, int n_fams, const Job* slots, const Step* slot_progs,
                              const FrameGeom& g, const float* tex, unsigned long long* accum, cudaStream_t s);
cudaError_t launch_trace_dump_f32(const Job* job, const Step* prog, const FrameGeom& g, int mode, const float* tex,
                                  lfb_ray_hit* out, cudaStream_t s);
cudaError_t launch_trace_dump_f64(const Job* job, const FrameGeom& g, int mode, const float* tex, lfb_ray_hit* out,
                                  cudaStream_t s);
cudaError_t launch_finalize(const unsigned long long* accum, int W, int H, double inv_scale, void* out,
                            size_t stride, int elem, int additive, cudaStream_t s);
cudaError_t launch_finalize_clear(unsigned long long* accum, int W, int H, double inv_scale, void* out, size_t stride, int elem,
                                  cudaStream_t s);
struct PeerAccums {  // the ranks' sensor accumulators as mapped in THIS process (NVLink peer memory)
  const unsigned long long* ptr[LFB_MAX_PEERS];
  int n;
};
struct PeerFlags {  // every rank's barrier flag array (LFB_MAX_PEERS u64 each) as mapped in THIS process
  unsigned long long* ptr[LFB_MAX_PEERS];
  int n, rank_self;
};
cudaError_t launch_peer_barrier(const PeerFlags& F, int rank, unsigned long long epoch, cudaStream_t s);
cudaError_t launch_reduce_finalize(const PeerAccums& P, const unsigned long long* mc, size_t p0, size_t p1, double inv_scale,
                                   void* out, size_t stride, int elem, cudaStream_t s);
// rect = {x0, y0, x1, y1} inclusive; out is PACKED: pixel (x, y) at ((y - y0) * (x1 - x0 + 1) + (x - x0)) * stride
cudaError_t launch_finalize_rect(const unsigned long long* accum, int W, const int rect[4], double inv_scale, void* out,
                                 size_t stride, int elem, cudaStream_t s);

// Grow the frame's bounding box (device int[4]) to cover [x0,x1] x [y0,y1].  Read first: after the first few CTAs
// almost no caller extends the box, so the atomics are rare.
__device__ __forceinline__ void grow_bbox(int* bb, int x0, int y0, int x1, int y1) {
  if (!bb) return;
  volatile int* v = bb;
  if (x0 < v[0]) atomicMin(bb + 0, x0);
  if (y0 < v[1]) atomicMin(bb + 1, y0);
  if (x1 > v[2]) atomicMax(bb + 2, x1);
  if (y1 > v[3]) atomicMax(bb + 3, y1);
}
cudaError_t launch_ref_setup(const RefFrame& f, const int* pairs, const float* rgb_weight, RefTri* tris,
                             lfb_ref_ghost* ghosts, int* bbox, cudaStream_t s);
cudaError_t launch_ref_raster(const RefFrame& f, const RefTri* tris, int n_tris, const float* tex, void* out,
                              size_t stride, int elem, int additive, const int* rect, cudaStream_t s);

// Starburst (starburst.cu): frame constants of PathTracer::raytrace_starburst (pathtracer.cpp:947-1004)
struct StarFrame {
  int W, H, tw, th;
  int bx0, by0, bw, bh;     // bounding box of mask texels > 0 (CameraApertureTexture min/max_x/y, camera.h:64-70)
  double lr, ud;            // compute_phase :918-934
  double org_x, org_y;      // flare origin in pixels (ceil(fo * size))
  double total;             // CameraApertureTexture::total_value
  double flare_radius, exponent;  // exponent = 3 - flare_intensity (2 when that is <= 0), :998-1001
  // the lattice the two matrix products are evaluated on: n_col columns / n_row rows; lattice_x / lattice_y: that axis is
  // the P-periodic class lattice (period P = W_t or 2 W_t), else one column / row per pixel
  // herm: both axes on the lattice -> |F(-r,-c)| = |F(r,c)| (real mask), only rows 0..period/2 are computed (n_row = P/2+1)
  int n_col, n_row, lattice_x, lattice_y, period, herm;
};
size_t starburst_scratch_bytes(const StarFrame& f);
cudaError_t launch_starburst(const StarFrame& f, const float* tex, void* scratch, const double* lights_dev, int n_lights,
                             const double rad_sum[3], void* out, size_t stride, int elem, int additive, cudaStream_t s, int* launches);

cudaError_t launch_to_color(const double* hdr, int W, int H, uint32_t* out, int flip, cudaStream_t s);

cudaError_t probe_peaks(int device, cudaStream_t s, double* fp32_flops, double* mufu_ops, double* sm_clock_hz);

}  // namespace lfb
