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
  double P_d, cell_d;  // the same in double (LFB_STRICT)
  float mask_su, mask_sv, mask_ou, mask_ov;  // aperture texel of a stop-plane point: u = x*su + ou, v = y*sv + ov
  const float4* prefix;  // cached forward sweeps: [slot][surface][part][ray of the half grid], see exact_trace.cuh
  int half_rays, n_surf;
  const float2* lut;  // reflectance tables over the cosine in the rarer medium, one per (wavelength, surface, direction):
                      // (R_i, R_{i+1} - R_i); the trace only reads them below kPolyV0 (steeper than ~53 degrees)
  int* bbox;        // device int[4] = {min_x, min_y, max_x, max_y} of every pixel the frame deposits into (or nullptr)
  unsigned* tile_bits;  // dirty-tile map of the accumulators: byte t != 0 = 16x16 sensor tile t received a deposit (or nullptr)
  int tiles_w;          // tiles per sensor row
  float poly_v0;        // weights come from the step's polynomial for cosines >= this (kPolyV0; > 1 forces the tables)
  int pad2, pad3;
  unsigned long long* stats;  // STATS instantiations only: [0] executed surface steps, [1] ray pairs started, [2] ray pairs landed
};

// Job headers of one launch of the ghost / family kernels, passed BY VALUE (kernel parameter space = constant bank): a CTA
// knows its job's cache slot, first-reflection surface and program length without a global load, so the ray-state and
// program loads of its prologue are issued at once -- one L2 round trip instead of two.  Launches take kHeadsPerLaunch jobs.
constexpr int kHeadsPerLaunch = 1024;
struct JobHeads {
  unsigned h[kHeadsPerLaunch];  // bits 0-15 slot + 1 (0: no cached sweep), 16-20 first-reflection surface, 21-27 program length
};
inline unsigned pack_head(int slot, int j_first, int n_steps) {
  return (unsigned)(slot + 1) | ((unsigned)(j_first < 0 ? 0 : j_first) << 16) | ((unsigned)n_steps << 21);
}

constexpr int kTilePxLog2 = 4;  // the dirty-tile map's tiles are 16 x 16 pixels
constexpr int kTilePx1 = 1 << kTilePxLog2;
inline int tiles_across(int px) { return (px + kTilePx1 - 1) >> kTilePxLog2; }

// A sensor accumulator buffer (lfb_accum_bytes): W*H*3 u64 fixed-point sums, then the dirty-tile map of those sums: one
// BYTE per 16 x 16 tile (non-zero = a splat deposited into it), handled four tiles to a 32-bit word.
struct AccumLayout {
  int tiles_w, tiles_h, n_tiles, n_words;
  size_t px_bytes, bits_off, total;
};
inline AccumLayout accum_layout(int W, int H) {
  AccumLayout a;
  a.tiles_w = tiles_across(W); a.tiles_h = tiles_across(H);
  a.n_tiles = a.tiles_w * a.tiles_h;
  a.n_words = (a.n_tiles + 3) / 4;
  a.px_bytes = sizeof(unsigned long long) * 3 * (size_t)W * (size_t)H;
  a.bits_off = (a.px_bytes + 255) & ~(size_t)255;
  a.total = a.bits_off + (((size_t)a.n_words * 4 + 255) & ~(size_t)255);
  return a;
}
// The state that goes with an OUTPUT buffer of the tile-sparse finalize (lfb_tile_state_bytes): which tiles the previous
// frame left non-zero in it.  [0] ticket, [1] tiles (staged units) of the last frame, [2] staging unit counter, [3] pad, then
// n_words words of the PREVIOUS map (per tile one byte: the mask of its 8 x 8 quadrants the output holds non-zero pixels in) and
// n_words words of the NEXT map (written during a launch, rolled over by its last CTA).
inline size_t tile_state_bytes(int W, int H) { return 16 + (((size_t)accum_layout(W, H).n_words * 8 + 255) & ~(size_t)255); }

// EXACT_GRID step program (exact_trace.cuh): a ghost flattened into straight-line steps with every ray-independent
// quantity precomputed on the host.  T = float (the FP32 throughput kernels, 64 B per step) or double (LFB_STRICT: FP64
// geometry, 80 B).  The weight factor of a step -- R at a reflection, 1 - R at a refraction -- is a degree-7 polynomial in
// x = (1 - v) * kPolyScale, v = the cosine of the ray's angle in the rarer medium, fitted on the host in double on
// v in [kPolyV0, 1] (|error| <= 3e-8: tighter than any table, and no memory access in the surface loop); steeper rays fall
// back to the interface's reflectance table.
enum StepOp { STEP_REFRACT = 0, STEP_REFLECT = 1, STEP_PASS = 2, STEP_STOP = 3, STEP_SENSOR = 4 };
#define LFB_MAX_STEPS (3 * LFB_MAX_SURFACES + 2)
constexpr int kLutSize = 1024;  // intervals of the table variable (cosine in the rarer medium) on [0, 1]
constexpr float kPolyV0 = 0.6f;
constexpr float kPolyScale = 2.5f;  // 1 / (1 - kPolyV0)
constexpr int kPolyN = 8;
template <typename T>
struct StepT {
  T c, dz, eta, eta2;  // curvature (0: plane); z of the previous vertex minus z of this one; n0/n2; (n0/n2)^2
  float semi2;         // clear radius^2
  int op, lut;         // StepOp (-1: a fork the family does not take); index of the interface's reflectance table
  float pad;
  float p[kPolyN];     // weight-factor polynomial, lowest order first
};
typedef StepT<float> StepF;
typedef StepT<double> StepD;
static_assert(sizeof(StepF) == 64 && sizeof(StepD) == 80, "step programs are staged as 16-byte words");

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
// PARAXIAL_GRID in FP32 / FP64 and the FP64 oracle-order EXACT_GRID parity kernels (ghost_grid_impl.cuh)
cudaError_t launch_trace_splat_f32(const Job* jobs, int n_jobs, const FrameGeom& g, int mode, const float* tex,
                                   unsigned long long* accum, cudaStream_t s);
cudaError_t launch_trace_splat_f64(const Job* jobs, int n_jobs, const FrameGeom& g, int mode, const float* tex,
                                   unsigned long long* accum, cudaStream_t s);
cudaError_t launch_trace_dump_f32(const Job* job, const FrameGeom& g, int mode, const float* tex, lfb_ray_hit* out, cudaStream_t s);
cudaError_t launch_trace_dump_f64(const Job* job, const FrameGeom& g, int mode, const float* tex, lfb_ray_hit* out,
                                  cudaStream_t s);
// The EXACT_GRID throughput kernels (exact_trace.cuh), instantiated for T = float (exact_f32.cu, LFB_FP32) and T = double
// (exact_f64.cu, LFB_STRICT: FP64 geometry, FP32 weights).  stats: the counting instantiation (FrameGeom::stats).
//   prefix    the forward sweep of every (light, lambda) slot, traced once; each ray's state on every surface is cached
//             (and the direct path splatted when accum_for_direct is given and the slot owns it)
//   ghosts    one job per ghost pair (i, j), starting ON surface j from the cache (or from the entrance without one)
//   families  one job per (light, lambda, first reflection j): forks at every second reflection
//   dump      per-ray records of one ghost (parity instrument)
template <typename T> int exact_prefix_parts();  // 16-byte words per cached ray state
template <typename T>
cudaError_t launch_exact_prefix(const Job* slots, const StepT<T>* progs, int n_slots, const FrameGeom& g, const float* tex, float4* prefix,
                                unsigned long long* accum_for_direct, bool stats, cudaStream_t s);
template <typename T>
cudaError_t launch_exact_ghosts(const Job* jobs, const StepT<T>* progs, const unsigned* heads, int n_jobs, const FrameGeom& g, const float* tex,
                                unsigned long long* accum, int ctas_per_sm, bool stats, cudaStream_t s);
template <typename T>
cudaError_t launch_exact_families(const Job* fams, const StepT<T>* fam_progs, const unsigned* heads, int n_fams, const Job* slots,
                                  const StepT<T>* slot_progs, const FrameGeom& g, const float* tex, unsigned long long* accum, int ctas_per_sm,
                                  bool stats, cudaStream_t s);
template <typename T>
cudaError_t launch_exact_dump(const Job* job, const StepT<T>* prog, const FrameGeom& g, const float* tex, lfb_ray_hit* out, cudaStream_t s);
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
// sparse.cu: dirty tiles of the ranks' accumulators -> pixels of `out` (n = 1: this GPU's finalize)
cudaError_t launch_tiles(const PeerAccums& P, int rank, int W, int H, double inv_scale, void* out, size_t stride, int elem,
                         unsigned* state, unsigned* count_out, int ctas, cudaStream_t s, void* stage = nullptr);
cudaError_t launch_tile_drain(int W, int H, void* out, size_t stride, const unsigned* state, const void* stage, int ctas, float gbps,
                              cudaStream_t s);
cudaError_t measure_host_write_gbps(cudaStream_t s, float* gbps);
// bytes of the optional device staging buffer of launch_tiles: every tile's pixels + one word per tile
inline size_t tile_stage_bytes(int W, int H, size_t stride) {
  const AccumLayout lay = accum_layout(W, H);
  return (size_t)lay.n_tiles * kTilePx1 * kTilePx1 * stride + sizeof(unsigned) * 4 * (size_t)lay.n_tiles;  // pixels + one word per quadrant
}
cudaError_t launch_peer_barrier(const PeerFlags& F, int rank, unsigned long long epoch, cudaStream_t s);
cudaError_t launch_reduce_finalize(const PeerAccums& P, const unsigned long long* mc, size_t p0, size_t p1, double inv_scale,
                                   void* out, size_t stride, int elem, int ctas, cudaStream_t s);
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
// The dirty-tile map: one byte per 16 x 16 sensor tile, non-zero = a splat deposited into it.  Marking is "look, then store
// the byte if it is still zero": no atomics (a lost race only repeats the store), and after the first few warps of a frame
// almost every look finds the byte set.  Measured on B200 at cfg2, ghost kernel under ncu: a bit map with a volatile read
// right before atomicOr 84 us (26 % of the stall samples on the read); unconditional atomicOr 1.5 ms (same-address atomics
// serialise in L2); unconditional byte stores 152 us (the hot sectors back the store path up: drain / MIO stalls
// everywhere).  The throughput kernels therefore issue the look early (tile_peek) and act on it after the splat (tile_mark_if).
// The EXACT_GRID landing code looks through L1 (__ldca) by default: within a kernel a byte only goes 0 -> 1, so a stale L1 line can
// only show 0 for a set byte (the store is repeated, harmless), never 1 for a clear one; the maps are cleared by earlier
// operations of the stream, and L1 is invalidated at every kernel launch.  cfg2 ghost kernel 0.091 -> 0.079 ms (tools/kernel_ab.py).
__device__ __forceinline__ unsigned char* tile_byte(unsigned* map, int tiles_w, int tx, int ty) {
  return reinterpret_cast<unsigned char*>(map) + ((unsigned)ty * (unsigned)tiles_w + (unsigned)tx);
}
__device__ __forceinline__ unsigned tile_peek(const unsigned char* p) { return __ldcg(p); }  // L2, not L1: other SMs' marks are seen
__device__ __forceinline__ void mark_tile(unsigned* map, int tiles_w, int tx, int ty) {
  unsigned char* p = tile_byte(map, tiles_w, tx, ty);
  if (!tile_peek(p)) *p = 1;
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
                             const double rad_sum[3], void* out, size_t stride, int elem, int additive, bool spectrum_cached, cudaStream_t s,
                             int* launches);

// scene.cu: the path-traced scene pass
struct SceneStore;
cudaError_t scene_upload(const lfb_scene* sc, SceneStore** out, const char** err_msg);
void scene_free(SceneStore* s);
cudaError_t launch_scene(const SceneStore* s, const lfb_camera* cam, int W, int H, void* out, size_t stride, int elem, int additive, cudaStream_t st);

cudaError_t launch_to_color(const double* hdr, int W, int H, uint32_t* out, int flip, cudaStream_t s);

cudaError_t probe_peaks(int device, cudaStream_t s, double* fp32_flops, double* mufu_ops, double* sm_clock_hz);

}  // namespace lfb
