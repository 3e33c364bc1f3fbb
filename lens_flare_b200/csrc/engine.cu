// engine.cu -- the C ABI of liblfb200.so (include/lfb200.h): host-side orchestration of the
// sm_100a ghost kernels.  Nothing here computes pixels or rays on the CPU: the host builds
// the job list (which ghost, which light, which wavelength, frame constants), uploads it,
// launches kernels on the engine's stream and copies results out.  There is no CPU fallback;
// without a CUDA device lfb_create fails with LFB_ERR_NO_DEVICE.
//
// Reference seam (paths under the reference tree): the caller is
// RaytracedRenderer::start_raytracing (src/pathtracer/raytraced_renderer.cpp:303-311), which
// runs find_sun_pos() and generate_ghost_buffer() once per render on its own thread; the
// result is PathTracer::ghost_buffer (src/pathtracer/pathtracer.h:54).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "lfb_internal.h"

using namespace lfb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int fail_cuda(cudaError_t err, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(err);
  return LFB_ERR_CUDA;
}
#define CU(call)                                        \
  do {                                                  \
    cudaError_t err_ = (call);                          \
    if (err_ != cudaSuccess) return fail_cuda(err_, #call); \
  } while (0)

struct JobId {
  int light, i, j, lambda;
};

// Ghost pairs in the reference's order: both reflections before the stop, then both after
// (pathtracer.cpp:735-762); LFB_PAIRS_ALL appends the pairs that straddle the stop.
int list_pairs(const lfb_lens& L, int pair_set, int pairs[][2]) {
  int n = 0;
  const int S = L.stop_index, N = L.n_surfaces;
  for (int i = 0; i < S; i++)
    for (int j = i + 1; j < S; j++) { pairs[n][0] = i; pairs[n][1] = j; n++; }
  for (int i = S + 1; i < N; i++)
    for (int j = i + 1; j < N; j++) { pairs[n][0] = i; pairs[n][1] = j; n++; }
  if (pair_set == LFB_PAIRS_ALL)
    for (int i = 0; i < S; i++)
      for (int j = S + 1; j < N; j++) { pairs[n][0] = i; pairs[n][1] = j; n++; }
  return n;
}

// Ray-surface interactions of one ray along ghost (i,j): forward 0..j, back j-1..i, forward
// i+1..n-1, sensor plane  =>  2(j-i) + n + 1; the direct path meets n surfaces + the sensor.
int interactions_of(const lfb_lens& L, int i, int j) { return i < 0 ? L.n_surfaces + 1 : 2 * (j - i) + L.n_surfaces + 1; }

// All (light, pair, lambda) jobs of a frame, then this shard's share (see the rule at the end).
void list_jobs(const lfb_lens& L, const lfb_params& P, int n_lights, std::vector<JobId>& out) {
  int pairs[LFB_MAX_SURFACES * LFB_MAX_SURFACES][2];
  std::vector<JobId> all;
  if (P.mode == LFB_MODE_REF_QUADS) {
    // one job per reference ghost quad (single sun: the last light wins, pathtracer.cpp:50-53)
    const int np = list_pairs(L, LFB_PAIRS_REF, pairs);
    if (n_lights > 0)
      for (int p = 0; p < np; p++)
        for (int c = 0; c < L.n_lambda; c++) all.push_back({n_lights - 1, pairs[p][0], pairs[p][1], c});
  } else {
    const int np = list_pairs(L, P.pair_set, pairs);
    for (int l = 0; l < n_lights; l++)
      for (int p = -1; p < np; p++) {
        if (p < 0 && !P.include_direct) continue;
        for (int c = 0; c < L.n_lambda; c++) all.push_back({l, p < 0 ? -1 : pairs[p][0], p < 0 ? -1 : pairs[p][1], c});
      }
  }
  out.clear();
  if (P.shard_count <= 1) { out = all; return; }
  // Sharding.  All ghosts of a (light, lambda) group share one forward sweep (the prefix cache), so groups are dealt WHOLE, in
  // CONTIGUOUS blocks (groups are light-major: a shard gets all wavelengths of a few lights, so its deposits -- and the tiles
  // the cross-GPU reduce has to fetch from it -- stay near those lights), as far as they divide evenly among the shards:
  // groups 0 .. floor(G / S) * S - 1.  The jobs of the remaining G mod S groups (all of them when G < S) are put in
  // longest-processing-time-first order (stable) and dealt one by one, so that e.g. one RGB light on two GPUs is split
  // 1.5 : 1.5 groups, not 2 : 1.
  const int S = P.shard_count;
  const int n_groups = P.mode == LFB_MODE_REF_QUADS ? 0 : n_lights * L.n_lambda;
  const int n_whole = (n_groups / S) * S;
  std::vector<JobId> rest;
  for (const JobId& id : all) {
    const int grp = id.light * L.n_lambda + id.lambda;
    if (grp < n_whole) { if (grp / (n_whole / S) == P.shard_index) out.push_back(id); }
    else rest.push_back(id);
  }
  std::stable_sort(rest.begin(), rest.end(), [&](const JobId& a, const JobId& b) {
    return interactions_of(L, a.i, a.j) > interactions_of(L, b.i, b.j);
  });
  for (size_t q = 0; q < rest.size(); q++)
    if ((int)(q % (size_t)S) == P.shard_index) out.push_back(rest[q]);
}

int check_lens(const lfb_lens* L) {
  if (!L) return fail(LFB_ERR_INVALID, "lens is NULL");
  if (L->n_surfaces < 2 || L->n_surfaces > LFB_MAX_SURFACES) return fail(LFB_ERR_INVALID, "n_surfaces out of range");
  if (L->n_lambda < 1 || L->n_lambda > LFB_MAX_LAMBDA) return fail(LFB_ERR_INVALID, "n_lambda out of range");
  if (L->stop_index < 0 || L->stop_index >= L->n_surfaces) return fail(LFB_ERR_INVALID, "stop_index out of range");
  if (!(L->entrance_half_height > 0) || !(L->stop_half_height > 0)) return fail(LFB_ERR_INVALID, "half heights must be > 0");
  return LFB_OK;
}

int check_params(const lfb_params* P, bool need_grid) {
  if (!P) return fail(LFB_ERR_INVALID, "params is NULL");
  if (P->mode < LFB_MODE_REF_QUADS || P->mode > LFB_MODE_EXACT_GRID) return fail(LFB_ERR_INVALID, "unknown mode");
  if (P->width < 1 || P->height < 1 || P->width > 32768 || P->height > 32768) return fail(LFB_ERR_INVALID, "sensor size out of range");
  if (need_grid || P->mode != LFB_MODE_REF_QUADS) {
    if (P->mode == LFB_MODE_REF_QUADS) return fail(LFB_ERR_INVALID, "this call needs a grid mode");
    if (P->grid_n < 1 || P->grid_n > 32768) return fail(LFB_ERR_INVALID, "grid_n out of range");
    if (P->pair_set != LFB_PAIRS_REF && P->pair_set != LFB_PAIRS_ALL) return fail(LFB_ERR_INVALID, "unknown pair_set");
    if (P->precision != LFB_FP32 && P->precision != LFB_FP64 && P->precision != LFB_STRICT) return fail(LFB_ERR_INVALID, "unknown precision");
    if (P->precision == LFB_STRICT && P->mode != LFB_MODE_EXACT_GRID) return fail(LFB_ERR_INVALID, "LFB_STRICT applies to LFB_MODE_EXACT_GRID only");
    if (P->splat != LFB_SPLAT_NEAREST && P->splat != LFB_SPLAT_BILINEAR) return fail(LFB_ERR_INVALID, "unknown splat");
    if (P->fixed_point_bits < 0 || P->fixed_point_bits > 56) return fail(LFB_ERR_INVALID, "fixed_point_bits out of range");
  }
  if (P->shard_count < 0 || (P->shard_count > 0 && (P->shard_index < 0 || P->shard_index >= P->shard_count)))
    return fail(LFB_ERR_INVALID, "bad shard_index/shard_count");
  return LFB_OK;
}

size_t elem_bytes(int elem) { return elem == LFB_F32x3 ? 12 : 24; }

}  // namespace

struct lfb_engine {
  int device = 0;
  lfb_options opt;  // as given to lfb_create_ex (0 = default everywhere)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_frame0 = nullptr, ev_trace0 = nullptr, ev_trace1 = nullptr, ev_frame1 = nullptr;
  bool has_lens = false, has_tex = false, timed = false;
  lfb_lens lens;
  DevLens dev_lens;
  float* d_tex = nullptr;  // the aperture mask the next frame reads: d_tex_ab[tex_cur]
  // lfb_set_aperture uploads into the OTHER of two textures on its own stream, so a host that changes the mask while earlier
  // frames are still in flight waits for the 1 MB copy only, not for those frames (ev_tex_free[k]: the frames that read k are done)
  float* d_tex_ab[2] = {nullptr, nullptr};
  int tex_cur = 0;
  cudaStream_t tex_stream = nullptr;
  cudaEvent_t ev_tex_up = nullptr, ev_tex_free[2] = {nullptr, nullptr};
  bool tex_free_valid[2] = {false, false};
  int tex_w = 0, tex_h = 0;
  // jobs of the current frame description (re-used while lights/params/lens do not change)
  Job* d_jobs = nullptr;
  Job* h_jobs = nullptr;  // pinned staging of the frame being built: alias of stage[stage_cur]
  // Two sets of page-locked staging buffers, used in turn: a host that enqueues frames back to back (moving lights) builds
  // frame k+1's tables while frame k's upload may still be queued behind frame k-1's kernels.  stage[b].ev = b's last upload done.
  struct HostStage {
    char* arena = nullptr;
    cudaEvent_t ev = nullptr;
    bool valid = false;
  } stage[2];
  int stage_cur = 0;
  char* d_arena = nullptr;  // all per-frame tables back to back (prepare_jobs): d_jobs, d_progs, d_slots, ... point into it
  size_t rec_off[3] = {0, 0, 0}, rec_bytes = 0;  // jobs / slots / families inside a record region; its size
  int rec_cur = 1;                                // the record region the current tables use
  cudaEvent_t ev_rec_free[2] = {nullptr, nullptr};  // on `stream`: the frames that read region r are done
  bool rec_free_valid[2] = {false, false};
  bool sweeps_aside = false;  // this structure's frames run their forward sweeps on prefix_stream (launch_exact_frame)
  size_t arena_cap = 0;
  int n_jobs = 0;
  std::vector<unsigned char> job_key;     // params + lights of the tables on the device
  std::vector<unsigned char> struct_key;  // params + light COUNT: the part the job lists and step programs depend on
  std::vector<char> job_tmpl;             // the job records of that structure (host copy)
  Job* d_dump_job = nullptr;
  // EXACT_GRID step programs, LFB_MAX_STEPS per job: StepF (LFB_FP32) or StepD (LFB_STRICT) records, buffers sized for StepD
  char* d_progs = nullptr;
  char* h_progs = nullptr;   // alias of stage[stage_cur].progs
  char* d_dump_prog = nullptr;
  bool frame_strict = false;  // the current job table holds StepD programs
  // prefix cache: one slot per (light, lambda) of the frame
  Job* d_slots = nullptr;  Job* h_slots = nullptr;
  char* d_slot_progs = nullptr;  char* h_slot_progs = nullptr;
  int n_slots = 0;
  float4* d_prefix = nullptr;
  size_t prefix_cap = 0;
  // prefix overlap: the forward sweeps of frame k+1 run on their own (high-priority) stream into the other of two caches while
  // the ghost kernel of frame k still reads its own, when a host enqueues frames back to back (options.prefix_overlap)
  float4* d_prefix2 = nullptr;
  size_t prefix2_cap = 0;
  bool frame_has_prefix2 = false;
  cudaStream_t prefix_stream = nullptr;
  cudaEvent_t ev_prefix_done[2] = {nullptr, nullptr}, ev_prefix_free[2] = {nullptr, nullptr}, ev_upload = nullptr;
  bool prefix_free_valid[2] = {false, false};
  bool upload_pending = true;  // tables / constants were (re)uploaded on `stream` since the last sweep on prefix_stream
  int prefix_flip = 0;
  size_t prefix_budget = (size_t)40 << 30;  // bytes of HBM the cache may take (options.prefix_budget_bytes)
  bool frame_has_prefix = false;
  // ghost families: one job per (light, lambda, first reflection j) of the frame
  Job* d_fams = nullptr;  Job* h_fams = nullptr;
  char* d_fam_progs = nullptr;  char* h_fam_progs = nullptr;
  int n_fams = 0;
  bool frame_has_family = false, last_families = false;
  std::vector<unsigned> job_heads, fam_heads;  // packed (slot, first reflection, program length) per ghost / family job
  float2* d_lut = nullptr;   // reflectance tables, kLutSize entries per (lambda, surface, direction)
  std::vector<float> poly;   // reflectance polynomials, kPolyN coefficients per (lambda, surface, direction)
  unsigned long long* d_stats = nullptr;  // options.collect_stats: executed steps / ray pairs started / landed of the last frame
  // owned buffers of the host-memory API
  unsigned long long* d_accum = nullptr;
  size_t accum_cap = 0;
  char* d_out = nullptr;
  size_t out_cap = 0;
  lfb_ray_hit* d_hits = nullptr;
  size_t hits_cap = 0;
  // REF_QUADS state
  RefTri* d_tris = nullptr;
  lfb_ref_ghost* d_ghosts = nullptr;
  int* d_pairs = nullptr;
  float* d_rgbw = nullptr;
  int n_ref_pairs = 0, n_ref_ghosts = 0;
  // starburst (lfb_set_starburst_aperture / lfb_render_starburst)
  float* d_star_tex = nullptr;
  int star_w = 0, star_h = 0, star_bbox[4] = {0, 0, -1, -1};
  double star_total = 0;
  char* d_star_scratch = nullptr;
  size_t star_scratch_cap = 0;
  double* d_star_lights = nullptr;
  size_t star_lights_cap = 0;
  bool star_spectrum_valid = false;  // d_star_scratch holds the lattice spectrum |F| of the current mask (options.starburst_cache)
  int star_spectrum_period = 0;
  // asynchronous host path (lfb_render_ghosts_async): the frame's device->host copy runs on its own stream out of one of
  // two device buffers, so it overlaps the next frame's trace
  cudaStream_t copy_stream = nullptr;
  char* d_out_ab[2] = {nullptr, nullptr};
  size_t out_ab_cap[2] = {0, 0};
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  bool copied_valid[2] = {false, false};
  int ab = 0;
  // display-frame path (lfb_render_frame_rgba8)
  double* d_hdr = nullptr;
  size_t hdr_cap = 0;
  uint32_t* d_rgba = nullptr;
  size_t rgba_cap = 0;
  // dirty-rectangle path (lfb_render_ghosts_rect)
  int* d_bbox = nullptr;      // device {min_x, min_y, max_x, max_y} of the pixels the frame deposits into
  int* h_bbox = nullptr;      // pinned: [0..3] reset pattern, [4..7] read-back
  bool track_bbox = false;
  int accum_dirty[4] = {0, 0, -1, -1};  // what the last rect frame left non-zero in d_accum
  int accum_dirty_w = 0, accum_dirty_h = 0;
  // tile-sparse host path (lfb_render_ghosts_sparse): the caller's buffer the engine last wrote and the tiles it left in it
  unsigned* d_sparse_state = nullptr;
  size_t sparse_state_cap = 0;
  const void* sparse_out = nullptr;
  int sparse_w = 0, sparse_h = 0, sparse_elem = 0;
  size_t sparse_stride = 0;
  unsigned* h_count = nullptr;  // page-locked, mapped: tiles written by the last sparse launch ([0]) / by slot s ([1 + s])
  // lfb_render_ghosts_sparse_begin / _end: up to LFB_SPARSE_SLOTS frames in flight.  Each slot owns an accumulator and the tile state of ITS host
  // frame; the tile kernel runs on fin_stream (highest priority), so its PCIe-bound stores overlap the next frame's trace.
  struct SparseSlot {
    unsigned long long* accum = nullptr; size_t accum_cap = 0; bool accum_clean = false; int accum_w = 0, accum_h = 0;
    unsigned* state = nullptr; size_t state_cap = 0;
    char* stage = nullptr; size_t stage_cap = 0;  // the frame's tiles on their way to host memory (sparse.cu: drain_kernel)
    const void* out = nullptr; int w = 0, h = 0, elem = 0; size_t stride = 0;
    cudaEvent_t ev_begin = nullptr, ev_traced = nullptr, ev_tile0 = nullptr, ev_staged = nullptr, ev_done = nullptr;  // timing events (lfb_sparse_slot_times)
    bool pending = false, done_valid = false;
  } slot[LFB_SPARSE_SLOTS];
  cudaStream_t fin_stream = nullptr, drain_stream = nullptr;
  float drain_gbps = -1.f;  // pace of the host-drain kernel; < 0: not decided yet (options.host_write_mbps)
  cudaEvent_t ev_epoch = nullptr;  // recorded at creation: the zero of lfb_sparse_slot_times
  bool accum_clean = false;     // d_accum (sums and bitmap) is all zeros for a accum_clean_w x accum_clean_h frame
  int accum_clean_w = 0, accum_clean_h = 0;
  SceneStore* scene = nullptr;  // lfb_set_scene
  uint64_t launches = 0;
  float last_trace_ms = 0, last_frame_ms = 0;
};

namespace {

// The parity kernels (PARAXIAL_GRID, the FP64 oracle-order EXACT_GRID, REF_QUADS) read the lens from __constant__ memory,
// one copy per device shared by every engine on it: the engine that last uploaded owns it, others re-upload before launching.
// Engines run on their own streams, so an upload must not overtake another engine's kernels that still read the previous
// lens: every launch of a constant-reading kernel records g_const_used[device] on its stream (note_const_use), and an upload
// by a non-owner first makes its stream wait for that event.  (The EXACT_GRID throughput kernels do not use the constants.)
const lfb_engine* g_const_owner[64] = {nullptr};
cudaEvent_t g_const_used[64] = {nullptr};
bool g_const_used_valid[64] = {false};
std::mutex g_const_mutex;  // engines are single-threaded, but two engines of one device may live on two threads

int bind(lfb_engine* e) {
  if (!e) return fail(LFB_ERR_INVALID, "engine is NULL");
  CU(cudaSetDevice(e->device));
  return LFB_OK;
}

int upload_constants(lfb_engine* e) {
  std::lock_guard<std::mutex> lock(g_const_mutex);
  if (e->device >= 64) return fail(LFB_ERR_INVALID, "device index >= 64");
  if (g_const_owner[e->device] == e) return LFB_OK;
  if (g_const_used_valid[e->device]) CU(cudaStreamWaitEvent(e->stream, g_const_used[e->device], 0));
  CU(upload_lens_f32(e->dev_lens, e->stream));
  CU(upload_lens_f64(e->dev_lens, e->stream));
  CU(upload_lens_ref(e->dev_lens, e->stream));
  g_const_owner[e->device] = e;
  e->upload_pending = true;
  return LFB_OK;
}

// call after enqueuing kernels that read the __constant__ lens
int note_const_use(lfb_engine* e) {
  std::lock_guard<std::mutex> lock(g_const_mutex);
  if (!g_const_used[e->device]) CU(cudaEventCreateWithFlags(&g_const_used[e->device], cudaEventDisableTiming));
  CU(cudaEventRecord(g_const_used[e->device], e->stream));
  g_const_used_valid[e->device] = true;
  return LFB_OK;
}

template <typename T>
int grow(T** p, size_t* cap, size_t need) {
  if (need <= *cap) return LFB_OK;
  if (*p) CU(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  cudaError_t err = cudaMalloc((void**)p, need);
  if (err == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(LFB_ERR_NOMEM, "cudaMalloc: out of device memory"); }
  if (err != cudaSuccess) return fail_cuda(err, "cudaMalloc");
  *cap = need;
  return LFB_OK;
}

FrameGeom make_geom(const lfb_engine* e, const lfb_params& P, unsigned* tile_bits) {
  FrameGeom g;
  memset(&g, 0, sizeof(g));
  g.W = P.width; g.H = P.height; g.N = P.grid_n; g.splat = P.splat;
  g.tiles_x = (P.grid_n + 15) / 16;
  g.tiles_per_job = g.tiles_x * g.tiles_x;
  g.tex_w = e->tex_w; g.tex_h = e->tex_h;
  g.fp_scale = ldexp(1.0, P.fixed_point_bits > 0 ? P.fixed_point_bits : 40);
  g.P = (float)e->lens.entrance_half_height; g.h_stop = (float)e->lens.stop_half_height;
  g.cell = 2.f * g.P / (float)P.grid_n;
  g.P_d = e->lens.entrance_half_height; g.cell_d = 2.0 * g.P_d / (double)P.grid_n;
  g.mask_su = 0.5f * (float)e->tex_w / g.h_stop; g.mask_ou = 0.5f * (float)e->tex_w;
  g.mask_sv = -0.5f * (float)e->tex_h / g.h_stop; g.mask_ov = 0.5f * (float)e->tex_h;
  g.lut = e->d_lut;
  g.prefix = nullptr;  // set by render_grid_device when the frame's jobs were built against the prefix cache
  g.half_rays = P.grid_n * ((P.grid_n + 1) / 2);
  g.n_surf = e->lens.n_surfaces;
  g.bbox = e->track_bbox ? e->d_bbox : nullptr;
  g.tile_bits = tile_bits;
  g.tiles_w = tiles_across(P.width);
  g.poly_v0 = e->opt.weights_table ? 2.f : kPolyV0;
  g.pad2 = e->opt.experiment;
  g.stats = e->opt.collect_stats ? e->d_stats : nullptr;
  return g;
}

bool is_exact_fast(const lfb_params& P) { return P.mode == LFB_MODE_EXACT_GRID && (P.precision == LFB_FP32 || P.precision == LFB_STRICT); }

// Polarisation-averaged reflectance of one interface as a function of s2 = sin^2(theta0): bare Fresnel, or the exact
// single-layer (Airy) film of index max(sqrt(n0 n2), 1.38) and quarter-wave thickness at lambda0.  Host-side, double:
// it feeds the polynomial fits and the tables of the throughput kernels; it is evaluated per node, never per ray.
double interface_reflectance(double n0, double n2, double s2, double lambda0, double lambda) {
  if (n0 == n2) return 0.0;
  const double cos0 = sqrt(1.0 - s2);
  const double e2 = n0 / n2, k2 = 1.0 - e2 * e2 * s2;
  if (k2 < 0) return 1.0;  // total internal reflection
  const double cos2 = sqrt(k2);
  if (lambda0 <= 0) {
    const double a = n0 * cos0 + n2 * cos2, b = n2 * cos0 + n0 * cos2;
    if (a == 0 || b == 0) return 1.0;
    const double rs = (n0 * cos0 - n2 * cos2) / a, rp = (n2 * cos0 - n0 * cos2) / b;
    return 0.5 * (rs * rs + rp * rp);
  }
  double n1 = sqrt(n0 * n2);
  if (n1 < 1.38) n1 = 1.38;
  const double e1 = n0 / n1, k1 = 1.0 - e1 * e1 * s2;
  if (k1 < 0) return 1.0;
  const double cos1 = sqrt(k1);
  const double cd = cos(3.14159265358979323846 * lambda0 * cos1 / lambda);  // 4 pi n1 d1 cos1 / lambda, d1 = lambda0 / (4 n1)
  auto ratio = [](double a, double b) { return (a + b) == 0 ? 1.0 : (a - b) / (a + b); };
  const double r01s = ratio(n0 * cos0, n1 * cos1), r12s = ratio(n1 * cos1, n2 * cos2);
  const double r01p = ratio(n1 * cos0, n0 * cos1), r12p = ratio(n2 * cos1, n1 * cos2);
  const double ps = r01s * r12s, pp = r01p * r12p;
  const double Rs = (r01s * r01s + r12s * r12s + 2 * ps * cd) / (1 + ps * ps + 2 * ps * cd);
  const double Rp = (r01p * r01p + r12p * r12p + 2 * pp * cd) / (1 + pp * pp + 2 * pp * cd);
  return 0.5 * (Rs + Rp);
}

// R of interface (lam, k, dir) over v = the cosine of the ray's angle in the RARER medium (c0 when entering the denser
// medium, c2 otherwise): R is analytic in v on [0, 1] -- R -> 1 linearly as v -> 0, at grazing incidence or at the critical
// angle -- whereas in sin^2(theta0) it has a square-root singularity at the critical angle.
struct Interface {
  double n0, n2, lambda0, lambda;
  double R(double v) const {
    const double s2 = n0 <= n2 ? 1.0 - v * v : (1.0 - v * v) * (n2 / n0) * (n2 / n0);
    return interface_reflectance(n0, n2, s2 < 0 ? 0 : s2, lambda0, lambda);
  }
};
Interface interface_of(const lfb_lens& L, int lam, int k, int dir) {
  const double na = k == 0 ? 1.0 : (double)L.ior[lam][k - 1], nb = (double)L.ior[lam][k];
  return {dir ? nb : na, dir ? na : nb, (double)L.coating_lambda0_nm[k], (double)L.lambda_nm[lam]};
}

// Tables and polynomials for every (wavelength, surface, direction): index (lam * n_surfaces + k) * 2 + (backward ? 1 : 0).
//   lut   kLutSize linear-interpolation intervals on v in [0, 1]: (R_i, R_{i+1} - R_i)
//   poly  degree kPolyN-1 interpolant of R at the Chebyshev nodes of v in [kPolyV0, 1], as monomial coefficients in
//         x = (1 - v) / (1 - kPolyV0) in [0, 1], lowest order first (|error| ~ 3e-8 for the air / glass interfaces)
void build_reflectance_tables(const lfb_lens& L, std::vector<float2>& lut, std::vector<float>& poly) {
  const size_t n_if = (size_t)L.n_lambda * L.n_surfaces * 2;
  lut.assign(n_if * kLutSize, make_float2(0.f, 0.f));
  poly.assign(n_if * kPolyN, 0.f);
  const double kPi = 3.14159265358979323846;
  for (int lam = 0; lam < L.n_lambda; lam++)
    for (int k = 0; k < L.n_surfaces; k++)
      for (int dir = 0; dir < 2; dir++) {
        const Interface I = interface_of(L, lam, k, dir);
        const size_t idx = (size_t)(lam * L.n_surfaces + k) * 2 + dir;
        float2* T = &lut[idx * kLutSize];
        double prev = I.R(0.0);
        for (int i = 0; i < kLutSize; i++) {
          const double next = I.R((double)(i + 1) / kLutSize);
          T[i] = make_float2((float)prev, (float)(next - prev));
          prev = next;
        }
        // Chebyshev interpolation on t in [-1, 1], x = (t + 1) / 2, v = 1 - x (1 - v0)
        double f[kPolyN], cheb[kPolyN];
        for (int m = 0; m < kPolyN; m++) {
          const double t = cos(kPi * (m + 0.5) / kPolyN);
          f[m] = I.R(1.0 - 0.5 * (t + 1.0) * (1.0 - (double)kPolyV0));
        }
        for (int q = 0; q < kPolyN; q++) {
          double sum = 0;
          for (int m = 0; m < kPolyN; m++) sum += f[m] * cos(kPi * q * (m + 0.5) / kPolyN);
          cheb[q] = sum * (q == 0 ? 1.0 : 2.0) / kPolyN;
        }
        // sum_q cheb[q] T_q(2x - 1) as monomials in x: T_0 = 1, T_1 = 2x - 1, T_{q+1} = 2 (2x - 1) T_q - T_{q-1}
        double mono[kPolyN] = {0}, t0[kPolyN] = {1.0}, t1[kPolyN] = {-1.0, 2.0}, t2[kPolyN];
        for (int q = 0; q < kPolyN; q++) {
          const double* tq = q == 0 ? t0 : t1;
          for (int d = 0; d < kPolyN; d++) mono[d] += cheb[q] * tq[d];
          if (q >= 1) {
            for (int d = 0; d < kPolyN; d++) t2[d] = 2.0 * ((d > 0 ? 2.0 * t1[d - 1] : 0.0) - t1[d]) - t0[d];
            memcpy(t0, t1, sizeof(t0));
            memcpy(t1, t2, sizeof(t1));
          }
        }
        for (int d = 0; d < kPolyN; d++) poly[idx * kPolyN + d] = (float)mono[d];
      }
}

// One step of a program: the ray arrives on surface k coming from surface prev_k (k = n_surfaces: the sensor plane).
// Everything ray-independent is computed here, once, in double; put_step narrows it to the kernel's geometry type.
StepD make_step(const lfb_engine* e, int lam, int k, int op, bool forward, int prev_k) {
  const lfb_lens& L = e->lens;
  const DevLens& D = e->dev_lens;
  const int n = L.n_surfaces, stop = L.stop_index;
  StepD S;
  memset(&S, 0, sizeof(S));
  S.lut = (lam * L.n_surfaces + (k < n ? k : 0)) * 2 + (forward ? 0 : 1);
  S.dz = D.zv_d[prev_k] - D.zv_d[k];
  S.eta = S.eta2 = 1.0;
  S.semi2 = INFINITY;  // the stop and the sensor are unbounded planes (the mask bounds the stop)
  if (k == n) { S.op = STEP_SENSOR; return S; }
  if (k == stop && op != STEP_REFLECT) { S.op = STEP_STOP; return S; }
  S.c = (double)L.curvature[k];
  S.semi2 = L.semi_aperture[k] * L.semi_aperture[k];
  const float na = k == 0 ? 1.f : L.ior[lam][k - 1], nb = L.ior[lam][k];
  const float n0 = forward ? na : nb, n2 = forward ? nb : na;
  S.eta = (double)n0 / (double)n2;
  S.op = (op == STEP_REFRACT && n0 == n2) ? STEP_PASS : op;
  // the weight factor: R at a reflection, 1 - R at a refraction
  const float* R = &e->poly[(size_t)S.lut * kPolyN];
  for (int d = 0; d < kPolyN; d++) S.p[d] = op == STEP_REFLECT ? R[d] : (d == 0 ? 1.f - R[d] : -R[d]);
  return S;
}

template <typename T>
void put_step(char* base, size_t index, const StepD& S) {
  StepT<T> o;
  memset(&o, 0, sizeof(o));
  o.c = (T)S.c; o.dz = (T)S.dz; o.eta = (T)S.eta; o.eta2 = o.eta * o.eta;
  o.semi2 = S.semi2; o.op = S.op; o.lut = S.lut;
  memcpy(o.p, S.p, sizeof(o.p));
  memcpy(base + index * sizeof(StepT<T>), &o, sizeof(o));
}
void put_steps(bool strict, char* base, size_t first, const StepD* steps, int n) {
  for (int s = 0; s < n; s++) {
    if (strict) put_step<double>(base, first + (size_t)s, steps[s]);
    else put_step<float>(base, first + (size_t)s, steps[s]);
  }
}

// Flatten ghost (i, j) at wavelength lam into a step program (exact_trace.cuh): the surface sequence forward 0..j-1,
// reflect at j, backward j-1..i+1, reflect at i, forward i+1..n-1, sensor plane (i < 0: the direct path).
// from_reflection: the program starts ON surface j with the first reflection (the forward sweep 0 .. j-1 comes from the
// prefix cache).  prefix_only: the forward sweep 0 .. n-1 followed by the sensor step (prefix_kernel traces the first n
// steps; the sensor step serves the direct path and the forks of the family kernel); returns n.
int build_program(const lfb_engine* e, int lam, int i, int j, StepD* out, bool from_reflection = false, bool prefix_only = false) {
  const int n = e->lens.n_surfaces;
  int ns = 0, prev = from_reflection ? j : 0;
  auto push = [&](int k, int op, bool forward) {
    out[ns++] = make_step(e, lam, k, op, forward, prev);
    prev = k;
  };
  if (prefix_only) {
    for (int k = 0; k < n; k++) push(k, STEP_REFRACT, true);
    push(n, STEP_SENSOR, true);
    return n;
  }
  if (i < 0) {
    for (int k = 0; k < n; k++) push(k, STEP_REFRACT, true);
  } else {
    if (!from_reflection)
      for (int k = 0; k < j; k++) push(k, STEP_REFRACT, true);
    push(j, STEP_REFLECT, true);
    for (int k = j - 1; k > i; k--) push(k, STEP_REFRACT, false);
    push(i, STEP_REFLECT, false);
    for (int k = i + 1; k < n; k++) push(k, STEP_REFRACT, true);
  }
  push(n, STEP_SENSOR, true);
  return ns;
}

// The program of a ghost family: first reflection at j, then for k = j-1 .. 0 the backward step at k followed by the
// fork step (reflection at k seen from behind; op = -1 when ghost (k, j) is not in `mask`).
int build_family_program(const lfb_engine* e, int lam, int j, unsigned mask, StepD* out) {
  int ns = 0, prev = j, kmin = 0;
  while (kmin < j && !((mask >> kmin) & 1u)) kmin++;  // the backward sweep ends at the lowest wanted second reflection
  out[ns++] = make_step(e, lam, j, STEP_REFLECT, true, j);
  for (int k = j - 1; k >= kmin; k--) {
    out[ns++] = make_step(e, lam, k, STEP_REFRACT, false, prev);
    prev = k;
    StepD fork = make_step(e, lam, k, STEP_REFLECT, false, k);
    if (!((mask >> k) & 1u) || k == e->lens.stop_index) fork.op = -1;
    out[ns++] = fork;
  }
  return ns;
}

// Frame constants of one job.  The libm calls (atan/cosf/sinf of frame constants,
// pathtracer.cpp:414, and sin/cos of the light's angle) are made here once per job, not per ray.
// The fields of a Job that depend on its LIGHT (and, through the colour weights, its wavelength): everything else is the
// frame's structure (lens, pairs, grid, shard) and survives a change of the lights alone.
struct LightFields {
  float theta;
  double cs, sn, sx, sy, ppu, sin_t, cos_t, inv_dist, area, scale;
};
LightFields light_fields(const lfb_engine* e, const lfb_params& P, const lfb_light& lt) {
  LightFields F;
  F.theta = lt.theta;
  const double dx = lt.ns_x - 0.5, dy = lt.ns_y - 0.5;
  if (P.physical_mapping) {
    // the flare axis runs from the image centre through the light: the rotation is the light's azimuth (atan2, not the
    // reference's atan, which folds left-of-centre suns onto the right) and the origin is the image centre, so that a ray
    // landing at local (xs, ys) is drawn at centre + ppu * R(phi) (xs, ys).  See lfb_params.physical_mapping.
    const double phi = (dx == 0 && dy == 0) ? 0.0 : atan2(dy * (double)P.height, dx * (double)P.width);
    F.cs = -cos(phi); F.sn = -sin(phi);  // to_pixel maps X = -ppu xs: fold the sign into the rotation
    F.sx = 0.5 * (double)P.width;
    F.sy = 0.5 * (double)P.height;
  } else {
    const float ang = (dx == 0 && dy == 0) ? 0.f : (float)atan(dy / dx);
    F.cs = cosf(ang); F.sn = sinf(ang);
    F.sx = ceil(lt.ns_x * (double)P.width);   // draw_ghost, pathtracer.cpp:462-463
    F.sy = ceil(lt.ns_y * (double)P.height);
  }
  F.ppu = P.px_per_unit > 0 ? P.px_per_unit : 0.4f;
  F.sin_t = sin((double)lt.theta); F.cos_t = cos((double)lt.theta);
  F.inv_dist = (lt.distance > 0 && std::isfinite(lt.distance)) ? 1.0 / lt.distance : 0.0;  // 0: directional
  const double cell = 2 * e->lens.entrance_half_height / P.grid_n;
  F.area = cell * cell * F.ppu * F.ppu;
  F.scale = ldexp(1.0, P.fixed_point_bits > 0 ? P.fixed_point_bits : 40);
  return F;
}
void apply_light(const lfb_engine* e, const lfb_light& lt, const LightFields& F, Job* J) {
  J->theta = F.theta;
  J->cs = F.cs; J->sn = F.sn; J->sx = F.sx; J->sy = F.sy; J->ppu = F.ppu;
  J->sin_t = F.sin_t; J->cos_t = F.cos_t; J->inv_dist = F.inv_dist;
  for (int c = 0; c < 3; c++) J->chan[c] = (double)lt.radiance[c] * (double)e->lens.rgb_weight[J->lambda][c] * F.area;
  J->f_sin_t = (float)J->sin_t; J->f_cos_t = (float)J->cos_t; J->f_inv_dist = (float)J->inv_dist;
  J->f_sx = (float)J->sx; J->f_sy = (float)J->sy; J->f_cs = (float)J->cs; J->f_sn = (float)J->sn; J->f_ppu = (float)J->ppu;
  for (int c = 0; c < 3; c++) J->f_chan[c] = (float)J->chan[c] * (float)F.scale;
}
void fill_job(const lfb_engine* e, const lfb_params& P, const lfb_light& lt, const JobId& id, Job* J) {
  memset(J, 0, sizeof(*J));
  J->light = id.light; J->i = id.i; J->j = id.j; J->lambda = id.lambda;
  apply_light(e, lt, light_fields(e, P, lt), J);
}

// the job / slot / family records the next launches read: region r of the arena's two
void set_record_region(lfb_engine* e, int r) {
  char* base = e->d_arena + (size_t)r * e->rec_bytes;
  e->d_jobs = (Job*)(base + e->rec_off[0]);
  e->d_slots = (Job*)(base + e->rec_off[1]);
  e->d_fams = (Job*)(base + e->rec_off[2]);
  e->rec_cur = r;
}

// Build + upload this shard's tables unless the frame description is unchanged.  All tables of a frame -- ghost jobs and their
// step programs, prefix slots and theirs, family jobs and theirs -- are laid out back to back in ONE arena and go up in ONE
// copy: six separate small copies cost ~8 us of stream time each (measured: 0.05 ms of a 0.26 ms cfg2 frame with a moving sun).
int prepare_jobs(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params& P, bool capturing) {
  std::vector<unsigned char> key(sizeof(lfb_params) + sizeof(lfb_light) * (size_t)n_lights);
  memcpy(key.data(), &P, sizeof(lfb_params));
  if (n_lights > 0) memcpy(key.data() + sizeof(lfb_params), lights, sizeof(lfb_light) * (size_t)n_lights);
  if (key == e->job_key) return LFB_OK;
  // Only the lights moved (the usual frame-to-frame change): the structure -- job lists, step programs, kernel choice, arena
  // layout -- stands, and so do the programs already on the device.  Refill the light-dependent fields of the job records
  // from the saved templates and upload just those (cfg2: 27 KB instead of 300 KB, ~8 us of host time instead of ~50).
  std::vector<unsigned char> skey(sizeof(lfb_params) + sizeof(int));
  memcpy(skey.data(), &P, sizeof(lfb_params));
  memcpy(skey.data() + sizeof(lfb_params), &n_lights, sizeof(int));
  if (!e->job_key.empty() && skey == e->struct_key) {
    e->stage_cur ^= 1;
    lfb_engine::HostStage& st = e->stage[e->stage_cur];
    if (st.valid) CU(cudaEventSynchronize(st.ev));
    const size_t bytes = e->job_tmpl.size();
    // With the forward sweeps on their own stream (launch_exact_frame), the records go up on THAT stream, into the record
    // region the previous frame is not using: upload and sweeps of frame k+1 then run under the ghost kernel of frame k.
    const bool aside = e->sweeps_aside && !capturing;
    cudaStream_t up_stream = aside ? e->prefix_stream : e->stream;
    if (bytes > 0) {
      const int r = e->rec_cur ^ 1;
      memcpy(st.arena, e->job_tmpl.data(), bytes);
      std::vector<LightFields> LF((size_t)n_lights);
      for (int l = 0; l < n_lights; l++) LF[l] = light_fields(e, P, lights[l]);
      Job* const arrays[3] = {(Job*)(st.arena + e->rec_off[0]), (Job*)(st.arena + e->rec_off[1]), (Job*)(st.arena + e->rec_off[2])};
      const int counts[3] = {e->n_jobs, e->n_slots, e->n_fams};
      for (int a = 0; a < 3; a++)
        for (int q = 0; q < counts[a]; q++) apply_light(e, lights[arrays[a][q].light], LF[arrays[a][q].light], &arrays[a][q]);
      if (aside && e->rec_free_valid[r]) CU(cudaStreamWaitEvent(up_stream, e->ev_rec_free[r], 0));
      CU(cudaMemcpyAsync(e->d_arena + (size_t)r * bytes, st.arena, bytes, cudaMemcpyHostToDevice, up_stream));
      set_record_region(e, r);
      if (e->n_jobs > 0 && P.mode == LFB_MODE_PARAXIAL_GRID) {
        CU(launch_paraxial_setup(e->d_jobs, e->n_jobs, P.physical_backward, e->stream));
        e->launches++;
        const int rcu = note_const_use(e);
        if (rcu) return rcu;
      }
    }
    CU(cudaEventRecord(st.ev, up_stream));
    st.valid = true;
    e->job_key.swap(key);
    if (!aside) e->upload_pending = true;
    return LFB_OK;
  }
  std::vector<JobId> ids;
  list_jobs(e->lens, P, n_lights, ids);
  // launch order = list order.  (Measured on cfg2: sorting the ghosts longest-first, or interleaving long and short
  // ones, is 20 % SLOWER than the reference order; the sensor sums are integers, so the order never changes the result.)
  const int n = (int)ids.size();
  int rc;
  const bool want_progs = is_exact_fast(P);
  const bool strict = P.precision == LFB_STRICT;
  e->frame_strict = strict;
  const size_t step_bytes = strict ? sizeof(StepD) : sizeof(StepF);
  const size_t prog_bytes = step_bytes * LFB_MAX_STEPS;  // per job
  const int parts = strict ? exact_prefix_parts<double>() : exact_prefix_parts<float>();
  const bool want_family = e->opt.kernel_select != 1;
  StepD prog[LFB_MAX_STEPS];
  // prefix cache: one slot per (light, lambda) that has ghost jobs in this shard
  std::vector<int> slot_of;   // [light * n_lambda + lambda] -> slot or -1
  std::vector<JobId> slot_ids;
  e->frame_has_prefix = false;
  e->frame_has_prefix2 = false;
  e->frame_has_family = false;
  e->n_fams = 0; e->n_slots = 0; e->n_jobs = 0;
  e->job_heads.clear(); e->fam_heads.clear();
  if (want_progs && e->opt.prefix_budget_bytes >= 0) {
    slot_of.assign((size_t)std::max(n_lights, 1) * e->lens.n_lambda, -1);
    for (int q = 0; q < n; q++)
      if (ids[q].i >= 0 || want_family) {  // with families the direct path is splatted by the prefix kernel itself
        int& sl = slot_of[(size_t)ids[q].light * e->lens.n_lambda + ids[q].lambda];
        if (sl < 0) { sl = (int)slot_ids.size(); slot_ids.push_back({ids[q].light, -1, -1, ids[q].lambda}); }
      }
    // The decision must not depend on the shard (a sharded frame and the whole frame pick the same kernels): it is made on
    // the cache the UNSHARDED frame would need.  (The frames have the same bits either way -- a job without a cached sweep
    // re-traces it with the same arithmetic -- this keeps the performance model simple.)
    const size_t half_rays = (size_t)P.grid_n * ((P.grid_n + 1) / 2);
    const size_t per_slot = (size_t)e->lens.n_surfaces * parts * half_rays * sizeof(float4);
    const size_t need_whole = (size_t)std::max(n_lights, 1) * e->lens.n_lambda * per_slot;
    const size_t need = slot_ids.size() * per_slot;
    if (!slot_ids.empty() && need_whole <= e->prefix_budget) {
      rc = grow(&e->d_prefix, &e->prefix_cap, need);
      if (rc == LFB_OK) e->frame_has_prefix = true;
      else if (rc != LFB_ERR_NOMEM) return rc;
      if (e->frame_has_prefix && e->opt.prefix_overlap >= 0 && e->prefix_stream && 2 * need_whole <= e->prefix_budget) {
        rc = grow(&e->d_prefix2, &e->prefix2_cap, need);
        if (rc == LFB_OK) e->frame_has_prefix2 = true;
        else if (rc != LFB_ERR_NOMEM) return rc;
      }
    }
  }
  const int ns = e->frame_has_prefix ? (int)slot_ids.size() : 0;
  // ghost families: the pairs (i, j) of a slot grouped by their first reflection j
  struct Fam { int slot, j; unsigned mask; };
  std::vector<Fam> fams;
  std::vector<char> slot_direct((size_t)ns, 0);  // this shard owns the slot's direct path
  bool use_fam = false;
  if (e->frame_has_prefix && want_family) {
    std::vector<int> fam_of((size_t)ns * LFB_MAX_SURFACES, -1);
    for (int q = 0; q < n; q++) {
      const int sl = slot_of[(size_t)ids[q].light * e->lens.n_lambda + ids[q].lambda];
      if (ids[q].i < 0) { slot_direct[sl] = 1; continue; }
      int& f = fam_of[(size_t)sl * LFB_MAX_SURFACES + ids[q].j];
      if (f < 0) { f = (int)fams.size(); fams.push_back({sl, ids[q].j, 0u}); }
      fams[f].mask |= 1u << ids[q].i;
    }
    // options.family_split = m: at most m forks per family job (more, shorter CTAs for small frames; the part of the
    // backward sweep above a chunk's forks is then repeated per chunk)
    if (e->opt.family_split > 0) {
      const int m = e->opt.family_split;
      std::vector<Fam> split;
      for (const Fam& f : fams) {
        unsigned cur = 0;
        int cnt = 0;
        for (int k = f.j - 1; k >= 0; k--) {
          if (!((f.mask >> k) & 1u)) continue;
          cur |= 1u << k;
          if (++cnt == m) { split.push_back({f.slot, f.j, cur}); cur = 0; cnt = 0; }
        }
        if (cnt) split.push_back({f.slot, f.j, cur});
      }
      fams.swap(split);
    }
    // Families do ~25 % less work but in 4x fewer, longer-lived CTAs: they win once the grid is many waves deep (cfg3 / cfg4:
    // x1.3), and lose ~8 % on a frame as small as cfg2 (5 376 CTAs = 3.6 waves), where the per-pair kernel keeps the SMs
    // fuller.  Decided here so that only the tables of the kernel that will run are built and uploaded.
    const long long fam_ctas = (long long)fams.size() * ((P.grid_n + 15) / 16) * (((P.grid_n + 1) / 2 + 7) / 8);
    use_fam = e->opt.kernel_select == 2 || fam_ctas >= 16384;
  }
  const int nf = use_fam ? (int)fams.size() : 0;
  const int n_build = use_fam ? 0 : n;  // a frame traced by families needs no per-pair jobs
  // ---- the arena: [jobs | slots | families || programs | slot programs | family programs], each 256-byte aligned; the job
  // records come first so that a change of the lights alone re-uploads only them ----
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  // (host staging: [records | programs]; device: [records 0 | records 1 | programs] -- two record regions, used in turn)
  const size_t off_jobs = 0;
  const size_t off_slots = up(off_jobs + sizeof(Job) * (size_t)n_build);
  const size_t off_fams = up(off_slots + sizeof(Job) * (size_t)ns);
  const size_t off_progs = up(off_fams + sizeof(Job) * (size_t)nf);  // = bytes of one record region
  const size_t off_slot_progs = up(off_progs + (want_progs ? prog_bytes * (size_t)n_build : 0));
  const size_t off_fam_progs = up(off_slot_progs + prog_bytes * (size_t)ns);
  const size_t host_total = up(off_fam_progs + prog_bytes * (size_t)nf);
  const size_t total = host_total + off_progs;
  if (total > e->arena_cap) {
    CU(cudaStreamSynchronize(e->stream));  // uploads out of the old staging buffers may still be queued, kernels may read the old tables
    if (e->prefix_stream) CU(cudaStreamSynchronize(e->prefix_stream));
    cudaFree(e->d_arena);
    e->d_arena = nullptr; e->arena_cap = 0;
    e->rec_free_valid[0] = e->rec_free_valid[1] = false;
    for (int b = 0; b < 2; b++) {
      if (e->stage[b].arena) cudaFreeHost(e->stage[b].arena);
      e->stage[b].arena = nullptr;
      e->stage[b].valid = false;
    }
    const size_t cap = total + total / 4;
    cudaError_t err = cudaMalloc((void**)&e->d_arena, cap);
    if (err == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(LFB_ERR_NOMEM, "cudaMalloc: out of device memory"); }
    if (err != cudaSuccess) return fail_cuda(err, "cudaMalloc");
    for (int b = 0; b < 2; b++) CU(cudaHostAlloc((void**)&e->stage[b].arena, cap, cudaHostAllocDefault));
    e->arena_cap = cap;
  }
  // Two page-locked staging arenas, used in turn: a host that enqueues frames back to back (moving lights) builds frame k+1's
  // tables while frame k's upload may still be queued behind frame k-1's kernels; the other arena's upload is two frames old.
  e->stage_cur ^= 1;
  lfb_engine::HostStage& st = e->stage[e->stage_cur];
  if (st.valid) CU(cudaEventSynchronize(st.ev));
  char* const h = st.arena;
  e->rec_off[0] = off_jobs; e->rec_off[1] = off_slots; e->rec_off[2] = off_fams; e->rec_bytes = off_progs;
  e->h_jobs = (Job*)(h + off_jobs); e->h_slots = (Job*)(h + off_slots); e->h_fams = (Job*)(h + off_fams);
  e->h_progs = h + off_progs; e->h_slot_progs = h + off_slot_progs; e->h_fam_progs = h + off_fam_progs;
  // a full build fills record region 1, so that [records | programs] of the staging arena go up as one contiguous copy
  set_record_region(e, 1);
  e->d_progs = e->d_arena + off_progs + off_progs;
  e->d_slot_progs = e->d_arena + off_slot_progs + off_progs;
  e->d_fam_progs = e->d_arena + off_fam_progs + off_progs;
  for (int sl = 0; sl < ns; sl++) {
    fill_job(e, P, lights[slot_ids[sl].light], slot_ids[sl], &e->h_slots[sl]);
    e->h_slots[sl].n_steps = build_program(e, slot_ids[sl].lambda, -1, -1, prog, false, true);
    put_steps(strict, e->h_slot_progs, (size_t)sl * LFB_MAX_STEPS, prog, e->h_slots[sl].n_steps + 1);
    e->h_slots[sl].i = slot_direct[sl];  // 1: the prefix kernel of a family frame also splats the slot's direct path
  }
  for (int f = 0; f < nf; f++) {
    const JobId& sid = slot_ids[fams[f].slot];
    Job* FJ = &e->h_fams[f];
    fill_job(e, P, lights[sid.light], sid, FJ);
    FJ->slot = fams[f].slot;
    FJ->j_first = fams[f].j;
    FJ->i = (int)fams[f].mask;
    FJ->j = fams[f].j;
    FJ->n_steps = build_family_program(e, sid.lambda, fams[f].j, fams[f].mask, prog);
    e->fam_heads.push_back(pack_head(FJ->slot, FJ->j_first, FJ->n_steps));
    put_steps(strict, e->h_fam_progs, (size_t)f * LFB_MAX_STEPS, prog, FJ->n_steps);
  }
  e->n_fams = nf;
  e->frame_has_family = use_fam;
  e->n_slots = ns;
  for (int q = 0; q < n_build; q++) {
    fill_job(e, P, lights[ids[q].light], ids[q], &e->h_jobs[q]);
    e->h_jobs[q].slot = -1;
    e->h_jobs[q].j_first = ids[q].j;  // jobs without a cached sweep re-trace it and canonicalise the state there
    if (want_progs) {
      const bool cached = e->frame_has_prefix && ids[q].i >= 0;
      if (cached) e->h_jobs[q].slot = slot_of[(size_t)ids[q].light * e->lens.n_lambda + ids[q].lambda];
      e->h_jobs[q].n_steps = build_program(e, ids[q].lambda, ids[q].i, ids[q].j, prog, cached);
      e->job_heads.push_back(pack_head(e->h_jobs[q].slot, ids[q].i < 0 ? -1 : ids[q].j, e->h_jobs[q].n_steps));
      put_steps(strict, e->h_progs, (size_t)q * LFB_MAX_STEPS, prog, e->h_jobs[q].n_steps);
    }
  }
  e->n_jobs = n_build;
  if (host_total > 0) CU(cudaMemcpyAsync(e->d_arena + off_progs, h, host_total, cudaMemcpyHostToDevice, e->stream));
  if (n_build > 0 && P.mode == LFB_MODE_PARAXIAL_GRID) {
    CU(launch_paraxial_setup(e->d_jobs, n_build, P.physical_backward, e->stream));
    e->launches++;
    const int rcu = note_const_use(e);
    if (rcu) return rcu;
  }
  CU(cudaEventRecord(st.ev, e->stream));
  st.valid = true;
  e->job_tmpl.assign(h, h + (host_total > 0 ? off_progs : 0));  // the job records, for the lights-only path above
  e->sweeps_aside = e->n_slots > 0 && e->frame_has_prefix && e->frame_has_prefix2 && !e->frame_has_family && !e->opt.collect_stats;
  e->job_key.swap(key);
  e->struct_key.swap(skey);
  e->upload_pending = true;
  return LFB_OK;
}

// The EXACT_GRID throughput kernels of one frame, for geometry type T.
template <typename T>
int launch_exact_frame(lfb_engine* e, const lfb_params& P, FrameGeom& g, unsigned long long* accum, bool capturing) {
  typedef StepT<T> S;
  const bool stats = e->opt.collect_stats != 0;
  const bool families = e->frame_has_prefix && e->frame_has_family;  // decided by prepare_jobs
  e->last_families = families;
  int overlap_buf = -1;
  if (e->n_slots > 0 && e->frame_has_prefix) {
    if (!families && !capturing && !stats && e->frame_has_prefix2) {
      // forward sweeps on their own stream, alternating caches: they do not touch the sensor, so the sweeps of this frame may
      // run while the previous frame's ghost kernel (reading the other cache) is still draining
      const int b = e->prefix_flip;
      e->prefix_flip ^= 1;
      float4* cache = b ? e->d_prefix2 : e->d_prefix;
      if (e->upload_pending) {  // the tables this sweep reads were just written on `stream`
        CU(cudaEventRecord(e->ev_upload, e->stream));
        CU(cudaStreamWaitEvent(e->prefix_stream, e->ev_upload, 0));
        e->upload_pending = false;
      }
      if (e->prefix_free_valid[b]) CU(cudaStreamWaitEvent(e->prefix_stream, e->ev_prefix_free[b], 0));
      CU(launch_exact_prefix<T>(e->d_slots, (const S*)e->d_slot_progs, e->n_slots, g, e->d_tex, cache, nullptr, stats, e->prefix_stream));
      CU(cudaEventRecord(e->ev_prefix_done[b], e->prefix_stream));
      CU(cudaStreamWaitEvent(e->stream, e->ev_prefix_done[b], 0));
      g.prefix = cache;
      overlap_buf = b;
    } else {
      // cache 0 on the engine stream; a sweep still running on prefix_stream from an earlier (overlapped) frame writes its own
      // cache and was already joined by that frame's ghost kernel, which precedes this launch in stream order
      CU(launch_exact_prefix<T>(e->d_slots, (const S*)e->d_slot_progs, e->n_slots, g, e->d_tex, e->d_prefix, families ? accum : nullptr, stats, e->stream));
      g.prefix = e->d_prefix;
      overlap_buf = 0;
    }
    e->launches++;
  }
  if (families) {  // one job per (light, lambda, first reflection); the direct path was splatted by the prefix kernel
    if (e->n_fams > 0) {
      CU(launch_exact_families<T>(e->d_fams, (const S*)e->d_fam_progs, e->fam_heads.data(), e->n_fams, e->d_slots, (const S*)e->d_slot_progs, g, e->d_tex, accum, e->opt.ctas_per_sm, stats, e->stream));
      e->launches++;
    }
  } else if (e->n_jobs > 0) {
    CU(launch_exact_ghosts<T>(e->d_jobs, (const S*)e->d_progs, e->job_heads.data(), e->n_jobs, g, e->d_tex, accum, e->opt.ctas_per_sm, stats, e->stream));
    e->launches++;
  }
  if (overlap_buf >= 0) {  // the cache may be rewritten once this frame's kernels are done with it
    CU(cudaEventRecord(e->ev_prefix_free[overlap_buf], e->stream));
    e->prefix_free_valid[overlap_buf] = true;
  }
  return LFB_OK;
}

int render_grid_device(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params& P,
                       unsigned long long* accum, int clear_first) {
  int rc = upload_constants(e);
  if (rc) return rc;
  // inside a CUDA-graph capture (a host may capture whole frames and replay them) the timing events are not recorded:
  // they could not be read back anyway
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  CU(cudaStreamIsCapturing(e->stream, &cap));
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  rc = prepare_jobs(e, lights, n_lights, P, capturing);
  if (rc) return rc;
  const AccumLayout lay = accum_layout(P.width, P.height);
  FrameGeom g = make_geom(e, P, (unsigned*)((char*)accum + lay.bits_off));
  if (clear_first) CU(cudaMemsetAsync(accum, 0, lay.total, e->stream));
  if (g.stats) CU(cudaMemsetAsync(e->d_stats, 0, 4 * sizeof(unsigned long long), e->stream));
  if (!capturing) CU(cudaEventRecord(e->ev_trace0, e->stream));
  if (is_exact_fast(P)) {
    rc = P.precision == LFB_STRICT ? launch_exact_frame<double>(e, P, g, accum, capturing) : launch_exact_frame<float>(e, P, g, accum, capturing);
    if (rc) return rc;
  } else if (e->n_jobs > 0) {
    if (P.precision == LFB_FP64) CU(launch_trace_splat_f64(e->d_jobs, e->n_jobs, g, P.mode, e->d_tex, accum, e->stream));
    else CU(launch_trace_splat_f32(e->d_jobs, e->n_jobs, g, P.mode, e->d_tex, accum, e->stream));
    e->launches++;
    rc = note_const_use(e);
    if (rc) return rc;
  }
  if (!capturing) {
    CU(cudaEventRecord(e->ev_trace1, e->stream));
    e->timed = true;
    CU(cudaEventRecord(e->ev_rec_free[e->rec_cur], e->stream));  // (a lights-only upload into this region waits for it)
    e->rec_free_valid[e->rec_cur] = true;
  }
  return LFB_OK;
}

int render_ref_device(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params& P, void* out_dev,
                      size_t stride, int elem, int additive) {
  int rc = upload_constants(e);
  if (rc) return rc;
  RefFrame f;
  memset(&f, 0, sizeof(f));
  f.W = P.width; f.H = P.height; f.tex_w = e->tex_w; f.tex_h = e->tex_h;
  f.n_pairs = e->n_ref_pairs; f.n_lambda = e->lens.n_lambda;
  e->n_ref_ghosts = 0;
  if (n_lights > 0) {
    const lfb_light& lt = lights[n_lights - 1];  // the last on-screen light wins, pathtracer.cpp:50-53
    f.has_sun = !(lt.ns_x == 0 && lt.ns_y == 0);  // :724-726
    f.angle_to_sun = lt.theta;
    const float ang = (float)atan((lt.ns_y - 0.5) / (lt.ns_x - 0.5));  // shift_vertex :414
    f.cs = cosf(ang); f.sn = sinf(ang);
    f.gb_mid_w = ceil(lt.ns_x * (double)P.width);
    f.gb_mid_h = ceil(lt.ns_y * (double)P.height);
  }
  CU(cudaEventRecord(e->ev_trace0, e->stream));
  const int n_ghosts = f.n_pairs * f.n_lambda;
  if (f.has_sun && n_ghosts > 0) {
    CU(launch_ref_setup(f, e->d_pairs, e->d_rgbw, e->d_tris, e->d_ghosts, e->track_bbox ? e->d_bbox : nullptr, e->stream));
    e->launches++;
    e->n_ref_ghosts = n_ghosts;
  }
  int rect[4];
  if (e->track_bbox) {  // rect path: raster only the ghosts' bounding rectangle, into a packed buffer
    CU(cudaMemcpyAsync(e->h_bbox + 4, e->d_bbox, 4 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    memcpy(rect, e->h_bbox + 4, sizeof(rect));
    if (rect[2] < rect[0] || rect[3] < rect[1]) { CU(cudaEventRecord(e->ev_trace1, e->stream)); e->timed = true; return note_const_use(e); }
  }
  CU(launch_ref_raster(f, e->d_tris, f.has_sun ? 2 * n_ghosts : 0, e->d_tex, out_dev, stride, elem, additive,
                       e->track_bbox ? rect : nullptr, e->stream));
  e->launches++;
  rc = note_const_use(e);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_trace1, e->stream));
  e->timed = true;
  return LFB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------
extern "C" int lfb_abi_version(void) { return LFB_ABI_VERSION; }

extern "C" const char* lfb_last_error(void) { return g_err.c_str(); }

extern "C" int lfb_create(lfb_engine** out, int device_id) { return lfb_create_ex(out, device_id, nullptr); }

extern "C" int lfb_create_ex(lfb_engine** out, int device_id, const lfb_options* options) {
  if (!out) return fail(LFB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  lfb_options opt;
  memset(&opt, 0, sizeof(opt));
  if (options) {
    if (options->struct_size < (int32_t)sizeof(int32_t) || options->struct_size > (int32_t)sizeof(lfb_options))
      return fail(LFB_ERR_INVALID, "lfb_options.struct_size must be sizeof(lfb_options) of the caller's header");
    memcpy(&opt, options, (size_t)options->struct_size);
  }
  opt.struct_size = (int32_t)sizeof(lfb_options);
  if (opt.kernel_select < 0 || opt.kernel_select > 2) return fail(LFB_ERR_INVALID, "lfb_options.kernel_select must be 0, 1 or 2");
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || n < 1) {
    cudaGetLastError();
    return fail(LFB_ERR_NO_DEVICE, std::string("no CUDA device (") + (err != cudaSuccess ? cudaGetErrorString(err) : "device count 0") +
                                       "); this engine has no CPU fallback");
  }
  if (device_id < 0) CU(cudaGetDevice(&device_id));
  if (device_id >= n) return fail(LFB_ERR_INVALID, "device_id out of range");
  CU(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10) return fail(LFB_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + "; the kernels are built for sm_100a only");
  lfb_engine* e = new (std::nothrow) lfb_engine();
  if (!e) return fail(LFB_ERR_NOMEM, "out of host memory");
  e->device = device_id;
  e->opt = opt;
  if (opt.prefix_budget_bytes > 0) e->prefix_budget = (size_t)opt.prefix_budget_bytes;
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  const int prio = opt.stream_priority == 1 ? prio_hi : prio_lo;
  cudaError_t rc = cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio);
  if (rc == cudaSuccess) rc = cudaStreamCreateWithPriority(&e->prefix_stream, cudaStreamNonBlocking, prio_hi);
  for (int k = 0; k < 2 && rc == cudaSuccess; k++) {
    rc = cudaEventCreateWithFlags(&e->ev_prefix_done[k], cudaEventDisableTiming);
    if (rc == cudaSuccess) rc = cudaEventCreateWithFlags(&e->ev_prefix_free[k], cudaEventDisableTiming);
  }
  if (rc == cudaSuccess) rc = cudaEventCreateWithFlags(&e->ev_upload, cudaEventDisableTiming);
  for (int k = 0; k < 2 && rc == cudaSuccess; k++) rc = cudaEventCreateWithFlags(&e->stage[k].ev, cudaEventDisableTiming);
  for (int k = 0; k < 2 && rc == cudaSuccess; k++) rc = cudaEventCreateWithFlags(&e->ev_rec_free[k], cudaEventDisableTiming);
  if (rc == cudaSuccess) rc = cudaStreamCreateWithFlags(&e->tex_stream, cudaStreamNonBlocking);
  if (rc == cudaSuccess) rc = cudaStreamCreateWithPriority(&e->fin_stream, cudaStreamNonBlocking, prio_hi);
  if (rc == cudaSuccess) rc = cudaStreamCreateWithPriority(&e->drain_stream, cudaStreamNonBlocking, prio_hi);
  for (int k = 0; k < LFB_SPARSE_SLOTS && rc == cudaSuccess; k++) {
    rc = cudaEventCreate(&e->slot[k].ev_traced);
    if (rc == cudaSuccess) rc = cudaEventCreate(&e->slot[k].ev_done);
    if (rc == cudaSuccess) rc = cudaEventCreate(&e->slot[k].ev_begin);
    if (rc == cudaSuccess) rc = cudaEventCreate(&e->slot[k].ev_tile0);
    if (rc == cudaSuccess) rc = cudaEventCreate(&e->slot[k].ev_staged);
  }
  if (rc == cudaSuccess) rc = cudaEventCreate(&e->ev_epoch);
  if (rc == cudaSuccess) rc = cudaEventRecord(e->ev_epoch, e->stream);
  if (rc == cudaSuccess) rc = cudaEventCreateWithFlags(&e->ev_tex_up, cudaEventDisableTiming);
  for (int k = 0; k < 2 && rc == cudaSuccess; k++) rc = cudaEventCreateWithFlags(&e->ev_tex_free[k], cudaEventDisableTiming);
  if (rc == cudaSuccess) rc = cudaEventCreate(&e->ev_frame0);
  if (rc == cudaSuccess) rc = cudaEventCreate(&e->ev_trace0);
  if (rc == cudaSuccess) rc = cudaEventCreate(&e->ev_trace1);
  if (rc == cudaSuccess) rc = cudaEventCreate(&e->ev_frame1);
  if (rc == cudaSuccess) rc = cudaMalloc((void**)&e->d_dump_job, sizeof(Job));
  if (rc == cudaSuccess) rc = cudaMalloc((void**)&e->d_dump_prog, sizeof(StepD) * LFB_MAX_STEPS);
  if (rc == cudaSuccess) rc = cudaMalloc((void**)&e->d_bbox, 4 * sizeof(int));
  if (rc == cudaSuccess) rc = cudaMalloc((void**)&e->d_stats, 4 * sizeof(unsigned long long));
  if (rc == cudaSuccess) rc = cudaMemset(e->d_stats, 0, 4 * sizeof(unsigned long long));
  if (rc == cudaSuccess) rc = cudaHostAlloc((void**)&e->h_bbox, 8 * sizeof(int), cudaHostAllocDefault);
  if (rc == cudaSuccess) rc = cudaHostAlloc((void**)&e->h_count, (1 + LFB_SPARSE_SLOTS) * sizeof(unsigned), cudaHostAllocMapped);
  if (rc == cudaSuccess) e->h_count[0] = 0;
  if (rc == cudaSuccess) { e->h_bbox[0] = e->h_bbox[1] = 0x7fffffff; e->h_bbox[2] = e->h_bbox[3] = -0x7fffffff; }
  if (rc != cudaSuccess) { lfb_destroy(e); return fail_cuda(rc, "lfb_create"); }
  *out = e;
  return LFB_OK;
}

extern "C" void lfb_destroy(lfb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (e->device < 64 && g_const_owner[e->device] == e) g_const_owner[e->device] = nullptr;
  }
  cudaFree(e->d_arena); cudaFree(e->d_dump_prog); cudaFree(e->d_bbox); cudaFree(e->d_lut); cudaFree(e->d_stats);
  cudaFree(e->d_prefix); cudaFree(e->d_prefix2);
  if (e->prefix_stream) { cudaStreamSynchronize(e->prefix_stream); cudaStreamDestroy(e->prefix_stream); }
  for (int k = 0; k < 2; k++) {
    if (e->ev_prefix_done[k]) cudaEventDestroy(e->ev_prefix_done[k]);
    if (e->ev_prefix_free[k]) cudaEventDestroy(e->ev_prefix_free[k]);
  }
  if (e->ev_upload) cudaEventDestroy(e->ev_upload);
  if (e->tex_stream) { cudaStreamSynchronize(e->tex_stream); cudaStreamDestroy(e->tex_stream); }
  if (e->fin_stream) { cudaStreamSynchronize(e->fin_stream); cudaStreamDestroy(e->fin_stream); }
  if (e->drain_stream) { cudaStreamSynchronize(e->drain_stream); cudaStreamDestroy(e->drain_stream); }
  for (int k = 0; k < LFB_SPARSE_SLOTS; k++) {
    cudaFree(e->slot[k].accum); cudaFree(e->slot[k].state); cudaFree(e->slot[k].stage);
    if (e->slot[k].ev_traced) cudaEventDestroy(e->slot[k].ev_traced);
    if (e->slot[k].ev_done) cudaEventDestroy(e->slot[k].ev_done);
    if (e->slot[k].ev_begin) cudaEventDestroy(e->slot[k].ev_begin);
    if (e->slot[k].ev_tile0) cudaEventDestroy(e->slot[k].ev_tile0);
    if (e->slot[k].ev_staged) cudaEventDestroy(e->slot[k].ev_staged);
  }
  if (e->ev_epoch) cudaEventDestroy(e->ev_epoch);
  if (e->ev_tex_up) cudaEventDestroy(e->ev_tex_up);
  for (int k = 0; k < 2; k++)
    if (e->ev_tex_free[k]) cudaEventDestroy(e->ev_tex_free[k]);
  for (int k = 0; k < 2; k++) {
    lfb_engine::HostStage& st = e->stage[k];
    if (st.arena) cudaFreeHost(st.arena);
    if (st.ev) cudaEventDestroy(st.ev);
    if (e->ev_rec_free[k]) cudaEventDestroy(e->ev_rec_free[k]);
  }
  if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
  for (int k = 0; k < 2; k++) {
    cudaFree(e->d_out_ab[k]);
    if (e->ev_ready[k]) cudaEventDestroy(e->ev_ready[k]);
    if (e->ev_copied[k]) cudaEventDestroy(e->ev_copied[k]);
  }
  cudaFree(e->d_star_tex); cudaFree(e->d_star_scratch); cudaFree(e->d_star_lights); cudaFree(e->d_hdr); cudaFree(e->d_rgba);
  if (e->h_bbox) cudaFreeHost(e->h_bbox);
  if (e->h_count) cudaFreeHost(e->h_count);
  scene_free(e->scene);
  cudaFree(e->d_sparse_state);
  cudaFree(e->d_tex_ab[0]); cudaFree(e->d_tex_ab[1]); cudaFree(e->d_dump_job); cudaFree(e->d_accum); cudaFree(e->d_out);
  cudaFree(e->d_hits); cudaFree(e->d_tris); cudaFree(e->d_ghosts); cudaFree(e->d_pairs); cudaFree(e->d_rgbw);
  if (e->ev_frame0) cudaEventDestroy(e->ev_frame0);
  if (e->ev_trace0) cudaEventDestroy(e->ev_trace0);
  if (e->ev_trace1) cudaEventDestroy(e->ev_trace1);
  if (e->ev_frame1) cudaEventDestroy(e->ev_frame1);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

// ---------------------------------------------------------------------------
// inputs
// ---------------------------------------------------------------------------
namespace {
// CIE 1931 colour matching functions (multi-lobe Gaussian fit of Wyman, Sloan & Shirley 2013)
// -> linear sRGB; used to weight spectral samples into R,G,B when n_lambda != 3.
double lobe(double lam, double mu, double s1, double s2) {
  const double t = (lam - mu) / (lam < mu ? s1 : s2);
  return exp(-0.5 * t * t);
}
void lambda_to_rgb(double lam, double rgb[3]) {
  const double X = 1.056 * lobe(lam, 599.8, 37.9, 31.0) + 0.362 * lobe(lam, 442.0, 16.0, 26.7) - 0.065 * lobe(lam, 501.1, 20.4, 26.2);
  const double Y = 0.821 * lobe(lam, 568.8, 46.9, 40.5) + 0.286 * lobe(lam, 530.9, 16.3, 31.1);
  const double Z = 1.217 * lobe(lam, 437.0, 11.8, 36.0) + 0.681 * lobe(lam, 459.0, 26.0, 13.8);
  rgb[0] = 3.2406 * X - 1.5372 * Y - 0.4986 * Z;
  rgb[1] = -0.9689 * X + 1.8758 * Y + 0.0415 * Z;
  rgb[2] = 0.0557 * X - 0.2040 * Y + 1.0570 * Z;
  for (int c = 0; c < 3; c++) rgb[c] = rgb[c] < 0 ? 0 : rgb[c];
}
}  // namespace

extern "C" int lfb_builtin_lens(lfb_lens* L, int n_lambda, float coating_lambda0_nm) {
  if (!L) return fail(LFB_ERR_INVALID, "lens is NULL");
  if (n_lambda < 1 || n_lambda > LFB_MAX_LAMBDA) return fail(LFB_ERR_INVALID, "n_lambda out of range");
  // The reference's hard-coded prescription, pathtracer.cpp:539-556 (thickness after surface k,
  // radius of curvature, index after surface k for its R, G, B tables).
  static const double thick[9] = {7.700, 1.850, 3.520, 1.850, 4.180, 3.000, 1.850, 7.270, 83.91};
  static const double radius[9] = {30.810, -89.350, 580.380, -80.630, 28.340, 0, 0, 32.190, -52.990};
  static const float n_rgb[3][9] = {{1.652f, 1.5991f, 1, 1.6396f, 1, 1, 1.5776f, 1.68990f, 1},
                                    {1.652f, 1.6113f, 1, 1.65f, 1, 1, 1.5885f, 1.6999f, 1},
                                    {1.652f, 1.6164f, 1, 1.6542f, 1, 1, 1.5930f, 1.7040f, 1}};
  static const double anchor_nm[3] = {650.0, 550.0, 450.0};
  memset(L, 0, sizeof(*L));
  L->n_surfaces = 9; L->stop_index = 5; L->n_lambda = n_lambda;
  L->entrance_half_height = 14.5;  // :737
  L->stop_half_height = 11.6;      // :621
  L->stop_half_height_neg = 11.5;  // :624
  for (int k = 0; k < 9; k++) {
    L->thickness[k] = (float)thick[k];
    L->curvature[k] = radius[k] == 0 ? 0.f : (float)(1 / radius[k]);
    L->semi_aperture[k] = 14.5f;
  }
  if (n_lambda == 3) {
    for (int l = 0; l < 3; l++) {
      L->lambda_nm[l] = (float)anchor_nm[l];
      for (int k = 0; k < 9; k++) L->ior[l][k] = n_rgb[l][k];
      L->rgb_weight[l][l] = 1.f;
    }
  } else {
    double sum[3] = {0, 0, 0}, w[LFB_MAX_LAMBDA][3];
    for (int l = 0; l < n_lambda; l++) {
      const double lam = 400.0 + 300.0 * (l + 0.5) / n_lambda;
      L->lambda_nm[l] = (float)lam;
      lambda_to_rgb(lam, w[l]);
      for (int c = 0; c < 3; c++) sum[c] += w[l][c];
    }
    for (int l = 0; l < n_lambda; l++)
      for (int c = 0; c < 3; c++) L->rgb_weight[l][c] = (float)(w[l][c] / sum[c]);
    // n(lambda): piecewise two-term Cauchy n = A + B/lam^2 through the R,G,B anchors (linear in
    // u = 1/lam^2 on [650,550] and [550,450] nm, each segment continued beyond its anchor): exact
    // at the reference's three indices and monotone, which a single 3-term fit of them is not.
    for (int k = 0; k < 9; k++) {
      double u[3], n[3];
      for (int a = 0; a < 3; a++) {
        const double lm = anchor_nm[a] * 1e-3;
        u[a] = 1.0 / (lm * lm);
        n[a] = n_rgb[a][k];
      }
      const double f01 = (n[1] - n[0]) / (u[1] - u[0]);
      const double f12 = (n[2] - n[1]) / (u[2] - u[1]);
      for (int l = 0; l < n_lambda; l++) {
        const double lm = (double)L->lambda_nm[l] * 1e-3;
        const double x = 1.0 / (lm * lm);
        L->ior[l][k] = (float)(x <= u[1] ? n[1] + (x - u[1]) * f01 : n[1] + (x - u[1]) * f12);
      }
    }
  }
  for (int k = 0; k < 9; k++) {
    bool interface = false;
    for (int l = 0; l < n_lambda; l++) {
      const float before = k == 0 ? 1.f : L->ior[l][k - 1];
      if (before != L->ior[l][k]) interface = true;
    }
    L->coating_lambda0_nm[k] = interface ? coating_lambda0_nm : 0.f;
  }
  return LFB_OK;
}

extern "C" int lfb_set_lens(lfb_engine* e, const lfb_lens* L) {
  int rc = bind(e);
  if (rc) return rc;
  rc = check_lens(L);
  if (rc) return rc;
  CU(cudaStreamSynchronize(e->stream));
  e->has_lens = false;  // until every table below is rebuilt
  e->lens = *L;
  DevLens& D = e->dev_lens;
  memset(&D, 0, sizeof(D));
  D.n_surfaces = L->n_surfaces; D.stop = L->stop_index; D.n_lambda = L->n_lambda;
  double z = 0;
  for (int k = 0; k < L->n_surfaces; k++) {
    D.c[k] = L->curvature[k]; D.d[k] = L->thickness[k];
    D.semi[k] = L->semi_aperture[k]; D.coat[k] = L->coating_lambda0_nm[k];
    D.zv_d[k] = z; D.zv[k] = (float)z;
    z += (double)L->thickness[k];
  }
  D.zv_d[L->n_surfaces] = z; D.zv[L->n_surfaces] = (float)z;
  memcpy(D.ior, L->ior, sizeof(D.ior));
  memcpy(D.lambda_nm, L->lambda_nm, sizeof(D.lambda_nm));
  D.P = L->entrance_half_height; D.h_stop = L->stop_half_height; D.h_stop_neg = L->stop_half_height_neg;
  {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (e->device < 64 && g_const_owner[e->device] == e) g_const_owner[e->device] = nullptr;
  }
  // REF_QUADS tables: the reference's pair list and per-wavelength colour basis
  int pairs[LFB_MAX_SURFACES * LFB_MAX_SURFACES][2];
  const int np = list_pairs(*L, LFB_PAIRS_REF, pairs);
  const int ng = np * L->n_lambda;
  cudaFree(e->d_pairs); cudaFree(e->d_rgbw); cudaFree(e->d_tris); cudaFree(e->d_ghosts);
  e->d_pairs = nullptr; e->d_rgbw = nullptr; e->d_tris = nullptr; e->d_ghosts = nullptr;
  CU(cudaMalloc((void**)&e->d_pairs, sizeof(int) * 2 * (size_t)std::max(np, 1)));
  CU(cudaMalloc((void**)&e->d_rgbw, sizeof(float) * 3 * (size_t)L->n_lambda));
  CU(cudaMalloc((void**)&e->d_tris, sizeof(RefTri) * 2 * (size_t)std::max(ng, 1)));
  CU(cudaMalloc((void**)&e->d_ghosts, sizeof(lfb_ref_ghost) * (size_t)std::max(ng, 1)));
  if (np > 0) CU(cudaMemcpy(e->d_pairs, pairs, sizeof(int) * 2 * (size_t)np, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->d_rgbw, L->rgb_weight, sizeof(float) * 3 * (size_t)L->n_lambda, cudaMemcpyHostToDevice));
  e->n_ref_pairs = np; e->n_ref_ghosts = 0;
  {
    std::vector<float2> lut;
    build_reflectance_tables(*L, lut, e->poly);
    cudaFree(e->d_lut);
    e->d_lut = nullptr;
    CU(cudaMalloc((void**)&e->d_lut, sizeof(float2) * lut.size()));
    CU(cudaMemcpy(e->d_lut, lut.data(), sizeof(float2) * lut.size(), cudaMemcpyHostToDevice));
  }
  e->job_key.clear();
  e->struct_key.clear();
  e->has_lens = true;
  return LFB_OK;
}

extern "C" int lfb_set_aperture(lfb_engine* e, const float* texels, int w, int h) {
  int rc = bind(e);
  if (rc) return rc;
  if (!texels || w < 1 || h < 1 || w > 16384 || h > 16384) return fail(LFB_ERR_INVALID, "bad aperture texture");
  const size_t bytes = sizeof(float) * (size_t)w * h;
  if (w != e->tex_w || h != e->tex_h) {
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaStreamSynchronize(e->tex_stream));
    for (int k = 0; k < 2; k++) {
      cudaFree(e->d_tex_ab[k]);
      e->d_tex_ab[k] = nullptr;
      e->tex_free_valid[k] = false;
    }
    e->d_tex = nullptr; e->has_tex = false;
    for (int k = 0; k < 2; k++) CU(cudaMalloc((void**)&e->d_tex_ab[k], bytes));
    e->tex_w = w; e->tex_h = h;
  }
  const int nb = e->tex_cur ^ 1;
  if (e->tex_free_valid[nb]) CU(cudaStreamWaitEvent(e->tex_stream, e->ev_tex_free[nb], 0));
  CU(cudaMemcpyAsync(e->d_tex_ab[nb], texels, bytes, cudaMemcpyHostToDevice, e->tex_stream));
  CU(cudaEventRecord(e->ev_tex_up, e->tex_stream));
  CU(cudaEventSynchronize(e->ev_tex_up));  // the caller may free texels on return; later launches need no device-side wait
  // everything enqueued so far reads the current texture (sweeps on prefix_stream are joined into `stream` by their frame)
  CU(cudaEventRecord(e->ev_tex_free[e->tex_cur], e->stream));
  e->tex_free_valid[e->tex_cur] = true;
  e->tex_cur = nb;
  e->d_tex = e->d_tex_ab[nb];
  e->has_tex = true;
  return LFB_OK;
}

// ---------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------
extern "C" size_t lfb_accum_bytes(int width, int height) {
  if (width < 1 || height < 1) return 0;
  return accum_layout(width, height).total;
}

extern "C" void* lfb_stream(lfb_engine* e) { return e ? (void*)e->stream : nullptr; }

extern "C" int lfb_sync(lfb_engine* e) {
  int rc = bind(e);
  if (rc) return rc;
  CU(cudaStreamSynchronize(e->stream));
  if (e->copy_stream) CU(cudaStreamSynchronize(e->copy_stream));
  if (e->fin_stream) CU(cudaStreamSynchronize(e->fin_stream));
  if (e->drain_stream) CU(cudaStreamSynchronize(e->drain_stream));
  return LFB_OK;
}

extern "C" int lfb_render_ghosts_device(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params* P,
                                        void* accum_dev, int clear_first) {
  int rc = bind(e);
  if (rc) return rc;
  if (!e->has_lens || !e->has_tex) return fail(LFB_ERR_STATE, "set the lens and the aperture first");
  rc = check_params(P, true);
  if (rc) return rc;
  if (n_lights < 0 || (n_lights > 0 && !lights) || !accum_dev) return fail(LFB_ERR_INVALID, "bad lights/accumulator");
  return render_grid_device(e, lights, n_lights, *P, (unsigned long long*)accum_dev, clear_first);
}

extern "C" int lfb_finalize_device(lfb_engine* e, const void* accum_dev, const lfb_params* P, void* out_dev,
                                   size_t out_stride_bytes, int out_elem) {
  int rc = bind(e);
  if (rc) return rc;
  rc = check_params(P, true);
  if (rc) return rc;
  if (!accum_dev || !out_dev) return fail(LFB_ERR_INVALID, "NULL device pointer");
  if (out_elem != LFB_F32x3 && out_elem != LFB_F64x3) return fail(LFB_ERR_INVALID, "unknown out_elem");
  if (out_stride_bytes < elem_bytes(out_elem) || out_stride_bytes % (out_elem == LFB_F32x3 ? 4 : 8)) return fail(LFB_ERR_INVALID, "bad out_stride_bytes");
  const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
  CU(launch_finalize((const unsigned long long*)accum_dev, P->width, P->height, inv, out_dev, out_stride_bytes, out_elem, 0, e->stream));
  e->launches++;
  return LFB_OK;
}

extern "C" int lfb_finalize_clear_device(lfb_engine* e, void* accum_dev, const lfb_params* P, void* out_dev,
                                         size_t out_stride_bytes, int out_elem) {
  int rc = bind(e);
  if (rc) return rc;
  rc = check_params(P, true);
  if (rc) return rc;
  if (!accum_dev || !out_dev) return fail(LFB_ERR_INVALID, "NULL device pointer");
  if (out_elem != LFB_F32x3 && out_elem != LFB_F64x3) return fail(LFB_ERR_INVALID, "unknown out_elem");
  if (out_stride_bytes < elem_bytes(out_elem) || out_stride_bytes % (out_elem == LFB_F32x3 ? 4 : 8)) return fail(LFB_ERR_INVALID, "bad out_stride_bytes");
  const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
  CU(launch_finalize_clear((unsigned long long*)accum_dev, P->width, P->height, inv, out_dev, out_stride_bytes, out_elem, e->stream));
  e->launches++;
  return LFB_OK;
}

extern "C" int lfb_peer_barrier(lfb_engine* e, void* const* flag_ptrs, int n_ranks, int rank, uint64_t epoch) {
  int rc = bind(e);
  if (rc) return rc;
  if (!flag_ptrs || n_ranks < 1 || n_ranks > LFB_MAX_PEERS || rank < 0 || rank >= n_ranks) return fail(LFB_ERR_INVALID, "bad peer set");
  PeerFlags F;
  F.n = n_ranks; F.rank_self = rank;
  for (int r = 0; r < n_ranks; r++) {
    if (!flag_ptrs[r]) return fail(LFB_ERR_INVALID, "NULL flag array");
    F.ptr[r] = (unsigned long long*)flag_ptrs[r];
  }
  CU(launch_peer_barrier(F, rank, (unsigned long long)epoch, e->stream));
  e->launches++;
  return LFB_OK;
}

extern "C" int lfb_reduce_finalize_peers(lfb_engine* e, const void* const* accum_ptrs, int n_ranks, int rank, const void* multicast_accum,
                                         const lfb_params* P, void* out_dev, size_t out_stride_bytes, int out_elem) {
  int rc = bind(e);
  if (rc) return rc;
  rc = check_params(P, true);
  if (rc) return rc;
  if (!accum_ptrs || n_ranks < 1 || n_ranks > LFB_MAX_PEERS || rank < 0 || rank >= n_ranks || !out_dev) return fail(LFB_ERR_INVALID, "bad peer set");
  if (out_elem != LFB_F32x3 && out_elem != LFB_F64x3) return fail(LFB_ERR_INVALID, "unknown out_elem");
  if (out_stride_bytes < elem_bytes(out_elem) || out_stride_bytes % (out_elem == LFB_F32x3 ? 4 : 8)) return fail(LFB_ERR_INVALID, "bad out_stride_bytes");
  PeerAccums A;
  A.n = n_ranks;
  for (int r = 0; r < n_ranks; r++) {
    if (!accum_ptrs[r]) return fail(LFB_ERR_INVALID, "NULL peer accumulator");
    A.ptr[r] = (const unsigned long long*)accum_ptrs[r];
  }
  const size_t npx = (size_t)P->width * P->height;
  const size_t p0 = npx * (size_t)rank / (size_t)n_ranks, p1 = npx * (size_t)(rank + 1) / (size_t)n_ranks;
  const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
  CU(launch_reduce_finalize(A, (const unsigned long long*)multicast_accum, p0, p1, inv, out_dev, out_stride_bytes, out_elem, e->opt.reduce_ctas, e->stream));
  e->launches++;
  return LFB_OK;
}

extern "C" int lfb_render_ghosts(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params* P, void* out,
                                 size_t stride, int elem, int additive) {
  int rc = bind(e);
  if (rc) return rc;
  if (!e->has_lens || !e->has_tex) return fail(LFB_ERR_STATE, "set the lens and the aperture first");
  rc = check_params(P, false);
  if (rc) return rc;
  if (n_lights < 0 || (n_lights > 0 && !lights) || !out) return fail(LFB_ERR_INVALID, "bad lights/out");
  if (elem != LFB_F32x3 && elem != LFB_F64x3) return fail(LFB_ERR_INVALID, "unknown out_elem");
  if (stride < elem_bytes(elem) || stride % (elem == LFB_F32x3 ? 4 : 8)) return fail(LFB_ERR_INVALID, "bad out_stride_bytes");
  const size_t npx = (size_t)P->width * P->height;
  const size_t out_bytes = (npx - 1) * stride + elem_bytes(elem);
  rc = grow(&e->d_out, &e->out_cap, out_bytes);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_frame0, e->stream));
  if (additive) CU(cudaMemcpyAsync(e->d_out, out, out_bytes, cudaMemcpyHostToDevice, e->stream));
  else if (stride != elem_bytes(elem)) CU(cudaMemsetAsync(e->d_out, 0, out_bytes, e->stream));  // padding bytes come back as 0
  if (P->mode == LFB_MODE_REF_QUADS) {
    rc = render_ref_device(e, lights, n_lights, *P, e->d_out, stride, elem, additive);
    if (rc) return rc;
  } else {
    rc = grow(&e->d_accum, &e->accum_cap, lfb_accum_bytes(P->width, P->height));
    if (rc) return rc;
    e->accum_clean = false;
    e->accum_dirty_w = e->accum_dirty_h = 0;  // the whole buffer is about to be used: the rect path must start fresh
    rc = render_grid_device(e, lights, n_lights, *P, e->d_accum, 1);
    if (rc) return rc;
    const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
    CU(launch_finalize(e->d_accum, P->width, P->height, inv, e->d_out, stride, elem, additive, e->stream));
    e->launches++;
  }
  CU(cudaMemcpyAsync(out, e->d_out, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaEventRecord(e->ev_frame1, e->stream));
  CU(cudaStreamSynchronize(e->stream));  // the reference's caller reads ghost_buffer right after the call
  CU(cudaEventElapsedTime(&e->last_frame_ms, e->ev_frame0, e->ev_frame1));
  CU(cudaEventElapsedTime(&e->last_trace_ms, e->ev_trace0, e->ev_trace1));
  return LFB_OK;
}

// ---------------------------------------------------------------------------
// tile-sparse back end (sparse.cu)
// ---------------------------------------------------------------------------
extern "C" size_t lfb_tile_state_bytes(int width, int height) {
  if (width < 1 || height < 1) return 0;
  return tile_state_bytes(width, height);
}

namespace {
int check_out_args(const void* out_dev, size_t stride, int elem) {
  if (!out_dev) return fail(LFB_ERR_INVALID, "NULL output pointer");
  if (elem != LFB_F32x3 && elem != LFB_F64x3) return fail(LFB_ERR_INVALID, "unknown out_elem");
  if (stride < elem_bytes(elem) || stride % (elem == LFB_F32x3 ? 4 : 8)) return fail(LFB_ERR_INVALID, "bad out_stride_bytes");
  return LFB_OK;
}
}  // namespace

extern "C" int lfb_finalize_tiles_device(lfb_engine* e, void* accum_dev, const lfb_params* P, void* out_dev, size_t stride, int elem,
                                         void* tile_state_dev) {
  int rc = bind(e);
  if (rc) return rc;
  rc = check_params(P, true);
  if (rc) return rc;
  if (!accum_dev || !tile_state_dev) return fail(LFB_ERR_INVALID, "NULL device pointer");
  rc = check_out_args(out_dev, stride, elem);
  if (rc) return rc;
  PeerAccums A;
  memset(&A, 0, sizeof(A));
  A.n = 1; A.ptr[0] = (const unsigned long long*)accum_dev;
  const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
  CU(launch_tiles(A, 0, P->width, P->height, inv, out_dev, stride, elem, (unsigned*)tile_state_dev, nullptr, 0, e->stream));
  e->launches++;
  return LFB_OK;
}

extern "C" int lfb_reduce_tiles_peers(lfb_engine* e, void* const* accum_ptrs, int n_ranks, int rank, const lfb_params* P, void* out_dev,
                                      size_t stride, int elem, void* tile_state_dev) {
  int rc = bind(e);
  if (rc) return rc;
  rc = check_params(P, true);
  if (rc) return rc;
  if (!accum_ptrs || n_ranks < 1 || n_ranks > LFB_MAX_PEERS || rank < 0 || rank >= n_ranks || !tile_state_dev) return fail(LFB_ERR_INVALID, "bad peer set");
  rc = check_out_args(out_dev, stride, elem);
  if (rc) return rc;
  PeerAccums A;
  memset(&A, 0, sizeof(A));
  A.n = n_ranks;
  for (int r = 0; r < n_ranks; r++) {
    if (!accum_ptrs[r]) return fail(LFB_ERR_INVALID, "NULL peer accumulator");
    A.ptr[r] = (const unsigned long long*)accum_ptrs[r];
  }
  const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
  CU(launch_tiles(A, rank, P->width, P->height, inv, out_dev, stride, elem, (unsigned*)tile_state_dev, nullptr, e->opt.reduce_ctas, e->stream));
  e->launches++;
  return LFB_OK;
}

// The two-stage form for an output frame in HOST memory written while other frames are in flight (include/lfb200.h):
// lfb_reduce_tiles_peers_staged leaves this rank's share of the dirty tiles as pixels in stage_dev, lfb_drain_tiles copies them
// into the host frame with a few clock-paced CTAs.
extern "C" size_t lfb_tile_stage_bytes(int width, int height, size_t stride) {
  if (width < 1 || height < 1 || stride < 12) return 0;
  return tile_stage_bytes(width, height, stride);
}

extern "C" int lfb_reduce_tiles_peers_staged(lfb_engine* e, void* const* accum_ptrs, int n_ranks, int rank, const lfb_params* P, void* out_dev,
                                             size_t stride, int elem, void* tile_state_dev, void* stage_dev) {
  int rc = bind(e);
  if (rc) return rc;
  rc = check_params(P, true);
  if (rc) return rc;
  if (!accum_ptrs || n_ranks < 1 || n_ranks > LFB_MAX_PEERS || rank < 0 || rank >= n_ranks || !tile_state_dev || !stage_dev) return fail(LFB_ERR_INVALID, "bad peer set");
  rc = check_out_args(out_dev, stride, elem);
  if (rc) return rc;
  PeerAccums A;
  memset(&A, 0, sizeof(A));
  A.n = n_ranks;
  for (int r = 0; r < n_ranks; r++) {
    if (!accum_ptrs[r]) return fail(LFB_ERR_INVALID, "NULL peer accumulator");
    A.ptr[r] = (const unsigned long long*)accum_ptrs[r];
  }
  const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
  CU(launch_tiles(A, rank, P->width, P->height, inv, out_dev, stride, elem, (unsigned*)tile_state_dev, nullptr, e->opt.reduce_ctas, e->stream, stage_dev));
  e->launches++;
  return LFB_OK;
}

extern "C" int lfb_drain_tiles(lfb_engine* e, const lfb_params* P, void* out_dev, size_t stride, const void* tile_state_dev, const void* stage_dev) {
  int rc = bind(e);
  if (rc) return rc;
  rc = check_params(P, true);
  if (rc) return rc;
  if (!out_dev || !tile_state_dev || !stage_dev || stride < 12) return fail(LFB_ERR_INVALID, "NULL pointer");
  if (e->drain_gbps < 0.f) {  // once per engine (as in lfb_render_ghosts_sparse_begin)
    if (e->opt.host_write_mbps > 0) e->drain_gbps = 1e-3f * (float)e->opt.host_write_mbps;
    else if (e->opt.host_write_mbps < 0) e->drain_gbps = 0.f;
    else {
      float link = 0.f;
      CU(measure_host_write_gbps(e->stream, &link));
      e->drain_gbps = link > 5.f ? 0.93f * link : 0.f;
    }
  }
  CU(launch_tile_drain(P->width, P->height, out_dev, stride, (const unsigned*)tile_state_dev, stage_dev, e->opt.reduce_ctas > 0 && e->opt.reduce_ctas <= 64 ? e->opt.reduce_ctas : 0, e->drain_gbps, e->stream));
  e->launches++;
  return LFB_OK;
}

extern "C" int lfb_render_ghosts_sparse(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params* P, void* out, size_t stride,
                                        int elem, int out_is_clear, int* tiles_written) {
  int rc = bind(e);
  if (rc) return rc;
  if (!e->has_lens || !e->has_tex) return fail(LFB_ERR_STATE, "set the lens and the aperture first");
  rc = check_params(P, true);
  if (rc) return rc;
  if (n_lights < 0 || (n_lights > 0 && !lights)) return fail(LFB_ERR_INVALID, "bad lights");
  rc = check_out_args(out, stride, elem);
  if (rc) return rc;
  if (tiles_written) *tiles_written = 0;
  // is the caller's buffer page-locked and mapped into this device?  (else: the full-frame copy)
  cudaPointerAttributes attr;
  memset(&attr, 0, sizeof(attr));
  const cudaError_t perr = cudaPointerGetAttributes(&attr, out);
  if (perr != cudaSuccess) cudaGetLastError();
  const bool same = out == e->sparse_out && P->width == e->sparse_w && P->height == e->sparse_h && stride == e->sparse_stride && elem == e->sparse_elem;
  if (!out_is_clear && !same) return fail(LFB_ERR_STATE, "lfb_render_ghosts_sparse: `out` is not the buffer of the previous sparse call; pass out_is_clear = 1 with a zeroed buffer");
  if (perr != cudaSuccess || attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
    e->sparse_out = nullptr;  // the dense call rewrites every pixel: nothing to remember
    if (tiles_written) *tiles_written = -1;
    return lfb_render_ghosts(e, lights, n_lights, P, out, stride, elem, 0);
  }
  const AccumLayout lay = accum_layout(P->width, P->height);
  rc = grow(&e->d_sparse_state, &e->sparse_state_cap, tile_state_bytes(P->width, P->height));
  if (rc) return rc;
  if (out_is_clear || !same) CU(cudaMemsetAsync(e->d_sparse_state, 0, tile_state_bytes(P->width, P->height), e->stream));
  e->sparse_out = out; e->sparse_w = P->width; e->sparse_h = P->height; e->sparse_stride = stride; e->sparse_elem = elem;
  if (lay.total > e->accum_cap) e->accum_clean = false;
  rc = grow(&e->d_accum, &e->accum_cap, lay.total);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_frame0, e->stream));
  const bool clean = e->accum_clean && e->accum_clean_w == P->width && e->accum_clean_h == P->height;
  e->accum_clean = false;
  e->accum_dirty_w = e->accum_dirty_h = 0;
  rc = render_grid_device(e, lights, n_lights, *P, e->d_accum, clean ? 0 : 1);
  if (rc) return rc;
  PeerAccums A;
  memset(&A, 0, sizeof(A));
  A.n = 1; A.ptr[0] = e->d_accum;
  const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
  unsigned* count_dev = nullptr;
  CU(cudaHostGetDevicePointer((void**)&count_dev, e->h_count, 0));
  CU(launch_tiles(A, 0, P->width, P->height, inv, attr.devicePointer, stride, elem, e->d_sparse_state, count_dev, 0, e->stream));
  e->launches++;
  CU(cudaEventRecord(e->ev_frame1, e->stream));
  CU(cudaStreamSynchronize(e->stream));  // the kernel's stores into the caller's memory are complete and visible
  e->accum_clean = true; e->accum_clean_w = P->width; e->accum_clean_h = P->height;
  if (tiles_written) *tiles_written = (int)e->h_count[0];
  CU(cudaEventElapsedTime(&e->last_frame_ms, e->ev_frame0, e->ev_frame1));
  CU(cudaEventElapsedTime(&e->last_trace_ms, e->ev_trace0, e->ev_trace1));
  return LFB_OK;
}

// Several frames in flight (include/lfb200.h).  begin(slot): trace on the engine stream into the slot's own accumulator, then the
// tile kernel on fin_stream writes the dirty tiles into `out` (page-locked, mapped); end(slot): wait for exactly that.
extern "C" int lfb_render_ghosts_sparse_begin(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params* P, void* out,
                                              size_t stride, int elem, int out_is_clear, int slot) {
  int rc = bind(e);
  if (rc) return rc;
  if (slot < 0 || slot >= LFB_SPARSE_SLOTS) return fail(LFB_ERR_INVALID, "slot must be in [0, LFB_SPARSE_SLOTS)");
  if (!e->has_lens || !e->has_tex) return fail(LFB_ERR_STATE, "set the lens and the aperture first");
  rc = check_params(P, true);
  if (rc) return rc;
  if (n_lights < 0 || (n_lights > 0 && !lights)) return fail(LFB_ERR_INVALID, "bad lights");
  rc = check_out_args(out, stride, elem);
  if (rc) return rc;
  lfb_engine::SparseSlot& S = e->slot[slot];
  if (S.pending) return fail(LFB_ERR_STATE, "lfb_render_ghosts_sparse_begin: the slot's previous frame was not collected (lfb_render_ghosts_sparse_end)");
  cudaPointerAttributes attr;
  memset(&attr, 0, sizeof(attr));
  if (cudaPointerGetAttributes(&attr, out) != cudaSuccess || attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
    cudaGetLastError();
    return fail(LFB_ERR_INVALID, "lfb_render_ghosts_sparse_begin: `out` must be page-locked host memory mapped into the device (lfb_host_alloc / lfb_host_register)");
  }
  const int W = P->width, H = P->height;
  const bool same = out == S.out && W == S.w && H == S.h && stride == S.stride && elem == S.elem;
  if (!out_is_clear && !same) return fail(LFB_ERR_STATE, "lfb_render_ghosts_sparse_begin: `out` is not the buffer of the slot's previous frame; pass out_is_clear = 1 with a zeroed buffer");
  const AccumLayout lay = accum_layout(W, H);
  const size_t st_bytes = tile_state_bytes(W, H);
  const size_t stage_bytes = tile_stage_bytes(W, H, stride);
  if (lay.total > S.accum_cap || st_bytes > S.state_cap || stage_bytes > S.stage_cap) {
    CU(cudaStreamSynchronize(e->fin_stream));
    CU(cudaStreamSynchronize(e->drain_stream));
    if (lay.total > S.accum_cap) S.accum_clean = false;
    if ((rc = grow(&S.accum, &S.accum_cap, lay.total))) return rc;
    if ((rc = grow(&S.state, &S.state_cap, st_bytes))) return rc;
    if ((rc = grow(&S.stage, &S.stage_cap, stage_bytes))) return rc;
  }
  if (e->drain_gbps < 0.f) {  // once per engine
    if (e->opt.host_write_mbps > 0) e->drain_gbps = 1e-3f * (float)e->opt.host_write_mbps;
    else if (e->opt.host_write_mbps < 0) e->drain_gbps = 0.f;  // unpaced
    else {
      float link = 0.f;
      CU(measure_host_write_gbps(e->fin_stream, &link));
      // measured (tools/e2e_probe.py, link ~50 GB/s): 0.146 ms / frame at 48 GB/s, 0.155 at 44, 0.19 at 32, 0.17 at 52, 0.19 unpaced
      e->drain_gbps = link > 5.f ? 0.93f * link : 0.f;
    }
  }
  // the slot's previous tile kernel (which also re-zeroed its accumulator) precedes this frame's trace
  if (S.done_valid) CU(cudaStreamWaitEvent(e->stream, S.ev_done, 0));
  CU(cudaEventRecord(S.ev_begin, e->stream));
  if (out_is_clear || !same) CU(cudaMemsetAsync(S.state, 0, st_bytes, e->stream));
  S.out = out; S.w = W; S.h = H; S.stride = stride; S.elem = elem;
  const bool clean = S.accum_clean && S.accum_w == W && S.accum_h == H;
  S.accum_clean = false;
  rc = render_grid_device(e, lights, n_lights, *P, S.accum, clean ? 0 : 1);
  if (rc) return rc;
  CU(cudaEventRecord(S.ev_traced, e->stream));
  CU(cudaStreamWaitEvent(e->fin_stream, S.ev_traced, 0));
  PeerAccums A;
  memset(&A, 0, sizeof(A));
  A.n = 1; A.ptr[0] = S.accum;
  const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
  unsigned* count_dev = nullptr;
  CU(cudaHostGetDevicePointer((void**)&count_dev, e->h_count + 1 + slot, 0));
  CU(cudaEventRecord(S.ev_tile0, e->fin_stream));
  // stage 1 (every SM, ~15 us): dirty tiles -> pixels in the slot's device staging; stage 2 (a few CTAs on their own stream,
  // paced): staging -> the caller's host frame.  Stage 1 of the next frame runs while this frame drains.
  CU(launch_tiles(A, 0, W, H, inv, attr.devicePointer, stride, elem, S.state, count_dev, 0, e->fin_stream, S.stage));
  CU(cudaEventRecord(S.ev_staged, e->fin_stream));
  CU(cudaStreamWaitEvent(e->drain_stream, S.ev_staged, 0));
  CU(launch_tile_drain(W, H, attr.devicePointer, stride, S.state, S.stage, e->opt.reduce_ctas, e->drain_gbps, e->drain_stream));
  e->launches += 2;
  CU(cudaEventRecord(S.ev_done, e->drain_stream));
  S.done_valid = true;
  S.pending = true;
  S.accum_clean = true; S.accum_w = W; S.accum_h = H;  // once ev_done has fired, which every later user of the slot waits for
  return LFB_OK;
}

extern "C" int lfb_sparse_slot_times(lfb_engine* e, int slot, float ms[5]) {
  int rc = bind(e);
  if (rc) return rc;
  if (slot < 0 || slot >= LFB_SPARSE_SLOTS || !ms) return fail(LFB_ERR_INVALID, "bad slot");
  lfb_engine::SparseSlot& S = e->slot[slot];
  if (!S.done_valid || S.pending) return fail(LFB_ERR_STATE, "lfb_sparse_slot_times: the slot has no collected frame");
  cudaEvent_t evs[5] = {S.ev_begin, S.ev_traced, S.ev_tile0, S.ev_staged, S.ev_done};
  for (int k = 0; k < 5; k++) CU(cudaEventElapsedTime(&ms[k], e->ev_epoch, evs[k]));
  return LFB_OK;
}

extern "C" int lfb_render_ghosts_sparse_end(lfb_engine* e, int slot, int* tiles_written) {
  int rc = bind(e);
  if (rc) return rc;
  if (slot < 0 || slot >= LFB_SPARSE_SLOTS) return fail(LFB_ERR_INVALID, "slot must be in [0, LFB_SPARSE_SLOTS)");
  lfb_engine::SparseSlot& S = e->slot[slot];
  if (!S.pending) return fail(LFB_ERR_STATE, "lfb_render_ghosts_sparse_end: no frame in flight in this slot");
  CU(cudaEventSynchronize(S.ev_done));  // the kernel's stores into the caller's memory are complete and visible
  S.pending = false;
  if (tiles_written) *tiles_written = (int)e->h_count[1 + slot];
  return LFB_OK;
}

// lfb_render_ghosts without the wait: the frame is traced and converted on the engine stream, then copied to `out` on a
// second stream out of one of two device buffers, so the 49.8 MB copy of frame k overlaps the trace of frame k+1 (the PCIe
// copy is ~8x longer than the trace).  `out` is complete after lfb_sync(); alternate between two host buffers to keep two
// frames in flight.  Grid modes only, overwrite semantics.
extern "C" int lfb_render_ghosts_async(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params* P, void* out,
                                       size_t stride, int elem) {
  int rc = bind(e);
  if (rc) return rc;
  if (!e->has_lens || !e->has_tex) return fail(LFB_ERR_STATE, "set the lens and the aperture first");
  rc = check_params(P, true);
  if (rc) return rc;
  if (n_lights < 0 || (n_lights > 0 && !lights) || !out) return fail(LFB_ERR_INVALID, "bad lights/out");
  if (elem != LFB_F32x3 && elem != LFB_F64x3) return fail(LFB_ERR_INVALID, "unknown out_elem");
  if (stride < elem_bytes(elem) || stride % (elem == LFB_F32x3 ? 4 : 8)) return fail(LFB_ERR_INVALID, "bad out_stride_bytes");
  if (!e->copy_stream) {
    CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      CU(cudaEventCreateWithFlags(&e->ev_ready[k], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&e->ev_copied[k], cudaEventDisableTiming));
    }
  }
  const int b = e->ab;
  e->ab ^= 1;
  const size_t npx = (size_t)P->width * P->height;
  const size_t out_bytes = (npx - 1) * stride + elem_bytes(elem);
  if (out_bytes > e->out_ab_cap[b]) {
    CU(cudaStreamSynchronize(e->copy_stream));
    rc = grow(&e->d_out_ab[b], &e->out_ab_cap[b], out_bytes);
    if (rc) return rc;
    e->copied_valid[b] = false;
  }
  rc = grow(&e->d_accum, &e->accum_cap, lfb_accum_bytes(P->width, P->height));
  if (rc) return rc;
  e->accum_clean = false;
  // buffer b's previous copy must have left the device buffer before it is overwritten
  if (e->copied_valid[b]) CU(cudaStreamWaitEvent(e->stream, e->ev_copied[b], 0));
  if (stride != elem_bytes(elem)) CU(cudaMemsetAsync(e->d_out_ab[b], 0, out_bytes, e->stream));
  e->accum_dirty_w = e->accum_dirty_h = 0;
  rc = render_grid_device(e, lights, n_lights, *P, e->d_accum, 1);
  if (rc) return rc;
  const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
  CU(launch_finalize(e->d_accum, P->width, P->height, inv, e->d_out_ab[b], stride, elem, 0, e->stream));
  e->launches++;
  CU(cudaEventRecord(e->ev_ready[b], e->stream));
  CU(cudaStreamWaitEvent(e->copy_stream, e->ev_ready[b], 0));
  CU(cudaMemcpyAsync(out, e->d_out_ab[b], out_bytes, cudaMemcpyDeviceToHost, e->copy_stream));
  CU(cudaEventRecord(e->ev_copied[b], e->copy_stream));
  e->copied_valid[b] = true;
  return LFB_OK;
}

// The dirty-rectangle form of lfb_render_ghosts: a flare covers a small part of the sensor, and the caller's
// buffer is normally already clear (the reference clears ghost_buffer right before, pathtracer.cpp:719-720), so only
// the bounding rectangle of the pixels the frame deposits into is converted and copied.  Pixels outside rect_out are
// NOT touched; inside it every pixel is overwritten (zeros where nothing landed).  rect_out = {x0, y0, x1, y1}
// inclusive, or {0, 0, -1, -1} for an empty frame.
extern "C" int lfb_render_ghosts_rect(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params* P, void* out,
                                      size_t stride, int elem, int* rect_out) {
  int rc = bind(e);
  if (rc) return rc;
  if (!e->has_lens || !e->has_tex) return fail(LFB_ERR_STATE, "set the lens and the aperture first");
  rc = check_params(P, false);
  if (rc) return rc;
  if (n_lights < 0 || (n_lights > 0 && !lights) || !out || !rect_out) return fail(LFB_ERR_INVALID, "bad lights/out/rect_out");
  if (elem != LFB_F32x3 && elem != LFB_F64x3) return fail(LFB_ERR_INVALID, "unknown out_elem");
  if (stride < elem_bytes(elem) || stride % (elem == LFB_F32x3 ? 4 : 8)) return fail(LFB_ERR_INVALID, "bad out_stride_bytes");
  rect_out[0] = rect_out[1] = 0; rect_out[2] = rect_out[3] = -1;
  const size_t npx = (size_t)P->width * P->height;
  rc = grow(&e->d_out, &e->out_cap, npx * stride);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_frame0, e->stream));
  CU(cudaMemcpyAsync(e->d_bbox, e->h_bbox, 4 * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  struct Track { lfb_engine* e; Track(lfb_engine* x) : e(x) { e->track_bbox = true; } ~Track() { e->track_bbox = false; } } track(e);
  int rect[4] = {0, 0, -1, -1};
  if (P->mode == LFB_MODE_REF_QUADS) {
    if (stride != elem_bytes(elem)) CU(cudaMemsetAsync(e->d_out, 0, npx * stride, e->stream));
    rc = render_ref_device(e, lights, n_lights, *P, e->d_out, stride, elem, 0);  // reads the box back itself
    if (rc) return rc;
    memcpy(rect, e->h_bbox + 4, sizeof(rect));
    if (n_lights == 0 || e->n_ref_ghosts == 0) { rect[0] = rect[1] = 0; rect[2] = rect[3] = -1; }
  } else {
    const size_t need = lfb_accum_bytes(P->width, P->height);
    const bool fresh = need > e->accum_cap || e->accum_dirty_w != P->width || e->accum_dirty_h != P->height;
    rc = grow(&e->d_accum, &e->accum_cap, need);
    if (rc) return rc;
    e->accum_clean = false;
    // the accumulators are zero except for the rectangle the previous rect frame left: clear only that
    if (fresh) {
      CU(cudaMemsetAsync(e->d_accum, 0, need, e->stream));
    } else if (e->accum_dirty[2] >= e->accum_dirty[0]) {
      const int* d = e->accum_dirty;
      CU(cudaMemset2DAsync(e->d_accum + 3 * ((size_t)d[0] + (size_t)d[1] * P->width), (size_t)P->width * 24, 0, (size_t)(d[2] - d[0] + 1) * 24,
                           (size_t)(d[3] - d[1] + 1), e->stream));
    }
    e->accum_dirty_w = P->width; e->accum_dirty_h = P->height;
    e->accum_dirty[0] = e->accum_dirty[1] = 0; e->accum_dirty[2] = P->width - 1; e->accum_dirty[3] = P->height - 1;  // until the box is known
    rc = render_grid_device(e, lights, n_lights, *P, e->d_accum, 0);
    if (rc) return rc;
    CU(cudaMemcpyAsync(e->h_bbox + 4, e->d_bbox, 4 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    memcpy(rect, e->h_bbox + 4, sizeof(rect));
    memcpy(e->accum_dirty, rect, sizeof(rect));
    if (rect[2] >= rect[0] && rect[3] >= rect[1]) {
      if (stride != elem_bytes(elem)) CU(cudaMemsetAsync(e->d_out, 0, (size_t)(rect[2] - rect[0] + 1) * (rect[3] - rect[1] + 1) * stride, e->stream));
      const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
      CU(launch_finalize_rect(e->d_accum, P->width, rect, inv, e->d_out, stride, elem, e->stream));
      e->launches++;
    }
  }
  if (rect[2] >= rect[0] && rect[3] >= rect[1]) {
    const size_t rw = (size_t)(rect[2] - rect[0] + 1), rh = (size_t)(rect[3] - rect[1] + 1);
    char* dst = (char*)out + ((size_t)rect[0] + (size_t)rect[1] * P->width) * stride;
    // each row: rw pixels; the last pixel's padding lane (stride > element) is left alone in the last row only
    CU(cudaMemcpy2DAsync(dst, (size_t)P->width * stride, e->d_out, rw * stride, (rw - 1) * stride + elem_bytes(elem), rh,
                         cudaMemcpyDeviceToHost, e->stream));
    memcpy(rect_out, rect, sizeof(rect));
  }
  CU(cudaEventRecord(e->ev_frame1, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaEventElapsedTime(&e->last_frame_ms, e->ev_frame0, e->ev_frame1));
  CU(cudaEventElapsedTime(&e->last_trace_ms, e->ev_trace0, e->ev_trace1));
  return LFB_OK;
}

extern "C" int lfb_dump_rays(lfb_engine* e, const lfb_light* light, const lfb_params* P, int i, int j, int lambda,
                             lfb_ray_hit* out, size_t cap) {
  int rc = bind(e);
  if (rc) return rc;
  if (!e->has_lens || !e->has_tex) return fail(LFB_ERR_STATE, "set the lens and the aperture first");
  rc = check_params(P, true);
  if (rc) return rc;
  if (!light || !out) return fail(LFB_ERR_INVALID, "NULL light/out");
  const size_t n = (size_t)P->grid_n * P->grid_n;
  if (cap < n) return fail(LFB_ERR_INVALID, "out holds fewer than grid_n^2 records");
  if (lambda < 0 || lambda >= e->lens.n_lambda) return fail(LFB_ERR_INVALID, "lambda out of range");
  const bool direct = i < 0 && j < 0;
  if (!direct) {
    const int S = e->lens.stop_index;
    if (i < 0 || j <= i || j >= e->lens.n_surfaces || i == S || j == S) return fail(LFB_ERR_INVALID, "bad ghost pair");
  }
  rc = upload_constants(e);
  if (rc) return rc;
  rc = grow(&e->d_hits, &e->hits_cap, sizeof(lfb_ray_hit) * n);
  if (rc) return rc;
  Job J;
  JobId id = {0, direct ? -1 : i, direct ? -1 : j, lambda};
  fill_job(e, *P, *light, id, &J);
  const bool fast = is_exact_fast(*P), strict = P->precision == LFB_STRICT;
  if (fast) {
    StepD prog[LFB_MAX_STEPS];
    J.n_steps = build_program(e, lambda, id.i, id.j, prog);
    std::vector<char> staged(sizeof(StepD) * LFB_MAX_STEPS);
    put_steps(strict, staged.data(), 0, prog, J.n_steps);
    CU(cudaMemcpyAsync(e->d_dump_prog, staged.data(), (strict ? sizeof(StepD) : sizeof(StepF)) * (size_t)J.n_steps, cudaMemcpyHostToDevice, e->stream));
  }
  CU(cudaMemcpyAsync(e->d_dump_job, &J, sizeof(Job), cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));  // J and the staged program live on this call's stack
  if (P->mode == LFB_MODE_PARAXIAL_GRID) {
    CU(launch_paraxial_setup(e->d_dump_job, 1, P->physical_backward, e->stream));
    e->launches++;
  }
  const FrameGeom g = make_geom(e, *P, nullptr);
  if (fast && strict) CU(launch_exact_dump<double>(e->d_dump_job, (const StepD*)e->d_dump_prog, g, e->d_tex, e->d_hits, e->stream));
  else if (fast) CU(launch_exact_dump<float>(e->d_dump_job, (const StepF*)e->d_dump_prog, g, e->d_tex, e->d_hits, e->stream));
  else if (P->precision == LFB_FP64) CU(launch_trace_dump_f64(e->d_dump_job, g, P->mode, e->d_tex, e->d_hits, e->stream));
  else CU(launch_trace_dump_f32(e->d_dump_job, g, P->mode, e->d_tex, e->d_hits, e->stream));
  if (!fast && (rc = note_const_use(e))) return rc;
  e->launches++;
  CU(cudaMemcpyAsync(out, e->d_hits, sizeof(lfb_ray_hit) * n, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return LFB_OK;
}

extern "C" int lfb_ref_ghosts(lfb_engine* e, lfb_ref_ghost* out, int cap) {
  int rc = bind(e);
  if (rc) return rc;
  if (!out || cap < 0) return fail(LFB_ERR_INVALID, "bad out/cap");
  const int n = std::min(cap, e->n_ref_ghosts);
  if (n > 0) {
    CU(cudaMemcpyAsync(out, e->d_ghosts, sizeof(lfb_ref_ghost) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
  }
  return e->n_ref_ghosts;
}

// ---------------------------------------------------------------------------
// starburst
// ---------------------------------------------------------------------------
extern "C" int lfb_set_starburst_aperture(lfb_engine* e, const float* texels, int w, int h) {
  int rc = bind(e);
  if (rc) return rc;
  if (!texels || w < 1 || h < 1 || w > 16384 || h > 16384) return fail(LFB_ERR_INVALID, "bad aperture texture");
  // what CameraApertureTexture::init keeps beside the texels (camera.h:61-73): total_value and the bbox of texels > 0
  double total = 0;
  int bb[4] = {w, w, -1, -1};  // sic: min_y also starts at the width (camera.h:55)
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const float v = texels[(size_t)y * w + x];
      total += v;
      if (v > 0) {
        if (x < bb[0]) bb[0] = x;
        if (y < bb[1]) bb[1] = y;
        if (x > bb[2]) bb[2] = x;
        if (y > bb[3]) bb[3] = y;
      }
    }
  CU(cudaStreamSynchronize(e->stream));
  cudaFree(e->d_star_tex);
  e->d_star_tex = nullptr;
  CU(cudaMalloc((void**)&e->d_star_tex, sizeof(float) * (size_t)w * h));
  CU(cudaMemcpy(e->d_star_tex, texels, sizeof(float) * (size_t)w * h, cudaMemcpyHostToDevice));
  e->star_w = w; e->star_h = h; e->star_total = total;
  e->star_spectrum_valid = false;
  memcpy(e->star_bbox, bb, sizeof(bb));
  return LFB_OK;
}

namespace {
// Enqueue the starburst of a frame into a device buffer (out_dev: width*height pixels, stride/elem as given).
int starburst_device(lfb_engine* e, const lfb_light* lights, int n_lights, int width, int height, double flare_radius,
                     double flare_intensity, void* out_dev, size_t stride, int elem, int additive) {
  if (!e->d_star_tex) return fail(LFB_ERR_STATE, "set the starburst aperture first");
  if (n_lights < 1 || !lights) return fail(LFB_ERR_INVALID, "the starburst needs at least one light (flare_origins[0], pathtracer.cpp:919)");
  if (e->star_bbox[2] < e->star_bbox[0] || !(e->star_total > 0)) return fail(LFB_ERR_INVALID, "the starburst aperture is empty");
  StarFrame f;
  memset(&f, 0, sizeof(f));
  f.W = width; f.H = height; f.tw = e->star_w; f.th = e->star_h;
  f.bx0 = e->star_bbox[0]; f.by0 = e->star_bbox[1];
  f.bw = e->star_bbox[2] - e->star_bbox[0] + 1; f.bh = e->star_bbox[3] - e->star_bbox[1] + 1;
  // compute_phase(0, ...) pathtracer.cpp:918-934: only the FIRST light's origin drives the diffraction pattern
  f.org_x = ceil(lights[0].ns_x * (double)width);
  f.org_y = ceil(lights[0].ns_y * (double)height);
  f.lr = f.org_x - (double)width / 2.0;
  f.ud = -f.org_y + (double)height / 2.0;
  f.total = e->star_total;
  f.flare_radius = flare_radius;
  f.exponent = -flare_intensity + 3.0;  // :998-1001
  if (f.exponent <= 0) f.exponent = 2.0;
  // the pattern is periodic in the (integer) pixel offsets with period W_t (2 W_t for an odd W_t): evaluate one period
  // when the frame is larger (options.starburst_lattice = -1 forces one column / row per pixel)
  f.period = (e->star_w % 2 == 0) ? e->star_w : 2 * e->star_w;
  const bool lattice_ok = e->opt.starburst_lattice >= 0;
  f.lattice_x = lattice_ok && width > f.period;
  f.lattice_y = lattice_ok && height > f.period;
  f.n_col = f.lattice_x ? f.period : width;
  f.n_row = f.lattice_y ? f.period : height;
  // a real mask makes the lattice spectrum Hermitian: F(P-r, P-c) = conj F(r, c), so |F| needs half the rows
  f.herm = (f.lattice_x && f.lattice_y) ? 1 : 0;
  if (f.herm) f.n_row = f.period / 2 + 1;
  std::vector<double> L(5 * (size_t)n_lights);
  double rad_sum[3] = {0, 0, 0};
  for (int l = 0; l < n_lights; l++) {
    L[5 * l] = lights[l].ns_x * (double)width;   // fo_s, :1041 (no ceil here)
    L[5 * l + 1] = lights[l].ns_y * (double)height;
    for (int c = 0; c < 3; c++) { L[5 * l + 2 + c] = lights[l].radiance[c]; rad_sum[c] += lights[l].radiance[c]; }
  }
  if (starburst_scratch_bytes(f) > e->star_scratch_cap) e->star_spectrum_valid = false;  // the scratch buffer is about to move
  int rc = grow(&e->d_star_scratch, &e->star_scratch_cap, starburst_scratch_bytes(f));
  if (rc) return rc;
  // both axes on the periodic lattice: |F| depends on the mask alone -- computed once per mask, then only the pixel kernel runs
  const bool cacheable = f.herm && e->opt.starburst_cache >= 0;
  const bool cached = cacheable && e->star_spectrum_valid && e->star_spectrum_period == f.period;
  rc = grow(&e->d_star_lights, &e->star_lights_cap, sizeof(double) * L.size());
  if (rc) return rc;
  CU(cudaMemcpyAsync(e->d_star_lights, L.data(), sizeof(double) * L.size(), cudaMemcpyHostToDevice, e->stream));
  int n = 0;
  CU(launch_starburst(f, e->d_star_tex, e->d_star_scratch, e->d_star_lights, n_lights, rad_sum, out_dev, stride, elem, additive, cached, e->stream, &n));
  e->launches += (uint64_t)n;
  e->star_spectrum_valid = cacheable;
  e->star_spectrum_period = f.period;
  return LFB_OK;
}
}  // namespace

extern "C" int lfb_render_starburst(lfb_engine* e, const lfb_light* lights, int n_lights, int width, int height,
                                    double flare_radius, double flare_intensity, void* out, size_t stride, int elem, int additive) {
  int rc = bind(e);
  if (rc) return rc;
  if (width < 1 || height < 1 || width > 32768 || height > 32768 || !out) return fail(LFB_ERR_INVALID, "bad frame/out");
  if (elem != LFB_F32x3 && elem != LFB_F64x3) return fail(LFB_ERR_INVALID, "unknown out_elem");
  if (stride < elem_bytes(elem) || stride % (elem == LFB_F32x3 ? 4 : 8)) return fail(LFB_ERR_INVALID, "bad out_stride_bytes");
  const size_t npx = (size_t)width * height;
  const size_t out_bytes = (npx - 1) * stride + elem_bytes(elem);
  rc = grow(&e->d_out, &e->out_cap, out_bytes);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_frame0, e->stream));
  if (additive) CU(cudaMemcpyAsync(e->d_out, out, out_bytes, cudaMemcpyHostToDevice, e->stream));
  else if (stride != elem_bytes(elem)) CU(cudaMemsetAsync(e->d_out, 0, out_bytes, e->stream));
  CU(cudaEventRecord(e->ev_trace0, e->stream));
  rc = starburst_device(e, lights, n_lights, width, height, flare_radius, flare_intensity, e->d_out, stride, elem, additive);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_trace1, e->stream));
  CU(cudaMemcpyAsync(out, e->d_out, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaEventRecord(e->ev_frame1, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaEventElapsedTime(&e->last_frame_ms, e->ev_frame0, e->ev_frame1));
  CU(cudaEventElapsedTime(&e->last_trace_ms, e->ev_trace0, e->ev_trace1));
  e->timed = true;
  return LFB_OK;
}

// The displayable flare frame in one call, everything device-resident until the 8-bit frame comes back:
//   hdr = [base] + ghosts (any mode) + [starburst]   ->   toColor (util/image.h:208-223)   ->   RGBA8 (ImageBuffer, :53-62)
// i.e. what raytrace_pixel (:881-891) + raytrace_tile's toColor + frameBuffer do for the flare terms.  base_hdr (optional,
// host, W*H F64x3 packed) is the path-traced radiance to composite over; flare_radius < 0 skips the starburst;
// flip_vertical = 1 applies save_image's row flip (raytraced_renderer.cpp:739-742).  4 bytes per pixel cross PCIe, not 24.
namespace {
// [base | scene pass] + ghosts + [starburst] -> toColor -> RGBA8, everything device-resident until the 8-bit frame comes back
int frame_rgba8(lfb_engine* e, const lfb_camera* cam, const lfb_light* lights, int n_lights, const lfb_params* P, double flare_radius,
                double flare_intensity, const double* base_hdr, uint32_t* out_rgba8, int flip_vertical) {
  if (!e->has_lens || !e->has_tex) return fail(LFB_ERR_STATE, "set the lens and the aperture first");
  int rc = check_params(P, false);
  if (rc) return rc;
  if (n_lights < 0 || (n_lights > 0 && !lights) || !out_rgba8) return fail(LFB_ERR_INVALID, "bad lights/out");
  if (cam && !e->scene) return fail(LFB_ERR_STATE, "set the scene first (lfb_set_scene)");
  const size_t npx = (size_t)P->width * P->height;
  rc = grow(&e->d_hdr, &e->hdr_cap, npx * 24);
  if (rc) return rc;
  rc = grow(&e->d_rgba, &e->rgba_cap, npx * 4);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_frame0, e->stream));
  const bool have_base = base_hdr || cam;
  if (base_hdr) CU(cudaMemcpyAsync(e->d_hdr, base_hdr, npx * 24, cudaMemcpyHostToDevice, e->stream));
  if (cam) {
    CU(launch_scene(e->scene, cam, P->width, P->height, e->d_hdr, 24, LFB_F64x3, base_hdr ? 1 : 0, e->stream));
    e->launches++;
  }
  if (P->mode == LFB_MODE_REF_QUADS) {
    rc = render_ref_device(e, lights, n_lights, *P, e->d_hdr, 24, LFB_F64x3, have_base ? 1 : 0);
    if (rc) return rc;
  } else {
    rc = grow(&e->d_accum, &e->accum_cap, lfb_accum_bytes(P->width, P->height));
    if (rc) return rc;
    e->accum_clean = false;
    e->accum_dirty_w = e->accum_dirty_h = 0;
    rc = render_grid_device(e, lights, n_lights, *P, e->d_accum, 1);
    if (rc) return rc;
    const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
    CU(launch_finalize(e->d_accum, P->width, P->height, inv, e->d_hdr, 24, LFB_F64x3, have_base ? 1 : 0, e->stream));
    e->launches++;
  }
  if (flare_radius >= 0 && n_lights > 0) {
    rc = starburst_device(e, lights, n_lights, P->width, P->height, flare_radius, flare_intensity, e->d_hdr, 24, LFB_F64x3, 1);
    if (rc) return rc;
  }
  CU(launch_to_color(e->d_hdr, P->width, P->height, e->d_rgba, flip_vertical, e->stream));
  e->launches++;
  CU(cudaMemcpyAsync(out_rgba8, e->d_rgba, npx * 4, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaEventRecord(e->ev_frame1, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaEventElapsedTime(&e->last_frame_ms, e->ev_frame0, e->ev_frame1));
  CU(cudaEventElapsedTime(&e->last_trace_ms, e->ev_trace0, e->ev_trace1));
  return LFB_OK;
}
}  // namespace

extern "C" int lfb_render_frame_rgba8(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params* P,
                                      double flare_radius, double flare_intensity, const double* base_hdr, uint32_t* out_rgba8,
                                      int flip_vertical) {
  int rc = bind(e);
  if (rc) return rc;
  return frame_rgba8(e, nullptr, lights, n_lights, P, flare_radius, flare_intensity, base_hdr, out_rgba8, flip_vertical);
}

// ---------------------------------------------------------------------------
// the path-traced scene pass (scene.cu)
// ---------------------------------------------------------------------------
extern "C" int lfb_set_scene(lfb_engine* e, const lfb_scene* sc) {
  int rc = bind(e);
  if (rc) return rc;
  if (!sc) return fail(LFB_ERR_INVALID, "scene is NULL");
  if (sc->n_tri < 0 || sc->n_sph < 0 || sc->n_mat < 1 || sc->n_lights < 0) return fail(LFB_ERR_INVALID, "negative counts (or no material)");
  if ((sc->n_tri > 0 && (!sc->tri_pos || !sc->tri_nrm || !sc->tri_mat)) || (sc->n_sph > 0 && (!sc->spheres || !sc->sph_mat)) || !sc->materials ||
      (sc->n_lights > 0 && !sc->lights))
    return fail(LFB_ERR_INVALID, "NULL scene array");
  CU(cudaStreamSynchronize(e->stream));
  SceneStore* st = nullptr;
  const char* msg = nullptr;
  const cudaError_t err = scene_upload(sc, &st, &msg);
  if (msg) return fail(LFB_ERR_INVALID, msg);
  if (err == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(LFB_ERR_NOMEM, "cudaMalloc: out of device memory"); }
  if (err != cudaSuccess) return fail_cuda(err, "lfb_set_scene");
  scene_free(e->scene);
  e->scene = st;
  return LFB_OK;
}

extern "C" int lfb_render_scene(lfb_engine* e, const lfb_camera* cam, int width, int height, void* out, size_t stride, int elem, int additive) {
  int rc = bind(e);
  if (rc) return rc;
  if (!e->scene) return fail(LFB_ERR_STATE, "set the scene first (lfb_set_scene)");
  if (!cam) return fail(LFB_ERR_INVALID, "camera is NULL");
  if (width < 1 || height < 1 || width > 32768 || height > 32768) return fail(LFB_ERR_INVALID, "frame size out of range");
  rc = check_out_args(out, stride, elem);
  if (rc) return rc;
  const size_t npx = (size_t)width * height;
  const size_t out_bytes = (npx - 1) * stride + elem_bytes(elem);
  rc = grow(&e->d_out, &e->out_cap, out_bytes);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev_frame0, e->stream));
  if (additive) CU(cudaMemcpyAsync(e->d_out, out, out_bytes, cudaMemcpyHostToDevice, e->stream));
  else if (stride != elem_bytes(elem)) CU(cudaMemsetAsync(e->d_out, 0, out_bytes, e->stream));
  CU(cudaEventRecord(e->ev_trace0, e->stream));
  CU(launch_scene(e->scene, cam, width, height, e->d_out, stride, elem, additive, e->stream));
  e->launches++;
  CU(cudaEventRecord(e->ev_trace1, e->stream));
  CU(cudaMemcpyAsync(out, e->d_out, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaEventRecord(e->ev_frame1, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaEventElapsedTime(&e->last_frame_ms, e->ev_frame0, e->ev_frame1));
  CU(cudaEventElapsedTime(&e->last_trace_ms, e->ev_trace0, e->ev_trace1));
  e->timed = true;
  return LFB_OK;
}

extern "C" int lfb_render_composite_rgba8(lfb_engine* e, const lfb_camera* cam, const lfb_light* lights, int n_lights, const lfb_params* P,
                                          double flare_radius, double flare_intensity, uint32_t* out_rgba8, int flip_vertical) {
  int rc = bind(e);
  if (rc) return rc;
  if (!cam) return fail(LFB_ERR_INVALID, "camera is NULL");
  return frame_rgba8(e, cam, lights, n_lights, P, flare_radius, flare_intensity, nullptr, out_rgba8, flip_vertical);
}

// ---------------------------------------------------------------------------
// single-process multi-GPU
// ---------------------------------------------------------------------------
namespace {
// Helper threads of a multi-GPU engine: the CALLER stays one thread (the reference's model), but enqueuing a frame's work on
// 8 devices from one thread costs ~0.1 ms per device (measured: 0.91 ms of a 0.98 ms frame), so each device gets its own
// enqueuing thread; run(f) executes f(d) for every device (d = 0 on the calling thread) and returns when all are done.
struct DevicePool {
  int n = 0;
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  const std::function<int(int)>* job = nullptr;
  unsigned long long generation = 0;
  int pending = 0;
  bool quit = false;
  std::vector<int> rc;
  std::vector<std::string> err;

  void start(int n_devices, const int* device_ids) {
    n = n_devices;
    rc.assign(n, 0);
    err.assign(n, std::string());
    for (int d = 1; d < n; d++) {
      const int dev = device_ids[d];
      threads.emplace_back([this, d, dev] {
        cudaSetDevice(dev);
        unsigned long long seen = 0;
        for (;;) {
          const std::function<int(int)>* f;
          {
            std::unique_lock<std::mutex> lock(mu);
            cv_go.wait(lock, [&] { return quit || generation != seen; });
            if (quit) return;
            seen = generation;
            f = job;
          }
          const int r = (*f)(d);
          {
            std::lock_guard<std::mutex> lock(mu);
            rc[d] = r;
            if (r) err[d] = g_err;
            if (--pending == 0) cv_done.notify_all();
          }
        }
      });
    }
  }
  int run(const std::function<int(int)>& f) {
    {
      std::lock_guard<std::mutex> lock(mu);
      job = &f;
      pending = n - 1;
      generation++;
    }
    cv_go.notify_all();
    rc[0] = f(0);
    if (rc[0]) err[0] = g_err;
    {
      std::unique_lock<std::mutex> lock(mu);
      cv_done.wait(lock, [&] { return pending == 0; });
    }
    for (int d = 0; d < n; d++)
      if (rc[d]) { g_err = err[d]; return rc[d]; }
    return LFB_OK;
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lock(mu);
      quit = true;
    }
    cv_go.notify_all();
    for (std::thread& t : threads) t.join();
    threads.clear();
  }
};
}  // namespace

struct lfb_multi {
  int n = 0;
  DevicePool pool;
  lfb_engine* eng[LFB_MAX_PEERS] = {nullptr};
  unsigned long long* accum[LFB_MAX_PEERS] = {nullptr};  // per device: sums + tile map (peer-accessible)
  size_t accum_cap = 0;
  bool accum_clean = false;
  int accum_w = 0, accum_h = 0;
  unsigned* state[LFB_MAX_PEERS] = {nullptr};            // per device: its tile state of the caller's buffer
  size_t state_cap = 0;
  unsigned* h_count = nullptr;                           // page-locked, portable: tiles written per device
  cudaEvent_t ev_traced[LFB_MAX_PEERS] = {nullptr}, ev_red0[LFB_MAX_PEERS] = {nullptr}, ev_red1[LFB_MAX_PEERS] = {nullptr};
  const void* out = nullptr;
  int out_w = 0, out_h = 0, out_elem = 0;
  size_t out_stride = 0;
  float stage_ms[4] = {0, 0, 0, 0};
};

extern "C" void lfb_destroy_multi(lfb_multi* m) {
  if (!m) return;
  m->pool.stop();
  for (int d = 0; d < m->n; d++) {
    if (!m->eng[d]) continue;
    cudaSetDevice(m->eng[d]->device);
    cudaStreamSynchronize(m->eng[d]->stream);
    cudaFree(m->accum[d]); cudaFree(m->state[d]);
    if (m->ev_traced[d]) cudaEventDestroy(m->ev_traced[d]);
    if (m->ev_red0[d]) cudaEventDestroy(m->ev_red0[d]);
    if (m->ev_red1[d]) cudaEventDestroy(m->ev_red1[d]);
    lfb_destroy(m->eng[d]);
  }
  if (m->h_count) cudaFreeHost(m->h_count);
  delete m;
}

extern "C" int lfb_create_multi(lfb_multi** out, const int* device_ids, int n_devices, const lfb_options* options) {
  if (!out) return fail(LFB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!device_ids || n_devices < 1 || n_devices > LFB_MAX_PEERS) return fail(LFB_ERR_INVALID, "bad device list");
  for (int a = 0; a < n_devices; a++)
    for (int b = a + 1; b < n_devices; b++)
      if (device_ids[a] == device_ids[b]) return fail(LFB_ERR_INVALID, "a device is listed twice");
  lfb_multi* m = new (std::nothrow) lfb_multi();
  if (!m) return fail(LFB_ERR_NOMEM, "out of host memory");
  m->n = n_devices;
  int rc = LFB_OK;
  for (int d = 0; d < n_devices && rc == LFB_OK; d++) rc = lfb_create_ex(&m->eng[d], device_ids[d], options);
  for (int d = 0; d < n_devices && rc == LFB_OK; d++) {
    cudaSetDevice(device_ids[d]);
    for (int q = 0; q < n_devices; q++) {
      if (q == d) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, device_ids[d], device_ids[q]);
      if (!can) { rc = fail(LFB_ERR_CUDA, "devices " + std::to_string(device_ids[d]) + " and " + std::to_string(device_ids[q]) + " have no peer access"); break; }
      const cudaError_t err = cudaDeviceEnablePeerAccess(device_ids[q], 0);
      if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) { rc = fail_cuda(err, "cudaDeviceEnablePeerAccess"); break; }
      cudaGetLastError();
    }
    if (rc) break;
    cudaError_t err = cudaEventCreateWithFlags(&m->ev_traced[d], cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreate(&m->ev_red0[d]);
    if (err == cudaSuccess) err = cudaEventCreate(&m->ev_red1[d]);
    if (err != cudaSuccess) rc = fail_cuda(err, "cudaEventCreate");
  }
  if (rc == LFB_OK && cudaHostAlloc((void**)&m->h_count, sizeof(unsigned) * LFB_MAX_PEERS, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess)
    rc = fail(LFB_ERR_NOMEM, "cudaHostAlloc");
  if (rc) { const std::string keep = g_err; lfb_destroy_multi(m); g_err = keep; return rc; }
  m->pool.start(n_devices, device_ids);
  *out = m;
  return LFB_OK;
}

extern "C" int lfb_multi_set_lens(lfb_multi* m, const lfb_lens* lens) {
  if (!m) return fail(LFB_ERR_INVALID, "multi engine is NULL");
  for (int d = 0; d < m->n; d++) {
    const int rc = lfb_set_lens(m->eng[d], lens);
    if (rc) return rc;
  }
  return LFB_OK;
}

extern "C" int lfb_multi_set_aperture(lfb_multi* m, const float* texels, int w, int h) {
  if (!m) return fail(LFB_ERR_INVALID, "multi engine is NULL");
  for (int d = 0; d < m->n; d++) {
    const int rc = lfb_set_aperture(m->eng[d], texels, w, h);
    if (rc) return rc;
  }
  return LFB_OK;
}

extern "C" int lfb_render_ghosts_multi(lfb_multi* m, const lfb_light* lights, int n_lights, const lfb_params* P, void* out, size_t stride,
                                       int elem, int out_is_clear, int* tiles_written) {
  if (!m) return fail(LFB_ERR_INVALID, "multi engine is NULL");
  int rc = check_params(P, true);
  if (rc) return rc;
  if (n_lights < 0 || (n_lights > 0 && !lights)) return fail(LFB_ERR_INVALID, "bad lights");
  rc = check_out_args(out, stride, elem);
  if (rc) return rc;
  for (int d = 0; d < m->n; d++)
    if (!m->eng[d]->has_lens || !m->eng[d]->has_tex) return fail(LFB_ERR_STATE, "set the lens and the aperture first");
  if (tiles_written) *tiles_written = 0;
  const int W = P->width, H = P->height, n = m->n;
  const bool same = out == m->out && W == m->out_w && H == m->out_h && stride == m->out_stride && elem == m->out_elem;
  if (!out_is_clear && !same) return fail(LFB_ERR_STATE, "lfb_render_ghosts_multi: `out` is not the buffer of the previous call; pass out_is_clear = 1 with a zeroed buffer");
  const auto host_t0 = std::chrono::steady_clock::now();
  // the caller's buffer as every device sees it (page-locked + mapped), or the fallback: device 0's own frame, copied whole
  void* out_dev[LFB_MAX_PEERS];
  bool zero_copy = true;
  for (int d = 0; d < n; d++) {
    CU(cudaSetDevice(m->eng[d]->device));
    out_dev[d] = nullptr;
    if (cudaHostGetDevicePointer(&out_dev[d], out, 0) != cudaSuccess || !out_dev[d]) { cudaGetLastError(); zero_copy = false; break; }
  }
  const AccumLayout lay = accum_layout(W, H);
  const size_t st_bytes = tile_state_bytes(W, H);
  lfb_engine* e0 = m->eng[0];
  if (!zero_copy) {  // the reduce writes a device-0 frame that is then copied whole: the engine keeps that frame's tile state
    CU(cudaSetDevice(e0->device));
    rc = grow(&e0->d_out, &e0->out_cap, ((size_t)W * H - 1) * stride + elem_bytes(elem));
    if (rc) return rc;
  }
  const bool fresh_state = out_is_clear || !same || st_bytes > m->state_cap || !zero_copy;
  const bool fresh_accum = lay.total > m->accum_cap || !(m->accum_clean && m->accum_w == W && m->accum_h == H);
  const size_t frame_bytes = ((size_t)W * H - 1) * stride + elem_bytes(elem);
  PeerAccums A;
  memset(&A, 0, sizeof(A));
  A.n = n;
  const double inv = ldexp(1.0, -(P->fixed_point_bits > 0 ? P->fixed_point_bits : 40));
  // stage 1 (one enqueuing thread per device): buffers, then every device traces its shard
  const std::function<int(int)> stage1 = [&](int d) -> int {
    lfb_engine* e = m->eng[d];
    CU(cudaSetDevice(e->device));
    if (lay.total > m->accum_cap) {
      cudaFree(m->accum[d]);
      m->accum[d] = nullptr;
      if (cudaMalloc((void**)&m->accum[d], lay.total) != cudaSuccess) { cudaGetLastError(); return fail(LFB_ERR_NOMEM, "cudaMalloc: out of device memory"); }
    }
    if (st_bytes > m->state_cap) {
      cudaFree(m->state[d]);
      m->state[d] = nullptr;
      if (cudaMalloc((void**)&m->state[d], st_bytes) != cudaSuccess) { cudaGetLastError(); return fail(LFB_ERR_NOMEM, "cudaMalloc: out of device memory"); }
    }
    if (fresh_accum) CU(cudaMemsetAsync(m->accum[d], 0, lay.total, e->stream));
    if (fresh_state) CU(cudaMemsetAsync(m->state[d], 0, st_bytes, e->stream));
    if (d == 0 && !zero_copy) CU(cudaMemsetAsync(e0->d_out, 0, frame_bytes, e0->stream));
    lfb_params Pd = *P;
    Pd.shard_index = n > 1 ? d : 0; Pd.shard_count = n > 1 ? n : 0;
    const int r = render_grid_device(e, lights, n_lights, Pd, m->accum[d], 0);
    if (r) return r;
    CU(cudaEventRecord(m->ev_traced[d], e->stream));
    return LFB_OK;
  };
  rc = m->pool.run(stage1);
  m->accum_cap = std::max(m->accum_cap, lay.total);
  m->state_cap = std::max(m->state_cap, st_bytes);
  m->accum_clean = false;
  if (rc) return rc;
  m->out = out; m->out_w = W; m->out_h = H; m->out_stride = stride; m->out_elem = elem;
  for (int d = 0; d < n; d++) A.ptr[d] = m->accum[d];
  // stage 2 (after every device's trace has been RECORDED): every device reduces its interleaved share of the dirty tiles over
  // peer memory, once all traces are done
  const std::function<int(int)> stage2 = [&](int d) -> int {
    lfb_engine* e = m->eng[d];
    CU(cudaSetDevice(e->device));
    for (int q = 0; q < n; q++)
      if (q != d) CU(cudaStreamWaitEvent(e->stream, m->ev_traced[q], 0));
    unsigned* count_dev = nullptr;
    CU(cudaHostGetDevicePointer((void**)&count_dev, m->h_count + d, 0));
    CU(cudaEventRecord(m->ev_red0[d], e->stream));
    CU(launch_tiles(A, d, W, H, inv, zero_copy ? out_dev[d] : (void*)e0->d_out, stride, elem, m->state[d], count_dev, e->opt.reduce_ctas, e->stream));
    e->launches++;
    CU(cudaEventRecord(m->ev_red1[d], e->stream));
    return LFB_OK;
  };
  rc = m->pool.run(stage2);
  if (rc) return rc;
  const auto host_t1 = std::chrono::steady_clock::now();
  for (int d = 0; d < n; d++) {
    CU(cudaSetDevice(m->eng[d]->device));
    CU(cudaStreamSynchronize(m->eng[d]->stream));
  }
  if (!zero_copy) {
    CU(cudaSetDevice(e0->device));
    CU(cudaMemcpy(out, e0->d_out, ((size_t)W * H - 1) * stride + elem_bytes(elem), cudaMemcpyDeviceToHost));
    m->out = nullptr;  // every pixel was rewritten: nothing to remember
  }
  const auto host_t2 = std::chrono::steady_clock::now();
  m->accum_clean = true; m->accum_w = W; m->accum_h = H;
  int total = 0;
  float t_trace = 0, t_red = 0;
  for (int d = 0; d < n; d++) {
    total += (int)m->h_count[d];
    float a = 0, b = 0;
    CU(cudaSetDevice(m->eng[d]->device));
    cudaEventElapsedTime(&a, m->eng[d]->ev_trace0, m->eng[d]->ev_trace1);
    cudaEventElapsedTime(&b, m->ev_red0[d], m->ev_red1[d]);
    cudaGetLastError();
    t_trace = std::max(t_trace, a); t_red = std::max(t_red, b);
  }
  m->stage_ms[0] = t_trace; m->stage_ms[1] = t_red;
  m->stage_ms[2] = std::chrono::duration<float, std::milli>(host_t2 - host_t0).count();
  m->stage_ms[3] = std::chrono::duration<float, std::milli>(host_t1 - host_t0).count();
  if (tiles_written) *tiles_written = zero_copy ? total : -1;
  return LFB_OK;
}

extern "C" int lfb_multi_stats(lfb_multi* m, float stage_ms[4], int* n_devices) {
  if (!m) return fail(LFB_ERR_INVALID, "multi engine is NULL");
  if (stage_ms) memcpy(stage_ms, m->stage_ms, sizeof(m->stage_ms));
  if (n_devices) *n_devices = m->n;
  return LFB_OK;
}

// ---------------------------------------------------------------------------
// accounting (host only; usable without a device)
// ---------------------------------------------------------------------------
extern "C" int lfb_count_work(const lfb_lens* L, const lfb_params* P, int n_lights, double* rays, double* interactions,
                              int* jobs) {
  int rc = check_lens(L);
  if (rc) return rc;
  rc = check_params(P, false);
  if (rc) return rc;
  if (n_lights < 0) return fail(LFB_ERR_INVALID, "n_lights < 0");
  std::vector<JobId> ids;
  list_jobs(*L, *P, n_lights, ids);
  const double per_job = P->mode == LFB_MODE_REF_QUADS ? 2.0 : (double)P->grid_n * P->grid_n;
  double r = 0, it = 0;
  for (const JobId& id : ids) { r += per_job; it += per_job * interactions_of(*L, id.i, id.j); }
  if (rays) *rays = r;
  if (interactions) *interactions = it;
  if (jobs) *jobs = (int)ids.size();
  return LFB_OK;
}

extern "C" int lfb_list_jobs(const lfb_lens* L, const lfb_params* P, int n_lights, int32_t* out, int cap) {
  int rc = check_lens(L);
  if (rc) return rc;
  rc = check_params(P, false);
  if (rc) return rc;
  if (n_lights < 0 || cap < 0 || (cap > 0 && !out)) return fail(LFB_ERR_INVALID, "bad arguments");
  std::vector<JobId> ids;
  list_jobs(*L, *P, n_lights, ids);
  for (int q = 0; q < (int)ids.size() && q < cap; q++) {
    out[4 * q] = ids[q].light; out[4 * q + 1] = ids[q].i; out[4 * q + 2] = ids[q].j; out[4 * q + 3] = ids[q].lambda;
  }
  return (int)ids.size();
}

extern "C" int lfb_stats(lfb_engine* e, uint64_t* kernel_launches, float* last_trace_ms, float* last_frame_ms) {
  int rc = bind(e);
  if (rc) return rc;
  if (kernel_launches) *kernel_launches = e->launches;
  if (last_trace_ms) {
    // the device-resident API does not synchronise: resolve the trace events on demand
    if (e->timed && cudaEventQuery(e->ev_trace1) == cudaSuccess) cudaEventElapsedTime(&e->last_trace_ms, e->ev_trace0, e->ev_trace1);
    cudaGetLastError();
    *last_trace_ms = e->last_trace_ms;
  }
  if (last_frame_ms) *last_frame_ms = e->last_frame_ms;
  return LFB_OK;
}

extern "C" int lfb_exec_stats(lfb_engine* e, uint64_t out[4]) {
  int rc = bind(e);
  if (rc) return rc;
  if (!out) return fail(LFB_ERR_INVALID, "out is NULL");
  if (!e->opt.collect_stats) return fail(LFB_ERR_STATE, "the engine was not created with lfb_options.collect_stats = 1");
  CU(cudaStreamSynchronize(e->stream));
  unsigned long long h[4];
  CU(cudaMemcpy(h, e->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
  out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; out[3] = e->last_families ? 1 : 0;
  return LFB_OK;
}

extern "C" int lfb_probe_peaks(lfb_engine* e, double* fp32_flops, double* mufu_ops, double* sm_clock_hz) {
  int rc = bind(e);
  if (rc) return rc;
  CU(probe_peaks(e->device, e->stream, fp32_flops, mufu_ops, sm_clock_hz));
  e->launches += 10;
  return LFB_OK;
}

// ---------------------------------------------------------------------------
// pinned host memory helpers
// ---------------------------------------------------------------------------
extern "C" void* lfb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

extern "C" void lfb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" int lfb_host_register(void* p, size_t bytes) {
  if (!p || bytes == 0) return fail(LFB_ERR_INVALID, "lfb_host_register: empty range");
  const cudaError_t rc = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
  if (rc != cudaSuccess) { cudaGetLastError(); return fail(LFB_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(rc)); }
  return LFB_OK;
}

extern "C" void* lfb_host_device_pointer(void* p) {
  void* d = nullptr;
  if (!p || cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return d;
}

extern "C" int lfb_host_unregister(void* p) {
  if (!p) return LFB_OK;
  const cudaError_t rc = cudaHostUnregister(p);
  if (rc != cudaSuccess) { cudaGetLastError(); return fail(LFB_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(rc)); }
  return LFB_OK;
}
