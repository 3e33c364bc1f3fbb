// ghost_grid_f32.cu -- FP32 kernels of the ray-grid path (the throughput path), FMA contraction on.
//   EXACT_GRID     exact_f32.cuh   step programs, fast scalar math, two passes with survivor compaction,
//                                  shared-memory sensor tile
//   PARAXIAL_GRID  ghost_grid_impl.cuh instantiated for float (a 2x2 matrix apply per ray)
#define LFB_TU f32
#include "ghost_grid_impl.cuh"
#include "exact_f32.cuh"
#include <stdlib.h>

namespace lfb {

cudaError_t upload_lens_f32(const DevLens& h, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(f32::c_lens, &h, sizeof(DevLens), 0, cudaMemcpyHostToDevice, s);
}

cudaError_t launch_trace_splat_f32(const Job* jobs, const Step* progs, int n_jobs, const FrameGeom& g, int mode,
                                   const float* tex, unsigned long long* accum, cudaStream_t s) {
  if (n_jobs <= 0) return cudaSuccess;
  if (mode == LFB_MODE_PARAXIAL_GRID) return f32::launch_trace_splat_t<float>(jobs, n_jobs, g, mode, tex, accum, s);
  auto blocks = [&](int rx, int ry) {  // patches tile the upper half of the grid (mirror symmetry, exact_f32.cuh)
    return (unsigned)n_jobs * (unsigned)(((g.N + 16 * rx - 1) / (16 * rx)) * (((g.N + 1) / 2 + 16 * ry - 1) / (16 * ry)));
  };
  if (g.prefix) {  // v5: ray states come from the prefix cache (launch_prefix_f32 ran first)
    // block size 64 / 128 / 256 (16 x 4 / 8 / 16 ray pairs per round); LFB_EXACT_BLOCK selects, default 128 (measured best)
    static int bt = 0;
    if (!bt) { const char* env = getenv("LFB_EXACT_BLOCK"); bt = env ? atoi(env) : 128; if (bt != 64 && bt != 256) bt = 128; }
    auto blocks2 = [&](int rx, int ry, int rows) {
      return (unsigned)n_jobs * (unsigned)(((g.N + 16 * rx - 1) / (16 * rx)) * (((g.N + 1) / 2 + rows * ry - 1) / (rows * ry)));
    };
#define LFB_LAUNCH2(RX, RY, MB, BT) xf32::exact_splat2_kernel<RX, RY, MB, BT><<<blocks2(RX, RY, BT / 16), BT, 0, s>>>(jobs, progs, g, tex, accum)
#define LFB_PATCH2(MB, BT)                        \
  do {                                            \
    if (g.patch >= 4) LFB_LAUNCH2(2, 2, MB, BT);  \
    else if (g.patch >= 2) LFB_LAUNCH2(2, 1, MB, BT); \
    else LFB_LAUNCH2(1, 1, MB, BT);               \
  } while (0)
    static int warp_mode = -1;
    if (warp_mode < 0) { const char* env = getenv("LFB_EXACT_WARP"); warp_mode = env ? atoi(env) : 1; }
    // read per launch (not cached) so that one process can compare the generations
    const char* ilp_env = getenv("LFB_EXACT_ILP");
    const int ilp_mode = ilp_env ? atoi(ilp_env) : 1;
    if (warp_mode && g.patch <= 1 && ilp_mode == 2) {  // v6i: two ray pairs per thread in lockstep (opt-in)
      const dim3 nb((g.N + 15) / 16, ((g.N + 1) / 2 + 15) / 16, n_jobs);  // 128 threads = 16 x 8 lanes x 2 rows each
      if (nb.y > 65535u || nb.z > 65535u) return cudaErrorInvalidConfiguration;
      if (g.pad >= 10) xf32::exact_splat4_kernel<10, 128><<<nb, 128, 0, s>>>(jobs, progs, g, tex, accum);
      else xf32::exact_splat4_kernel<8, 128><<<nb, 128, 0, s>>>(jobs, progs, g, tex, accum);
      return cudaGetLastError();
    }
    if (warp_mode && g.patch <= 1) {  // v6: warp-autonomous splat (default)
      const int rows = bt / 16;
      const dim3 nb((g.N + 15) / 16, ((g.N + 1) / 2 + rows - 1) / rows, n_jobs);  // (patch column, patch row, job)
      if (nb.y > 65535u || nb.z > 65535u) return cudaErrorInvalidConfiguration;
      if (bt == 64) xf32::exact_splat3_kernel<24, 64><<<nb, 64, 0, s>>>(jobs, progs, g, tex, accum);
      else if (bt == 256) xf32::exact_splat3_kernel<6, 256><<<nb, 256, 0, s>>>(jobs, progs, g, tex, accum);
      else xf32::exact_splat3_kernel<12, 128><<<nb, 128, 0, s>>>(jobs, progs, g, tex, accum);
      return cudaGetLastError();
    }
    if (bt == 64) {
      LFB_PATCH2(24, 64);
    } else if (bt == 128) {
      if (g.pad >= 7) LFB_PATCH2(16, 128); else if (g.pad >= 6) LFB_PATCH2(12, 128);
      else LFB_PATCH2(8, 128);
    } else {
      if (g.pad >= 6) LFB_PATCH2(6, 256);
      else if (g.pad == 5) LFB_PATCH2(5, 256);
      else LFB_PATCH2(4, 256);
    }
#undef LFB_PATCH2
#undef LFB_LAUNCH2
    return cudaGetLastError();
  }
  // patch shape (rays per thread in pass 1) x resident CTAs per SM the register allocation targets
  const int minb = g.pad;  // 0 (default) -> 4
  // g.lut set: the one-pass kernel with tabulated reflectances (v4); else the two-pass closed-form kernel (v3)
#define LFB_LAUNCH(RX, RY, MB)                                                                                              \
  do {                                                                                                                      \
    if (g.lut) xf32::exact_splat1_kernel<RX, RY, MB><<<blocks(RX, RY), xf32::kThreads, 0, s>>>(jobs, progs, g, tex, accum); \
    else xf32::exact_splat_kernel<RX, RY, MB><<<blocks(RX, RY), xf32::kThreads, 0, s>>>(jobs, progs, g, tex, accum);        \
  } while (0)
#define LFB_PATCH(MB)                       \
  do {                                      \
    if (g.patch >= 4) LFB_LAUNCH(2, 2, MB); \
    else if (g.patch >= 2) LFB_LAUNCH(2, 1, MB); \
    else LFB_LAUNCH(1, 1, MB);              \
  } while (0)
  if (minb >= 6) LFB_PATCH(6);
  else if (minb == 5) LFB_PATCH(5);
  else LFB_PATCH(4);
#undef LFB_PATCH
#undef LFB_LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_prefix_f32(const Job* slots, const Step* progs, int n_slots, const FrameGeom& g, const float* tex, float4* prefix,
                              unsigned long long* accum_for_direct, cudaStream_t s) {
  if (n_slots <= 0) return cudaSuccess;
  const dim3 nb((g.N + 15) / 16, ((g.N + 1) / 2 + 15) / 16, n_slots);  // (patch column, patch row, slot)
  if (nb.z > 65535u) return cudaErrorInvalidConfiguration;
  xf32::prefix_kernel<<<nb, xf32::kThreads, 0, s>>>(slots, progs, g, tex, prefix, accum_for_direct);
  return cudaGetLastError();
}

cudaError_t launch_family_f32(const Job* fams, const Step* fam_progs, int n_fams, const Job* slots, const Step* slot_progs,
                              const FrameGeom& g, const float* tex, unsigned long long* accum, cudaStream_t s) {
  if (n_fams <= 0) return cudaSuccess;
  static int cfg = -1;  // LFB_FAMILY_CFG: 0 = 128 threads / 10 CTAs per SM (default), 1 = 128 / 12, 2 = 64 / 20, 3 = 64 / 24, 4 = 256 / 5
  if (cfg < 0) { const char* env = getenv("LFB_FAMILY_CFG"); cfg = env ? atoi(env) : 0; }
  const int bt = (cfg == 2 || cfg == 3) ? 64 : (cfg == 4 ? 256 : 128);
  const int rows = bt / 16;
  const dim3 nb((g.N + 15) / 16, ((g.N + 1) / 2 + rows - 1) / rows, n_fams);  // (patch column, patch row, family)
  if (nb.y > 65535u || nb.z > 65535u) return cudaErrorInvalidConfiguration;
  if (cfg == 1) xf32::exact_family_kernel<12, 128><<<nb, 128, 0, s>>>(fams, fam_progs, slots, slot_progs, g, tex, accum);
  else if (cfg == 2) xf32::exact_family_kernel<20, 64><<<nb, 64, 0, s>>>(fams, fam_progs, slots, slot_progs, g, tex, accum);
  else if (cfg == 3) xf32::exact_family_kernel<24, 64><<<nb, 64, 0, s>>>(fams, fam_progs, slots, slot_progs, g, tex, accum);
  else if (cfg == 4) xf32::exact_family_kernel<5, 256><<<nb, 256, 0, s>>>(fams, fam_progs, slots, slot_progs, g, tex, accum);
  else xf32::exact_family_kernel<10, 128><<<nb, 128, 0, s>>>(fams, fam_progs, slots, slot_progs, g, tex, accum);
  return cudaGetLastError();
}

cudaError_t launch_trace_dump_f32(const Job* job, const Step* prog, const FrameGeom& g, int mode, const float* tex,
                                  lfb_ray_hit* out, cudaStream_t s) {
  if (mode == LFB_MODE_PARAXIAL_GRID) return f32::launch_trace_dump_t<float>(job, g, mode, tex, out, s);
  xf32::exact_dump_kernel<<<(unsigned)g.tiles_per_job, xf32::kThreads, 0, s>>>(job, prog, g, tex, out);
  return cudaGetLastError();
}

}  // namespace lfb
