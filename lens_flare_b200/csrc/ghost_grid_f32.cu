// ghost_grid_f32.cu -- FP32 instantiation of the oracle-order ray-grid kernels (ghost_grid_impl.cuh), FMA contraction on:
// PARAXIAL_GRID in FP32 (a 2x2 matrix apply per ray).  EXACT_GRID in FP32 is exact_f32.cu.
#define LFB_TU f32
#include "ghost_grid_impl.cuh"

namespace lfb {

cudaError_t upload_lens_f32(const DevLens& h, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(f32::c_lens, &h, sizeof(DevLens), 0, cudaMemcpyHostToDevice, s);
}

cudaError_t launch_trace_splat_f32(const Job* jobs, int n_jobs, const FrameGeom& g, int mode, const float* tex,
                                   unsigned long long* accum, cudaStream_t s) {
  if (mode != LFB_MODE_PARAXIAL_GRID) return cudaErrorInvalidValue;  // EXACT_GRID FP32: launch_exact_*<float>
  return f32::launch_trace_splat_t<float>(jobs, n_jobs, g, mode, tex, accum, s);
}

cudaError_t launch_trace_dump_f32(const Job* job, const FrameGeom& g, int mode, const float* tex, lfb_ray_hit* out, cudaStream_t s) {
  if (mode != LFB_MODE_PARAXIAL_GRID) return cudaErrorInvalidValue;
  return f32::launch_trace_dump_t<float>(job, g, mode, tex, out, s);
}

}  // namespace lfb
