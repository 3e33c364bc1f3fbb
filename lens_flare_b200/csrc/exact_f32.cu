// exact_f32.cu -- the EXACT_GRID throughput kernels (exact_trace.cuh) for T = float (LFB_FP32), FMA contraction on.
#include "exact_trace.cuh"

namespace lfb {

template <> int exact_prefix_parts<float>() { return xt::PrefixIO<float>::kParts; }
template <>
cudaError_t launch_exact_prefix<float>(const Job* slots, const StepT<float>* progs, int n_slots, const FrameGeom& g, const float* tex,
                                     float4* prefix, unsigned long long* accum_for_direct, bool stats, cudaStream_t s) {
  return xt::launch_prefix_t<float>(slots, progs, n_slots, g, tex, prefix, accum_for_direct, stats, s);
}
template <>
cudaError_t launch_exact_ghosts<float>(const Job* jobs, const StepT<float>* progs, const unsigned* heads, int n_jobs, const FrameGeom& g,
                                     const float* tex, unsigned long long* accum, int ctas_per_sm, bool stats, cudaStream_t s) {
  return xt::launch_ghosts_t<float>(jobs, progs, heads, n_jobs, g, tex, accum, ctas_per_sm, stats, s);
}
template <>
cudaError_t launch_exact_families<float>(const Job* fams, const StepT<float>* fam_progs, const unsigned* heads, int n_fams, const Job* slots,
                                       const StepT<float>* slot_progs, const FrameGeom& g, const float* tex, unsigned long long* accum,
                                       int ctas_per_sm, bool stats, cudaStream_t s) {
  return xt::launch_families_t<float>(fams, fam_progs, heads, n_fams, slots, slot_progs, g, tex, accum, ctas_per_sm, stats, s);
}
template <>
cudaError_t launch_exact_dump<float>(const Job* job, const StepT<float>* prog, const FrameGeom& g, const float* tex, lfb_ray_hit* out,
                                   cudaStream_t s) {
  return xt::launch_dump_t<float>(job, prog, g, tex, out, s);
}

}  // namespace lfb
