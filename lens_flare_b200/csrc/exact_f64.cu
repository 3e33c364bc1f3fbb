// exact_f64.cu -- the EXACT_GRID throughput kernels (exact_trace.cuh) for T = double (LFB_STRICT: FP64 geometry, FP32 weights), FMA contraction on.
#include "exact_trace.cuh"

namespace lfb {

template <> int exact_prefix_parts<double>() { return xt::PrefixIO<double>::kParts; }
template <>
cudaError_t launch_exact_prefix<double>(const Job* slots, const StepT<double>* progs, int n_slots, const FrameGeom& g, const float* tex,
                                     float4* prefix, unsigned long long* accum_for_direct, bool stats, cudaStream_t s) {
  return xt::launch_prefix_t<double>(slots, progs, n_slots, g, tex, prefix, accum_for_direct, stats, s);
}
template <>
cudaError_t launch_exact_ghosts<double>(const Job* jobs, const StepT<double>* progs, const unsigned* heads, int n_jobs, const FrameGeom& g,
                                     const float* tex, unsigned long long* accum, int ctas_per_sm, bool stats, cudaStream_t s) {
  return xt::launch_ghosts_t<double>(jobs, progs, heads, n_jobs, g, tex, accum, ctas_per_sm, stats, s);
}
template <>
cudaError_t launch_exact_families<double>(const Job* fams, const StepT<double>* fam_progs, const unsigned* heads, int n_fams, const Job* slots,
                                       const StepT<double>* slot_progs, const FrameGeom& g, const float* tex, unsigned long long* accum,
                                       int ctas_per_sm, bool stats, cudaStream_t s) {
  return xt::launch_families_t<double>(fams, fam_progs, heads, n_fams, slots, slot_progs, g, tex, accum, ctas_per_sm, stats, s);
}
template <>
cudaError_t launch_exact_dump<double>(const Job* job, const StepT<double>* prog, const FrameGeom& g, const float* tex, lfb_ray_hit* out,
                                   cudaStream_t s) {
  return xt::launch_dump_t<double>(job, prog, g, tex, out, s);
}

}  // namespace lfb
