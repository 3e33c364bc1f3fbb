// peaks.cu -- live roofline denominators for the scalar pipes the ghost trace is bound by.
// MEASURED_PEAKS.json (driver-written) holds HBM and tensor peaks only; the ray trace is
// FP32-FMA / MUFU bound, so the bench measures those two pipes on the same GPU, under the
// same clocks, right beside the number it normalises:
//   fma_peak_kernel   8 independent FFMA chains per thread (register-only)
//   mufu_peak_kernel  8 independent MUFU.RSQ chains per thread
// The FMA kernel's block 0 also reports the SM cycles (clock64) and the nanoseconds (globaltimer) its own threads took:
// their ratio is the SM clock under load.  (Dividing block 0's cycles by the WHOLE kernel's duration, as round 1 did, is
// wrong: the first wave's blocks finish long before the kernel does.)
#include "lfb_internal.h"

namespace lfb {

namespace {
constexpr int kIters = 4096, kChains = 8;

__global__ void __launch_bounds__(256) fma_peak_kernel(float* __restrict__ sink, long long* __restrict__ cycles, float a, float b) {
  float x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) x[c] = (float)(threadIdx.x + c);
  // the chains depend on t0 and t1 depends on the chains, so the compiler cannot move work across the clock reads
  long long t0, t1;
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0));
  x[0] += (float)(t0 >> 63);
#pragma unroll 16
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int c = 0; c < kChains; c++) x[c] = fmaf(x[c], a, b);
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < kChains; c++) s += x[c];
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "f"(s) : "memory");
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1) : "l"(t1) : "memory");
  if (s == 12345.678f) sink[0] = s;  // keeps the chains alive; never true in practice
  if (blockIdx.x == 0 && threadIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = (long long)(g1 - g0); }
}

__global__ void __launch_bounds__(256) mufu_peak_kernel(float* __restrict__ sink, long long* __restrict__ cycles) {
  float x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) x[c] = 1.5f + (float)(threadIdx.x + c);
  long long t0, t1;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0));
  x[0] += (float)(t0 >> 63);
#pragma unroll 16
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int c = 0; c < kChains; c++) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < kChains; c++) s += x[c];
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "f"(s) : "memory");
  if (s == 12345.678f) sink[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}
}  // namespace

// Returns FP32 FLOP/s (FMA = 2), MUFU op/s and the SM clock (Hz) seen while the FMA kernel ran.
cudaError_t probe_peaks(int device, cudaStream_t s, double* fp32_flops, double* mufu_ops, double* sm_clock_hz) {
  cudaDeviceProp prop;
  cudaError_t err = cudaGetDeviceProperties(&prop, device);
  if (err != cudaSuccess) return err;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  float* sink = nullptr;
  long long* cyc = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if ((err = cudaMalloc((void**)&sink, sizeof(float))) != cudaSuccess) return err;
  if ((err = cudaMalloc((void**)&cyc, 2 * sizeof(long long))) != cudaSuccess) { cudaFree(sink); return err; }
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double ops = (double)blocks * threads * kIters * kChains;
  double best_fma = 0, best_mufu = 0, clock_hz = 0;
  for (int rep = 0; rep < 5; rep++) {
    float ms = 0;
    long long h_cyc[2] = {0, 0};
    cudaEventRecord(e0, s);
    fma_peak_kernel<<<blocks, threads, 0, s>>>(sink, cyc, 1.0000001f, 1e-7f);
    cudaEventRecord(e1, s);
    if ((err = cudaStreamSynchronize(s)) != cudaSuccess) break;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h_cyc, cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost);
    if (rep > 0 && 2 * ops / (ms * 1e-3) > best_fma) {
      best_fma = 2 * ops / (ms * 1e-3);
      if (h_cyc[1] > 0) clock_hz = (double)h_cyc[0] / ((double)h_cyc[1] * 1e-9);  // block 0: its own cycles over its own nanoseconds
    }
    cudaEventRecord(e0, s);
    mufu_peak_kernel<<<blocks, threads, 0, s>>>(sink, cyc);
    cudaEventRecord(e1, s);
    if ((err = cudaStreamSynchronize(s)) != cudaSuccess) break;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best_mufu = best_mufu > ops / (ms * 1e-3) ? best_mufu : ops / (ms * 1e-3);
  }
  if (err == cudaSuccess) err = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink); cudaFree(cyc);
  if (fp32_flops) *fp32_flops = best_fma;
  if (mufu_ops) *mufu_ops = best_mufu;
  if (sm_clock_hz) *sm_clock_hz = clock_hz;
  return err;
}

}  // namespace lfb
