// ghost_grid_f64.cu -- FP64 instantiation of the ray-grid kernels (the parity path), the
// paraxial set-up kernel and the accumulator finalize kernel.  Compiled with --fmad=false:
// the CPU oracle (and the x86-64 reference) round every product and sum separately, and the
// parity tests compare fixed-point sums bit for bit.
#define LFB_TU f64
#include "ghost_grid_impl.cuh"
#include "ref_abcd.cuh"

namespace lfb {

cudaError_t upload_lens_f64(const DevLens& h, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(f64::c_lens, &h, sizeof(DevLens), 0, cudaMemcpyHostToDevice, s);
}
cudaError_t launch_trace_splat_f64(const Job* jobs, int n_jobs, const FrameGeom& g, int mode, const float* tex,
                                   unsigned long long* accum, cudaStream_t s) {
  return f64::launch_trace_splat_t<double>(jobs, n_jobs, g, mode, tex, accum, s);
}
cudaError_t launch_trace_dump_f64(const Job* job, const FrameGeom& g, int mode, const float* tex, lfb_ray_hit* out,
                                  cudaStream_t s) {
  return f64::launch_trace_dump_t<double>(job, g, mode, tex, out, s);
}

// ---------------------------------------------------------------------------
// paraxial set-up (restates lfo_paraxial_system / pathtracer.cpp:588-689)
// ---------------------------------------------------------------------------
__global__ void paraxial_setup_kernel(Job* __restrict__ jobs, int n_jobs, int physical_backward) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_jobs) return;
  const DevLens& L = f64::c_lens;
  const int n = L.n_surfaces, stop = L.stop;
  const int i = jobs[q].i, j = jobs[q].j, lam = jobs[q].lambda;
  int nc = 0;
  m2 M = make2(1.f, 0.f, 0.f, 1.f);
  auto record = [&](const m2& X) {
    jobs[q].cross[nc][0] = X.a; jobs[q].cross[nc][1] = X.b;
    jobs[q].cross[nc][2] = X.c; jobs[q].cross[nc][3] = X.d;
    nc++;
  };
  if (i < 0) {
    for (int k = 0; k < n; k++) {
      if (k == stop) { record(M); M = mmul(mT(L.d[k]), M); continue; }
      M = step_TR(L, lam, k, M);
    }
  } else {
    for (int k = 0; k < j; k++) {
      if (k == stop) { record(M); M = mmul(mT(L.d[k]), M); continue; }
      M = step_TR(L, lam, k, M);
    }
    M = mmul(mL(L.c[j]), M);
    for (int k = j - 1; k > i; k--) {
      M = step_back(L, lam, k, M, physical_backward);
      if (k == stop) record(M);
    }
    M = step_second_reflection(L, i, M);
    for (int k = i + 1; k < n; k++) {
      if (k == stop) { record(M); M = mmul(mT(L.d[k]), M); continue; }
      M = step_TR(L, lam, k, M);
    }
  }
  jobs[q].n_cross = nc;
  jobs[q].full[0] = M.a; jobs[q].full[1] = M.b; jobs[q].full[2] = M.c; jobs[q].full[3] = M.d;
}

cudaError_t launch_paraxial_setup(Job* jobs, int n_jobs, int physical_backward, cudaStream_t s) {
  if (n_jobs <= 0) return cudaSuccess;
  paraxial_setup_kernel<<<(n_jobs + 63) / 64, 64, 0, s>>>(jobs, n_jobs, physical_backward);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// finalize: u64 fixed point -> pixels.  One thread per pixel; the accumulators are read
// once (24 B/px) and the output written once (12 or 24 B/px): purely HBM-bound.
// ---------------------------------------------------------------------------
// CLEAR: the accumulators are zeroed as they are read, so a pipelined host needs no separate 24 B/px memset per frame.
template <bool CLEAR>
__global__ void __launch_bounds__(256) finalize_kernel(unsigned long long* __restrict__ accum, size_t npx,
                                                       double inv_scale, char* __restrict__ out, size_t stride,
                                                       int elem, int additive) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a[3] = {0ull, 0ull, 0ull};
  if (p < npx) {
#pragma unroll
    for (int c = 0; c < 3; c++) a[c] = CLEAR ? accum[3 * p + c] : __ldg(accum + 3 * p + c);  // read-only path only when nothing is written
  }
  if (CLEAR) {
    // the CTA's 256 pixels are one contiguous, 16-byte aligned 6 KB range (256 * 24 B): once every thread has read its
    // pixel, zero the range with coalesced 16-byte stores instead of three 8-byte stores at a 24-byte stride per thread
    __syncthreads();
    const size_t first = (size_t)blockIdx.x * blockDim.x * 3;                 // first u64 of this CTA's range (even)
    const size_t last = (first + (size_t)blockDim.x * 3 < npx * 3) ? first + (size_t)blockDim.x * 3 : npx * 3;
    ulonglong2* z = reinterpret_cast<ulonglong2*>(accum + first);
    const size_t pairs = (last - first) / 2;
    for (size_t q = threadIdx.x; q < pairs; q += blockDim.x) z[q] = make_ulonglong2(0ull, 0ull);
    if (threadIdx.x == 0 && ((last - first) & 1)) accum[last - 1] = 0ull;
  }
  if (p >= npx) return;
  double v[3];
#pragma unroll
  for (int c = 0; c < 3; c++) v[c] = (double)(long long)a[c] * inv_scale;
  if (elem == LFB_F32x3) {
    float* o = reinterpret_cast<float*>(out + p * stride);
    if (additive) { o[0] += (float)v[0]; o[1] += (float)v[1]; o[2] += (float)v[2]; }
    else { o[0] = (float)v[0]; o[1] = (float)v[1]; o[2] = (float)v[2]; }
  } else {
    double* o = reinterpret_cast<double*>(out + p * stride);
    if (additive) { o[0] += v[0]; o[1] += v[1]; o[2] += v[2]; }
    else { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
  }
}

__global__ void __launch_bounds__(256) finalize_rect_kernel(const unsigned long long* __restrict__ accum, int W, int x0, int y0,
                                                            int rw, int rh, double inv_scale, char* __restrict__ out,
                                                            size_t stride, int elem) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)rw * rh) return;
  const int ry = (int)(q / rw), rx = (int)(q - (size_t)ry * rw);
  const size_t p = (size_t)(x0 + rx) + (size_t)(y0 + ry) * W;
  double v[3];
#pragma unroll
  for (int c = 0; c < 3; c++) v[c] = (double)(long long)accum[3 * p + c] * inv_scale;
  if (elem == LFB_F32x3) {
    float* o = reinterpret_cast<float*>(out + q * stride);
    o[0] = (float)v[0]; o[1] = (float)v[1]; o[2] = (float)v[2];
  } else {
    double* o = reinterpret_cast<double*>(out + q * stride);
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
  }
}

cudaError_t launch_finalize_rect(const unsigned long long* accum, int W, const int rect[4], double inv_scale, void* out,
                                 size_t stride, int elem, cudaStream_t s) {
  const int rw = rect[2] - rect[0] + 1, rh = rect[3] - rect[1] + 1;
  if (rw <= 0 || rh <= 0) return cudaSuccess;
  const size_t n = (size_t)rw * rh;
  finalize_rect_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(accum, W, rect[0], rect[1], rw, rh, inv_scale, (char*)out, stride, elem);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// fused cross-GPU reduce + finalize over NVLink peer memory.  Every rank runs this kernel on ITS slice of the frame's
// pixels [p0, p1): it sums the ranks' u64 accumulators -- plain loads from the peers' mapped buffers, or ONE
// multimem.ld_reduce per value when the buffers are bound to an NVSwitch multicast object (the switch adds) -- converts
// to pixels and stores them straight into the owner rank's output buffer (a peer store for every other rank).
// Integer sums: the result is bit-identical to a single-GPU frame for any rank count.  Replaces NCCL reduce + finalize.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long multimem_add_u64(const unsigned long long* mc) {
  unsigned long long v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(mc) : "memory");
  return v;
}

// Device-side barrier across the ranks of one node, for stream-ordered use: rank r publishes `epoch` into slot r of every
// peer's flag array (system-scope release stores over NVLink) and then waits until all slots of ITS OWN array reached the
// epoch.  Everything a rank enqueued before the barrier (its trace kernel's sums) is complete and visible to the peers'
// kernels enqueued after it.  One warp; ranks run on different GPUs, so the spin cannot starve the peers it waits for.
__global__ void peer_barrier_kernel(PeerFlags F, int rank, unsigned long long epoch) {
  const int r = threadIdx.x;
  if (r < F.n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(F.ptr[r] + rank), "l"(epoch) : "memory");
    unsigned long long seen = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(F.ptr[F.rank_self] + r) : "memory");
    } while (seen < epoch);
  }
  __syncwarp();
}

cudaError_t launch_peer_barrier(const PeerFlags& F, int rank, unsigned long long epoch, cudaStream_t s) {
  peer_barrier_kernel<<<1, 32, 0, s>>>(F, rank, epoch);
  return cudaGetLastError();
}

// The accumulators are treated as a flat u64 array: each thread owns 4 consecutive values (two 16-byte loads per rank,
// all issued before any is used, so a warp keeps 32 x 32 B x n_ranks in flight across NVLink).
__device__ __forceinline__ void store_value(char* out, size_t e, double v, size_t stride, int elem, bool packed) {
  if (packed) {
    if (elem == LFB_F32x3) reinterpret_cast<float*>(out)[e] = (float)v;
    else reinterpret_cast<double*>(out)[e] = v;
  } else {
    const size_t p = e / 3;
    const int c = (int)(e - 3 * p);
    if (elem == LFB_F32x3) reinterpret_cast<float*>(out + p * stride)[c] = (float)v;
    else reinterpret_cast<double*>(out + p * stride)[c] = v;
  }
}

// Persistent form: a fixed grid (2 CTAs per SM) strides over the slice, so the kernel co-runs with the next frame's trace
// kernel instead of queueing thousands of CTAs behind it; each thread keeps UNROLL x 32 B per rank in flight.
template <int UNROLL>
__global__ void __launch_bounds__(256) reduce_finalize_kernel(PeerAccums P, const unsigned long long* __restrict__ mc, size_t e0, size_t e1,
                                                              double inv_scale, char* __restrict__ out, size_t stride, int elem) {
  const bool packed = stride == (elem == LFB_F32x3 ? 12u : 24u);
  const size_t chunk = 4;  // values per thread per load group
  const size_t n_groups = (e1 - e0) / chunk;  // full groups; the tail is handled by thread 0 of block 0
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
  for (size_t g0 = tid; g0 < n_groups; g0 += nthreads * UNROLL) {
    unsigned long long sum[UNROLL][4];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) { sum[u][0] = sum[u][1] = sum[u][2] = sum[u][3] = 0; }
    if (mc) {
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const size_t g = g0 + (size_t)u * nthreads;
        if (g < n_groups) {
#pragma unroll
          for (int k = 0; k < 4; k++) sum[u][k] = multimem_add_u64(mc + e0 + g * chunk + k);
        }
      }
    } else {
      for (int r = 0; r < P.n; r++) {
        ulonglong2 v[UNROLL][2];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
          const size_t g = g0 + (size_t)u * nthreads;
          if (g < n_groups) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(P.ptr[r] + e0 + g * chunk);
            v[u][0] = src[0]; v[u][1] = src[1];
          } else {
            v[u][0] = v[u][1] = make_ulonglong2(0, 0);
          }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) { sum[u][0] += v[u][0].x; sum[u][1] += v[u][0].y; sum[u][2] += v[u][1].x; sum[u][3] += v[u][1].y; }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t g = g0 + (size_t)u * nthreads;
      if (g < n_groups) {
#pragma unroll
        for (int k = 0; k < 4; k++) store_value(out, e0 + g * chunk + k, (double)(long long)sum[u][k] * inv_scale, stride, elem, packed);
      }
    }
  }
  if (tid == 0) {
    for (size_t q = e0 + n_groups * chunk; q < e1; q++) {
      unsigned long long t = 0;
      if (mc) t = multimem_add_u64(mc + q);
      else for (int r = 0; r < P.n; r++) t += P.ptr[r][q];
      store_value(out, q, (double)(long long)t * inv_scale, stride, elem, packed);
    }
  }
}

cudaError_t launch_reduce_finalize(const PeerAccums& P, const unsigned long long* mc, size_t p0, size_t p1, double inv_scale,
                                   void* out, size_t stride, int elem, int ctas, cudaStream_t s) {
  if (p1 <= p0) return cudaSuccess;
  size_t e0 = 3 * p0, e1 = 3 * p1;
  if (e0 & 1) {  // keep the 16-byte loads aligned: the first (odd) value goes through the scalar tail of a 1-thread launch
    reduce_finalize_kernel<1><<<1, 1, 0, s>>>(P, mc, e0, e0 + 1, inv_scale, (char*)out, stride, elem);
    e0 += 1;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // few, fat CTAs: the kernel is bound by NVLink latency, not by SM resources, and must leave the SMs to the trace
  // kernel it overlaps with (lfb_options.reduce_ctas overrides the CTA count for tuning)
  if (ctas < 1) ctas = sms;
  if (e1 > e0) reduce_finalize_kernel<8><<<ctas, 256, 0, s>>>(P, mc, e0, e1, inv_scale, (char*)out, stride, elem);
  return cudaGetLastError();
}

cudaError_t launch_finalize(const unsigned long long* accum, int W, int H, double inv_scale, void* out,
                            size_t stride, int elem, int additive, cudaStream_t s) {
  size_t npx = (size_t)W * H;
  finalize_kernel<false><<<(unsigned)((npx + 255) / 256), 256, 0, s>>>(const_cast<unsigned long long*>(accum), npx, inv_scale, (char*)out,
                                                                       stride, elem, additive);
  return cudaGetLastError();
}

cudaError_t launch_finalize_clear(unsigned long long* accum, int W, int H, double inv_scale, void* out, size_t stride, int elem,
                                  cudaStream_t s) {
  size_t npx = (size_t)W * H;
  finalize_kernel<true><<<(unsigned)((npx + 255) / 256), 256, 0, s>>>(accum, npx, inv_scale, (char*)out, stride, elem, 0);
  return cudaGetLastError();
}

}  // namespace lfb
