// pcie_write_probe.cu -- how fast can SM stores fill page-locked host memory, against the copy engine?  A measurement tool
// behind DESIGN.md's e2e section (the tile kernel's zero-copy stores ARE the device->host transfer of a sparse frame).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/pcie_write_probe tools/pcie_write_probe.cu
//   tools/_build/pcie_write_probe            (prints one JSON line per case)
// Cases: cudaMemcpyAsync D2H of the same bytes; contiguous 16-byte stores from G CTAs; the tile pattern of a 1920 x 1080
// F64x3 frame (T dirty 16 x 16 tiles, rows of 384 contiguous bytes at a 46 080-byte pitch), rows by 16-byte lanes, for
// several grid sizes and with 1 or 4 tiles in flight per CTA.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void fill_contig(uint4* out, size_t n16, unsigned v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = make_uint4(v, v, v, v);
}

// tile q of `tiles` (tile index in a tiles_w-wide map); rows of 384 bytes = 24 lanes of 16 bytes; 256 threads cover 384 chunks
template <int INFLIGHT>
__global__ void fill_tiles(char* out, const int* tiles, int n_tiles, int tiles_w, size_t pitch, unsigned v) {
  for (int q0 = blockIdx.x * INFLIGHT; q0 < n_tiles; q0 += gridDim.x * INFLIGHT) {
#pragma unroll
    for (int k = 0; k < INFLIGHT; k++) {
      const int q = q0 + k;
      if (q >= n_tiles) break;
      const int t = tiles[q], ty = t / tiles_w, tx = t - ty * tiles_w;
      char* base = out + (size_t)ty * 16 * pitch + (size_t)tx * 384;
      for (int c = threadIdx.x; c < 384; c += blockDim.x) {
        const int row = c / 24, col = c - row * 24;
        *reinterpret_cast<uint4*>(base + (size_t)row * pitch + (size_t)col * 16) = make_uint4(v, v, v, v);
      }
    }
  }
}

// the same bytes, each warp writing one whole tile row pair (768 B) -- fewer, longer runs per warp
__global__ void fill_tiles_warp_rows(char* out, const int* tiles, int n_tiles, int tiles_w, size_t pitch, unsigned v) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int q = blockIdx.x; q < n_tiles; q += gridDim.x) {
    const int t = tiles[q], ty = t / tiles_w, tx = t - ty * tiles_w;
    char* base = out + (size_t)ty * 16 * pitch + (size_t)tx * 384;
    for (int row = warp; row < 16; row += nw)
      if (lane < 24) *reinterpret_cast<uint4*>(base + (size_t)row * pitch + (size_t)lane * 16) = make_uint4(v, v, v, v);
  }
}


// ---- interference: what does a kernel pay for running next to the host-memory writer? ----
__global__ void busy_fma(float* sink, int iters) {
  float x = threadIdx.x, y = blockIdx.x;
  for (int i = 0; i < iters; i++) { x = fmaf(x, 1.0000001f, 1e-7f); y = fmaf(y, 0.9999999f, 1e-7f); }
  if (x + y == 12345.f) sink[0] = x;
}
__global__ void busy_atomics(unsigned long long* buf, size_t n, int iters) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 97 % n;
  for (int k = 0; k < iters; k++) { atomicAdd(buf + i, 1ull); i = (i + 7919) % n; }
}
__global__ void busy_loads(const uint4* buf, size_t n, int iters, uint4* sink) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) % n;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int k = 0; k < iters; k++) { const uint4 v = __ldcg(buf + i); acc.x += v.x; acc.y ^= v.y; i = (i + 4099 + (v.x & 1)) % n; }
  if (acc.x == 0x12345678u) sink[0] = acc;
}

__global__ void fill_tiles_rep(char* out, const int* tiles, int n_tiles, int tiles_w, size_t pitch, unsigned v, int reps) {
  for (int r = 0; r < reps; r++)
    for (int q = blockIdx.x; q < n_tiles; q += gridDim.x) {
      const int t = tiles[q], ty = t / tiles_w, tx = t - ty * tiles_w;
      char* base = out + (size_t)ty * 16 * pitch + (size_t)tx * 384;
      for (int c = threadIdx.x; c < 384; c += blockDim.x) {
        const int row = c / 24, col = c - row * 24;
        asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(base + (size_t)row * pitch + (size_t)col * 16), "r"(v + r), "r"(v), "r"(v), "r"(v) : "memory");
      }
    }
}

// a PACED writer: n CTAs, each storing 32 KB per iteration and then waiting for its next time slot, so that the link runs at
// `gbps` and the queue of posted writes in front of the GPU's own PCIe READS (command fetches: reads may not pass posted
// writes) stays shallow
__global__ void fill_tiles_paced(char* out, const int* tiles, int n_tiles, int tiles_w, size_t pitch, unsigned v, float gbps) {
  const unsigned cpt = 384;  // 16-byte chunks per tile
  const size_t n_chunks = (size_t)n_tiles * cpt;
  const size_t per_iter = (size_t)gridDim.x * blockDim.x * 8;
  const float ns_per_iter = (float)(per_iter * 16) / gbps;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  int it = 0;
  for (size_t c0 = (size_t)blockIdx.x * blockDim.x * 8 + threadIdx.x; c0 < n_chunks; c0 += per_iter, it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const size_t c = c0 + (size_t)u * blockDim.x;
      if (c >= n_chunks) continue;
      const unsigned q = (unsigned)(c / cpt), r = (unsigned)(c - (size_t)q * cpt), row = r / 24, col = r - row * 24;
      const int t = tiles[q], ty = t / tiles_w, tx = t - ty * tiles_w;
      *reinterpret_cast<uint4*>(out + ((size_t)ty * 16 + row) * pitch + (size_t)tx * 384 + (size_t)col * 16) = make_uint4(v, v, v, v);
    }
    const unsigned long long due = t0 + (unsigned long long)((float)(it + 1) * ns_per_iter);
    unsigned long long now;
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); } while (now < due);
  }
}

static void interference(char* hd, char* dd, const int* d_tiles, int n_tiles, int tiles_w, size_t pitch) {
  cudaStream_t sa, sb;
  int lo, hi;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CK(cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, hi));
  CK(cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, lo));
  cudaEvent_t a0, a1, b0, b1;
  CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1)); CK(cudaEventCreate(&b0)); CK(cudaEventCreate(&b1));
  float* sink; unsigned long long* abuf; uint4* lbuf;
  const size_t na = (size_t)4 << 20, nl = (size_t)2 << 20;  // 32 MB each: L2-resident
  CK(cudaMalloc((void**)&sink, 64)); CK(cudaMalloc((void**)&abuf, na * 8)); CK(cudaMalloc((void**)&lbuf, nl * 16));
  CK(cudaMemset(abuf, 0, na * 8)); CK(cudaMemset(lbuf, 0, nl * 16));
  for (int which : {0, 1, 3}) {
    const char* name = which == 0 ? "fma" : which == 1 ? "l2_atomics" : which == 2 ? "l2_loads" : which == 3 ? "fma_one_wave_long_ctas" : which == 4 ? "fma_many_short_ctas" : "fma_4_ctas_per_sm";
    auto launch_b = [&]() {
      if (which == 0) busy_fma<<<148 * 8, 256, 0, sb>>>(sink, 6000);
      else if (which == 3) busy_fma<<<148, 256, 0, sb>>>(sink, 48000);
      else if (which == 4) busy_fma<<<148 * 8 * 16, 256, 0, sb>>>(sink, 6000 / 16);
      else if (which == 5) busy_fma<<<148 * 4, 256, 0, sb>>>(sink, 12000);
      else if (which == 1) busy_atomics<<<148 * 8, 256, 0, sb>>>(abuf, na, 40);
      else busy_loads<<<148 * 8, 256, 0, sb>>>(lbuf, nl, 60, (uint4*)sink);
    };
    for (int wc : {0, 16, -296, -1000, -2030, -2040, -2045, -2048, -2060}) {
      const bool one_writer = wc < 0;
      const bool dev_writer = wc == -1000;  // the same kernel shape writing DEVICE memory, repeated to last about as long
      const bool paced = wc <= -2000;       // 8 CTAs paced at (-wc - 2000) GB/s
      const int writer_ctas = paced ? 8 : dev_writer ? 296 : wc < 0 ? -wc : wc;
      float tb = 0, ta = 0;
      for (int rep = 0; rep < 4; rep++) {
        CK(cudaDeviceSynchronize());
        if (writer_ctas) {
          CK(cudaEventRecord(a0, sa));
          if (paced) fill_tiles_paced<<<writer_ctas, 256, 0, sa>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 3u, (float)(-wc - 2000));
          else if (dev_writer) fill_tiles_rep<<<writer_ctas, 256, 0, sa>>>(dd, d_tiles, n_tiles, tiles_w, pitch, 3u, 60);
          else if (one_writer) fill_tiles<1><<<writer_ctas, 256, 0, sa>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 3u);
          else for (int k = 0; k < 3; k++) fill_tiles<1><<<writer_ctas, 256, 0, sa>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 3u + k);
          CK(cudaEventRecord(a1, sa));
        }
        CK(cudaEventRecord(b0, sb));
        launch_b();
        CK(cudaEventRecord(b1, sb));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&tb, b0, b1));
        if (writer_ctas) CK(cudaEventElapsedTime(&ta, a0, a1));
      }
      printf("{\"case\": \"interference\", \"kernel\": \"%s\", \"writer\": \"%s\", \"writer_ctas\": %d, \"writer_kernels\": %d, \"kernel_us\": %.1f, \"writer_us\": %.1f}\n", name, paced ? (wc == -2030 ? "host paced 30" : wc == -2040 ? "host paced 40" : wc == -2045 ? "host paced 45" : wc == -2048 ? "host paced 48" : "host paced 60") : dev_writer ? "device" : "host", writer_ctas, writer_ctas ? (one_writer ? 1 : 3) : 0, tb * 1e3, ta * 1e3);
      fflush(stdout);
    }
  }
}

// paced writer, second try: small bursts (U chunks per thread), ONE polling thread per CTA that sleeps between looks
template <int U>
__global__ void fill_tiles_paced2(char* out, const int* tiles, int n_tiles, int tiles_w, size_t pitch, unsigned v, float gbps) {
  const unsigned cpt = 384;
  const size_t n_chunks = (size_t)n_tiles * cpt;
  const size_t per_iter = (size_t)gridDim.x * blockDim.x * U;
  const float ns_per_iter = (float)(per_iter * 16) / gbps;
  __shared__ unsigned long long t0s;
  if (threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); t0s = t; }
  __syncthreads();
  int it = 0;
  for (size_t c0 = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; c0 < n_chunks; c0 += per_iter, it++) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t c = c0 + (size_t)u * blockDim.x;
      if (c >= n_chunks) continue;
      const unsigned q = (unsigned)(c / cpt), r = (unsigned)(c - (size_t)q * cpt), row = r / 24, col = r - row * 24;
      const int t = tiles[q], ty = t / tiles_w, tx = t - ty * tiles_w;
      *reinterpret_cast<uint4*>(out + ((size_t)ty * 16 + row) * pitch + (size_t)tx * 384 + (size_t)col * 16) = make_uint4(v, v, v, v);
    }
    if (threadIdx.x == 0) {
      const unsigned long long due = t0s + (unsigned long long)((float)(it + 1) * ns_per_iter);
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      while (now < due) { __nanosleep(200); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); }
    }
    __syncthreads();
  }
}

// ---- how long do stream-order and cross-stream (event) dependencies take while a host writer saturates the link? ----
__global__ void tiny(int* p) { if (threadIdx.x == 0) atomicAdd(p, 1); }
static void chains(char* hd, const int* d_tiles, int n_tiles, int tiles_w, size_t pitch) {
  cudaStream_t sa, x, y;
  int lo, hi;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CK(cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, hi));
  CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&y, cudaStreamNonBlocking));
  cudaEvent_t t0, t1, w0, w1, hop[64];
  CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1)); CK(cudaEventCreate(&w0)); CK(cudaEventCreate(&w1));
  for (auto& h : hop) CK(cudaEventCreateWithFlags(&h, cudaEventDisableTiming));
  int* cnt;
  CK(cudaMalloc((void**)&cnt, 4));
  for (int writer : {0, 12, 296, -25, -35, -42, -46, -50, -1025, -1042, -1046}) {
    for (int mode = 0; mode < 3; mode++) {  // 0: 20 kernels in one stream; 1: 20 kernels ping-ponging between two streams by events; 2: kernel, 64 KB H2D copy, kernel ... x 10
      float ms = 0, wms = 0;
      static char* pinned = nullptr;
      static char* dbuf = nullptr;
      if (!pinned) { CK(cudaHostAlloc((void**)&pinned, 1 << 20, cudaHostAllocDefault)); CK(cudaMalloc((void**)&dbuf, 1 << 20)); }
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaDeviceSynchronize());
        if (writer) CK(cudaEventRecord(w0, sa));
        if (writer > 0) for (int k = 0; k < 3; k++) fill_tiles<1><<<writer, 256, 0, sa>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 3u + k);
        if (writer < 0 && writer > -1000) for (int k = 0; k < 3; k++) fill_tiles_paced2<2><<<12, 256, 0, sa>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 3u + k, (float)-writer);
        if (writer <= -1000) for (int k = 0; k < 3; k++) fill_tiles_paced2<1><<<4, 256, 0, sa>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 3u + k, (float)(-writer - 1000));
        if (writer) CK(cudaEventRecord(w1, sa));
        CK(cudaEventRecord(t0, x));
        if (mode == 0) {
          for (int k = 0; k < 20; k++) tiny<<<1, 32, 0, x>>>(cnt);
        } else if (mode == 1) {
          for (int k = 0; k < 20; k++) {
            cudaStream_t cur = (k & 1) ? y : x, nxt = (k & 1) ? x : y;
            tiny<<<1, 32, 0, cur>>>(cnt);
            CK(cudaEventRecord(hop[k], cur));
            CK(cudaStreamWaitEvent(nxt, hop[k], 0));
          }
        } else {
          for (int k = 0; k < 10; k++) {
            tiny<<<1, 32, 0, x>>>(cnt);
            CK(cudaMemcpyAsync(dbuf, pinned, 64 << 10, cudaMemcpyHostToDevice, x));
          }
        }
        CK(cudaEventRecord(t1, x));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, t0, t1));
        if (writer) { CK(cudaEventElapsedTime(&wms, w0, w1)); }
      }
      printf("{\"case\": \"chain\", \"mode\": \"%s\", \"writer\": %d, \"chain_us\": %.1f, \"three_writers_us\": %.1f}\n",
             mode == 0 ? "20 kernels, one stream" : mode == 1 ? "20 kernels, two streams, event hops" : "10 x (kernel + 64 KB H2D)", writer, ms * 1e3, wms * 1e3);
      fflush(stdout);
    }
  }
}

int main(int argc, char** argv) {
  const int W = 1920, H = 1080, tiles_w = W / 16, tiles_h = (H + 15) / 16;
  const size_t pitch = (size_t)W * 24, frame = pitch * H;
  int n_tiles = argc > 1 ? atoi(argv[1]) : 886;
  char *h = nullptr, *hd = nullptr, *d = nullptr;
  CK(cudaHostAlloc((void**)&h, frame, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void**)&hd, h, 0));
  CK(cudaMalloc((void**)&d, frame));
  CK(cudaMemset(d, 1, frame));
  // a compact blob of dirty tiles around the frame centre (what a flare looks like), whole tiles only
  std::vector<int> tiles;
  {
    const int side = (int)ceil(sqrt((double)n_tiles));
    const int x0 = (tiles_w - side) / 2, y0 = std::max(0, (tiles_h - 1 - side) / 2);
    for (int y = 0; y < side && (int)tiles.size() < n_tiles; y++)
      for (int x = 0; x < side && (int)tiles.size() < n_tiles; x++) tiles.push_back((y0 + y) * tiles_w + x0 + x);
  }
  n_tiles = (int)tiles.size();
  int* d_tiles = nullptr;
  CK(cudaMalloc((void**)&d_tiles, sizeof(int) * n_tiles));
  CK(cudaMemcpy(d_tiles, tiles.data(), sizeof(int) * n_tiles, cudaMemcpyHostToDevice));
  const size_t bytes = (size_t)n_tiles * 256 * 24;
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int reps = 50;
  auto report = [&](const char* name, int grid, int threads, float ms, size_t b) {
    printf("{\"case\": \"%s\", \"grid\": %d, \"threads\": %d, \"bytes\": %zu, \"us\": %.2f, \"GB_per_s\": %.2f}\n", name, grid, threads, b, ms * 1e3 / reps,
           (double)b * reps / (ms * 1e-3) / 1e9);
    fflush(stdout);
  };
  float ms;
  if (argc > 3) { chains(hd, d_tiles, n_tiles, tiles_w, pitch); return 0; }
  if (argc > 2) { interference(hd, d, d_tiles, n_tiles, tiles_w, pitch); return 0; }
  // copy engine
  for (size_t b : {bytes, frame}) {
    for (int k = 0; k < 3; k++) CK(cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(e0, s));
    for (int k = 0; k < reps; k++) CK(cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    report("memcpy_d2h", 0, 0, ms, b);
  }
  for (int grid : {37, 74, 148, 296, 592, 1184}) {
    for (int k = 0; k < 3; k++) fill_contig<<<grid, 256, 0, s>>>((uint4*)hd, bytes / 16, 7u);
    CK(cudaEventRecord(e0, s));
    for (int k = 0; k < reps; k++) fill_contig<<<grid, 256, 0, s>>>((uint4*)hd, bytes / 16, 7u + k);
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    report("sm_contig_16B", grid, 256, ms, bytes);
  }
  for (int grid : {148, 592}) {
    for (int k = 0; k < 3; k++) fill_contig<<<grid, 256, 0, s>>>((uint4*)hd, frame / 16, 7u);
    CK(cudaEventRecord(e0, s));
    for (int k = 0; k < 10; k++) fill_contig<<<grid, 256, 0, s>>>((uint4*)hd, frame / 16, 7u + k);
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"case\": \"sm_contig_16B_frame\", \"grid\": %d, \"bytes\": %zu, \"us\": %.2f, \"GB_per_s\": %.2f}\n", grid, frame, ms * 1e3 / 10, (double)frame * 10 / (ms * 1e-3) / 1e9);
  }
  for (int grid : {74, 148, 296, 592, 886}) {
    for (int k = 0; k < 3; k++) fill_tiles<1><<<grid, 256, 0, s>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 7u);
    CK(cudaEventRecord(e0, s));
    for (int k = 0; k < reps; k++) fill_tiles<1><<<grid, 256, 0, s>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 7u + k);
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    report("sm_tiles_1", grid, 256, ms, bytes);
  }
  for (int grid : {74, 148, 296}) {
    for (int k = 0; k < 3; k++) fill_tiles<4><<<grid, 256, 0, s>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 7u);
    CK(cudaEventRecord(e0, s));
    for (int k = 0; k < reps; k++) fill_tiles<4><<<grid, 256, 0, s>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 7u + k);
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    report("sm_tiles_4", grid, 256, ms, bytes);
  }
  for (int grid : {148, 296, 886}) {
    for (int k = 0; k < 3; k++) fill_tiles_warp_rows<<<grid, 512, 0, s>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 7u);
    CK(cudaEventRecord(e0, s));
    for (int k = 0; k < reps; k++) fill_tiles_warp_rows<<<grid, 512, 0, s>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 7u + k);
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    report("sm_tiles_warp_rows", grid, 512, ms, bytes);
  }
  // one kernel, timed alone with a sync each side (what a blocking frame sees)
  for (int k = 0; k < 5; k++) {
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    fill_tiles<1><<<296, 256, 0, s>>>(hd, d_tiles, n_tiles, tiles_w, pitch, 9u + k);
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"case\": \"sm_tiles_1_single_launch\", \"grid\": 296, \"bytes\": %zu, \"us\": %.2f, \"GB_per_s\": %.2f}\n", bytes, ms * 1e3, (double)bytes / (ms * 1e-3) / 1e9);
  }
  return 0;
}
