// sparse.cu -- the tile-sparse back end of the ghost path: fixed-point sums -> pixels for the DIRTY 16 x 16 sensor tiles only,
// on one GPU or summed across GPUs over NVLink peer memory.
//
// A flare frame is ~99 % zeros (cfg2: 24 725 non-zero pixels of 2 073 600).  The splat kernels mark every tile they deposit
// into in a byte map behind the accumulators (lfb_internal.h: AccumLayout, mark_tile: a plain store per tile); this kernel then
//   * visits the tiles that are dirty NOW (convert, store) or were dirty in the output buffer's PREVIOUS frame (store zeros),
//     so the output buffer always holds exactly this frame -- the reference's "ghost_buffer fully rewritten on return"
//     (pathtracer.cpp:719-720) without touching the other 99 %;
//   * zeroes the accumulator tiles it reads, so the next frame needs no 49.8 MB memset either;
//   * with n_ranks > 1 is the cross-GPU reduce as well: rank r owns the map words w = r (mod n) -- interleaved strips of
//     4 tiles -- ORs the ranks' marks for them, sums the dirty tiles of exactly those ranks that have them (peer loads),
//     zeroes them (peer stores) and writes the pixels into the owner's output buffer (peer stores).  Integer sums: the frame
//     has the same bits for any rank count.  This is the "reduce" of SURVEY 8e fused with finalize, moving ~1 % of the bytes
//     an all-pixels reduce moves.
// The output may be DEVICE memory or page-locked HOST memory mapped into the device (zero-copy): the stores of the dirty
// tiles then ARE the device->host transfer -- no host-side scatter, no second pass over the frame.  For a host that waits for
// each frame (lfb_render_ghosts_sparse) tiles_kernel stores into the host frame itself: 5.44 MB in 118 us at cfg2, the link's
// rate.  For frames in flight it stages the tiles in device memory and drain_kernel (below) copies them out, paced: unpaced
// stores into host memory stall everything else the GPU is asked to do while they drain.
//
// One tiles_kernel launch per frame; the CTA that finishes last rolls the tile maps over (previous = current, current = 0).
#include "lfb_internal.h"

namespace lfb {

namespace {

constexpr int kThreads = 256;  // one thread per pixel of a tile

struct TileArgs {
  PeerAccums acc;        // the ranks' accumulator buffers (pixel sums at offset 0)
  size_t bits_off;       // byte offset of the dirty-tile map inside each
  unsigned* state;       // THIS rank's tile state of the output buffer: [0] ticket, [1] tiles written, [2..3] pad, [4..] previous bits
  int rank;              // this rank owns the tile-map words w with w % acc.n == rank
  int W, H, tiles_w, n_words;
  double inv_scale;
  char* out;             // the output frame: pixel (x, y) at (x + y * W) * stride
  size_t stride;
  int elem;
  unsigned* count_out;   // optional (mapped host memory): the number of tiles this launch wrote
  char* stage;           // optional: the quadrants to write go HERE instead, unit u at u * 8 * 8 * stride, rows packed (see drain_kernel) ...
  unsigned* stage_units; // ... and [u] = the quadrant's byte offset in `out` / 16
};

__device__ __forceinline__ unsigned* bits_of(const TileArgs& A, int r) {
  return reinterpret_cast<unsigned*>(reinterpret_cast<char*>(const_cast<unsigned long long*>(A.acc.ptr[r])) + A.bits_off);
}

// Word w of every rank's tile map, loaded with all (peer) loads in flight at once; ranks >= n read as 0.
__device__ __forceinline__ void load_ranks(const TileArgs& A, int w, unsigned* words) {
#pragma unroll
  for (int r = 0; r < LFB_MAX_PEERS; r++) words[r] = r < A.acc.n ? bits_of(A, r)[w] : 0u;
}
__device__ __forceinline__ unsigned or_of_ranks(const TileArgs& A, int w) {
  unsigned words[LFB_MAX_PEERS], m = 0;
  load_ranks(A, w, words);
#pragma unroll
  for (int r = 0; r < LFB_MAX_PEERS; r++) m |= words[r];
  return m;
}

// one bit (0, 8, 16, 24) per non-zero byte of a tile-map word
__device__ __forceinline__ unsigned nonzero_bytes(unsigned m) { return __vcmpne4(m, 0u) & 0x01010101u; }

__global__ void __launch_bounds__(kThreads) tiles_kernel(TileArgs A) {
  extern __shared__ unsigned s_pre[];  // [n_words + 1] exclusive prefix of the per-word tile counts
  __shared__ __align__(16) char s_rows[kTilePx1 * kTilePx1 * 32];  // one tile's pixels as they go out, row by row (stride <= 32 B)
  __shared__ ulonglong2 s_acc[kTilePx1 * kTilePx1 * 24 / 16];  // one tile's summed accumulators (6 KB)
  __shared__ unsigned s_warp[kThreads / 32];
  __shared__ unsigned s_total;
  __shared__ bool s_last;
  __shared__ unsigned s_quads;  // which 8 x 8 quadrants of the CTA's current tile hold a non-zero pixel (bit 2 * (y / 8) + x / 8)
  __shared__ unsigned s_unit0;  // staging: the first unit of the CTA's current tile
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // Tile state: [0] ticket, [1] tiles (or staged units) of the last frame, [2] staging unit counter, [3] pad, then the PREVIOUS
  // map -- per tile, one byte: the quadrants the output frame holds non-zero pixels in -- and the NEXT map being written by
  // this launch (the last CTA rolls it over: the previous map is still being read by other CTAs until then).
  unsigned* prev = A.state + 4;
  unsigned char* next = reinterpret_cast<unsigned char*>(A.state + 4 + A.n_words);
  const int n = A.acc.n;
  const size_t row_bytes = (size_t)kTilePx1 * A.stride;  // a tile row in the output: 16 pixels, contiguous
  const bool row_vec = A.stride <= 32 && ((size_t)A.out % 16) == 0 && (((size_t)A.W * A.stride) % 16) == 0;

  // ---- 1. how many tiles does each word of the tile map hold (current of any rank, or previous)?  owned words only ----
  const int per = (A.n_words + kThreads - 1) / kThreads;  // contiguous words per thread
  const int w0 = tid * per, w1 = min(w0 + per, A.n_words);
  unsigned mine = 0;
  if (n == 1) {  // one GPU: the words of this thread's run eight at a time, their loads in flight together
    const unsigned* cur = bits_of(A, 0);
    for (int wb = w0; wb < w1; wb += 8) {
      unsigned pv[8], cv[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const bool in = wb + k < w1;
        pv[k] = in ? prev[wb + k] : 0u;
        cv[k] = in ? cur[wb + k] : 0u;
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (wb + k >= w1) break;
        const unsigned c = __popc(nonzero_bytes(pv[k] | cv[k]));
        s_pre[wb + k] = c;
        mine += c;
      }
    }
  } else {
    for (int w = w0; w < w1; w++) {
      unsigned m = 0;
      if (w % n == A.rank) m = prev[w] | or_of_ranks(A, w);
      const unsigned c = __popc(nonzero_bytes(m));
      s_pre[w] = c;
      mine += c;
    }
  }
  // block-exclusive scan of the per-thread sums
  unsigned incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    unsigned v = lane < kThreads / 32 ? s_warp[lane] : 0, inc = v;
#pragma unroll
    for (int d = 1; d < kThreads / 32; d <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += u;
    }
    if (lane < kThreads / 32) s_warp[lane] = inc - v;
    if (lane == kThreads / 32 - 1) s_total = inc;
  }
  __syncthreads();
  unsigned run = s_warp[wid] + incl - mine;
  for (int w = w0; w < w1; w++) {
    const unsigned c = s_pre[w];
    s_pre[w] = run;
    run += c;
  }
  if (tid == kThreads - 1 || (w0 < A.n_words && w1 == A.n_words)) s_pre[A.n_words] = s_total;
  __syncthreads();
  const unsigned total = s_total;

  // ---- 2. tiles q = blockIdx.x, + gridDim.x, ... in tile-map order ----
  for (unsigned q = blockIdx.x; q < total; q += gridDim.x) {
    int lo = 0, hi = A.n_words;  // the word w with s_pre[w] <= q < s_pre[w + 1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_pre[mid] <= q) lo = mid; else hi = mid;
    }
    const int w = lo;
    unsigned words[LFB_MAX_PEERS];  // every rank's map word, all loads in flight at once (peer loads: ~2 us each over NVLink)
    load_ranks(A, w, words);
    unsigned have = 0;  // bit r set = rank r's tile is dirty
    unsigned m = prev[w];
#pragma unroll
    for (int r = 0; r < LFB_MAX_PEERS; r++) m |= words[r];
    const int byte = __fns(nonzero_bytes(m), 0, (int)(q - s_pre[w]) + 1) >> 3;  // the (q - s_pre[w])-th marked tile of the word
#pragma unroll
    for (int r = 0; r < LFB_MAX_PEERS; r++) have |= (((words[r] >> (8 * byte)) & 0xffu) ? 1u : 0u) << r;
    const int t = w * 4 + byte;
    const int ty = t / A.tiles_w, tx = t - ty * A.tiles_w;
    const int lx = tid & (kTilePx1 - 1), ly = tid >> kTilePxLog2;
    const int x = tx * kTilePx1 + lx, y = ty * kTilePx1 + ly;
    const bool inside = x < A.W && y < A.H;
    const size_t p = (size_t)x + (size_t)y * A.W;
    unsigned long long s0 = 0, s1 = 0, s2 = 0;
    const bool whole = (tx + 1) * kTilePx1 <= A.W && (ty + 1) * kTilePx1 <= A.H;
    if (whole && (A.W & 1) == 0) {
      // a whole tile of an even-width frame: its 16 rows are 384 contiguous, 16-byte aligned bytes each in every rank's
      // accumulators -> read (and zero) them as 16-byte chunks, 24 per row: full sectors over NVLink / out of L2.  Thread t
      // owns chunks t and t + 256 for every rank, so the ranks' sums build up in registers.
      constexpr int kChunks = kTilePx1 * (kTilePx1 * 24 / 16);  // 384
      ulonglong2 c0 = make_ulonglong2(0ull, 0ull), c1 = c0;
      const size_t row0 = 3 * ((size_t)tx * kTilePx1 + (size_t)ty * kTilePx1 * A.W);  // u64 index of the tile's first value
      for (int r = 0; r < n; r++) {
        if (!((have >> r) & 1u)) continue;
        unsigned long long* base = const_cast<unsigned long long*>(A.acc.ptr[r]) + row0;
        {
          const int row = tid / 24, col = tid - row * 24;
          ulonglong2* src = reinterpret_cast<ulonglong2*>(base + (size_t)row * 3 * A.W) + col;
          const ulonglong2 v = *src;
          *src = make_ulonglong2(0ull, 0ull);  // the next frame finds the accumulators clear
          c0.x += v.x; c0.y += v.y;
        }
        if (tid + kThreads < kChunks) {
          const int c = tid + kThreads, row = c / 24, col = c - row * 24;
          ulonglong2* src = reinterpret_cast<ulonglong2*>(base + (size_t)row * 3 * A.W) + col;
          const ulonglong2 v = *src;
          *src = make_ulonglong2(0ull, 0ull);
          c1.x += v.x; c1.y += v.y;
        }
      }
      s_acc[tid] = c0;
      if (tid + kThreads < kChunks) s_acc[tid + kThreads] = c1;
      __syncthreads();
      const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(s_acc) + (ly * kTilePx1 + lx) * 3;
      s0 = mine[0]; s1 = mine[1]; s2 = mine[2];
      __syncthreads();  // s_acc is reused by the CTA's next tile
    } else if (inside) {
      for (int r = 0; r < n; r++) {
        if (!((have >> r) & 1u)) continue;
        unsigned long long* a = const_cast<unsigned long long*>(A.acc.ptr[r]) + 3 * p;
        s0 += a[0]; s1 += a[1]; s2 += a[2];
        a[0] = 0ull; a[1] = 0ull; a[2] = 0ull;  // the next frame finds the accumulators clear
      }
    }
    const double v0 = (double)(long long)s0 * A.inv_scale, v1 = (double)(long long)s1 * A.inv_scale, v2 = (double)(long long)s2 * A.inv_scale;
    // Quadrant sparsity: a dirty 16 x 16 tile is mostly zeros too (cfg2: 24 719 non-zero pixels lie in 128 512 pixels of dirty
    // tiles, but in 76 288 pixels of dirty 8 x 8 quadrants).  Only the quadrants that hold a non-zero pixel now, or held one in
    // the output's previous frame (the tile's byte in the previous map), are written: -41 % of the bytes that cross PCIe.
    // A warp is two tile rows: lanes 0-7 / 16-23 are the left quadrant, 8-15 / 24-31 the right one; warps 0-3 the upper half.
    __syncthreads();  // (every thread is done with the previous tile's s_quads / s_unit0)
    if (tid == 0) s_quads = 0u;
    __syncthreads();
    {
      const unsigned bal = __ballot_sync(0xffffffffu, inside && (s0 | s1 | s2) != 0ull);
      if (lane == 0 && bal) atomicOr(&s_quads, (((bal & 0x00ff00ffu) ? 1u : 0u) | ((bal & 0xff00ff00u) ? 2u : 0u)) << (2 * (wid >> 2)));
    }
    __syncthreads();
    const unsigned q_now = s_quads;
    const unsigned q_prev = (prev[w] >> (8 * byte)) & 0xffu;
    // A full tile whose rows are 16-byte aligned in the output goes out as 16-byte chunks, row by row (16 pixels x stride bytes
    // are contiguous): whole 128-byte lines per warp -- what a PCIe (zero-copy host frame) or NVLink (peer frame) write wants;
    // 24-byte pixels stored 8 bytes per lane would go out as partial sectors (measured: 17 GB/s into host memory).
    const bool full = whole && row_vec;
    // full tiles go out by quadrant; the others (frame edge, unaligned rows) whole, pixel by pixel -- their byte says "all four"
    const unsigned q_write = full ? (q_now | (q_prev & 0xfu)) : 0xfu;
    if (tid == 0) {
      next[t] = (unsigned char)(full ? q_now : (q_now ? 0xfu : 0u));
      if (A.stage) s_unit0 = full && q_write ? atomicAdd(A.state + 2, (unsigned)__popc(q_write)) : 0u;
    }
    if (full) {
      char* mine = s_rows + (size_t)ly * row_bytes + (size_t)lx * A.stride;
      if (A.elem == LFB_F32x3) {
        float* o = reinterpret_cast<float*>(mine);
        o[0] = (float)v0; o[1] = (float)v1; o[2] = (float)v2;
        for (size_t q = 3; q < A.stride / 4; q++) o[q] = 0.f;  // padding lanes read back as zero
      } else {
        double* o = reinterpret_cast<double*>(mine);
        o[0] = v0; o[1] = v1; o[2] = v2;
        for (size_t q = 3; q < A.stride / 8; q++) o[q] = 0.0;
      }
      __syncthreads();
      const int chunks_per_row = (int)(row_bytes / 16), half = chunks_per_row / 2;  // (16 px * stride / 16 = stride: even)
      const size_t tile_off = ((size_t)tx * kTilePx1 + (size_t)ty * kTilePx1 * A.W) * A.stride;
      const size_t pitch = (size_t)A.W * A.stride;
      const unsigned unit0 = s_unit0;
      if (A.stage && tid < 4 && ((q_write >> tid) & 1u))  // where each staged quadrant goes in the output, in 16-byte chunks
        A.stage_units[unit0 + __popc(q_write & ((1u << tid) - 1u))] =
            (unsigned)((tile_off + (size_t)(tid >> 1) * 8 * pitch + (size_t)(tid & 1) * half * 16) >> 4);
      for (int c = tid; c < kTilePx1 * chunks_per_row; c += kThreads) {
        const int row = c / chunks_per_row, col = c - row * chunks_per_row;
        const int quad = 2 * (row >> 3) + (col >= half ? 1 : 0);
        if (!((q_write >> quad) & 1u)) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(s_rows + (size_t)row * row_bytes + (size_t)col * 16);
        if (A.stage) {  // unit = one quadrant: 8 rows of `half` chunks, packed
          const unsigned u = unit0 + __popc(q_write & ((1u << quad) - 1u));
          reinterpret_cast<uint4*>(A.stage)[(size_t)u * 8 * half + (size_t)(row & 7) * half + (col >= half ? col - half : col)] = v;
        } else {
          *reinterpret_cast<uint4*>(A.out + tile_off + (size_t)row * pitch + (size_t)col * 16) = v;
        }
      }
      __syncthreads();  // s_rows (and s_quads / s_unit0) are reused by the CTA's next tile
    } else if (inside) {
      if (A.elem == LFB_F32x3) {
        float* o = reinterpret_cast<float*>(A.out + p * A.stride);
        o[0] = (float)v0; o[1] = (float)v1; o[2] = (float)v2;
      } else {
        double* o = reinterpret_cast<double*>(A.out + p * A.stride);
        o[0] = v0; o[1] = v1; o[2] = v2;
      }
    }
  }

  // ---- 3. the CTA that finishes last rolls the tile maps over ----
  // (no CTA-wide fence before the ticket: what the last CTA must not overtake are the other CTAs' READS of the maps -- completed,
  // their values were consumed above -- and their bytes of the NEXT map, which thread 0 wrote and fences itself; the pixel
  // stores need no ordering against the roll, and the kernel's end publishes them.  A fence by all 256 threads here was 26 %
  // of the kernel's stall samples.)
  __syncthreads();
  if (tid == 0) {
    __threadfence();  // this thread's bytes of the next map, before the ticket
    s_last = atomicAdd(A.state, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  unsigned* next_w = A.state + 4 + A.n_words;
  if (n == 1) {
    unsigned* cur = bits_of(A, 0);
    for (int wb = tid; wb < A.n_words; wb += 8 * kThreads) {
      unsigned nv[8];  // eight words per thread, their loads in flight together
#pragma unroll
      for (int k = 0; k < 8; k++) nv[k] = wb + k * kThreads < A.n_words ? __ldcg(next_w + wb + k * kThreads) : 0u;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int w = wb + k * kThreads;
        if (w >= A.n_words) break;
        cur[w] = 0u;
        prev[w] = nv[k];
        next_w[w] = 0u;
      }
    }
  } else {
    for (int w = tid; w < A.n_words; w += kThreads) {
      if (w % n != A.rank) continue;
      for (int r = 0; r < n; r++) bits_of(A, r)[w] = 0u;
      prev[w] = __ldcg(next_w + w);
      next_w[w] = 0u;
    }
  }
  if (tid == 0) {
    A.state[0] = 0u;
    A.state[1] = A.stage ? __ldcg(A.state + 2) : total;  // what drain_kernel copies: the staged units
    A.state[2] = 0u;
    if (A.count_out) *A.count_out = total;
  }
}

// The second stage of a frame that goes to HOST memory while other work runs: a FEW CTAs copy the staged tiles into the
// caller's page-locked frame, PACED just below the link rate.  Measured on B200 (tools/pcie_write_probe.cu):
//  * stores into host memory are posted PCIe writes; SMs issue them far faster than the link drains them (~50 GB/s), and the
//    backlog sits in front of every PCIe READ the GPU makes -- reads may not pass posted writes -- which is how the GPU fetches
//    its command stream: 20 back-to-back tiny kernels take 80 us alone and 315 us next to an unpaced writer (they complete when
//    it does), i.e. the next frame's kernels cannot even be launched while this frame drains;
//  * an SM whose store path is backed up is also all but lost to the other CTAs on it.
// A few CTAs storing one 16-byte chunk per thread and round, each round released by the clock just below the link rate, keep
// the backlog at a microsecond or two: in the probe the same 20 kernels then take 110 us and the writer still gets 43 of the
// 48 GB/s it gets unpaced; in the frame pipeline (tools/e2e_probe.py, cfg2, sun moving every frame) 0.146 ms per frame at 48 GB/s
// against 0.19 unpaced (and 0.17 at 52 GB/s, 0.19 at 32: the optimum sits right under the link rate) and 0.27 for one blocking
// call per frame.
constexpr int kDrainDepth = 8;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void __launch_bounds__(kThreads) drain_kernel(const char* __restrict__ stage, const unsigned* __restrict__ stage_units,
                                                         const unsigned* __restrict__ state, char* __restrict__ out, int W, size_t stride,
                                                         float gbps) {
  __shared__ unsigned long long s_t0;
  const unsigned total = state[1];  // staged units (8 x 8 quadrants) of this frame (tiles_kernel's last CTA)
  const unsigned half = (unsigned)stride / 2u, cpu = 8u * half;  // 16-byte chunks per quadrant row (8 px * stride / 16) / per unit
  const unsigned n_chunks = total * cpu;                          // < 2^32: launch_tiles checks
  const uint4* src = reinterpret_cast<const uint4*>(stage);
  const unsigned per_round = gridDim.x * kThreads;  // chunks per round: one per thread (16 CTAs: 64 KB)
  const float ns_per_round = gbps > 0.f ? (float)per_round * 16.f / gbps : 0.f;
  const bool paced = ns_per_round > 0.f;
  const unsigned pitch16 = (unsigned)(((size_t)W * stride) >> 4);  // output row pitch in chunks
  if (threadIdx.x == 0) s_t0 = global_ns();
  __syncthreads();
  const unsigned lane0 = blockIdx.x * kThreads + threadIdx.x;
  // this thread's chunk as (unit q, row, col), advanced by one round per step without divisions
  unsigned q = lane0 / cpu, row = (lane0 - q * cpu) / half, col = lane0 - q * cpu - row * half;
  const unsigned dq = per_round / cpu, drow = (per_round - dq * cpu) / half, dcol = per_round - dq * cpu - drow * half;
  uint4 cur[kDrainDepth], nxt[kDrainDepth];
#pragma unroll
  for (int d = 0; d < kDrainDepth; d++) {
    const unsigned c = lane0 + (unsigned)d * per_round;
    cur[d] = c < n_chunks ? __ldcg(src + c) : make_uint4(0u, 0u, 0u, 0u);
  }
  unsigned round = 0;
  for (unsigned g = 0; g < n_chunks; g += per_round * kDrainDepth) {  // uniform over the CTA: it synchronises every round
#pragma unroll
    for (int d = 0; d < kDrainDepth; d++) {  // the next group's loads are in flight while this group drains
      const unsigned c = g + lane0 + (unsigned)(kDrainDepth + d) * per_round;
      nxt[d] = c < n_chunks ? __ldcg(src + c) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int d = 0; d < kDrainDepth; d++) {
      const unsigned c = g + lane0 + (unsigned)d * per_round;
      if (c < n_chunks) {
        const unsigned off16 = __ldca(stage_units + q);  // the quadrant's first chunk in the output
        reinterpret_cast<uint4*>(out)[(size_t)off16 + (size_t)row * pitch16 + col] = cur[d];
      }
      col += dcol; row += drow; q += dq;
      if (col >= half) { col -= half; row++; }
      if (row >= 8u) { row -= 8u; q++; }
      if (paced) {
        round++;
        if (threadIdx.x == 0) {
          const unsigned long long due = s_t0 + (unsigned long long)((float)round * ns_per_round);
          while (global_ns() < due) {}  // one spinning thread per CTA; __nanosleep overshoots by more than a round
        }
        __syncthreads();
      }
    }
#pragma unroll
    for (int d = 0; d < kDrainDepth; d++) cur[d] = nxt[d];
  }
}

// calibration: how fast do unpaced stores from a few CTAs fill page-locked host memory?
__global__ void __launch_bounds__(kThreads) host_fill_kernel(uint4* out, size_t n16, unsigned v) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n16; i += (size_t)gridDim.x * kThreads) out[i] = make_uint4(v, v, v, v);
}

}  // namespace

// accum_ptrs[r]: rank r's accumulator buffer as mapped in this process (n_ranks = 1: this GPU's own); state: this rank's
// tile state of `out` (tile_state_bytes, zero-initialised once while `out` is clear).
// stage (optional, lfb_internal.h: tile_stage_bytes): whole tiles are staged there instead of stored into `out`; launch_tile_drain
// copies them out -- for a host-memory `out` written while other kernels run.
cudaError_t launch_tiles(const PeerAccums& P, int rank, int W, int H, double inv_scale, void* out, size_t stride, int elem,
                         unsigned* state, unsigned* count_out, int ctas, cudaStream_t s, void* stage) {
  const AccumLayout lay = accum_layout(W, H);
  TileArgs A;
  A.acc = P;
  A.bits_off = lay.bits_off;
  A.state = state;
  A.rank = rank;
  A.W = W; A.H = H; A.tiles_w = lay.tiles_w; A.n_words = lay.n_words;
  A.inv_scale = inv_scale;
  A.out = (char*)out; A.stride = stride; A.elem = elem;
  A.count_out = count_out;
  A.stage = (char*)stage;
  A.stage_units = stage ? reinterpret_cast<unsigned*>((char*)stage + (size_t)lay.n_tiles * kTilePx1 * kTilePx1 * stride) : nullptr;
  const size_t smem = sizeof(unsigned) * ((size_t)lay.n_words + 1);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;  // > 1.6 M tiles: use the dense finalize
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
  }
  if (ctas < 1) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    ctas = P.n > 1 ? sms : 2 * sms;  // cross-GPU: its CTAs mostly wait for NVLink; fewer of them leave the co-running trace more of the SMs (N = 8: 0.1026 vs 0.1076 ms / step; issuing all ranks' tile loads of a chunk before consuming any -- 80 registers -- was slower there: 0.1177)
  }
  if (ctas > lay.n_tiles) ctas = lay.n_tiles;
  tiles_kernel<<<ctas, kThreads, smem, s>>>(A);
  if (stage && (((size_t)W * H * stride) >> 4) >= 0xffffffffull) return cudaErrorInvalidConfiguration;  // chunk offsets are 32-bit
  return cudaGetLastError();
}

// The staged tiles of launch_tiles(..., stage) -> `out` (page-locked host memory), by `ctas` CTAs paced at gbps (<= 0: unpaced).
// Any stream, ordered after that launch_tiles.
cudaError_t launch_tile_drain(int W, int H, void* out, size_t stride, const unsigned* state, const void* stage, int ctas, float gbps,
                              cudaStream_t s) {
  const AccumLayout lay = accum_layout(W, H);
  const char* st = (const char*)stage;
  const unsigned* stage_units = reinterpret_cast<const unsigned*>(st + (size_t)lay.n_tiles * kTilePx1 * kTilePx1 * stride);
  drain_kernel<<<ctas > 0 ? ctas : 16, kThreads, 0, s>>>(st, stage_units, state, (char*)out, W, stride, gbps);
  return cudaGetLastError();
}

}  // namespace lfb

namespace lfb {
// GB/s that unpaced SM stores reach into page-locked host memory on this device's link (best of 3 fills of 8 MB).
cudaError_t measure_host_write_gbps(cudaStream_t s, float* gbps) {
  const size_t bytes = (size_t)8 << 20;
  char *h = nullptr, *hd = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t err = cudaHostAlloc((void**)&h, bytes, cudaHostAllocMapped);
  if (err != cudaSuccess) return err;
  if ((err = cudaHostGetDevicePointer((void**)&hd, h, 0)) != cudaSuccess) { cudaFreeHost(h); return err; }
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 0.f;
  for (int rep = 0; rep < 4 && err == cudaSuccess; rep++) {
    cudaEventRecord(e0, s);
    host_fill_kernel<<<64, kThreads, 0, s>>>(reinterpret_cast<uint4*>(hd), bytes / 16, (unsigned)rep);  // (16 CTAs reach only ~35 of the link's ~50 GB/s)
    cudaEventRecord(e1, s);
    err = cudaStreamSynchronize(s);
    float ms = 0.f;
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms > 0.f) best = fmaxf(best, (float)bytes / (ms * 1e6f));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFreeHost(h);
  *gbps = best;
  return err;
}
}  // namespace lfb
