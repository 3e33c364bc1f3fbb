// starburst.cu -- the diffraction starburst, sm_100a, FP64.
//
// Reference: PathTracer::raytrace_starburst (src/pathtracer/pathtracer.cpp:947-1004) evaluates, PER PIXEL, a brute-force
// 2-D DFT of the aperture mask over its bounding box (~127 k texels, two sincos each: 8.9 ms per pixel, ~5 h per 1080p
// frame on one core), then a radial suppression / amplification, a power law and a 1/r^1.5 falloff (:1030-1052).
//
// The phase of texel (xc, yc) at pixel (x, y) is 2 pi [u (lr - x') + v (ud - y')] with u = xc/W_t - 1/2, v = yc/W_t - 1/2
// (sic: both over the width, :963-964), x' = convertCoordinate(x) (:936-945), (lr, ud) from compute_phase (:918-934):
// it SEPARATES, so the whole frame is two complex matrix products
//     G[yc, x] = sum_xc A[yc, xc] E1[xc, x]          E1 = exp(j 2 pi u_xc (lr - x'(x)))          (bh x bw) . (bw x W)
//     F[y,  x] = sum_yc E2[y, yc] G[yc, x]           E2 = exp(j 2 pi v_yc (ud - y'(y)))          (H x bh) . (bh x W)
// -- 6.6 GFLOP for a 1080p frame instead of 2.6e11 sincos pairs -- followed by a per-pixel epilogue kernel.  FP64
// throughout (the reference is double; the parity tolerance is 1e-9 relative on the DFT scalar).
//
// PERIODICITY.  alpha(x) = lr - x'(x) and beta(y) = ud - y'(y) are always INTEGERS (the W/2 and H/2 cancel), and
// u_xc alpha = (xc - W_t/2) alpha / W_t, so E1 only depends on alpha mod P with P = W_t (W_t even) or 2 W_t; likewise E2 on
// beta mod P.  When the frame is wider / taller than P (1080p vs the 500-texel masks: P = 500) the products are evaluated
// on the P-periodic lattice only -- 7x fewer FLOPs at 1080p -- and every pixel looks its |F| up at
// (beta(y) mod P, alpha(x) mod P); with the reduced arguments the twiddles are also more accurate than the reference's.
//
// HERMITIAN SYMMETRY.  The mask is real and u_xc P, v_yc P are integers, so on the lattice E1[xc, P - c] = conj E1[xc, c],
// G[yc, P - c] = conj G[yc, c] and F[P - r, P - c] = conj F[r, c]: the first product computes columns 0..P/2 (two FMAs per
// term, mirrored on store), the second rows 0..P/2, and a pixel in the other half reads |F| at (P - r, P - c).
//
//   twiddle_kernel            E1 / E2 / complex copy of the mask's bounding box
//   zgemm_kernel<...>         tiled complex FP64 GEMM (32x32 or 64x64 tile per CTA, 4x4 outputs per thread, K step 16,
//                             K-split thread groups for the small lattice products)
//                             EPILOGUE: stores |F| only
//   star_pixels_kernel        |F| / total -> suppression (dist > W_t/2: factor^8) / amplification (dist <= radius:
//                             I^(dist/radius)) -> I^(3 - flare_intensity) -> x sum of radiance + falloff -> caller layout
#include "lfb_internal.h"

namespace lfb {
namespace {

constexpr int TK = 16;

__device__ __forceinline__ double convert_coordinate(int pixel, int length, bool is_y) {  // :936-945
  const double cc = is_y ? -((double)(float)pixel) + ((double)(float)length / 2.0) : ((double)(float)pixel) - ((double)(float)length / 2.0);
  return cc >= 0 ? cc : (double)length + cc;
}

// alpha of lattice column c / beta of lattice row r: the class index itself on the periodic lattice, else the pixel's value
__device__ __forceinline__ double star_alpha(const StarFrame& f, int c) {
  return f.lattice_x ? (double)c : f.lr - convert_coordinate(c, f.W, false);
}
__device__ __forceinline__ double star_beta(const StarFrame& f, int r) {
  return f.lattice_y ? (double)r : f.ud - convert_coordinate(r, f.H, true);
}
// lattice column of pixel x / row of pixel y (the class index in [0, P); the offsets are integers well inside int range)
__device__ __forceinline__ int star_col(const StarFrame& f, int x) {
  if (!f.lattice_x) return x;
  const int m = __double2int_rn(f.lr - convert_coordinate(x, f.W, false)) % f.period;
  return m < 0 ? m + f.period : m;
}
__device__ __forceinline__ int star_row(const StarFrame& f, int y) {
  if (!f.lattice_y) return y;
  const int m = __double2int_rn(f.ud - convert_coordinate(y, f.H, true)) % f.period;
  return m < 0 ? m + f.period : m;
}

// E1[xc_i][col]  (rows = bw, cols = n_col);  E2[row][yc_i]  (rows = n_row, cols = bh);
// Ac = complex copy of the mask's bounding box A[yc_i][xc_i] (rows = bh, cols = bw)
__global__ void __launch_bounds__(256) twiddle_kernel(StarFrame f, const float* __restrict__ tex, double2* __restrict__ E1,
                                                      double2* __restrict__ E2, double2* __restrict__ Ac) {
  // one launch for the three tables: [0, n0) -> E1, [n0, n0 + n1) -> E2, the rest -> Ac
  const size_t n0 = (size_t)f.bw * f.n_col, n1 = (size_t)f.n_row * f.bh, n2 = (size_t)f.bh * f.bw;
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n0 + n1 + n2) return;
  const int which = q < n0 ? 0 : (q < n0 + n1 ? 1 : 2);
  if (which == 1) q -= n0;
  if (which == 2) q -= n0 + n1;
  const unsigned cols = which == 0 ? f.n_col : (which == 1 ? f.bh : f.bw);
  const int r = (int)((unsigned)q / cols), c = (int)((unsigned)q - (unsigned)r * cols);  // every table is far below 2^32 entries
  double2 v;
  if (which == 2) {
    v.x = (double)tex[(size_t)(f.by0 + r) * f.tw + (f.bx0 + c)];
    v.y = 0.0;
  } else {
    double e;
    if (which == 0) e = (((double)(f.bx0 + r) / (double)f.tw) - 0.5) * star_alpha(f, c);
    else e = (((double)(f.by0 + c) / (double)f.tw) - 0.5) * star_beta(f, r);
    sincospi(2.0 * e, &v.y, &v.x);
  }
  (which == 0 ? E1 : (which == 1 ? E2 : Ac))[q] = v;
}

struct StarEpilogue {
  char* out;
  size_t stride;
  int elem, additive, n_lights;
  const double* lights;  // per light: fo_x * W, fo_y * H, r, g, b
  double rad_sum[3];
};

// 1/sqrt(x) for a normal positive x: the hardware's FP64 seed (MUFU.RSQ64H, ~2^-22 relative) + one Newton step (1.5 e^2:
// < 1e-13 relative) with none of rsqrt()'s special-case branches.  Used by the falloff term only, whose 4x4 stratified
// quadrature already differs from the reference's 16 random samples by ~1e-2; the oracle pin on it is 1e-11 absolute.
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y * fma(-0.5 * x, y * y, 1.5);
}

__device__ __forceinline__ void star_pixel(const StarFrame& f, const StarEpilogue& E, int x, int y, double mag) {
  double I = mag / f.total;
  const double dx = f.org_x - (double)x, dy = f.org_y - (double)y, dist = sqrt(dx * dx + dy * dy);
  if (dist > (double)f.tw / 2.0) {  // suppression :983-989
    const double factor = ((double)f.tw / 2.0) / dist;
    const double f2 = factor * factor, f4 = f2 * f2;
    I = (f4 * f4) * I;  // pow(factor, 8.0)
  } else if (dist <= f.flare_radius) {  // amplification :990-996
    I = pow(I, dist / f.flare_radius);
  }
  const double s = f.exponent == 2.0 ? I * I : (f.exponent == 1.0 ? I : pow(I, f.exponent));  // the usual -i values, exactly
  // calculate_irradiance_falloff :1030-1052; the reference's 16 random samples of the pixel -> its 4x4 stratified midpoints
  double fall[3] = {0.0, 0.0, 0.0};
  for (int l = 0; l < E.n_lights; l++) {
    const double* L = E.lights + 5 * l;
    double ox2[4], oy2[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double ox = L[0] - ((double)x + (k + 0.5) / 4), oy = L[1] - ((double)y + (k + 0.5) / 4);
      ox2[k] = ox * ox; oy2[k] = oy * oy;
    }
    double acc = 0.0;
#pragma unroll
    for (int sy = 0; sy < 4; sy++)
#pragma unroll
      for (int sx = 0; sx < 4; sx++) {
        // r^-1.5 with r = 1 + max(0, |o| - 5), through two branch-free reciprocal square roots instead of sqrt, sqrt and
        // a division: this loop is the kernel's FP64 budget (16 samples x lights per pixel)
        const double d2 = (ox2[sx] + oy2[sy]) + 1e-30;
        const double r = 1.0 + fmax(0.0, d2 * rsqrt_pos(d2) - 5.0);
        const double q = rsqrt_pos(r);
        acc = fma(q * q, q, acc);
      }
    acc *= (1.0 / 16.0);
    fall[0] += L[2] * acc; fall[1] += L[3] * acc; fall[2] += L[4] * acc;
  }
  const double v0 = s * E.rad_sum[0] + fall[0], v1 = s * E.rad_sum[1] + fall[1], v2 = s * E.rad_sum[2] + fall[2];
  char* o = E.out + ((size_t)x + (size_t)y * f.W) * E.stride;
  if (E.elem == LFB_F32x3) {
    float* p = reinterpret_cast<float*>(o);
    if (E.additive) { p[0] += (float)v0; p[1] += (float)v1; p[2] += (float)v2; }
    else { p[0] = (float)v0; p[1] = (float)v1; p[2] = (float)v2; }
  } else {
    double* p = reinterpret_cast<double*>(o);
    if (E.additive) { p[0] += v0; p[1] += v1; p[2] += v2; }
    else { p[0] = v0; p[1] = v1; p[2] = v2; }
  }
}

// C[M x N] = A[M x K] . B[K x N], complex FP64, row-major.  EPILOGUE: only |C| is stored (as a double array in C).
// BM x BN tile per CTA, 4 x 4 outputs per thread, K step 16.  The lattice products are small (250..1000 per side, K ~ 350):
//   * 32 x 32 tiles give 130..260 CTAs for the 148 SMs where 64 x 64 gave 48..64;
//   * KG thread groups share each staged K tile and take every KG-th k of it, so a CTA runs KG x 2 warps instead of 2 and its
//     serial FMA chain is KG times shorter (one warp per scheduler cannot hide the LDS -> DFMA latency); the groups' partial
//     tiles are summed through shared memory in a fixed order at the end (deterministic);
//   * REAL_A: the mask is real, so the first product needs two FMAs per term instead of four.
template <int BM, int BN, int KG, bool REAL_A, bool EPILOGUE>
__global__ void __launch_bounds__((BM / 4) * (BN / 4) * KG, KG > 1 ? 2 : 1)
zgemm_kernel(const double2* __restrict__ A, const double2* __restrict__ B, double2* __restrict__ C, int M, int N, int K, int ldb, int ldc,
             int mirror_p) {
  constexpr int NT0 = (BM / 4) * (BN / 4), NT = NT0 * KG, SM_ROWS = BM / 4, SN_COLS = BN / 4;
  constexpr int A_DOUBLES = TK * (BM + 1) * (REAL_A ? 1 : 2), B_DOUBLES = TK * BN * 2, RED_DOUBLES = KG > 1 ? NT0 * 32 : 0;
  constexpr int SMEM_DOUBLES = (A_DOUBLES + B_DOUBLES) > RED_DOUBLES ? (A_DOUBLES + B_DOUBLES) : RED_DOUBLES;
  __shared__ __align__(16) double smem[SMEM_DOUBLES];
  double2 (*sB)[BN] = reinterpret_cast<double2 (*)[BN]>(smem);                    // [TK][BN]
  double (*sAr)[BM + 1] = reinterpret_cast<double (*)[BM + 1]>(smem + B_DOUBLES);  // transposed: [TK][BM + 1], real parts
  double (*sAi)[BM + 1] = sAr + TK;                                                // imaginary parts (absent for a real A)
  const int grp = threadIdx.x / NT0, lt = threadIdx.x % NT0;
  const int tx = lt % SN_COLS, ty = lt / SN_COLS;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  double2 acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = make_double2(0.0, 0.0);
  // the next K tile travels global -> registers while the current one is multiplied (register double buffering)
  constexpr int A_PER = (BM * TK + NT - 1) / NT, B_PER = (TK * BN + NT - 1) / NT;
  double2 pa[A_PER], pb[B_PER];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < A_PER; q++) {  // A tile: BM x 16, 16 consecutive k per row
      const int e = threadIdx.x + q * NT, am = e / TK, ak = e % TK;
      const int gm = m0 + am, gk = k0 + ak;
      pa[q] = (e < BM * TK && gm < M && gk < K) ? A[(size_t)gm * K + gk] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int q = 0; q < B_PER; q++) {  // B tile: 16 x BN
      const int e = threadIdx.x + q * NT, bk = e / BN, bn = e % BN;
      const int gk = k0 + bk, gn = n0 + bn;
      pb[q] = (e < TK * BN && gk < K && gn < N) ? B[(size_t)gk * ldb + gn] : make_double2(0.0, 0.0);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int q = 0; q < A_PER; q++) {
      const int e = threadIdx.x + q * NT;
      if (e < BM * TK) {
        sAr[e % TK][e / TK] = pa[q].x;
        if (!REAL_A) sAi[e % TK][e / TK] = pa[q].y;
      }
    }
#pragma unroll
    for (int q = 0; q < B_PER; q++) {
      const int e = threadIdx.x + q * NT;
      if (e < TK * BN) sB[e / BN][e % BN] = pb[q];
    }
    if (k0 + TK < K) fetch(k0 + TK);
    __syncthreads();
#pragma unroll
    for (int k = grp; k < TK; k += KG) {
      double ar[4], ai[4];
      double2 b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        ar[i] = sAr[k][ty + SM_ROWS * i];
        ai[i] = REAL_A ? 0.0 : sAi[k][ty + SM_ROWS * i];
      }
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = sB[k][tx + SN_COLS * j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (REAL_A) {
            acc[i][j].x = fma(ar[i], b[j].x, acc[i][j].x);
            acc[i][j].y = fma(ar[i], b[j].y, acc[i][j].y);
          } else {
            acc[i][j].x = fma(ar[i], b[j].x, fma(-ai[i], b[j].y, acc[i][j].x));
            acc[i][j].y = fma(ar[i], b[j].y, fma(ai[i], b[j].x, acc[i][j].y));
          }
        }
    }
    __syncthreads();
  }
  if (KG > 1) {  // group 0 += group 1, 2, ... (fixed order), through the (now idle) tile storage
    double2* red = reinterpret_cast<double2*>(smem);
    for (int g = 1; g < KG; g++) {
      if (grp == g) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) red[(i * 4 + j) * NT0 + lt] = acc[i][j];
      }
      __syncthreads();
      if (grp == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const double2 v = red[(i * 4 + j) * NT0 + lt];
            acc[i][j].x += v.x; acc[i][j].y += v.y;
          }
      }
      __syncthreads();
    }
    if (grp != 0) return;
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gm = m0 + ty + SM_ROWS * i, gn = n0 + tx + SN_COLS * j;
      if (gm >= M || gn >= N) continue;
      if (EPILOGUE) {
        reinterpret_cast<double*>(C)[(size_t)gm * ldc + gn] = sqrt(acc[i][j].x * acc[i][j].x + acc[i][j].y * acc[i][j].y);
      } else {
        C[(size_t)gm * ldc + gn] = acc[i][j];
        // real A, lattice columns: C[., P - c] = conj C[., c]; only columns 0..P/2 were computed
        if (mirror_p && gn > 0 && 2 * gn < mirror_p) C[(size_t)gm * ldc + (mirror_p - gn)] = make_double2(acc[i][j].x, -acc[i][j].y);
      }
    }
}

}  // namespace

// one thread per pixel: |F| from the lattice, then suppression / amplification / power law / falloff / output
__global__ void __launch_bounds__(256) star_pixels_kernel(StarFrame f, StarEpilogue E, const double* __restrict__ mag) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;  // one row per blockIdx.y: no 64-bit divisions
  if (x >= f.W) return;
  int row = star_row(f, y), col = star_col(f, x);
  if (f.herm && row > f.period / 2) {  // |F(r, c)| = |F(P - r, P - c)|
    row = f.period - row;
    col = col ? f.period - col : 0;
  }
  star_pixel(f, E, x, y, mag[(size_t)row * f.n_col + col]);
}

// HDRImageBuffer::toColor (util/image.h:208-223: gamma 2.2, exposure sqrt(2), clamp) + ImageBuffer::update_pixel
// (:53-62: truncating 8-bit pack 0xFFBBGGRR); flip = 1 also applies save_image's vertical flip (raytraced_renderer.cpp:739-742).
__global__ void __launch_bounds__(256) to_color_kernel(const double* __restrict__ hdr, int W, int H, uint32_t* __restrict__ out, int flip) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (size_t)W * H) return;
  const float one_over_gamma = 1.0f / 2.2f;
  const float exposure = (float)sqrt(pow(2.0, (double)1.0f));
  uint32_t px = 0xFF000000u;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double s = hdr[3 * p + k] * (double)exposure;
    double v;
    if (s == 0.0) v = 0.0;       // pow(+-0, y > 0) = +0: the empty part of a flare frame skips the FP64 pow
    else if (s >= 1.0) v = 1.0;  // pow(s >= 1, y > 0) >= 1, clamped below anyway
    else {
      v = pow(s, (double)one_over_gamma);
      v = (1.0 < v) ? 1.0 : v;  // std::min(pow, 1.0): NaN passes
      v = (0.0 < v) ? v : 0.0;  // std::max(0.0, .): NaN -> 0
    }
    const float c = (float)v;
    px |= ((uint32_t)(fminf(fmaxf(c, 0.f), 1.f) * 255.f)) << (8 * k);
  }
  const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
  out[(size_t)x + (size_t)(flip ? H - 1 - y : y) * W] = px;
}

cudaError_t launch_to_color(const double* hdr, int W, int H, uint32_t* out, int flip, cudaStream_t s) {
  const size_t n = (size_t)W * H;
  to_color_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(hdr, W, H, out, flip);
  return cudaGetLastError();
}

// scratch: E1 (bw*n_col) | E2 (n_row*bh) | Ac (bh*bw) | G (bh*n_col) complex doubles | |F| (n_row*n_col) doubles
size_t starburst_scratch_bytes(const StarFrame& f) {
  return sizeof(double2) * ((size_t)f.bw * f.n_col + (size_t)f.n_row * f.bh + (size_t)f.bh * f.bw + (size_t)f.bh * f.n_col) +
         sizeof(double) * (size_t)f.n_row * f.n_col;
}

cudaError_t launch_starburst(const StarFrame& f, const float* tex, void* scratch, const double* lights_dev, int n_lights,
                             const double rad_sum[3], void* out, size_t stride, int elem, int additive, bool spectrum_cached, cudaStream_t s,
                             int* launches) {
  double2* E1 = (double2*)scratch;
  double2* E2 = E1 + (size_t)f.bw * f.n_col;
  double2* Ac = E2 + (size_t)f.n_row * f.bh;
  double2* G = Ac + (size_t)f.bh * f.bw;
  double* mag = (double*)(G + (size_t)f.bh * f.n_col);
  auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
  StarEpilogue E;
  E.out = (char*)out; E.stride = stride; E.elem = elem; E.additive = additive; E.n_lights = n_lights; E.lights = lights_dev;
  E.rad_sum[0] = rad_sum[0]; E.rad_sum[1] = rad_sum[1]; E.rad_sum[2] = rad_sum[2];
  // On the periodic lattice of both axes the spectrum |F| is a property of the MASK alone (the light and the frame only
  // decide which lattice class a pixel reads): the engine keeps it between frames and only the pixel kernel runs.
  if (!spectrum_cached) {
  twiddle_kernel<<<blocks((size_t)f.bw * f.n_col + (size_t)f.n_row * f.bh + (size_t)f.bh * f.bw), 256, 0, s>>>(f, tex, E1, E2, Ac);
  // tile size by problem size: 64 x 64 tiles once they fill the GPU twice over, 32 x 32 below that
  auto big = [](int m, int n) { return (size_t)((m + 63) / 64) * ((n + 63) / 64) >= 2 * 148; };
  {  // G = Ac . E1   (bh x bw) . (bw x n_col); Ac is real.  On the column lattice G[., P - c] = conj G[., c]: half the columns
    const int n_half = f.lattice_x ? f.period / 2 + 1 : f.n_col, mirror = f.lattice_x ? f.period : 0;
    if (big(f.bh, n_half)) {
      dim3 grid((n_half + 63) / 64, (f.bh + 63) / 64);
      zgemm_kernel<64, 64, 1, true, false><<<grid, 256, 0, s>>>(Ac, E1, G, f.bh, n_half, f.bw, f.n_col, f.n_col, mirror);
    } else {
      dim3 grid((n_half + 31) / 32, (f.bh + 31) / 32);
      zgemm_kernel<32, 32, 4, true, false><<<grid, 256, 0, s>>>(Ac, E1, G, f.bh, n_half, f.bw, f.n_col, f.n_col, mirror);
    }
  }
  {  // |F| = |E2 . G|   (n_row x bh) . (bh x n_col)
    if (big(f.n_row, f.n_col)) {
      dim3 grid((f.n_col + 63) / 64, (f.n_row + 63) / 64);
      zgemm_kernel<64, 64, 1, false, true><<<grid, 256, 0, s>>>(E2, G, (double2*)mag, f.n_row, f.n_col, f.bh, f.n_col, f.n_col, 0);
    } else {
      dim3 grid((f.n_col + 31) / 32, (f.n_row + 31) / 32);
      zgemm_kernel<32, 32, 4, false, true><<<grid, 256, 0, s>>>(E2, G, (double2*)mag, f.n_row, f.n_col, f.bh, f.n_col, f.n_col, 0);
    }
  }
  if (launches) *launches += 3;
  }
  star_pixels_kernel<<<dim3((unsigned)((f.W + 255) / 256), (unsigned)f.H), 256, 0, s>>>(f, E, mag);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace lfb
