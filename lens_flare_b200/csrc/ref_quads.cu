// ref_quads.cu -- REF_QUADS mode: PathTracer::generate_ghost_buffer on the device,
// bit-for-bit (compiled with --fmad=false; see the Makefile).
//
//   ref_setup_kernel    one thread per ghost (pair x colour): the two marginal rays through
//                       trace_ray_auto_before/after (pathtracer.cpp:588-689), then draw_ghost
//                       (:433-508) -> two y-sorted, half-pixel-shifted triangles with their
//                       loop bounds (:346-393).
//   ref_raster_kernel   GATHER formulation of rasterize_textured_triangle/fill_textured_pixel
//                       (:305-410): one thread owns one sensor pixel and walks the triangles
//                       in the reference's draw order, accumulating in double exactly like
//                       HDRImageBuffer::update_pixel_additive (util/image.h:145-147).  Same
//                       operations in the same order => the same bits, with no atomics, and
//                       the reference's dominant cost (zero-filling the buffer twice,
//                       :719-720) is fused into the single output write.
//
// The only libm calls of this path -- atan/cosf/sinf of frame constants (:414) -- are made
// once per frame on the host (RefFrame), with the same libm the reference links.
#include "lfb_internal.h"
#include "ref_abcd.cuh"

namespace lfb {

__constant__ DevLens c_lens_ref;

cudaError_t upload_lens_ref(const DevLens& h, cudaStream_t s) {
  return cudaMemcpyToSymbolAsync(c_lens_ref, &h, sizeof(DevLens), 0, cudaMemcpyHostToDevice, s);
}

namespace {

// the stop re-aim of pathtracer.cpp:618-629 / :656-667
__device__ void reaim(const DevLens& L, const m2& M, float r, float theta, double& ray_r, double& ray_t) {
  double ap = mrow0(M, ray_r, ray_t);
  if (ap > L.h_stop || ap < -L.h_stop) {
    float r_a = (float)L.h_stop;
    if (r < 0.f) r_a = (float)(-L.h_stop_neg);
    float r_e = (float)(((double)r_a - M.b * (double)theta) / M.a);
    ray_r = (double)r_e;
    ray_t = (double)theta;
  }
}

// trace_ray_auto_before (which = 0) / trace_ray_auto_after (which = 1); returns the sensor height
__device__ double trace_ray_auto(const DevLens& L, int lam, int which, float r, float theta, int i, int j) {
  const int n = L.n_surfaces, stop = L.stop;
  double ray_r = (double)r, ray_t = (double)theta;
  m2 M = make2(1.f, 0.f, 0.f, 1.f);
  if (which == 0) {
    for (int k = 0; k < j; k++) M = step_TR(L, lam, k, M);
    M = mmul(mL(L.c[j]), M);
    for (int k = j - 1; k > i; k--) M = step_back(L, lam, k, M, 0);
    M = step_second_reflection(L, i, M);
    for (int k = i + 1; k < n; k++) {
      if (k == stop) { reaim(L, M, r, theta, ray_r, ray_t); M = mmul(mT(L.d[k]), M); continue; }
      M = step_TR(L, lam, k, M);
    }
  } else {
    for (int k = 0; k < j; k++) {
      if (k == stop) { reaim(L, M, r, theta, ray_r, ray_t); M = mmul(mT(L.d[k]), M); continue; }
      M = step_TR(L, lam, k, M);
    }
    M = mmul(mL(L.c[j]), M);
    for (int k = j - 1; k > i; k--) M = step_back(L, lam, k, M, 0);
    M = step_second_reflection(L, i, M);
    for (int k = i + 1; k < n; k++) M = step_TR(L, lam, k, M);
  }
  return mrow0(M, ray_r, ray_t);
}

struct m3 { double m[3][3]; };

__device__ m3 m3mul(const m3& A, const m3& B) {  // CGL/src/matrix3x3.cpp:99-114 ordering
  m3 C;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++)
      C.m[r][c] = (B.m[0][c] * A.m[r][0] + B.m[1][c] * A.m[r][1]) + B.m[2][c] * A.m[r][2];
  return C;
}

// shift_vertex :412-430 (cs/sn = cosf/sinf(float(atan(...))) from the host)
__device__ void shift_vertex(float x, float y, float scale, float shift_amount, float cs, float sn, double out[2]) {
  m3 scaling = {{{(double)scale, 0, 0}, {0, (double)scale, 0}, {0, 0, 1}}};
  m3 rotation = {{{(double)cs, (double)(-sn), 0}, {(double)sn, (double)cs, 0}, {0, 0, 1}}};
  m3 shift = {{{1, 0, (double)(shift_amount * cs)}, {0, 1, (double)(shift_amount * sn)}, {0, 0, 1}}};
  m3 sr = m3mul(shift, rotation);
  m3 srs = m3mul(sr, scaling);
  double vx = (double)x, vy = (double)y, vz = 1.0;
  out[0] = (vx * srs.m[0][0] + vy * srs.m[0][1]) + vz * srs.m[0][2];
  out[1] = (vx * srs.m[1][0] + vy * srs.m[1][1]) + vz * srs.m[1][2];
}

__device__ __forceinline__ void swapf(float& a, float& b) { float t = a; a = b; b = t; }

// the prologue of rasterize_textured_triangle :350-393
__device__ void make_tri(RefTri& T, int W, int H, float x0, float y0, float u0, float v0, float x1, float y1,
                         float u1, float v1, float x2, float y2, float u2, float v2, const double col[3]) {
  if (y1 < y0) { swapf(x0, x1); swapf(y0, y1); swapf(u0, u1); swapf(v0, v1); }
  if (y2 < y0) { swapf(x0, x2); swapf(y0, y2); swapf(u0, u2); swapf(v0, v2); }
  if (y2 < y1) { swapf(x1, x2); swapf(y1, y2); swapf(u1, u2); swapf(v1, v2); }
  x0 = (float)((double)x0 - 0.5); y0 = (float)((double)y0 - 0.5);
  x1 = (float)((double)x1 - 0.5); y1 = (float)((double)y1 - 0.5);
  x2 = (float)((double)x2 - 0.5); y2 = (float)((double)y2 - 0.5);
  int lo, hi;
  lo = (int)floorf(fminf(fminf(x0, x1), x2)); float min_x = (float)(lo > 0 ? lo : 0);
  hi = (int)ceilf(fmaxf(fmaxf(x0, x1), x2));  float max_x = (float)(hi < W - 1 ? hi : W - 1);
  lo = (int)floorf(y0);                       float min_y = (float)(lo > 0 ? lo : 0);
  hi = (int)ceilf(y2);                        float max_y = (float)(hi < H - 1 ? hi : H - 1);
  T.x0 = x0; T.y0 = y0; T.u0 = u0; T.v0 = v0;
  T.x1 = x1; T.y1 = y1; T.u1 = u1; T.v1 = v1;
  T.x2 = x2; T.y2 = y2; T.u2 = u2; T.v2 = v2;
  // for (int y = min_y; y < max_y; y++): int start, float upper bound
  T.min_x = (int)min_x; T.min_y = (int)min_y;
  T.max_x = (int)ceilf(max_x); T.max_y = (int)ceilf(max_y);  // y < max (float) <=> y < ceil(max) for ints
  T.col[0] = col[0]; T.col[1] = col[1]; T.col[2] = col[2];
}

__global__ void ref_setup_kernel(RefFrame f, const int* __restrict__ pairs, const float* __restrict__ rgb_weight,
                                 RefTri* __restrict__ tris, lfb_ref_ghost* __restrict__ ghosts, int* bbox) {
  const DevLens& L = c_lens_ref;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= f.n_pairs * f.n_lambda) return;
  const int p = g / f.n_lambda, c = g - p * f.n_lambda;  // colour is the inner loop, :737-746
  const int i = pairs[2 * p], j = pairs[2 * p + 1];
  const int which = i > L.stop ? 1 : 0;
  const float P = (float)L.P;
  const double s1 = trace_ray_auto(L, c, which, P, f.angle_to_sun, i, j);
  const double s2 = trace_ray_auto(L, c, which, -P, f.angle_to_sun, i, j);
  // draw_ghost(color, float r1, float r2) :433-508
  const float r1 = (float)s1, r2 = (float)s2;
  const float shift_amt = (float)((double)(-(r1 + r2) / 2.f) * 0.4);
  const float scale_amt = (float)((double)fabsf(r2 - r1) * 0.2);
  double ul[2], ll[2], ur[2], lr[2];
  shift_vertex(-1.f, 1.f, scale_amt, shift_amt, f.cs, f.sn, ul);
  shift_vertex(-1.f, -1.f, scale_amt, shift_amt, f.cs, f.sn, ll);
  shift_vertex(1.f, 1.f, scale_amt, shift_amt, f.cs, f.sn, ur);
  shift_vertex(1.f, -1.f, scale_amt, shift_amt, f.cs, f.sn, lr);
  const float intensity_scalar = 10.f;
  const float size_scalar = 1.f / (scale_amt * scale_amt);
  const float kk = intensity_scalar * size_scalar;
  double col[3] = {(double)rgb_weight[3 * c] * (double)kk, (double)rgb_weight[3 * c + 1] * (double)kk,
                   (double)rgb_weight[3 * c + 2] * (double)kk};
  const float tw = (float)f.tex_w, th = (float)f.tex_h;
  const float ulx = (float)(f.gb_mid_w + ul[0]), uly = (float)(f.gb_mid_h + ul[1]);
  const float llx = (float)(f.gb_mid_w + ll[0]), lly = (float)(f.gb_mid_h + ll[1]);
  const float urx = (float)(f.gb_mid_w + ur[0]), ury = (float)(f.gb_mid_h + ur[1]);
  const float lrx = (float)(f.gb_mid_w + lr[0]), lry = (float)(f.gb_mid_h + lr[1]);
  make_tri(tris[2 * g], f.W, f.H, ulx, uly, 0.f, 0.f, llx, lly, 0.f, th, urx, ury, tw, 0.f, col);
  make_tri(tris[2 * g + 1], f.W, f.H, lrx, lry, 0.f, 0.f, llx, lly, 0.f, th, urx, ury, tw, 0.f, col);  // sic :498
  for (int t = 0; t < 2; t++) {  // the pixels the reference's bbox loops visit (upper bounds exclusive)
    const RefTri& T = tris[2 * g + t];
    if (T.max_x > T.min_x && T.max_y > T.min_y) grow_bbox(bbox, T.min_x, T.min_y, T.max_x - 1, T.max_y - 1);
  }
  lfb_ref_ghost& G = ghosts[g];
  G.i = i; G.j = j; G.colour = c; G.pad = 0; G.r1 = s1; G.r2 = s2;
  G.verts[0][0] = ulx; G.verts[0][1] = uly; G.verts[1][0] = llx; G.verts[1][1] = lly;
  G.verts[2][0] = urx; G.verts[2][1] = ury; G.verts[3][0] = lrx; G.verts[3][1] = lry;
  G.scale = scale_amt; G.shift = shift_amt;
}

constexpr int kTileW = 32, kTileH = 8, kMaxTris = 2 * 64 * LFB_MAX_LAMBDA;

// fill_textured_pixel :305-343 for one triangle at pixel (x, y): returns texel (0 when outside)
__device__ __forceinline__ bool fill_sample(const RefTri& T, int x, int y, const float* __restrict__ tex, int tw,
                                            int th, float& sample) {
  const float fx = (float)x, fy = (float)y;
  float xy_to_01 = -(T.y1 - T.y0) * (fx - T.x0) + (T.x1 - T.x0) * (fy - T.y0);
  float two_to_01 = -(T.y1 - T.y0) * (T.x2 - T.x0) + (T.x1 - T.x0) * (T.y2 - T.y0);
  float alpha = xy_to_01 / two_to_01;
  float xy_to_12 = -(T.y2 - T.y1) * (fx - T.x1) + (T.x2 - T.x1) * (fy - T.y1);
  float zero_to_12 = -(T.y2 - T.y1) * (T.x0 - T.x1) + (T.x2 - T.x1) * (T.y0 - T.y1);
  float beta = xy_to_12 / zero_to_12;
  float gamma = 1.f - alpha - beta;
  if (!(gamma >= 0.f && alpha >= 0.f && beta >= 0.f)) return false;
  float u = T.u2 * alpha + T.u0 * beta + T.u1 * gamma;
  float v = T.v2 * alpha + T.v0 * beta + T.v1 * gamma;
  double fidx = floor((double)v) * (double)tw + (double)u;
  sample = 0.f;  // the reference reads past the texture here (UB); defined as 0
  if (fidx >= 0.0 && fidx < (double)tw * (double)th) sample = __ldg(tex + (int)fidx);
  return true;
}

__global__ void __launch_bounds__(kTileW * kTileH) ref_raster_kernel(RefFrame f, const RefTri* __restrict__ tris,
                                                                     int n_tris, const float* __restrict__ tex,
                                                                     char* __restrict__ out, size_t stride, int elem,
                                                                     int additive, int rx0, int ry0, int rw, int rh) {
  __shared__ unsigned short s_list[kMaxTris];
  __shared__ int s_count;
  const int tid = threadIdx.y * kTileW + threadIdx.x;
  // rw > 0: only the rectangle [rx0, rx0+rw) x [ry0, ry0+rh) is rastered, into a PACKED rw x rh output
  const int x0 = rx0 + blockIdx.x * kTileW, y0 = ry0 + blockIdx.y * kTileH;
  // triangles whose loop bounds touch this tile, kept in draw order: every thread tests its share of the triangles in
  // parallel (flags in shared memory), then one thread compacts the flags in index order
  __shared__ unsigned char s_hit[kMaxTris];
  for (int t = tid; t < n_tris; t += kTileW * kTileH) {
    const RefTri& T = tris[t];
    s_hit[t] = T.min_x < x0 + kTileW && T.max_x > x0 && T.min_y < y0 + kTileH && T.max_y > y0;
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int t = 0; t < n_tris; t++)
      if (s_hit[t]) s_list[n++] = (unsigned short)t;
    s_count = n;
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= f.W || y >= f.H) return;
  if (rw > 0 && (x >= rx0 + rw || y >= ry0 + rh)) return;
  double acc[3] = {0.0, 0.0, 0.0};
  const int n = f.has_sun ? s_count : 0;
  for (int q = 0; q < n; q++) {
    const RefTri& T = tris[s_list[q]];
    if (x < T.min_x || x >= T.max_x || y < T.min_y || y >= T.max_y) continue;
    float sample;
    if (!fill_sample(T, x, y, tex, f.tex_w, f.tex_h, sample)) continue;
    acc[0] += (double)sample * T.col[0];
    acc[1] += (double)sample * T.col[1];
    acc[2] += (double)sample * T.col[2];
  }
  char* o = out + (rw > 0 ? ((size_t)(x - rx0) + (size_t)(y - ry0) * rw) : ((size_t)x + (size_t)y * f.W)) * stride;
  if (elem == LFB_F32x3) {
    float* p = reinterpret_cast<float*>(o);
    if (additive) { p[0] += (float)acc[0]; p[1] += (float)acc[1]; p[2] += (float)acc[2]; }
    else { p[0] = (float)acc[0]; p[1] = (float)acc[1]; p[2] = (float)acc[2]; }
  } else {
    double* p = reinterpret_cast<double*>(o);
    if (additive) { p[0] += acc[0]; p[1] += acc[1]; p[2] += acc[2]; }
    else { p[0] = acc[0]; p[1] = acc[1]; p[2] = acc[2]; }
  }
}

}  // namespace

cudaError_t launch_ref_setup(const RefFrame& f, const int* pairs, const float* rgb_weight, RefTri* tris,
                             lfb_ref_ghost* ghosts, int* bbox, cudaStream_t s) {
  const int n = f.n_pairs * f.n_lambda;
  if (n <= 0) return cudaSuccess;
  ref_setup_kernel<<<(n + 31) / 32, 32, 0, s>>>(f, pairs, rgb_weight, tris, ghosts, bbox);
  return cudaGetLastError();
}

cudaError_t launch_ref_raster(const RefFrame& f, const RefTri* tris, int n_tris, const float* tex, void* out,
                              size_t stride, int elem, int additive, const int* rect, cudaStream_t s) {
  if (n_tris > kMaxTris) return cudaErrorInvalidValue;
  const int rx0 = rect ? rect[0] : 0, ry0 = rect ? rect[1] : 0;
  const int rw = rect ? rect[2] - rect[0] + 1 : 0, rh = rect ? rect[3] - rect[1] + 1 : 0;
  if (rect && (rw <= 0 || rh <= 0)) return cudaSuccess;
  const int gw = rect ? rw : f.W, gh = rect ? rh : f.H;
  dim3 block(kTileW, kTileH), grid((gw + kTileW - 1) / kTileW, (gh + kTileH - 1) / kTileH);
  ref_raster_kernel<<<grid, block, 0, s>>>(f, tris, n_tris, tex, (char*)out, stride, elem, additive, rx0, ry0, rw, rh);
  return cudaGetLastError();
}

}  // namespace lfb
