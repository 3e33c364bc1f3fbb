/* oracle/lf_oracle.c -- TEST INFRASTRUCTURE ONLY.  See lf_oracle.h for the parity
 * status of each function.  Compiled with -ffp-contract=off: the reference is built
 * for x86-64 without FMA, so every float/double operation below rounds exactly where
 * the reference's does.  Reference citations are file:line under /root/reference. */
#define _POSIX_C_SOURCE 200809L
#define _DEFAULT_SOURCE
#include "lf_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------- */
/* built-in prescription (data of src/pathtracer/pathtracer.cpp:539-556)      */
/* ------------------------------------------------------------------------- */
static const double k_thick[9] = {7.700, 1.850, 3.520, 1.850, 4.180, 3.000, 1.850, 7.270, 83.91};
static const double k_radius[9] = {30.810, -89.350, 580.380, -80.630, 28.340, 0, 0, 32.190, -52.990};
static const float k_n_rgb[3][9] = {
    {1.652f, 1.5991f, 1, 1.6396f, 1, 1, 1.5776f, 1.68990f, 1},
    {1.652f, 1.6113f, 1, 1.65f, 1, 1, 1.5885f, 1.6999f, 1},
    {1.652f, 1.6164f, 1, 1.6542f, 1, 1, 1.5930f, 1.7040f, 1}};
static const double k_anchor_nm[3] = {650.0, 550.0, 450.0};

static double lobe(double lam, double mu, double s1, double s2) {
  double t = (lam - mu) / (lam < mu ? s1 : s2);
  return exp(-0.5 * t * t);
}

/* CIE 1931 colour matching functions, multi-lobe Gaussian fit (Wyman, Sloan, Shirley
 * 2013), then XYZ -> linear sRGB; negative lobes clamped. */
static void lambda_to_rgb(double lam, double rgb[3]) {
  double X = 1.056 * lobe(lam, 599.8, 37.9, 31.0) + 0.362 * lobe(lam, 442.0, 16.0, 26.7) -
             0.065 * lobe(lam, 501.1, 20.4, 26.2);
  double Y = 0.821 * lobe(lam, 568.8, 46.9, 40.5) + 0.286 * lobe(lam, 530.9, 16.3, 31.1);
  double Z = 1.217 * lobe(lam, 437.0, 11.8, 36.0) + 0.681 * lobe(lam, 459.0, 26.0, 13.8);
  rgb[0] = 3.2406 * X - 1.5372 * Y - 0.4986 * Z;
  rgb[1] = -0.9689 * X + 1.8758 * Y + 0.0415 * Z;
  rgb[2] = 0.0557 * X - 0.2040 * Y + 1.0570 * Z;
  for (int c = 0; c < 3; c++)
    if (rgb[c] < 0) rgb[c] = 0;
}

/* The file-scope prescription of pathtracer.cpp:539-556 as an lfb_lens (+ ours: wavelengths other than the reference's
 * three tables, and the coating flag -- neither exists in the reference). */
int lfo_builtin_lens(lfb_lens* L, int n_lambda, float coating_lambda0_nm) {
  if (!L || n_lambda < 1 || n_lambda > LFB_MAX_LAMBDA) return LFB_ERR_INVALID;
  memset(L, 0, sizeof(*L));
  L->n_surfaces = 9;
  L->stop_index = 5;
  L->n_lambda = n_lambda;
  L->entrance_half_height = 14.5; /* pathtracer.cpp:737 */
  L->stop_half_height = 11.6;     /* :621 */
  L->stop_half_height_neg = 11.5; /* :624 */
  for (int k = 0; k < 9; k++) {
    L->thickness[k] = (float)k_thick[k];
    L->curvature[k] = k_radius[k] == 0 ? 0.f : (float)(1 / k_radius[k]); /* :556 */
    L->semi_aperture[k] = 14.5f;
  }
  if (n_lambda == 3) {
    for (int l = 0; l < 3; l++) {
      L->lambda_nm[l] = (float)k_anchor_nm[l];
      for (int k = 0; k < 9; k++) L->ior[l][k] = k_n_rgb[l][k];
      L->rgb_weight[l][l] = 1.f;
    }
  } else {
    double sum[3] = {0, 0, 0};
    double w[LFB_MAX_LAMBDA][3];
    for (int l = 0; l < n_lambda; l++) {
      double lam = 400.0 + 300.0 * (l + 0.5) / n_lambda;
      L->lambda_nm[l] = (float)lam;
      lambda_to_rgb(lam, w[l]);
      for (int c = 0; c < 3; c++) sum[c] += w[l][c];
    }
    for (int l = 0; l < n_lambda; l++)
      for (int c = 0; c < 3; c++) L->rgb_weight[l][c] = (float)(w[l][c] / sum[c]);
    /* n(lambda): piecewise two-term Cauchy n = A + B/lam^2 through the R,G,B anchors (linear in
     * u = 1/lam^2 on [650,550] and [550,450] nm, each segment continued beyond its anchor).  The
     * reference's three indices per glass are not consistent with one smooth 3-term Cauchy curve
     * (it turns over below 450 nm); this form passes through them exactly and stays monotone. */
    for (int k = 0; k < 9; k++) {
      double u[3], n[3];
      for (int a = 0; a < 3; a++) {
        double lm = k_anchor_nm[a] * 1e-3;
        u[a] = 1.0 / (lm * lm);
        n[a] = k_n_rgb[a][k];
      }
      double f01 = (n[1] - n[0]) / (u[1] - u[0]);
      double f12 = (n[2] - n[1]) / (u[2] - u[1]);
      for (int l = 0; l < n_lambda; l++) {
        double lm = (double)L->lambda_nm[l] * 1e-3;
        double x = 1.0 / (lm * lm);
        L->ior[l][k] = (float)(x <= u[1] ? n[1] + (x - u[1]) * f01 : n[1] + (x - u[1]) * f12);
      }
    }
  }
  for (int k = 0; k < 9; k++) {
    int glass_interface = 0;
    for (int l = 0; l < n_lambda; l++) {
      float before = k == 0 ? 1.f : L->ior[l][k - 1];
      if (before != L->ior[l][k]) glass_interface = 1;
    }
    L->coating_lambda0_nm[k] = glass_interface ? coating_lambda0_nm : 0.f;
  }
  return LFB_OK;
}

/* ------------------------------------------------------------------------- */
/* 2x2 ray-transfer algebra, with the rounding sequence of the reference      */
/* ------------------------------------------------------------------------- */
typedef struct { double a, b, c, d; } m2; /* [[a,b],[c,d]] */

/* make_2_matrix(float...) pathtracer.cpp:511-516: float arguments stored as doubles */
static m2 make2(float a, float b, float c, float d) { m2 m = {a, b, c, d}; return m; }
static m2 mT(float d) { return make2(1, d, 0, 1); }                                   /* :527-529 */
static m2 mR(float c, float n1, float n2) { return make2(1, 0, c * (n1 - n2) / n2, n1 / n2); } /* :531-533 float math */
static m2 mL(float c) { return make2(1, 0, 2 * c, 1); }                               /* :535-537 */

/* Matrix3x3::operator*(Matrix3x3) CGL/src/matrix3x3.cpp:99-108 with operator*(Vector3D)
 * :110-114: C(r,c) = B(0,c)*A(r,0) + B(1,c)*A(r,1) (+ 0*0 from the zero padding). */
static m2 mmul(m2 A, m2 B) {
  m2 C;
  C.a = B.a * A.a + B.c * A.b;
  C.c = B.a * A.c + B.c * A.d;
  C.b = B.b * A.a + B.d * A.b;
  C.d = B.b * A.c + B.d * A.d;
  return C;
}
/* invert2x2 pathtracer.cpp:519-525: float temporaries, float determinant, double scale */
static m2 minv(m2 m) {
  float a = (float)m.a, b = (float)m.b, c = (float)m.c, d = (float)m.d;
  double s = 1.0 / (a * d - b * c);
  m2 r = {d * s, (-b) * s, (-c) * s, a * s};
  return r;
}
/* M * Vector3D(r, theta, 0): x*col0 + y*col1 */
static void mapply(m2 M, double r, double th, double out[2]) {
  out[0] = r * M.a + th * M.b;
  out[1] = r * M.c + th * M.d;
}

static float n_before(const lfb_lens* L, int lambda, int k) { return k == 0 ? 1.00f : L->ior[lambda][k - 1]; }
static m2 surf_R(const lfb_lens* L, int lambda, int k) { /* create_Rs_for_color :559-569 */
  return mR(L->curvature[k], n_before(L, lambda, k), L->ior[lambda][k]);
}

/* The reference's globals Ts / Ls / R_<colour> (pathtracer.cpp:559-586) as 2x2 blocks, for the golden comparison. */
void lfo_prescription(const lfb_lens* L, int lambda, double* t, double* l, double* r) {
  for (int k = 0; k < L->n_surfaces; k++) {
    m2 T = mT(L->thickness[k]), Lm = mL(L->curvature[k]), R = surf_R(L, lambda, k);
    t[4 * k] = T.a; t[4 * k + 1] = T.b; t[4 * k + 2] = T.c; t[4 * k + 3] = T.d;
    l[4 * k] = Lm.a; l[4 * k + 1] = Lm.b; l[4 * k + 2] = Lm.c; l[4 * k + 3] = Lm.d;
    r[4 * k] = R.a; r[4 * k + 1] = R.b; r[4 * k + 2] = R.c; r[4 * k + 3] = R.d;
  }
}

/* The stop re-aim of pathtracer.cpp:618-629 / :656-667. */
static void reaim(const lfb_lens* L, m2 M, float r, float theta, double ray[2]) {
  double ap[2];
  mapply(M, ray[0], ray[1], ap);
  if (ap[0] > L->stop_half_height || ap[0] < -L->stop_half_height) {
    float r_a = (float)L->stop_half_height;
    if (r < 0) r_a = (float)-L->stop_half_height_neg;
    float r_e = (float)((r_a - M.b * theta) / M.a);
    ray[0] = r_e;
    ray[1] = theta;
  }
}

/* trace_ray_auto_before (pathtracer.cpp:588-641, which = 0) and trace_ray_auto_after (:643-689, which = 1), including the
 * by-value matrix products in the reference's order and the stop re-aim; pinned bit for bit by tests/golden/ref_vectors.npz. */
void lfo_trace_ray_auto(const lfb_lens* L, int lambda, int which, float r, float theta, int i,
                        int j, double out[2]) {
  const int n = L->n_surfaces, stop = L->stop_index;
  double ray[2] = {r, theta};
  if (i > j) { int t = i; i = j; j = t; }
  m2 M = make2(1, 0, 0, 1);
  if (which == 0) { /* trace_ray_auto_before :588-641 */
    for (int k = 0; k < j; k++) M = mmul(mmul(mT(L->thickness[k]), surf_R(L, lambda, k)), M);
    M = mmul(mL(L->curvature[j]), M);
    for (int k = j - 1; k > i; k--) M = mmul(mmul(minv(surf_R(L, lambda, k)), mT(L->thickness[k])), M);
    M = mmul(mmul(mmul(mT(L->thickness[i]), minv(mL(L->curvature[i]))), mT(L->thickness[i])), M);
    for (int k = i + 1; k < n; k++) {
      if (k == stop) {
        reaim(L, M, r, theta, ray);
        M = mmul(mT(L->thickness[k]), M);
        continue;
      }
      M = mmul(mmul(mT(L->thickness[k]), surf_R(L, lambda, k)), M);
    }
  } else { /* trace_ray_auto_after :643-689 */
    for (int k = 0; k < j; k++) {
      if (k == stop) {
        reaim(L, M, r, theta, ray);
        M = mmul(mT(L->thickness[k]), M);
        continue;
      }
      M = mmul(mmul(mT(L->thickness[k]), surf_R(L, lambda, k)), M);
    }
    M = mmul(mL(L->curvature[j]), M);
    for (int k = j - 1; k > i; k--) M = mmul(mmul(minv(surf_R(L, lambda, k)), mT(L->thickness[k])), M);
    M = mmul(mmul(mmul(mT(L->thickness[i]), minv(mL(L->curvature[i]))), mT(L->thickness[i])), M);
    for (int k = i + 1; k < n; k++) M = mmul(mmul(mT(L->thickness[k]), surf_R(L, lambda, k)), M);
  }
  mapply(M, ray[0], ray[1], out);
}

/* ------------------------------------------------------------------------- */
/* ghost quads: draw_ghost / shift_vertex / rasterize / fill                  */
/* ------------------------------------------------------------------------- */
typedef struct { double m[3][3]; } m3; /* m[row][col] */

static m3 m3mul(const m3* A, const m3* B) { /* matrix3x3.cpp:99-114 ordering */
  m3 C;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++)
      C.m[r][c] = (B->m[0][c] * A->m[r][0] + B->m[1][c] * A->m[r][1]) + B->m[2][c] * A->m[r][2];
  return C;
}

/* shift_vertex pathtracer.cpp:412-430.  cos/sin take a float argument, so the
 * reference resolves them to the float overloads (cosf/sinf). */
static void shift_vertex(float x, float y, float scale, float shift_amount, double ax, double ay,
                         double out[2]) {
  float ang = (float)atan((ay - 0.5) / (ax - 0.5));
  float cs = cosf(ang), sn = sinf(ang);
  m3 scaling = {{{scale, 0, 0}, {0, scale, 0}, {0, 0, 1}}};
  m3 rotation = {{{cs, -sn, 0}, {sn, cs, 0}, {0, 0, 1}}};
  m3 shift = {{{1, 0, shift_amount * cs}, {0, 1, shift_amount * sn}, {0, 0, 1}}};
  m3 sr = m3mul(&shift, &rotation);
  m3 srs = m3mul(&sr, &scaling);
  double v[3] = {x, y, 1};
  out[0] = (v[0] * srs.m[0][0] + v[1] * srs.m[0][1]) + v[2] * srs.m[0][2];
  out[1] = (v[0] * srs.m[1][0] + v[1] * srs.m[1][1]) + v[2] * srs.m[1][2];
}

typedef struct {
  const float* tex; int tw, th;
  int W, H;
  double* buf; /* W*H*3 */
} raster_ctx;

/* fill_textured_pixel pathtracer.cpp:305-343 (all float arithmetic). */
static void fill_pixel(raster_ctx* R, float x0, float y0, float u0, float v0, float x1, float y1,
                       float u1, float v1, float x2, float y2, float u2, float v2, int x, int y,
                       const double col[3]) {
  float xy_to_01 = -(y1 - y0) * (x - x0) + (x1 - x0) * (y - y0);
  float two_to_01 = -(y1 - y0) * (x2 - x0) + (x1 - x0) * (y2 - y0);
  float alpha = xy_to_01 / two_to_01;
  float xy_to_12 = -(y2 - y1) * (x - x1) + (x2 - x1) * (y - y1);
  float zero_to_12 = -(y2 - y1) * (x0 - x1) + (x2 - x1) * (y0 - y1);
  float beta = xy_to_12 / zero_to_12;
  float gamma = 1 - alpha - beta;
  if (gamma >= 0 && alpha >= 0 && beta >= 0) {
    float u = u2 * alpha + u0 * beta + u1 * gamma;
    float v = v2 * alpha + v0 * beta + v1 * gamma;
    double du = u, dv = v;
    double fidx = floor(dv) * (double)R->tw + du;
    float sample = 0.f; /* the reference reads out of bounds here (UB); defined as 0 */
    if (fidx >= 0 && fidx < (double)R->tw * R->th) sample = R->tex[(int)fidx];
    double* p = R->buf + 3 * ((size_t)x + (size_t)y * R->W);
    p[0] += sample * col[0];
    p[1] += sample * col[1];
    p[2] += sample * col[2];
  }
}

#define SWAPF(a, b) do { float t_ = a; a = b; b = t_; } while (0)
static float fminf3(float a, float b, float c) { float m = a < b ? a : b; return m < c ? m : c; }
static float fmaxf3(float a, float b, float c) { float m = a > b ? a : b; return m > c ? m : c; }

/* rasterize_textured_triangle pathtracer.cpp:346-410. */
static void raster_tri(raster_ctx* R, float x0, float y0, float u0, float v0, float x1, float y1,
                       float u1, float v1, float x2, float y2, float u2, float v2,
                       const double col[3]) {
  if (y1 < y0) { SWAPF(x0, x1); SWAPF(y0, y1); SWAPF(u0, u1); SWAPF(v0, v1); }
  if (y2 < y0) { SWAPF(x0, x2); SWAPF(y0, y2); SWAPF(u0, u2); SWAPF(v0, v2); }
  if (y2 < y1) { SWAPF(x1, x2); SWAPF(y1, y2); SWAPF(u1, u2); SWAPF(v1, v2); }
  x0 = (float)(x0 - 0.5); y0 = (float)(y0 - 0.5);
  x1 = (float)(x1 - 0.5); y1 = (float)(y1 - 0.5);
  x2 = (float)(x2 - 0.5); y2 = (float)(y2 - 0.5);
  int lo, hi;
  lo = (int)floorf(fminf3(x0, x1, x2)); float min_x = (float)(lo > 0 ? lo : 0);
  hi = (int)ceilf(fmaxf3(x0, x1, x2));  float max_x = (float)(hi < R->W - 1 ? hi : R->W - 1);
  lo = (int)floorf(y0);                 float min_y = (float)(lo > 0 ? lo : 0);
  hi = (int)ceilf(y2);                  float max_y = (float)(hi < R->H - 1 ? hi : R->H - 1);
  for (int y = (int)min_y; y < max_y; y++)
    for (int x = (int)min_x; x < max_x; x++)
      fill_pixel(R, x0, y0, u0, v0, x1, y1, u1, v1, x2, y2, u2, v2, x, y, col);
}

/* draw_ghost pathtracer.cpp:433-508 */
static void draw_ghost(raster_ctx* R, const float rgb[3], float r1, float r2, double ax, double ay,
                       lfb_ref_ghost* rec) {
  float shift_amt = (float)(-(r1 + r2) / 2 * 0.4);
  float scale_amt = (float)(fabsf(r2 - r1) * 0.2);
  double gb_mid_w = ceil(ax * (double)R->W);
  double gb_mid_h = ceil(ay * (double)R->H);
  double ul[2], ll[2], ur[2], lr[2];
  shift_vertex(-1, 1, scale_amt, shift_amt, ax, ay, ul);
  shift_vertex(-1, -1, scale_amt, shift_amt, ax, ay, ll);
  shift_vertex(1, 1, scale_amt, shift_amt, ax, ay, ur);
  shift_vertex(1, -1, scale_amt, shift_amt, ax, ay, lr);
  float intensity_scalar = 10;
  float size_scalar = 1 / (scale_amt * scale_amt);
  float k = intensity_scalar * size_scalar;
  double col[3] = {(double)rgb[0] * k, (double)rgb[1] * k, (double)rgb[2] * k};
  float tw = (float)R->tw, th = (float)R->th;
  float ulx = (float)(gb_mid_w + ul[0]), uly = (float)(gb_mid_h + ul[1]);
  float llx = (float)(gb_mid_w + ll[0]), lly = (float)(gb_mid_h + ll[1]);
  float urx = (float)(gb_mid_w + ur[0]), ury = (float)(gb_mid_h + ur[1]);
  float lrx = (float)(gb_mid_w + lr[0]), lry = (float)(gb_mid_h + lr[1]);
  if (rec) {
    rec->verts[0][0] = ulx; rec->verts[0][1] = uly; rec->verts[1][0] = llx; rec->verts[1][1] = lly;
    rec->verts[2][0] = urx; rec->verts[2][1] = ury; rec->verts[3][0] = lrx; rec->verts[3][1] = lry;
    rec->scale = scale_amt; rec->shift = shift_amt;
  }
  raster_tri(R, ulx, uly, 0, 0, llx, lly, 0, th, urx, ury, tw, 0, col);
  raster_tri(R, lrx, lry, 0, 0, llx, lly, 0, th, urx, ury, tw, 0, col); /* sic: mirrored uv, :498 */
}

/* Pair enumeration shared by every mode: REF = same side of the stop, in the
 * reference's order (before-stop pairs, then after-stop pairs, :735-762). */
static int list_pairs(const lfb_lens* L, int pair_set, int pairs[][2]) {
  int n = 0, S = L->stop_index, N = L->n_surfaces;
  for (int i = 0; i < S; i++)
    for (int j = i + 1; j < S; j++) { pairs[n][0] = i; pairs[n][1] = j; n++; }
  for (int i = S + 1; i < N; i++)
    for (int j = i + 1; j < N; j++) { pairs[n][0] = i; pairs[n][1] = j; n++; }
  if (pair_set == LFB_PAIRS_ALL)
    for (int i = 0; i < S; i++)
      for (int j = S + 1; j < N; j++) { pairs[n][0] = i; pairs[n][1] = j; n++; }
  return n;
}

/* PathTracer::generate_ghost_buffer (pathtracer.cpp:714-762): 13 pairs x 3 colours of marginal rays -> draw_ghost ->
 * two textured triangles each, accumulated in double; pinned bit for bit to the compiled reference (tests/golden/ref_vectors.npz). */
int lfo_generate_ghost_buffer(const lfb_lens* L, const float* tex, int tw, int th, int W, int H,
                              double ax, double ay, float angle, double* out,
                              lfb_ref_ghost* ghosts, int cap) {
  raster_ctx R = {tex, tw, th, W, H, out};
  memset(out, 0, sizeof(double) * 3 * (size_t)W * H); /* clear + resize :719-720 */
  if (ax == 0 && ay == 0) return 0;                    /* :724-726 */
  int pairs[LFB_MAX_SURFACES * LFB_MAX_SURFACES][2];
  int np = list_pairs(L, LFB_PAIRS_REF, pairs), ng = 0;
  const float P = (float)L->entrance_half_height;
  for (int p = 0; p < np; p++) {
    int i = pairs[p][0], j = pairs[p][1];
    int which = i > L->stop_index ? 1 : 0;
    for (int c = 0; c < L->n_lambda; c++) {
      double s1[2], s2[2];
      lfo_trace_ray_auto(L, c, which, P, angle, i, j, s1);
      lfo_trace_ray_auto(L, c, which, -P, angle, i, j, s2);
      lfb_ref_ghost* rec = (ghosts && ng < cap) ? &ghosts[ng] : NULL;
      if (rec) { rec->i = i; rec->j = j; rec->colour = c; rec->pad = 0; rec->r1 = s1[0]; rec->r2 = s2[0]; }
      draw_ghost(&R, L->rgb_weight[c], (float)s1[0], (float)s2[0], ax, ay, rec);
      ng++;
    }
  }
  return ng;
}

/* ------------------------------------------------------------------------- */
/* PARAXIAL_GRID: per-ghost system matrices                                   */
/* ------------------------------------------------------------------------- */
/* The matrix chain of trace_ray_auto_* (pathtracer.cpp:588-689) WITHOUT the stop re-aim, split at every stop crossing;
 * physical_backward = 1 replaces the reference's R_k^-1 on the backward legs (:607-608) by the physical refraction (ours). */
int lfo_paraxial_system(const lfb_lens* L, int lambda, int i, int j, int physical_backward,
                        double cross[3][4], double full[4]) {
  const int n = L->n_surfaces, stop = L->stop_index;
  int nc = 0;
  m2 M = make2(1, 0, 0, 1);
#define RECORD() do { cross[nc][0] = M.a; cross[nc][1] = M.b; cross[nc][2] = M.c; cross[nc][3] = M.d; nc++; } while (0)
  if (i < 0) { /* direct path */
    for (int k = 0; k < n; k++) {
      if (k == stop) { RECORD(); M = mmul(mT(L->thickness[k]), M); continue; }
      M = mmul(mmul(mT(L->thickness[k]), surf_R(L, lambda, k)), M);
    }
  } else {
    for (int k = 0; k < j; k++) {
      if (k == stop) { RECORD(); M = mmul(mT(L->thickness[k]), M); continue; }
      M = mmul(mmul(mT(L->thickness[k]), surf_R(L, lambda, k)), M);
    }
    M = mmul(mL(L->curvature[j]), M);
    for (int k = j - 1; k > i; k--) {
      m2 back = physical_backward
                    ? mR(-L->curvature[k], L->ior[lambda][k], n_before(L, lambda, k))
                    : minv(surf_R(L, lambda, k));
      M = mmul(mmul(back, mT(L->thickness[k])), M);
      if (k == stop) RECORD();
    }
    M = mmul(mmul(mmul(mT(L->thickness[i]), minv(mL(L->curvature[i]))), mT(L->thickness[i])), M);
    for (int k = i + 1; k < n; k++) {
      if (k == stop) { RECORD(); M = mmul(mT(L->thickness[k]), M); continue; }
      M = mmul(mmul(mT(L->thickness[k]), surf_R(L, lambda, k)), M);
    }
  }
#undef RECORD
  full[0] = M.a; full[1] = M.b; full[2] = M.c; full[3] = M.d;
  return nc;
}

/* ------------------------------------------------------------------------- */
/* shared by the grid modes: aperture lookup, pixel mapping                   */
/* ------------------------------------------------------------------------- */
static double mask_lookup(const lfb_lens* L, const float* tex, int tw, int th, double xa, double ya) {
  double h = L->stop_half_height;
  double u = (xa / h * 0.5 + 0.5) * tw;
  double v = (0.5 - ya / h * 0.5) * th;
  double fu = floor(u), fv = floor(v);
  if (!(fu >= 0 && fu < tw && fv >= 0 && fv < th)) return 0.0;
  return tex[(int)fv * tw + (int)fu];
}

typedef struct { double sx, sy, cs, sn, ppu; } pixmap;

static pixmap make_pixmap(const lfb_light* lt, const lfb_params* P) {
  pixmap m;
  double dx = lt->ns_x - 0.5, dy = lt->ns_y - 0.5;
  if (P->physical_mapping) {
    /* ours (lfb_params.physical_mapping, no reference counterpart): origin at the image centre, meridional axis along the
     * light's azimuth; the sign of to_pixel's X = -ppu xs is folded into the rotation */
    double phi = (dx == 0 && dy == 0) ? 0.0 : atan2(dy * (double)P->height, dx * (double)P->width);
    m.cs = -cos(phi); m.sn = -sin(phi);
    m.sx = 0.5 * (double)P->width; m.sy = 0.5 * (double)P->height;
  } else {
    float ang = (dx == 0 && dy == 0) ? 0.f : (float)atan(dy / dx); /* shift_vertex :414 */
    m.cs = cosf(ang); m.sn = sinf(ang);
    m.sx = ceil(lt->ns_x * (double)P->width);  /* draw_ghost :462-463 */
    m.sy = ceil(lt->ns_y * (double)P->height);
  }
  m.ppu = P->px_per_unit > 0 ? P->px_per_unit : 0.4f;
  return m;
}

static void to_pixel(const pixmap* m, double xs, double ys, double* px, double* py) {
  double X = -m->ppu * xs, Y = m->ppu * ys;
  *px = m->sx + (X * m->cs - Y * m->sn);
  *py = m->sy + (X * m->sn + Y * m->cs);
}

/* ------------------------------------------------------------------------- */
/* EXACT_GRID physics (absent from the reference; see lf_oracle.h)            */
/* ------------------------------------------------------------------------- */
/* Fresnel reflectance of a bare interface, or of one carrying a quarter-wave film designed for lambda0 (Hullin et al. 2011,
 * the formulas quoted in SURVEY.md 8c).  PARITY UNPINNED by the reference (its Fresnel code is an empty stub,
 * advanced_bsdf.cpp:52-60); pinned by the analytic invariants of tests/test_oracle_physics.py. */
double lfo_reflectance(double n0, double n2, double cos0, double lambda0, double lambda) {
  if (n0 == n2) return 0.0;
  double sin2 = 1.0 - cos0 * cos0;
  double e2 = n0 / n2, k2 = 1.0 - e2 * e2 * sin2;
  if (k2 < 0) return 1.0; /* total internal reflection */
  double cos2 = sqrt(k2);
  if (lambda0 <= 0) {
    double rs = (n0 * cos0 - n2 * cos2) / (n0 * cos0 + n2 * cos2);
    double rp = (n2 * cos0 - n0 * cos2) / (n2 * cos0 + n0 * cos2);
    return 0.5 * (rs * rs + rp * rp);
  }
  double n1 = sqrt(n0 * n2);
  if (n1 < 1.38) n1 = 1.38; /* MgF2 floor */
  double d1 = lambda0 / (4.0 * n1);
  double e1 = n0 / n1, k1 = 1.0 - e1 * e1 * sin2;
  if (k1 < 0) return 1.0;
  double cos1 = sqrt(k1);
  double cd = cos(4.0 * M_PI * n1 * d1 * cos1 / lambda);
  double r01s = (n0 * cos0 - n1 * cos1) / (n0 * cos0 + n1 * cos1);
  double r12s = (n1 * cos1 - n2 * cos2) / (n1 * cos1 + n2 * cos2);
  double r01p = (n1 * cos0 - n0 * cos1) / (n1 * cos0 + n0 * cos1);
  double r12p = (n2 * cos1 - n1 * cos2) / (n2 * cos1 + n1 * cos2);
  double ps = r01s * r12s, pp = r01p * r12p;
  double Rs = (r01s * r01s + r12s * r12s + 2 * ps * cd) / (1 + ps * ps + 2 * ps * cd);
  double Rp = (r01p * r01p + r12p * r12p + 2 * pp * cd) / (1 + pp * pp + 2 * pp * cd);
  return 0.5 * (Rs + Rp);
}

typedef struct {
  double o[3], d[3], w;
  unsigned flags;
  double xa, ya;
} ray_t;

/* intersect surface k; on success o is the hit point, nrm the unit normal facing
 * against d and *cos0 = -nrm.d */
static int hit_surface(const lfb_lens* L, const double* zv, int k, ray_t* r, double nrm[3], double* cos0) {
  double c = L->curvature[k];
  double p[3] = {r->o[0], r->o[1], r->o[2] - zv[k]};
  double pd = p[0] * r->d[0] + p[1] * r->d[1] + p[2] * r->d[2];
  double pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  double B = c * pd - r->d[2];
  double Cq = c * pp - 2 * p[2];
  double disc = B * B - c * Cq;
  if (disc < 0) { r->flags |= LFB_RAY_MISSED; return 0; }
  double sq = sqrt(disc);
  double t = -Cq / (B + (B < 0 ? -sq : sq));
  double h[3] = {p[0] + t * r->d[0], p[1] + t * r->d[1], p[2] + t * r->d[2]};
  r->o[0] = h[0]; r->o[1] = h[1]; r->o[2] = h[2] + zv[k];
  double sa = L->semi_aperture[k];
  if (h[0] * h[0] + h[1] * h[1] > sa * sa) { r->flags |= LFB_RAY_VIGNETTED; return 0; }
  nrm[0] = -c * h[0]; nrm[1] = -c * h[1]; nrm[2] = 1 - c * h[2];
  double nd = nrm[0] * r->d[0] + nrm[1] * r->d[1] + nrm[2] * r->d[2];
  if (nd > 0) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; nd = -nd; }
  *cos0 = -nd;
  return 1;
}

static int refract_at(const lfb_lens* L, const double* zv, int lambda, int k, int forward, ray_t* r) {
  double nrm[3], cos0;
  if (!hit_surface(L, zv, k, r, nrm, &cos0)) return 0;
  double na = n_before(L, lambda, k), nb = L->ior[lambda][k];
  double n0 = forward ? na : nb, n2 = forward ? nb : na;
  if (n0 == n2) return 1;
  double eta = n0 / n2, k2 = 1 - eta * eta * (1 - cos0 * cos0);
  if (k2 < 0) { r->flags |= LFB_RAY_TIR; return 0; }
  double cos2 = sqrt(k2), f = eta * cos0 - cos2;
  for (int a = 0; a < 3; a++) r->d[a] = eta * r->d[a] + f * nrm[a];
  r->w *= 1.0 - lfo_reflectance(n0, n2, cos0, L->coating_lambda0_nm[k], L->lambda_nm[lambda]);
  return 1;
}

static int reflect_at(const lfb_lens* L, const double* zv, int lambda, int k, int forward, ray_t* r) {
  double nrm[3], cos0;
  if (!hit_surface(L, zv, k, r, nrm, &cos0)) return 0;
  double na = n_before(L, lambda, k), nb = L->ior[lambda][k];
  double n0 = forward ? na : nb, n2 = forward ? nb : na;
  for (int a = 0; a < 3; a++) r->d[a] = r->d[a] + 2 * cos0 * nrm[a];
  r->w *= lfo_reflectance(n0, n2, cos0, L->coating_lambda0_nm[k], L->lambda_nm[lambda]);
  return 1;
}

static void to_plane(ray_t* r, double z) {
  double t = (z - r->o[2]) / r->d[2];
  r->o[0] += t * r->d[0]; r->o[1] += t * r->d[1]; r->o[2] = z;
}

static void cross_stop(const lfb_lens* L, const double* zv, const float* tex, int tw, int th, ray_t* r) {
  to_plane(r, zv[L->stop_index]);
  r->xa = r->o[0]; r->ya = r->o[1];
  double m = mask_lookup(L, tex, tw, th, r->xa, r->ya);
  if (m == 0) r->flags |= LFB_RAY_STOPPED;
  r->w *= m;
}

static void trace_exact(const lfb_lens* L, const float* tex, int tw, int th, int lambda, int i, int j,
                        double x, double y, double theta, double inv_dist, ray_t* r) {
  const int n = L->n_surfaces, stop = L->stop_index;
  double zv[LFB_MAX_SURFACES + 1];
  zv[0] = 0;
  for (int k = 0; k < n; k++) zv[k + 1] = zv[k] + (double)L->thickness[k];
  r->o[0] = x; r->o[1] = y; r->o[2] = 0;
  r->d[0] = sin(theta); r->d[1] = 0; r->d[2] = cos(theta);
  r->w = 1; r->flags = 0; r->xa = r->ya = NAN;
  if (inv_dist != 0) {
    /* point light at -D (sin t, 0, cos t) (lfb_light.distance, include/lfb200.h): direction (E - P)/D normalised; weight =
     * irradiance at the entrance point relative to the vertex = (D/|E-P|)^2 * cos(incidence)/cos t = |v|^-3 */
    double vx = x * inv_dist + r->d[0], vy = y * inv_dist, vz = r->d[2];
    double q = vx * vx + vy * vy + vz * vz;
    double len = sqrt(q);
    r->d[0] = vx / len; r->d[1] = vy / len; r->d[2] = vz / len;
    r->w = 1 / (q * len);
  }
#define STEP(expr) do { if (!(expr)) { r->w = 0; r->o[0] = r->o[1] = NAN; return; } } while (0)
  if (i < 0) {
    for (int k = 0; k < n; k++) {
      if (k == stop) cross_stop(L, zv, tex, tw, th, r); else STEP(refract_at(L, zv, lambda, k, 1, r));
    }
  } else {
    for (int k = 0; k < j; k++) {
      if (k == stop) cross_stop(L, zv, tex, tw, th, r); else STEP(refract_at(L, zv, lambda, k, 1, r));
    }
    STEP(reflect_at(L, zv, lambda, j, 1, r));
    for (int k = j - 1; k > i; k--) {
      if (k == stop) cross_stop(L, zv, tex, tw, th, r); else STEP(refract_at(L, zv, lambda, k, 0, r));
    }
    STEP(reflect_at(L, zv, lambda, i, 0, r));
    for (int k = i + 1; k < n; k++) {
      if (k == stop) cross_stop(L, zv, tex, tw, th, r); else STEP(refract_at(L, zv, lambda, k, 1, r));
    }
  }
#undef STEP
  to_plane(r, zv[n]);
}

/* ------------------------------------------------------------------------- */
/* grid tracing                                                               */
/* ------------------------------------------------------------------------- */
static double light_inv_dist(const lfb_light* lt) {
  return (lt->distance > 0 && isfinite(lt->distance)) ? 1.0 / lt->distance : 0.0;
}

static void trace_one(const lfb_lens* L, const float* tex, int tw, int th, const lfb_light* lt,
                      const lfb_params* P, const pixmap* pm, int i, int j, int lambda, int nc,
                      double cross[3][4], const double full[4], int a, int b, lfb_ray_hit* h) {
  const int N = P->grid_n;
  const double Pe = L->entrance_half_height;
  double x = -Pe + (a + 0.5) * (2 * Pe / N);
  double y = -Pe + (b + 0.5) * (2 * Pe / N);
  memset(h, 0, sizeof(*h));
  if (P->mode == LFB_MODE_PARAXIAL_GRID) {
    double th_ = lt->theta, thy = 0, w = 1;
    const double inv_dist = light_inv_dist(lt);
    if (inv_dist != 0) { th_ = x * inv_dist + th_; thy = y * inv_dist; }  /* paraxial point light */
    h->x_ap = h->y_ap = NAN;
    for (int c = 0; c < nc; c++) {
      double xa = x * cross[c][0] + th_ * cross[c][1];
      double ya = y * cross[c][0];
      if (inv_dist != 0) ya = ya + thy * cross[c][1];
      double m = mask_lookup(L, tex, tw, th, xa, ya);
      if (m == 0) h->flags |= LFB_RAY_STOPPED;
      w *= m;
      h->x_ap = xa; h->y_ap = ya;
    }
    h->x_s = x * full[0] + th_ * full[1];
    h->y_s = y * full[0];
    if (inv_dist != 0) h->y_s = h->y_s + thy * full[1];
    h->weight = w;
  } else {
    ray_t r;
    trace_exact(L, tex, tw, th, lambda, i, j, x, y, lt->theta, light_inv_dist(lt), &r);
    h->x_s = r.o[0]; h->y_s = r.o[1]; h->x_ap = r.xa; h->y_ap = r.ya;
    h->weight = r.w; h->flags = r.flags;
  }
  if (h->x_s == h->x_s) {
    to_pixel(pm, h->x_s, h->y_s, &h->px, &h->py);
    double fx = floor(h->px), fy = floor(h->py);
    if (!(fx >= 0 && fx < P->width && fy >= 0 && fy < P->height)) h->flags |= LFB_RAY_OFF_SENSOR;
  } else {
    h->px = h->py = NAN;
  }
}

/* Per-ray records of one ghost's N x N grid (the checker for lfb_dump_rays): PARAXIAL_GRID applies the reference's system
 * matrices per axis (pinned through lfo_trace_ray_auto); EXACT_GRID is ours (unpinned by the reference, see above). */
int lfo_trace_grid(const lfb_lens* L, const float* tex, int tw, int th, const lfb_light* lt,
                   const lfb_params* P, int i, int j, int lambda, lfb_ray_hit* out) {
  if (P->mode != LFB_MODE_PARAXIAL_GRID && P->mode != LFB_MODE_EXACT_GRID) return LFB_ERR_INVALID;
  double cross[3][4], full[4];
  int nc = 0;
  if (P->mode == LFB_MODE_PARAXIAL_GRID)
    nc = lfo_paraxial_system(L, lambda, i, j, P->physical_backward, cross, full);
  pixmap pm = make_pixmap(lt, P);
  const int N = P->grid_n;
  for (int b = 0; b < N; b++)
    for (int a = 0; a < N; a++)
      trace_one(L, tex, tw, th, lt, P, &pm, i, j, lambda, nc, cross, full, a, b, &out[(size_t)b * N + a]);
  return LFB_OK;
}

/* fixed-point deposit shared by nearest / bilinear splats */
static void deposit(int64_t* acc, const lfb_params* P, long ix, long iy, const double val[3], double scale) {
  if (ix < 0 || ix >= P->width || iy < 0 || iy >= P->height) return;
  int64_t* p = acc + 3 * ((size_t)ix + (size_t)iy * P->width);
  for (int c = 0; c < 3; c++) p[c] += (int64_t)llrint(val[c] * scale);
}

static void render_job(const lfb_lens* L, const float* tex, int tw, int th, const lfb_light* lt,
                       const lfb_params* P, int i, int j, int lambda, int64_t* acc) {
  double cross[3][4], full[4];
  int nc = 0;
  if (P->mode == LFB_MODE_PARAXIAL_GRID)
    nc = lfo_paraxial_system(L, lambda, i, j, P->physical_backward, cross, full);
  pixmap pm = make_pixmap(lt, P);
  const int N = P->grid_n;
  const int bits = P->fixed_point_bits > 0 ? P->fixed_point_bits : 40;
  const double scale = ldexp(1.0, bits);
  const double cell = 2 * L->entrance_half_height / N;
  const double area = cell * cell * pm.ppu * pm.ppu;
  double chan[3];
  for (int c = 0; c < 3; c++) chan[c] = (double)lt->radiance[c] * (double)L->rgb_weight[lambda][c] * area;
  for (int b = 0; b < N; b++)
    for (int a = 0; a < N; a++) {
      lfb_ray_hit h;
      trace_one(L, tex, tw, th, lt, P, &pm, i, j, lambda, nc, cross, full, a, b, &h);
      if (!(h.weight > 0) || h.px != h.px) continue;
      if (P->splat == LFB_SPLAT_NEAREST) {
        double v[3] = {h.weight * chan[0], h.weight * chan[1], h.weight * chan[2]};
        deposit(acc, P, (long)floor(h.px), (long)floor(h.py), v, scale);
      } else {
        double qx = h.px - 0.5, qy = h.py - 0.5;
        double fx0 = floor(qx), fy0 = floor(qy);
        double fx = qx - fx0, fy = qy - fy0;
        double wt[4] = {(1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy, fx * fy};
        for (int t = 0; t < 4; t++) {
          double ww = h.weight * wt[t];
          double v[3] = {ww * chan[0], ww * chan[1], ww * chan[2]};
          deposit(acc, P, (long)fx0 + (t & 1), (long)fy0 + (t >> 1), v, scale);
        }
      }
    }
}

typedef struct { int light, i, j, lambda; } job_t;

static int job_cost(const lfb_lens* L, const job_t* q) { /* ray-surface interactions per ray */
  return q->i < 0 ? L->n_surfaces + 1 : 2 * (q->j - q->i) + L->n_surfaces + 1;
}

/* All jobs in (light, pair, lambda) order.  Sharding: whole (light, lambda) groups round-robin as far as they divide evenly
 * among the shards; the jobs of the remaining groups are put in longest-processing-time-first order (stable) and dealt
 * one by one (q % shard_count == shard_index). */
static int list_jobs(const lfb_lens* L, const lfb_params* P, int n_lights, job_t** out) {
  int pairs[LFB_MAX_SURFACES * LFB_MAX_SURFACES][2];
  int np = list_pairs(L, P->pair_set, pairs);
  int total = n_lights * (np + (P->include_direct ? 1 : 0)) * L->n_lambda, n = 0;
  job_t* J = (job_t*)malloc(sizeof(job_t) * (total > 0 ? total : 1));
  for (int l = 0; l < n_lights; l++)
    for (int p = -1; p < np; p++) {
      if (p < 0 && !P->include_direct) continue;
      for (int c = 0; c < L->n_lambda; c++) {
        job_t q = {l, p < 0 ? -1 : pairs[p][0], p < 0 ? -1 : pairs[p][1], c};
        J[n++] = q;
      }
    }
  if (P->shard_count > 1) {
    /* whole (light, lambda) groups in contiguous blocks as far as they divide evenly (groups 0 .. floor(G/S)*S - 1); the jobs of the
     * remaining groups in longest-processing-time-first order (stable), dealt one by one */
    const int S = P->shard_count, G = n_lights * L->n_lambda, n_whole = (G / S) * S;
    job_t* keep = (job_t*)malloc(sizeof(job_t) * (n > 0 ? n : 1));
    job_t* rest = (job_t*)malloc(sizeof(job_t) * (n > 0 ? n : 1));
    int m = 0, nr = 0;
    for (int q = 0; q < n; q++) {
      int grp = J[q].light * L->n_lambda + J[q].lambda;
      if (grp < n_whole) { if (grp / (n_whole / S) == P->shard_index) keep[m++] = J[q]; }
      else rest[nr++] = J[q];
    }
    for (int a = 1; a < nr; a++) { /* stable insertion sort, decreasing cost */
      job_t q = rest[a];
      int b = a - 1;
      while (b >= 0 && job_cost(L, &rest[b]) < job_cost(L, &q)) { rest[b + 1] = rest[b]; b--; }
      rest[b + 1] = q;
    }
    for (int q = 0; q < nr; q++)
      if (q % S == P->shard_index) keep[m++] = rest[q];
    memcpy(J, keep, sizeof(job_t) * (size_t)m);
    n = m;
    free(keep); free(rest);
  }
  *out = J;
  return n;
}

/* A whole frame in any mode, with the engine's job list, sharding and u64 fixed-point deposits restated (ours: the
 * reference has none of these; REF_QUADS delegates to lfo_generate_ghost_buffer). */
int lfo_render(const lfb_lens* L, const float* tex, int tw, int th, const lfb_light* lights,
               int n_lights, const lfb_params* P, double* out, int64_t* accum) {
  const size_t npx = (size_t)P->width * P->height;
  if (P->mode == LFB_MODE_REF_QUADS) {
    /* the reference handles a single sun: the last on-screen light wins (:50-53) */
    memset(out, 0, sizeof(double) * 3 * npx);
    if (n_lights < 1) return LFB_OK;
    const lfb_light* lt = &lights[n_lights - 1];
    lfo_generate_ghost_buffer(L, tex, tw, th, P->width, P->height, lt->ns_x, lt->ns_y, lt->theta, out, NULL, 0);
    return LFB_OK;
  }
  int64_t* acc = accum ? accum : (int64_t*)malloc(sizeof(int64_t) * 3 * npx);
  if (!acc) return LFB_ERR_NOMEM;
  memset(acc, 0, sizeof(int64_t) * 3 * npx);
  job_t* J;
  int nj = list_jobs(L, P, n_lights, &J);
  for (int q = 0; q < nj; q++)
    render_job(L, tex, tw, th, &lights[J[q].light], P, J[q].i, J[q].j, J[q].lambda, acc);
  free(J);
  const int bits = P->fixed_point_bits > 0 ? P->fixed_point_bits : 40;
  const double inv = ldexp(1.0, -bits);
  if (out)
    for (size_t p = 0; p < 3 * npx; p++) out[p] = (double)acc[p] * inv;
  if (!accum) free(acc);
  return LFB_OK;
}

/* ------------------------------------------------------------------------- */
/* starburst: pathtracer.cpp:901-1052                           */
/* ------------------------------------------------------------------------- */
static double convert_coordinate(int pixel_coord, int length, int is_y) { /* :936-945, float arithmetic as written */
  double cc;
  if (is_y) cc = -((float)pixel_coord) + ((float)length / 2.0);
  else cc = ((float)pixel_coord) - ((float)length / 2.0);
  return cc >= 0 ? cc : length + cc;
}

/* PathTracer::raytrace_starburst (pathtracer.cpp:947-1004) + compute_phase (:918-934) + convertCoordinate (:936-945) +
 * calculate_irradiance_falloff (:1030-1052, its 16 random samples replaced by the 4x4 stratified midpoints), brute force per
 * pixel like the reference; the DFT scalar is pinned to the compiled reference by tests/golden/starburst.npz. */
int lfo_starburst_pixels(const float* tex, int tw, int th, int W, int H, int n_lights, const double* fo,
                         const double* rad, double flare_radius, double flare_intensity, const int* xs,
                         const int* ys, int n, double* out_dft, double* out_falloff) {
  if (n_lights < 1) return LFB_ERR_INVALID;
  /* CameraApertureTexture::init (camera.h:61-73): total and bbox of texels > 0 */
  double total = 0;
  int min_x = tw, min_y = tw, max_x = -1, max_y = -1;
  for (int y = 0; y < th; y++)
    for (int x = 0; x < tw; x++) {
      float v = tex[(size_t)y * tw + x];
      total += v;
      if (v > 0) {
        if (x < min_x) min_x = x;
        if (y < min_y) min_y = y;
        if (x > max_x) max_x = x;
        if (y > max_y) max_y = y;
      }
    }
  /* compute_phase(0, ...) :918-934 */
  double lr = ceil(fo[0] * (double)W), ud = ceil(fo[1] * (double)H);
  const double org_x = lr, org_y = ud;
  lr -= (double)W / 2.0;
  ud = -ud + (double)H / 2.0;
  for (int k = 0; k < n; k++) {
    const int x = xs[k], y = ys[k];
    const double xprime = convert_coordinate(x, W, 0), yprime = convert_coordinate(y, H, 1);
    double re = 0, im = 0;
    for (int yc = min_y; yc <= max_y; yc++)
      for (int xc = min_x; xc <= max_x; xc++) {
        const double a = (double)tex[(size_t)yc * tw + xc];
        const double u = ((double)xc / (double)tw) - 0.5;
        const double v = ((double)yc / (double)tw) - 0.5; /* sic: divided by the WIDTH, :964 */
        const double e1 = u * xprime + v * yprime, e2 = u * lr + v * ud;
        /* a * additional_phase * complex_exponential, complex products as std::complex does */
        const double c1 = cos(2.0 * M_PI * e1), s1 = -sin(2.0 * M_PI * e1);
        const double c2 = cos(2.0 * M_PI * e2), s2 = sin(2.0 * M_PI * e2);
        const double pr = a * c2, pi = a * s2;
        re += pr * c1 - pi * s1;
        im += pr * s1 + pi * c1;
      }
    double I = sqrt(re * re + im * im) / total;
    const double dx = org_x - x, dy = org_y - y, dist = sqrt(dx * dx + dy * dy);
    if (dist > (double)tw / 2.0) { /* suppression :983-989 */
      const double factor = ((double)tw / 2.0) / dist;
      I = pow(factor, 8.0) * I;
    } else if (dist <= flare_radius) { /* amplification :990-996 */
      I = pow(I, dist / flare_radius);
    }
    double intensity = -flare_intensity + 3.0;
    if (intensity <= 0) intensity = 2.0;
    out_dft[k] = pow(I, intensity);
    if (out_falloff) { /* :1030-1052 with 4x4 stratified midpoints instead of 16 random samples */
      double f[3] = {0, 0, 0};
      for (int sy = 0; sy < 4; sy++)
        for (int sx = 0; sx < 4; sx++) {
          const double px = x + (sx + 0.5) / 4, py = y + (sy + 0.5) / 4;
          for (int l = 0; l < n_lights; l++) {
            const double ox = fo[2 * l] * (double)W - px, oy = fo[2 * l + 1] * (double)H - py;
            double r = sqrt(ox * ox + oy * oy) - 5.0;
            r = 1 + (r > 0 ? r : 0);
            const double r2 = pow(r, 1.5);
            for (int c = 0; c < 3; c++) f[c] += rad[3 * l + c] / r2;
          }
        }
      for (int c = 0; c < 3; c++) out_falloff[3 * k + c] = f[c] / 16.0;
    }
  }
  return LFB_OK;
}

/* HDRImageBuffer::toColor (util/image.h:208-223) + ImageBuffer::update_pixel (:53-62); pinned bit for bit by tests/golden/tocolor.npz. */
void lfo_to_color(const double* hdr, int W, int H, uint32_t* out) {
  const float gamma = 2.2f, level = 1.0f;
  const float one_over_gamma = 1.0f / gamma;
  const float exposure = (float)sqrt(pow(2, level));
  for (size_t p = 0; p < (size_t)W * H; p++) {
    float c[3];
    for (int k = 0; k < 3; k++) {
      double v = pow(hdr[3 * p + k] * exposure, one_over_gamma);
      v = (1.0 < v) ? 1.0 : v; /* std::min(pow, 1.0): a NaN pow (negative radiance) passes through */
      v = (0.0 < v) ? v : 0.0; /* std::max(0.0, .): ... and becomes 0 here */
      c[k] = (float)v;
    }
    uint32_t px = 0;
    for (int k = 0; k < 3; k++) {
      float v = c[k] < 0.f ? 0.f : (c[k] > 1.f ? 1.f : c[k]);
      if (!(c[k] == c[k])) v = 0.f;
      px |= ((uint32_t)(v * 255)) << (8 * k); /* r | g << 8 | b << 16 */
    }
    out[p] = px | 0xFF000000u;
  }
}

/* ------------------------------------------------------------------------- */
/* timed multi-threaded render (cpu_baseline "port")                          */
/* ------------------------------------------------------------------------- */
typedef struct {
  const lfb_lens* L; const float* tex; int tw, th; const lfb_light* lights; const lfb_params* P;
  job_t* J; int nj, tid, nt; int64_t* acc;
} worker_t;

static void* worker(void* arg) {
  worker_t* w = (worker_t*)arg;
  for (int q = w->tid; q < w->nj; q += w->nt)
    render_job(w->L, w->tex, w->tw, w->th, &w->lights[w->J[q].light], w->P, w->J[q].i, w->J[q].j,
               w->J[q].lambda, w->acc);
  return NULL;
}

/* lfo_render's ghost jobs on nthreads pthreads (jobs dealt round-robin, one private accumulator buffer per thread, summed
 * as integers: the same bits as the single-threaded lfo_render).  Returns the wall-clock seconds; acc_out (optional)
 * receives the W*H*3 sums. */
static double render_threads(const lfb_lens* L, const float* tex, int tw, int th, const lfb_light* lights, int n_lights,
                             const lfb_params* P, int nthreads, int64_t* acc_out, double* checksum) {
  if (nthreads < 1) nthreads = 1;
  const size_t npx = (size_t)P->width * P->height;
  job_t* J;
  int nj = list_jobs(L, P, n_lights, &J);
  worker_t* W = (worker_t*)calloc(nthreads, sizeof(worker_t));
  pthread_t* T = (pthread_t*)calloc(nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) {
    worker_t w = {L, tex, tw, th, lights, P, J, nj, t, nthreads, (int64_t*)calloc(3 * npx, sizeof(int64_t))};
    W[t] = w;
  }
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 1; t < nthreads; t++) pthread_create(&T[t], NULL, worker, &W[t]);
  worker(&W[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(T[t], NULL);
  for (int t = 1; t < nthreads; t++)
    for (size_t p = 0; p < 3 * npx; p++) W[0].acc[p] += W[t].acc[p];
  clock_gettime(CLOCK_MONOTONIC, &t1);
  double s = 0;
  for (size_t p = 0; p < 3 * npx; p++) s += (double)W[0].acc[p];
  if (checksum) *checksum = s;
  if (acc_out) memcpy(acc_out, W[0].acc, sizeof(int64_t) * 3 * npx);
  for (int t = 0; t < nthreads; t++) free(W[t].acc);
  free(W); free(T); free(J);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* Wall-clock of lfo_render's ghost jobs on nthreads pthreads: bench.py's cpu_baseline "port" leg (no reference counterpart). */
double lfo_time_render(const lfb_lens* L, const float* tex, int tw, int th, const lfb_light* lights,
                       int n_lights, const lfb_params* P, int nthreads, double* checksum) {
  return render_threads(L, tex, tw, th, lights, n_lights, P, nthreads, NULL, checksum);
}

/* lfo_render (grid modes) on nthreads pthreads, for the full-size parity tests: out = W*H*3 doubles, the same bits as
 * lfo_render's. */
int lfo_render_mt(const lfb_lens* L, const float* tex, int tw, int th, const lfb_light* lights, int n_lights,
                  const lfb_params* P, int nthreads, double* out) {
  if (P->mode != LFB_MODE_PARAXIAL_GRID && P->mode != LFB_MODE_EXACT_GRID) return LFB_ERR_INVALID;
  const size_t npx = (size_t)P->width * P->height;
  int64_t* acc = (int64_t*)malloc(sizeof(int64_t) * 3 * npx);
  if (!acc) return LFB_ERR_NOMEM;
  render_threads(L, tex, tw, th, lights, n_lights, P, nthreads, acc, NULL);
  const int bits = P->fixed_point_bits > 0 ? P->fixed_point_bits : 40;
  const double inv = ldexp(1.0, -bits);
  for (size_t p = 0; p < 3 * npx; p++) out[p] = (double)acc[p] * inv;
  free(acc);
  return LFB_OK;
}

/* ------------------------------------------------------------------------- */
/* path-traced scene pass, restated                             */
/* ------------------------------------------------------------------------- */
/* PathTracer::est_radiance_global_illumination (pathtracer.cpp:279-302) as the reference has it TODAY: zero_bounce_radiance
 * (:213-218, the surface's emission) + one_bounce_radiance (:220-231) = estimate_direct_lighting_importance (:136-211); the
 * indirect bounces are commented out there (:299).  For delta lights (DirectionalLight / PointLight, scene/light.cpp:11-24,
 * 49-60: one sample, pdf 1) the estimate is deterministic.  Primary rays: Camera::generate_ray (camera.cpp:278-305) through
 * the pixel centres.  Nearest hit by brute force over Triangle::intersect (scene/triangle.cpp:26-113: Moeller-Trumbore, hits
 * with t in [min_t, max_t], barycentrics in [0, 1], interpolated unit normal) and Sphere::intersect (scene/sphere.cpp:11-108);
 * the reference's BVH (scene/bvh.cpp) only changes the order of equal-t ties.  DiffuseBSDF::f = reflectance / pi,
 * EmissionBSDF::f = 0 (pathtracer/bsdf.cpp:58-61, 87-89).  Array layouts: oracle/ref_shim.cpp ref_scene_radiance.
 * PINNED against the compiled reference (tests/test_scene_pass.py, tests/golden/scene.npz). */
typedef struct { double o[3], d[3], min_t, max_t; } sray_t;
typedef struct { double t, n[3]; int mat; } shit_t;

static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static void unit3(const double* a, double* u) { /* vector3D.h:215-218: multiply by the reciprocal norm */
  double r = 1. / sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  u[0] = a[0] * r; u[1] = a[1] * r; u[2] = a[2] * r;
}

static int tri_hit(const double* P, const double* N, sray_t* r, shit_t* h) { /* triangle.cpp:26-113 */
  const double *p0 = P, *p1 = P + 3, *p2 = P + 6;
  double e1[3], e2[3], s[3], s1[3], s2[3];
  for (int a = 0; a < 3; a++) { e1[a] = p1[a] - p0[a]; e2[a] = p2[a] - p0[a]; s[a] = r->o[a] - p0[a]; }
  cross3(r->d, e2, s1);
  cross3(s, e1, s2);
  double den = dot3(s1, e1);
  double t = dot3(s2, e2) / den, b1 = dot3(s1, s) / den, b2 = dot3(s2, r->d) / den;
  if (t < r->min_t || t > r->max_t) return 0;
  if (b1 < 0 || b1 > 1) return 0;
  if (b2 < 0 || b2 > 1) return 0;
  if (b1 + b2 > 1) return 0;
  if (!(t == t) || !(b1 == b1) || !(b2 == b2)) return 0; /* NaN (a ray parallel to a degenerate configuration): not a hit here */
  if (!h) return 1;
  double b0 = 1 - b1 - b2, n[3];
  for (int a = 0; a < 3; a++) n[a] = b0 * N[a] + b1 * N[3 + a] + b2 * N[6 + a];
  r->max_t = t;
  h->t = t;
  unit3(n, h->n);
  return 1;
}

static int sph_hit(const double* S, sray_t* r, shit_t* h) { /* sphere.cpp:11-108 */
  double oc[3] = {r->o[0] - S[0], r->o[1] - S[1], r->o[2] - S[2]};
  double a = dot3(r->d, r->d), b = 2 * dot3(oc, r->d), c = dot3(oc, oc) - S[3] * S[3];
  double t1;
  if (b * b < 4.0 * a * c) return 0;
  if (b * b == 4.0 * a * c) {
    double root = (-b) / (2.0 * a);
    if (root < r->min_t || root > r->max_t) return 0;
    t1 = root;
  } else {
    double q = sqrt(b * b - 4.0 * a * c);
    double r1 = (-b - q) / (2.0 * a), r2 = (-b + q) / (2.0 * a);
    double lo = r1 < r2 ? r1 : r2, hi = r1 < r2 ? r2 : r1;
    if (lo > r->max_t || hi < r->min_t) return 0;
    if (lo < r->min_t) {
      if (hi > r->max_t) return 0;
      t1 = hi;
    } else {
      t1 = lo;
    }
  }
  if (!h) return 1;
  r->max_t = t1;
  h->t = t1;
  double p[3] = {r->o[0] + t1 * r->d[0] - S[0], r->o[1] + t1 * r->d[1] - S[1], r->o[2] + t1 * r->d[2] - S[2]};
  unit3(p, h->n);
  return 1;
}

typedef struct {
  const double *tri_pos, *tri_nrm, *sph; const int *tri_mat, *sph_mat; int nt, ns;
} sscene_t;

static int scene_hit(const sscene_t* S, sray_t* r, shit_t* h) {
  int hit = 0;
  for (int t = 0; t < S->nt; t++)
    if (tri_hit(S->tri_pos + 9 * t, S->tri_nrm + 9 * t, r, h)) { hit = 1; if (h) h->mat = S->tri_mat[t]; else return 1; }
  for (int k = 0; k < S->ns; k++)
    if (sph_hit(S->sph + 4 * k, r, h)) { hit = 1; if (h) h->mat = S->sph_mat[k]; else return 1; }
  return hit;
}

int lfo_scene_radiance(const double* tri_pos, const double* tri_nrm, const int* tri_mat, int nt, const double* sph,
                       const int* sph_mat, int ns, const double* mats, int nm, const double* lights, int nl,
                       const double* cam, int W, int H, double* out) {
  (void)nm;
  const sscene_t S = {tri_pos, tri_nrm, sph, tri_mat, sph_mat, nt, ns};
  const double pi = 3.14159265358979323846264338327950288; /* CGL's PI */
  const double ex = tan(0.5 * (cam[12] * (pi / 180.0))), ey = tan(0.5 * (cam[13] * (pi / 180.0)));
  const float eps_f = 0.00001f; /* EPS_F, CGL/misc.h:13 */
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      double* o = out + 3 * ((size_t)x + (size_t)y * W);
      o[0] = o[1] = o[2] = 0;
      /* Camera::generate_ray */
      double cx = ex * (2 * ((x + 0.5) / (double)W) - 1), cy = ey * (2 * ((y + 0.5) / (double)H) - 1);
      double dcam[3] = {cx, cy, -1}, du[3];
      unit3(dcam, du);
      sray_t r;
      for (int a = 0; a < 3; a++) {
        r.o[a] = cam[a];
        r.d[a] = cam[3 + 3 * a] * du[0] + cam[3 + 3 * a + 1] * du[1] + cam[3 + 3 * a + 2] * du[2];
      }
      r.min_t = cam[14]; r.max_t = cam[15];
      shit_t h;
      if (!scene_hit(&S, &r, &h)) continue; /* envLight == NULL: black */
      const double* M = mats + 6 * h.mat;
      const int emissive = M[3] > 0 || M[4] > 0 || M[5] > 0;
      double Lz[3] = {emissive ? M[3] : 0, emissive ? M[4] : 0, emissive ? M[5] : 0}; /* zero bounce */
      /* make_coord_space (bsdf.cpp:20-43) */
      double z[3] = {h.n[0], h.n[1], h.n[2]}, hh[3] = {h.n[0], h.n[1], h.n[2]}, xx[3], yy[3];
      if (fabs(hh[0]) <= fabs(hh[1]) && fabs(hh[0]) <= fabs(hh[2])) hh[0] = 1.0;
      else if (fabs(hh[1]) <= fabs(hh[0]) && fabs(hh[1]) <= fabs(hh[2])) hh[1] = 1.0;
      else hh[2] = 1.0;
      { double nz = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]); z[0] /= nz; z[1] /= nz; z[2] /= nz; }
      cross3(hh, z, yy);
      { double ny = sqrt(yy[0] * yy[0] + yy[1] * yy[1] + yy[2] * yy[2]); yy[0] /= ny; yy[1] /= ny; yy[2] /= ny; }
      cross3(z, yy, xx);
      { double nx = sqrt(xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2]); xx[0] /= nx; xx[1] /= nx; xx[2] /= nx; }
      double hit_p[3] = {r.o[0] + r.d[0] * h.t, r.o[1] + r.d[1] * h.t, r.o[2] + r.d[2] * h.t};
      double Ld[3] = {0, 0, 0};
      for (int l = 0; l < nl; l++) {
        const double* A = lights + 7 * l;
        double wi[3], dist;
        if (A[0] == 0) { /* DirectionalLight: dirToLight = -lightDir.unit() */
          double u[3];
          unit3(A + 4, u);
          wi[0] = -u[0]; wi[1] = -u[1]; wi[2] = -u[2];
          dist = INFINITY;
        } else { /* PointLight */
          double d[3] = {A[4] - hit_p[0], A[5] - hit_p[1], A[6] - hit_p[2]};
          unit3(d, wi);
          dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        }
        double wo[3] = {dot3(xx, wi), dot3(yy, wi), dot3(z, wi)}; /* w2o * wi */
        if (wo[2] < 0) continue;
        sray_t sh;
        for (int a = 0; a < 3; a++) { sh.o[a] = hit_p[a]; sh.d[a] = wi[a]; }
        sh.min_t = eps_f; sh.max_t = dist - eps_f;
        if (scene_hit(&S, &sh, NULL)) continue;
        double wu[3];
        unit3(wo, wu);
        if (!emissive)
          for (int c = 0; c < 3; c++) Ld[c] += (((1.0 / pi) * M[c]) * A[1 + c] * wu[2]) / 1.0;
      }
      for (int c = 0; c < 3; c++) o[c] = Lz[c] + (nl > 0 ? Ld[c] / (double)nl : 0.0);
    }
  return LFB_OK;
}
