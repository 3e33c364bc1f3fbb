/* oracle/lf_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement (C99, double precision) of the lens-flare ghost path, used as
 * the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * Parity status:
 *   - lfo_prescription / lfo_trace_ray_auto / lfo_generate_ghost_buffer restate
 *     the reference (src/pathtracer/pathtracer.cpp) and are PINNED bit-for-bit
 *     against the compiled reference (oracle/_ref, built by oracle/Makefile) and
 *     against tests/golden/ (vectors produced by tools/make_golden.py from it).
 *   - lfo_paraxial_system (PARAXIAL_GRID) is pinned against the reference's
 *     trace_ray_auto_before/after per ray (<= 1e-9 lens units).
 *   - the EXACT_GRID physics (sphere/plane intersection, vector Snell, Fresnel,
 *     quarter-wave coating) does not exist in the reference (its BSDF::refract /
 *     reflect / Fresnel are empty stubs, advanced_bsdf.cpp:52-169): PARITY
 *     UNPINNED by the reference; it is pinned by analytic invariants instead
 *     (tests/test_oracle_physics.py).
 * Struct definitions are shared with the product header include/lfb200.h.
 */
#ifndef LF_ORACLE_H
#define LF_ORACLE_H
#include "../include/lfb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* pathtracer.cpp:539-556 as data (+ Cauchy n(lambda) when n_lambda != 3). */
int lfo_builtin_lens(lfb_lens* lens, int n_lambda, float coating_lambda0_nm);

/* pathtracer.cpp:511-586: T_k, L_k, R_k(lambda) as (m00,m01,m10,m11) per surface. */
void lfo_prescription(const lfb_lens* lens, int lambda, double* t, double* l, double* r);

/* pathtracer.cpp:588-641 (which=0, trace_ray_auto_before) / :643-689 (which=1, _after). */
void lfo_trace_ray_auto(const lfb_lens* lens, int lambda, int which, float r, float theta,
                        int i, int j, double out[2]);

/* pathtracer.cpp:714-762 incl. draw_ghost :433-508, shift_vertex :412-430,
 * rasterize_textured_triangle :346-410, fill_textured_pixel :305-343.
 * out = W*H*3 doubles.  ghosts (optional, cap rows) receives the per-ghost records. */
int lfo_generate_ghost_buffer(const lfb_lens* lens, const float* tex, int tw, int th,
                              int W, int H, double axis_x, double axis_y, float angle_to_sun,
                              double* out, lfb_ref_ghost* ghosts, int cap);

/* Per-ghost ABCD system for PARAXIAL_GRID: n_cross stop crossings (entrance -> stop
 * plane, 2x2 each as m00,m01,m10,m11) and the entrance -> sensor matrix. */
int lfo_paraxial_system(const lfb_lens* lens, int lambda, int i, int j, int physical_backward,
                        double cross[3][4], double full[4]);

/* Trace the N x N grid of one ghost (i=j=-1: direct path); out has N*N records, row-major b*N+a. */
int lfo_trace_grid(const lfb_lens* lens, const float* tex, int tw, int th,
                   const lfb_light* light, const lfb_params* params, int i, int j, int lambda,
                   lfb_ray_hit* out);

/* Full frame (grid modes or REF_QUADS): out = W*H*3 doubles; accum (optional) = W*H*3 int64
 * fixed-point sums exactly as the engine keeps them. */
int lfo_render(const lfb_lens* lens, const float* tex, int tw, int th,
               const lfb_light* lights, int n_lights, const lfb_params* params,
               double* out, int64_t* accum);

/* Single-surface reflectance used by EXACT_GRID (Airy single-layer film, or bare
 * Fresnel when lambda0 = 0).  cos0 = cosine of the incidence angle in medium n0. */
double lfo_reflectance(double n0, double n2, double cos0, double coating_lambda0_nm,
                       double lambda_nm);

/* pathtracer.cpp:947-1004 (raytrace_starburst) restated per pixel, brute force like the reference, for the pixels
 * (xs[k], ys[k]).  fo = normalised flare origins (origin 0 drives the DFT phase, :918-934), rad = radiance per light.
 *   out_dft[k]      = pow(|sum A e^{j..}| / total_value, with suppression / amplification applied, 3 - flare_intensity)
 *                     -- the scalar the reference multiplies with the sum of the lights' radiance
 *   out_falloff[3k] = calculate_irradiance_falloff (:1030-1052) with the 16 random samples replaced by the 4x4 stratified
 *                     midpoints of the pixel (the reference draws from a process-global RNG: only its expectation is defined)
 * PINNED against the compiled reference: the DFT scalar exactly (ref_starburst_multi), the falloff statistically. */
int lfo_starburst_pixels(const float* tex, int tw, int th, int W, int H, int n_lights, const double* fo_xy,
                         const double* radiance3, double flare_radius, double flare_intensity, const int* xs,
                         const int* ys, int n, double* out_dft, double* out_falloff);

/* util/image.h:208-223 (toColor) + :53-62 (ImageBuffer::update_pixel): HDR doubles -> 0xAABBGGRR, alpha 0xFF. */
void lfo_to_color(const double* hdr, int W, int H, uint32_t* out);

/* Timed render on nthreads pthreads (jobs split round-robin); returns seconds. */
double lfo_time_render(const lfb_lens* lens, const float* tex, int tw, int th,
                       const lfb_light* lights, int n_lights, const lfb_params* params,
                       int nthreads, double* checksum);

/* lfo_render (grid modes) on nthreads pthreads: the same bits (integer sums), for the full-size parity tests. */
int lfo_render_mt(const lfb_lens* lens, const float* tex, int tw, int th, const lfb_light* lights, int n_lights,
                  const lfb_params* params, int nthreads, double* out);

/* The path-traced scene pass as the reference computes it today (pathtracer.cpp:279-302: emission + direct lighting by
 * light sampling; delta lights only -> deterministic), per pixel centre, brute-force nearest hit.  Layouts as
 * oracle/ref_shim.cpp ref_scene_radiance.  PINNED against the compiled reference. */
int lfo_scene_radiance(const double* tri_pos, const double* tri_nrm, const int* tri_mat, int nt, const double* sph,
                       const int* sph_mat, int ns, const double* mats, int nm, const double* lights, int nl,
                       const double* cam, int W, int H, double* out);

#ifdef __cplusplus
}
#endif
#endif
