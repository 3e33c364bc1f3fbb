/* lfb200.h -- C ABI of the B200-native lens-flare ghost engine (liblfb200.so).
 *
 * The reference (aatifjiwani/lens-flare) has no plugin/FFI boundary: its ghost
 * path is a set of C++ members and file-scope functions in namespace CGL.  This
 * header is the seam a maintainer cuts at: each
 * entry point names the reference interface it replaces (file:line relative to
 * the reference tree).  Plain pointers and sizes only; no C++/torch types.
 *
 * Threading: like the reference (raytraced_renderer.cpp:303-311 calls the path
 * once per render on the caller's thread) an engine is used by one thread at a
 * time.  All calls are blocking unless the name ends in _device/_async.
 * Errors: int status, 0 = LFB_OK, negative = failure; lfb_last_error() returns a
 * thread-local message.  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with LFB_ERR_NO_DEVICE.
 */
#ifndef LFB200_H
#define LFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFB_ABI_VERSION 3
#define LFB_MAX_SURFACES 16
#define LFB_MAX_LAMBDA 64
#define LFB_MAX_PEERS 16

typedef struct lfb_engine lfb_engine;

enum lfb_status {
  LFB_OK = 0,
  LFB_ERR_INVALID = -1,   /* bad argument */
  LFB_ERR_NO_DEVICE = -2, /* no CUDA device / driver: the engine never falls back to the CPU */
  LFB_ERR_CUDA = -3,      /* a CUDA runtime call failed (message has the CUDA error string) */
  LFB_ERR_STATE = -4,     /* lens or aperture not set yet */
  LFB_ERR_NOMEM = -5
};

/* Rendering modes.
 * REF_QUADS      bit-faithful PathTracer::generate_ghost_buffer (pathtracer.cpp:714-762):
 *                two marginal rays per ghost through the reference's ABCD chain
 *                (:588-689), ghost quad (:433-508) and textured-triangle raster (:305-410).
 * PARAXIAL_GRID  N x N ray bundle per ghost pushed through the SAME ABCD matrices
 *                (per ray, both meridional axes), aperture-mask lookup at every
 *                stop crossing, splat on the sensor.
 * EXACT_GRID     N x N ray bundle traced through the real prescription: sphere/plane
 *                intersection, vector Snell refraction with n(lambda), Fresnel +
 *                quarter-wave coating reflectance (physics absent from the reference,
 *                SURVEY.md 8a row a13), aperture lookup, splat. */
enum lfb_mode { LFB_MODE_REF_QUADS = 0, LFB_MODE_PARAXIAL_GRID = 1, LFB_MODE_EXACT_GRID = 2 };

/* Ghost pair sets.  REF13 = what generate_ghost_buffer draws (both reflections on the
 * same side of the stop, pathtracer.cpp:735-762); ALL = every i<j of non-stop surfaces
 * (28 for the built-in lens). */
enum lfb_pairset { LFB_PAIRS_REF = 0, LFB_PAIRS_ALL = 1 };

/* Output element of lfb_render_ghosts.  F64x3 with stride 24 or 32 is the in-memory
 * layout of the reference's HDRImageBuffer::data (util/image.h:239; sizeof(Vector3D)
 * is 24, or 32 when the reference is built with -mavx, CGL/include/CGL/vector3D.h:30-43). */
enum lfb_elem { LFB_F32x3 = 0, LFB_F64x3 = 1 };

/* Arithmetic of the grid modes.
 * LFB_FP32    the throughput kernels north_star asks for: FP32 geometry and weights (images <= 1e-3 relative L2 of the
 *             double oracle; per-ray sensor positions ~1e-5 RELATIVE, i.e. up to ~1e-2 lens units on the most defocused ghosts).
 * LFB_FP64    the oracle-order parity kernels: every operation in double, in the order of oracle/lf_oracle.c, so that the
 *             fixed-point sensor sums agree bit for bit.  Instruments, not throughput kernels.
 * LFB_STRICT  EXACT_GRID only: the throughput kernels with FP64 positions and directions (FP32 weights): per-ray sensor
 *             positions within 1e-5 lens units ABSOLUTE of the oracle (north_star's per-ray bar; measured ~1e-9). */
enum lfb_precision { LFB_FP32 = 0, LFB_FP64 = 1, LFB_STRICT = 2 };
enum lfb_splat { LFB_SPLAT_NEAREST = 0, LFB_SPLAT_BILINEAR = 1 };

/* Per-ray flags reported by lfb_dump_rays. */
enum lfb_ray_flags {
  LFB_RAY_OK = 0,
  LFB_RAY_MISSED = 1,      /* no intersection with a surface */
  LFB_RAY_VIGNETTED = 2,   /* outside a surface's clear semi-aperture */
  LFB_RAY_TIR = 4,         /* total internal reflection at a refracting surface */
  LFB_RAY_STOPPED = 8,     /* aperture mask was 0 (or outside the mask) at a stop crossing */
  LFB_RAY_OFF_SENSOR = 16  /* landed outside the W x H sensor */
};

/* The lens prescription: replaces the file-scope globals of pathtracer.cpp:539-556
 * (Ts, red/green/blue_refr, curvatures) and the constants 14.5 (:737), 11.6/-11.5
 * (:621-625).  ior[l][k] is the index AFTER surface k at wavelength l (the medium
 * before surface 0 is air, 1.0, as in create_Rs_for_color :560-566). */
typedef struct lfb_lens {
  int32_t n_surfaces;                       /* <= LFB_MAX_SURFACES */
  int32_t stop_index;                       /* surface index of the aperture stop */
  int32_t n_lambda;                         /* <= LFB_MAX_LAMBDA */
  int32_t reserved0;
  float curvature[LFB_MAX_SURFACES];        /* c = 1/R, 0 for planes */
  float thickness[LFB_MAX_SURFACES];        /* axial distance after surface k (last: to the sensor) */
  float semi_aperture[LFB_MAX_SURFACES];    /* clear radius of each element (EXACT_GRID vignetting) */
  float coating_lambda0_nm[LFB_MAX_SURFACES]; /* quarter-wave design wavelength; 0 = uncoated */
  float ior[LFB_MAX_LAMBDA][LFB_MAX_SURFACES];
  float lambda_nm[LFB_MAX_LAMBDA];
  float rgb_weight[LFB_MAX_LAMBDA][3];      /* how wavelength l adds into R,G,B */
  double entrance_half_height;              /* 14.5: half-side of the entrance ray grid */
  double stop_half_height;                  /* 11.6: the aperture mask spans [-h,h]^2 at the stop */
  double stop_half_height_neg;              /* 11.5: the reference's asymmetric re-aim height for r<0 */
} lfb_lens;

/* One light.  Replaces PathTracer::axis_ray / angle_to_sun / flare_radiance
 * (pathtracer.h:131-135), which find_sun_pos (pathtracer.cpp:32-64) fills for the
 * single sun of the reference.
 *
 * distance: 0 (a zero-initialised struct), negative or infinite = a directional light, the
 * only kind the reference's flares fire for (pathtracer.cpp:35): every ray of the bundle
 * enters along theta.  > 0 = a point light (the reference's PointLight, scene/light.h,
 * which its flare code ignores) that many lens units in front of the first surface
 * vertex, in the direction theta: the grid modes aim every entrance ray away from that
 * point, and EXACT_GRID weighs it by the inverse-square/cosine falloff relative to the
 * vertex, so distance -> infinity converges to the directional frame.  REF_QUADS and the
 * starburst do not use it (the reference has no counterpart there). */
typedef struct lfb_light {
  double ns_x, ns_y;  /* normalised screen position in [0,1]^2 (axis_ray) */
  float theta;        /* ray angle in the meridional plane, radians (angle_to_sun) */
  float radiance[3];
  double distance;    /* 0: directional; > 0: point light at this distance (lens units) */
} lfb_light;

typedef struct lfb_params {
  int32_t mode;            /* enum lfb_mode */
  int32_t pair_set;        /* enum lfb_pairset */
  int32_t include_direct;  /* grid modes: also trace the unreflected path */
  int32_t grid_n;          /* grid modes: N rays per axis per ghost */
  int32_t width, height;   /* sensor size in pixels (sampleBuffer.w/h, pathtracer.cpp:720) */
  int32_t precision;       /* enum lfb_precision (grid modes) */
  int32_t splat;           /* enum lfb_splat (grid modes) */
  int32_t fixed_point_bits; /* sensor accumulators are u64 fixed point with this many fractional bits; 0 -> 40 */
  int32_t physical_backward; /* PARAXIAL_GRID only: 0 = the reference's R^-1 on backward legs
                                (pathtracer.cpp:607-608), 1 = physically consistent backward refraction */
  int32_t shard_index, shard_count; /* this engine renders its share of the (light x pair x lambda) job list: whole
                                       (light, lambda) groups in contiguous blocks (a shard gets all wavelengths of a few
                                       lights) as far as they divide evenly among the shards, the jobs of the remaining
                                       groups one by one, longest first; 0,0 -> all */
  float px_per_unit;       /* sensor pixels per lens unit; 0 -> 0.4 (pathtracer.cpp:457-463) */
  int32_t physical_mapping; /* grid modes.  0 = the reference's screen mapping (draw_ghost / shift_vertex, pathtracer.cpp:412-430,
                               457-463): origin at the sun pixel, ghosts laid out along atan((ay-.5)/(ax-.5)) -- an angle mod pi, so
                               the ghosts of a sun left of centre are mirrored through the sun.  1 = the physical mapping: origin at
                               the image centre, the meridional axis along the light's azimuth atan2, a sensor point (xs, ys)
                               drawn at centre + ppu (xs u + ys v): with px_per_unit = W / (2 f tan(hFov/2)) the direct image of
                               the light lands on its own pixel (Camera::analyze_world_coord, camera.cpp:245-273). */
  float reserved[2];
} lfb_params;

/* Per-ray record of lfb_dump_rays (parity instrument). */
typedef struct lfb_ray_hit {
  double x_s, y_s;    /* sensor-plane position, lens units, local (meridional, sagittal) frame */
  double x_ap, y_ap;  /* position at the LAST stop crossing */
  double px, py;      /* continuous pixel position on the sensor */
  double weight;      /* scalar weight before the light's radiance and rgb_weight */
  uint32_t flags;     /* enum lfb_ray_flags */
  uint32_t pad;
} lfb_ray_hit;

/* Per-ghost record of lfb_ref_ghosts (REF_QUADS introspection). */
typedef struct lfb_ref_ghost {
  int32_t i, j, colour, pad;
  double r1, r2;          /* sensor heights of the +/-entrance marginal rays (trace_ray_auto_*) */
  float verts[4][2];      /* ul, ll, ur, lr screen-space vertices of draw_ghost */
  float scale, shift;
} lfb_ref_ghost;

/* Engine options (lfb_create_ex).  Zero-initialise, set struct_size = sizeof(lfb_options), change what you need: 0 always
 * means "the default".  No reference counterpart (the reference has no tunables on this path); the library reads NO
 * environment variables. */
typedef struct lfb_options {
  int32_t struct_size;        /* sizeof(lfb_options) of the caller (fields beyond it are taken as 0) */
  int32_t stream_priority;    /* 0: default priority; 1: highest -- for a second engine whose short kernels (finalize, reduce)
                                 must slip in between the CTAs of another engine's long trace kernel on the same device */
  int32_t kernel_select;      /* EXACT_GRID throughput kernels: 0 = by frame size (one job per ghost pair for small frames,
                                 ghost families from 16 384 family CTAs up), 1 = always per pair, 2 = always families */
  int32_t family_split;       /* > 0: at most this many forks per family job */
  int32_t ctas_per_sm;        /* build variant of the ghost / family kernels: 0 = default, 1 = the alternative register-allocation
                                 target, 2 = per-pair kernel without program staging (no CTA barrier; an experiment), 3 = per-pair kernel
                                 with the landing code out of line (an experiment) */
  int32_t prefix_overlap;     /* forward sweeps of frame k+1 overlap the ghost kernel of frame k: 0 = on, -1 = off */
  int32_t starburst_lattice;  /* starburst on the aperture's periodic lattice: 0 = when the frame is larger than the period, -1 = never */
  int32_t starburst_cache;    /* keep the lattice spectrum |F| between frames (it depends on the mask alone): 0 = on, -1 = off */
  int32_t reduce_ctas;        /* CTAs of the cross-GPU tile reduce (0 = one per SM; the one-GPU finalize always runs two per SM) and
                                 of the host-drain kernel of lfb_render_ghosts_sparse_begin / lfb_drain_tiles (0 = 16; at most 64) */
  int32_t collect_stats;      /* 1: run the counting instantiation of the EXACT_GRID kernels (lfb_exec_stats); slower */
  int64_t prefix_budget_bytes; /* device memory the cached forward sweeps may take: 0 = 40 GiB; < 0 = no cache */
  int32_t weights_table;      /* 1: Fresnel / coating weights from the 1024-interval tables for every ray (round 1's scheme,
                                 kept for A/B measurements) instead of the per-step polynomials */
  int32_t experiment;         /* bit mask of measurement switches that never change a frame's bits (tools/kernel_ab.py): 1 = look at the
                                 dirty-tile bytes in L2 (ld.global.cg) instead of through L1 (the default: 13 % faster at cfg2) */
  int32_t host_write_mbps;    /* lfb_render_ghosts_sparse_begin, lfb_drain_tiles: the pace, in MB/s, at which a frame's tiles are stored into host memory
                                 while other frames are in flight: 0 = 93 % of what unpaced stores reach on this link (measured once,
                                 at the first call); < 0 = unpaced (the next frame's kernels then wait for the stores: see sparse.cu) */
  int32_t reserved[5];
} lfb_options;

/* ---- lifecycle -------------------------------------------------------- */
/* LFB_ABI_VERSION of the library that was loaded (no reference counterpart: the reference is one binary). */
int lfb_abi_version(void);
/* Replaces `new PathTracer` (raytraced_renderer.cpp:58) for the ghost path.  One engine per CUDA device (one
 * process per GPU under torchrun; a C++ host may create several).  device_id < 0 -> current device.  Fails with
 * LFB_ERR_NO_DEVICE without an sm_100 GPU: there is no CPU fallback. */
int lfb_create(lfb_engine** out, int device_id);
/* lfb_create with options (NULL = defaults).  No reference counterpart. */
int lfb_create_ex(lfb_engine** out, int device_id, const lfb_options* options);
/* Replaces `delete pt` (raytraced_renderer.cpp:104). */
void lfb_destroy(lfb_engine* e);
/* The message of this thread's last failing call.  The reference has no error channel (it prints and goes on,
 * camera.h:40-44, or exit()s in its parsers); every int-returning entry point here returns 0 or a negative lfb_status. */
const char* lfb_last_error(void);

/* ---- inputs ----------------------------------------------------------- */
/* Fills `lens` with the reference's built-in prescription (pathtracer.cpp:539-556)
 * for n_lambda = 3 (the reference's R,G,B index tables) or, for other n_lambda,
 * wavelengths uniform on [400,700] nm with a piecewise two-term Cauchy n(lambda)
 * (linear in 1/lambda^2) through the R,G,B anchors at 650/550/450 nm.  coating_lambda0_nm > 0 coats every glass surface. */
int lfb_builtin_lens(lfb_lens* lens, int n_lambda, float coating_lambda0_nm);
/* Replaces the file-scope prescription and its derived matrices: Ts, red/green/blue_refr, curvatures, Ls, R_red/green/blue
 * (pathtracer.cpp:539-586) and the constants 14.5 / 11.6 / -11.5 (:737, :621-625).  The engine copies the struct. */
int lfb_set_lens(lfb_engine* e, const lfb_lens* lens);
/* Replaces Camera::ghost_aperture_texture (camera.h:175, filled by
 * CameraApertureTexture::init, camera.h:26-83): row-major y*w+x, values in [0,1]. */
int lfb_set_aperture(lfb_engine* e, const float* texels, int w, int h);

/* ---- the hot path ------------------------------------------------------ */
/* Replaces PathTracer::generate_ghost_buffer (pathtracer.cpp:714-762).  Renders all
 * ghosts of all lights into `out` (host memory; pageable or pinned), pixel (x,y) at
 * byte offset (x + y*width)*out_stride_bytes holding 3 floats or 3 doubles.
 * additive = 0 overwrites (the reference clears first, :719-720), 1 adds. */
int lfb_render_ghosts(lfb_engine* e, const lfb_light* lights, int n_lights,
                      const lfb_params* params, void* out, size_t out_stride_bytes,
                      int out_elem, int additive);

/* No reference counterpart (the reference renders one frame per call, raytraced_renderer.cpp:303-311).
 * lfb_render_ghosts without the wait (grid modes, overwrite semantics): the frame's device->host copy runs on a second
 * stream out of one of two device buffers and overlaps the next frame's trace.  `out` (pinned memory for a truly
 * asynchronous copy) is complete after lfb_sync(); alternate between two host buffers to keep two frames in flight. */
int lfb_render_ghosts_async(lfb_engine* e, const lfb_light* lights, int n_lights,
                            const lfb_params* params, void* out, size_t out_stride_bytes, int out_elem);

/* Dirty-rectangle form of lfb_render_ghosts.  A flare covers a small part of the sensor and the caller's buffer is
 * normally already clear (the reference clears ghost_buffer right before drawing, pathtracer.cpp:719-720; util/image.h:126-131),
 * so only the bounding rectangle of the pixels this frame deposits into is converted and copied back: pixels outside
 * rect_out are NOT touched, pixels inside are overwritten (zero where nothing landed).  rect_out = {x0, y0, x1, y1},
 * inclusive; {0, 0, -1, -1} when nothing landed. */
int lfb_render_ghosts_rect(lfb_engine* e, const lfb_light* lights, int n_lights,
                           const lfb_params* params, void* out, size_t out_stride_bytes,
                           int out_elem, int* rect_out);

/* Tile-sparse form of lfb_render_ghosts (grid modes, overwrite semantics).  A flare frame is ~99 % zeros: the splat kernels
 * mark the 16 x 16 sensor tiles they deposit into, and only the non-zero 8 x 8 quadrants of those tiles -- plus the quadrants the
 * PREVIOUS frame left non-zero in this same buffer, which are re-zeroed -- are converted and written, straight into the caller's memory from the device
 * (zero-copy over PCIe: `out` must be page-locked, lfb_host_alloc / lfb_host_register; pageable memory falls back to the
 * full-frame copy of lfb_render_ghosts and reports *tiles_written = -1).  On return `out` holds exactly this frame, bit for
 * bit what lfb_render_ghosts writes -- the reference's ghost_buffer after generate_ghost_buffer (pathtracer.cpp:714-762),
 * whose clear() + resize() (:719-720) is what the bookkeeping below replaces.
 * Contract: out_is_clear = 1 says every pixel of `out` is zero now (a fresh HDRImageBuffer::resize + clear); with 0, `out`
 * must be the buffer of this engine's previous lfb_render_ghosts_sparse call, same size / stride / element, not written by
 * anyone else since (else LFB_ERR_STATE).  n_lights = 0 just re-zeroes what the previous frame wrote. */
int lfb_render_ghosts_sparse(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params* params,
                             void* out, size_t out_stride_bytes, int out_elem, int out_is_clear, int* tiles_written);

/* lfb_render_ghosts_sparse with SEVERAL FRAMES IN FLIGHT (no reference counterpart: the reference renders and shows one frame
 * at a time, raytraced_renderer.cpp:303-311; this is for a host that keeps two or three ghost_buffers in rotation).  The tile
 * kernel's stores into host memory are PCIe-bound (~0.11 ms for cfg2's 5.4 MB) and the trace is SM-bound (~0.09 ms): in flight
 * together they cost the longer of the two, not the sum.  _begin(slot) enqueues the frame -- trace on the engine stream into the
 * slot's own accumulator, tile kernel on a second, highest-priority stream -- and returns; _end(slot) blocks until that frame's
 * pixels are in `out` and reports the tiles written.  slot is in [0, LFB_SPARSE_SLOTS); each slot remembers ITS `out` buffer
 * (same contract as above: out_is_clear = 1 the first time, then the same buffer, untouched by others); a slot must be
 * collected by _end before its next _begin (LFB_ERR_STATE).  Three slots keep the device busy while the host prepares the
 * next frame (cfg2 with the sun moving every frame, 0.23 ms per blocking call, 0.14 ms per frame in flight).  `out` must be page-locked and mapped
 * (LFB_ERR_INVALID otherwise: there is no staged fallback here).  The frames are bit for bit those of lfb_render_ghosts. */
#define LFB_SPARSE_SLOTS 4
int lfb_render_ghosts_sparse_begin(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params* params,
                                   void* out, size_t out_stride_bytes, int out_elem, int out_is_clear, int slot);
int lfb_render_ghosts_sparse_end(lfb_engine* e, int slot, int* tiles_written);
/* Device timeline of the slot's last collected frame, in ms since the engine was created (CUDA events; no reference
 * counterpart): ms[0] the engine stream reached the frame, ms[1] its trace kernels were done, ms[2] the tile kernel's stream
 * reached it, ms[3] its tiles were staged (pixels, in device memory), ms[4] they were in `out`.  For benches and
 * tools/e2e_probe.py. */
int lfb_sparse_slot_times(lfb_engine* e, int slot, float ms[5]);

/* Parity instrument (no reference counterpart; PARAXIAL_GRID records are what trace_ray_auto_before / _after,
 * pathtracer.cpp:588-689, return per axis): trace the N x N grid of one ghost (i, j, lambda) of one light
 * and return every ray's record (grid modes only).  i = j = -1 selects the direct path. */
int lfb_dump_rays(lfb_engine* e, const lfb_light* light, const lfb_params* params,
                  int i, int j, int lambda, lfb_ray_hit* out, size_t cap);

/* REF_QUADS introspection (no reference counterpart): the ghosts the device set up for the last REF_QUADS frame --
 * the trace_ray_auto_* results (pathtracer.cpp:738-757) and the draw_ghost vertices (:433-508).  Returns the count or <0. */
int lfb_ref_ghosts(lfb_engine* e, lfb_ref_ghost* out, int cap);

/* ---- starburst (the diffraction pattern of the aperture) ----------------- */
/* Replaces Camera::aperture_texture (camera.h:173, the -x PNG) for the starburst: texels as CameraApertureTexture::init
 * produces them; total_value and the bbox of texels > 0 (camera.h:61-73) are derived here. */
int lfb_set_starburst_aperture(lfb_engine* e, const float* texels, int w, int h);
/* Replaces PathTracer::raytrace_starburst (pathtracer.cpp:947-1004, called per pixel from raytrace_pixel :881) and
 * calculate_irradiance_falloff (:1030-1052) for the WHOLE frame: the per-pixel brute-force DFT of the mask becomes two
 * complex FP64 matrix products on the device.  lights[0] is flare_origins[0] (it alone drives the pattern, :919); every
 * light's radiance and falloff are summed like the reference does.  flare_radius / flare_intensity are the -n / -i
 * options (main.cpp:135-152).  The falloff's 16 random samples per pixel become the pixel's 4x4 stratified midpoints.
 * out: as lfb_render_ghosts (additive = 1 adds to the caller's sampleBuffer-like buffer, as raytrace_pixel :891 does). */
int lfb_render_starburst(lfb_engine* e, const lfb_light* lights, int n_lights, int width, int height,
                         double flare_radius, double flare_intensity, void* out, size_t out_stride_bytes,
                         int out_elem, int additive);

/* The displayable flare frame in one call: [base] + ghosts + [starburst] composited on the device as
 * raytrace_pixel does (pathtracer.cpp:881-891), tone-mapped by HDRImageBuffer::toColor (util/image.h:208-223) and packed
 * like ImageBuffer::update_pixel (util/image.h:53-62, 0xFFBBGGRR).  base_hdr: optional host W*H packed F64x3 (the path-traced
 * radiance); flare_radius < 0 skips the starburst; flip_vertical = 1 applies save_image's row flip
 * (raytraced_renderer.cpp:739-742).  Only 4 bytes per pixel come back over PCIe. */
int lfb_render_frame_rgba8(lfb_engine* e, const lfb_light* lights, int n_lights, const lfb_params* params,
                           double flare_radius, double flare_intensity, const double* base_hdr,
                           uint32_t* out_rgba8, int flip_vertical);

/* ---- the path-traced scene pass (BASELINE config 5) ------- */
/* A scene as plain arrays (the reference's COLLADA loader, GUI and tile pool are out of scope; the application hands over what
 * it loaded).  Replaces what PathTracer::set_scene / build_accel give raytrace_pixel: the primitives of Scene::objects
 * (Triangle, scene/triangle.h; Sphere, scene/sphere.h), their BSDFs (DiffuseBSDF / EmissionBSDF, pathtracer/bsdf.h) and
 * Scene::lights (DirectionalLight / PointLight, scene/light.h; area lights are sampled randomly in the reference and are not
 * supported).  The engine copies everything and builds its own BVH. */
typedef struct lfb_scene {
  const double* tri_pos;    /* [n_tri][3][3] vertex positions (Triangle::p1, p2, p3) */
  const double* tri_nrm;    /* [n_tri][3][3] vertex normals (Triangle::n1, n2, n3) */
  const int32_t* tri_mat;   /* [n_tri] material index */
  int32_t n_tri, n_sph;
  const double* spheres;    /* [n_sph][4] centre, radius */
  const int32_t* sph_mat;   /* [n_sph] */
  const double* materials;  /* [n_mat][6] reflectance rgb, emission rgb: any emission > 0 -> EmissionBSDF, else DiffuseBSDF */
  const double* lights;     /* [n_lights][7] kind (0 directional: vec = the direction the light travels; 1 point: vec = position),
                               radiance rgb, vec xyz */
  int32_t n_mat, n_lights;
} lfb_scene;
/* The pinhole camera of Camera::generate_ray (camera.cpp:278-305): position, camera-to-world rotation (row-major; the camera
 * looks down its -z), fields of view in degrees, clip distances. */
typedef struct lfb_camera {
  double pos[3];
  double c2w[9];
  double hfov_deg, vfov_deg, nclip, fclip;
} lfb_camera;
int lfb_set_scene(lfb_engine* e, const lfb_scene* scene);
/* Replaces the path-traced term of PathTracer::raytrace_pixel (pathtracer.cpp:819-899) for the WHOLE frame:
 * est_radiance_global_illumination (:279-302) as the reference has it -- the hit surface's emission plus direct lighting by
 * light sampling with shadow rays (:136-231; the indirect bounces are commented out there) -- along the camera ray through
 * every pixel CENTRE (the reference jitters ns_aa samples with its global RNG).  out: as lfb_render_ghosts; additive = 1 adds
 * into the caller's sampleBuffer-like frame. */
int lfb_render_scene(lfb_engine* e, const lfb_camera* camera, int width, int height, void* out, size_t out_stride_bytes,
                     int out_elem, int additive);
/* BASELINE config 5 in one call: scene pass + ghosts + [starburst] composited on the device exactly as raytrace_pixel sums them
 * (:881-891), then HDRImageBuffer::toColor and the RGBA8 pack (as lfb_render_frame_rgba8, whose other arguments these are). */
int lfb_render_composite_rgba8(lfb_engine* e, const lfb_camera* camera, const lfb_light* lights, int n_lights,
                               const lfb_params* params, double flare_radius, double flare_intensity, uint32_t* out_rgba8,
                               int flip_vertical);

/* ---- device-resident API (multi-GPU sharding, benchmarking) ------------- */
/* No reference counterparts in this section: the reference is a single-process CPU program.  Together these calls are
 * generate_ghost_buffer (pathtracer.cpp:714-762) split into its device stages so that one process per GPU can shard a
 * frame by (light, ghost pair, wavelength) and sum the shards. */
/* A sensor accumulator buffer, owned by the caller: width*height*3 u64 fixed-point sums followed by the dirty-tile map
 * of those sums (one byte per 16 x 16 pixel tile, set by the splat kernels).  Zero it once; the tile-sparse calls below
 * leave it zeroed again. */
size_t lfb_accum_bytes(int width, int height);
/* The state that goes with an OUTPUT buffer of the tile-sparse calls (which tiles the previous frame left non-zero in it).
 * Zero it once, while the output buffer is all zeros. */
size_t lfb_tile_state_bytes(int width, int height);
/* Tile-sparse lfb_finalize_device + accumulator clear in one launch: converts the dirty tiles of accum_dev into out_dev
 * (a full W x H frame in device memory, or page-locked host memory as mapped on the device), re-zeroes the tiles the previous
 * frame left in out_dev, zeroes the accumulator tiles it read and rolls the tile maps over.  out_dev then holds exactly the
 * frame lfb_finalize_device would write, given that it was all zeros when tile_state_dev was. */
int lfb_finalize_tiles_device(lfb_engine* e, void* accum_dev, const lfb_params* params, void* out_dev, size_t out_stride_bytes,
                              int out_elem, void* tile_state_dev);
/* Multi-GPU form of the same launch: the cross-GPU reduce of SURVEY.md 8e fused with finalize, over NVLink peer memory, on
 * dirty tiles only.  accum_ptrs[r] is rank r's accumulator buffer as mapped in this process; this rank handles the tile-map
 * words w = rank (mod n_ranks): it sums the tiles that are dirty on any rank (reading only the ranks that have them),
 * zeroes them, and stores the pixels into out_dev, the OWNER's output frame as mapped here.  tile_state_dev is this rank's
 * own state for that frame.  Order it after every rank's lfb_render_ghosts_device of the frame (lfb_peer_barrier, or
 * events in a single-process host).  Integer sums: the frame has the same bits for any rank count. */
int lfb_reduce_tiles_peers(lfb_engine* e, void* const* accum_ptrs, int n_ranks, int rank, const lfb_params* params,
                           void* out_dev, size_t out_stride_bytes, int out_elem, void* tile_state_dev);
/* The same reduce for an output frame in page-locked HOST memory that is written while other frames are in flight (DESIGN.md
 * 5b: stores into host memory must be paced, or the GPU cannot fetch its next commands while they drain).  Two launches:
 * lfb_reduce_tiles_peers_staged leaves this rank's share of the frame's dirty tiles as pixels in stage_dev (device memory,
 * lfb_tile_stage_bytes; only partial tiles at the frame's edge go to out_dev directly); lfb_drain_tiles -- on another engine's
 * stream, ordered after it by the caller -- copies them into out_dev with a few CTAs paced just under the link rate
 * (lfb_options.host_write_mbps).  tile_state_dev and stage_dev must not be reused before the drain has run. */
size_t lfb_tile_stage_bytes(int width, int height, size_t out_stride_bytes);
int lfb_reduce_tiles_peers_staged(lfb_engine* e, void* const* accum_ptrs, int n_ranks, int rank, const lfb_params* params,
                                  void* out_dev, size_t out_stride_bytes, int out_elem, void* tile_state_dev, void* stage_dev);
int lfb_drain_tiles(lfb_engine* e, const lfb_params* params, void* out_dev, size_t out_stride_bytes, const void* tile_state_dev,
                    const void* stage_dev);
/* The engine's CUDA stream (cudaStream_t) so callers can order work / record events. */
void* lfb_stream(lfb_engine* e);
/* Zero the accumulators, trace + splat this shard's ghosts (grid modes), all on the
 * engine stream; returns without synchronising.  accum_dev is a device pointer. */
int lfb_render_ghosts_device(lfb_engine* e, const lfb_light* lights, int n_lights,
                             const lfb_params* params, void* accum_dev, int clear_first);
/* accum (u64 fixed point) -> packed F32x3 / F64x3 pixels in device memory. */
int lfb_finalize_device(lfb_engine* e, const void* accum_dev, const lfb_params* params,
                        void* out_dev, size_t out_stride_bytes, int out_elem);
/* The same, and the accumulators are left zeroed (cleared as they are read): one launch instead of finalize +
 * 24 B/pixel memset before the buffer's next frame (lfb_render_ghosts_device, clear_first = 0).  For serial hosts;
 * a host that overlaps this with the next frame's trace is better off with the separate memset (measured). */
int lfb_finalize_clear_device(lfb_engine* e, void* accum_dev, const lfb_params* params,
                              void* out_dev, size_t out_stride_bytes, int out_elem);
/* Wait for everything enqueued on the engine's streams (what the blocking calls do before they return). */
int lfb_sync(lfb_engine* e);
/* Device-side barrier across the ranks of one node, enqueued on lfb_stream: flag_ptrs[r] is rank r's flag array
 * (LFB_MAX_PEERS zero-initialised u64, symmetric / IPC memory) as mapped in this process; epoch must increase by one per
 * barrier.  Work enqueued before it on every rank is complete and visible to work enqueued after it on any rank. */
int lfb_peer_barrier(lfb_engine* e, void* const* flag_ptrs, int n_ranks, int rank, uint64_t epoch);
/* Multi-GPU: fused reduce + finalize over NVLink peer memory (replaces an NCCL reduce followed by lfb_finalize_device).
 * accum_ptrs[r] is rank r's accumulator buffer AS MAPPED IN THIS PROCESS (CUDA IPC / symmetric memory); multicast_accum,
 * when not NULL, is the same buffer's NVSwitch multicast address (the switch performs the additions).  This rank converts
 * the pixels [rank*npx/n, (rank+1)*npx/n) and stores them into out_dev, the OWNER rank's output buffer as mapped here.
 * The caller orders this call after every rank's lfb_render_ghosts_device (lfb_peer_barrier) and reads
 * the owner's buffer after another one.  Integer sums: the frame has the same bits for any rank count. */
int lfb_reduce_finalize_peers(lfb_engine* e, const void* const* accum_ptrs, int n_ranks, int rank,
                              const void* multicast_accum, const lfb_params* params, void* out_dev,
                              size_t out_stride_bytes, int out_elem);

/* ---- single-process multi-GPU ------------------------- */
/* The reference's caller is ONE thread of ONE process (RaytracedRenderer::start_raytracing, raytraced_renderer.cpp:303-311):
 * this is the multi-GPU frame for that caller -- no torchrun, no NCCL, no IPC.  lfb_create_multi opens one engine per device
 * of one node and enables peer access between them; lfb_render_ghosts_multi shards the (light x pair x lambda) jobs over the
 * devices (lfb_params.shard_* are set internally), every device traces its share into its own accumulators, and then every
 * device runs the tile-sparse reduce (lfb_reduce_tiles_peers) for its interleaved share of the dirty tiles: it sums them over
 * NVLink peer memory and writes the pixels STRAIGHT INTO THE CALLER'S page-locked buffer over its own PCIe link.  Ordering is
 * by CUDA events between the devices' streams.  Same contract and result as lfb_render_ghosts_sparse (out_is_clear,
 * tiles_written; bit-identical frames for any device count); pageable `out` falls back to a full-frame copy from device 0. */
typedef struct lfb_multi lfb_multi;
int lfb_create_multi(lfb_multi** out, const int* device_ids, int n_devices, const lfb_options* options);
void lfb_destroy_multi(lfb_multi* m);
/* lfb_set_lens / lfb_set_aperture on every device. */
int lfb_multi_set_lens(lfb_multi* m, const lfb_lens* lens);
int lfb_multi_set_aperture(lfb_multi* m, const float* texels, int w, int h);
int lfb_render_ghosts_multi(lfb_multi* m, const lfb_light* lights, int n_lights, const lfb_params* params, void* out,
                            size_t out_stride_bytes, int out_elem, int out_is_clear, int* tiles_written);
/* Device times of the last lfb_render_ghosts_multi, ms: stage_ms[0] = slowest device's trace kernels, [1] = slowest device's
 * reduce kernel, [2] = first enqueue to last completion as seen by the host thread, [3] = host time spent enqueuing. */
int lfb_multi_stats(lfb_multi* m, float stage_ms[4], int* n_devices);

/* ---- accounting (no reference counterparts) ------------------------------ */
/* Ray-surface interactions (I(i,j) = 2(j-i) + n_surfaces + 1 per ray,
 * n_surfaces + 1 for the direct path) and rays this shard traces for one frame. */
int lfb_count_work(const lfb_lens* lens, const lfb_params* params, int n_lights,
                   double* rays, double* interactions, int* jobs);
/* The (light, i, j, lambda) jobs this shard renders, in launch order (host-only; no
 * device needed).  out holds cap rows of 4 ints; returns the job count (may exceed cap). */
int lfb_list_jobs(const lfb_lens* lens, const lfb_params* params, int n_lights,
                  int32_t* out, int cap);
/* Kernels launched by this engine since creation, and the device time (ms, CUDA
 * events on the engine stream) of the trace/splat and raster kernels of the last frame. */
int lfb_stats(lfb_engine* e, uint64_t* kernel_launches, float* last_trace_ms, float* last_frame_ms);
/* What the EXACT_GRID throughput kernels actually executed for the last frame rendered by an engine created with
 * options.collect_stats = 1 (the counting instantiation; waits for the engine's stream): out[0] = surface steps executed
 * (one ray reaching one surface or the sensor; mirror pairs, the shared forward sweep and early deaths make this ~10x
 * fewer than the NOMINAL interactions lfb_count_work reports), out[1] = ray pairs started, out[2] = ray pairs landed.
 * out[3] = 1 when the frame ran the family kernel, 0 for the per-pair kernel. */
int lfb_exec_stats(lfb_engine* e, uint64_t out[4]);

/* Live roofline denominators for the scalar pipes the trace is bound by: FP32 FLOP/s (FMA = 2)
 * and MUFU op/s of this GPU measured by two register-only micro-kernels, plus the SM clock
 * (Hz) seen while they ran. */
int lfb_probe_peaks(lfb_engine* e, double* fp32_flops, double* mufu_ops, double* sm_clock_hz);

/* ---- pinned host memory helpers (no reference counterparts) -------------- */
/* cudaHostAlloc / cudaFreeHost for callers that stage frames themselves; NULL on failure. */
void* lfb_host_alloc(size_t bytes);
void lfb_host_free(void* p);
/* Page-lock memory the caller already owns -- e.g. the storage of the reference's
 * HDRImageBuffer::data (std::vector<Vector3D>, util/image.h:239) -- so that lfb_render_ghosts
 * copies the frame back at PCIe rate instead of through the driver's pageable staging.
 * One-time cost of the order of the copy itself; undo before the memory is freed or
 * reallocated.  Returns LFB_OK or LFB_ERR_CUDA (the memory then simply stays pageable). */
int lfb_host_register(void* p, size_t bytes);
int lfb_host_unregister(void* p);
/* The address at which the current device sees page-locked host memory (cudaHostGetDevicePointer): what the device-resident
 * calls (lfb_finalize_tiles_device, lfb_reduce_tiles_peers) take as out_dev to write a frame's tiles straight into host
 * memory.  NULL when p is not page-locked / mapped. */
void* lfb_host_device_pointer(void* p);

#ifdef __cplusplus
}
#endif
#endif /* LFB200_H */
