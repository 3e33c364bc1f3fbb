// flare_pathtracer.hpp -- C++ host facade over the C ABI (include/lfb200.h) that keeps the reference's
// spellings for the ghost path, so that the reference application's call sites keep reading the same:
//
//   lfb::CameraApertureTexture::init(path)        <- CGL::CameraApertureTexture::init   src/pathtracer/camera.h:26-83
//   lfb::Camera::analyze_world_coord              <- CGL::Camera::analyze_world_coord   src/pathtracer/camera.cpp:245-273
//   lfb::DirectionalLight                         <- SceneObjects::DirectionalLight     src/scene/light.cpp:11-16
//   lfb::HDRImageBuffer                           <- CGL::HDRImageBuffer                src/util/image.h:105-242
//   lfb::PathTracer::set_frame_size               <- CGL::PathTracer::set_frame_size    src/pathtracer/pathtracer.cpp:66-69
//   lfb::PathTracer::find_sun_pos                 <- CGL::PathTracer::find_sun_pos      src/pathtracer/pathtracer.cpp:32-64
//   lfb::PathTracer::generate_ghost_buffer        <- CGL::PathTracer::generate_ghost_buffer  :714-762
//   lfb::PathTracer::ghost_buffer / axis_ray / angle_to_sun / flare_origins / flare_radiance   pathtracer.h:54, 131-135
//
// The facade owns no pixel or ray arithmetic: generate_ghost_buffer() hands the frame description to
// liblfb200.so, whose sm_100a kernels write straight into ghost_buffer.data (Vector3D[], stride 24).
// Errors: the reference prints and carries on; the facade throws lfb::Error (there is no CPU fallback to
// carry on with).  Header + flare_pathtracer.cpp + png_reader.cpp; link with -llfb200 -lz.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lfb200.h"

namespace lfb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

struct Vector2D {
  double x = 0, y = 0;
  Vector2D() {}
  Vector2D(double x_, double y_) : x(x_), y(y_) {}
};

struct Vector3D {  // CGL::Vector3D without AVX: three doubles, 24 bytes
  double x = 0, y = 0, z = 0;
  Vector3D() {}
  Vector3D(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};
static_assert(sizeof(Vector3D) == 24, "HDRImageBuffer::data must be tightly packed doubles");

struct Matrix3x3 {  // row-major here; CGL's is column-major, the accessors below hide the difference
  double m[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double operator()(int r, int c) const { return m[r][c]; }
  double& operator()(int r, int c) { return m[r][c]; }
};

// util/image.h:105-242
struct HDRImageBuffer {
  size_t w = 0, h = 0;
  std::vector<Vector3D> data;
  void resize(size_t w_, size_t h_) { w = w_; h = h_; data.assign(w * h, Vector3D()); }
  void clear() { data.clear(); w = h = 0; }
  void update_pixel_additive(const Vector3D& s, size_t x, size_t y) {
    Vector3D& p = data[x + y * w];
    p.x += s.x; p.y += s.y; p.z += s.z;
  }
  const Vector3D& get_pixel_value(size_t x, size_t y) const { return data[x + y * w]; }
};

// camera.h:18-88
class CameraApertureTexture {
 public:
  size_t width = 0, height = 0;
  std::vector<float> aperture;  // row-major y*width+x, red byte * float(1/255)
  double total_value = 0;
  int min_x = 0, min_y = 0, max_x = 0, max_y = 0;
  void init(const std::string& aperture_filename);            // PNG file (8-bit gray / RGB / RGBA / palette)
  void init_from_bytes(const unsigned char* red, size_t w, size_t h);
};

// scene/light.cpp:11-16 (note the sign flips the reference applies)
struct DirectionalLight {
  Vector3D radiance, posLight, dirToLight;
  DirectionalLight(const Vector3D& rad, const Vector3D& posLight, const Vector3D& lightDir);
};

// scene/light.h:49-59, light.cpp:47-48.  The reference's flare code skips point lights (pathtracer.cpp:35 only takes
// DirectionalLight); here one becomes an lfb_light with a finite distance.
struct PointLight {
  Vector3D radiance, position;
  PointLight(const Vector3D& rad, const Vector3D& pos) : radiance(rad), position(pos) {}
};

// the members of CGL::Camera the ghost path reads (camera.h:93-200)
class Camera {
 public:
  Matrix3x3 c2w;
  Vector3D pos;
  double hFov = 50, vFov = 35;
  CameraApertureTexture* aperture_texture = nullptr;
  CameraApertureTexture* ghost_aperture_texture = nullptr;
  void analyze_world_coord(const Vector3D& pos_world, double& ns_x, double& ns_y) const;
};

struct Scene {
  std::vector<DirectionalLight*> lights;
  std::vector<PointLight*> point_lights;
};

class PathTracer {
 public:
  explicit PathTracer(int device_id = -1);
  // Several GPUs of one node behind the same single-threaded call surface (lfb_create_multi): generate_ghost_buffer() shards
  // the grid modes' ghosts over them and every GPU writes its tiles into ghost_buffer.data.  REF_QUADS, the starburst and
  // render_frame run on the first device.
  explicit PathTracer(const std::vector<int>& device_ids);
  ~PathTracer();
  PathTracer(const PathTracer&) = delete;
  PathTracer& operator=(const PathTracer&) = delete;

  // --- the reference's members -------------------------------------------------------------
  Scene* scene = nullptr;
  Camera* camera = nullptr;
  HDRImageBuffer ghost_buffer;
  std::vector<Vector2D> flare_origins;
  std::vector<Vector3D> flare_radiance;
  std::vector<double> flare_distance;  // ours: lens units per flare origin, 0 = directional
  double scene_unit = 1000.0;          // ours: lens units (mm) per scene unit, for PointLight distances
  Vector2D axis_ray;
  float angle_to_sun = 0;
  void set_frame_size(size_t width, size_t height) { frame_w_ = width; frame_h_ = height; }
  double flare_radius = 30.0;     // pathtracer.h:93 (-n)
  double flare_intensity = 1.0;   // pathtracer.h:94 (-i)
  void find_sun_pos();
  void generate_ghost_buffer();
  // PathTracer::raytrace_starburst (pathtracer.cpp:947-1004) for the whole frame at once (the reference evaluates it per
  // pixel from raytrace_pixel :881).  additive = true adds into `target` like sampleBuffer.update_pixel(total + ghost +
  // starburst) (:891); camera->aperture_texture is the mask (-x).  Uses flare_origins / flare_radiance from find_sun_pos().
  void render_starburst(HDRImageBuffer& target, bool additive);
  // The displayable frame: [base] + ghosts + [starburst] -> HDRImageBuffer::toColor (util/image.h:208-223) -> 0xFFBBGGRR
  // pixels (ImageBuffer, util/image.h:53-62), composited and tone-mapped on the device.  base may be null.
  void render_frame(std::vector<uint32_t>& rgba8, const HDRImageBuffer* base, bool with_starburst, bool flip_vertical);

  // --- what the reference hard-codes, exposed ------------------------------------------------
  // Rendering mode of generate_ghost_buffer: LFB_MODE_REF_QUADS (default) reproduces the reference bit for
  // bit; the grid modes trace N x N ray bundles (see include/lfb200.h).
  lfb_params params;
  void set_lens(const lfb_lens& lens);  // default: the built-in prescription, RGB
  const lfb_lens& lens() const { return lens_; }
  // ghost_buffer stays allocated between renders of the same size and the facade tracks what the last frame left in it, so
  // a render clears only that instead of zero-filling W x H x 24 bytes (the reference's clear() + resize(), :719-720).
  // ghost_buffer must then only be written by generate_ghost_buffer(); call invalidate_ghost_buffer() after touching it.
  // Dirty-rectangle mode (REF_QUADS, whose ghosts are compact quads; grid modes when pinning is off): only the bounding
  // rectangle of the frame's deposits is cleared / converted / copied (lfb_render_ghosts_rect).
  bool dirty_rect_mode = true;
  // Grid modes: page-lock ghost_buffer's storage once (lfb_host_register) and render tile-sparse (lfb_render_ghosts_sparse):
  // the device writes this frame's dirty 16 x 16 tiles straight into ghost_buffer.data and re-zeroes the previous frame's.
  // The storage is unregistered when it moves, changes size, or the PathTracer dies; do not reallocate ghost_buffer.data
  // behind the PathTracer's back while this is on.
  bool pin_ghost_buffer = true;
  // Frames in flight (grid modes, one GPU; lfb_render_ghosts_sparse_begin / _end): a ring of ghost buffers for a host that
  // renders a sequence.  begin_ghost_frame(slot) is generate_ghost_buffer() for ghost_ring[slot] -- current flare_origins /
  // axis_ray from find_sun_pos(), page-locked on first use -- without the wait; end_ghost_frame(slot) blocks until that frame
  // is complete in ghost_ring[slot] and returns the 16 x 16 tiles it wrote.  A slot is collected before it is begun again;
  // ghost_ring[slot] must only be written through these calls (resize it, or the frame, and it starts clear again).
  HDRImageBuffer ghost_ring[LFB_SPARSE_SLOTS];
  void begin_ghost_frame(int slot);
  int end_ghost_frame(int slot);
  void invalidate_ghost_buffer() { buffer_state_ = kUnknown; }
  int last_tiles_written() const { return last_tiles_; }  // tiles the last sparse render wrote (-1: full-frame fallback)
  // stats of the last generate_ghost_buffer(): device ms of the trace kernels and of the whole call
  float last_trace_ms() const;
  float last_frame_ms() const;

 private:
  void ensure_engine();
  void* ring_pinned_[LFB_SPARSE_SLOTS] = {};
  size_t ring_bytes_[LFB_SPARSE_SLOTS] = {};
  bool ring_pending_[LFB_SPARSE_SLOTS] = {};
  lfb_engine* engine_ = nullptr;
  lfb_multi* multi_ = nullptr;
  std::vector<int> device_ids_;
  bool multi_lens_dirty_ = true;
  const CameraApertureTexture* multi_uploaded_ = nullptr;
  int device_id_;
  lfb_lens lens_;
  bool lens_dirty_ = true;
  const CameraApertureTexture* uploaded_ = nullptr;
  const CameraApertureTexture* uploaded_star_ = nullptr;
  std::vector<lfb_light> make_lights(bool for_ghosts) const;
  void upload_textures(bool ghost, bool star);
  size_t frame_w_ = 0, frame_h_ = 0;
  int dirty_[4] = {0, 0, -1, -1};  // kRectDirty: what the last frame wrote into ghost_buffer
  // what is known about ghost_buffer's content: all zeros / zeros outside dirty_ / the engine tracks the non-zero tiles of
  // the storage at tracked_ / anything (a full-frame render, or the application wrote into it)
  enum BufferState { kAllClear, kRectDirty, kTilesTracked, kUnknown };
  BufferState buffer_state_ = kUnknown;
  const void* tracked_ = nullptr;
  int last_tiles_ = 0;
  void* pinned_ = nullptr;         // ghost_buffer storage currently page-locked
  size_t pinned_bytes_ = 0;
  void pin_storage();
  void unpin_storage();
};

}  // namespace lfb
