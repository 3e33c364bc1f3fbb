#include "flare_pathtracer.hpp"

#include <cmath>
#include <cstring>

#include "png_reader.hpp"

namespace lfb {
namespace {
void check(int rc, const char* what) {
  if (rc < 0) throw Error(rc, std::string(what) + ": " + lfb_last_error());
}
const double kPi = 3.14159265358979323846;
}  // namespace

// ---- CameraApertureTexture (camera.h:26-83) -------------------------------------------------
void CameraApertureTexture::init(const std::string& aperture_filename) {
  std::vector<unsigned char> red;
  unsigned w = 0, h = 0;
  read_png_red(aperture_filename, red, w, h);  // throws: the reference prints "UNABLE TO LOAD" and then indexes an empty vector
  init_from_bytes(red.data(), w, h);
}

void CameraApertureTexture::init_from_bytes(const unsigned char* red, size_t w, size_t h) {
  width = w; height = h;
  aperture.resize(w * h);
  const float inv = 1.0 / 255.0;  // Color(const uchar*), CGL/src/color.cpp:16-21
  min_x = min_y = (int)w; max_x = max_y = -1;
  total_value = 0.0;
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) {
      const float v = red[y * w + x] * inv;
      aperture[y * w + x] = v;
      total_value += v;
      if (v > 0) {
        if ((int)x < min_x) min_x = (int)x;
        if ((int)y < min_y) min_y = (int)y;
        if ((int)x > max_x) max_x = (int)x;
        if ((int)y > max_y) max_y = (int)y;
      }
    }
}

// ---- DirectionalLight (light.cpp:11-16) ------------------------------------------------------
DirectionalLight::DirectionalLight(const Vector3D& rad, const Vector3D& posLight_, const Vector3D& lightDir) : radiance(rad) {
  posLight = Vector3D(-posLight_.x, -posLight_.y, -posLight_.z);
  const double n = std::sqrt(lightDir.x * lightDir.x + lightDir.y * lightDir.y + lightDir.z * lightDir.z);
  dirToLight = Vector3D(-lightDir.x / n, -lightDir.y / n, -lightDir.z / n);
}

// ---- Camera::analyze_world_coord (camera.cpp:245-273) ------------------------------------------
void Camera::analyze_world_coord(const Vector3D& p, double& ns_x, double& ns_y) const {
  const double edge_x = std::tan(0.5 * (hFov * (kPi / 180.0)));
  const double edge_y = std::tan(0.5 * (vFov * (kPi / 180.0)));
  const double dx = p.x - pos.x, dy = p.y - pos.y, dz = p.z - pos.z;
  // c2w.T() * (p - pos)
  const double cx = c2w(0, 0) * dx + c2w(1, 0) * dy + c2w(2, 0) * dz;
  const double cy = c2w(0, 1) * dx + c2w(1, 1) * dy + c2w(2, 1) * dz;
  const double cz = c2w(0, 2) * dx + c2w(1, 2) * dy + c2w(2, 2) * dz;
  const double a = std::fabs(cz);
  ns_x = ((cx / a / edge_x) + 1) / 2.0;
  ns_y = ((cy / a / edge_y) + 1) / 2.0;
}

// ---- PathTracer ----------------------------------------------------------------------------
PathTracer::PathTracer(int device_id) : device_id_(device_id) {
  std::memset(&params, 0, sizeof(params));
  params.mode = LFB_MODE_REF_QUADS;
  params.pair_set = LFB_PAIRS_REF;
  params.grid_n = 256;
  params.precision = LFB_FP32;
  params.splat = LFB_SPLAT_BILINEAR;
  check(lfb_builtin_lens(&lens_, 3, 0.f), "lfb_builtin_lens");
}

PathTracer::PathTracer(const std::vector<int>& device_ids) : PathTracer(device_ids.empty() ? -1 : device_ids[0]) {
  if (device_ids.size() > 1) device_ids_ = device_ids;
}

PathTracer::~PathTracer() {
  for (int s = 0; s < LFB_SPARSE_SLOTS; s++)
    if (ring_pending_[s] && engine_) lfb_render_ghosts_sparse_end(engine_, s, nullptr);  // nothing may still be writing into the ring
  unpin_storage();
  lfb_destroy_multi(multi_);
  lfb_destroy(engine_);
  for (int s = 0; s < LFB_SPARSE_SLOTS; s++)
    if (ring_pinned_[s]) lfb_host_unregister(ring_pinned_[s]);
}

void PathTracer::begin_ghost_frame(int slot) {
  if (slot < 0 || slot >= LFB_SPARSE_SLOTS) throw Error(LFB_ERR_INVALID, "begin_ghost_frame: slot out of range");
  if (params.mode == LFB_MODE_REF_QUADS) throw Error(LFB_ERR_STATE, "begin_ghost_frame: grid modes only (REF_QUADS frames are compact quads: generate_ghost_buffer)");
  if (!device_ids_.empty()) throw Error(LFB_ERR_STATE, "begin_ghost_frame: one GPU only");
  if (ring_pending_[slot]) throw Error(LFB_ERR_STATE, "begin_ghost_frame: the slot's previous frame was not collected (end_ghost_frame)");
  HDRImageBuffer& B = ghost_ring[slot];
  const size_t bytes = frame_w_ * frame_h_ * sizeof(Vector3D);
  const bool same = B.w == frame_w_ && B.h == frame_h_ && B.data.size() == frame_w_ * frame_h_ && frame_w_ > 0 &&
                    static_cast<void*>(B.data.data()) == ring_pinned_[slot] && bytes == ring_bytes_[slot];
  int clear = 0;
  if (!same) {  // (re)allocate, zero, page-lock: the engine's tile bookkeeping for this slot starts over
    if (ring_pinned_[slot]) lfb_host_unregister(ring_pinned_[slot]);
    ring_pinned_[slot] = nullptr;
    B.clear();
    B.resize(frame_w_, frame_h_);
    if (B.data.empty()) throw Error(LFB_ERR_INVALID, "begin_ghost_frame: set_frame_size first");
    ensure_engine();
    check(lfb_host_register(B.data.data(), bytes), "lfb_host_register");
    ring_pinned_[slot] = B.data.data();
    ring_bytes_[slot] = bytes;
    clear = 1;
  }
  const bool sun = !(axis_ray.x == 0 && axis_ray.y == 0);
  if (sun) upload_textures(true, false);
  else ensure_engine();
  params.width = (int)frame_w_;
  params.height = (int)frame_h_;
  std::vector<lfb_light> lights = sun ? make_lights(true) : std::vector<lfb_light>();
  check(lfb_render_ghosts_sparse_begin(engine_, lights.data(), (int)lights.size(), &params, B.data.data(), sizeof(Vector3D), LFB_F64x3, clear, slot),
        "lfb_render_ghosts_sparse_begin");
  ring_pending_[slot] = true;
}

int PathTracer::end_ghost_frame(int slot) {
  if (slot < 0 || slot >= LFB_SPARSE_SLOTS) throw Error(LFB_ERR_INVALID, "end_ghost_frame: slot out of range");
  if (!ring_pending_[slot]) throw Error(LFB_ERR_STATE, "end_ghost_frame: no frame in flight in this slot");
  int tiles = 0;
  check(lfb_render_ghosts_sparse_end(engine_, slot, &tiles), "lfb_render_ghosts_sparse_end");
  ring_pending_[slot] = false;
  return tiles;
}

void PathTracer::unpin_storage() {
  if (pinned_) lfb_host_unregister(pinned_);  // best effort: a failure leaves nothing to undo
  pinned_ = nullptr;
  pinned_bytes_ = 0;
}

void PathTracer::pin_storage() {
  void* p = ghost_buffer.data.empty() ? nullptr : static_cast<void*>(ghost_buffer.data.data());
  const size_t bytes = ghost_buffer.data.size() * sizeof(Vector3D);
  if (p == pinned_ && bytes == pinned_bytes_) return;
  unpin_storage();
  if (p && lfb_host_register(p, bytes) == LFB_OK) { pinned_ = p; pinned_bytes_ = bytes; }  // else: stays pageable, still correct
}

void PathTracer::set_lens(const lfb_lens& lens) {
  lens_ = lens;
  lens_dirty_ = true;
  multi_lens_dirty_ = true;
}

void PathTracer::ensure_engine() {
  if (!engine_) check(lfb_create(&engine_, device_id_), "lfb_create");
  if (lens_dirty_) {
    check(lfb_set_lens(engine_, &lens_), "lfb_set_lens");
    lens_dirty_ = false;
  }
}

// pathtracer.cpp:32-64: every on-screen directional light is recorded; axis_ray / angle_to_sun describe the last one
void PathTracer::find_sun_pos() {
  if (!scene || !camera) return;
  for (DirectionalLight* light : scene->lights) {
    double ns_x, ns_y;
    camera->analyze_world_coord(light->posLight, ns_x, ns_y);
    if ((ns_x >= 0 && ns_x <= 1) && (ns_y >= 0 && ns_y <= 1)) {
      flare_origins.emplace_back(ns_x, ns_y);
      flare_radiance.push_back(light->radiance);
      flare_distance.push_back(0.0);
      angle_to_sun = (float)std::atan(ns_y / ns_x);
      axis_ray = Vector2D(ns_x, ns_y);
    }
  }
  for (PointLight* light : scene->point_lights) {  // ours: same projection, finite distance
    double ns_x, ns_y;
    camera->analyze_world_coord(light->position, ns_x, ns_y);
    if ((ns_x >= 0 && ns_x <= 1) && (ns_y >= 0 && ns_y <= 1)) {
      const double dx = light->position.x - camera->pos.x, dy = light->position.y - camera->pos.y, dz = light->position.z - camera->pos.z;
      flare_origins.emplace_back(ns_x, ns_y);
      flare_radiance.push_back(light->radiance);
      flare_distance.push_back(std::sqrt(dx * dx + dy * dy + dz * dz) * scene_unit);
      angle_to_sun = (float)std::atan(ns_y / ns_x);
      axis_ray = Vector2D(ns_x, ns_y);
    }
  }
}

// pathtracer.cpp:714-762.  The reference clears and resizes ghost_buffer on every call (:719-720; 18.5 ms of zero fill at
// 1080p) and then draws.  Here ghost_buffer stays allocated while the frame size is unchanged and the
// facade keeps track of what is non-zero in it (buffer_state_), so that each path clears only what the previous frame wrote:
//   grid modes (PARAXIAL / EXACT), pinned   lfb_render_ghosts_sparse: the engine re-zeroes the previous frame's 16 x 16 tiles
//                                           and writes this frame's, from the device straight into ghost_buffer.data
//   REF_QUADS (or grid modes with pinning off), dirty_rect_mode
//                                           lfb_render_ghosts_rect: the previous frame's rectangle is cleared on the host
//   otherwise                               lfb_render_ghosts: every pixel is rewritten
// Whatever path the PREVIOUS call took, the buffer is first brought to the state the new path needs -- a buffer of unknown
// content is cleared as the reference does.
void PathTracer::generate_ghost_buffer() {
  const bool same = ghost_buffer.w == frame_w_ && ghost_buffer.h == frame_h_ && ghost_buffer.data.size() == frame_w_ * frame_h_ && frame_w_ > 0;
  if (!same) {
    unpin_storage();  // the vector may reallocate
    ghost_buffer.clear();
    ghost_buffer.resize(frame_w_, frame_h_);
    buffer_state_ = kAllClear;
  }
  const bool sun = !(axis_ray.x == 0 && axis_ray.y == 0);
  const bool grid = params.mode != LFB_MODE_REF_QUADS;
  enum { kSparse, kRect, kDense } path = (grid && pin_ghost_buffer) ? kSparse : (dirty_rect_mode ? kRect : kDense);
  auto clear_all = [&] {
    if (!ghost_buffer.data.empty()) std::memset(static_cast<void*>(ghost_buffer.data.data()), 0, ghost_buffer.data.size() * sizeof(Vector3D));
    buffer_state_ = kAllClear;
  };
  auto clear_rect = [&] {
    for (int y = dirty_[1]; y <= dirty_[3]; y++)
      std::memset(static_cast<void*>(&ghost_buffer.data[(size_t)dirty_[0] + (size_t)y * frame_w_]), 0, sizeof(Vector3D) * (size_t)(dirty_[2] - dirty_[0] + 1));
    dirty_[0] = dirty_[1] = 0; dirty_[2] = dirty_[3] = -1;
    buffer_state_ = kAllClear;
  };
  // bring the buffer to what the chosen path expects
  if (buffer_state_ == kRectDirty) clear_rect();
  if (buffer_state_ == kUnknown && (path != kDense || !sun)) clear_all();
  if (buffer_state_ == kTilesTracked && (path != kSparse || ghost_buffer.data.data() != tracked_)) {
    if (path == kDense && sun) buffer_state_ = kUnknown;  // about to be rewritten in full
    else clear_all();
  }
  if (!sun && buffer_state_ != kTilesTracked) return;  // :724-726 (the buffer is clear)
  if (sun && (!camera || !camera->ghost_aperture_texture || camera->ghost_aperture_texture->aperture.empty()))
    throw Error(LFB_ERR_STATE, "camera->ghost_aperture_texture is not loaded");
  if (sun) upload_textures(true, false);
  else ensure_engine();
  params.width = (int)frame_w_;
  params.height = (int)frame_h_;
  std::vector<lfb_light> lights = sun ? make_lights(true) : std::vector<lfb_light>();
  if (path == kSparse) {
    pin_storage();
    int tiles = 0;
    const int was_clear = buffer_state_ == kAllClear ? 1 : 0;
    if (!device_ids_.empty()) {  // several GPUs: one engine each, the same call
      if (!multi_) check(lfb_create_multi(&multi_, device_ids_.data(), (int)device_ids_.size(), nullptr), "lfb_create_multi");
      if (multi_lens_dirty_) { check(lfb_multi_set_lens(multi_, &lens_), "lfb_multi_set_lens"); multi_lens_dirty_ = false; }
      const CameraApertureTexture* t = camera->ghost_aperture_texture;
      if (sun && multi_uploaded_ != t) {
        check(lfb_multi_set_aperture(multi_, t->aperture.data(), (int)t->width, (int)t->height), "lfb_multi_set_aperture");
        multi_uploaded_ = t;
      }
      check(lfb_render_ghosts_multi(multi_, lights.data(), (int)lights.size(), &params, ghost_buffer.data.data(), sizeof(Vector3D), LFB_F64x3,
                                    was_clear, &tiles),
            "lfb_render_ghosts_multi");
    } else
    check(lfb_render_ghosts_sparse(engine_, lights.data(), (int)lights.size(), &params, ghost_buffer.data.data(), sizeof(Vector3D), LFB_F64x3,
                                   was_clear, &tiles),
          "lfb_render_ghosts_sparse");
    last_tiles_ = tiles;
    tracked_ = ghost_buffer.data.data();
    buffer_state_ = tiles < 0 ? kUnknown : kTilesTracked;  // < 0: pageable storage, the engine rewrote every pixel
  } else if (path == kRect) {
    check(lfb_render_ghosts_rect(engine_, lights.data(), (int)lights.size(), &params, ghost_buffer.data.data(), sizeof(Vector3D), LFB_F64x3, dirty_),
          "lfb_render_ghosts_rect");
    buffer_state_ = dirty_[2] >= dirty_[0] ? kRectDirty : kAllClear;
  } else {
    if (pin_ghost_buffer) pin_storage(); else unpin_storage();
    check(lfb_render_ghosts(engine_, lights.data(), (int)lights.size(), &params, ghost_buffer.data.data(), sizeof(Vector3D), LFB_F64x3, 0),
          "lfb_render_ghosts");
    buffer_state_ = kUnknown;
  }
}

void PathTracer::upload_textures(bool ghost, bool star) {
  ensure_engine();
  if (ghost) {
    const CameraApertureTexture* t = camera ? camera->ghost_aperture_texture : nullptr;
    if (!t || t->aperture.empty()) throw Error(LFB_ERR_STATE, "camera->ghost_aperture_texture is not loaded");
    if (uploaded_ != t) {
      check(lfb_set_aperture(engine_, t->aperture.data(), (int)t->width, (int)t->height), "lfb_set_aperture");
      uploaded_ = t;
    }
  }
  if (star) {
    const CameraApertureTexture* t = camera ? camera->aperture_texture : nullptr;
    if (!t || t->aperture.empty()) throw Error(LFB_ERR_STATE, "camera->aperture_texture is not loaded");
    if (uploaded_star_ != t) {
      check(lfb_set_starburst_aperture(engine_, t->aperture.data(), (int)t->width, (int)t->height), "lfb_set_starburst_aperture");
      uploaded_star_ = t;
    }
  }
}

std::vector<lfb_light> PathTracer::make_lights(bool for_ghosts) const {
  std::vector<lfb_light> lights;
  if ((for_ghosts && params.mode == LFB_MODE_REF_QUADS) || flare_origins.empty()) {
    lfb_light lt;
    std::memset(&lt, 0, sizeof(lt));
    lt.ns_x = axis_ray.x; lt.ns_y = axis_ray.y; lt.theta = angle_to_sun;
    lt.radiance[0] = lt.radiance[1] = lt.radiance[2] = 1.f;
    if (!for_ghosts && !flare_radiance.empty()) {
      lt.radiance[0] = (float)flare_radiance[0].x; lt.radiance[1] = (float)flare_radiance[0].y; lt.radiance[2] = (float)flare_radiance[0].z;
    }
    lights.push_back(lt);
    return lights;
  }
  for (size_t l = 0; l < flare_origins.size(); l++) {
    lfb_light lt;
    std::memset(&lt, 0, sizeof(lt));
    lt.ns_x = flare_origins[l].x; lt.ns_y = flare_origins[l].y;
    if (params.mode == LFB_MODE_EXACT_GRID && camera) {
      // real refraction needs the real off-axis angle of the light (the inverse of analyze_world_coord), not the
      // reference's screen-space atan(ns_y/ns_x)
      const double tx = (2 * lt.ns_x - 1) * std::tan(0.5 * camera->hFov * kPi / 180.0);
      const double ty = (2 * lt.ns_y - 1) * std::tan(0.5 * camera->vFov * kPi / 180.0);
      lt.theta = (float)std::atan(std::sqrt(tx * tx + ty * ty));
    } else {
      lt.theta = (float)std::atan(lt.ns_y / lt.ns_x);
    }
    lt.radiance[0] = (float)flare_radiance[l].x; lt.radiance[1] = (float)flare_radiance[l].y; lt.radiance[2] = (float)flare_radiance[l].z;
    lt.distance = l < flare_distance.size() ? flare_distance[l] : 0.0;
    lights.push_back(lt);
  }
  return lights;
}

void PathTracer::render_starburst(HDRImageBuffer& target, bool additive) {
  if (flare_origins.empty()) return;  // the reference dereferences flare_origins[0] unconditionally (pathtracer.cpp:919): UB there
  if (target.w != frame_w_ || target.h != frame_h_ || target.data.size() != frame_w_ * frame_h_) {
    target.resize(frame_w_, frame_h_);
  }
  upload_textures(false, true);
  std::vector<lfb_light> lights = make_lights(false);
  check(lfb_render_starburst(engine_, lights.data(), (int)lights.size(), (int)frame_w_, (int)frame_h_, flare_radius, flare_intensity,
                             target.data.data(), sizeof(Vector3D), LFB_F64x3, additive ? 1 : 0),
        "lfb_render_starburst");
}

void PathTracer::render_frame(std::vector<uint32_t>& rgba8, const HDRImageBuffer* base, bool with_starburst, bool flip_vertical) {
  rgba8.assign(frame_w_ * frame_h_, 0xFF000000u);
  if (base && (base->w != frame_w_ || base->h != frame_h_)) throw Error(LFB_ERR_INVALID, "base frame size mismatch");
  upload_textures(true, with_starburst);
  params.width = (int)frame_w_;
  params.height = (int)frame_h_;
  const bool sun = !(axis_ray.x == 0 && axis_ray.y == 0);
  std::vector<lfb_light> lights = sun ? make_lights(true) : std::vector<lfb_light>();
  check(lfb_render_frame_rgba8(engine_, lights.data(), (int)lights.size(), &params, with_starburst && sun ? flare_radius : -1.0, flare_intensity,
                               base ? &base->data[0].x : nullptr, rgba8.data(), flip_vertical ? 1 : 0),
        "lfb_render_frame_rgba8");
}

float PathTracer::last_trace_ms() const {
  float t = 0;
  if (engine_) lfb_stats(engine_, nullptr, &t, nullptr);
  return t;
}
float PathTracer::last_frame_ms() const {
  float t = 0;
  if (engine_) lfb_stats(engine_, nullptr, nullptr, &t);
  return t;
}

}  // namespace lfb
