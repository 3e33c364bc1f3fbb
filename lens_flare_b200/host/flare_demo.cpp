// flare_demo.cpp -- the reference application's flare flags on top of the C++ facade:
//   flare_demo -r W H -y ghost_aperture.png [-x starburst_aperture.png] [-s ns_x ns_y] [-m ref|paraxial|exact] [-g N] [-f out.pfm|out.png]
// (-r, -x, -y, -f as in src/application/main.cpp:87, 135-152).  Prints frame statistics as one JSON line.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "flare_pathtracer.hpp"
#include "png_reader.hpp"
#include "png_writer.hpp"

int main(int argc, char** argv) {
  size_t W = 512, H = 512;
  std::string ghost_png, star_png, out;
  double sx = 0.45, sy = 0.55;
  int mode = LFB_MODE_REF_QUADS, grid = 256;
  double flare_intensity = 1.0, flare_radius = 30.0;  // -i / -n, main.cpp:135-152
  int repeat = 0, pin = 1;                            // --repeat K: time K more generate_ghost_buffer() calls on the host clock
  int frames = 0, in_flight = 3;                      // --frames N [--in-flight R]: an N-frame sweep of the sun, blocking and with R frames in flight
  int gpus = 1;                                       // --gpus N: devices 0 .. N-1 behind the one PathTracer (lfb_create_multi)
  if (argc == 3 && std::string(argv[1]) == "--png-info") {  // host-only: what CameraApertureTexture::init decodes
    try {
      lfb::CameraApertureTexture t;
      t.init(argv[2]);
      unsigned long long bytes = 0;
      for (float v : t.aperture) bytes += (unsigned long long)(v * 255.0f + 0.5f);
      std::printf("{\"w\": %zu, \"h\": %zu, \"total\": %.17g, \"bbox\": [%d, %d, %d, %d], \"byte_sum\": %llu}\n", t.width, t.height,
                  t.total_value, t.min_x, t.min_y, t.max_x, t.max_y, bytes);
      return 0;
    } catch (const std::exception& e) {
      std::fprintf(stderr, "flare_demo: %s\n", e.what());
      return 1;
    }
  }
  if (argc == 4 && std::string(argv[1]) == "--png-copy") {  // host-only: decode -> save_image (bottom-up rows, alpha 255)
    try {
      std::vector<unsigned char> red;
      unsigned w = 0, h = 0;
      lfb::read_png_red(argv[2], red, w, h);
      std::vector<uint32_t> frame((size_t)w * h);
      for (size_t i = 0; i < frame.size(); i++) frame[i] = red[i] | ((uint32_t)(red[i] / 2) << 8) | ((uint32_t)(255 - red[i]) << 16);
      lfb::save_image(argv[3], frame.data(), w, h, true);
      return 0;
    } catch (const std::exception& e) {
      std::fprintf(stderr, "flare_demo: %s\n", e.what());
      return 1;
    }
  }
  for (int a = 1; a < argc; a++) {
    const std::string k = argv[a];
    if (k == "-r" && a + 2 < argc) { W = std::strtoul(argv[a + 1], nullptr, 10); H = std::strtoul(argv[a + 2], nullptr, 10); a += 2; }
    else if (k == "-y" && a + 1 < argc) ghost_png = argv[++a];
    else if (k == "-x" && a + 1 < argc) star_png = argv[++a];
    else if (k == "-f" && a + 1 < argc) out = argv[++a];
    else if (k == "-s" && a + 2 < argc) { sx = std::atof(argv[a + 1]); sy = std::atof(argv[a + 2]); a += 2; }
    else if (k == "-g" && a + 1 < argc) grid = std::atoi(argv[++a]);
    else if (k == "-i" && a + 1 < argc) flare_intensity = std::atof(argv[++a]);
    else if (k == "-n" && a + 1 < argc) flare_radius = std::atof(argv[++a]);
    else if (k == "--repeat" && a + 1 < argc) repeat = std::atoi(argv[++a]);
    else if (k == "--no-pin") pin = 0;
    else if (k == "--gpus" && a + 1 < argc) gpus = std::atoi(argv[++a]);
    else if (k == "--frames" && a + 1 < argc) frames = std::atoi(argv[++a]);
    else if (k == "--in-flight" && a + 1 < argc) in_flight = std::atoi(argv[++a]);
    else if (k == "-m" && a + 1 < argc) {
      const std::string m = argv[++a];
      mode = m == "exact" ? LFB_MODE_EXACT_GRID : (m == "paraxial" ? LFB_MODE_PARAXIAL_GRID : LFB_MODE_REF_QUADS);
    } else { std::fprintf(stderr, "unknown argument %s\n", argv[a]); return 2; }
  }
  if (ghost_png.empty()) { std::fprintf(stderr, "usage: flare_demo -r W H -y ghost_aperture.png [-s ns_x ns_y] [-m ref|paraxial|exact] [-g N] [-f out.pfm]\n"); return 2; }
  try {
    lfb::CameraApertureTexture ghost_tex;
    ghost_tex.init(ghost_png);
    lfb::CameraApertureTexture star_tex;
    if (!star_png.empty()) star_tex.init(star_png);
    lfb::Camera camera;
    camera.ghost_aperture_texture = &ghost_tex;
    camera.aperture_texture = star_png.empty() ? &ghost_tex : &star_tex;
    // a directional light whose image lands at (sx, sy): invert analyze_world_coord for the identity camera
    const double kPi = 3.14159265358979323846;
    const double ex = std::tan(0.5 * camera.hFov * kPi / 180.0), ey = std::tan(0.5 * camera.vFov * kPi / 180.0);
    lfb::DirectionalLight sun(lfb::Vector3D(1, 1, 1), lfb::Vector3D(-(2 * sx - 1) * ex, -(2 * sy - 1) * ey, 1.0), lfb::Vector3D(0, 0, -1));
    lfb::Scene scene;
    scene.lights.push_back(&sun);
    std::vector<int> devices;
    for (int d = 0; d < (gpus > 1 ? gpus : 1); d++) devices.push_back(d);
    lfb::PathTracer pt(devices);
    pt.scene = &scene;
    pt.camera = &camera;
    pt.params.mode = mode;
    pt.params.grid_n = grid;
    if (mode != LFB_MODE_REF_QUADS) { pt.params.pair_set = LFB_PAIRS_ALL; pt.params.include_direct = 1; }
    pt.pin_ghost_buffer = pin != 0;
    pt.set_frame_size(W, H);
    pt.find_sun_pos();
    pt.generate_ghost_buffer();
    double host_ms = 0;
    if (repeat > 0) {  // steady state of the drop-in call: what the application's render loop pays per frame
      const auto t0 = std::chrono::steady_clock::now();
      for (int r = 0; r < repeat; r++) pt.generate_ghost_buffer();
      host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / repeat;
    }
    double sum[3] = {0, 0, 0}, l2 = 0;
    size_t nz = 0;
    for (const lfb::Vector3D& v : pt.ghost_buffer.data) {
      sum[0] += v.x; sum[1] += v.y; sum[2] += v.z;
      l2 += v.x * v.x + v.y * v.y + v.z * v.z;
      nz += (v.x != 0 || v.y != 0 || v.z != 0);
    }
    std::printf("{\"w\": %zu, \"h\": %zu, \"axis_ray\": [%.17g, %.17g], \"angle_to_sun\": %.9g, \"sum\": [%.17g, %.17g, %.17g], "
                "\"l2\": %.17g, \"nonzero\": %zu, \"trace_ms\": %.4f, \"frame_ms\": %.4f, \"host_ms_per_render\": %.4f, \"tex_total\": %.17g, "
                "\"tiles_written\": %d, \"gpus\": %d}\n",
                W, H, pt.axis_ray.x, pt.axis_ray.y, pt.angle_to_sun, sum[0], sum[1], sum[2], std::sqrt(l2), nz, pt.last_trace_ms(),
                pt.last_frame_ms(), host_ms, ghost_tex.total_value, pt.last_tiles_written(), gpus > 1 ? gpus : 1);
    if (!star_png.empty()) {  // the rest of raytrace_pixel's flare terms (:881-891) and the displayable frame
      pt.flare_radius = flare_radius;
      pt.flare_intensity = flare_intensity;
      lfb::HDRImageBuffer sample = pt.ghost_buffer;
      pt.render_starburst(sample, true);
      double ssum[3] = {0, 0, 0};
      for (const lfb::Vector3D& v : sample.data) { ssum[0] += v.x; ssum[1] += v.y; ssum[2] += v.z; }
      std::vector<uint32_t> rgba;
      pt.render_frame(rgba, nullptr, true, true);
      unsigned long long bytes = 0;
      for (uint32_t px : rgba) bytes += (px & 0xFF) + ((px >> 8) & 0xFF) + ((px >> 16) & 0xFF);
      std::printf("{\"ghost_plus_starburst_sum\": [%.17g, %.17g, %.17g], \"starburst_ms\": %.4f, \"rgba8_byte_sum\": %llu, \"frame_ms\": %.4f}\n",
                  ssum[0], ssum[1], ssum[2], pt.last_trace_ms(), bytes, pt.last_frame_ms());
    }
    if (frames > 0 && mode != LFB_MODE_REF_QUADS && gpus <= 1) {
      // a sequence: the sun moves on a small circle; every frame once through the blocking call, once through the ring of
      // ghost buffers with `in_flight` frames in flight -- same pixels, less time per frame
      if (in_flight < 1) in_flight = 1;
      if (in_flight > LFB_SPARSE_SLOTS) in_flight = LFB_SPARSE_SLOTS;
      auto place_sun = [&](int k) {
        const double a = 2 * kPi * k / 16.0, x = sx + 0.03 * std::cos(a), y = sy + 0.03 * std::sin(a);
        sun.posLight = lfb::Vector3D((2 * x - 1) * ex, (2 * y - 1) * ey, -1.0);  // (the constructor negates its argument)
        pt.flare_origins.clear(); pt.flare_radiance.clear(); pt.flare_distance.clear();  // find_sun_pos() appends, as the reference's does
        pt.find_sun_pos();
      };
      auto checksum = [](const lfb::HDRImageBuffer& b) {
        double s = 0;
        for (const lfb::Vector3D& v : b.data) s += v.x + 2 * v.y + 3 * v.z;
        return s;
      };
      std::vector<double> want((size_t)frames), got((size_t)frames);
      auto t0 = std::chrono::steady_clock::now();
      for (int k = 0; k < frames; k++) { place_sun(k); pt.generate_ghost_buffer(); want[(size_t)k] = checksum(pt.ghost_buffer); }
      const double blocking_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / frames;
      t0 = std::chrono::steady_clock::now();
      for (int k = 0; k < frames; k++) {
        const int s = k % in_flight;
        if (k >= in_flight) { pt.end_ghost_frame(s); got[(size_t)(k - in_flight)] = checksum(pt.ghost_ring[s]); }
        place_sun(k);
        pt.begin_ghost_frame(s);
      }
      for (int k = std::max(0, frames - in_flight); k < frames; k++) { pt.end_ghost_frame(k % in_flight); got[(size_t)k] = checksum(pt.ghost_ring[k % in_flight]); }
      const double flight_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / frames;
      int equal = 0;
      for (int k = 0; k < frames; k++) equal += want[(size_t)k] == got[(size_t)k];
      std::printf("{\"frames\": %d, \"in_flight\": %d, \"frames_equal\": %d, \"blocking_ms_per_frame_incl_checksum\": %.4f, "
                  "\"in_flight_ms_per_frame_incl_checksum_and_ring_setup\": %.4f}\n", frames, in_flight, equal, blocking_ms, flight_ms);
    }
    const bool out_png = out.size() > 4 && out.compare(out.size() - 4, 4, ".png") == 0;
    if (out_png) {  // the tone-mapped flare frame, as save_image writes it (raytraced_renderer.cpp:717-755)
      std::vector<uint32_t> rgba;
      pt.render_frame(rgba, nullptr, !star_png.empty(), true);  // rows already bottom-up
      lfb::save_image(out, rgba.data(), W, H, false);
    } else if (!out.empty()) {  // PFM, bottom row first -- the same vertical flip the reference applies on save (raytraced_renderer.cpp:739-742)
      FILE* f = std::fopen(out.c_str(), "wb");
      if (!f) { std::fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
      std::fprintf(f, "PF\n%zu %zu\n-1.0\n", W, H);
      for (size_t y = 0; y < H; y++)
        for (size_t x = 0; x < W; x++) {
          const lfb::Vector3D& v = pt.ghost_buffer.get_pixel_value(x, y);
          const float rgb[3] = {(float)v.x, (float)v.y, (float)v.z};
          std::fwrite(rgb, sizeof(float), 3, f);
        }
      std::fclose(f);
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "flare_demo: %s\n", e.what());
    return 1;
  }
  return 0;
}
