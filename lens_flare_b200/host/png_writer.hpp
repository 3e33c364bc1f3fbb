// png_writer.hpp -- RGBA8 PNG encoder on zlib, for the frame save at the end of a render.
// Stands in for lodepng::encode as RaytracedRenderer::save_image uses it
// (src/pathtracer/raytraced_renderer.cpp:717-755): 8-bit RGBA, non-interlaced; pixels are stored
// losslessly, so any decoder returns the bytes the reference's file would hold.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace lfb {
// `rgba` is ImageBuffer-style: one uint32 per pixel, R in the low byte (src/util/image.h:53-62), row 0 first.
// Throws std::runtime_error on failure.
void write_png_rgba8(const std::string& path, const uint32_t* rgba, size_t width, size_t height);

// save_image's pixel handling (:737-746): rows written bottom-up (unless the frame is already flipped) and alpha
// forced to 255, then encoded.
void save_image(const std::string& path, const uint32_t* frame, size_t width, size_t height, bool flip_vertical);
}  // namespace lfb
