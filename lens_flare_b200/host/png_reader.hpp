// png_reader.hpp -- minimal PNG decoder for the aperture masks (8-bit gray, gray+alpha, RGB, RGBA,
// palette; non-interlaced), on zlib.  The reference decodes with its vendored lodepng to RGBA8 and keeps
// byte 0 of each texel (src/pathtracer/camera.h:36-59); any conforming decoder yields the same bytes.
#pragma once
#include <string>
#include <vector>

namespace lfb {
// Decodes `path` and returns the RED channel, row-major.  Throws std::runtime_error on failure.
void read_png_red(const std::string& path, std::vector<unsigned char>& red, unsigned& width, unsigned& height);
}  // namespace lfb
