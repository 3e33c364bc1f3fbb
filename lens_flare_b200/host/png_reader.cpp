#include "png_reader.hpp"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace lfb {
namespace {

unsigned be32(const unsigned char* p) { return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

void read_png_red(const std::string& path, std::vector<unsigned char>& red, unsigned& width, unsigned& height) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<unsigned char> file;
  unsigned char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
  std::fclose(f);
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (file.size() < 33 || std::memcmp(file.data(), sig, 8) != 0) throw std::runtime_error(path + ": not a PNG file");

  unsigned bit_depth = 0, colour = 0, interlace = 0;
  std::vector<unsigned char> idat, palette;
  size_t pos = 8;
  bool have_ihdr = false, done = false;
  while (!done && pos + 12 <= file.size()) {
    const unsigned len = be32(&file[pos]);
    const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
    if (pos + 12 + (size_t)len > file.size()) throw std::runtime_error(path + ": truncated chunk");
    const unsigned char* data = &file[pos + 8];
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len < 13) throw std::runtime_error(path + ": bad IHDR");
      width = be32(data); height = be32(data + 4);
      bit_depth = data[8]; colour = data[9]; interlace = data[12];
      have_ihdr = true;
    } else if (!std::memcmp(type, "PLTE", 4)) {
      palette.assign(data, data + len);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      done = true;
    }
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr || idat.empty()) throw std::runtime_error(path + ": missing IHDR/IDAT");
  if (bit_depth != 8 || interlace != 0) throw std::runtime_error(path + ": only 8-bit non-interlaced PNGs are supported");
  unsigned channels;
  switch (colour) {
    case 0: channels = 1; break;  // gray
    case 2: channels = 3; break;  // RGB
    case 3: channels = 1; break;  // palette index
    case 4: channels = 2; break;  // gray + alpha
    case 6: channels = 4; break;  // RGBA
    default: throw std::runtime_error(path + ": unknown colour type");
  }
  if (width == 0 || height == 0 || width > 16384 || height > 16384) throw std::runtime_error(path + ": unreasonable size");
  const size_t stride = (size_t)width * channels;
  std::vector<unsigned char> raw((stride + 1) * height);
  uLongf raw_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size())
    throw std::runtime_error(path + ": zlib inflate failed");

  // undo the scanline filters in place
  std::vector<unsigned char> prev(stride, 0), cur(stride);
  red.assign((size_t)width * height, 0);
  for (unsigned y = 0; y < height; y++) {
    const unsigned char* line = &raw[(stride + 1) * y];
    const unsigned filter = line[0];
    for (size_t x = 0; x < stride; x++) {
      const int a = x >= channels ? cur[x - channels] : 0, b = prev[x], c = x >= channels ? prev[x - channels] : 0;
      int v = line[1 + x];
      switch (filter) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: throw std::runtime_error(path + ": bad scanline filter");
      }
      cur[x] = (unsigned char)v;
    }
    for (unsigned x = 0; x < width; x++) {
      unsigned char r = cur[(size_t)x * channels];
      if (colour == 3) {
        if ((size_t)r * 3 + 2 >= palette.size()) throw std::runtime_error(path + ": palette index out of range");
        r = palette[(size_t)r * 3];
      }
      red[(size_t)y * width + x] = r;
    }
    prev.swap(cur);
  }
}

}  // namespace lfb
