#include "png_writer.hpp"

#include <zlib.h>

#include <cstdio>
#include <stdexcept>
#include <vector>

namespace lfb {
namespace {

void put_be32(std::vector<unsigned char>& v, uint32_t x) {
  v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16));
  v.push_back((unsigned char)(x >> 8));  v.push_back((unsigned char)x);
}

void put_chunk(std::vector<unsigned char>& file, const char type[4], const unsigned char* data, size_t len) {
  put_be32(file, (uint32_t)len);
  const size_t start = file.size();
  file.insert(file.end(), type, type + 4);
  if (len) file.insert(file.end(), data, data + len);
  put_be32(file, (uint32_t)crc32(0L, &file[start], (uInt)(len + 4)));  // CRC covers type + data
}

}  // namespace

void write_png_rgba8(const std::string& path, const uint32_t* rgba, size_t width, size_t height) {
  if (!rgba || width == 0 || height == 0 || width > 0x7fffffffu || height > 0x7fffffffu)
    throw std::runtime_error("write_png_rgba8: bad image size");
  // filter type 0 on every scanline: flare frames are mostly flat, deflate does the work
  const size_t row = 1 + width * 4;
  std::vector<unsigned char> raw(row * height);
  for (size_t y = 0; y < height; y++) {
    unsigned char* d = &raw[y * row];
    *d++ = 0;
    for (size_t x = 0; x < width; x++) {
      const uint32_t p = rgba[y * width + x];
      *d++ = (unsigned char)p; *d++ = (unsigned char)(p >> 8); *d++ = (unsigned char)(p >> 16); *d++ = (unsigned char)(p >> 24);
    }
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<unsigned char> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw std::runtime_error("write_png_rgba8: deflate failed");

  std::vector<unsigned char> file = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  std::vector<unsigned char> ihdr;
  put_be32(ihdr, (uint32_t)width);
  put_be32(ihdr, (uint32_t)height);
  const unsigned char tail[5] = {8, 6, 0, 0, 0};  // 8 bits, RGBA, deflate, adaptive filtering, no interlace
  ihdr.insert(ihdr.end(), tail, tail + 5);
  put_chunk(file, "IHDR", ihdr.data(), ihdr.size());
  put_chunk(file, "IDAT", z.data(), zlen);
  put_chunk(file, "IEND", nullptr, 0);

  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write " + path);
  const bool ok = std::fwrite(file.data(), 1, file.size(), f) == file.size();
  if (std::fclose(f) != 0 || !ok) throw std::runtime_error("short write to " + path);
}

void save_image(const std::string& path, const uint32_t* frame, size_t width, size_t height, bool flip_vertical) {
  std::vector<uint32_t> out(width * height);
  for (size_t y = 0; y < height; y++) {
    const uint32_t* src = frame + (flip_vertical ? height - 1 - y : y) * width;
    for (size_t x = 0; x < width; x++) out[y * width + x] = src[x] | 0xFF000000u;
  }
  write_png_rgba8(path, out.data(), width, height);
}

}  // namespace lfb
