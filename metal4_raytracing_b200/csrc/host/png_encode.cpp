// png_encode.cpp — minimal RGBA8 -> PNG writer on top of zlib's deflate: the image-writer end of the post chain
// (the reference presents to a drawable instead, MetalRaytracing/FramePresenter.swift).
// 8-bit RGBA, colour type 6, filter 0 on every row, one IDAT chunk.
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/rt_scene.h"

namespace {

void put32(std::vector<uint8_t> &v, uint32_t x) {
  v.push_back(uint8_t(x >> 24));
  v.push_back(uint8_t(x >> 16));
  v.push_back(uint8_t(x >> 8));
  v.push_back(uint8_t(x));
}

void chunk(std::vector<uint8_t> &out, const char type[4], const uint8_t *data, size_t n) {
  put32(out, uint32_t(n));
  const size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  if (n) out.insert(out.end(), data, data + n);
  put32(out, uint32_t(crc32(0L, out.data() + start, uInt(out.size() - start))));
}

} // namespace

extern "C" int rts_write_png(const char *path, const uint8_t *rgba8, int width, int height) {
  if (!path || !rgba8 || width <= 0 || height <= 0) return 2;
  std::vector<uint8_t> raw(size_t(height) * (size_t(width) * 4 + 1));
  for (int y = 0; y < height; ++y) {
    uint8_t *row = raw.data() + size_t(y) * (size_t(width) * 4 + 1);
    row[0] = 0; // filter: none
    std::memcpy(row + 1, rgba8 + size_t(y) * size_t(width) * 4, size_t(width) * 4);
  }
  uLongf bound = compressBound(uLong(raw.size()));
  std::vector<uint8_t> z(bound);
  if (compress2(z.data(), &bound, raw.data(), uLong(raw.size()), 6) != Z_OK) return 1;
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  std::vector<uint8_t> ihdr;
  put32(ihdr, uint32_t(width));
  put32(ihdr, uint32_t(height));
  const uint8_t tail[5] = {8, 6, 0, 0, 0}; // bit depth 8, RGBA, deflate, adaptive filtering, no interlace
  ihdr.insert(ihdr.end(), tail, tail + 5);
  chunk(out, "IHDR", ihdr.data(), ihdr.size());
  chunk(out, "IDAT", z.data(), bound);
  chunk(out, "IEND", nullptr, 0);
  FILE *f = std::fopen(path, "wb");
  if (!f) return 1;
  const size_t wrote = std::fwrite(out.data(), 1, out.size(), f);
  std::fclose(f);
  return wrote == out.size() ? 0 : 1;
}
