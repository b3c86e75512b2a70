// png_decode.cpp — minimal PNG -> RGBA8 decoder on top of zlib's inflate (stands in for MTKTextureLoader,
// SubMesh.swift:69-116). Supports 8- and 16-bit, colour types 0/2/3/4/6, non-interlaced — enough for the
// reference's material maps. 16-bit samples keep their high byte.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>

#include "scene.h"

namespace rts {

namespace {
uint32_t be32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
int paeth(int a, int b, int c) {
  int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
} // namespace

// The file is untrusted: chunk lengths, the IHDR fields and every size derived from them are checked before use,
// dimensions are capped (kMaxSide) so that no size computation can wrap, and allocation failure is reported as a
// decode failure instead of escaping through the extern "C" callers.
static bool decodePngChecked(const std::string &path, Texture &out);
bool decodePng(const std::string &path, Texture &out) {
  try {
    return decodePngChecked(path, out);
  } catch (const std::exception &) { // std::bad_alloc / length_error on hostile sizes
    return false;
  }
}

static bool decodePngChecked(const std::string &path, Texture &out) {
  constexpr uint32_t kMaxSide = 16384; // the reference's largest map is 4096^2
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> file(sz > 0 ? sz : 0);
  size_t got = std::fread(file.data(), 1, file.size(), f);
  std::fclose(f);
  static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  if (got < 8 || std::memcmp(file.data(), sig, 8) != 0) return false;
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = 0, interlace = 0;
  bool haveHeader = false;
  std::vector<uint8_t> idat, palette, trns;
  size_t pos = 8;
  while (pos + 12 <= file.size()) {
    uint32_t len = be32(&file[pos]);
    const uint8_t *type = &file[pos + 4];
    const uint8_t *data = &file[pos + 8];
    if (len > file.size() || pos + 12 + size_t(len) > file.size()) return false;
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len != 13 || haveHeader) return false;
      haveHeader = true;
      w = be32(data);
      h = be32(data + 4);
      depth = data[8];
      ctype = data[9];
      interlace = data[12];
    } else if (!std::memcmp(type, "PLTE", 4)) {
      palette.assign(data, data + len);
    } else if (!std::memcmp(type, "tRNS", 4)) {
      trns.assign(data, data + len);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + len;
  }
  if (!haveHeader || !w || !h || w > kMaxSide || h > kMaxSide || interlace || (depth != 8 && depth != 16)) return false;
  int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (!channels || (ctype == 3 && depth != 8)) return false;
  size_t bpp = size_t(channels) * (depth / 8), stride = bpp * w;
  std::vector<uint8_t> raw((stride + 1) * h);
  uLongf rawLen = raw.size();
  if (uncompress(raw.data(), &rawLen, idat.data(), idat.size()) != Z_OK || rawLen != raw.size()) return false;
  std::vector<uint8_t> img(stride * h);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t *src = &raw[(stride + 1) * y];
    uint8_t *dst = &img[stride * y];
    const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
    int ft = src[0];
    if (ft > 4) return false;
    ++src;
    for (size_t x = 0; x < stride; ++x) {
      int a = x >= bpp ? dst[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
      int v = src[x];
      switch (ft) {
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: break;
      }
      dst[x] = uint8_t(v);
    }
  }
  out.width = int(w);
  out.height = int(h);
  out.rgba.resize(size_t(w) * h * 4);
  size_t step = depth / 8;
  for (size_t i = 0; i < size_t(w) * h; ++i) {
    const uint8_t *p = &img[i * bpp];
    uint8_t r, g, b, a = 255;
    switch (ctype) {
      case 0: r = g = b = p[0]; break;
      case 2: r = p[0]; g = p[step]; b = p[2 * step]; break;
      case 3: {
        size_t k = p[0];
        r = k * 3 + 2 < palette.size() ? palette[k * 3] : 0;
        g = k * 3 + 2 < palette.size() ? palette[k * 3 + 1] : 0;
        b = k * 3 + 2 < palette.size() ? palette[k * 3 + 2] : 0;
        a = k < trns.size() ? trns[k] : 255;
        break;
      }
      case 4: r = g = b = p[0]; a = p[step]; break;
      default: r = p[0]; g = p[step]; b = p[2 * step]; a = p[3 * step]; break;
    }
    out.rgba[i * 4 + 0] = r;
    out.rgba[i * 4 + 1] = g;
    out.rgba[i * 4 + 2] = b;
    out.rgba[i * 4 + 3] = a;
  }
  return true;
}

} // namespace rts
