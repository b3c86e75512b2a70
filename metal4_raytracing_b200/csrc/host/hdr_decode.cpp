// hdr_decode.cpp — Radiance RGBE (.hdr) reader for the environment extension (include/rt_b200.h rt_environment).
// BASELINE's K3 names vulture_hide_4k.hdr as its environment light; the reference ships that file but never loads it
// (SURVEY.md F5), and it is absent from this mount, so the reader is exercised by files the tests write.
// Supported: "#?RADIANCE" / "#?RGBE" header, FORMAT=32-bit_rle_rgbe, resolution "-Y h +X w" (rows top to bottom),
// flat pixels and new-style run-length-encoded scanlines. A pixel (r, g, b, e) decodes to (r, g, b) * 2^(e - 136),
// e = 0 to black.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace rts {

bool decodeHdr(const std::string &path, int &width, int &height, std::vector<float> &rgba, std::string &err) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) {
    err = "cannot open " + path;
    return false;
  }
  std::vector<uint8_t> data;
  {
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + n);
    std::fclose(f);
  }
  size_t pos = 0;
  auto line = [&](std::string &out) {
    out.clear();
    if (pos >= data.size()) return false;
    while (pos < data.size() && data[pos] != '\n') out.push_back(char(data[pos++]));
    if (pos < data.size()) ++pos;
    if (!out.empty() && out.back() == '\r') out.pop_back();
    return true;
  };
  std::string l;
  if (!line(l) || (l.rfind("#?RADIANCE", 0) != 0 && l.rfind("#?RGBE", 0) != 0)) {
    err = path + ": not a Radiance picture";
    return false;
  }
  bool formatOk = false;
  while (line(l) && !l.empty())
    if (l.rfind("FORMAT=", 0) == 0) formatOk = l == "FORMAT=32-bit_rle_rgbe";
  if (!formatOk) {
    err = path + ": FORMAT=32-bit_rle_rgbe expected";
    return false;
  }
  int w = 0, h = 0;
  if (!line(l) || std::sscanf(l.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0 || w > 65536 || h > 65536) {
    err = path + ": resolution line \"-Y h +X w\" expected";
    return false;
  }
  std::vector<uint8_t> scan(size_t(w) * 4);
  rgba.assign(size_t(w) * size_t(h) * 4, 1.0f);
  for (int y = 0; y < h; ++y) {
    const bool rle = w >= 8 && w < 32768 && pos + 4 <= data.size() && data[pos] == 2 && data[pos + 1] == 2 &&
                     (int(data[pos + 2]) << 8 | int(data[pos + 3])) == w;
    if (rle) {
      pos += 4;
      for (int c = 0; c < 4; ++c) { // the four channels of the row follow one another, each run-length coded
        int x = 0;
        while (x < w) {
          if (pos >= data.size()) {
            err = path + ": truncated scanline";
            return false;
          }
          int count = data[pos++];
          if (count > 128) { // a run
            count -= 128;
            if (x + count > w || pos >= data.size()) {
              err = path + ": bad run in a scanline";
              return false;
            }
            const uint8_t v = data[pos++];
            for (int i = 0; i < count; ++i) scan[size_t(x++) * 4 + c] = v;
          } else { // literals
            if (count == 0 || x + count > w || pos + size_t(count) > data.size()) {
              err = path + ": bad literal block in a scanline";
              return false;
            }
            for (int i = 0; i < count; ++i) scan[size_t(x++) * 4 + c] = data[pos++];
          }
        }
      }
    } else {
      if (pos + size_t(w) * 4 > data.size()) {
        err = path + ": truncated pixel data";
        return false;
      }
      std::memcpy(scan.data(), data.data() + pos, size_t(w) * 4);
      pos += size_t(w) * 4;
    }
    float *row = rgba.data() + size_t(y) * size_t(w) * 4;
    for (int x = 0; x < w; ++x) {
      const uint8_t *p = scan.data() + size_t(x) * 4;
      const float scale = p[3] ? std::ldexp(1.0f, int(p[3]) - 136) : 0.0f;
      row[4 * x] = float(p[0]) * scale, row[4 * x + 1] = float(p[1]) * scale, row[4 * x + 2] = float(p[2]) * scale;
    }
  }
  width = w, height = h;
  return true;
}

// Writer: flat (not run-length-encoded) RGBE pixels, rows top to bottom; the shared exponent is that of the largest
// channel, mantissas are truncated (Ward's float2rgbe), so decode(encode(x)) is within 1/128 of the largest channel.
bool encodeHdr(const std::string &path, int width, int height, const float *rgba, int channels, std::string &err) {
  if (width <= 0 || height <= 0 || !rgba || (channels != 3 && channels != 4)) {
    err = "empty image or unsupported channel count";
    return false;
  }
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) {
    err = "cannot write " + path;
    return false;
  }
  std::fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", height, width);
  std::vector<uint8_t> row(size_t(width) * 4);
  bool ok = true;
  for (int y = 0; y < height && ok; ++y) {
    for (int x = 0; x < width; ++x) {
      const float *p = rgba + (size_t(y) * size_t(width) + size_t(x)) * size_t(channels);
      const float r = p[0] > 0.0f ? p[0] : 0.0f, g = p[1] > 0.0f ? p[1] : 0.0f, b = p[2] > 0.0f ? p[2] : 0.0f;
      const float m = r > g ? (r > b ? r : b) : (g > b ? g : b);
      uint8_t *o = row.data() + size_t(x) * 4;
      if (!(m > 1e-32f) || !std::isfinite(m)) {
        o[0] = o[1] = o[2] = o[3] = 0;
      } else {
        int e = 0;
        const float scale = std::frexp(m, &e) * 256.0f / m; // m = mantissa * 2^e, mantissa in [0.5, 1)
        o[0] = uint8_t(r * scale), o[1] = uint8_t(g * scale), o[2] = uint8_t(b * scale);
        o[3] = uint8_t(e + 128);
      }
    }
    ok = std::fwrite(row.data(), 1, row.size(), f) == row.size();
  }
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) err = "write failed: " + path;
  return ok;
}

} // namespace rts
