// post.cu — display transform of a rendered frame.
//
// The reference shows the accumulation image through a full-screen quad whose fragment shader applies
// color / (1 + color) (MetalRaytracing/Shaders.metal:38-52) and whose uv = position * 0.5 + 0.5 (:30-35) puts image
// row 0 — the bottom of the view, the kernel has no y flip (Raytracing.metal:272-292) — at the bottom of the
// screen. rt_tonemap does the same per pixel into an RGBA8 buffer for an image writer: Reinhard, optional sRGB
// transfer (what an *_srgb drawable applies on store) and optional row flip so that row 0 is the top of the picture.
// Pure streaming: 8 or 16 B read, 4 B written per pixel.
#include <cuda_fp16.h>

#include "shade.cuh"

namespace rtb {

namespace {

__device__ __forceinline__ float3 loadRgb(const rt_image &img, size_t i) {
  if (img.format == RT_FORMAT_RGBA32_FLOAT) {
    const float4 v = static_cast<const float4 *>(img.data)[i];
    return make_float3(v.x, v.y, v.z);
  }
  const uint2 raw = static_cast<const uint2 *>(img.data)[i]; // rgba16f
  const __half2 a = *reinterpret_cast<const __half2 *>(&raw.x), b = *reinterpret_cast<const __half2 *>(&raw.y);
  return make_float3(__low2float(a), __high2float(a), __low2float(b));
}

// sRGB opto-electronic transfer, evaluated in double and rounded once (tests compare against numpy float64)
__device__ __forceinline__ float srgbEncode(float c) {
  const double x = double(c);
  return float(x <= 0.0031308 ? 12.92 * x : 1.055 * pow(x, 1.0 / 2.4) - 0.055);
}

__device__ __forceinline__ uint32_t toByte(float c, bool srgb) {
  c = fminf(fmaxf(c, 0.0f), 1.0f); // NaN -> 0
  if (srgb) c = srgbEncode(c);
  return uint32_t(c * 255.0f + 0.5f);
}

__global__ void k_tonemap(const rt_image src, uchar4 *__restrict__ dst, uint32_t flags) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= src.width || y >= src.height) return;
  float3 c = loadRgb(src, size_t(y) * size_t(src.width) + size_t(x));
  c.x = fmaxf(c.x, 0.0f), c.y = fmaxf(c.y, 0.0f), c.z = fmaxf(c.z, 0.0f);
  c.x = c.x / (1.0f + c.x), c.y = c.y / (1.0f + c.y), c.z = c.z / (1.0f + c.z); // Shaders.metal:48
  const bool srgb = (flags & RT_TONEMAP_SRGB) != 0;
  const int outY = (flags & RT_TONEMAP_FLIP_Y) ? src.height - 1 - y : y;
  dst[size_t(outY) * size_t(src.width) + size_t(x)] =
      make_uchar4(uint8_t(toByte(c.x, srgb)), uint8_t(toByte(c.y, srgb)), uint8_t(toByte(c.z, srgb)), 255);
}

// ---- temporal reprojection filter -----------------------------------------------------------------------------
// The consumer of the kernel's depth / motion / normal outputs (Raytracing.metal:341-389, 506-515), the role the
// reference hands to MetalFX's temporal denoiser (FramePresenter.swift:435-521, closed source). Specified here:
//   prev = (x - motion.x, y + motion.y)        the kernel stores motion in pixels with +y down, image rows grow upward
//   history = bilinear(historyColor, prev), taken only if prev is inside the image and the nearest history texel has
//             |historyDepth - depth| <= depthTolerance * depth and dot(normal, historyNormal) >= normalThreshold
//   history is clamped to the min / max of the current frame's 3x3 neighbourhood (anti-ghosting)
//   out = color + (history - color) * historyWeight      (no valid history: out = color)
// One thread per pixel; reads ~10 texels, writes one. Arithmetic in the association order written (-fmad=false).
__global__ void k_temporal_filter(const rt_denoise_frame cur, const rt_denoise_frame hist, const rt_image out,
                                  float historyWeight, float depthTolerance, float normalThreshold) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const int w = cur.color.width, h = cur.color.height;
  if (x >= w || y >= h) return;
  const f4 c4 = readImage(cur.color, x, y);
  f3 c = mk3(c4.x, c4.y, c4.z);
  f3 result = c;
  const f4 mv = readImage(cur.motion, x, y);
  const float px = float(x) - mv.x, py = float(y) + mv.y;
  const float depth = readImage(cur.depth, x, y).x;
  if (hist.color.data != nullptr && px >= 0.0f && py >= 0.0f && px <= float(w - 1) && py <= float(h - 1) &&
      depth < 1.0e7f) {
    const int nx = int(floorf(px + 0.5f)), ny = int(floorf(py + 0.5f));
    const float hd = readImage(hist.depth, nx, ny).x;
    const f4 n4 = readImage(cur.normal, x, y), hn4 = readImage(hist.normal, nx, ny);
    const f3 n = mk3(n4.x, n4.y, n4.z) * 2.0f - mk3(1.0f), hn = mk3(hn4.x, hn4.y, hn4.z) * 2.0f - mk3(1.0f);
    if (fabsf(hd - depth) <= depthTolerance * depth && dot(n, hn) >= normalThreshold) {
      const float fx0 = floorf(px), fy0 = floorf(py);
      const float tx = px - fx0, ty = py - fy0;
      const int x0 = int(fx0), y0 = int(fy0), x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
      const f4 a = readImage(hist.color, x0, y0), b = readImage(hist.color, x1, y0);
      const f4 d = readImage(hist.color, x0, y1), e = readImage(hist.color, x1, y1);
      const f3 top = mix(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), tx);
      const f3 bottom = mix(mk3(d.x, d.y, d.z), mk3(e.x, e.y, e.z), tx);
      f3 history = mix(top, bottom, ty);
      f3 lo = c, hi = c;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int qx = min(max(x + dx, 0), w - 1), qy = min(max(y + dy, 0), h - 1);
          const f4 q = readImage(cur.color, qx, qy);
          lo = mk3(fminf(lo.x, q.x), fminf(lo.y, q.y), fminf(lo.z, q.z));
          hi = mk3(fmaxf(hi.x, q.x), fmaxf(hi.y, q.y), fmaxf(hi.z, q.z));
        }
      history = mk3(fminf(fmaxf(history.x, lo.x), hi.x), fminf(fmaxf(history.y, lo.y), hi.y),
                    fminf(fmaxf(history.z, lo.z), hi.z));
      result = mix(c, history, historyWeight);
    }
  }
  writeImage(out, x, y, {result.x, result.y, result.z, 1.0f});
}

// rt_spatial_filter (rt_b200.h): one a-trous pass, one thread per pixel
__global__ void k_spatial_filter(const rt_denoise_frame fr, const rt_image out, int step, float depthSigma,
                                 int normalSquarings, float colorSigma) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const int w = fr.color.width, h = fr.color.height;
  if (x >= w || y >= h) return;
  const f4 c4 = readImage(fr.color, x, y);
  const f3 c = mk3(c4.x, c4.y, c4.z);
  const float depth = readImage(fr.depth, x, y).x;
  f3 result = c;
  if (depth < 1.0e7f) {
    const f4 n4 = readImage(fr.normal, x, y);
    const f3 n = mk3(n4.x, n4.y, n4.z) * 2.0f - mk3(1.0f);
    const float kernel1[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    const float depthScale = depthSigma * depth;
    const float colorScale = colorSigma * colorSigma;
    f3 sum = mk3(0.0f);
    float weightSum = 0.0f;
    for (int dy = -2; dy <= 2; ++dy)
      for (int dx = -2; dx <= 2; ++dx) {
        const int qx = x + dx * step, qy = y + dy * step;
        if (qx < 0 || qy < 0 || qx >= w || qy >= h) continue;
        const float qd = readImage(fr.depth, qx, qy).x;
        if (!(qd < 1.0e7f)) continue;
        const f4 q4 = readImage(fr.color, qx, qy);
        const f3 q = mk3(q4.x, q4.y, q4.z);
        const f4 qn4 = readImage(fr.normal, qx, qy);
        const f3 qn = mk3(qn4.x, qn4.y, qn4.z) * 2.0f - mk3(1.0f);
        float wn = fmaxf(dot(n, qn), 0.0f);
        for (int k = 0; k < normalSquarings; ++k) wn = wn * wn;
        const float dz = fabsf(depth - qd) / depthScale;
        const float wz = 1.0f / (1.0f + dz * dz);
        float wc = 1.0f;
        if (colorSigma > 0.0f) {
          const f3 dc = c - q;
          wc = 1.0f / (1.0f + dot(dc, dc) / colorScale);
        }
        const float wgt = ((kernel1[dx + 2] * kernel1[dy + 2]) * wn) * (wz * wc);
        sum = sum + q * wgt;
        weightSum = weightSum + wgt;
      }
    if (weightSum > 0.0f) result = sum / weightSum;
  }
  writeImage(out, x, y, {result.x, result.y, result.z, 1.0f});
}

} // namespace

int launchSpatialFilter(rt_context *ctx, const rt_denoise_frame *fr, const rt_image *out, int step, float depthSigma,
                        int normalSquarings, float colorSigma) {
  RT_CHECK(fr && out && fr->color.data && fr->depth.data && fr->normal.data && out->data,
           "rt_spatial_filter: colour, depth, normal and output images must be bound");
  const int w = fr->color.width, h = fr->color.height;
  auto same = [&](const rt_image &i) { return i.width == w && i.height == h; };
  RT_CHECK(w > 0 && h > 0 && same(fr->depth) && same(fr->normal) && same(*out), "rt_spatial_filter: image sizes differ");
  RT_CHECK(step >= 1 && step <= 1024, "rt_spatial_filter: step is 1..1024");
  RT_CHECK(depthSigma > 0.0f && normalSquarings >= 0 && normalSquarings <= 8,
           "rt_spatial_filter: depthSigma > 0, normalSquarings 0..8");
  RT_CHECK(out->data != fr->color.data, "rt_spatial_filter: output aliases the input");
  const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
  k_spatial_filter<<<grid, block, 0, ctx->stream>>>(*fr, *out, step, depthSigma, normalSquarings, colorSigma);
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

int launchTemporalFilter(rt_context *ctx, const rt_denoise_frame *cur, const rt_denoise_frame *hist, const rt_image *out,
                         float historyWeight, float depthTolerance, float normalThreshold) {
  RT_CHECK(cur && out && cur->color.data && cur->motion.data && cur->depth.data && cur->normal.data && out->data,
           "rt_temporal_filter: current frame images and the output must be bound");
  const int w = cur->color.width, h = cur->color.height;
  auto same = [&](const rt_image &i) { return i.width == w && i.height == h; };
  RT_CHECK(w > 0 && h > 0 && same(cur->motion) && same(cur->depth) && same(cur->normal) && same(*out),
           "rt_temporal_filter: image sizes differ");
  rt_denoise_frame none{};
  if (hist == nullptr || hist->color.data == nullptr) hist = &none;
  else
    RT_CHECK(hist->depth.data && hist->normal.data && same(hist->color) && same(hist->depth) && same(hist->normal),
             "rt_temporal_filter: history needs colour, depth and normal images of the same size");
  RT_CHECK(out->data != cur->color.data && out->data != hist->color.data, "rt_temporal_filter: output aliases an input");
  const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
  k_temporal_filter<<<grid, block, 0, ctx->stream>>>(*cur, *hist, *out, historyWeight, depthTolerance, normalThreshold);
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

int launchTonemap(rt_context *ctx, const rt_image *src, uint8_t *dst, uint32_t flags) {
  RT_CHECK(src && src->data && dst, "rt_tonemap: null pointer");
  RT_CHECK(src->format == RT_FORMAT_RGBA16_FLOAT || src->format == RT_FORMAT_RGBA32_FLOAT,
           "rt_tonemap: source must be rgba16f or rgba32f");
  RT_CHECK(src->width > 0 && src->height > 0, "rt_tonemap: empty image");
  const dim3 block(32, 8), grid((src->width + 31) / 32, (src->height + 7) / 8);
  k_tonemap<<<grid, block, 0, ctx->stream>>>(*src, reinterpret_cast<uchar4 *>(dst), flags);
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

} // namespace rtb
