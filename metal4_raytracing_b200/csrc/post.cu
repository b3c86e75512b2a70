// post.cu — display transform of a rendered frame.
//
// The reference shows the accumulation image through a full-screen quad whose fragment shader applies
// color / (1 + color) (MetalRaytracing/Shaders.metal:38-52) and whose uv = position * 0.5 + 0.5 (:30-35) puts image
// row 0 — the bottom of the view, the kernel has no y flip (Raytracing.metal:272-292) — at the bottom of the
// screen. rt_tonemap does the same per pixel into an RGBA8 buffer for an image writer: Reinhard, optional sRGB
// transfer (what an *_srgb drawable applies on store) and optional row flip so that row 0 is the top of the picture.
// Pure streaming: 8 or 16 B read, 4 B written per pixel.
#include <cuda_fp16.h>

#include "common.cuh"

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

} // namespace

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
