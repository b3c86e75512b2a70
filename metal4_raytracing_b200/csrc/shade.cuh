// shade.cuh — device-side math, sampling, texture and image helpers of the path-tracing kernels.
//
// These implement the per-hit work of the reference's raytracingKernel (MetalRaytracing/Raytracing.metal:28-218
// helpers, :391-774 shading) on sm_100a. Arithmetic follows the numeric contract in DESIGN.md: every float
// operation rounded on its own in the association order written here (translation units are compiled with
// -fmad=false), pow(x,5) by repeated multiplication, sin/cos evaluated in double and rounded to float.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace rtb {

struct f2 {
  float x, y;
};
struct f3 {
  float x, y, z;
};
struct f4 {
  float x, y, z, w;
};

__device__ __forceinline__ f2 mk2(float x, float y) { return {x, y}; }
__device__ __forceinline__ f3 mk3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ f3 mk3(float s) { return {s, s, s}; }
__device__ __forceinline__ f3 mk3(const rt_float3 &v) { return {v.x, v.y, v.z}; }

__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ f2 operator*(f2 a, float s) { return {a.x * s, a.y * s}; }
__device__ __forceinline__ f2 operator*(float s, f2 a) { return {s * a.x, s * a.y}; }
__device__ __forceinline__ f2 operator/(f2 a, f2 b) { return {a.x / b.x, a.y / b.y}; }
__device__ __forceinline__ f2 operator/(f2 a, float s) { return {a.x / s, a.y / s}; }

__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ f3 operator+(f3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
__device__ __forceinline__ f3 operator-(float s, f3 a) { return {s - a.x, s - a.y, s - a.z}; }
__device__ __forceinline__ f3 &operator+=(f3 &a, f3 b) {
  a = a + b;
  return a;
}
__device__ __forceinline__ f3 &operator*=(f3 &a, f3 b) {
  a = a * b;
  return a;
}
__device__ __forceinline__ f3 &operator*=(f3 &a, float s) {
  a = a * s;
  return a;
}

__device__ __forceinline__ float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float length(f3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ float length(f2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
__device__ __forceinline__ f3 normalize(f3 a) {
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float saturatef(float x) { return clampf(x, 0.0f, 1.0f); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a + (b - a) * t; }
__device__ __forceinline__ f3 mix(f3 a, f3 b, float t) { return a + (b - a) * t; }
__device__ __forceinline__ float pow5(float x) {
  float x2 = x * x;
  float x4 = x2 * x2;
  return x4 * x;
}
__device__ __forceinline__ float sinDet(float x) { return float(sin(double(x))); }
__device__ __forceinline__ float cosDet(float x) { return float(cos(double(x))); }

constexpr float kPi = 3.14159265358979323846f;

struct M34 { // object->world, columns c0..c3 (rows 0..2), from the 72-byte instance descriptor
  f3 c0, c1, c2, c3;
};
__device__ __forceinline__ M34 loadInstanceMatrix(const rt_instance_descriptor *d) {
  const float *m = &d->transformationMatrix[0][0];
  M34 r;
  r.c0 = mk3(__ldg(m + 0), __ldg(m + 1), __ldg(m + 2));
  r.c1 = mk3(__ldg(m + 3), __ldg(m + 4), __ldg(m + 5));
  r.c2 = mk3(__ldg(m + 6), __ldg(m + 7), __ldg(m + 8));
  r.c3 = mk3(__ldg(m + 9), __ldg(m + 10), __ldg(m + 11));
  return r;
}
__device__ __forceinline__ f3 mulPoint(const M34 &m, f3 v) { return ((m.c0 * v.x + m.c1 * v.y) + m.c2 * v.z) + m.c3; }
__device__ __forceinline__ f3 mulDir(const M34 &m, f3 v) { return (m.c0 * v.x + m.c1 * v.y) + m.c2 * v.z; }

// ---- Halton (Raytracing.metal:28-57) -------------------------------------------------------------------------
__constant__ short c_primes[100] = {
    2,   3,   5,   7,   11,  13,  17,  19,  23,  29,  31,  37,  41,  43,  47,  53,  59,  61,  67,  71,
    73,  79,  83,  89,  97,  101, 103, 107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173,
    179, 181, 191, 193, 197, 199, 211, 223, 227, 229, 233, 239, 241, 251, 257, 263, 269, 271, 277, 281,
    283, 293, 307, 311, 313, 317, 331, 337, 347, 349, 353, 359, 367, 373, 379, 383, 389, 397, 401, 409,
    419, 421, 431, 433, 439, 443, 449, 457, 461, 463, 467, 479, 487, 491, 499, 503, 509, 521, 523, 541};

__device__ __forceinline__ float halton(int i, int d) {
  const int b = c_primes[((d % 100) + 100) % 100]; // the reference indexes past the table for d > 99 (F9)
  float f = 1.0f;
  const float invB = 1.0f / float(b);
  float r = 0.0f;
  while (i > 0) {
    const int q = i / b;
    const int digit = i - q * b;
    f = f * invB;
    r = r + f * float(digit);
    i = q;
  }
  return r;
}

// ---- material textures: bilinear, repeat, LOD 0 (Raytracing.metal:421) ---------------------------------------
__device__ __forceinline__ f4 fetchTexel(const rt_texture2d &t, int x, int y, const float *__restrict__ srgbLut) {
  const uchar4 p = __ldg(reinterpret_cast<const uchar4 *>(t.texels) + (size_t(y) * size_t(t.width) + size_t(x)));
  f4 c;
  if (t.srgb) {
    c.x = __ldg(srgbLut + p.x);
    c.y = __ldg(srgbLut + p.y);
    c.z = __ldg(srgbLut + p.z);
  } else {
    c.x = float(p.x) / 255.0f;
    c.y = float(p.y) / 255.0f;
    c.z = float(p.z) / 255.0f;
  }
  c.w = float(p.w) / 255.0f;
  return c;
}
__device__ __forceinline__ int wrapIndex(int i, int n) {
  int m = i % n;
  return m < 0 ? m + n : m;
}
__device__ __forceinline__ f4 sampleTexture(const rt_texture2d *tp, f2 uv, const float *__restrict__ srgbLut) {
  const rt_texture2d t = *tp;
  float x = uv.x * float(t.width) - 0.5f, y = uv.y * float(t.height) - 0.5f;
  x = fminf(fmaxf(x, -1.0e9f), 1.0e9f);
  y = fminf(fmaxf(y, -1.0e9f), 1.0e9f);
  const float fx0 = floorf(x), fy0 = floorf(y);
  const float fx = x - fx0, fy = y - fy0;
  const int x0 = wrapIndex(int(fx0), t.width), y0 = wrapIndex(int(fy0), t.height);
  const int x1 = wrapIndex(x0 + 1, t.width), y1 = wrapIndex(y0 + 1, t.height);
  const f4 t00 = fetchTexel(t, x0, y0, srgbLut), t10 = fetchTexel(t, x1, y0, srgbLut),
           t01 = fetchTexel(t, x0, y1, srgbLut), t11 = fetchTexel(t, x1, y1, srgbLut);
  const float wx0 = 1.0f - fx, wy0 = 1.0f - fy;
  f4 r;
  r.x = (t00.x * wx0 + t10.x * fx) * wy0 + (t01.x * wx0 + t11.x * fx) * fy;
  r.y = (t00.y * wx0 + t10.y * fx) * wy0 + (t01.y * wx0 + t11.y * fx) * fy;
  r.z = (t00.z * wx0 + t10.z * fx) * wy0 + (t01.z * wx0 + t11.z * fx) * fy;
  r.w = (t00.w * wx0 + t10.w * fx) * wy0 + (t01.w * wx0 + t11.w * fx) * fy;
  return r;
}

// ---- render-target access in the bound format -------------------------------------------------------------------
__device__ __forceinline__ f4 readImage(const rt_image &img, int x, int y) {
  const size_t i = size_t(y) * size_t(img.width) + size_t(x);
  switch (img.format) {
    case RT_FORMAT_RGBA16_FLOAT: {
      const uint2 raw = reinterpret_cast<const uint2 *>(img.data)[i];
      const __half2 a = *reinterpret_cast<const __half2 *>(&raw.x), b = *reinterpret_cast<const __half2 *>(&raw.y);
      return {__low2float(a), __high2float(a), __low2float(b), __high2float(b)};
    }
    case RT_FORMAT_RGBA32_FLOAT: {
      const float4 v = reinterpret_cast<const float4 *>(img.data)[i];
      return {v.x, v.y, v.z, v.w};
    }
    case RT_FORMAT_RG16_FLOAT: {
      const __half2 a = reinterpret_cast<const __half2 *>(img.data)[i];
      return {__low2float(a), __high2float(a), 0.0f, 1.0f};
    }
    case RT_FORMAT_RG32_FLOAT: {
      const float2 v = reinterpret_cast<const float2 *>(img.data)[i];
      return {v.x, v.y, 0.0f, 1.0f};
    }
    case RT_FORMAT_R32_FLOAT:
      return {reinterpret_cast<const float *>(img.data)[i], 0.0f, 0.0f, 1.0f};
    case RT_FORMAT_R16_FLOAT:
      return {__half2float(reinterpret_cast<const __half *>(img.data)[i]), 0.0f, 0.0f, 1.0f};
    default:
      return {0.0f, 0.0f, 0.0f, 0.0f};
  }
}

__device__ __forceinline__ void writeImageAt(void *data, int format, size_t i, f4 v) {
  switch (format) {
    case RT_FORMAT_RGBA16_FLOAT: {
      const __half2 a = __halves2half2(__float2half_rn(v.x), __float2half_rn(v.y));
      const __half2 b = __halves2half2(__float2half_rn(v.z), __float2half_rn(v.w));
      uint2 raw;
      raw.x = *reinterpret_cast<const uint32_t *>(&a);
      raw.y = *reinterpret_cast<const uint32_t *>(&b);
      reinterpret_cast<uint2 *>(data)[i] = raw;
      break;
    }
    case RT_FORMAT_RGBA32_FLOAT:
      reinterpret_cast<float4 *>(data)[i] = make_float4(v.x, v.y, v.z, v.w);
      break;
    case RT_FORMAT_RG16_FLOAT:
      reinterpret_cast<__half2 *>(data)[i] = __halves2half2(__float2half_rn(v.x), __float2half_rn(v.y));
      break;
    case RT_FORMAT_RG32_FLOAT:
      reinterpret_cast<float2 *>(data)[i] = make_float2(v.x, v.y);
      break;
    case RT_FORMAT_R32_FLOAT:
      reinterpret_cast<float *>(data)[i] = v.x;
      break;
    case RT_FORMAT_R16_FLOAT:
      reinterpret_cast<__half *>(data)[i] = __float2half_rn(v.x);
      break;
    default:
      break;
  }
}
__device__ __forceinline__ void writeImage(const rt_image &img, int x, int y, f4 v) {
  if (img.data == nullptr) return;
  writeImageAt(img.data, img.format, size_t(y) * size_t(img.width) + size_t(x), v);
}

// ---- BRDF + sampling (Raytracing.metal:79-166) ------------------------------------------------------------------
__device__ __forceinline__ f3 sampleCosineWeightedHemisphere(f2 u) {
  const float phi = 2.0f * kPi * u.x;
  const float cosPhi = cosDet(phi), sinPhi = sinDet(phi);
  const float cosTheta = sqrtf(u.y);
  const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
  return mk3(sinTheta * cosPhi, cosTheta, sinTheta * sinPhi);
}
__device__ __forceinline__ f3 alignHemisphereWithNormal(f3 sample, f3 normal) {
  const f3 up = normal;
  const f3 right = normalize(cross(normal, mk3(0.0072f, 1.0f, 0.0034f)));
  const f3 forward = cross(right, up);
  return sample.x * right + sample.y * up + sample.z * forward;
}
__device__ __forceinline__ float distributionGGX(float NdotH, float alpha) {
  const float a2 = alpha * alpha;
  const float denom = (NdotH * NdotH) * (a2 - 1.0f) + 1.0f;
  return a2 / fmaxf(kPi * denom * denom, 1e-7f);
}
__device__ __forceinline__ float geometrySchlickGGX(float NdotV, float k) {
  return NdotV / fmaxf(NdotV * (1.0f - k) + k, 1e-7f);
}
__device__ __forceinline__ float geometrySmith(float NdotV, float NdotL, float k) {
  return geometrySchlickGGX(NdotV, k) * geometrySchlickGGX(NdotL, k);
}
__device__ __forceinline__ f3 fresnelSchlick(float cosTheta, f3 F0) {
  return F0 + (1.0f - F0) * pow5(clampf(1.0f - cosTheta, 0.0f, 1.0f));
}

} // namespace rtb
