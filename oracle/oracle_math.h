// oracle_math.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product; see oracle/README.md.
//
// MSL-like float2/3/4 + float4x4 with the arithmetic *defined* op by op, so that a second implementation
// (the CUDA kernels) can reproduce it bit for bit. Must be compiled with -ffp-contract=off.
//
// Definitions that the reference leaves to Metal's fast-math library (PBX:406 MTL_FAST_MATH=YES) and that
// this oracle therefore pins (DESIGN.md "numeric contract"):
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z
//   length(v)     = sqrtf(dot(v,v));   normalize(v) = v * (1.0f / sqrtf(dot(v,v)))
//   mix(a,b,t)    = a + (b - a) * t;   clamp(x,lo,hi) = fminf(fmaxf(x,lo),hi)
//   pow(x,5)      = x2 = x*x; x4 = x2*x2; x4*x
//   sin/cos(x)    = (float)sin((double)x) / (float)cos((double)x)
//   M * (v,1)     = ((c0*v.x + c1*v.y) + c2*v.z) + c3;   M * (v,0) = (c0*v.x + c1*v.y) + c2*v.z
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

struct float2 {
  float x, y;
};
struct float3 {
  float x, y, z;
};
struct float4 {
  float x, y, z, w;
};

inline float2 make2(float x, float y) { return {x, y}; }
inline float3 make3(float x, float y, float z) { return {x, y, z}; }
inline float3 make3(float s) { return {s, s, s}; }

inline float2 operator+(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
inline float2 operator-(float2 a, float2 b) { return {a.x - b.x, a.y - b.y}; }
inline float2 operator*(float2 a, float s) { return {a.x * s, a.y * s}; }
inline float2 operator*(float s, float2 a) { return {s * a.x, s * a.y}; }
inline float2 operator/(float2 a, float2 b) { return {a.x / b.x, a.y / b.y}; }
inline float2 operator/(float2 a, float s) { return {a.x / s, a.y / s}; }

inline float3 operator+(float3 a, float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(float3 a, float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(float3 a) { return {-a.x, -a.y, -a.z}; }
inline float3 operator*(float3 a, float3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator*(float3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float s, float3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline float3 operator/(float3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float3 operator+(float3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline float3 operator-(float s, float3 a) { return {s - a.x, s - a.y, s - a.z}; }
inline float3 &operator+=(float3 &a, float3 b) {
  a = a + b;
  return a;
}
inline float3 &operator*=(float3 &a, float3 b) {
  a = a * b;
  return a;
}
inline float3 &operator*=(float3 &a, float s) {
  a = a * s;
  return a;
}

inline float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float3 cross(float3 a, float3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(float3 a) { return sqrtf(dot(a, a)); }
inline float length(float2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
inline float3 normalize(float3 a) {
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline float saturate(float x) { return clampf(x, 0.0f, 1.0f); }
inline float mixf(float a, float b, float t) { return a + (b - a) * t; }
inline float3 mix(float3 a, float3 b, float t) { return a + (b - a) * t; }
inline float pow5(float x) {
  float x2 = x * x;
  float x4 = x2 * x2;
  return x4 * x;
}
inline float sin_det(float x) { return static_cast<float>(std::sin(static_cast<double>(x))); }
inline float cos_det(float x) { return static_cast<float>(std::cos(static_cast<double>(x))); }

constexpr float kPi = 3.14159265358979323846f; // M_PI_F

struct float4x4 { // column-major like metal::float4x4: c[column]
  float4 c[4];
};
inline float3 mulPoint(const float4x4 &m, float3 v) { // (M * float4(v, 1)).xyz
  float3 c0{m.c[0].x, m.c[0].y, m.c[0].z}, c1{m.c[1].x, m.c[1].y, m.c[1].z}, c2{m.c[2].x, m.c[2].y, m.c[2].z},
      c3{m.c[3].x, m.c[3].y, m.c[3].z};
  return ((c0 * v.x + c1 * v.y) + c2 * v.z) + c3;
}
inline float3 mulDir(const float4x4 &m, float3 v) { // (M * float4(v, 0)).xyz
  float3 c0{m.c[0].x, m.c[0].y, m.c[0].z}, c1{m.c[1].x, m.c[1].y, m.c[1].z}, c2{m.c[2].x, m.c[2].y, m.c[2].z};
  return (c0 * v.x + c1 * v.y) + c2 * v.z;
}

// IEEE binary16 <-> binary32, round-to-nearest-even (what an rgba16Float texture write/read does).
inline uint16_t floatToHalf(float f) {
  uint32_t x;
  __builtin_memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t mant = x & 0x007FFFFFu;
  int exp = int((x >> 23) & 0xFF);
  if (exp == 255) return uint16_t(sign | 0x7C00u | (mant ? 0x200u | (mant >> 13) : 0));
  int e = exp - 127 + 15;
  if (e >= 31) return uint16_t(sign | 0x7C00u); // overflow -> inf
  if (e <= 0) {
    if (e < -10) return uint16_t(sign); // underflow -> signed zero
    mant |= 0x00800000u;
    int shift = 14 - e; // 14..24
    uint32_t half = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half & 1))) ++half;
    return uint16_t(sign | half);
  }
  uint32_t half = (uint32_t(e) << 10) | (mant >> 13);
  uint32_t rem = mant & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) ++half; // may carry into exponent: correct
  return uint16_t(sign | half);
}
inline float halfToFloat(uint16_t h) {
  uint32_t sign = uint32_t(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1F, mant = h & 0x3FFu, x;
  if (exp == 0) {
    if (mant == 0) {
      x = sign;
    } else {
      int e = -1;
      do {
        ++e;
        mant <<= 1;
      } while (!(mant & 0x400u));
      x = sign | uint32_t(127 - 15 - e) << 23 | (mant & 0x3FFu) << 13;
    }
  } else if (exp == 31) {
    x = sign | 0x7F800000u | (mant << 13);
  } else {
    x = sign | (exp + 127 - 15) << 23 | (mant << 13);
  }
  float f;
  __builtin_memcpy(&f, &x, 4);
  return f;
}

} // namespace orc
