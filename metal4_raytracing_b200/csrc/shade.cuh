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
#ifdef RT_FAST_SHADE
// fast-shade build (csrc/Makefile `fast`): float library functions (<= 2 ulp) instead of double evaluation rounded once
__device__ __forceinline__ float sinDet(float x) { return sinf(x); }
__device__ __forceinline__ float cosDet(float x) { return cosf(x); }
__device__ __forceinline__ float atan2Det(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ float acosDet(float x) { return acosf(x); }
#else
__device__ __forceinline__ float sinDet(float x) { return float(sin(double(x))); }
__device__ __forceinline__ float cosDet(float x) { return float(cos(double(x))); }
__device__ __forceinline__ float atan2Det(float y, float x) { return float(atan2(double(y), double(x))); }
__device__ __forceinline__ float acosDet(float x) { return float(acos(double(x))); }
#endif

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
// {prime, magic, shift, bits of 1.0f / float(prime)}: i / prime == __umulhi(i, magic) >> shift for 0 <= i < 2^31
// (round-up reciprocal, Granlund & Montgomery 1994; generated and checked by the script in the comment of halton())
__constant__ uint4 c_haltonBase[100] = {
    {2u, 0x80000001u, 0u, 0x3F000000u}, {3u, 0xAAAAAAABu, 1u, 0x3EAAAAABu}, {5u, 0xCCCCCCCDu, 2u, 0x3E4CCCCDu},
    {7u, 0x92492493u, 2u, 0x3E124925u}, {11u, 0xBA2E8BA3u, 3u, 0x3DBA2E8Cu}, {13u, 0x9D89D89Eu, 3u, 0x3D9D89D9u},
    {17u, 0xF0F0F0F1u, 4u, 0x3D70F0F1u}, {19u, 0xD79435E6u, 4u, 0x3D579436u}, {23u, 0xB21642C9u, 4u, 0x3D321643u},
    {29u, 0x8D3DCB09u, 4u, 0x3D0D3DCBu}, {31u, 0x84210843u, 4u, 0x3D042108u}, {37u, 0xDD67C8A7u, 5u, 0x3CDD67C9u},
    {41u, 0xC7CE0C7Du, 5u, 0x3CC7CE0Cu}, {43u, 0xBE82FA0Cu, 5u, 0x3CBE82FAu}, {47u, 0xAE4C415Du, 5u, 0x3CAE4C41u},
    {53u, 0x9A90E7DAu, 5u, 0x3C9A90E8u}, {59u, 0x8AD8F2FCu, 5u, 0x3C8AD8F3u}, {61u, 0x864B8A7Eu, 5u, 0x3C864B8Au},
    {67u, 0xF4898D60u, 6u, 0x3C74898Du}, {71u, 0xE6C2B449u, 6u, 0x3C66C2B4u}, {73u, 0xE070381Du, 6u, 0x3C607038u},
    {79u, 0xCF6474A9u, 6u, 0x3C4F6475u}, {83u, 0xC565C87Cu, 6u, 0x3C4565C8u}, {89u, 0xB81702E1u, 6u, 0x3C381703u},
    {97u, 0xA8E83F58u, 6u, 0x3C28E83Fu}, {101u, 0xA237C32Cu, 6u, 0x3C2237C3u}, {103u, 0x9F1165E8u, 6u, 0x3C1F1166u},
    {107u, 0x991F1A52u, 6u, 0x3C191F1Au}, {109u, 0x964FDA6Du, 6u, 0x3C164FDAu}, {113u, 0x90FDBC0Au, 6u, 0x3C10FDBCu},
    {127u, 0x81020409u, 6u, 0x3C010204u}, {131u, 0xFA232CF3u, 7u, 0x3BFA232Du}, {137u, 0xEF2EB720u, 7u, 0x3BEF2EB7u},
    {139u, 0xEBBDB2A6u, 7u, 0x3BEBBDB3u}, {149u, 0xDBEB61EFu, 7u, 0x3BDBEB62u}, {151u, 0xD901B204u, 7u, 0x3BD901B2u},
    {157u, 0xD0B69FCCu, 7u, 0x3BD0B6A0u}, {163u, 0xC907DA4Fu, 7u, 0x3BC907DAu}, {167u, 0xC4372F86u, 7u, 0x3BC43730u},
    {173u, 0xBD691048u, 7u, 0x3BBD6910u}, {179u, 0xB70FBB5Bu, 7u, 0x3BB70FBBu}, {181u, 0xB509E68Bu, 7u, 0x3BB509E7u},
    {191u, 0xAB8F69E3u, 7u, 0x3BAB8F6Au}, {193u, 0xA9C84A48u, 7u, 0x3BA9C84Au}, {197u, 0xA655C43Au, 7u, 0x3BA655C4u},
    {199u, 0xA4A9CF1Eu, 7u, 0x3BA4A9CFu}, {211u, 0x9B4C6F9Fu, 7u, 0x3B9B4C70u}, {223u, 0x92F11385u, 7u, 0x3B92F114u},
    {227u, 0x905A3864u, 7u, 0x3B905A38u}, {229u, 0x8F1779DAu, 7u, 0x3B8F177Au}, {233u, 0x8CA29C05u, 7u, 0x3B8CA29Cu},
    {239u, 0x891AC73Bu, 7u, 0x3B891AC7u}, {241u, 0x87F78088u, 7u, 0x3B87F781u}, {251u, 0x828CBFBFu, 7u, 0x3B828CC0u},
    {257u, 0xFF00FF01u, 8u, 0x3B7F00FFu}, {263u, 0xF92FB222u, 8u, 0x3B792FB2u}, {269u, 0xF3A0D52Du, 8u, 0x3B73A0D5u},
    {271u, 0xF1D48BCFu, 8u, 0x3B71D48Cu}, {277u, 0xEC979119u, 8u, 0x3B6C9791u}, {281u, 0xE9396520u, 8u, 0x3B693965u},
    {283u, 0xE79372E3u, 8u, 0x3B679373u}, {293u, 0xDFAC1F75u, 8u, 0x3B5FAC1Fu}, {307u, 0xD578E97Du, 8u, 0x3B5578E9u},
    {311u, 0xD2BA083Cu, 8u, 0x3B52BA08u}, {313u, 0xD161543Fu, 8u, 0x3B516154u}, {317u, 0xCEBCF8BCu, 8u, 0x3B4EBCF9u},
    {331u, 0xC5FE7404u, 8u, 0x3B45FE74u}, {337u, 0xC2780614u, 8u, 0x3B427806u}, {347u, 0xBCDD535Eu, 8u, 0x3B3CDD53u},
    {349u, 0xBBC8408Du, 8u, 0x3B3BC841u}, {353u, 0xB9A7862Bu, 8u, 0x3B39A786u}, {359u, 0xB68D3135u, 8u, 0x3B368D31u},
    {367u, 0xB2927C2Au, 8u, 0x3B32927Cu}, {373u, 0xAFB321A2u, 8u, 0x3B2FB322u}, {379u, 0xACEB0F8Au, 8u, 0x3B2CEB10u},
    {383u, 0xAB1CBDD4u, 8u, 0x3B2B1CBEu}, {389u, 0xA8791709u, 8u, 0x3B287917u}, {397u, 0xA513FD6Cu, 8u, 0x3B2513FDu},
    {401u, 0xA36E71A3u, 8u, 0x3B236E72u}, {409u, 0xA03C1689u, 8u, 0x3B203C17u}, {419u, 0x9C69169Cu, 8u, 0x3B1C6917u},
    {421u, 0x9BAADE8Fu, 8u, 0x3B1BAADFu}, {431u, 0x980E4157u, 8u, 0x3B180E41u}, {433u, 0x975A7510u, 8u, 0x3B175A75u},
    {439u, 0x9548E498u, 8u, 0x3B1548E5u}, {443u, 0x93EFD1C6u, 8u, 0x3B13EFD2u}, {449u, 0x91F5BCB9u, 8u, 0x3B11F5BDu},
    {457u, 0x8F67A1E4u, 8u, 0x3B0F67A2u}, {461u, 0x8E2917E1u, 8u, 0x3B0E2918u}, {463u, 0x8D8BE340u, 8u, 0x3B0D8BE3u},
    {467u, 0x8C55841Du, 8u, 0x3B0C5584u}, {479u, 0x88D180CEu, 8u, 0x3B08D181u}, {487u, 0x869222B2u, 8u, 0x3B069223u},
    {491u, 0x85797B92u, 8u, 0x3B05797Cu}, {499u, 0x8355ACE4u, 8u, 0x3B0355ADu}, {503u, 0x824A4E61u, 8u, 0x3B024A4Eu},
    {509u, 0x80C121B3u, 8u, 0x3B00C122u}, {521u, 0xFB93E673u, 9u, 0x3AFB93E6u}, {523u, 0xFA9D9D20u, 9u, 0x3AFA9D9Du},
    {541u, 0xF246FACCu, 9u, 0x3AF246FBu}};

// Same value, bit for bit, as the reference loop `f *= 1/b; r += f * (i % b); i /= b` (Raytracing.metal:42-57): the
// quotient comes from a multiply-high with a per-prime reciprocal instead of an integer division by a runtime
// divisor (the dominant cost of the shading kernel before), the float accumulation is unchanged.
// Table: tools/gen_halton_table.py.
__device__ __forceinline__ float halton(int i, int d) {
  // base 2 (dimension 0, the pixel jitter of every path): the digit loop is a bit reversal. Every partial sum of the
  // loop is exactly representable while i < 2^24, so float(brev(i)) * 2^-32 is the same float, without 21 iterations.
  if (d == 0 && i > 0 && i < (1 << 24)) return float(__brev(uint32_t(i))) * 2.3283064365386963e-10f;
  const uint4 base = c_haltonBase[((d % 100) + 100) % 100]; // the reference indexes past the table for d > 99 (F9)
  const float invB = __uint_as_float(base.w);
  float f = 1.0f;
  float r = 0.0f;
  uint32_t n = i > 0 ? uint32_t(i) : 0u;
  // Two digits per trip. A digit taken after n has reached 0 is 0 and adds f * 0 = 0, which leaves r untouched, so
  // running past the last digit changes nothing and the exit test is only needed every other digit. The digit goes
  // to float through the 2^23 trick (exact below 2^23; LOP3 + FADD) instead of an I2F, which issues at quarter rate.
  while (n != 0u) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint32_t q = __umulhi(n, base.y) >> base.z;
      const uint32_t digit = n - q * base.x;
      f = f * invB;
      r = r + f * (__uint_as_float(0x4B000000u | digit) - 8388608.0f);
      n = q;
    }
  }
  return r;
}

// ---- environment (extension, rt_b200.h rt_environment) -----------------------------------------------------------
// atan2 / acos are evaluated in double and rounded once, like sin / cos above, so the CPU oracle gets the same
// floats; the bilinear blend uses the same a + (b - a) t form as the material textures.
constexpr float kInvTwoPi = 0.15915494309189535f, kInvPi = 0.3183098861837907f;
__device__ __forceinline__ f3 sampleEnvironment(const rt_environment &env, f3 d) {
  const float phi = atan2Det(d.z, d.x);
  const float theta = acosDet(clampf(d.y, -1.0f, 1.0f));
  const float u = phi * kInvTwoPi + 0.5f, v = theta * kInvPi;
  const float x = u * float(env.width) - 0.5f, y = v * float(env.height) - 0.5f;
  const float fx = floorf(x), fy = floorf(y);
  const float tx = x - fx, ty = y - fy;
  int x0 = int(fx), y0 = int(fy);
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = ((x0 % env.width) + env.width) % env.width;
  x1 = ((x1 % env.width) + env.width) % env.width;
  y0 = min(max(y0, 0), env.height - 1);
  y1 = min(max(y1, 0), env.height - 1);
  const float4 *t = reinterpret_cast<const float4 *>(env.texelsDev);
  const float4 a = __ldg(t + size_t(y0) * env.width + x0), b = __ldg(t + size_t(y0) * env.width + x1);
  const float4 c = __ldg(t + size_t(y1) * env.width + x0), e = __ldg(t + size_t(y1) * env.width + x1);
  const f3 top = mix(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), tx);
  const f3 bottom = mix(mk3(c.x, c.y, c.z), mk3(e.x, e.y, e.z), tx);
  return mix(top, bottom, ty) * env.intensity;
}

// RT_ENV_IMPORTANCE (rt_b200.h): the environment as a light. Table = rt_environment_cdf's output.
constexpr float kTwoPiSquared = 19.739208802178716f;
// largest i in [0, n - 1] with c[i] <= xi (c[0] = 0, c[n] = 1 > xi): the selected cell always has weight
__device__ __forceinline__ int cdfFind(const float *__restrict__ c, int n, float xi) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(c + mid) <= xi) lo = mid;
    else hi = mid;
  }
  return lo;
}
// The same cell, found from a narrower start (RT_ENV_GUIDED): guide[k] = largest i with c[i] <= k / 64, so for xi in
// bucket k, c[guide[k]] <= xi and c[guide[k + 1] + 1] > xi — the invariant of the search above, entered later.
__device__ __forceinline__ int cdfFindGuided(const float *__restrict__ c, int n, float xi, const float *__restrict__ guide) {
  const int k = min(int(xi * float(RT_ENV_GUIDE_CELLS)), RT_ENV_GUIDE_CELLS - 1);
  int lo = int(__float_as_uint(__ldg(guide + k))), hi = min(int(__float_as_uint(__ldg(guide + k + 1))) + 1, n);
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(c + mid) <= xi) lo = mid;
    else hi = mid;
  }
  return lo;
}
// solid-angle density of the texel (x, y) at polar sine sinTheta
__device__ __forceinline__ float environmentTexelPdf(const rt_environment &env, int x, int y, float sinTheta) {
  const float *marginal = env.cdfDev;
  const float *row = env.cdfDev + (env.height + 1) + size_t(y) * size_t(env.width + 1);
  const float pr = (__ldg(marginal + y + 1) - __ldg(marginal + y)) * float(env.height);
  const float pc = (__ldg(row + x + 1) - __ldg(row + x)) * float(env.width);
  return (pr * pc) / (kTwoPiSquared * fmaxf(sinTheta, 1e-6f));
}
__device__ __forceinline__ float environmentPdf(const rt_environment &env, f3 d) {
  const float phi = atan2Det(d.z, d.x);
  const float theta = acosDet(clampf(d.y, -1.0f, 1.0f));
  const float u = phi * kInvTwoPi + 0.5f, v = theta * kInvPi;
  const int x = min(max(int(floorf(u * float(env.width))), 0), env.width - 1);
  const int y = min(max(int(floorf(v * float(env.height))), 0), env.height - 1);
  return environmentTexelPdf(env, x, y, sinDet(theta));
}
// direction drawn from the table with (xi.x -> row, xi.y -> column); returns its density
__device__ __forceinline__ float sampleEnvironmentDirection(const rt_environment &env, f2 xi, f3 &dir) {
  const float *marginal = env.cdfDev;
  const bool guided = (env.flags & RT_ENV_GUIDED) != 0u;
  const float *guides = env.cdfDev + (env.height + 1) + size_t(env.height) * size_t(env.width + 1);
  const int y = guided ? cdfFindGuided(marginal, env.height, xi.x, guides) : cdfFind(marginal, env.height, xi.x);
  const float m0 = __ldg(marginal + y), m1 = __ldg(marginal + y + 1);
  const float dy = (xi.x - m0) / (m1 - m0);
  const float *row = env.cdfDev + (env.height + 1) + size_t(y) * size_t(env.width + 1);
  const int x = guided ? cdfFindGuided(row, env.width, xi.y, guides + size_t(y + 1) * size_t(RT_ENV_GUIDE_CELLS + 1))
                       : cdfFind(row, env.width, xi.y);
  const float c0 = __ldg(row + x), c1 = __ldg(row + x + 1);
  const float dx = (xi.y - c0) / (c1 - c0);
  const float u = (float(x) + dx) / float(env.width), v = (float(y) + dy) / float(env.height);
  const float phi = (u - 0.5f) * (2.0f * kPi), theta = v * kPi;
  const float sinTheta = sinDet(theta), cosTheta = cosDet(theta);
  dir = mk3(sinTheta * cosDet(phi), cosTheta, sinTheta * sinDet(phi));
  return environmentTexelPdf(env, x, y, sinTheta);
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
