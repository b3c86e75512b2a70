// oracle_kernels.cpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Statement-by-statement C++ restatement of the two Metal kernels on the hot path:
//   raytracingKernel   MetalRaytracing/Raytracing.metal:220-831 (helpers :28-218)
//   skinningKernel     MetalRaytracing/Skinning.metal:7-49
// Quirks kept on purpose: miss => black (F5), EMA accumulation with <= 0.95 history (F7),
// stride-5/stride-6 Halton dimension mismatch (F9; prime index taken mod 100 where the reference would read
// out of bounds), maxBounces counts closest-hit segments (F13), shadow ray traced even when NdotL == 0,
// a bound normal map doubles as the opacity map. motionTex is zero-initialised by the caller (F12).
// PARITY UNPINNED: the reference has no tests or golden vectors and cannot run here (Swift/Metal only).
#include <atomic>
#include <cmath>
#include <cstring>

#include "oracle_internal.h"

namespace orc {

// ---- halton (Raytracing.metal:28-57) ---------------------------------------------------------------------
static const short kPrimes[100] = {
    2,   3,   5,   7,   11,  13,  17,  19,  23,  29,  31,  37,  41,  43,  47,  53,  59,  61,  67,  71,
    73,  79,  83,  89,  97,  101, 103, 107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173,
    179, 181, 191, 193, 197, 199, 211, 223, 227, 229, 233, 239, 241, 251, 257, 263, 269, 271, 277, 281,
    283, 293, 307, 311, 313, 317, 331, 337, 347, 349, 353, 359, 367, 373, 379, 383, 389, 397, 401, 409,
    419, 421, 431, 433, 439, 443, 449, 457, 461, 463, 467, 479, 487, 491, 499, 503, 509, 521, 523, 541};

float halton(int i, int d) {
  short b = kPrimes[((d % 100) + 100) % 100];
  float f = 1.0f;
  float invB = 1.0f / float(b);
  float r = 0;
  while (i > 0) {
    f = f * invB;
    r = r + f * float(i % b);
    i = i / b;
  }
  return r;
}

// ---- texture2d<float>::sample with linear filter, repeat addressing, LOD 0 (Raytracing.metal:421) --------
static float g_srgbLut[256];
static std::atomic<bool> g_srgbReady{false};
static void initSrgb() {
  if (g_srgbReady.load()) return;
  for (int i = 0; i < 256; ++i) {
    double c = double(i) / 255.0;
    g_srgbLut[i] = float(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
  }
  g_srgbReady.store(true);
}

static inline float4 fetchTexel(const rt_texture2d *t, int x, int y) {
  const uint8_t *p = t->texels + (size_t(y) * size_t(t->width) + size_t(x)) * 4;
  float4 c;
  if (t->srgb) {
    c.x = g_srgbLut[p[0]];
    c.y = g_srgbLut[p[1]];
    c.z = g_srgbLut[p[2]];
  } else {
    c.x = float(p[0]) / 255.0f;
    c.y = float(p[1]) / 255.0f;
    c.z = float(p[2]) / 255.0f;
  }
  c.w = float(p[3]) / 255.0f;
  return c;
}

static inline int wrapIndex(int i, int n) {
  int m = i % n;
  return m < 0 ? m + n : m;
}

float4 sampleTexture(const rt_texture2d *t, float2 uv) {
  initSrgb();
  float x = uv.x * float(t->width) - 0.5f, y = uv.y * float(t->height) - 0.5f;
  x = fminf(fmaxf(x, -1.0e9f), 1.0e9f);
  y = fminf(fmaxf(y, -1.0e9f), 1.0e9f);
  float fx0 = floorf(x), fy0 = floorf(y);
  float fx = x - fx0, fy = y - fy0;
  int x0 = wrapIndex(int(fx0), t->width), y0 = wrapIndex(int(fy0), t->height);
  int x1 = wrapIndex(x0 + 1, t->width), y1 = wrapIndex(y0 + 1, t->height);
  float4 t00 = fetchTexel(t, x0, y0), t10 = fetchTexel(t, x1, y0), t01 = fetchTexel(t, x0, y1),
         t11 = fetchTexel(t, x1, y1);
  auto lerp2 = [&](float a, float b, float c, float d) {
    return (a * (1.0f - fx) + b * fx) * (1.0f - fy) + (c * (1.0f - fx) + d * fx) * fy;
  };
  return {lerp2(t00.x, t10.x, t01.x, t11.x), lerp2(t00.y, t10.y, t01.y, t11.y), lerp2(t00.z, t10.z, t01.z, t11.z),
          lerp2(t00.w, t10.w, t01.w, t11.w)};
}

// ---- image read / write in the bound format --------------------------------------------------------------
static inline float4 readImage(const rt_image &img, int x, int y) {
  size_t i = size_t(y) * size_t(img.width) + size_t(x);
  switch (img.format) {
    case RT_FORMAT_RGBA16_FLOAT: {
      const uint16_t *p = static_cast<const uint16_t *>(img.data) + i * 4;
      return {halfToFloat(p[0]), halfToFloat(p[1]), halfToFloat(p[2]), halfToFloat(p[3])};
    }
    case RT_FORMAT_RGBA32_FLOAT: {
      const float *p = static_cast<const float *>(img.data) + i * 4;
      return {p[0], p[1], p[2], p[3]};
    }
    case RT_FORMAT_RG16_FLOAT: {
      const uint16_t *p = static_cast<const uint16_t *>(img.data) + i * 2;
      return {halfToFloat(p[0]), halfToFloat(p[1]), 0, 1};
    }
    case RT_FORMAT_RG32_FLOAT: {
      const float *p = static_cast<const float *>(img.data) + i * 2;
      return {p[0], p[1], 0, 1};
    }
    case RT_FORMAT_R32_FLOAT:
      return {static_cast<const float *>(img.data)[i], 0, 0, 1};
    case RT_FORMAT_R16_FLOAT:
      return {halfToFloat(static_cast<const uint16_t *>(img.data)[i]), 0, 0, 1};
    default:
      return {0, 0, 0, 0};
  }
}

static inline void writeImage(const rt_image &img, int x, int y, float4 v) {
  if (!img.data) return;
  size_t i = size_t(y) * size_t(img.width) + size_t(x);
  switch (img.format) {
    case RT_FORMAT_RGBA16_FLOAT: {
      uint16_t *p = static_cast<uint16_t *>(img.data) + i * 4;
      p[0] = floatToHalf(v.x), p[1] = floatToHalf(v.y), p[2] = floatToHalf(v.z), p[3] = floatToHalf(v.w);
      break;
    }
    case RT_FORMAT_RGBA32_FLOAT: {
      float *p = static_cast<float *>(img.data) + i * 4;
      p[0] = v.x, p[1] = v.y, p[2] = v.z, p[3] = v.w;
      break;
    }
    case RT_FORMAT_RG16_FLOAT: {
      uint16_t *p = static_cast<uint16_t *>(img.data) + i * 2;
      p[0] = floatToHalf(v.x), p[1] = floatToHalf(v.y);
      break;
    }
    case RT_FORMAT_RG32_FLOAT: {
      float *p = static_cast<float *>(img.data) + i * 2;
      p[0] = v.x, p[1] = v.y;
      break;
    }
    case RT_FORMAT_R32_FLOAT:
      static_cast<float *>(img.data)[i] = v.x;
      break;
    case RT_FORMAT_R16_FLOAT:
      static_cast<uint16_t *>(img.data)[i] = floatToHalf(v.x);
      break;
    default:
      break;
  }
}

// ---- helpers (Raytracing.metal:59-218) --------------------------------------------------------------------
static inline float3 ld3(const rt_float3 *a, unsigned i) { return {a[i].x, a[i].y, a[i].z}; }

// interpolateVertexAttribute<float3> (Raytracing.metal:61-74)
static inline float3 interpolateFloat3(const rt_float3 *attributes, const Hit &hit, const int32_t *vertexIndices) {
  float3 uvw;
  uvw.x = hit.u;
  uvw.y = hit.v;
  uvw.z = 1.0f - uvw.x - uvw.y;
  unsigned triangleIndex = hit.primitive;
  unsigned index1 = unsigned(vertexIndices[triangleIndex * 3 + 1]);
  unsigned index2 = unsigned(vertexIndices[triangleIndex * 3 + 2]);
  unsigned index3 = unsigned(vertexIndices[triangleIndex * 3 + 0]);
  float3 T0 = ld3(attributes, index1), T1 = ld3(attributes, index2), T2 = ld3(attributes, index3);
  return uvw.x * T0 + uvw.y * T1 + uvw.z * T2;
}

// interpolateVertexAttribute<float2> (Raytracing.metal:61-74)
static inline float2 interpolateFloat2(const float *attributes, const Hit &hit, const int32_t *vertexIndices) {
  float ux = hit.u, uy = hit.v, uz = 1.0f - ux - uy;
  unsigned triangleIndex = hit.primitive;
  unsigned index1 = unsigned(vertexIndices[triangleIndex * 3 + 1]);
  unsigned index2 = unsigned(vertexIndices[triangleIndex * 3 + 2]);
  unsigned index3 = unsigned(vertexIndices[triangleIndex * 3 + 0]);
  float2 T0{attributes[2 * index1], attributes[2 * index1 + 1]}, T1{attributes[2 * index2], attributes[2 * index2 + 1]},
      T2{attributes[2 * index3], attributes[2 * index3 + 1]};
  return ux * T0 + uy * T1 + uz * T2;
}

// Raytracing.metal:79-89
static inline float3 sampleCosineWeightedHemisphere(float2 u) {
  float phi = 2.0f * kPi * u.x;
  float cos_phi = cos_det(phi);
  float sin_phi = sin_det(phi);
  float cos_theta = sqrtf(u.y);
  float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
  return {sin_theta * cos_phi, cos_theta, sin_theta * sin_phi};
}

// Raytracing.metal:95-129
static inline void sampleAreaLight(const rt_light &light, float2 u, float3 position, float3 &lightDirection,
                                   float3 &lightColor, float &lightDistance) {
  u = u * 2.0f - make2(1.0f, 1.0f);
  float3 lp{light.position.x, light.position.y, light.position.z}, lr{light.right.x, light.right.y, light.right.z},
      lu{light.up.x, light.up.y, light.up.z}, lf{light.forward.x, light.forward.y, light.forward.z};
  float3 samplePosition = lp + lr * u.x + lu * u.y;
  lightDirection = samplePosition - position;
  lightDistance = length(lightDirection);
  float inverseLightDistance = 1.0f / fmaxf(lightDistance, 1e-3f);
  lightDirection *= inverseLightDistance;
  lightColor = make3(light.color.x, light.color.y, light.color.z);
  lightColor *= (inverseLightDistance * inverseLightDistance);
  lightColor *= saturate(dot(-lightDirection, lf));
}

// Raytracing.metal:133-148
static inline float3 alignHemisphereWithNormal(float3 sample, float3 normal) {
  float3 up = normal;
  float3 right = normalize(cross(normal, make3(0.0072f, 1.0f, 0.0034f)));
  float3 forward = cross(right, up);
  return sample.x * right + sample.y * up + sample.z * forward;
}

// Raytracing.metal:150-154
static inline float distributionGGX(float NdotH, float alpha) {
  float a2 = alpha * alpha;
  float denom = (NdotH * NdotH) * (a2 - 1.0f) + 1.0f;
  return a2 / fmaxf(kPi * denom * denom, 1e-7f);
}
// Raytracing.metal:156-158
static inline float geometrySchlickGGX(float NdotV, float k) { return NdotV / fmaxf(NdotV * (1.0f - k) + k, 1e-7f); }
// Raytracing.metal:160-162
static inline float geometrySmith(float NdotV, float NdotL, float k) {
  return geometrySchlickGGX(NdotV, k) * geometrySchlickGGX(NdotL, k);
}
// Raytracing.metal:164-166
static inline float3 fresnelSchlick(float cosTheta, float3 F0) {
  return F0 + (1.0f - F0) * pow5(clampf(1.0f - cosTheta, 0.0f, 1.0f));
}

// Raytracing.metal:185-218
static inline bool computeTangentBasis(const rt_float3 *positions, const float *uvs, const Hit &hit,
                                       const int32_t *vertexIndices, float3 &tangent, float3 &bitangent) {
  unsigned triangleIndex = hit.primitive;
  unsigned index1 = unsigned(vertexIndices[triangleIndex * 3 + 1]);
  unsigned index2 = unsigned(vertexIndices[triangleIndex * 3 + 2]);
  unsigned index3 = unsigned(vertexIndices[triangleIndex * 3 + 0]);
  float3 p0 = ld3(positions, index1), p1 = ld3(positions, index2), p2 = ld3(positions, index3);
  float2 uv0{uvs[2 * index1], uvs[2 * index1 + 1]}, uv1{uvs[2 * index2], uvs[2 * index2 + 1]},
      uv2{uvs[2 * index3], uvs[2 * index3 + 1]};
  float3 e1 = p1 - p0, e2 = p2 - p0;
  float2 dUV1 = uv1 - uv0, dUV2 = uv2 - uv0;
  float denom = dUV1.x * dUV2.y - dUV1.y * dUV2.x;
  if (fabsf(denom) < 1e-8f) return false;
  float r = 1.0f / denom;
  tangent = (e1 * dUV2.y - e2 * dUV1.y) * r;
  bitangent = (e2 * dUV1.x - e1 * dUV2.x) * r;
  return (length(tangent) > 1e-8f) && (length(bitangent) > 1e-8f);
}

// Raytracing.metal:324-330: 4x4 from the descriptor's packed 4x3 (columns, rows 0..2)
static inline float4x4 instanceMatrix(const rt_instance_descriptor &d) {
  float4x4 m;
  for (int c = 0; c < 4; ++c) m.c[c] = {d.transformationMatrix[c][0], d.transformationMatrix[c][1],
                                        d.transformationMatrix[c][2], c == 3 ? 1.0f : 0.0f};
  return m;
}
static inline float3 f3(const rt_float3 &v) { return {v.x, v.y, v.z}; }

// ---- raytracingKernel, one thread (Raytracing.metal:220-831) ----------------------------------------------
// Environment extension: equirectangular lookup with the arithmetic spelled out in include/rt_b200.h; mirrors
// sampleEnvironment() in csrc/shade.cuh operation for operation.
float3 sampleEnvironment(const rt_environment &env, float3 d) {
  const float kInvTwoPi = 0.15915494309189535f, kInvPi = 0.3183098861837907f;
  const float phi = static_cast<float>(std::atan2(static_cast<double>(d.z), static_cast<double>(d.x)));
  const float theta = static_cast<float>(std::acos(static_cast<double>(clampf(d.y, -1.0f, 1.0f))));
  const float u = phi * kInvTwoPi + 0.5f, v = theta * kInvPi;
  const float x = u * float(env.width) - 0.5f, y = v * float(env.height) - 0.5f;
  const float fx = std::floor(x), fy = std::floor(y);
  const float tx = x - fx, ty = y - fy;
  int x0 = int(fx), y0 = int(fy);
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = ((x0 % env.width) + env.width) % env.width;
  x1 = ((x1 % env.width) + env.width) % env.width;
  y0 = std::min(std::max(y0, 0), env.height - 1);
  y1 = std::min(std::max(y1, 0), env.height - 1);
  auto texel = [&](int xx, int yy) {
    const float *t = env.texelsDev + (size_t(yy) * env.width + xx) * 4;
    return make3(t[0], t[1], t[2]);
  };
  const float3 top = mix(texel(x0, y0), texel(x1, y0), tx);
  const float3 bottom = mix(texel(x0, y1), texel(x1, y1), tx);
  return mix(top, bottom, ty) * env.intensity;
}

// RT_ENV_IMPORTANCE (include/rt_b200.h): the environment as one more light; mirrors cdfFind / environmentTexelPdf /
// environmentPdf / sampleEnvironmentDirection of csrc/shade.cuh operation for operation.
static inline bool environmentIsLight(const rt_environment &env) {
  return (env.flags & RT_ENV_IMPORTANCE) != 0u && env.cdfDev != nullptr && env.texelsDev != nullptr;
}
static inline int cdfFind(const float *c, int n, float xi) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (c[mid] <= xi) lo = mid;
    else hi = mid;
  }
  return lo;
}
static inline float environmentTexelPdf(const rt_environment &env, int x, int y, float sinTheta) {
  const float kTwoPiSquared = 19.739208802178716f;
  const float *marginal = env.cdfDev;
  const float *row = env.cdfDev + (env.height + 1) + size_t(y) * size_t(env.width + 1);
  const float pr = (marginal[y + 1] - marginal[y]) * float(env.height);
  const float pc = (row[x + 1] - row[x]) * float(env.width);
  return (pr * pc) / (kTwoPiSquared * fmaxf(sinTheta, 1e-6f));
}
static inline float environmentPdf(const rt_environment &env, float3 d) {
  const float kInvTwoPi = 0.15915494309189535f, kInvPi = 0.3183098861837907f;
  const float phi = static_cast<float>(std::atan2(static_cast<double>(d.z), static_cast<double>(d.x)));
  const float theta = static_cast<float>(std::acos(static_cast<double>(clampf(d.y, -1.0f, 1.0f))));
  const float u = phi * kInvTwoPi + 0.5f, v = theta * kInvPi;
  const int x = std::min(std::max(int(std::floor(u * float(env.width))), 0), env.width - 1);
  const int y = std::min(std::max(int(std::floor(v * float(env.height))), 0), env.height - 1);
  return environmentTexelPdf(env, x, y, sin_det(theta));
}
static inline float sampleEnvironmentDirection(const rt_environment &env, float2 xi, float3 &dir) {
  const float *marginal = env.cdfDev;
  const int y = cdfFind(marginal, env.height, xi.x);
  const float m0 = marginal[y], m1 = marginal[y + 1];
  const float dy = (xi.x - m0) / (m1 - m0);
  const float *row = env.cdfDev + (env.height + 1) + size_t(y) * size_t(env.width + 1);
  const int x = cdfFind(row, env.width, xi.y);
  const float c0 = row[x], c1 = row[x + 1];
  const float dx = (xi.y - c0) / (c1 - c0);
  const float u = (float(x) + dx) / float(env.width), v = (float(y) + dy) / float(env.height);
  const float phi = (u - 0.5f) * (2.0f * kPi), theta = v * kPi;
  const float sinTheta = sin_det(theta), cosTheta = cos_det(theta);
  dir = make3(sinTheta * cos_det(phi), cosTheta, sinTheta * sin_det(phi));
  return environmentTexelPdf(env, x, y, sinTheta);
}

void raytracingKernelPixel(int tidx, int tidy, const KernelArgs &a, PixelStats &stats) {
  const rt_uniforms &uniforms = *a.uniforms;
  if (!(tidx < uniforms.width && tidy < uniforms.height)) return;
  unsigned int offset = static_cast<const uint32_t *>(a.textures[RT_TEXTURE_RANDOM].data)[size_t(tidy) * uniforms.width + tidx];

  float3 totalColor = make3(0.0f);
  float4 pm = readImage(a.textures[RT_TEXTURE_MOTION], tidx, tidy);
  float2 prevMotion{pm.x, pm.y};

  float primaryDepth = 1.0e8f;
  float2 motionVector{0.0f, 0.0f};
  bool hadPrimaryHit = false;

  float4 outDiffuseAlbedo{0, 0, 0, 0}, outSpecularAlbedo{0, 0, 0, 0}, outNormal{0, 0, 0, 0}, outRoughness{0, 0, 0, 0};
  bool wroteGBuffer = false;

  int baseSamples = std::max(uniforms.samplesPerPixel, 1);
  int maxExtraSamples = (uniforms.enableMotionAdaptiveSampling != 0) ? std::max(uniforms.motionSamplingMaxExtraSamples, 0) : 0;
  int sampleStride = baseSamples + maxExtraSamples;
  int totalSamples = baseSamples;

  for (int sampleIndex = 0; sampleIndex < totalSamples; sampleIndex++) {
    int frameOffset = int(uniforms.frameIndex * unsigned(sampleStride) + unsigned(sampleIndex));
    int haltonIndex = int(offset + unsigned(frameOffset));

    float2 r{halton(haltonIndex, 0), halton(haltonIndex, 1)};
    float2 samplePixel = make2(float(tidx), float(tidy)) + r;
    float2 uv = samplePixel / make2(float(uniforms.width), float(uniforms.height));
    uv = uv * 2.0f - make2(1.0f, 1.0f);

    const rt_camera &camera = uniforms.camera;
    float3 rayOrigin = f3(camera.position);
    float3 rayDirection = normalize(uv.x * f3(camera.right) + uv.y * f3(camera.up) + f3(camera.forward));
    float rayMax = INFINITY;

    float3 color = make3(1.0f);
    float3 accumulatedColor = make3(0.0f);

    int bounce = 0;
    int step = 0;
    int transparencyPasses = 0;
    const bool envLight = environmentIsLight(a.env); // extension: the environment is light number lightCount
    const float kInvPi = 0.3183098861837907f;
    float bsdfPdf = 0.0f; // density of the cosine bounce that produced rayDirection; 0 = camera ray or glass
    while (bounce < uniforms.maxBounces) {
      ++stats.closestRays;
      Hit intersection = traceClosest(*a.tlas, rayOrigin, rayDirection, 0.0f, rayMax);
      if (a.primaryIds && sampleIndex == 0 && step == 0 && bounce == 0 && transparencyPasses == 0) {
        uint32_t *pid = a.primaryIds + (size_t(tidy) * uniforms.width + tidx) * 4;
        if (intersection.valid) {
          pid[0] = intersection.instance, pid[1] = intersection.geometry, pid[2] = intersection.primitive;
          std::memcpy(&pid[3], &intersection.t, 4);
        } else {
          pid[0] = pid[1] = pid[2] = pid[3] = 0xFFFFFFFFu;
        }
      }
      if (!intersection.valid) {
        // extension (include/rt_b200.h rt_environment); the reference only breaks here (Raytracing.metal:320-322)
        if (a.env.texelsDev) {
          if (envLight && bsdfPdf > 0.0f) { // balance heuristic against the light sample of the previous hit
            const float envPdf = environmentPdf(a.env, rayDirection) / float(uniforms.lightCount + 1);
            const float weight = bsdfPdf / (bsdfPdf + envPdf);
            accumulatedColor = accumulatedColor + color * (sampleEnvironment(a.env, rayDirection) * weight);
          } else {
            accumulatedColor = accumulatedColor + color * sampleEnvironment(a.env, rayDirection);
          }
        }
        break;
      }
      ++stats.hits;

      int instanceIndex = int(intersection.instance);
      int geometryIndex2 = int(intersection.geometry);
      float4x4 objectToWorldSpaceTransform = instanceMatrix(a.instances[instanceIndex]);

      float3 worldSpaceIntersectionPoint = rayOrigin + rayDirection * intersection.t;
      int resourceIndex = instanceIndex * a.maxSubmeshes + geometryIndex2;
      const rt_resource &resource = a.resources[resourceIndex];

      if (bounce == 0 && sampleIndex == 0) {
        float3 objectSpacePos = interpolateFloat3(resource.positions, intersection, resource.indices);
        float3 prevObjectSpacePos = interpolateFloat3(resource.previousPositions, intersection, resource.indices);
        float3 worldPos = mulPoint(objectToWorldSpaceTransform, objectSpacePos);
        float4x4 prevObjectToWorldSpaceTransform = instanceMatrix(a.prevInstances[instanceIndex]);
        float3 prevWorldPos = mulPoint(prevObjectToWorldSpaceTransform, prevObjectSpacePos);

        float3 viewPos = worldPos - f3(camera.position);
        float2 screenPos;
        screenPos.x = dot(viewPos, f3(camera.right));
        screenPos.y = dot(viewPos, f3(camera.up));
        float depth = dot(viewPos, f3(camera.forward));
        primaryDepth = fmaxf(depth, 1.0e-3f);
        screenPos = screenPos / fmaxf(depth, 0.001f);

        const rt_camera &prevCamera = uniforms.previousCamera;
        float3 prevViewPos = prevWorldPos - f3(prevCamera.position);
        float2 prevScreenPos;
        prevScreenPos.x = dot(prevViewPos, f3(prevCamera.right));
        prevScreenPos.y = dot(prevViewPos, f3(prevCamera.up));
        float prevDepth = dot(prevViewPos, f3(prevCamera.forward));
        prevScreenPos = prevScreenPos / fmaxf(prevDepth, 0.001f);

        float2 motionNdc = screenPos - prevScreenPos;
        float rightScale = fmaxf(length(f3(camera.right)), 1e-5f);
        float upScale = fmaxf(length(f3(camera.up)), 1e-5f);
        float2 motionPixels{motionNdc.x * (float(uniforms.width) / (2.0f * rightScale)),
                            motionNdc.y * (float(uniforms.height) / (2.0f * upScale))};
        motionPixels.y = -motionPixels.y;
        motionVector = motionPixels;
        hadPrimaryHit = true;
      }

      float3 objectSpaceSurfaceNormal = interpolateFloat3(resource.normals, intersection, resource.indices);
      float3 worldSpaceSurfaceNormal = mulDir(objectToWorldSpaceTransform, objectSpaceSurfaceNormal);
      worldSpaceSurfaceNormal = normalize(worldSpaceSurfaceNormal);
      if (length(objectSpaceSurfaceNormal) < 1e-10f) worldSpaceSurfaceNormal = -rayDirection;

      const rt_material &material = *resource.material;
      float3 albedo = f3(material.baseColor);
      uint32_t textureFlags = material.textureFlags;
      bool hasBaseColorMap = (textureFlags & RT_MATERIAL_TEXTURE_BASECOLOR) != 0;
      bool hasNormalMap = (textureFlags & RT_MATERIAL_TEXTURE_NORMAL) != 0;
      bool hasRoughnessMap = (textureFlags & RT_MATERIAL_TEXTURE_ROUGHNESS) != 0;
      bool hasMetallicMap = (textureFlags & RT_MATERIAL_TEXTURE_METALLIC) != 0;
      // #if ENABLE_AO (Raytracing.metal:405-409): a build-time switch in the reference, a per-context one here
      bool hasAOMap = a.enableAO && (textureFlags & RT_MATERIAL_TEXTURE_AO) != 0;
      bool hasOpacityMap = (textureFlags & RT_MATERIAL_TEXTURE_OPACITY) != 0;
      bool hasEmissionMap = (textureFlags & RT_MATERIAL_TEXTURE_EMISSION) != 0;

      float2 texCoord{0.0f, 0.0f};
      if (hasBaseColorMap || hasNormalMap || hasRoughnessMap || hasMetallicMap || hasAOMap || hasOpacityMap || hasEmissionMap) {
        texCoord = interpolateFloat2(resource.uvs, intersection, resource.indices);
        texCoord.y = 1.0f - texCoord.y;
      }

      float4 baseColorSample{1, 1, 1, 1};
      if (hasBaseColorMap) {
        baseColorSample = sampleTexture(resource.baseColorMap, texCoord);
        albedo *= make3(baseColorSample.x, baseColorSample.y, baseColorSample.z);
      }
      float roughness = 1.0f;
      if (hasRoughnessMap) roughness = sampleTexture(resource.roughnessMap, texCoord).x;
      float metallic = 0.0f;
      if (hasMetallicMap) metallic = sampleTexture(resource.metallicMap, texCoord).x;
      float ao = 1.0f;
      if (hasAOMap) ao = sampleTexture(resource.aoMap, texCoord).x; // #if ENABLE_AO (Raytracing.metal:442-446)
      float opacity = clampf(material.opacity, 0.0f, 1.0f);
      if (hasOpacityMap) opacity *= sampleTexture(resource.opacityMap, texCoord).x;
      float3 emission = f3(material.emission);
      if (hasEmissionMap) {
        float4 e = sampleTexture(resource.emissionMap, texCoord);
        emission = make3(e.x, e.y, e.z);
      }

      if (uniforms.debugTextureMode != RT_DEBUG_NONE) {
        float3 debugColor = make3(0.0f);
        if (uniforms.debugTextureMode == RT_DEBUG_BASECOLOR) {
          debugColor = hasBaseColorMap ? make3(baseColorSample.x, baseColorSample.y, baseColorSample.z) : make3(1.0f, 0.0f, 1.0f);
        } else if (uniforms.debugTextureMode == RT_DEBUG_NORMAL) {
          if (hasNormalMap) {
            float4 n = sampleTexture(resource.normalMap, texCoord);
            debugColor = make3(n.x, n.y, n.z);
          } else {
            debugColor = worldSpaceSurfaceNormal * 0.5f + 0.5f;
          }
        } else if (uniforms.debugTextureMode == RT_DEBUG_ROUGHNESS) {
          debugColor = make3(roughness);
        } else if (uniforms.debugTextureMode == RT_DEBUG_METALLIC) {
          debugColor = make3(metallic);
        } else if (uniforms.debugTextureMode == RT_DEBUG_AO) {
          debugColor = a.enableAO ? make3(ao) : make3(1.0f, 0.0f, 1.0f); // #if ENABLE_AO (Raytracing.metal:475-479)
        } else if (uniforms.debugTextureMode == RT_DEBUG_EMISSION) {
          debugColor = emission;
        } else if (uniforms.debugTextureMode == RT_DEBUG_MOTION) {
          float2 motionPixels = hadPrimaryHit ? motionVector : prevMotion;
          float2 scaled{clampf(motionPixels.x * 0.05f, -1.0f, 1.0f), clampf(motionPixels.y * 0.05f, -1.0f, 1.0f)};
          float mag = clampf(length(motionPixels) * 0.1f, 0.0f, 1.0f);
          debugColor = make3(scaled.x * 0.5f + 0.5f, scaled.y * 0.5f + 0.5f, mag);
        }
        accumulatedColor = debugColor;
        break;
      }

      float3 shadingNormal = worldSpaceSurfaceNormal;
      if (hasNormalMap) {
        float3 tangent, bitangent;
        if (computeTangentBasis(resource.positions, resource.uvs, intersection, resource.indices, tangent, bitangent)) {
          float3 worldT = mulDir(objectToWorldSpaceTransform, tangent);
          worldT = normalize(worldT - worldSpaceSurfaceNormal * dot(worldT, worldSpaceSurfaceNormal));
          float3 worldBOrtho = normalize(cross(worldSpaceSurfaceNormal, worldT));
          float4 ns = sampleTexture(resource.normalMap, texCoord);
          float3 nMap = make3(ns.x, ns.y, ns.z) * 2.0f - make3(1.0f);
          shadingNormal = normalize(nMap.x * worldT + nMap.y * worldBOrtho + nMap.z * worldSpaceSurfaceNormal);
        }
      }

      if (uniforms.enableDenoiseGBuffer != 0 && !wroteGBuffer && sampleIndex == 0) {
        float roughnessForOutput = clampf(roughness, 0.0f, 1.0f);
        float3 diffuseAlbedo = albedo * (1.0f - metallic);
        float3 specularAlbedo = mix(make3(0.04f), albedo, metallic);
        outDiffuseAlbedo = {diffuseAlbedo.x, diffuseAlbedo.y, diffuseAlbedo.z, 1.0f};
        outSpecularAlbedo = {specularAlbedo.x, specularAlbedo.y, specularAlbedo.z, 1.0f};
        float3 nn = shadingNormal * 0.5f + 0.5f;
        outNormal = {nn.x, nn.y, nn.z, 1.0f};
        outRoughness = {roughnessForOutput, 0.0f, 0.0f, 1.0f};
        wroteGBuffer = true;
      }

      float clampedOpacity = clampf(opacity, 0.0f, 1.0f);
      float ior = fmaxf(material.refractionIndex, 1.0f);
      bool consumeBounce = true;
      bool skipLighting = false;
      if (clampedOpacity < 0.999f || ior > 1.01f) {
        float3 N = shadingNormal;
        float3 I = rayDirection;
        float cosi = clampf(dot(-I, N), -1.0f, 1.0f);
        float etaI = 1.0f;
        float etaT = ior;
        if (cosi < 0.0f) {
          cosi = -cosi;
          N = -N;
          float tmp = etaI;
          etaI = etaT;
          etaT = tmp;
        }
        float eta = etaI / etaT;
        float k = 1.0f - eta * eta * (1.0f - cosi * cosi);
        float f0 = (etaT - etaI) / (etaT + etaI);
        f0 = f0 * f0;
        float F = f0 + (1.0f - f0) * pow5(clampf(1.0f - cosi, 0.0f, 1.0f));
        float transmission = 1.0f - clampedOpacity;
        float reflectWeight = F;
        float refractWeight = (1.0f - F) * transmission;
        float totalWeight = fmaxf(reflectWeight + refractWeight, 1e-4f);
        float reflectProb = reflectWeight / totalWeight;
        float choice = halton(haltonIndex, 2 + step * 6 + 5);
        if (k < 0.0f || choice < reflectProb) {
          float3 reflectDir = normalize(I - 2.0f * dot(I, N) * N);
          rayOrigin = worldSpaceIntersectionPoint + reflectDir * 1e-3f;
          rayDirection = reflectDir;
          color *= totalWeight;
        } else {
          float cosT = sqrtf(fmaxf(k, 0.0f));
          float3 refractDir = normalize(eta * I + (eta * cosi - cosT) * N);
          rayOrigin = worldSpaceIntersectionPoint + refractDir * 1e-3f;
          rayDirection = refractDir;
          color *= totalWeight * albedo;
          consumeBounce = false;
        }
        skipLighting = true;
      }

      if (skipLighting) {
        bsdfPdf = 0.0f;
        step++;
        if (consumeBounce) {
          bounce++;
          transparencyPasses = 0;
        } else {
          transparencyPasses++;
          if (transparencyPasses > uniforms.maxBounces) {
            bounce++;
            transparencyPasses = 0;
          }
        }
        continue;
      }

      float perceptualRoughness = clampf(roughness, 0.04f, 1.0f);
      float alpha = perceptualRoughness * perceptualRoughness;
      float3 diffuseColor = albedo;
      float3 F0 = mix(make3(0.04f), albedo, metallic);
      float3 V = normalize(-rayDirection);

      accumulatedColor += color * emission;

      float lightSample = halton(haltonIndex, 2 + step * 6 + 0);
      const int pickCount = uniforms.lightCount + (envLight ? 1 : 0);
      int lightIndex = std::min(int(lightSample * float(pickCount)), pickCount - 1);
      const bool pickedEnvironment = envLight && lightIndex == uniforms.lightCount;
      const rt_light &light = a.lights[pickedEnvironment ? 0 : lightIndex];

      float3 worldSpaceLightDirection;
      float lightDistance;
      float3 lightColor;

      if (pickedEnvironment) { // include/rt_b200.h RT_ENV_IMPORTANCE
        r = make2(halton(haltonIndex, 2 + step * 6 + 1), halton(haltonIndex, 2 + step * 6 + 2));
        const float envPdf = sampleEnvironmentDirection(a.env, r, worldSpaceLightDirection);
        lightDistance = INFINITY;
        const float bouncePdf = saturate(dot(shadingNormal, worldSpaceLightDirection)) * kInvPi;
        lightColor = sampleEnvironment(a.env, worldSpaceLightDirection) / (envPdf / float(pickCount) + bouncePdf);
      } else if (light.type == RT_LIGHT_AREA) {
        r = make2(halton(haltonIndex, 2 + step * 6 + 1), halton(haltonIndex, 2 + step * 6 + 2));
        sampleAreaLight(light, r, worldSpaceIntersectionPoint, worldSpaceLightDirection, lightColor, lightDistance);
      } else if (light.type == RT_LIGHT_SPOT) {
        worldSpaceLightDirection = f3(light.position) - worldSpaceIntersectionPoint;
        lightDistance = length(worldSpaceLightDirection);
        float inverseLightDistance = 1.0f / fmaxf(lightDistance, 1e-3f);
        worldSpaceLightDirection *= inverseLightDistance;
        lightColor = make3(0.0f);
        float3 coneDirection = normalize(f3(light.direction));
        float spotResult = dot(-worldSpaceLightDirection, coneDirection);
        if (spotResult > cos_det(light.coneAngle)) lightColor = f3(light.color) * inverseLightDistance * inverseLightDistance;
      } else if (light.type == RT_LIGHT_POINT) {
        worldSpaceLightDirection = f3(light.position) - worldSpaceIntersectionPoint;
        lightDistance = length(worldSpaceLightDirection);
        float inverseLightDistance = 1.0f / fmaxf(lightDistance, 1e-3f);
        worldSpaceLightDirection *= inverseLightDistance;
        lightColor = f3(light.color) * inverseLightDistance * inverseLightDistance;
      } else {
        worldSpaceLightDirection = -normalize(f3(light.direction));
        lightDistance = INFINITY;
        lightColor = f3(light.color);
      }
      if (!pickedEnvironment) lightColor *= float(pickCount);

      if (uniforms.shadingMode == RT_SHADING_LEGACY) {
        float3 L = normalize(worldSpaceLightDirection);
        float NdotL = saturate(dot(shadingNormal, L));
        float3 legacyColor = color * albedo;
        if (length(legacyColor) < 0.001f) break;
        if (length(lightColor) > 0.0001f && NdotL > 0.0f) {
          float3 so = worldSpaceIntersectionPoint + worldSpaceSurfaceNormal * 1e-3f;
          ++stats.anyRays;
          if (!traceAny(*a.tlas, so, worldSpaceLightDirection, 0.0f, lightDistance - 1e-3f))
            accumulatedColor += legacyColor * lightColor * NdotL;
        }
        color = legacyColor * ao;
        if (length(color) < 0.001f) break;
        r = make2(halton(haltonIndex, 2 + step * 5 + 3), halton(haltonIndex, 2 + step * 5 + 4));
        float3 worldSpaceSampleDirection = sampleCosineWeightedHemisphere(r);
        worldSpaceSampleDirection = alignHemisphereWithNormal(worldSpaceSampleDirection, shadingNormal);
        bsdfPdf = envLight ? saturate(dot(shadingNormal, worldSpaceSampleDirection)) * kInvPi : 0.0f;
        rayOrigin = worldSpaceIntersectionPoint + worldSpaceSurfaceNormal * 1e-3f;
        rayDirection = worldSpaceSampleDirection;
        step++;
        bounce++;
        transparencyPasses = 0;
        continue;
      }

      if (length(lightColor) > 0.0001f) {
        float3 L = normalize(worldSpaceLightDirection);
        float3 H = normalize(V + L);
        float NdotL = saturate(dot(shadingNormal, L));
        float NdotV = saturate(dot(shadingNormal, V));
        float NdotH = saturate(dot(shadingNormal, H));
        float VdotH = saturate(dot(V, H));

        float3 F = fresnelSchlick(VdotH, F0);
        float D = distributionGGX(NdotH, alpha);
        float k = (perceptualRoughness + 1.0f);
        k = (k * k) / 8.0f;
        float G = geometrySmith(NdotV, NdotL, k);

        float3 specular = (D * G) * F / fmaxf(4.0f * NdotV * NdotL, 1e-4f);
        float3 kS = F;
        float3 kD = (1.0f - kS) * (1.0f - metallic);
        float3 diffuse = kD * diffuseColor / kPi;
        float3 direct = (diffuse + specular) * lightColor * NdotL;

        float3 so = worldSpaceIntersectionPoint + worldSpaceSurfaceNormal * 1e-3f;
        ++stats.anyRays;
        if (!traceAny(*a.tlas, so, worldSpaceLightDirection, 0.0f, lightDistance - 1e-3f)) accumulatedColor += color * direct;
      }

      color *= diffuseColor * (1.0f - metallic) * ao;
      if (length(color) < 0.001f) break;

      r = make2(halton(haltonIndex, 2 + step * 5 + 3), halton(haltonIndex, 2 + step * 5 + 4));
      float3 worldSpaceSampleDirection = sampleCosineWeightedHemisphere(r);
      worldSpaceSampleDirection = alignHemisphereWithNormal(worldSpaceSampleDirection, shadingNormal);
      bsdfPdf = envLight ? saturate(dot(shadingNormal, worldSpaceSampleDirection)) * kInvPi : 0.0f;
      rayOrigin = worldSpaceIntersectionPoint + worldSpaceSurfaceNormal * 1e-3f;
      rayDirection = worldSpaceSampleDirection;

      step++;
      bounce++;
      transparencyPasses = 0;
    }

    totalColor += accumulatedColor;

    if (sampleIndex == 0 && maxExtraSamples > 0) {
      float motionMag = fmaxf(length(motionVector), length(prevMotion));
      float low = fmaxf(uniforms.motionSamplingLowThresholdPixels, 0.0f);
      float high = fmaxf(uniforms.motionSamplingHighThresholdPixels, low + 1e-3f);
      float t = clampf((motionMag - low) / (high - low), 0.0f, 1.0f);
      int extraSamples = int(roundf(t * float(maxExtraSamples)));
      extraSamples = std::min(std::max(extraSamples, 0), maxExtraSamples);
      totalSamples = baseSamples + extraSamples;
    }
  }

  totalColor = totalColor / float(std::max(totalSamples, 1));

  if (uniforms.frameIndex > 0) {
    float4 pc = readImage(a.textures[RT_TEXTURE_ACCUMULATION], tidx, tidy);
    float3 prevColor{pc.x, pc.y, pc.z};
    float historyWeight = clampf(uniforms.accumulationWeight, 0.0f, 0.95f);
    if (uniforms.enableMotionAdaptiveAccumulation != 0) {
      float motionMag = fmaxf(length(motionVector), length(prevMotion));
      float low = fmaxf(uniforms.motionAccumulationLowThresholdPixels, 0.0f);
      float high = fmaxf(uniforms.motionAccumulationHighThresholdPixels, low + 1e-3f);
      float t = clampf((motionMag - low) / (high - low), 0.0f, 1.0f);
      float minWeight = clampf(uniforms.motionAccumulationMinWeight, 0.0f, 0.95f);
      minWeight = fminf(minWeight, historyWeight);
      historyWeight = mixf(historyWeight, minWeight, t);
    }
    totalColor = mix(totalColor, prevColor, historyWeight);
  }

  writeImage(a.textures[RT_TEXTURE_PREVIOUS_ACCUMULATION], tidx, tidy, {totalColor.x, totalColor.y, totalColor.z, 1.0f});
  writeImage(a.textures[RT_TEXTURE_DEPTH], tidx, tidy, {primaryDepth, 0, 0, 0});
  writeImage(a.textures[RT_TEXTURE_MOTION], tidx, tidy, {motionVector.x, motionVector.y, 0.0f, 0.0f});
  if (uniforms.enableDenoiseGBuffer != 0) {
    writeImage(a.textures[RT_TEXTURE_DIFFUSE_ALBEDO], tidx, tidy, outDiffuseAlbedo);
    writeImage(a.textures[RT_TEXTURE_SPECULAR_ALBEDO], tidx, tidy, outSpecularAlbedo);
    writeImage(a.textures[RT_TEXTURE_NORMAL], tidx, tidy, outNormal);
    writeImage(a.textures[RT_TEXTURE_ROUGHNESS], tidx, tidy, outRoughness);
  }
}

// ---- skinningKernel, one thread (Skinning.metal:7-49) -----------------------------------------------------
void skinningKernelVertex(uint32_t vertexID, const void *const *buffers, uint32_t vertexCount) {
  if (vertexID >= vertexCount) return;
  const rt_float3 *restPositions = static_cast<const rt_float3 *>(buffers[RT_BUFFER_REST_POSITIONS]);
  const rt_float3 *restNormals = static_cast<const rt_float3 *>(buffers[RT_BUFFER_REST_NORMALS]);
  const uint16_t *jointIndices = static_cast<const uint16_t *>(buffers[RT_BUFFER_JOINT_INDICES]);
  const float *jointWeights = static_cast<const float *>(buffers[RT_BUFFER_JOINT_WEIGHTS]);
  const float *jointMatrices = static_cast<const float *>(buffers[RT_BUFFER_JOINT_MATRICES]);
  rt_float3 *skinnedPositions = static_cast<rt_float3 *>(const_cast<void *>(buffers[RT_BUFFER_SKINNED_POSITIONS]));
  rt_float3 *skinnedNormals = static_cast<rt_float3 *>(const_cast<void *>(buffers[RT_BUFFER_SKINNED_NORMALS]));

  float3 position = ld3(restPositions, vertexID);
  float3 normal = ld3(restNormals, vertexID);
  uint16_t idx[4];
  float w[4];
  for (int k = 0; k < 4; ++k) {
    idx[k] = jointIndices[4 * vertexID + k];
    w[k] = jointWeights[4 * vertexID + k];
  }
  float weightSum = w[0] + w[1] + w[2] + w[3];
  if (weightSum < 0.0001f) {
    w[0] = 1.0f;
    w[1] = w[2] = w[3] = 0.0f;
  }
  float3 skinnedPos = make3(0.0f);
  float3 skinnedNrm = make3(0.0f);
  for (int k = 0; k < 4; ++k) {
    float4x4 m;
    std::memcpy(&m, jointMatrices + size_t(idx[k]) * 16, 64);
    skinnedPos += w[k] * mulPoint(m, position);
  }
  for (int k = 0; k < 4; ++k) {
    float4x4 m;
    std::memcpy(&m, jointMatrices + size_t(idx[k]) * 16, 64);
    skinnedNrm += w[k] * mulDir(m, normal);
  }
  skinnedPositions[vertexID] = {skinnedPos.x, skinnedPos.y, skinnedPos.z, 0.0f};
  skinnedNormals[vertexID] = {skinnedNrm.x, skinnedNrm.y, skinnedNrm.z, 0.0f};
}

} // namespace orc
