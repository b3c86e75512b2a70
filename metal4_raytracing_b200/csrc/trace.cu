// trace.cu — the per-pixel path-tracing kernel for sm_100a, megakernel layout.
//
// B200-native counterpart of the reference's raytracingKernel (MetalRaytracing/Raytracing.metal:220-831) behind
// the same argument table (buffers by BufferIndex, images by TextureIndex, Uniforms, function constant
// maxSubmeshes). One thread owns one pixel for all of its samples and path segments; a CTA is one 16x16 screen
// tile (the reference's threadgroup, Renderer.swift:1445-1451) whose eight warps each cover an 8x4 pixel block so
// primary rays of a warp stay coherent. Hardware intersection is replaced by traverse.cuh. Tiles are taken from
// the list this rank owns (tile % tileModulo == tileRemainder), so the same kernel serves 1..8 GPUs.
//
// Behaviour kept from the reference on purpose: a miss ends the path with no contribution (no environment
// lookup), EMA accumulation with history weight <= 0.95, the stride-5 / stride-6 Halton dimension mix,
// maxBounces counting closest-hit segments, a shadow ray whenever the light colour is non-negligible, glass
// refraction not consuming a bounce until transparencyPasses > maxBounces.
// Compile with -fmad=false (numeric contract).
#include <cstring>

#include "shade.cuh"
#include "traverse.cuh"

namespace rtb {

struct TraceParams {
  rt_uniforms uniforms;
  const TlasHeader *tlas;
  const rt_resource *resources;
  const rt_instance_descriptor *instances;
  const rt_instance_descriptor *prevInstances;
  const rt_light *lights;
  rt_image images[RT_TEXTURE_COUNT];
  const float *srgbLut;
  int maxSubmeshes;
  int tileModulo, tileRemainder;
  int tilesX, tilesY;
  uint32_t *primaryIds;
  unsigned long long *rayCounters;
  void *peerAccumulation[8];
  int peerCount;
};

struct SurfaceHit { // what shading needs from a closest hit
  f3 position;      // world
  f3 geometricNormal; // interpolated vertex normal in world space (or -direction fallback)
  f3 shadingNormal;
  f3 albedo, emission;
  float roughness, metallic, opacity, ior;
};

__device__ __forceinline__ f3 interpolate3(const rt_float3 *attr, const int32_t *indices, const RayHit &h) {
  const float wx = h.u, wy = h.v, wz = 1.0f - wx - wy;
  const uint32_t i1 = uint32_t(__ldg(indices + h.primitive * 3 + 1));
  const uint32_t i2 = uint32_t(__ldg(indices + h.primitive * 3 + 2));
  const uint32_t i0 = uint32_t(__ldg(indices + h.primitive * 3 + 0));
  const float4 a = __ldg(reinterpret_cast<const float4 *>(attr) + i1);
  const float4 b = __ldg(reinterpret_cast<const float4 *>(attr) + i2);
  const float4 c = __ldg(reinterpret_cast<const float4 *>(attr) + i0);
  return wx * mk3(a.x, a.y, a.z) + wy * mk3(b.x, b.y, b.z) + wz * mk3(c.x, c.y, c.z);
}

__device__ __forceinline__ f2 interpolate2(const float *attr, const int32_t *indices, const RayHit &h) {
  const float wx = h.u, wy = h.v, wz = 1.0f - wx - wy;
  const uint32_t i1 = uint32_t(__ldg(indices + h.primitive * 3 + 1));
  const uint32_t i2 = uint32_t(__ldg(indices + h.primitive * 3 + 2));
  const uint32_t i0 = uint32_t(__ldg(indices + h.primitive * 3 + 0));
  const float2 a = __ldg(reinterpret_cast<const float2 *>(attr) + i1);
  const float2 b = __ldg(reinterpret_cast<const float2 *>(attr) + i2);
  const float2 c = __ldg(reinterpret_cast<const float2 *>(attr) + i0);
  return wx * mk2(a.x, a.y) + wy * mk2(b.x, b.y) + wz * mk2(c.x, c.y);
}

// Per-triangle tangent frame from position / uv deltas (Raytracing.metal:185-218).
__device__ __forceinline__ bool tangentBasis(const rt_resource &res, const RayHit &h, f3 &tangent, f3 &bitangent) {
  const uint32_t i1 = uint32_t(__ldg(res.indices + h.primitive * 3 + 1));
  const uint32_t i2 = uint32_t(__ldg(res.indices + h.primitive * 3 + 2));
  const uint32_t i0 = uint32_t(__ldg(res.indices + h.primitive * 3 + 0));
  const float4 a = __ldg(reinterpret_cast<const float4 *>(res.positions) + i1);
  const float4 b = __ldg(reinterpret_cast<const float4 *>(res.positions) + i2);
  const float4 c = __ldg(reinterpret_cast<const float4 *>(res.positions) + i0);
  const float2 ta = __ldg(reinterpret_cast<const float2 *>(res.uvs) + i1);
  const float2 tb = __ldg(reinterpret_cast<const float2 *>(res.uvs) + i2);
  const float2 tc = __ldg(reinterpret_cast<const float2 *>(res.uvs) + i0);
  const f3 p0 = mk3(a.x, a.y, a.z), p1 = mk3(b.x, b.y, b.z), p2 = mk3(c.x, c.y, c.z);
  const f3 e1 = p1 - p0, e2 = p2 - p0;
  const f2 d1 = mk2(tb.x, tb.y) - mk2(ta.x, ta.y), d2 = mk2(tc.x, tc.y) - mk2(ta.x, ta.y);
  const float denom = d1.x * d2.y - d1.y * d2.x;
  if (fabsf(denom) < 1e-8f) return false;
  const float r = 1.0f / denom;
  tangent = (e1 * d2.y - e2 * d1.y) * r;
  bitangent = (e2 * d1.x - e1 * d2.x) * r;
  return (length(tangent) > 1e-8f) && (length(bitangent) > 1e-8f);
}

__global__ void __launch_bounds__(256) k_trace_megakernel(const __grid_constant__ TraceParams P) {
  const rt_uniforms &U = P.uniforms;
  // tile owned by this CTA; warps are 8x4 pixel blocks inside the 16x16 tile
  const int ownedIndex = blockIdx.x;
  const int tile = ownedIndex * P.tileModulo + P.tileRemainder;
  if (tile >= P.tilesX * P.tilesY) return;
  const int tileX = tile % P.tilesX, tileY = tile / P.tilesX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int px = tileX * 16 + (warp & 1) * 8 + (lane & 7);
  const int py = tileY * 16 + (warp >> 1) * 4 + (lane >> 3);
  if (!(px < U.width && py < U.height)) return;
  const size_t pixelIndex = size_t(py) * size_t(U.width) + size_t(px);

  const uint32_t offset = reinterpret_cast<const uint32_t *>(P.images[RT_TEXTURE_RANDOM].data)[pixelIndex];
  f3 totalColor = mk3(0.0f);
  const f4 pm = readImage(P.images[RT_TEXTURE_MOTION], px, py);
  const f2 prevMotion = mk2(pm.x, pm.y);

  float primaryDepth = 1.0e8f;
  f2 motionVector = mk2(0.0f, 0.0f);
  bool hadPrimaryHit = false;
  f4 gDiffuse = {0, 0, 0, 0}, gSpecular = {0, 0, 0, 0}, gNormal = {0, 0, 0, 0}, gRoughness = {0, 0, 0, 0};
  bool wroteGBuffer = false;

  unsigned long long nClosest = 0, nAny = 0, nHits = 0;

  const int baseSamples = max(U.samplesPerPixel, 1);
  const int maxExtraSamples = (U.enableMotionAdaptiveSampling != 0) ? max(U.motionSamplingMaxExtraSamples, 0) : 0;
  const int sampleStride = baseSamples + maxExtraSamples;
  int totalSamples = baseSamples;

  const f3 camPos = mk3(U.camera.position), camRight = mk3(U.camera.right), camUp = mk3(U.camera.up),
           camFwd = mk3(U.camera.forward);

  for (int sampleIndex = 0; sampleIndex < totalSamples; ++sampleIndex) {
    const int frameOffset = int(U.frameIndex * uint32_t(sampleStride) + uint32_t(sampleIndex));
    const int hIndex = int(offset + uint32_t(frameOffset));

    f2 r = mk2(halton(hIndex, 0), halton(hIndex, 1));
    const f2 samplePixel = mk2(float(px), float(py)) + r;
    f2 uv = samplePixel / mk2(float(U.width), float(U.height));
    uv = uv * 2.0f - mk2(1.0f, 1.0f);

    f3 rayOrigin = camPos;
    f3 rayDir = normalize(uv.x * camRight + uv.y * camUp + camFwd);

    f3 throughput = mk3(1.0f);
    f3 radiance = mk3(0.0f);
    int bounce = 0, step = 0, transparencyPasses = 0;

    while (bounce < U.maxBounces) {
      RayHit hit;
      ++nClosest;
      const bool found = traverseScene<false>(P.tlas, rayOrigin.x, rayOrigin.y, rayOrigin.z, rayDir.x, rayDir.y,
                                              rayDir.z, 0.0f, INFINITY, hit);
      if (P.primaryIds != nullptr && sampleIndex == 0 && step == 0) {
        uint4 id = found ? make_uint4(hit.instance, hit.geometry, hit.primitive, __float_as_uint(hit.t))
                         : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        reinterpret_cast<uint4 *>(P.primaryIds)[pixelIndex] = id;
      }
      if (!found) break;
      ++nHits;

      const int instanceIndex = int(hit.instance);
      const M34 objectToWorld = loadInstanceMatrix(P.instances + instanceIndex);
      const f3 hitPoint = rayOrigin + rayDir * hit.t;
      const rt_resource res = P.resources[instanceIndex * P.maxSubmeshes + int(hit.geometry)];

      if (bounce == 0 && sampleIndex == 0) { // depth + motion vector of the primary hit
        const f3 objPos = interpolate3(res.positions, res.indices, hit);
        const f3 prevObjPos = interpolate3(res.previousPositions, res.indices, hit);
        const f3 worldPos = mulPoint(objectToWorld, objPos);
        const M34 prevObjectToWorld = loadInstanceMatrix(P.prevInstances + instanceIndex);
        const f3 prevWorldPos = mulPoint(prevObjectToWorld, prevObjPos);
        const f3 viewPos = worldPos - camPos;
        f2 screenPos = mk2(dot(viewPos, camRight), dot(viewPos, camUp));
        const float depth = dot(viewPos, camFwd);
        primaryDepth = fmaxf(depth, 1.0e-3f);
        screenPos = screenPos / fmaxf(depth, 0.001f);
        const f3 pcPos = mk3(U.previousCamera.position), pcRight = mk3(U.previousCamera.right),
                 pcUp = mk3(U.previousCamera.up), pcFwd = mk3(U.previousCamera.forward);
        const f3 prevViewPos = prevWorldPos - pcPos;
        f2 prevScreenPos = mk2(dot(prevViewPos, pcRight), dot(prevViewPos, pcUp));
        const float prevDepth = dot(prevViewPos, pcFwd);
        prevScreenPos = prevScreenPos / fmaxf(prevDepth, 0.001f);
        const f2 motionNdc = screenPos - prevScreenPos;
        const float rightScale = fmaxf(length(camRight), 1e-5f);
        const float upScale = fmaxf(length(camUp), 1e-5f);
        f2 motionPixels = mk2(motionNdc.x * (float(U.width) / (2.0f * rightScale)),
                              motionNdc.y * (float(U.height) / (2.0f * upScale)));
        motionPixels.y = -motionPixels.y;
        motionVector = motionPixels;
        hadPrimaryHit = true;
      }

      const f3 objNormal = interpolate3(res.normals, res.indices, hit);
      f3 surfaceNormal = normalize(mulDir(objectToWorld, objNormal));
      if (length(objNormal) < 1e-10f) surfaceNormal = -rayDir;

      // material + textures (Raytracing.metal:399-456)
      const rt_material mat = *res.material;
      f3 albedo = mk3(mat.baseColor);
      const uint32_t flags = mat.textureFlags;
      const bool hasBase = (flags & RT_MATERIAL_TEXTURE_BASECOLOR) != 0, hasNormalMap = (flags & RT_MATERIAL_TEXTURE_NORMAL) != 0;
      const bool hasRough = (flags & RT_MATERIAL_TEXTURE_ROUGHNESS) != 0, hasMetal = (flags & RT_MATERIAL_TEXTURE_METALLIC) != 0;
      const bool hasOpacityMap = (flags & RT_MATERIAL_TEXTURE_OPACITY) != 0, hasEmissionMap = (flags & RT_MATERIAL_TEXTURE_EMISSION) != 0;
      f2 texCoord = mk2(0.0f, 0.0f);
      if (hasBase || hasNormalMap || hasRough || hasMetal || hasOpacityMap || hasEmissionMap) {
        texCoord = interpolate2(res.uvs, res.indices, hit);
        texCoord.y = 1.0f - texCoord.y;
      }
      f4 baseSample = {1.0f, 1.0f, 1.0f, 1.0f};
      if (hasBase) {
        baseSample = sampleTexture(res.baseColorMap, texCoord, P.srgbLut);
        albedo *= mk3(baseSample.x, baseSample.y, baseSample.z);
      }
      float roughness = 1.0f;
      if (hasRough) roughness = sampleTexture(res.roughnessMap, texCoord, P.srgbLut).x;
      float metallic = 0.0f;
      if (hasMetal) metallic = sampleTexture(res.metallicMap, texCoord, P.srgbLut).x;
      const float ao = 1.0f; // ENABLE_AO == 0 in the reference build
      float opacity = clampf(mat.opacity, 0.0f, 1.0f);
      if (hasOpacityMap) opacity *= sampleTexture(res.opacityMap, texCoord, P.srgbLut).x;
      f3 emission = mk3(mat.emission);
      if (hasEmissionMap) {
        const f4 e = sampleTexture(res.emissionMap, texCoord, P.srgbLut);
        emission = mk3(e.x, e.y, e.z);
      }

      if (U.debugTextureMode != RT_DEBUG_NONE) { // debug views (Raytracing.metal:458-490)
        f3 dbg = mk3(0.0f);
        switch (U.debugTextureMode) {
          case RT_DEBUG_BASECOLOR: dbg = hasBase ? mk3(baseSample.x, baseSample.y, baseSample.z) : mk3(1.0f, 0.0f, 1.0f); break;
          case RT_DEBUG_NORMAL:
            if (hasNormalMap) {
              const f4 n = sampleTexture(res.normalMap, texCoord, P.srgbLut);
              dbg = mk3(n.x, n.y, n.z);
            } else {
              dbg = surfaceNormal * 0.5f + 0.5f;
            }
            break;
          case RT_DEBUG_ROUGHNESS: dbg = mk3(roughness); break;
          case RT_DEBUG_METALLIC: dbg = mk3(metallic); break;
          case RT_DEBUG_AO: dbg = mk3(1.0f, 0.0f, 1.0f); break;
          case RT_DEBUG_EMISSION: dbg = emission; break;
          case RT_DEBUG_MOTION: {
            const f2 mp = hadPrimaryHit ? motionVector : prevMotion;
            const f2 scaled = mk2(clampf(mp.x * 0.05f, -1.0f, 1.0f), clampf(mp.y * 0.05f, -1.0f, 1.0f));
            const float mag = clampf(length(mp) * 0.1f, 0.0f, 1.0f);
            dbg = mk3(scaled.x * 0.5f + 0.5f, scaled.y * 0.5f + 0.5f, mag);
            break;
          }
          default: break;
        }
        radiance = dbg;
        break;
      }

      f3 shadingNormal = surfaceNormal;
      if (hasNormalMap) {
        f3 tangent, bitangent;
        if (tangentBasis(res, hit, tangent, bitangent)) {
          f3 worldT = mulDir(objectToWorld, tangent);
          worldT = normalize(worldT - surfaceNormal * dot(worldT, surfaceNormal));
          const f3 worldB = normalize(cross(surfaceNormal, worldT));
          const f4 ns = sampleTexture(res.normalMap, texCoord, P.srgbLut);
          const f3 nMap = mk3(ns.x, ns.y, ns.z) * 2.0f - mk3(1.0f);
          shadingNormal = normalize(nMap.x * worldT + nMap.y * worldB + nMap.z * surfaceNormal);
        }
      }

      if (U.enableDenoiseGBuffer != 0 && !wroteGBuffer && sampleIndex == 0) {
        const f3 diffuseAlbedo = albedo * (1.0f - metallic);
        const f3 specularAlbedo = mix(mk3(0.04f), albedo, metallic);
        const f3 nn = shadingNormal * 0.5f + 0.5f;
        gDiffuse = {diffuseAlbedo.x, diffuseAlbedo.y, diffuseAlbedo.z, 1.0f};
        gSpecular = {specularAlbedo.x, specularAlbedo.y, specularAlbedo.z, 1.0f};
        gNormal = {nn.x, nn.y, nn.z, 1.0f};
        gRoughness = {clampf(roughness, 0.0f, 1.0f), 0.0f, 0.0f, 1.0f};
        wroteGBuffer = true;
      }

      // glass: Fresnel-weighted choice between mirror reflection and refraction (Raytracing.metal:517-576)
      const float clampedOpacity = clampf(opacity, 0.0f, 1.0f);
      const float ior = fmaxf(mat.refractionIndex, 1.0f);
      if (clampedOpacity < 0.999f || ior > 1.01f) {
        f3 N = shadingNormal;
        const f3 I = rayDir;
        float cosi = clampf(dot(-I, N), -1.0f, 1.0f);
        float etaI = 1.0f, etaT = ior;
        if (cosi < 0.0f) {
          cosi = -cosi;
          N = -N;
          const float tmp = etaI;
          etaI = etaT;
          etaT = tmp;
        }
        const float eta = etaI / etaT;
        const float k = 1.0f - eta * eta * (1.0f - cosi * cosi);
        float f0 = (etaT - etaI) / (etaT + etaI);
        f0 = f0 * f0;
        const float F = f0 + (1.0f - f0) * pow5(clampf(1.0f - cosi, 0.0f, 1.0f));
        const float transmission = 1.0f - clampedOpacity;
        const float reflectWeight = F;
        const float refractWeight = (1.0f - F) * transmission;
        const float totalWeight = fmaxf(reflectWeight + refractWeight, 1e-4f);
        const float reflectProb = reflectWeight / totalWeight;
        const float choice = halton(hIndex, 2 + step * 6 + 5);
        bool consumeBounce = true;
        if (k < 0.0f || choice < reflectProb) {
          const f3 reflectDir = normalize(I - 2.0f * dot(I, N) * N);
          rayOrigin = hitPoint + reflectDir * 1e-3f;
          rayDir = reflectDir;
          throughput *= totalWeight;
        } else {
          const float cosT = sqrtf(fmaxf(k, 0.0f));
          const f3 refractDir = normalize(eta * I + (eta * cosi - cosT) * N);
          rayOrigin = hitPoint + refractDir * 1e-3f;
          rayDir = refractDir;
          throughput *= totalWeight * albedo;
          consumeBounce = false;
        }
        ++step;
        if (consumeBounce) {
          ++bounce;
          transparencyPasses = 0;
        } else {
          ++transparencyPasses;
          if (transparencyPasses > U.maxBounces) {
            ++bounce;
            transparencyPasses = 0;
          }
        }
        continue;
      }

      const float perceptualRoughness = clampf(roughness, 0.04f, 1.0f);
      const float alpha = perceptualRoughness * perceptualRoughness;
      const f3 F0 = mix(mk3(0.04f), albedo, metallic);
      const f3 V = normalize(-rayDir);

      radiance += throughput * emission;

      // one light, picked uniformly (Raytracing.metal:587-647)
      const float lightSample = halton(hIndex, 2 + step * 6 + 0);
      const int lightIndex = min(int(lightSample * float(U.lightCount)), U.lightCount - 1);
      const rt_light *light = P.lights + lightIndex;
      const int lightType = light->type;
      f3 L, lightColor;
      float lightDistance;
      if (lightType == RT_LIGHT_AREA) {
        r = mk2(halton(hIndex, 2 + step * 6 + 1), halton(hIndex, 2 + step * 6 + 2));
        const f2 s = r * 2.0f - mk2(1.0f, 1.0f);
        const f3 samplePosition = mk3(light->position) + mk3(light->right) * s.x + mk3(light->up) * s.y;
        L = samplePosition - hitPoint;
        lightDistance = length(L);
        const float inv = 1.0f / fmaxf(lightDistance, 1e-3f);
        L *= inv;
        lightColor = mk3(light->color);
        lightColor *= (inv * inv);
        lightColor *= saturatef(dot(-L, mk3(light->forward)));
      } else if (lightType == RT_LIGHT_SPOT) {
        L = mk3(light->position) - hitPoint;
        lightDistance = length(L);
        const float inv = 1.0f / fmaxf(lightDistance, 1e-3f);
        L *= inv;
        lightColor = mk3(0.0f);
        const f3 coneDirection = normalize(mk3(light->direction));
        const float spotResult = dot(-L, coneDirection);
        if (spotResult > cosDet(light->coneAngle)) lightColor = mk3(light->color) * inv * inv;
      } else if (lightType == RT_LIGHT_POINT) {
        L = mk3(light->position) - hitPoint;
        lightDistance = length(L);
        const float inv = 1.0f / fmaxf(lightDistance, 1e-3f);
        L *= inv;
        lightColor = mk3(light->color) * inv * inv;
      } else { // sun
        L = -normalize(mk3(light->direction));
        lightDistance = INFINITY;
        lightColor = mk3(light->color);
      }
      lightColor *= float(U.lightCount);

      const f3 shadowOrigin = hitPoint + surfaceNormal * 1e-3f;

      if (U.shadingMode == RT_SHADING_LEGACY) { // Lambert branch (Raytracing.metal:649-690)
        const f3 Ln = normalize(L);
        const float NdotL = saturatef(dot(shadingNormal, Ln));
        const f3 legacyColor = throughput * albedo;
        if (length(legacyColor) < 0.001f) break;
        if (length(lightColor) > 0.0001f && NdotL > 0.0f) {
          ++nAny;
          RayHit sh;
          if (!traverseScene<true>(P.tlas, shadowOrigin.x, shadowOrigin.y, shadowOrigin.z, L.x, L.y, L.z, 0.0f,
                                   lightDistance - 1e-3f, sh))
            radiance += legacyColor * lightColor * NdotL;
        }
        throughput = legacyColor * ao;
        if (length(throughput) < 0.001f) break;
      } else { // Cook-Torrance (Raytracing.metal:692-753)
        if (length(lightColor) > 0.0001f) {
          const f3 Ln = normalize(L);
          const f3 H = normalize(V + Ln);
          const float NdotL = saturatef(dot(shadingNormal, Ln));
          const float NdotV = saturatef(dot(shadingNormal, V));
          const float NdotH = saturatef(dot(shadingNormal, H));
          const float VdotH = saturatef(dot(V, H));
          const f3 F = fresnelSchlick(VdotH, F0);
          const float D = distributionGGX(NdotH, alpha);
          float k = (perceptualRoughness + 1.0f);
          k = (k * k) / 8.0f;
          const float G = geometrySmith(NdotV, NdotL, k);
          const f3 specular = (D * G) * F / fmaxf(4.0f * NdotV * NdotL, 1e-4f);
          const f3 kD = (1.0f - F) * (1.0f - metallic);
          const f3 diffuse = kD * albedo / kPi;
          const f3 direct = (diffuse + specular) * lightColor * NdotL;
          ++nAny;
          RayHit sh;
          if (!traverseScene<true>(P.tlas, shadowOrigin.x, shadowOrigin.y, shadowOrigin.z, L.x, L.y, L.z, 0.0f,
                                   lightDistance - 1e-3f, sh))
            radiance += throughput * direct;
        }
        throughput *= albedo * (1.0f - metallic) * ao;
        if (length(throughput) < 0.001f) break;
      }

      // cosine-weighted bounce (Raytracing.metal:763-774); note the stride-5 dimension index
      r = mk2(halton(hIndex, 2 + step * 5 + 3), halton(hIndex, 2 + step * 5 + 4));
      const f3 local = sampleCosineWeightedHemisphere(r);
      rayDir = alignHemisphereWithNormal(local, shadingNormal);
      rayOrigin = shadowOrigin;
      ++step;
      ++bounce;
      transparencyPasses = 0;
    }

    totalColor += radiance;

    if (sampleIndex == 0 && maxExtraSamples > 0) { // motion-adaptive sample count (Raytracing.metal:779-789)
      const float motionMag = fmaxf(length(motionVector), length(prevMotion));
      const float low = fmaxf(U.motionSamplingLowThresholdPixels, 0.0f);
      const float high = fmaxf(U.motionSamplingHighThresholdPixels, low + 1e-3f);
      const float t = clampf((motionMag - low) / (high - low), 0.0f, 1.0f);
      int extra = int(roundf(t * float(maxExtraSamples)));
      extra = min(max(extra, 0), maxExtraSamples);
      totalSamples = baseSamples + extra;
    }
  }

  totalColor = totalColor / float(max(totalSamples, 1));

  if (U.frameIndex > 0) { // exponential moving average with the history image (Raytracing.metal:796-817)
    const f4 pc = readImage(P.images[RT_TEXTURE_ACCUMULATION], px, py);
    float historyWeight = clampf(U.accumulationWeight, 0.0f, 0.95f);
    if (U.enableMotionAdaptiveAccumulation != 0) {
      const float motionMag = fmaxf(length(motionVector), length(prevMotion));
      const float low = fmaxf(U.motionAccumulationLowThresholdPixels, 0.0f);
      const float high = fmaxf(U.motionAccumulationHighThresholdPixels, low + 1e-3f);
      const float t = clampf((motionMag - low) / (high - low), 0.0f, 1.0f);
      float minWeight = clampf(U.motionAccumulationMinWeight, 0.0f, 0.95f);
      minWeight = fminf(minWeight, historyWeight);
      historyWeight = mixf(historyWeight, minWeight, t);
    }
    totalColor = mix(totalColor, mk3(pc.x, pc.y, pc.z), historyWeight);
  }

  const f4 outColor = {totalColor.x, totalColor.y, totalColor.z, 1.0f};
  writeImage(P.images[RT_TEXTURE_PREVIOUS_ACCUMULATION], px, py, outColor);
  for (int p = 0; p < P.peerCount; ++p) // multi-GPU: publish owned pixels into every rank's frame over NVLink
    if (P.peerAccumulation[p] != nullptr && P.peerAccumulation[p] != P.images[RT_TEXTURE_PREVIOUS_ACCUMULATION].data)
      writeImageAt(P.peerAccumulation[p], P.images[RT_TEXTURE_PREVIOUS_ACCUMULATION].format, pixelIndex, outColor);
  writeImage(P.images[RT_TEXTURE_DEPTH], px, py, {primaryDepth, 0.0f, 0.0f, 0.0f});
  writeImage(P.images[RT_TEXTURE_MOTION], px, py, {motionVector.x, motionVector.y, 0.0f, 0.0f});
  if (U.enableDenoiseGBuffer != 0) {
    writeImage(P.images[RT_TEXTURE_DIFFUSE_ALBEDO], px, py, gDiffuse);
    writeImage(P.images[RT_TEXTURE_SPECULAR_ALBEDO], px, py, gSpecular);
    writeImage(P.images[RT_TEXTURE_NORMAL], px, py, gNormal);
    writeImage(P.images[RT_TEXTURE_ROUGHNESS], px, py, gRoughness);
  }

  if (P.rayCounters != nullptr) { // probe: warp-aggregated counters
    const unsigned mask = __activemask();
    for (int o = 16; o > 0; o >>= 1) {
      nClosest += __shfl_down_sync(mask, nClosest, o);
      nAny += __shfl_down_sync(mask, nAny, o);
      nHits += __shfl_down_sync(mask, nHits, o);
    }
    if (lane == __ffs(int(mask)) - 1) {
      atomicAdd(P.rayCounters + 0, nClosest);
      atomicAdd(P.rayCounters + 1, nAny);
      atomicAdd(P.rayCounters + 2, nHits);
    }
  }
}

int launchTrace(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], const rt_image textures[RT_TEXTURE_COUNT],
                int maxSubmeshes, const rt_trace_options *opt) {
  RT_CHECK(buffers != nullptr && textures != nullptr, "rt_trace: null argument table");
  RT_CHECK(buffers[RT_BUFFER_UNIFORMS] != nullptr, "rt_trace: buffer 0 (Uniforms) is not bound");
  RT_CHECK(buffers[RT_BUFFER_ACCELERATION_STRUCTURE] != nullptr, "rt_trace: buffer 8 (acceleration structure) is not bound");
  RT_CHECK(buffers[RT_BUFFER_RESOURCES] != nullptr, "rt_trace: buffer 5 (resources) is not bound");
  RT_CHECK(buffers[RT_BUFFER_INSTANCE_DESCRIPTORS] != nullptr, "rt_trace: buffer 9 (instance descriptors) is not bound");
  RT_CHECK(buffers[RT_BUFFER_PREVIOUS_INSTANCE_DESCRIPTORS] != nullptr, "rt_trace: buffer 17 (previous instance descriptors) is not bound");
  RT_CHECK(buffers[RT_BUFFER_LIGHTS] != nullptr, "rt_trace: buffer 6 (lights) is not bound");
  RT_CHECK(maxSubmeshes >= 1, "rt_trace: function constant maxSubmeshes must be >= 1");
  TraceParams P{};
  std::memcpy(&P.uniforms, buffers[RT_BUFFER_UNIFORMS], sizeof(rt_uniforms));
  RT_CHECK(P.uniforms.width > 0 && P.uniforms.height > 0, "rt_trace: empty render target");
  RT_CHECK(P.uniforms.lightCount >= 1, "rt_trace: lightCount must be >= 1");
  uint64_t tlasId = uint64_t(reinterpret_cast<uintptr_t>(buffers[RT_BUFFER_ACCELERATION_STRUCTURE]));
  auto it = ctx->accels.find(tlasId);
  RT_CHECK(it != ctx->accels.end() && it->second->isTlas, "rt_trace: buffer 8 is not a TLAS id from rt_tlas_build");
  P.tlas = static_cast<const TlasHeader *>(it->second->headerDev);
  P.resources = static_cast<const rt_resource *>(buffers[RT_BUFFER_RESOURCES]);
  P.instances = static_cast<const rt_instance_descriptor *>(buffers[RT_BUFFER_INSTANCE_DESCRIPTORS]);
  P.prevInstances = static_cast<const rt_instance_descriptor *>(buffers[RT_BUFFER_PREVIOUS_INSTANCE_DESCRIPTORS]);
  P.lights = static_cast<const rt_light *>(buffers[RT_BUFFER_LIGHTS]);
  for (int i = 0; i < RT_TEXTURE_COUNT; ++i) P.images[i] = textures[i];
  const rt_image &rnd = textures[RT_TEXTURE_RANDOM], &dst = textures[RT_TEXTURE_PREVIOUS_ACCUMULATION];
  RT_CHECK(rnd.data && rnd.format == RT_FORMAT_R32_UINT && rnd.width == P.uniforms.width && rnd.height == P.uniforms.height,
           "rt_trace: texture 2 (random) must be an r32uint image of the render size");
  RT_CHECK(dst.data && dst.width == P.uniforms.width && dst.height == P.uniforms.height,
           "rt_trace: texture 1 (destination accumulation) is not bound at the render size");
  RT_CHECK(textures[RT_TEXTURE_MOTION].data && textures[RT_TEXTURE_DEPTH].data, "rt_trace: depth/motion images are not bound");
  RT_CHECK(P.uniforms.frameIndex == 0 || textures[RT_TEXTURE_ACCUMULATION].data, "rt_trace: texture 0 (history) is not bound");
  P.srgbLut = ctx->srgbLutDev;
  P.maxSubmeshes = maxSubmeshes;
  P.tileModulo = (opt && opt->tileModulo > 1) ? opt->tileModulo : 1;
  P.tileRemainder = (opt && opt->tileModulo > 1) ? opt->tileRemainder : 0;
  RT_CHECK(P.tileRemainder >= 0 && P.tileRemainder < P.tileModulo, "rt_trace: tileRemainder out of range");
  P.tilesX = (P.uniforms.width + 15) / 16;
  P.tilesY = (P.uniforms.height + 15) / 16;
  P.primaryIds = opt ? opt->primaryIdsDev : nullptr;
  P.rayCounters = opt ? reinterpret_cast<unsigned long long *>(opt->rayCountersDev) : nullptr;
  P.peerCount = 0;
  if (opt && opt->peerAccumulation) {
    RT_CHECK(P.tileModulo <= 8, "rt_trace: at most 8 peers");
    P.peerCount = P.tileModulo;
    for (int p = 0; p < P.peerCount; ++p) P.peerAccumulation[p] = opt->peerAccumulation[p];
  }
  const int tileCount = P.tilesX * P.tilesY;
  const int owned = (tileCount - P.tileRemainder + P.tileModulo - 1) / P.tileModulo;
  if (owned <= 0) return 0;
  k_trace_megakernel<<<owned, 256, 0, ctx->stream>>>(P);
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

} // namespace rtb
