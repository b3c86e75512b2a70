// path_step.cuh — the per-segment work of the path tracer, shared by both kernel layouts (trace.cu megakernel,
// trace_wavefront.cu): camera-ray generation, the shading of one closest hit (everything the reference's bounce
// loop does between two intersect calls, MetalRaytracing/Raytracing.metal:324-774) and the per-pixel resolve
// (sample mean + EMA + image writes, :792-829). One implementation => both layouts produce the same bits.
//
// Behaviour kept from the reference on purpose: a miss ends the path with no contribution (no environment
// lookup), EMA accumulation with history weight <= 0.95, the stride-5 / stride-6 Halton dimension mix,
// maxBounces counting closest-hit segments, a shadow ray whenever the light colour is non-negligible, glass
// refraction not consuming a bounce until transparencyPasses > maxBounces.
// The shadow ray of a segment is returned to the caller instead of being traced in place: its only effect is
// `radiance += contribution` when unoccluded, and nothing reads radiance in between, so tracing it right after
// shadeSegment() returns is arithmetically identical to the reference order.
#pragma once
#include "shade.cuh"
#include "traverse.cuh"

namespace rtb {

struct TraceParams {
  rt_uniforms uniforms;
  TlasHeader tlas; // by value: the traversal reads its pointers from the constant bank
  const float4 *tlasRootBox; // exact box of the TLAS root (lo, hi): grid for the ray-reordering keys
  const rt_resource *resources;
  const rt_instance_descriptor *instances;
  const rt_instance_descriptor *prevInstances;
  const rt_light *lights;
  const float4 *lightDerived; // per light: normalize(direction).xyz, cos(coneAngle) — hoisted out of the per-hit code
  const float4 *nodeUnionBox; // flat TLAS: bounding sphere (centre, radius) of the instances whose BLAS has nodes (k_prepare_classes)
  rt_image images[RT_TEXTURE_COUNT];
  const float *srgbLut;
  int maxSubmeshes;
  int tileModulo, tileRemainder;
  int sampleModulo, sampleRemainder; // sample partition (rt_trace_options): this dispatch owns samples s % modulo == remainder
  int tilesX, tilesY;
  uint32_t *primaryIds;
  unsigned long long *rayCounters;
  void *peerAccumulation[8];
  int peerCount;
  rt_environment env; // texelsDev == nullptr: off (reference behaviour)
  uint32_t hints;     // RT_TRACE_HINT_*
};

// Pixel owned by slot `ownedIndex * 256 + t`: CTA-sized 16x16 tiles, eight 8x4-pixel warps per tile.
__device__ __forceinline__ bool ownedPixel(const TraceParams &P, int ownedTile, int t, int &px, int &py) {
  const int tile = ownedTile * P.tileModulo + P.tileRemainder;
  if (tile >= P.tilesX * P.tilesY) return false;
  const int tileX = tile % P.tilesX, tileY = tile / P.tilesX;
  const int warp = t >> 5, lane = t & 31;
  px = tileX * 16 + (warp & 1) * 8 + (lane & 7);
  py = tileY * 16 + (warp >> 1) * 4 + (lane >> 3);
  return px < P.uniforms.width && py < P.uniforms.height;
}

struct PathState { // registers of one path (one sample of one pixel)
  f3 origin, dir;
  f3 throughput; // `color` in the reference
  f3 radiance;   // `accumulatedColor`
  int bounce, step, transparencyPasses;
  float bsdfPdf; // RT_ENV_IMPORTANCE: density the current direction was drawn with by the cosine bounce; 0 = camera ray
                 // or glass (not sampled by the environment light, so a miss takes the environment at full weight)
};

__host__ __device__ __forceinline__ bool environmentIsLight(const TraceParams &P) {
  return (P.env.flags & RT_ENV_IMPORTANCE) != 0u && P.env.cdfDev != nullptr && P.env.texelsDev != nullptr;
}

struct PrimaryOutputs { // per pixel, filled by the first hit of sample 0
  float depth;
  f2 motion;
  bool hadPrimaryHit;
  bool wroteGBuffer;
  f4 gDiffuse, gSpecular, gNormal, gRoughness;
};

struct ShadowRequest {
  bool valid;
  f3 origin, dir;
  float tmax;
  f3 contribution;
};

__device__ __forceinline__ int haltonIndex(const rt_uniforms &U, uint32_t offset, int sampleStride, int sampleIndex) {
  const int frameOffset = int(U.frameIndex * uint32_t(sampleStride) + uint32_t(sampleIndex));
  return int(offset + uint32_t(frameOffset));
}

// Raytracing.metal:272-297
__device__ __forceinline__ void startPath(const rt_uniforms &U, int px, int py, int hIndex, PathState &s) {
  const f2 r = mk2(halton(hIndex, 0), halton(hIndex, 1));
  const f2 samplePixel = mk2(float(px), float(py)) + r;
  f2 uv = samplePixel / mk2(float(U.width), float(U.height));
  uv = uv * 2.0f - mk2(1.0f, 1.0f);
  s.origin = mk3(U.camera.position);
  s.dir = normalize(uv.x * mk3(U.camera.right) + uv.y * mk3(U.camera.up) + mk3(U.camera.forward));
  s.throughput = mk3(1.0f);
  s.radiance = mk3(0.0f);
  s.bounce = s.step = s.transparencyPasses = 0;
  s.bsdfPdf = 0.0f;
}

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

// Shades one closest hit and advances the path. Returns true when the path continues with (s.origin, s.dir).
// `shadow` is the shadow ray to trace for this segment (valid == false: none).
// kTextures = false compiles the material-map paths out (RT_TRACE_HINT_UNTEXTURED); kPlain = true compiles the debug
// views and the Legacy branch out (the launcher checks debugTextureMode == 0 and shadingMode == PBR). The generic
// instantiation <true, false> is what the megakernel uses.
template <bool kTextures = true, bool kPlain = false>
__device__ __forceinline__ bool shadeSegment(const TraceParams &P, PathState &s, const RayHit &hit, int hIndex,
                                             int sampleIndex, const f2 &prevMotion, PrimaryOutputs &prim,
                                             ShadowRequest &shadow) {
  const rt_uniforms &U = P.uniforms;
  shadow.valid = false;
  const int instanceIndex = int(hit.instance);
  const M34 objectToWorld = loadInstanceMatrix(P.instances + instanceIndex);
  const f3 hitPoint = s.origin + s.dir * hit.t;
  // a reference, not a copy: only the pointers a material actually uses are loaded (the 104-byte row has seven
  // texture slots that an untextured material never touches, and holding all 13 pointers costs 26 registers)
  const rt_resource &res = P.resources[instanceIndex * P.maxSubmeshes + int(hit.geometry)];

  if (s.bounce == 0 && sampleIndex == 0) { // depth + motion vector of the primary hit (Raytracing.metal:341-389)
    const f3 camPos = mk3(U.camera.position), camRight = mk3(U.camera.right), camUp = mk3(U.camera.up),
             camFwd = mk3(U.camera.forward);
    const f3 objPos = interpolate3(res.positions, res.indices, hit);
    const f3 prevObjPos = interpolate3(res.previousPositions, res.indices, hit);
    const f3 worldPos = mulPoint(objectToWorld, objPos);
    const M34 prevObjectToWorld = loadInstanceMatrix(P.prevInstances + instanceIndex);
    const f3 prevWorldPos = mulPoint(prevObjectToWorld, prevObjPos);
    const f3 viewPos = worldPos - camPos;
    f2 screenPos = mk2(dot(viewPos, camRight), dot(viewPos, camUp));
    const float depth = dot(viewPos, camFwd);
    prim.depth = fmaxf(depth, 1.0e-3f);
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
    prim.motion = motionPixels;
    prim.hadPrimaryHit = true;
  }

  const f3 objNormal = interpolate3(res.normals, res.indices, hit);
  f3 surfaceNormal = normalize(mulDir(objectToWorld, objNormal));
  if (length(objNormal) < 1e-10f) surfaceNormal = -s.dir;

  // material + textures (Raytracing.metal:399-456)
  const rt_material mat = *res.material;
  f3 albedo = mk3(mat.baseColor);
  const uint32_t flags = kTextures ? mat.textureFlags : 0u;
  const bool hasBase = (flags & RT_MATERIAL_TEXTURE_BASECOLOR) != 0, hasNormalMap = (flags & RT_MATERIAL_TEXTURE_NORMAL) != 0;
  const bool hasRough = (flags & RT_MATERIAL_TEXTURE_ROUGHNESS) != 0, hasMetal = (flags & RT_MATERIAL_TEXTURE_METALLIC) != 0;
  const bool hasOpacityMap = (flags & RT_MATERIAL_TEXTURE_OPACITY) != 0, hasEmissionMap = (flags & RT_MATERIAL_TEXTURE_EMISSION) != 0;
  // the reference's compile-time ENABLE_AO (ShaderTypes.h:155-157, default 0; Raytracing.metal:405-409) is a per-dispatch
  // switch here (RT_TRACE_ENABLE_AO in rt_trace_options.hints): uniform over the launch, textured build only
  const bool enableAO = kTextures && (P.hints & RT_TRACE_ENABLE_AO) != 0u;
  const bool hasAO = enableAO && (flags & RT_MATERIAL_TEXTURE_AO) != 0;
  f2 texCoord = mk2(0.0f, 0.0f);
  if (hasBase || hasNormalMap || hasRough || hasMetal || hasAO || hasOpacityMap || hasEmissionMap) {
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
  float ao = 1.0f;
  if (hasAO) ao = sampleTexture(res.aoMap, texCoord, P.srgbLut).x; // #if ENABLE_AO (Raytracing.metal:442-446)
  float opacity = clampf(mat.opacity, 0.0f, 1.0f);
  if (hasOpacityMap) opacity *= sampleTexture(res.opacityMap, texCoord, P.srgbLut).x;
  f3 emission = mk3(mat.emission);
  if (hasEmissionMap) {
    const f4 e = sampleTexture(res.emissionMap, texCoord, P.srgbLut);
    emission = mk3(e.x, e.y, e.z);
  }

  if (!kPlain && U.debugTextureMode != RT_DEBUG_NONE) { // debug views end the path (Raytracing.metal:458-490)
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
      case RT_DEBUG_AO: dbg = enableAO ? mk3(ao) : mk3(1.0f, 0.0f, 1.0f); break; // Raytracing.metal:475-479
      case RT_DEBUG_EMISSION: dbg = emission; break;
      case RT_DEBUG_MOTION: {
        const f2 mp = prim.hadPrimaryHit ? prim.motion : prevMotion;
        const f2 scaled = mk2(clampf(mp.x * 0.05f, -1.0f, 1.0f), clampf(mp.y * 0.05f, -1.0f, 1.0f));
        const float mag = clampf(length(mp) * 0.1f, 0.0f, 1.0f);
        dbg = mk3(scaled.x * 0.5f + 0.5f, scaled.y * 0.5f + 0.5f, mag);
        break;
      }
      default: break;
    }
    s.radiance = dbg;
    return false;
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

  if (U.enableDenoiseGBuffer != 0 && !prim.wroteGBuffer && sampleIndex == 0) {
    const f3 diffuseAlbedo = albedo * (1.0f - metallic);
    const f3 specularAlbedo = mix(mk3(0.04f), albedo, metallic);
    const f3 nn = shadingNormal * 0.5f + 0.5f;
    prim.gDiffuse = {diffuseAlbedo.x, diffuseAlbedo.y, diffuseAlbedo.z, 1.0f};
    prim.gSpecular = {specularAlbedo.x, specularAlbedo.y, specularAlbedo.z, 1.0f};
    prim.gNormal = {nn.x, nn.y, nn.z, 1.0f};
    prim.gRoughness = {clampf(roughness, 0.0f, 1.0f), 0.0f, 0.0f, 1.0f};
    prim.wroteGBuffer = true;
  }

  // glass: Fresnel-weighted choice between mirror reflection and refraction (Raytracing.metal:517-576)
  const float clampedOpacity = clampf(opacity, 0.0f, 1.0f);
  const float ior = fmaxf(mat.refractionIndex, 1.0f);
  if (clampedOpacity < 0.999f || ior > 1.01f) {
    f3 N = shadingNormal;
    const f3 I = s.dir;
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
    const float choice = halton(hIndex, 2 + s.step * 6 + 5);
    bool consumeBounce = true;
    if (k < 0.0f || choice < reflectProb) {
      const f3 reflectDir = normalize(I - 2.0f * dot(I, N) * N);
      s.origin = hitPoint + reflectDir * 1e-3f;
      s.dir = reflectDir;
      s.throughput *= totalWeight;
    } else {
      const float cosT = sqrtf(fmaxf(k, 0.0f));
      const f3 refractDir = normalize(eta * I + (eta * cosi - cosT) * N);
      s.origin = hitPoint + refractDir * 1e-3f;
      s.dir = refractDir;
      s.throughput *= totalWeight * albedo;
      consumeBounce = false;
    }
    s.bsdfPdf = 0.0f;
    ++s.step;
    if (consumeBounce) {
      ++s.bounce;
      s.transparencyPasses = 0;
    } else {
      ++s.transparencyPasses;
      if (s.transparencyPasses > U.maxBounces) {
        ++s.bounce;
        s.transparencyPasses = 0;
      }
    }
    return s.bounce < U.maxBounces;
  }

  const float perceptualRoughness = clampf(roughness, 0.04f, 1.0f);
  const float alpha = perceptualRoughness * perceptualRoughness;
  const f3 F0 = mix(mk3(0.04f), albedo, metallic);
  const f3 V = normalize(-s.dir);

  s.radiance += s.throughput * emission;

  // one light, picked uniformly (Raytracing.metal:587-647)
  const float lightSample = halton(hIndex, 2 + s.step * 6 + 0);
  // extension: the environment is light number lightCount. Only the general build of the shade kernel carries the
  // code (the launcher routes dispatches with RT_ENV_IMPORTANCE to it), the plain-PBR build stays as small as it was
  const bool envLight = !kPlain && environmentIsLight(P);
  const int pickCount = U.lightCount + (envLight ? 1 : 0);
  const int lightIndex = min(int(lightSample * float(pickCount)), pickCount - 1);
  const bool pickedEnvironment = envLight && lightIndex == U.lightCount;
  const rt_light *light = P.lights + (pickedEnvironment ? 0 : lightIndex);
  const int lightType = pickedEnvironment ? -1 : light->type;
  f3 L, lightColor;
  float lightDistance;
  if (pickedEnvironment) { // rt_b200.h RT_ENV_IMPORTANCE
    const f2 r = mk2(halton(hIndex, 2 + s.step * 6 + 1), halton(hIndex, 2 + s.step * 6 + 2));
    const float envPdf = sampleEnvironmentDirection(P.env, r, L);
    lightDistance = INFINITY;
    const float bouncePdf = saturatef(dot(shadingNormal, L)) * kInvPi;
    lightColor = sampleEnvironment(P.env, L) / (envPdf / float(pickCount) + bouncePdf);
  } else if (lightType == RT_LIGHT_AREA) {
    const f2 r = mk2(halton(hIndex, 2 + s.step * 6 + 1), halton(hIndex, 2 + s.step * 6 + 2));
    const f2 sq = r * 2.0f - mk2(1.0f, 1.0f);
    const f3 samplePosition = mk3(light->position) + mk3(light->right) * sq.x + mk3(light->up) * sq.y;
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
    const float4 ld = __ldg(P.lightDerived + lightIndex);
    const f3 coneDirection = mk3(ld.x, ld.y, ld.z);
    const float spotResult = dot(-L, coneDirection);
    if (spotResult > ld.w) lightColor = mk3(light->color) * inv * inv;
  } else if (lightType == RT_LIGHT_POINT) {
    L = mk3(light->position) - hitPoint;
    lightDistance = length(L);
    const float inv = 1.0f / fmaxf(lightDistance, 1e-3f);
    L *= inv;
    lightColor = mk3(light->color) * inv * inv;
  } else { // sun
    const float4 ld = __ldg(P.lightDerived + lightIndex);
    L = -mk3(ld.x, ld.y, ld.z);
    lightDistance = INFINITY;
    lightColor = mk3(light->color);
  }
  if (!pickedEnvironment) lightColor *= float(pickCount);

  const f3 shadowOrigin = hitPoint + surfaceNormal * 1e-3f;

  if (!kPlain && U.shadingMode == RT_SHADING_LEGACY) { // Lambert branch (Raytracing.metal:649-690)
    const f3 Ln = normalize(L);
    const float NdotL = saturatef(dot(shadingNormal, Ln));
    const f3 legacyColor = s.throughput * albedo;
    if (length(legacyColor) < 0.001f) return false;
    if (length(lightColor) > 0.0001f && NdotL > 0.0f) {
      shadow.valid = true;
      shadow.origin = shadowOrigin;
      shadow.dir = L;
      shadow.tmax = lightDistance - 1e-3f;
      shadow.contribution = legacyColor * lightColor * NdotL;
    }
    s.throughput = legacyColor * ao;
    if (length(s.throughput) < 0.001f) return false;
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
      shadow.valid = true;
      shadow.origin = shadowOrigin;
      shadow.dir = L;
      shadow.tmax = lightDistance - 1e-3f;
      shadow.contribution = s.throughput * direct;
    }
    s.throughput *= albedo * (1.0f - metallic) * ao;
    if (length(s.throughput) < 0.001f) return false;
  }

  // cosine-weighted bounce (Raytracing.metal:763-774); note the stride-5 dimension index
  const f2 r = mk2(halton(hIndex, 2 + s.step * 5 + 3), halton(hIndex, 2 + s.step * 5 + 4));
  const f3 local = sampleCosineWeightedHemisphere(r);
  s.dir = alignHemisphereWithNormal(local, shadingNormal);
  s.bsdfPdf = envLight ? saturatef(dot(shadingNormal, s.dir)) * kInvPi : 0.0f;
  s.origin = shadowOrigin;
  ++s.step;
  ++s.bounce;
  s.transparencyPasses = 0;
  return s.bounce < U.maxBounces;
}

// Extension (rt_b200.h rt_environment): what a path that leaves the scene picks up. Returns black when no
// environment is bound, which is the reference's behaviour.
__device__ __forceinline__ void shadeMiss(const TraceParams &P, PathState &s) {
  if (P.env.texelsDev == nullptr) return;
  if (environmentIsLight(P) && s.bsdfPdf > 0.0f) { // the light sample of the previous hit covers part of this
    const float envPdf = environmentPdf(P.env, s.dir) / float(P.uniforms.lightCount + 1);
    const float weight = s.bsdfPdf / (s.bsdfPdf + envPdf);
    s.radiance += s.throughput * (sampleEnvironment(P.env, s.dir) * weight);
  } else {
    s.radiance += s.throughput * sampleEnvironment(P.env, s.dir);
  }
}

// Motion-adaptive sample count, evaluated after sample 0 (Raytracing.metal:779-789).
__device__ __forceinline__ int adaptiveSampleCount(const rt_uniforms &U, int baseSamples, int maxExtraSamples,
                                                   const f2 &motion, const f2 &prevMotion) {
  const float motionMag = fmaxf(length(motion), length(prevMotion));
  const float low = fmaxf(U.motionSamplingLowThresholdPixels, 0.0f);
  const float high = fmaxf(U.motionSamplingHighThresholdPixels, low + 1e-3f);
  const float t = clampf((motionMag - low) / (high - low), 0.0f, 1.0f);
  int extra = int(roundf(t * float(maxExtraSamples)));
  extra = min(max(extra, 0), maxExtraSamples);
  return baseSamples + extra;
}

// Sample mean, EMA with the history image and all image writes of one pixel (Raytracing.metal:792-829).
__device__ __forceinline__ void resolvePixel(const TraceParams &P, int px, int py, f3 totalColor, int totalSamples,
                                             const f2 &prevMotion, const PrimaryOutputs &prim, bool writeGBuffer) {
  const rt_uniforms &U = P.uniforms;
  const size_t pixelIndex = size_t(py) * size_t(U.width) + size_t(px);
  totalColor = totalColor / float(max(totalSamples, 1));
  if (P.sampleModulo > 1) {
    // sample partition: this dispatch's share of the frame; the shares of all dispatches add up to mix(mean, history, w)
    if (U.frameIndex > 0) {
      const float historyWeight = clampf(U.accumulationWeight, 0.0f, 0.95f);
      totalColor = totalColor * (1.0f - historyWeight);
      if (P.sampleRemainder == 0) {
        const f4 pc = readImage(P.images[RT_TEXTURE_ACCUMULATION], px, py);
        totalColor = totalColor + mk3(pc.x, pc.y, pc.z) * historyWeight;
      }
    }
  } else if (U.frameIndex > 0) {
    const f4 pc = readImage(P.images[RT_TEXTURE_ACCUMULATION], px, py);
    float historyWeight = clampf(U.accumulationWeight, 0.0f, 0.95f);
    if (U.enableMotionAdaptiveAccumulation != 0) {
      const float motionMag = fmaxf(length(prim.motion), length(prevMotion));
      const float low = fmaxf(U.motionAccumulationLowThresholdPixels, 0.0f);
      const float high = fmaxf(U.motionAccumulationHighThresholdPixels, low + 1e-3f);
      const float t = clampf((motionMag - low) / (high - low), 0.0f, 1.0f);
      float minWeight = clampf(U.motionAccumulationMinWeight, 0.0f, 0.95f);
      minWeight = fminf(minWeight, historyWeight);
      historyWeight = mixf(historyWeight, minWeight, t);
    }
    totalColor = mix(totalColor, mk3(pc.x, pc.y, pc.z), historyWeight);
  }
  // alpha is 1 (Raytracing.metal:819); under the sample partition only the share that owns sample 0 carries it, so the sum does
  const f4 outColor = {totalColor.x, totalColor.y, totalColor.z, (P.sampleModulo > 1 && P.sampleRemainder != 0) ? 0.0f : 1.0f};
  writeImage(P.images[RT_TEXTURE_PREVIOUS_ACCUMULATION], px, py, outColor);
  for (int p = 0; p < P.peerCount; ++p) // multi-GPU: publish owned pixels into every rank's frame over NVLink
    if (P.peerAccumulation[p] != nullptr && P.peerAccumulation[p] != P.images[RT_TEXTURE_PREVIOUS_ACCUMULATION].data)
      writeImageAt(P.peerAccumulation[p], P.images[RT_TEXTURE_PREVIOUS_ACCUMULATION].format, pixelIndex, outColor);
  writeImage(P.images[RT_TEXTURE_DEPTH], px, py, {prim.depth, 0.0f, 0.0f, 0.0f});
  writeImage(P.images[RT_TEXTURE_MOTION], px, py, {prim.motion.x, prim.motion.y, 0.0f, 0.0f});
  if (writeGBuffer && U.enableDenoiseGBuffer != 0) {
    writeImage(P.images[RT_TEXTURE_DIFFUSE_ALBEDO], px, py, prim.gDiffuse);
    writeImage(P.images[RT_TEXTURE_SPECULAR_ALBEDO], px, py, prim.gSpecular);
    writeImage(P.images[RT_TEXTURE_NORMAL], px, py, prim.gNormal);
    writeImage(P.images[RT_TEXTURE_ROUGHNESS], px, py, prim.gRoughness);
  }
}

__device__ __forceinline__ PrimaryOutputs emptyPrimaryOutputs() {
  PrimaryOutputs p;
  p.depth = 1.0e8f;
  p.motion = mk2(0.0f, 0.0f);
  p.hadPrimaryHit = false;
  p.wroteGBuffer = false;
  p.gDiffuse = p.gSpecular = p.gNormal = p.gRoughness = {0.0f, 0.0f, 0.0f, 0.0f};
  return p;
}

int fillTraceParams(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], const rt_image textures[RT_TEXTURE_COUNT],
                    int maxSubmeshes, const rt_trace_options *opt, TraceParams &P);
int launchTraceWavefront(rt_context *ctx, const TraceParams &P);

} // namespace rtb
