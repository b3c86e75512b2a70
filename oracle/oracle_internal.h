// oracle_internal.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
#pragma once
#include <algorithm>
#include <cstdint>

#include "../include/rt_b200.h"
#include "../include/rt_types.h"
#include "oracle_bvh.h"
#include "oracle_math.h"

namespace orc {

// The kernel's argument table (Raytracing.metal:221-236) with host pointers.
struct KernelArgs {
  const rt_uniforms *uniforms;            // buffer 0
  const Tlas *tlas;                       // buffer 8
  const rt_resource *resources;           // buffer 5
  const rt_instance_descriptor *instances;     // buffer 9
  const rt_instance_descriptor *prevInstances; // buffer 17
  const rt_light *lights;                 // buffer 6
  rt_image textures[RT_TEXTURE_COUNT];    // textures 0..8
  int maxSubmeshes;                       // function constant 1
  uint32_t *primaryIds;                   // optional probe: 4 x u32 per pixel (instance, geometry, primitive, t bits)
  rt_environment env{};                   // extension (rt_b200.h): texelsDev == nullptr means off (reference behaviour)
  bool enableAO = false;                  // the reference's compile-time ENABLE_AO (ShaderTypes.h:155-157), default 0
};

struct PixelStats {
  uint64_t closestRays = 0, anyRays = 0, hits = 0;
};

float halton(int i, int d);
float4 sampleTexture(const rt_texture2d *t, float2 uv);
float3 sampleEnvironment(const rt_environment &env, float3 d);
void raytracingKernelPixel(int tidx, int tidy, const KernelArgs &a, PixelStats &stats);
void skinningKernelVertex(uint32_t vertexID, const void *const *buffers, uint32_t vertexCount);

} // namespace orc
