/*
 * rt_types.h — frozen binary layouts shared by host code, the CPU oracle and the sm_100a kernels.
 *
 * These are the structs and binding indices the reference shares between its Swift host and its Metal
 * kernels (MetalRaytracing/ShaderTypes.h:35-145), re-typed without <simd/simd.h>: a `vector_float3` there
 * is 16 bytes in size and alignment, which is what `rt_float3` reproduces. The instance descriptor is the
 * Apple SDK layout the reference fills in MetalRaytracing/Renderer.swift:547-556 and reads in
 * MetalRaytracing/Raytracing.metal:331-333; the `Resource` row is MetalRaytracing/Raytracing.metal:168-183.
 *
 * Every offset is pinned by a static assertion so a drift breaks the build, not the image.
 * Compiles as C11, C++17 and CUDA C++.
 */
#ifndef RT_TYPES_H
#define RT_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define RT_STATIC_ASSERT(c, m) static_assert(c, m)
#define RT_ALIGNAS(n) alignas(n)
#else
#define RT_STATIC_ASSERT(c, m) _Static_assert(c, m)
#define RT_ALIGNAS(n) _Alignas(n)
#endif

/* ---- binding indices (ShaderTypes.h:35-70) --------------------------------------------------------- */
enum rt_buffer_index {
  RT_BUFFER_UNIFORMS = 0,
  RT_BUFFER_INSTANCE_ACCELERATION_STRUCTURE = 1, /* legacy, never bound */
  RT_BUFFER_RANDOM = 2,                          /* legacy, never bound */
  RT_BUFFER_VERTEX_COLOR = 3,                    /* legacy, never bound */
  RT_BUFFER_VERTEX_NORMALS = 4,                  /* legacy, never bound */
  RT_BUFFER_RESOURCES = 5,
  RT_BUFFER_LIGHTS = 6,
  RT_BUFFER_INSTANCES = 7,                       /* legacy, never bound */
  RT_BUFFER_ACCELERATION_STRUCTURE = 8,
  RT_BUFFER_INSTANCE_DESCRIPTORS = 9,
  RT_BUFFER_REST_POSITIONS = 10,
  RT_BUFFER_REST_NORMALS = 11,
  RT_BUFFER_JOINT_INDICES = 12,
  RT_BUFFER_JOINT_WEIGHTS = 13,
  RT_BUFFER_JOINT_MATRICES = 14,
  RT_BUFFER_SKINNED_POSITIONS = 15,
  RT_BUFFER_SKINNED_NORMALS = 16,
  RT_BUFFER_PREVIOUS_INSTANCE_DESCRIPTORS = 17,
  RT_BUFFER_COUNT = 18
};

enum rt_texture_index {
  RT_TEXTURE_ACCUMULATION = 0,          /* READ by the kernel (history), despite the name */
  RT_TEXTURE_PREVIOUS_ACCUMULATION = 1, /* WRITTEN by the kernel; host swaps 0<->1 after each frame */
  RT_TEXTURE_RANDOM = 2,
  RT_TEXTURE_DEPTH = 3,
  RT_TEXTURE_MOTION = 4,
  RT_TEXTURE_DIFFUSE_ALBEDO = 5,
  RT_TEXTURE_SPECULAR_ALBEDO = 6,
  RT_TEXTURE_NORMAL = 7,
  RT_TEXTURE_ROUGHNESS = 8,
  RT_TEXTURE_COUNT = 9
};

enum rt_light_type {
  RT_LIGHT_UNUSED = 0,
  RT_LIGHT_SUN = 1,
  RT_LIGHT_SPOT = 2,
  RT_LIGHT_POINT = 3,
  RT_LIGHT_AREA = 4
};

enum rt_shading_mode { RT_SHADING_PBR = 0, RT_SHADING_LEGACY = 1 };

enum rt_debug_texture_mode {
  RT_DEBUG_NONE = 0,
  RT_DEBUG_BASECOLOR = 1,
  RT_DEBUG_NORMAL = 2,
  RT_DEBUG_ROUGHNESS = 3,
  RT_DEBUG_METALLIC = 4,
  RT_DEBUG_AO = 5,
  RT_DEBUG_EMISSION = 6,
  RT_DEBUG_MOTION = 7
};

#define RT_MATERIAL_TEXTURE_BASECOLOR (1u << 0)
#define RT_MATERIAL_TEXTURE_NORMAL (1u << 1)
#define RT_MATERIAL_TEXTURE_ROUGHNESS (1u << 2)
#define RT_MATERIAL_TEXTURE_METALLIC (1u << 3)
#define RT_MATERIAL_TEXTURE_AO (1u << 4) /* compiled out in the reference build (ENABLE_AO == 0) */
#define RT_MATERIAL_TEXTURE_EMISSION (1u << 5)
#define RT_MATERIAL_TEXTURE_OPACITY (1u << 6)

/* ---- vector_float3: 16-byte size and alignment ------------------------------------------------------ */
#ifdef __cplusplus
typedef struct alignas(16) rt_float3 {
  float x, y, z, _pad;
} rt_float3;
#else
typedef struct rt_float3 {
  _Alignas(16) float x;
  float y, z, _pad;
} rt_float3;
#endif
RT_STATIC_ASSERT(sizeof(rt_float3) == 16, "vector_float3 is 16 bytes");

/* ---- Camera (ShaderTypes.h:80-85) ------------------------------------------------------------------- */
typedef struct rt_camera {
  rt_float3 position;
  rt_float3 right;   /* pre-scaled by tan(fov/2)*aspect (Scene.swift:126-159) */
  rt_float3 up;      /* pre-scaled by tan(fov/2) */
  rt_float3 forward; /* unit */
} rt_camera;
RT_STATIC_ASSERT(sizeof(rt_camera) == 64, "Camera");
RT_STATIC_ASSERT(offsetof(rt_camera, right) == 16 && offsetof(rt_camera, up) == 32 &&
                     offsetof(rt_camera, forward) == 48,
                 "Camera offsets");

/* ---- Light (ShaderTypes.h:95-106) ------------------------------------------------------------------- */
typedef struct rt_light {
  int32_t type; /* rt_light_type */
  int32_t _pad0[3];
  rt_float3 position;
  rt_float3 color;
  rt_float3 forward; /* area light */
  rt_float3 right;
  rt_float3 up;
  float coneAngle; /* spot light */
  float _pad1[3];
  rt_float3 direction;
} rt_light;
RT_STATIC_ASSERT(sizeof(rt_light) == 128, "Light");
RT_STATIC_ASSERT(offsetof(rt_light, position) == 16 && offsetof(rt_light, color) == 32 &&
                     offsetof(rt_light, forward) == 48 && offsetof(rt_light, right) == 64 &&
                     offsetof(rt_light, up) == 80 && offsetof(rt_light, coneAngle) == 96 &&
                     offsetof(rt_light, direction) == 112,
                 "Light offsets");

/* ---- Uniforms (ShaderTypes.h:108-130) --------------------------------------------------------------- */
typedef struct rt_uniforms {
  int32_t width;
  int32_t height;
  int32_t blocksWide; /* written by the host, never read by the kernel */
  uint32_t frameIndex;
  int32_t lightCount;
  int32_t samplesPerPixel;
  int32_t maxBounces;
  int32_t _pad0;
  rt_camera camera;
  rt_camera previousCamera;
  int32_t debugTextureMode;
  float accumulationWeight;
  int32_t enableDenoiseGBuffer;
  int32_t shadingMode;
  int32_t enableMotionAdaptiveAccumulation;
  float motionAccumulationMinWeight;
  float motionAccumulationLowThresholdPixels;
  float motionAccumulationHighThresholdPixels;
  int32_t enableMotionAdaptiveSampling;
  int32_t motionSamplingMaxExtraSamples;
  float motionSamplingLowThresholdPixels;
  float motionSamplingHighThresholdPixels;
} rt_uniforms;
RT_STATIC_ASSERT(sizeof(rt_uniforms) == 208, "Uniforms");
RT_STATIC_ASSERT(offsetof(rt_uniforms, frameIndex) == 12 && offsetof(rt_uniforms, maxBounces) == 24 &&
                     offsetof(rt_uniforms, camera) == 32 && offsetof(rt_uniforms, previousCamera) == 96 &&
                     offsetof(rt_uniforms, debugTextureMode) == 160 &&
                     offsetof(rt_uniforms, accumulationWeight) == 164 &&
                     offsetof(rt_uniforms, shadingMode) == 172 &&
                     offsetof(rt_uniforms, enableMotionAdaptiveSampling) == 192 &&
                     offsetof(rt_uniforms, motionSamplingHighThresholdPixels) == 204,
                 "Uniforms offsets");

/* ---- Material (ShaderTypes.h:137-145) --------------------------------------------------------------- */
typedef struct rt_material {
  rt_float3 baseColor;
  rt_float3 specular; /* loaded, never read by the kernel */
  rt_float3 emission;
  float specularExponent; /* loaded, never read by the kernel */
  float refractionIndex;
  float opacity;
  uint32_t textureFlags;
} rt_material;
RT_STATIC_ASSERT(sizeof(rt_material) == 64, "Material");
RT_STATIC_ASSERT(offsetof(rt_material, emission) == 32 && offsetof(rt_material, specularExponent) == 48 &&
                     offsetof(rt_material, refractionIndex) == 52 && offsetof(rt_material, opacity) == 56 &&
                     offsetof(rt_material, textureFlags) == 60,
                 "Material offsets");

/* ---- MTLIndirectAccelerationStructureInstanceDescriptor (72 B; Renderer.swift:547-556,1393-1401) ---- */
typedef struct rt_instance_descriptor {
  float transformationMatrix[4][3]; /* [column][row]: 4 columns x packed float3, rows 0..2 of a 4x4 */
  uint32_t options;
  uint32_t mask;
  uint32_t intersectionFunctionTableOffset;
  uint32_t userID;
  uint64_t accelerationStructureID; /* handle returned by rt_blas_build */
} rt_instance_descriptor;
RT_STATIC_ASSERT(sizeof(rt_instance_descriptor) == 72, "instance descriptor");
RT_STATIC_ASSERT(offsetof(rt_instance_descriptor, options) == 48 &&
                     offsetof(rt_instance_descriptor, mask) == 52 &&
                     offsetof(rt_instance_descriptor, userID) == 60 &&
                     offsetof(rt_instance_descriptor, accelerationStructureID) == 64,
                 "instance descriptor offsets");

/* ---- sampled material texture (stands in for a texture2d<float> argument-buffer slot) --------------- */
/* A Resource texture slot holds an 8-byte handle. Here the handle is the address of one of these records
 * (device address on the GPU path, host address in the oracle). Texels are RGBA8, row 0 first; `srgb` != 0
 * means R,G,B are sRGB-encoded and decoded to linear on fetch (MTKTextureLoader .SRGB option,
 * SubMesh.swift:79-96). Filtering is bilinear at LOD 0 with repeat addressing (Raytracing.metal:421). */
typedef struct rt_texture2d {
  const uint8_t *texels; /* width*height*4 bytes */
  int32_t width;
  int32_t height;
  int32_t srgb;
  int32_t _pad;
} rt_texture2d;
RT_STATIC_ASSERT(sizeof(rt_texture2d) == 24, "texture record");

/* ---- Resource argument-buffer row (Raytracing.metal:168-183), 13 x 8 B = 104 B ---------------------- */
typedef struct rt_resource {
  const rt_float3 *positions;         /* id 0: float3 stride 16 */
  const rt_float3 *previousPositions; /* id 1 */
  const rt_float3 *normals;           /* id 2 */
  const int32_t *indices;             /* id 3: 3 per triangle, 32-bit */
  const rt_material *material;        /* id 4 */
  const float *uvs;                   /* id 5: float2 stride 8 (normals buffer reused as dummy if none) */
  const rt_texture2d *baseColorMap;   /* id 6 */
  const rt_texture2d *normalMap;      /* id 7 */
  const rt_texture2d *roughnessMap;   /* id 8 */
  const rt_texture2d *metallicMap;    /* id 9 */
  const rt_texture2d *aoMap;          /* id 10 */
  const rt_texture2d *opacityMap;     /* id 11 */
  const rt_texture2d *emissionMap;    /* id 12 */
} rt_resource;
RT_STATIC_ASSERT(sizeof(rt_resource) == 104, "Resource");

/* ---- render-target image bound at a TextureIndex ---------------------------------------------------- */
enum rt_image_format {
  RT_FORMAT_NONE = 0,
  RT_FORMAT_R32_UINT = 1,   /* random seeds */
  RT_FORMAT_R32_FLOAT = 2,  /* depth */
  RT_FORMAT_RG16_FLOAT = 3, /* motion (reference format) */
  RT_FORMAT_RGBA16_FLOAT = 4, /* accumulation, G-buffer (reference format) */
  RT_FORMAT_R16_FLOAT = 5,  /* roughness */
  RT_FORMAT_RG32_FLOAT = 6, /* fp32 alternatives, selectable per image */
  RT_FORMAT_RGBA32_FLOAT = 7
};

typedef struct rt_image {
  void *data;      /* tightly packed rows, row 0 first */
  int32_t width;
  int32_t height;
  int32_t format;  /* rt_image_format */
  int32_t _pad;
} rt_image;
RT_STATIC_ASSERT(sizeof(rt_image) == 24, "image record");

/* ---- acceleration-structure geometry descriptor (Mesh.swift:84-101) ---------------------------------- */
typedef struct rt_triangle_geometry {
  const void *vertexBuffer; /* float3 positions */
  uint32_t vertexStride;    /* 16 in the reference */
  uint32_t vertexCount;
  const void *indexBuffer;  /* u16 or u32 */
  uint32_t indexStride;     /* 2 or 4 */
  uint32_t triangleCount;
} rt_triangle_geometry;

#define RT_AS_FLAG_COMPACT 1u    /* static mesh: build best-quality, drop build scratch */
#define RT_AS_FLAG_REFITTABLE 2u /* skinned mesh: keep topology + scratch so rt_blas_refit works */

#endif /* RT_TYPES_H */
