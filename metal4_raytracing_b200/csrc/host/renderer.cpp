// renderer.cpp — host-side frame orchestrator (include/rt_renderer.h) on top of the rt_* C-ABI.
// Mirrors the hot-path half of MetalRaytracing/Renderer.swift and MetalRaytracing/SkinningPass.swift; every GPU
// action goes through rt_b200.h, so this file contains no CUDA and no compute of its own.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/rt_renderer.h"

namespace {

thread_local std::string g_err;

struct DeviceMesh {
  uint32_t vertexCount = 0;
  bool skinned = false;
  uint32_t jointCount = 0;
  void *restPositions = nullptr, *restNormals = nullptr; // also the live streams of a static mesh
  void *positions = nullptr, *prevPositions = nullptr, *normals = nullptr; // live streams the kernel reads
  void *uvs = nullptr;
  void *jointIndices = nullptr, *jointWeights = nullptr, *jointMatrices = nullptr;
  void *jointParents = nullptr, *jointInverseBind = nullptr, *jointLocalTRS = nullptr; // RTR_FLAG_GPU_SKELETON
  bool inverseBindUploaded = false;
  std::vector<void *> indices;
  std::vector<uint32_t> triangleCounts;
  void *materials = nullptr; // rt_material per submesh
  uint64_t blas = 0;
  std::vector<rt_triangle_geometry> geoms;
};

size_t formatBytes(int f) {
  switch (f) {
    case RT_FORMAT_R32_UINT:
    case RT_FORMAT_R32_FLOAT:
    case RT_FORMAT_RG16_FLOAT: return 4;
    case RT_FORMAT_RGBA16_FLOAT:
    case RT_FORMAT_RG32_FLOAT: return 8;
    case RT_FORMAT_R16_FLOAT: return 2;
    case RT_FORMAT_RGBA32_FLOAT: return 16;
    default: return 0;
  }
}

} // namespace

struct rtr_renderer {
  rt_context *ctx = nullptr;
  int width = 0, height = 0;
  uint32_t flags = 0;
  uint32_t maxSubmeshes = 1;
  std::vector<DeviceMesh> meshes;
  std::vector<const rt_texture2d *> textures;
  std::vector<uint32_t> instanceMesh;
  void *resources = nullptr;
  void *descriptors = nullptr, *prevDescriptors = nullptr;
  void *lights = nullptr;
  uint32_t lightCapacity = 0;
  uint32_t lightCount = 0; // lights resident in `lights` (what uniforms.lightCount may address)
  uint64_t tlas = 0;
  rt_image images[RT_TEXTURE_COUNT]{};
  // pinned staging for per-frame uploads: a ring of three arenas (the reference triple-buffers its per-frame host
  // data, Renderer.swift:208-212), each guarded by a stream fence, so rtr_update never waits for the GPU to drain
  static constexpr int kStageSlots = 3;
  uint8_t *stage[kStageSlots] = {};
  uint64_t stageFence[kStageSlots] = {};
  size_t stageBytes = 0, stageUsed = 0;
  uint64_t updates = 0;
  rt_instance_descriptor *stageDescriptors = nullptr; // inside the current arena
  size_t stagePaletteFloats = 0;                       // sum over skinned meshes of 16 floats per joint
  bool untextured = true; // no submesh material has a textureFlags bit: rtr_draw passes RT_TRACE_HINT_UNTEXTURED
  bool noGlass = true;    // no submesh material can take the glass branch: rtr_draw passes RT_TRACE_HINT_NO_GLASS
};

#define RTR_TRY(expr)                  \
  do {                                 \
    int _r = (expr);                   \
    if (_r != 0) {                     \
      g_err = rt_last_error();         \
      return _r;                       \
    }                                  \
  } while (0)

static int uploadNew(rt_context *ctx, const void *src, size_t bytes, void **dev) {
  int r = rt_malloc(ctx, bytes, dev);
  if (r) return r;
  return bytes ? rt_upload(ctx, *dev, src, bytes) : 0;
}

static void packDescriptor(const float m[16], uint64_t blas, rt_instance_descriptor &d) {
  // packedFloat4x3 (Renderer.swift:1393-1401): columns 0..3, rows 0..2 ; mask 0xFF, options 0 (:547-556)
  std::memset(&d, 0, sizeof d);
  for (int c = 0; c < 4; ++c)
    for (int row = 0; row < 3; ++row) d.transformationMatrix[c][row] = m[c * 4 + row];
  d.mask = 0xFF;
  d.accelerationStructureID = blas;
}

static void fillGeometries(DeviceMesh &dm) {
  dm.geoms.resize(dm.indices.size());
  for (size_t k = 0; k < dm.indices.size(); ++k) {
    dm.geoms[k].vertexBuffer = dm.positions;
    dm.geoms[k].vertexStride = 16;
    dm.geoms[k].vertexCount = dm.vertexCount;
    dm.geoms[k].indexBuffer = dm.indices[k];
    dm.geoms[k].indexStride = 4;
    dm.geoms[k].triangleCount = dm.triangleCounts[k];
  }
}

static int skinMesh(rtr_renderer *r, DeviceMesh &dm) {
  const void *table[RT_BUFFER_COUNT] = {};
  table[RT_BUFFER_REST_POSITIONS] = dm.restPositions;
  table[RT_BUFFER_REST_NORMALS] = dm.restNormals;
  table[RT_BUFFER_JOINT_INDICES] = dm.jointIndices;
  table[RT_BUFFER_JOINT_WEIGHTS] = dm.jointWeights;
  table[RT_BUFFER_JOINT_MATRICES] = dm.jointMatrices;
  table[RT_BUFFER_SKINNED_POSITIONS] = dm.positions;
  table[RT_BUFFER_SKINNED_NORMALS] = dm.normals;
  return rt_skin(r->ctx, table, dm.vertexCount);
}

extern "C" {

const char *rtr_last_error(void) { return g_err.c_str(); }

int rtr_create(rt_context *ctx, const rt_scene_desc *scene, int width, int height, uint32_t flags,
               rtr_renderer **out) {
  if (!ctx || !scene || !out || width <= 0 || height <= 0) {
    g_err = "rtr_create: bad arguments";
    return 2;
  }
  rtr_renderer *r = new rtr_renderer();
  r->ctx = ctx;
  r->width = width;
  r->height = height;
  r->flags = flags;
  r->maxSubmeshes = scene->maxSubmeshes ? scene->maxSubmeshes : 1;
  *out = r;
  // textures
  r->textures.resize(scene->textureCount);
  for (uint32_t i = 0; i < scene->textureCount; ++i)
    RTR_TRY(rt_texture_create(ctx, scene->textures[i].texels, scene->textures[i].width, scene->textures[i].height,
                              scene->textures[i].srgb, &r->textures[i]));
  // meshes: vertex streams, indices, materials; skinned meshes get their own live streams + palette
  r->meshes.resize(scene->meshCount);
  for (uint32_t m = 0; m < scene->meshCount; ++m) {
    const rt_scene_mesh &sm = scene->meshes[m];
    DeviceMesh &dm = r->meshes[m];
    dm.vertexCount = sm.vertexCount;
    size_t vb = size_t(sm.vertexCount) * 16;
    RTR_TRY(uploadNew(ctx, sm.positions, vb, &dm.restPositions));
    RTR_TRY(uploadNew(ctx, sm.normals, vb, &dm.restNormals));
    if (sm.uvs) RTR_TRY(uploadNew(ctx, sm.uvs, size_t(sm.vertexCount) * 8, &dm.uvs));
    dm.skinned = sm.jointIndices != nullptr && sm.jointCount > 0;
    dm.jointCount = sm.jointCount;
    if (dm.skinned) {
      RTR_TRY(uploadNew(ctx, sm.jointIndices, size_t(sm.vertexCount) * 8, &dm.jointIndices));
      RTR_TRY(uploadNew(ctx, sm.jointWeights, size_t(sm.vertexCount) * 16, &dm.jointWeights));
      RTR_TRY(uploadNew(ctx, sm.jointMatrices, size_t(sm.jointCount) * 64, &dm.jointMatrices));
      if ((flags & RTR_FLAG_GPU_SKELETON) && sm.jointParents && sm.jointCount <= 1024) {
        RTR_TRY(uploadNew(ctx, sm.jointParents, size_t(sm.jointCount) * 4, &dm.jointParents));
        RTR_TRY(rt_malloc(ctx, size_t(sm.jointCount) * 64, &dm.jointInverseBind));
        RTR_TRY(rt_malloc(ctx, size_t(sm.jointCount) * 40, &dm.jointLocalTRS));
      }
      RTR_TRY(rt_malloc(ctx, vb, &dm.positions));
      RTR_TRY(rt_malloc(ctx, vb, &dm.prevPositions));
      RTR_TRY(rt_malloc(ctx, vb, &dm.normals));
      RTR_TRY(skinMesh(r, dm)); // initial skinning pass (Renderer.swift:470-494)
      RTR_TRY(rt_copy(ctx, dm.prevPositions, dm.positions, vb));
      r->stagePaletteFloats += size_t(sm.jointCount) * 16;
    } else {
      dm.positions = dm.restPositions;
      dm.prevPositions = dm.restPositions; // SubMesh.swift:60: previousPositionBuffer = positionBuffer
      dm.normals = dm.restNormals;
    }
    std::vector<rt_material> mats(sm.submeshCount);
    dm.indices.resize(sm.submeshCount);
    dm.triangleCounts.resize(sm.submeshCount);
    for (uint32_t k = 0; k < sm.submeshCount; ++k) {
      mats[k] = sm.submeshes[k].material;
      if (mats[k].textureFlags != 0u) r->untextured = false;
      // the kernel's test is clamp(opacity [* map]) < 0.999 || max(ior, 1) > 1.01 (Raytracing.metal:517-519); an
      // opacity map can only lower opacity, so a bound one counts as possible glass
      const float opacity = std::min(std::max(mats[k].opacity, 0.0f), 1.0f);
      if (!(opacity >= 0.999f) || mats[k].refractionIndex > 1.01f || (mats[k].textureFlags & RT_MATERIAL_TEXTURE_OPACITY))
        r->noGlass = false;
      dm.triangleCounts[k] = sm.submeshes[k].triangleCount;
      RTR_TRY(uploadNew(ctx, sm.submeshes[k].indices, size_t(sm.submeshes[k].triangleCount) * 12, &dm.indices[k]));
    }
    RTR_TRY(uploadNew(ctx, mats.data(), mats.size() * sizeof(rt_material), &dm.materials));
    fillGeometries(dm);
    RTR_TRY(rt_blas_build(ctx, dm.geoms.data(), uint32_t(dm.geoms.size()),
                          dm.skinned ? RT_AS_FLAG_REFITTABLE : RT_AS_FLAG_COMPACT, &dm.blas));
  }
  // resource rows: instance * maxSubmeshes + submesh (Renderer.swift:346-409)
  std::vector<rt_resource> rows(size_t(scene->instanceCount) * r->maxSubmeshes);
  std::memset(rows.data(), 0, rows.size() * sizeof(rt_resource));
  r->instanceMesh.resize(scene->instanceCount);
  for (uint32_t i = 0; i < scene->instanceCount; ++i) {
    uint32_t m = scene->instances[i].meshIndex;
    if (m >= scene->meshCount) {
      g_err = "rtr_create: instance references a mesh out of range";
      return 2;
    }
    r->instanceMesh[i] = m;
    const rt_scene_mesh &sm = scene->meshes[m];
    DeviceMesh &dm = r->meshes[m];
    for (uint32_t k = 0; k < sm.submeshCount; ++k) {
      rt_resource &row = rows[size_t(i) * r->maxSubmeshes + k];
      row.positions = static_cast<const rt_float3 *>(dm.positions);
      row.previousPositions = static_cast<const rt_float3 *>(dm.prevPositions);
      row.normals = static_cast<const rt_float3 *>(dm.normals);
      row.indices = static_cast<const int32_t *>(dm.indices[k]);
      row.material = static_cast<const rt_material *>(dm.materials) + k;
      row.uvs = dm.uvs ? static_cast<const float *>(dm.uvs) : static_cast<const float *>(dm.normals);
      const int32_t *ti = sm.submeshes[k].textureIndex;
      for (int slot = 0; slot < RT_SLOT_COUNT; ++slot)
        if (ti[slot] < 0 || uint32_t(ti[slot]) >= scene->textureCount) {
          g_err = "rtr_create: mesh " + std::to_string(m) + " submesh " + std::to_string(k) + " texture slot " +
                  std::to_string(slot) + " has index " + std::to_string(ti[slot]) + " outside the scene's " +
                  std::to_string(scene->textureCount) + " textures (rt_scene.h: every slot names a texture, 1x1 fallbacks included)";
          return 2;
        }
      row.baseColorMap = r->textures[ti[RT_SLOT_BASECOLOR]];
      row.normalMap = r->textures[ti[RT_SLOT_NORMAL]];
      row.roughnessMap = r->textures[ti[RT_SLOT_ROUGHNESS]];
      row.metallicMap = r->textures[ti[RT_SLOT_METALLIC]];
      row.aoMap = r->textures[ti[RT_SLOT_AO]];
      row.opacityMap = r->textures[ti[RT_SLOT_OPACITY]];
      row.emissionMap = r->textures[ti[RT_SLOT_EMISSION]];
    }
  }
  RTR_TRY(uploadNew(ctx, rows.data(), rows.size() * sizeof(rt_resource), &r->resources));
  // instance descriptors (current + previous) and TLAS
  size_t descBytes = size_t(scene->instanceCount) * sizeof(rt_instance_descriptor);
  r->lightCapacity = scene->lightCount ? scene->lightCount : 1;
  r->stageBytes = ((descBytes + 255) & ~size_t(255)) + ((size_t(r->lightCapacity) * sizeof(rt_light) + 255) & ~size_t(255)) +
                  r->stagePaletteFloats * 4 + 256 * (scene->meshCount + 2);
  for (int k = 0; k < rtr_renderer::kStageSlots; ++k)
    RTR_TRY(rt_malloc_host(ctx, r->stageBytes, reinterpret_cast<void **>(&r->stage[k])));
  r->stageDescriptors = reinterpret_cast<rt_instance_descriptor *>(r->stage[0]);
  RTR_TRY(rt_malloc(ctx, descBytes, &r->descriptors));
  RTR_TRY(rt_malloc(ctx, descBytes, &r->prevDescriptors));
  for (uint32_t i = 0; i < scene->instanceCount; ++i)
    packDescriptor(scene->instances[i].previousTransform, r->meshes[r->instanceMesh[i]].blas, r->stageDescriptors[i]);
  RTR_TRY(rt_upload(ctx, r->prevDescriptors, r->stageDescriptors, descBytes));
  RTR_TRY(rt_sync(ctx));
  for (uint32_t i = 0; i < scene->instanceCount; ++i)
    packDescriptor(scene->instances[i].transform, r->meshes[r->instanceMesh[i]].blas, r->stageDescriptors[i]);
  RTR_TRY(rt_upload(ctx, r->descriptors, r->stageDescriptors, descBytes));
  RTR_TRY(rt_tlas_build(ctx, static_cast<const rt_instance_descriptor *>(r->descriptors), scene->instanceCount, &r->tlas));
  // lights
  RTR_TRY(rt_malloc(ctx, size_t(r->lightCapacity) * sizeof(rt_light), &r->lights));
  if (scene->lightCount) {
    RTR_TRY(rt_sync(ctx)); // the descriptor upload above still reads arena 0
    std::memcpy(r->stage[1], scene->lights, size_t(scene->lightCount) * sizeof(rt_light));
    RTR_TRY(rt_upload(ctx, r->lights, r->stage[1], size_t(scene->lightCount) * sizeof(rt_light)));
  }
  r->lightCount = scene->lightCount;
  // images (Renderer.swift:685-799)
  const bool fp32 = (flags & RTR_FLAG_FP32_IMAGES) != 0;
  const int rgba = fp32 ? RT_FORMAT_RGBA32_FLOAT : RT_FORMAT_RGBA16_FLOAT;
  const int rg = fp32 ? RT_FORMAT_RG32_FLOAT : RT_FORMAT_RG16_FLOAT;
  const int formats[RT_TEXTURE_COUNT] = {rgba, rgba, RT_FORMAT_R32_UINT, RT_FORMAT_R32_FLOAT, rg, rgba, rgba, rgba,
                                         fp32 ? RT_FORMAT_R32_FLOAT : RT_FORMAT_R16_FLOAT};
  for (int i = 0; i < RT_TEXTURE_COUNT; ++i) {
    size_t bytes = size_t(width) * height * formatBytes(formats[i]);
    void *p = nullptr;
    RTR_TRY(rt_malloc(ctx, bytes, &p));
    RTR_TRY(rt_memset(ctx, p, 0, bytes)); // the reference never clears motionTex before its first read (F12)
    r->images[i] = {p, width, height, formats[i], 0};
  }
  RTR_TRY(rt_sync(ctx));
  return 0;
}

int rtr_destroy(rtr_renderer *r) {
  if (!r) return 0;
  rt_context *ctx = r->ctx;
  rt_sync(ctx);
  if (r->tlas) rt_tlas_destroy(ctx, r->tlas);
  for (auto &dm : r->meshes) {
    if (dm.blas) rt_blas_destroy(ctx, dm.blas);
    rt_free(ctx, dm.restPositions);
    rt_free(ctx, dm.restNormals);
    rt_free(ctx, dm.uvs);
    if (dm.skinned) {
      rt_free(ctx, dm.positions);
      rt_free(ctx, dm.prevPositions);
      rt_free(ctx, dm.normals);
      rt_free(ctx, dm.jointIndices);
      rt_free(ctx, dm.jointWeights);
      rt_free(ctx, dm.jointMatrices);
      rt_free(ctx, dm.jointParents);
      rt_free(ctx, dm.jointInverseBind);
      rt_free(ctx, dm.jointLocalTRS);
    }
    for (void *p : dm.indices) rt_free(ctx, p);
    rt_free(ctx, dm.materials);
  }
  for (auto *t : r->textures) rt_texture_destroy(ctx, t);
  rt_free(ctx, r->resources);
  rt_free(ctx, r->descriptors);
  rt_free(ctx, r->prevDescriptors);
  rt_free(ctx, r->lights);
  for (auto &img : r->images) rt_free(ctx, img.data);
  for (uint8_t *p : r->stage) rt_free_host(ctx, p);
  delete r;
  return 0;
}

int rtr_set_seeds(rtr_renderer *r, const uint32_t *seedsHost) {
  RTR_TRY(rt_upload(r->ctx, r->images[RT_TEXTURE_RANDOM].data, seedsHost, size_t(r->width) * r->height * 4));
  RTR_TRY(rt_sync(r->ctx));
  return 0;
}

int rtr_update(rtr_renderer *r, const rt_scene_desc *scene) {
  rt_context *ctx = r->ctx;
  if (scene->meshCount != r->meshes.size() || scene->instanceCount != r->instanceMesh.size()) {
    g_err = "rtr_update: scene topology changed; create a new renderer";
    return 2;
  }
  // updateInstanceDescriptors (Renderer.swift:937-973): previous <- current, current <- mesh transforms
  size_t descBytes = size_t(scene->instanceCount) * sizeof(rt_instance_descriptor);
  // this update's staging arena: wait only for the uploads that used it three updates ago
  const int slot = int(r->updates % rtr_renderer::kStageSlots);
  if (r->stageFence[slot]) RTR_TRY(rt_fence_wait(ctx, r->stageFence[slot]));
  uint8_t *arena = r->stage[slot];
  size_t used = 0;
  auto take = [&](size_t bytes) {
    uint8_t *p = arena + used;
    used += (bytes + 255) & ~size_t(255);
    return p;
  };
  r->stageDescriptors = reinterpret_cast<rt_instance_descriptor *>(take(descBytes));
  RTR_TRY(rt_copy(ctx, r->prevDescriptors, r->descriptors, descBytes));
  for (uint32_t i = 0; i < scene->instanceCount; ++i)
    packDescriptor(scene->instances[i].transform, r->meshes[r->instanceMesh[i]].blas, r->stageDescriptors[i]);
  RTR_TRY(rt_upload(ctx, r->descriptors, r->stageDescriptors, descBytes));
  if (scene->lightCount > r->lightCapacity) {
    // more lights than the renderer was created with: the staging arenas and the device array were sized for
    // lightCapacity (the reference's light list is fixed after Scene.init, Scene.swift:82-93)
    g_err = "rtr_update: scene has " + std::to_string(scene->lightCount) + " lights, renderer was created for " +
            std::to_string(r->lightCapacity) + "; create a new renderer";
    return 2;
  }
  r->lightCount = scene->lightCount;
  if (scene->lightCount) {
    uint8_t *lights = take(size_t(scene->lightCount) * sizeof(rt_light));
    std::memcpy(lights, scene->lights, size_t(scene->lightCount) * sizeof(rt_light));
    RTR_TRY(rt_upload(ctx, r->lights, lights, size_t(scene->lightCount) * sizeof(rt_light)));
  }
  // skinned meshes: prev <- cur, new palette, skin, refit (Renderer.swift:1290-1326)
  for (uint32_t m = 0; m < scene->meshCount; ++m) {
    DeviceMesh &dm = r->meshes[m];
    if (!dm.skinned) continue;
    const rt_scene_mesh &sm = scene->meshes[m];
    size_t vb = size_t(dm.vertexCount) * 16;
    RTR_TRY(rt_copy(ctx, dm.prevPositions, dm.positions, vb));
    uint8_t *palette = take(size_t(dm.jointCount) * 64);
    if (dm.jointLocalTRS && sm.jointLocalTRS && sm.jointInverseBind) {
      // palette on the device: 40 B per joint of local TRS go up instead of a 64 B matrix, hierarchy + inverse bind
      // products run in rt_joint_palette (bit-identical to the host palette)
      if (!dm.inverseBindUploaded) {
        std::memcpy(palette, sm.jointInverseBind, size_t(dm.jointCount) * 64);
        RTR_TRY(rt_upload(ctx, dm.jointInverseBind, palette, size_t(dm.jointCount) * 64));
        RTR_TRY(rt_sync(ctx)); // once per mesh: the same staging bytes are reused just below
        dm.inverseBindUploaded = true;
      }
      std::memcpy(palette, sm.jointLocalTRS, size_t(dm.jointCount) * 40);
      RTR_TRY(rt_upload(ctx, dm.jointLocalTRS, palette, size_t(dm.jointCount) * 40));
      RTR_TRY(rt_joint_palette(ctx, static_cast<const float *>(dm.jointLocalTRS), static_cast<const int32_t *>(dm.jointParents),
                               static_cast<const float *>(dm.jointInverseBind), dm.jointCount,
                               static_cast<float *>(dm.jointMatrices)));
    } else {
      std::memcpy(palette, sm.jointMatrices, size_t(dm.jointCount) * 64);
      RTR_TRY(rt_upload(ctx, dm.jointMatrices, palette, size_t(dm.jointCount) * 64));
    }
    RTR_TRY(skinMesh(r, dm));
    if (r->flags & RTR_FLAG_REBUILD_SKINNED) {
      uint64_t fresh = 0;
      RTR_TRY(rt_blas_build(ctx, dm.geoms.data(), uint32_t(dm.geoms.size()), RT_AS_FLAG_REFITTABLE, &fresh));
      RTR_TRY(rt_blas_destroy(ctx, dm.blas));
      dm.blas = fresh;
      for (uint32_t i = 0; i < scene->instanceCount; ++i)
        if (r->instanceMesh[i] == m) r->stageDescriptors[i].accelerationStructureID = fresh;
      RTR_TRY(rt_upload(ctx, r->descriptors, r->stageDescriptors, descBytes));
    } else {
      RTR_TRY(rt_blas_refit(ctx, dm.blas, dm.geoms.data(), uint32_t(dm.geoms.size())));
    }
  }
  if (used > r->stageBytes) {
    g_err = "rtr_update: internal staging arena overflow";
    return 2;
  }
  RTR_TRY(rt_fence(ctx, &r->stageFence[slot]));
  ++r->updates;
  // instance AS: refit in place when only transforms / BLAS contents changed — which is all rtr_update allows — as the
  // reference does on devices that support it, otherwise (or on request) rebuild (Renderer.swift:1084-1202)
  if (r->flags & RTR_FLAG_TLAS_REBUILD)
    RTR_TRY(rt_tlas_update(ctx, r->tlas, static_cast<const rt_instance_descriptor *>(r->descriptors), scene->instanceCount));
  else
    RTR_TRY(rt_tlas_refit(ctx, r->tlas, static_cast<const rt_instance_descriptor *>(r->descriptors), scene->instanceCount));
  return 0;
}

int rtr_draw(rtr_renderer *r, const rt_uniforms *uniforms, const rt_trace_options *options) {
  if (!r || !uniforms) {
    g_err = "rtr_draw: null argument";
    return 2;
  }
  if (uniforms->lightCount < 0 || uint32_t(uniforms->lightCount) > r->lightCount) {
    g_err = "rtr_draw: uniforms.lightCount " + std::to_string(uniforms->lightCount) + " exceeds the " +
            std::to_string(r->lightCount) + " lights resident on the device";
    return 2;
  }
  const void *buffers[RT_BUFFER_COUNT] = {};
  buffers[RT_BUFFER_UNIFORMS] = uniforms;
  buffers[RT_BUFFER_RESOURCES] = r->resources;
  buffers[RT_BUFFER_LIGHTS] = r->lights;
  buffers[RT_BUFFER_ACCELERATION_STRUCTURE] = reinterpret_cast<const void *>(static_cast<uintptr_t>(r->tlas));
  buffers[RT_BUFFER_INSTANCE_DESCRIPTORS] = r->descriptors;
  buffers[RT_BUFFER_PREVIOUS_INSTANCE_DESCRIPTORS] = r->prevDescriptors;
  rt_trace_options withHints{};
  if (options) withHints = *options;
  if (r->untextured) withHints.hints |= RT_TRACE_HINT_UNTEXTURED; // known from the materials uploaded at creation
  if (r->noGlass) withHints.hints |= RT_TRACE_HINT_NO_GLASS;
  if (r->flags & RTR_FLAG_ENABLE_AO) withHints.hints |= RT_TRACE_ENABLE_AO;
  RTR_TRY(rt_trace(r->ctx, buffers, r->images, int(sizeof(rt_resource)), int(r->maxSubmeshes), &withHints));
  std::swap(r->images[RT_TEXTURE_ACCUMULATION], r->images[RT_TEXTURE_PREVIOUS_ACCUMULATION]); // Renderer.swift:1492-1494
  return 0;
}

int rtr_read_image(rtr_renderer *r, int textureIndex, void *dstHost, size_t bytes) {
  if (textureIndex < 0 || textureIndex >= RT_TEXTURE_COUNT) {
    g_err = "rtr_read_image: texture index out of range";
    return 2;
  }
  size_t have = size_t(r->width) * r->height * formatBytes(r->images[textureIndex].format);
  if (bytes != have) {
    g_err = "rtr_read_image: size mismatch";
    return 2;
  }
  RTR_TRY(rt_download(r->ctx, dstHost, r->images[textureIndex].data, bytes));
  return 0;
}

int rtr_image_info(rtr_renderer *r, int textureIndex, rt_image *out) {
  if (textureIndex < 0 || textureIndex >= RT_TEXTURE_COUNT) return 2;
  *out = r->images[textureIndex];
  return 0;
}

int rtr_reset_accumulation(rtr_renderer *r) {
  for (int i : {RT_TEXTURE_ACCUMULATION, RT_TEXTURE_PREVIOUS_ACCUMULATION, RT_TEXTURE_MOTION, RT_TEXTURE_DEPTH})
    RTR_TRY(rt_memset(r->ctx, r->images[i].data, 0, size_t(r->width) * r->height * formatBytes(r->images[i].format)));
  return 0;
}

int rtr_read_mesh_streams(rtr_renderer *r, int mesh, float *positions4, float *normals4, float *prevPositions4) {
  if (mesh < 0 || size_t(mesh) >= r->meshes.size()) return 2;
  DeviceMesh &dm = r->meshes[mesh];
  size_t vb = size_t(dm.vertexCount) * 16;
  if (positions4) RTR_TRY(rt_download(r->ctx, positions4, dm.positions, vb));
  if (normals4) RTR_TRY(rt_download(r->ctx, normals4, dm.normals, vb));
  if (prevPositions4) RTR_TRY(rt_download(r->ctx, prevPositions4, dm.prevPositions, vb));
  return 0;
}

int rtr_get_blas_id(rtr_renderer *r, int mesh, uint64_t *id) {
  if (mesh < 0 || size_t(mesh) >= r->meshes.size()) return 2;
  *id = r->meshes[mesh].blas;
  return 0;
}

int rtr_get_tlas_id(rtr_renderer *r, uint64_t *id) {
  *id = r->tlas;
  return 0;
}

int rtr_mesh_count(rtr_renderer *r) { return int(r->meshes.size()); }

} // extern "C"
