// oracle_api.cpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Host-side restatement of what the reference's Renderer does around the two kernels, for a flat
// rt_scene_desc: resource table (Renderer.swift:342-409: row = instance * maxSubmeshes + submesh; skinned
// meshes point ids 0,1,2 at skinned / previous-skinned / skinned-normal buffers), skinning dispatch
// (SkinningPass.swift:160-211), BLAS per mesh with one geometry per submesh (Mesh.swift:84-101), instance
// descriptors (Renderer.swift:547-556, 937-973) and the 16x16-tile dispatch (Renderer.swift:1445-1451).
#include <omp.h>

#include <cstring>
#include <memory>
#include <vector>

#include "oracle.h"
#include "oracle_internal.h"

using namespace orc;

struct OracleMesh {
  std::vector<rt_float3> positions, prevPositions, normals; // current (skinned) streams
  std::vector<rt_material> materials;
  std::unique_ptr<Blas> blas;
  bool skinned = false;
};

struct oracle_ctx {
  int threads = 1;
  std::vector<OracleMesh> meshes;
  std::vector<rt_texture2d> textures;
  std::vector<rt_resource> resources;
  std::vector<rt_instance_descriptor> instances, prevInstances;
  std::vector<rt_light> lights;
  Tlas tlas;
  int maxSubmeshes = 1;
  rt_environment env{};
  bool enableAO = false;
};

static void packDescriptor(const float m[16], uint64_t asId, rt_instance_descriptor &d) {
  std::memset(&d, 0, sizeof d);
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 3; ++r) d.transformationMatrix[c][r] = m[c * 4 + r];
  d.mask = 0xFF;
  d.accelerationStructureID = asId;
}

static void skinMesh(oracle_ctx *c, const rt_scene_mesh &sm, OracleMesh &om) {
  const void *buffers[RT_BUFFER_COUNT] = {};
  buffers[RT_BUFFER_REST_POSITIONS] = sm.positions;
  buffers[RT_BUFFER_REST_NORMALS] = sm.normals;
  buffers[RT_BUFFER_JOINT_INDICES] = sm.jointIndices;
  buffers[RT_BUFFER_JOINT_WEIGHTS] = sm.jointWeights;
  buffers[RT_BUFFER_JOINT_MATRICES] = sm.jointMatrices;
  buffers[RT_BUFFER_SKINNED_POSITIONS] = om.positions.data();
  buffers[RT_BUFFER_SKINNED_NORMALS] = om.normals.data();
  uint32_t n = sm.vertexCount;
#pragma omp parallel for num_threads(c->threads) schedule(static)
  for (int64_t v = 0; v < int64_t(n); ++v) skinningKernelVertex(uint32_t(v), buffers, n);
}

static void buildBlas(const rt_scene_mesh &sm, OracleMesh &om) {
  std::vector<rt_triangle_geometry> geoms(sm.submeshCount);
  for (uint32_t k = 0; k < sm.submeshCount; ++k) {
    geoms[k].vertexBuffer = om.positions.data();
    geoms[k].vertexStride = 16;
    geoms[k].vertexCount = sm.vertexCount;
    geoms[k].indexBuffer = sm.submeshes[k].indices;
    geoms[k].indexStride = 4;
    geoms[k].triangleCount = sm.submeshes[k].triangleCount;
  }
  om.blas.reset(new Blas());
  om.blas->build(geoms.data(), int(geoms.size()));
}

static void refreshInstances(oracle_ctx *c, const rt_scene_desc *scene) {
  c->instances.resize(scene->instanceCount);
  c->prevInstances.resize(scene->instanceCount);
  for (uint32_t i = 0; i < scene->instanceCount; ++i) {
    const rt_scene_instance &in = scene->instances[i];
    uint64_t id = uint64_t(reinterpret_cast<uintptr_t>(c->meshes[in.meshIndex].blas.get()));
    packDescriptor(in.transform, id, c->instances[i]);
    packDescriptor(in.previousTransform, id, c->prevInstances[i]);
  }
  c->tlas.build(c->instances.data(), uint32_t(c->instances.size()));
}

extern "C" {

oracle_ctx *oracle_create(const rt_scene_desc *scene, int threads) {
  oracle_ctx *c = new oracle_ctx();
  // all processors by default, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its children)
  c->threads = threads > 0 ? threads : omp_get_num_procs();
  c->maxSubmeshes = int(scene->maxSubmeshes);
  c->textures.resize(scene->textureCount);
  for (uint32_t i = 0; i < scene->textureCount; ++i)
    c->textures[i] = {scene->textures[i].texels, scene->textures[i].width, scene->textures[i].height,
                      scene->textures[i].srgb, 0};
  c->lights.assign(scene->lights, scene->lights + scene->lightCount);
  c->meshes.resize(scene->meshCount);
  for (uint32_t m = 0; m < scene->meshCount; ++m) {
    const rt_scene_mesh &sm = scene->meshes[m];
    OracleMesh &om = c->meshes[m];
    om.skinned = sm.jointIndices != nullptr && sm.jointCount > 0;
    om.positions.assign(sm.positions, sm.positions + sm.vertexCount);
    om.normals.assign(sm.normals, sm.normals + sm.vertexCount);
    if (om.skinned) skinMesh(c, sm, om);
    om.prevPositions = om.positions;
    om.materials.resize(sm.submeshCount);
    for (uint32_t k = 0; k < sm.submeshCount; ++k) om.materials[k] = sm.submeshes[k].material;
    buildBlas(sm, om);
  }
  c->resources.assign(size_t(scene->instanceCount) * c->maxSubmeshes, rt_resource{});
  for (uint32_t i = 0; i < scene->instanceCount; ++i) {
    uint32_t m = scene->instances[i].meshIndex;
    const rt_scene_mesh &sm = scene->meshes[m];
    OracleMesh &om = c->meshes[m];
    for (uint32_t k = 0; k < sm.submeshCount; ++k) {
      rt_resource &r = c->resources[size_t(i) * c->maxSubmeshes + k];
      r.positions = om.positions.data();
      r.previousPositions = om.prevPositions.data();
      r.normals = om.normals.data();
      r.indices = sm.submeshes[k].indices;
      r.material = &om.materials[k];
      r.uvs = sm.uvs ? sm.uvs : reinterpret_cast<const float *>(om.normals.data());
      const int32_t *ti = sm.submeshes[k].textureIndex;
      r.baseColorMap = &c->textures[ti[RT_SLOT_BASECOLOR]];
      r.normalMap = &c->textures[ti[RT_SLOT_NORMAL]];
      r.roughnessMap = &c->textures[ti[RT_SLOT_ROUGHNESS]];
      r.metallicMap = &c->textures[ti[RT_SLOT_METALLIC]];
      r.aoMap = &c->textures[ti[RT_SLOT_AO]];
      r.opacityMap = &c->textures[ti[RT_SLOT_OPACITY]];
      r.emissionMap = &c->textures[ti[RT_SLOT_EMISSION]];
    }
  }
  refreshInstances(c, scene);
  return c;
}

void oracle_destroy(oracle_ctx *c) { delete c; }

int oracle_update(oracle_ctx *c, const rt_scene_desc *scene) {
  if (scene->meshCount != c->meshes.size()) return -1;
  for (uint32_t m = 0; m < scene->meshCount; ++m) {
    OracleMesh &om = c->meshes[m];
    if (!om.skinned) continue;
    const rt_scene_mesh &sm = scene->meshes[m];
    std::memcpy(om.prevPositions.data(), om.positions.data(), om.positions.size() * sizeof(rt_float3));
    skinMesh(c, sm, om);
    buildBlas(sm, om); // vectors keep their addresses, so resource rows stay valid
  }
  c->lights.assign(scene->lights, scene->lights + scene->lightCount);
  refreshInstances(c, scene);
  return 0;
}

int oracle_render(oracle_ctx *c, const rt_uniforms *uniforms, const rt_image textures[9], uint32_t *primaryIds,
                  uint64_t stats[3], int tileModulo, int tileRemainder) {
  KernelArgs a{};
  a.uniforms = uniforms;
  a.tlas = &c->tlas;
  a.resources = c->resources.data();
  a.instances = c->instances.data();
  a.prevInstances = c->prevInstances.data();
  a.lights = c->lights.data();
  for (int i = 0; i < RT_TEXTURE_COUNT; ++i) a.textures[i] = textures[i];
  a.maxSubmeshes = c->maxSubmeshes;
  a.primaryIds = primaryIds;
  a.env = c->env;
  a.enableAO = c->enableAO;
  if (!textures[RT_TEXTURE_RANDOM].data || !textures[RT_TEXTURE_PREVIOUS_ACCUMULATION].data) return -1;
  if (tileModulo < 1) tileModulo = 1;
  int tilesX = (uniforms->width + 15) / 16, tilesY = (uniforms->height + 15) / 16;
  int tileCount = tilesX * tilesY;
  uint64_t closest = 0, any = 0, hits = 0;
#pragma omp parallel for num_threads(c->threads) schedule(dynamic, 1) reduction(+ : closest, any, hits)
  for (int tile = 0; tile < tileCount; ++tile) {
    if (tile % tileModulo != tileRemainder) continue;
    PixelStats ps;
    int tx = tile % tilesX, ty = tile / tilesX;
    for (int y = ty * 16; y < ty * 16 + 16; ++y)
      for (int x = tx * 16; x < tx * 16 + 16; ++x) raytracingKernelPixel(x, y, a, ps);
    closest += ps.closestRays;
    any += ps.anyRays;
    hits += ps.hits;
  }
  if (stats) {
    stats[0] = closest;
    stats[1] = any;
    stats[2] = hits;
  }
  return 0;
}

int oracle_get_mesh_streams(oracle_ctx *c, int mesh, float *positions4, float *normals4, float *prevPositions4) {
  if (mesh < 0 || size_t(mesh) >= c->meshes.size()) return -1;
  OracleMesh &om = c->meshes[mesh];
  if (positions4) std::memcpy(positions4, om.positions.data(), om.positions.size() * 16);
  if (normals4) std::memcpy(normals4, om.normals.data(), om.normals.size() * 16);
  if (prevPositions4) std::memcpy(prevPositions4, om.prevPositions.data(), om.prevPositions.size() * 16);
  return 0;
}

void oracle_skin(const void *const buffers[18], uint32_t vertexCount) {
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < int64_t(vertexCount); ++v) skinningKernelVertex(uint32_t(v), buffers, vertexCount);
}

float oracle_halton(int i, int d) { return halton(i, d); }

int oracle_intersect_triangle(const float origin[3], const float dir[3], const float v0[3], const float v1[3],
                              const float v2[3], float tmin, float tmax, float out_tuv[3]) {
  RayPrecalc rp = precalcRay(dir);
  float t, u, v;
  if (!intersectTriangle(origin, rp, v0, v1, v2, tmin, tmax, t, u, v)) return 0;
  out_tuv[0] = t;
  out_tuv[1] = u;
  out_tuv[2] = v;
  return 1;
}

void oracle_sample_texture(const rt_texture2d *t, float u, float v, float out[4]) {
  float4 c = sampleTexture(t, {u, v});
  out[0] = c.x, out[1] = c.y, out[2] = c.z, out[3] = c.w;
}

void oracle_invert_affine(const float m4x3[12], float inv[12]) {
  float m[4][3];
  std::memcpy(m, m4x3, sizeof m);
  invertAffine4x3(m, inv);
}

uint16_t oracle_float_to_half(float f) { return floatToHalf(f); }
float oracle_half_to_float(uint16_t h) { return halfToFloat(h); }

int oracle_trace_ray(oracle_ctx *c, const float origin[3], const float dir[3], float tmin, float tmax,
                     uint32_t out_ids[4], float out_tuv[3]) {
  Hit h = traceClosest(c->tlas, {origin[0], origin[1], origin[2]}, {dir[0], dir[1], dir[2]}, tmin, tmax);
  out_ids[0] = h.valid ? 1u : 0u;
  out_ids[1] = h.instance;
  out_ids[2] = h.geometry;
  out_ids[3] = h.primitive;
  out_tuv[0] = h.t;
  out_tuv[1] = h.u;
  out_tuv[2] = h.v;
  return h.valid ? 1 : 0;
}

int oracle_thread_count(oracle_ctx *c) { return c->threads; }

void oracle_sample_environment(const rt_environment *env, const float dir[3], float out_rgb[3]) {
  const float3 c = sampleEnvironment(*env, make3(dir[0], dir[1], dir[2]));
  out_rgb[0] = c.x, out_rgb[1] = c.y, out_rgb[2] = c.z;
}

/* The sampling table of RT_ENV_IMPORTANCE as include/rt_b200.h specifies it (rt_environment_cdf): marginal over rows,
 * then one conditional distribution per row; running sums in double, normalised, stored as float. */
int oracle_environment_cdf(const float *texels, int width, int height, float *out) {
  if (!texels || !out || width <= 0 || height <= 0) return -1;
  const double pi = 3.14159265358979323846;
  float *marginal = out, *rows = out + (height + 1);
  std::vector<double> rowSum(static_cast<size_t>(height)), w(static_cast<size_t>(width));
  double total = 0.0;
  for (int y = 0; y < height; ++y) {
    const double sinTheta = std::sin(pi * (double(y) + 0.5) / double(height));
    double sum = 0.0;
    for (int x = 0; x < width; ++x) {
      const float *t = texels + (size_t(y) * width + x) * 4;
      const double lum = 0.2126 * double(t[0]) + 0.7152 * double(t[1]) + 0.0722 * double(t[2]);
      w[x] = (lum > 0.0 ? lum : 0.0) * sinTheta;
      sum += w[x];
    }
    rowSum[y] = sum;
    total += sum;
    float *row = rows + size_t(y) * (width + 1);
    row[0] = 0.0f;
    double run = 0.0;
    for (int x = 0; x < width; ++x) {
      run += w[x];
      row[x + 1] = sum > 0.0 ? float(run / sum) : float(double(x + 1) / double(width));
    }
    row[width] = 1.0f;
  }
  marginal[0] = 0.0f;
  double run = 0.0;
  for (int y = 0; y < height; ++y) {
    run += rowSum[y];
    marginal[y + 1] = total > 0.0 ? float(run / total) : float(double(y + 1) / double(height));
  }
  marginal[height] = 1.0f;
  return 0;
}

void oracle_set_enable_ao(oracle_ctx *c, int enable) {
  if (c) c->enableAO = enable != 0;
}

int oracle_set_environment(oracle_ctx *c, const rt_environment *env) {
  if (!c) return -1;
  c->env = rt_environment{};
  if (env && env->texelsDev) {
    if (env->width <= 0 || env->height <= 0) return -1;
    c->env = *env;
  }
  return 0;
}

} // extern "C"
