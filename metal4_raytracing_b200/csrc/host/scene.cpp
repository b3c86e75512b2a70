// scene.cpp — Scene container, animation, named benchmark scenes and the rts_* C-ABI (include/rt_scene.h).
#include "scene.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace rts {

bool decodeHdr(const std::string &path, int &width, int &height, std::vector<float> &rgba, std::string &err); // hdr_decode.cpp
bool encodeHdr(const std::string &path, int width, int height, const float *rgba, int channels, std::string &err);

static thread_local std::string g_error;

Scene::Scene() {
  // 1x1 fallbacks exactly as SubMesh.swift:176-241 creates them.
  Texture white;
  white.rgba = {255, 255, 255, 255};
  Texture neutral;
  neutral.rgba = {128, 128, 255, 255}; // 0xFFFF8080 little-endian: (0.5, 0.5, 1.0, 1.0)
  Texture black;
  black.rgba = {0, 0, 0, 255}; // 0x000000FF little-endian is R=255,G=B=A=0 in the reference; the kernel never
                               // samples it (flags are off), so the documented intent (black, opaque) is kept
  texWhite = addTexture(white);
  texNeutralNormal = addTexture(neutral);
  texBlack = addTexture(black);
}

int Scene::addTexture(Texture t) {
  textures.push_back(std::move(t));
  return int(textures.size()) - 1;
}

void Scene::initSubmeshDefaults(Submesh &sm) const {
  sm.texture[RT_SLOT_BASECOLOR] = texWhite;
  sm.texture[RT_SLOT_NORMAL] = texNeutralNormal;
  sm.texture[RT_SLOT_ROUGHNESS] = texWhite;
  sm.texture[RT_SLOT_METALLIC] = texBlack;
  sm.texture[RT_SLOT_AO] = texWhite;
  sm.texture[RT_SLOT_OPACITY] = texWhite;
  sm.texture[RT_SLOT_EMISSION] = texBlack;
  sm.material = rt_material{};
  sm.material.refractionIndex = 1.0f;
  sm.material.opacity = 1.0f;
}

int Scene::maxSubmeshes() const {
  size_t n = 1;
  for (auto &m : meshes) n = std::max(n, m.submeshes.size());
  return int(n);
}

// Model.update (Model.swift:207-261) + SkinningPass.updateSkinningJointMatrices (SkinningPass.swift:123-157)
// at absolute clip time t. geometryBindTransform is identity for the procedural asset.
void Scene::animate(double t) {
  for (auto &m : meshes) animateMesh(m, t);
}

void Scene::animateMesh(Mesh &m, double t) {
  {
    if (!m.skinned()) return;
    const Skeleton &sk = m.skeleton;
    size_t J = sk.parent.size();
    double ct = sk.duration > 0 ? std::fmod(t, sk.duration) : 0.0;
    std::vector<rth::M4> local(J), global(J);
    m.jointLocalTRS.resize(J * 10);
    m.jointInverseBind.resize(J * 16);
    // keyed clip: the two keys around the clip time (AnimationClip.sample, Model.swift:405-411; ModelIO's
    // interpolation is not visible in the reference, here: translation and scale linear, rotation normalised-linear
    // along the shorter arc)
    const bool keyed = !sk.keyTimes.empty() || !sk.restTRS.empty();
    size_t k0 = 0, k1 = 0;
    float blend = 0.0f;
    if (!sk.keyTimes.empty()) {
      const size_t K = sk.keyTimes.size();
      const float ts = sk.keyTimes.front() + float(ct);
      while (k0 + 1 < K && sk.keyTimes[k0 + 1] <= ts) ++k0;
      k1 = k0 + 1 < K ? k0 + 1 : k0;
      const float span = sk.keyTimes[k1] - sk.keyTimes[k0];
      blend = span > 0.0f ? (ts - sk.keyTimes[k0]) / span : 0.0f;
    }
    for (size_t j = 0; j < J; ++j) {
      float qx, qy, qz, qw;
      rth::V3 offset, scl{1, 1, 1};
      if (keyed) {
        const float *a = sk.keyTimes.empty() ? &sk.restTRS[j * 10] : &sk.keyTRS[(k0 * J + j) * 10];
        const float *b = sk.keyTimes.empty() ? a : &sk.keyTRS[(k1 * J + j) * 10];
        auto lerp = [&](int c) { return a[c] + (b[c] - a[c]) * blend; };
        offset = {lerp(0), lerp(1), lerp(2)};
        scl = {lerp(7), lerp(8), lerp(9)};
        const float d = ((a[3] * b[3] + a[4] * b[4]) + a[5] * b[5]) + a[6] * b[6];
        const float sgn = d < 0.0f ? -1.0f : 1.0f; // q and -q are the same rotation: take the shorter arc
        qx = a[3] + (sgn * b[3] - a[3]) * blend, qy = a[4] + (sgn * b[4] - a[4]) * blend;
        qz = a[5] + (sgn * b[5] - a[5]) * blend, qw = a[6] + (sgn * b[6] - a[6]) * blend;
      } else {
        float ang = sk.amplitude[j] * std::sin(float(2.0 * 3.14159265358979323846 * sk.freq[j] * ct) + sk.phase[j]) -
                    sk.amplitude[j] * std::sin(sk.phase[j]); // zero at t = 0 so frame 0 is the bind pose
        rth::V3 ax = rth::normalize(sk.axis[j]);
        float sh = std::sin(ang * 0.5f), ch = std::cos(ang * 0.5f);
        qx = ax.x * sh, qy = ax.y * sh, qz = ax.z * sh, qw = ch;
        offset = sk.restOffset[j];
      }
      { // the same inputs, as handed to the device-side palette evaluation (rt_joint_palette)
        float *t = &m.jointLocalTRS[j * 10];
        t[0] = offset.x, t[1] = offset.y, t[2] = offset.z;
        t[3] = qx, t[4] = qy, t[5] = qz, t[6] = qw;
        t[7] = scl.x, t[8] = scl.y, t[9] = scl.z;
        std::copy(sk.inverseBind[j].m, sk.inverseBind[j].m + 16, &m.jointInverseBind[j * 16]);
      }
      float ql = std::sqrt(((qw * qw + qx * qx) + qy * qy) + qz * qz);
      if (ql > 0.0001f) {
        qx /= ql;
        qy /= ql;
        qz /= ql;
        qw /= ql;
      } else {
        qx = qy = qz = 0;
        qw = 1;
      }
      // matrix4x4_trs (Model.swift:497-506): T * R * S
      local[j] = rth::mul(rth::mul(rth::translate(offset), rth::fromQuat(qx, qy, qz, qw)), rth::scale(scl));
    }
    global = local;
    for (size_t j = 0; j < J; ++j) { // Skeleton.computeGlobalTransforms: parents precede children
      int p = sk.parent[j];
      if (p >= 0 && size_t(p) < j) global[j] = rth::mul(global[p], local[j]);
    }
    m.jointMatrices.resize(J * 16);
    for (size_t j = 0; j < J; ++j) {
      rth::M4 skin = rth::mul(global[j], sk.inverseBind[j]);
      std::copy(skin.m, skin.m + 16, &m.jointMatrices[j * 16]);
    }
  }
}

void Scene::flatten(rt_scene_desc *out) {
  flatMeshes.resize(meshes.size());
  flatSubmeshes.resize(meshes.size());
  for (size_t i = 0; i < meshes.size(); ++i) {
    Mesh &m = meshes[i];
    flatSubmeshes[i].resize(m.submeshes.size());
    for (size_t k = 0; k < m.submeshes.size(); ++k) {
      rt_scene_submesh &fs = flatSubmeshes[i][k];
      std::memset(&fs, 0, sizeof fs);
      fs.indices = m.submeshes[k].indices.data();
      fs.triangleCount = uint32_t(m.submeshes[k].indices.size() / 3);
      fs.material = m.submeshes[k].material;
      for (int q = 0; q < RT_SLOT_COUNT; ++q) fs.textureIndex[q] = m.submeshes[k].texture[q];
    }
    rt_scene_mesh &fm = flatMeshes[i];
    std::memset(&fm, 0, sizeof fm);
    fm.vertexCount = uint32_t(m.positions.size());
    fm.submeshCount = uint32_t(m.submeshes.size());
    fm.positions = m.positions.data();
    fm.normals = m.normals.data();
    fm.uvs = m.uvs.empty() ? nullptr : m.uvs.data();
    fm.jointIndices = m.skinned() ? m.jointIndices.data() : nullptr;
    fm.jointWeights = m.skinned() ? m.jointWeights.data() : nullptr;
    fm.jointCount = m.skinned() ? uint32_t(m.skeleton.parent.size()) : 0;
    fm.jointMatrices = m.skinned() ? m.jointMatrices.data() : nullptr;
    fm.jointParents = m.skinned() ? m.skeleton.parent.data() : nullptr;
    fm.jointInverseBind = m.skinned() && !m.jointInverseBind.empty() ? m.jointInverseBind.data() : nullptr;
    fm.jointLocalTRS = m.skinned() && !m.jointLocalTRS.empty() ? m.jointLocalTRS.data() : nullptr;
    fm.submeshes = flatSubmeshes[i].data();
  }
  flatTextures.resize(textures.size());
  for (size_t i = 0; i < textures.size(); ++i)
    flatTextures[i] = {textures[i].rgba.data(), textures[i].width, textures[i].height, textures[i].srgb ? 1 : 0, 0};
  flatInstances.resize(instances.size());
  for (size_t i = 0; i < instances.size(); ++i) {
    flatInstances[i].meshIndex = uint32_t(instances[i].mesh);
    flatInstances[i]._pad = 0;
    std::copy(instances[i].transform.m, instances[i].transform.m + 16, flatInstances[i].transform);
    std::copy(instances[i].previous.m, instances[i].previous.m + 16, flatInstances[i].previousTransform);
  }
  out->meshCount = uint32_t(meshes.size());
  out->textureCount = uint32_t(textures.size());
  out->instanceCount = uint32_t(instances.size());
  out->lightCount = uint32_t(lights.size());
  out->maxSubmeshes = uint32_t(maxSubmeshes());
  out->_pad = 0;
  out->meshes = flatMeshes.data();
  out->textures = flatTextures.data();
  out->instances = flatInstances.data();
  out->lights = lights.data();
}

static rt_light areaLight(rth::V3 pos, rth::V3 fwd, rth::V3 right, rth::V3 up, rth::V3 color) {
  rt_light l{};
  l.type = RT_LIGHT_AREA;
  l.position = {pos.x, pos.y, pos.z, 0};
  l.forward = {fwd.x, fwd.y, fwd.z, 0};
  l.right = {right.x, right.y, right.z, 0};
  l.up = {up.x, up.y, up.z, 0};
  l.color = {color.x, color.y, color.z, 0};
  return l;
}

static int addInstance(Scene &s, int mesh, rth::V3 p, rth::V3 r, float sc) {
  Instance in;
  in.mesh = mesh;
  in.transform = rth::trs(p, r, sc);
  in.previous = in.transform;
  s.instances.push_back(in);
  return int(s.instances.size()) - 1;
}

} // namespace rts

using namespace rts;

extern "C" {

const char *rts_last_error(void) { return g_error.c_str(); }

rts_scene *rts_scene_create(void) { return new rts_scene(); }
void rts_scene_destroy(rts_scene *s) { delete s; }

int rts_add_mesh_obj(rts_scene *s, const char *objPath, int glass) {
  std::string err;
  int r = loadObj(s->s, objPath, glass != 0, err);
  if (r < 0) g_error = err;
  return r;
}

int rts_add_mesh_procedural(rts_scene *s, const char *kind, int p0, int p1, int p2, int p3) {
  (void)p3;
  std::string k = kind ? kind : "";
  if (k == "plane") return addPlane(s->s);
  if (k == "uvsphere") return addUvSphere(s->s, p0, p1);
  if (k == "icosphere_bumpy") return addBumpyIcosphere(s->s, p0, p1);
  if (k == "torusknot") return addTorusKnot(s->s, p0, p1, p2);
  if (k == "humanoid") return addHumanoid(s->s, p0, p1);
  g_error = "unknown procedural mesh kind: " + k;
  return -1;
}

int rts_add_mesh_raw(rts_scene *s, const float *positions3, const float *normals3, const float *uvs2,
                     uint32_t vertexCount, const int32_t *indices, uint32_t triangleCount) {
  if (!positions3 || !indices || !vertexCount || !triangleCount) {
    g_error = "rts_add_mesh_raw: empty mesh";
    return -1;
  }
  for (uint32_t i = 0; i < triangleCount * 3; ++i)
    if (indices[i] < 0 || uint32_t(indices[i]) >= vertexCount) {
      g_error = "rts_add_mesh_raw: index out of range";
      return -1;
    }
  Mesh m;
  m.name = "raw";
  for (uint32_t i = 0; i < vertexCount; ++i) {
    m.positions.push_back({positions3[3 * i], positions3[3 * i + 1], positions3[3 * i + 2], 0});
    if (normals3)
      m.normals.push_back({normals3[3 * i], normals3[3 * i + 1], normals3[3 * i + 2], 0});
    else
      m.normals.push_back({0, 0, 0, 0});
    if (uvs2) m.uvs.insert(m.uvs.end(), {uvs2[2 * i], uvs2[2 * i + 1]});
  }
  m.submeshes.emplace_back();
  s->s.initSubmeshDefaults(m.submeshes.back());
  m.submeshes.back().material.baseColor = {0.8f, 0.8f, 0.8f, 0};
  m.submeshes.back().indices.assign(indices, indices + size_t(triangleCount) * 3);
  s->s.meshes.push_back(std::move(m));
  return int(s->s.meshes.size()) - 1;
}

// ---- skinned meshes from caller data / from a file (the USD-free path for animated assets) ----
int rts_add_mesh_skinned(rts_scene *s, const float *positions3, const float *normals3, const float *uvs2,
                         const uint16_t *jointIndices4, const float *jointWeights4, uint32_t vertexCount,
                         const int32_t *indices, uint32_t triangleCount, uint32_t jointCount, const int32_t *parents,
                         const float *restTRS, const float *inverseBind) {
  if (!jointIndices4 || !jointWeights4 || !jointCount || !parents || !restTRS || !inverseBind) {
    g_error = "rts_add_mesh_skinned: joint data missing";
    return -1;
  }
  for (uint32_t j = 0; j < jointCount; ++j)
    if (parents[j] >= int32_t(j)) {
      g_error = "rts_add_mesh_skinned: parents must precede children (-1 = root)";
      return -1;
    }
  for (size_t i = 0; i < size_t(vertexCount) * 4; ++i)
    if (jointIndices4[i] >= jointCount) {
      g_error = "rts_add_mesh_skinned: joint index out of range";
      return -1;
    }
  const int mesh = rts_add_mesh_raw(s, positions3, normals3, uvs2, vertexCount, indices, triangleCount);
  if (mesh < 0) return -1;
  Mesh &m = s->s.meshes[size_t(mesh)];
  m.name = "skinned";
  m.jointIndices.assign(jointIndices4, jointIndices4 + size_t(vertexCount) * 4);
  m.jointWeights.assign(jointWeights4, jointWeights4 + size_t(vertexCount) * 4);
  Skeleton &sk = m.skeleton;
  sk.parent.assign(parents, parents + jointCount);
  sk.restTRS.assign(restTRS, restTRS + size_t(jointCount) * 10);
  sk.inverseBind.resize(jointCount);
  for (uint32_t j = 0; j < jointCount; ++j) std::copy(inverseBind + 16 * j, inverseBind + 16 * j + 16, sk.inverseBind[j].m);
  sk.duration = 0.0;
  s->s.animateMesh(m, 0.0); // palette of the rest pose
  return mesh;
}

int rts_set_animation_keys(rts_scene *s, int mesh, uint32_t keyCount, const float *times, const float *trs) {
  if (mesh < 0 || size_t(mesh) >= s->s.meshes.size() || !s->s.meshes[size_t(mesh)].skinned()) {
    g_error = "rts_set_animation_keys: not a skinned mesh";
    return -1;
  }
  Skeleton &sk = s->s.meshes[size_t(mesh)].skeleton;
  if (keyCount == 0) { // back to the rest pose (or the procedural clip of a stand-in)
    sk.keyTimes.clear();
    sk.keyTRS.clear();
    return 0;
  }
  if (!times || !trs) {
    g_error = "rts_set_animation_keys: null arrays";
    return -1;
  }
  for (uint32_t k = 1; k < keyCount; ++k)
    if (!(times[k] > times[k - 1])) {
      g_error = "rts_set_animation_keys: key times must ascend";
      return -1;
    }
  const size_t J = sk.parent.size();
  sk.keyTimes.assign(times, times + keyCount);
  sk.keyTRS.assign(trs, trs + size_t(keyCount) * J * 10);
  if (sk.restTRS.empty()) sk.restTRS.assign(trs, trs + J * 10);
  sk.duration = double(times[keyCount - 1]) - double(times[0]); // AnimationClip.duration, Model.swift:399-401
  return 0;
}

// File layout (little endian): "RTSK1\0\0\0", u32 vertexCount, triangleCount, jointCount, keyCount, hasUvs, 3 x u32 zero;
// positions (3 f32 / vertex), normals (3), [uvs (2)], joint indices (4 u16), joint weights (4 f32), indices (3 i32 /
// triangle), parents (i32 / joint), rest TRS (10 f32 / joint), inverse bind (16 f32 / joint), key times (f32 / key),
// key TRS (10 f32 / joint / key).
int rts_save_skinned_mesh(const rts_scene *s, int mesh, const char *path) {
  if (mesh < 0 || size_t(mesh) >= s->s.meshes.size() || !s->s.meshes[size_t(mesh)].skinned() || !path) {
    g_error = "rts_save_skinned_mesh: not a skinned mesh";
    return -1;
  }
  const Mesh &m = s->s.meshes[size_t(mesh)];
  const Skeleton &sk = m.skeleton;
  if (sk.restTRS.empty()) {
    g_error = "rts_save_skinned_mesh: the mesh has a procedural clip, not keys";
    return -1;
  }
  FILE *f = std::fopen(path, "wb");
  if (!f) {
    g_error = std::string("rts_save_skinned_mesh: cannot write ") + path;
    return -1;
  }
  std::vector<int32_t> indices;
  for (const Submesh &sm : m.submeshes) indices.insert(indices.end(), sm.indices.begin(), sm.indices.end());
  const uint32_t V = uint32_t(m.positions.size()), T = uint32_t(indices.size() / 3), J = uint32_t(sk.parent.size()),
                 K = uint32_t(sk.keyTimes.size());
  const uint32_t head[8] = {V, T, J, K, m.uvs.empty() ? 0u : 1u, 0u, 0u, 0u};
  bool ok = std::fwrite("RTSK1\0\0\0", 1, 8, f) == 8 && std::fwrite(head, 4, 8, f) == 8;
  auto put = [&](const void *p, size_t bytes) { ok = ok && (bytes == 0 || std::fwrite(p, 1, bytes, f) == bytes); };
  std::vector<float> p3(size_t(V) * 3), n3(size_t(V) * 3);
  for (uint32_t i = 0; i < V; ++i) {
    p3[3 * i] = m.positions[i].x, p3[3 * i + 1] = m.positions[i].y, p3[3 * i + 2] = m.positions[i].z;
    n3[3 * i] = m.normals[i].x, n3[3 * i + 1] = m.normals[i].y, n3[3 * i + 2] = m.normals[i].z;
  }
  put(p3.data(), p3.size() * 4);
  put(n3.data(), n3.size() * 4);
  put(m.uvs.data(), m.uvs.size() * 4);
  put(m.jointIndices.data(), m.jointIndices.size() * 2);
  put(m.jointWeights.data(), m.jointWeights.size() * 4);
  put(indices.data(), indices.size() * 4);
  put(sk.parent.data(), sk.parent.size() * 4);
  put(sk.restTRS.data(), sk.restTRS.size() * 4);
  for (uint32_t j = 0; j < J; ++j) put(sk.inverseBind[j].m, 64);
  put(sk.keyTimes.data(), sk.keyTimes.size() * 4);
  put(sk.keyTRS.data(), sk.keyTRS.size() * 4);
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) g_error = std::string("rts_save_skinned_mesh: write failed: ") + path;
  return ok ? 0 : -1;
}

int rts_load_skinned_mesh(rts_scene *s, const char *path) {
  FILE *f = path ? std::fopen(path, "rb") : nullptr;
  if (!f) {
    g_error = std::string("rts_load_skinned_mesh: cannot open ") + (path ? path : "(null)");
    return -1;
  }
  char magic[8];
  uint32_t head[8];
  bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "RTSK1\0\0\0", 8) == 0 && std::fread(head, 4, 8, f) == 8;
  const uint32_t V = ok ? head[0] : 0, T = ok ? head[1] : 0, J = ok ? head[2] : 0, K = ok ? head[3] : 0;
  ok = ok && V > 0 && T > 0 && J > 0 && V <= (1u << 26) && T <= (1u << 26) && J <= 1024 && K <= (1u << 20);
  std::vector<float> p3, n3, uv, w4, rest, bind, times, keys;
  std::vector<uint16_t> j4;
  std::vector<int32_t> indices, parents;
  auto get = [&](auto &v, size_t count) {
    v.resize(count);
    ok = ok && (count == 0 || std::fread(v.data(), sizeof(v[0]), count, f) == count);
  };
  if (ok) {
    get(p3, size_t(V) * 3);
    get(n3, size_t(V) * 3);
    get(uv, head[4] ? size_t(V) * 2 : 0);
    get(j4, size_t(V) * 4);
    get(w4, size_t(V) * 4);
    get(indices, size_t(T) * 3);
    get(parents, J);
    get(rest, size_t(J) * 10);
    get(bind, size_t(J) * 16);
    get(times, K);
    get(keys, size_t(K) * J * 10);
  }
  std::fclose(f);
  if (!ok) {
    g_error = std::string("rts_load_skinned_mesh: not a skinned-mesh file or truncated: ") + path;
    return -1;
  }
  const int mesh = rts_add_mesh_skinned(s, p3.data(), n3.data(), uv.empty() ? nullptr : uv.data(), j4.data(), w4.data(), V,
                                        indices.data(), T, J, parents.data(), rest.data(), bind.data());
  if (mesh < 0) return -1;
  if (K > 0 && rts_set_animation_keys(s, mesh, K, times.data(), keys.data()) != 0) return -1;
  return mesh;
}

int rts_load_hdr(const char *path, int *width, int *height, float **rgbaOut) {
  if (!path || !width || !height || !rgbaOut) {
    g_error = "rts_load_hdr: null argument";
    return -1;
  }
  std::vector<float> texels;
  std::string err;
  if (!rts::decodeHdr(path, *width, *height, texels, err)) {
    g_error = "rts_load_hdr: " + err;
    return -1;
  }
  *rgbaOut = static_cast<float *>(std::malloc(texels.size() * sizeof(float)));
  if (!*rgbaOut) {
    g_error = "rts_load_hdr: out of memory";
    return -1;
  }
  std::memcpy(*rgbaOut, texels.data(), texels.size() * sizeof(float));
  return 0;
}

void rts_free(void *p) { std::free(p); }

int rts_write_hdr(const char *path, const float *rgba32f, int width, int height) {
  std::string err;
  if (!path || !rts::encodeHdr(path, width, height, rgba32f, 4, err)) {
    g_error = "rts_write_hdr: " + (path ? err : std::string("null path"));
    return -1;
  }
  return 0;
}

static Submesh *findSubmesh(rts_scene *s, int mesh, int submesh) {
  if (mesh < 0 || size_t(mesh) >= s->s.meshes.size() || submesh < 0 ||
      size_t(submesh) >= s->s.meshes[mesh].submeshes.size()) {
    g_error = "mesh/submesh index out of range";
    return nullptr;
  }
  return &s->s.meshes[mesh].submeshes[submesh];
}

int rts_set_material(rts_scene *s, int mesh, int submesh, const rt_material *m) {
  Submesh *sm = findSubmesh(s, mesh, submesh);
  if (!sm) return -1;
  sm->material = *m;
  return 0;
}

int rts_get_material(const rts_scene *s, int mesh, int submesh, rt_material *m) {
  Submesh *sm = findSubmesh(const_cast<rts_scene *>(s), mesh, submesh);
  if (!sm) return -1;
  *m = sm->material;
  return 0;
}

int rts_add_texture_rgba8(rts_scene *s, const uint8_t *texels, int width, int height, int srgb) {
  if (!texels || width <= 0 || height <= 0) {
    g_error = "rts_add_texture_rgba8: bad arguments";
    return -1;
  }
  Texture t;
  t.width = width;
  t.height = height;
  t.srgb = srgb != 0;
  t.rgba.assign(texels, texels + size_t(width) * height * 4);
  return s->s.addTexture(std::move(t));
}

int rts_add_texture_procedural(rts_scene *s, const char *kind, int width, int height, int seed, int srgb) {
  return s->s.addTexture(makeProceduralTexture(kind ? kind : "", width, height, seed, srgb != 0));
}

int rts_bind_texture(rts_scene *s, int mesh, int submesh, int slot, int texture) {
  Submesh *sm = findSubmesh(s, mesh, submesh);
  if (!sm) return -1;
  if (slot < 0 || slot >= RT_SLOT_COUNT || texture < 0 || size_t(texture) >= s->s.textures.size()) {
    g_error = "rts_bind_texture: slot/texture out of range";
    return -1;
  }
  static const uint32_t flag[RT_SLOT_COUNT] = {RT_MATERIAL_TEXTURE_BASECOLOR, RT_MATERIAL_TEXTURE_NORMAL,
                                               RT_MATERIAL_TEXTURE_ROUGHNESS, RT_MATERIAL_TEXTURE_METALLIC,
                                               RT_MATERIAL_TEXTURE_AO,        RT_MATERIAL_TEXTURE_OPACITY,
                                               RT_MATERIAL_TEXTURE_EMISSION};
  sm->texture[slot] = texture;
  sm->material.textureFlags |= flag[slot];
  if (slot == RT_SLOT_BASECOLOR) sm->material.baseColor = {1, 1, 1, 0};
  return 0;
}

int rts_add_instance(rts_scene *s, int mesh, const float position[3], const float rotation[3], float scale) {
  if (mesh < 0 || size_t(mesh) >= s->s.meshes.size()) {
    g_error = "rts_add_instance: mesh out of range";
    return -1;
  }
  return addInstance(s->s, mesh, {position[0], position[1], position[2]}, {rotation[0], rotation[1], rotation[2]}, scale);
}

int rts_set_instance_transform(rts_scene *s, int instance, const float position[3], const float rotation[3],
                               float scale) {
  if (instance < 0 || size_t(instance) >= s->s.instances.size()) {
    g_error = "rts_set_instance_transform: instance out of range";
    return -1;
  }
  Instance &in = s->s.instances[instance];
  in.previous = in.transform;
  in.transform = rth::trs({position[0], position[1], position[2]}, {rotation[0], rotation[1], rotation[2]}, scale);
  return 0;
}

int rts_add_light(rts_scene *s, const rt_light *l) {
  s->s.lights.push_back(*l);
  return int(s->s.lights.size()) - 1;
}

void rts_clear_lights(rts_scene *s) { s->s.lights.clear(); }

void rts_default_lights(rts_scene *s) { // Scene.swift:82-93,161-169
  s->s.lights.clear();
  s->s.lights.push_back(areaLight({0, 1.98f, 0}, {0, -1, 0}, {0.25f, 0, 0}, {0, 0, 0.25f}, {4, 4, 4}));
  rt_light spot{};
  spot.type = RT_LIGHT_SPOT;
  spot.position = {2, 1, 4, 0};
  spot.direction = {-1.5f, -0.5f, -1.5f, 0};
  spot.coneAngle = 25.0f / 180.0f * 3.14159265358979323846f;
  spot.color = {4, 4, 4, 0};
  s->s.lights.push_back(spot);
}

void rts_make_orbit_camera(float width, float height, const float target[3], float azimuth, float elevation,
                           float distance, float fovDegrees, rt_camera *out) {
  // Scene.makeOrbitCamera, Scene.swift:126-159
  const float pi = 3.14159265358979323846f;
  float safeDistance = std::max(0.001f, distance);
  float limit = (pi / 2.0f) - 0.001f;
  float el = std::max(-limit, std::min(limit, elevation));
  float x = safeDistance * std::cos(el) * std::sin(azimuth);
  float y = safeDistance * std::sin(el);
  float z = safeDistance * std::cos(el) * std::cos(azimuth);
  rth::V3 tgt{target[0], target[1], target[2]};
  rth::V3 position = tgt + rth::V3{x, y, z};
  rth::V3 forward = rth::normalize(tgt - position);
  rth::V3 right = rth::normalize(rth::cross(forward, {0, 1, 0}));
  if (rth::length(right) < 0.0001f) right = {1, 0, 0};
  rth::V3 up = rth::normalize(rth::cross(right, forward));
  float fieldOfView = fovDegrees * (pi / 180.0f);
  float aspect = width / height;
  float imagePlaneHeight = std::tan(fieldOfView / 2.0f);
  float imagePlaneWidth = aspect * imagePlaneHeight;
  std::memset(out, 0, sizeof *out);
  out->position = {position.x, position.y, position.z, 0};
  out->right = {right.x * imagePlaneWidth, right.y * imagePlaneWidth, right.z * imagePlaneWidth, 0};
  out->up = {up.x * imagePlaneHeight, up.y * imagePlaneHeight, up.z * imagePlaneHeight, 0};
  out->forward = {forward.x, forward.y, forward.z, 0};
}

void rts_default_camera(float width, float height, rt_camera *out) {
  // Scene.setupCamera, Scene.swift:111-123
  const float target[3] = {0, 0, 0};
  rth::V3 offset{0.0f, 1.0f, 5.38f};
  float distance = std::max(0.001f, rth::length(offset));
  float azimuth = std::atan2(offset.x, offset.z);
  float elevation = std::asin(offset.y / distance);
  rts_make_orbit_camera(width, height, target, azimuth, elevation, distance, 45.0f, out);
}

int rts_animate(rts_scene *s, double timeSeconds) {
  s->s.animate(timeSeconds);
  return 0;
}

void rts_default_uniforms(int width, int height, rt_uniforms *u) {
  std::memset(u, 0, sizeof *u);
  u->width = width;
  u->height = height;
  u->blocksWide = (width + 15) / 16;
  u->frameIndex = 0;
  u->lightCount = 0;
  u->samplesPerPixel = 2;
  u->maxBounces = 2;
  rts_default_camera(float(width), float(height), &u->camera);
  u->previousCamera = u->camera;
  u->debugTextureMode = 0;
  u->accumulationWeight = 0.9f;
  u->enableDenoiseGBuffer = 0;
  u->shadingMode = RT_SHADING_PBR;
  u->enableMotionAdaptiveAccumulation = 1;
  u->motionAccumulationMinWeight = 0.1f;
  u->motionAccumulationLowThresholdPixels = 0.5f;
  u->motionAccumulationHighThresholdPixels = 4.0f;
  u->enableMotionAdaptiveSampling = 1;
  u->motionSamplingMaxExtraSamples = 2;
  u->motionSamplingLowThresholdPixels = 1.0f;
  u->motionSamplingHighThresholdPixels = 6.0f;
}

void rts_fill_seed_image(uint32_t *dst, int width, int height, uint32_t seed) {
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) dst[size_t(y) * width + x] = hash32(uint32_t(y * width + x), seed) & 0xFFFFFu;
}

int rts_scene_get_desc(rts_scene *s, rt_scene_desc *out) {
  s->s.flatten(out);
  return 0;
}

rts_scene *rts_scene_create_named(const char *name, const char *assetDir, int width, int height,
                                  rt_uniforms *uniformsOut, uint32_t *seedOut) {
  std::string n = name ? name : "";
  std::string dir = assetDir ? assetDir : "";
  rts_scene *sc = new rts_scene();
  Scene &s = sc->s;
  rt_uniforms u;
  rts_default_uniforms(width, height, &u);
  // throughput configs run with the motion-adaptive paths off; K5 turns accumulation on
  u.enableMotionAdaptiveSampling = 0;
  u.enableMotionAdaptiveAccumulation = 0;
  uint32_t seed = 0;
  std::string err;
  auto obj = [&](const char *file, bool glass) -> int {
    if (dir.empty()) {
      err = std::string("scene ") + n + " needs asset " + file + " but no asset directory was given";
      return -1;
    }
    return loadObj(s, dir + "/" + file, glass, err);
  };
  bool ok = true;
  if (n == "K1") {
    int plane = obj("plane.obj", false), sphere = plane >= 0 ? obj("sphere.obj", false) : -1;
    ok = plane >= 0 && sphere >= 0;
    if (ok) {
      addInstance(s, plane, {0, 0, 0}, {0, 0, 0}, 10.0f);
      addInstance(s, sphere, {-1.9f, 0.0f, 0.3f}, {0, 0, 0}, 1.0f);
      rt_light pl{};
      pl.type = RT_LIGHT_POINT;
      pl.position = {1, 1, 1, 0};
      pl.color = {4, 4, 4, 0};
      s.lights.push_back(pl);
      u.samplesPerPixel = 1;
      u.maxBounces = 1;
      u.accumulationWeight = 0.0f;
      seed = 0xC0FFEEu;
    }
  } else if (n == "K2" || n == "K2tex") {
    int bunny = addBumpyIcosphere(s, 6, 2);
    int plane = dir.empty() ? addPlane(s) : obj("plane.obj", false);
    ok = plane >= 0;
    if (ok) {
      addInstance(s, bunny, {0, 0.5f, 0}, {0, 0, 0}, 0.5f);
      addInstance(s, plane, {0, 0, 0}, {0, 0, 0}, 10.0f);
      rts_default_lights(sc);
      u.samplesPerPixel = 4;
      u.maxBounces = 2;
      seed = 1;
      if (n == "K2tex") {
        int base = s.addTexture(makeProceduralTexture("uvgrid", 1024, 1024, 3, true));
        int rough = s.addTexture(makeProceduralTexture("valuenoise", 1024, 1024, 7, false));
        int metal = s.addTexture(makeProceduralTexture("checker", 512, 512, 5, false));
        int bump = s.addTexture(makeProceduralTexture("bump", 1024, 1024, 9, false));
        rts_bind_texture(sc, bunny, 0, RT_SLOT_BASECOLOR, base);
        rts_bind_texture(sc, bunny, 0, RT_SLOT_ROUGHNESS, rough);
        rts_bind_texture(sc, bunny, 0, RT_SLOT_METALLIC, metal);
        rts_bind_texture(sc, bunny, 0, RT_SLOT_NORMAL, bump);
      }
    }
  } else if (n == "K3" || n == "K3glass" || n == "K3small") {
    // K3small: same layout with a 132 x 33 knot (8,712 tris) for CPU-sized tests
    int dragon = n == "K3small" ? addTorusKnot(s, 132, 33, 3) : addTorusKnot(s, 1320, 330, 3);
    int plane = dir.empty() ? addPlane(s) : obj("plane.obj", false);
    int back = -1;
    if (plane >= 0 && !dir.empty()) back = obj("plane-back.obj", false);
    ok = plane >= 0 && (dir.empty() || back >= 0);
    if (ok) {
      if (n == "K3glass")
        for (auto &sm : s.meshes[dragon].submeshes) {
          sm.material.baseColor = {0.95f, 0.98f, 1.0f, 0};
          sm.material.refractionIndex = 1.52f;
          sm.material.opacity = 0.08f;
        }
      addInstance(s, dragon, {0.3f, 0.38f, 2.5f}, {0, 3.14159265358979323846f / 2 * 1.2f, 0}, 1.2f);
      addInstance(s, plane, {0, 0, 0}, {0, 0, 0}, 10.0f);
      if (back >= 0) addInstance(s, back, {0, 0, -1.5f}, {0, 0, 0}, 10.0f);
      rts_default_lights(sc);
      u.samplesPerPixel = 16;
      u.maxBounces = 3;
      seed = 3;
    }
  } else if (n == "K4" || n == "K4small") {
    int tree = obj("treefir.obj", false), train = tree >= 0 ? obj("train.obj", false) : -1,
        teapot = train >= 0 ? obj("teapot.obj", false) : -1, plane = teapot >= 0 ? obj("plane.obj", false) : -1;
    ok = plane >= 0;
    if (ok) {
      int grid = n == "K4small" ? 8 : 64;
      const float spacing = 1.25f;
      for (int gz = 0; gz < grid; ++gz)
        for (int gx = 0; gx < grid; ++gx) {
          uint32_t h = hash32(uint32_t(gz * grid + gx), 11);
          int which = int(h % 3u);
          float jx = (float((h >> 4) & 0xFF) / 255.0f - 0.5f) * 0.5f;
          float jz = (float((h >> 12) & 0xFF) / 255.0f - 0.5f) * 0.5f;
          float yaw = float((h >> 20) & 0xFF) / 255.0f * 6.2831853f;
          float sc2 = 0.8f + 0.4f * float((h >> 28) & 0xF) / 15.0f;
          rth::V3 p{(float(gx) - 0.5f * float(grid - 1)) * spacing + jx, 0.0f,
                    (float(gz) - 0.5f * float(grid - 1)) * spacing + jz};
          if (which == 0)
            addInstance(s, tree, p, {0, yaw, 0}, 0.7f * sc2);
          else if (which == 1)
            addInstance(s, train, p, {0, yaw, 0}, 0.5f * sc2);
          else
            addInstance(s, teapot, p, {0, yaw, 0}, 0.006f * sc2); // teapot.obj is ~100 units wide
        }
      addInstance(s, plane, {0, 0, 0}, {0, 0, 0}, float(grid) * spacing);
      rt_light sun{};
      sun.type = RT_LIGHT_SUN;
      sun.direction = {-1, -2, 0, 0};
      sun.color = {1, 1, 1, 0};
      s.lights.push_back(sun);
      float half = 0.5f * float(grid) * spacing;
      s.lights.push_back(areaLight({0, 6.0f, 0}, {0, -1, 0}, {half * 0.25f, 0, 0}, {0, 0, half * 0.25f}, {40, 40, 40}));
      const float target[3] = {0, 0, 0};
      rts_make_orbit_camera(float(width), float(height), target, 0.6f, 0.55f, float(grid) * spacing * 0.9f, 45.0f,
                            &u.camera);
      u.previousCamera = u.camera;
      u.samplesPerPixel = 8;
      u.maxBounces = 2;
      seed = 4;
    }
  } else if (n == "K5" || n == "K5small") {
    int robot = n == "K5small" ? addHumanoid(s, 4000, 64) : addHumanoid(s, 100000, 64);
    int plane = dir.empty() ? addPlane(s) : obj("plane.obj", false);
    ok = plane >= 0;
    if (ok) {
      addInstance(s, robot, {-0.5f, 0.0f, 1.0f}, {0, 0, 0}, 0.01f);
      addInstance(s, plane, {0, 0, 0}, {0, 0, 0}, 10.0f);
      rts_default_lights(sc);
      u.samplesPerPixel = 2;
      u.maxBounces = 2;
      u.accumulationWeight = 0.9f;
      u.enableMotionAdaptiveAccumulation = 1;
      seed = 5;
    }
  } else if (n == "appscene") {
    // AppScene.swift:14-28 with stand-ins for robot and dragon; defaults of Renderer.swift:117-192
    int robot = addHumanoid(s, 100000, 64);
    int dragon = addTorusKnot(s, 1320, 330, 3);
    for (auto &sm : s.meshes[dragon].submeshes) {
      sm.material.baseColor = {0.95f, 0.98f, 1.0f, 0};
      sm.material.refractionIndex = 1.52f;
      sm.material.opacity = 0.08f;
    }
    int train = obj("train.obj", false), tree = train >= 0 ? obj("treefir.obj", false) : -1,
        plane = tree >= 0 ? obj("plane.obj", false) : -1, sph1 = plane >= 0 ? obj("sphere.obj", false) : -1,
        sph2 = sph1 >= 0 ? obj("sphere.obj", false) : -1, back = sph2 >= 0 ? obj("plane-back.obj", false) : -1;
    ok = back >= 0;
    if (ok) {
      addInstance(s, robot, {-0.5f, 0, 1.0f}, {0, 0, 0}, 0.01f);
      addInstance(s, dragon, {0.3f, 0.38f, 2.5f}, {0, 3.14159265358979323846f / 2 * 1.2f, 0}, 1.2f);
      addInstance(s, train, {-0.3f, 0, 0.4f}, {0, 0, 0}, 0.5f);
      addInstance(s, tree, {0.5f, 0, -0.2f}, {0, 0, 0}, 0.7f);
      addInstance(s, plane, {0, 0, 0}, {0, 0, 0}, 10.0f);
      addInstance(s, sph1, {-1.9f, 0, 0.3f}, {0, 0, 0}, 1.0f);
      addInstance(s, sph2, {2.9f, 0, -0.5f}, {0, 0, 0}, 2.0f);
      addInstance(s, back, {0, 0, -1.5f}, {0, 0, 0}, 10.0f);
      rts_default_lights(sc);
      rts_default_uniforms(width, height, &u); // the app's own defaults, adaptive paths ON
      seed = 6;
    }
  } else {
    err = "unknown scene name: " + n;
    ok = false;
  }
  if (!ok) {
    g_error = err.empty() ? std::string("failed to build scene ") + n : err;
    delete sc;
    return nullptr;
  }
  u.lightCount = int32_t(s.lights.size());
  if (uniformsOut) *uniformsOut = u;
  if (seedOut) *seedOut = seed;
  return sc;
}

} // extern "C"
