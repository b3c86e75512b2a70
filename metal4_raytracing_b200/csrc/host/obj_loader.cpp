// obj_loader.cpp — Wavefront OBJ + MTL reader producing the SoA streams the kernels read.
//
// The reference hands this job to ModelIO (MDLAsset + a 5-stream vertex descriptor, Model.swift:60-81,
// 304-341) whose welding/triangulation rules are not visible in source, so this file *defines* them:
//   * one vertex per distinct (v, vt, vn) index tuple, numbered in first-appearance order;
//   * polygons are fan-triangulated (0, i, i+1);
//   * one submesh per `usemtl` run, in file order (MDLSubmesh per material group);
//   * absent normals are zero-filled (the kernel then falls back to -ray.direction, Raytracing.metal:395-397),
//     absent uvs mean "no uv stream" (Renderer.swift:346-409 binds the normals buffer as a dummy);
//   * MTL -> Material follows SubMesh.swift:291-323: Kd->baseColor, Ke->emission, Ks->specular,
//     Ni->refractionIndex, d->opacity (clamped); Ns is NOT taken (the reference tests .float3 on a scalar).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <tuple>
#include <unordered_map>

#include "scene.h"

namespace rts {

namespace {

struct MtlEntry {
  rt_material m{};
  std::string mapKd, mapBump;
};

std::string dirOf(const std::string &p) {
  size_t k = p.find_last_of('/');
  return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

rt_material defaultMaterial() {
  rt_material m{};
  m.baseColor = {0.8f, 0.8f, 0.8f, 0.0f}; // no MTL record: documented default
  m.refractionIndex = 1.0f;
  m.opacity = 1.0f;
  m.textureFlags = 0;
  return m;
}

void loadMtl(const std::string &path, std::map<std::string, MtlEntry> &out) {
  std::ifstream f(path);
  if (!f) return;
  std::string line, cur;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    std::string key;
    if (!(ss >> key) || key[0] == '#') continue;
    if (key == "newmtl") {
      ss >> cur;
      MtlEntry e;
      e.m = defaultMaterial();
      e.m.baseColor = {0, 0, 0, 0}; // Material() zero-init; Kd fills it
      out[cur] = e;
      continue;
    }
    if (cur.empty()) continue;
    MtlEntry &e = out[cur];
    if (key == "Kd") ss >> e.m.baseColor.x >> e.m.baseColor.y >> e.m.baseColor.z;
    else if (key == "Ke") ss >> e.m.emission.x >> e.m.emission.y >> e.m.emission.z;
    else if (key == "Ks") ss >> e.m.specular.x >> e.m.specular.y >> e.m.specular.z;
    else if (key == "Ni") {
      ss >> e.m.refractionIndex;
    } else if (key == "d") {
      float d = 1.0f;
      ss >> d;
      e.m.opacity = d < 0 ? 0 : (d > 1 ? 1 : d);
    } else if (key == "map_Kd") ss >> e.mapKd;
    else if (key == "map_bump" || key == "bump" || key == "norm") ss >> e.mapBump;
  }
}

struct TupleHash {
  size_t operator()(const std::tuple<int, int, int> &t) const {
    uint64_t a = static_cast<uint32_t>(std::get<0>(t)), b = static_cast<uint32_t>(std::get<1>(t)),
             c = static_cast<uint32_t>(std::get<2>(t));
    uint64_t h = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full ^
                 (c + 0x165667B1ull) * 0xFF51AFD7ED558CCDull;
    return static_cast<size_t>(h ^ (h >> 29));
  }
};

} // namespace

bool decodePng(const std::string &path, Texture &out); // png_decode.cpp

int loadObj(Scene &s, const std::string &path, bool glass, std::string &err) {
  FILE *fp = std::fopen(path.c_str(), "rb");
  if (!fp) {
    err = "cannot open " + path;
    return -1;
  }
  std::vector<float> P, T, N;
  std::map<std::string, MtlEntry> mtl;
  std::unordered_map<std::tuple<int, int, int>, int, TupleHash> weld;
  Mesh mesh;
  size_t slash = path.find_last_of('/');
  mesh.name = slash == std::string::npos ? path : path.substr(slash + 1);
  bool anyUv = false, anyNormal = false;
  Submesh *cur = nullptr;
  std::vector<int> poly;
  char buf[4096];
  auto startSubmesh = [&](const std::string &name) {
    mesh.submeshes.emplace_back();
    cur = &mesh.submeshes.back();
    cur->name = name;
    s.initSubmeshDefaults(*cur);
    auto it = mtl.find(name);
    cur->material = it != mtl.end() ? it->second.m : defaultMaterial();
  };
  while (std::fgets(buf, sizeof buf, fp)) {
    char *p = buf;
    while (*p == ' ' || *p == '\t') ++p;
    if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
      float x = 0, y = 0, z = 0;
      std::sscanf(p + 2, "%f %f %f", &x, &y, &z);
      P.insert(P.end(), {x, y, z});
    } else if (p[0] == 'v' && p[1] == 't') {
      float u = 0, v = 0;
      std::sscanf(p + 3, "%f %f", &u, &v);
      T.insert(T.end(), {u, v});
    } else if (p[0] == 'v' && p[1] == 'n') {
      float x = 0, y = 0, z = 0;
      std::sscanf(p + 3, "%f %f %f", &x, &y, &z);
      N.insert(N.end(), {x, y, z});
    } else if (std::strncmp(p, "mtllib", 6) == 0) {
      char name[1024];
      if (std::sscanf(p + 6, "%1023s", name) == 1) loadMtl(dirOf(path) + "/" + name, mtl);
    } else if (std::strncmp(p, "usemtl", 6) == 0) {
      char name[1024] = "";
      std::sscanf(p + 6, "%1023s", name);
      startSubmesh(name);
    } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
      if (!cur) startSubmesh("");
      poly.clear();
      char *q = p + 1;
      while (*q) {
        while (*q == ' ' || *q == '\t') ++q;
        if (*q == '\0' || *q == '\n' || *q == '\r') break;
        int vi = 0, ti = 0, ni = 0;
        vi = static_cast<int>(std::strtol(q, &q, 10));
        if (*q == '/') {
          ++q;
          if (*q != '/') ti = static_cast<int>(std::strtol(q, &q, 10));
          if (*q == '/') {
            ++q;
            ni = static_cast<int>(std::strtol(q, &q, 10));
          }
        }
        int nP = static_cast<int>(P.size() / 3), nT = static_cast<int>(T.size() / 2),
            nN = static_cast<int>(N.size() / 3);
        vi = vi < 0 ? nP + vi : vi - 1;
        ti = ti < 0 ? nT + ti : ti - 1; // absent -> -1
        ni = ni < 0 ? nN + ni : ni - 1;
        if (vi < 0 || vi >= nP) {
          err = "bad vertex index in " + path;
          std::fclose(fp);
          return -1;
        }
        if (ti >= nT) ti = -1;
        if (ni >= nN) ni = -1;
        auto key = std::make_tuple(vi, ti, ni);
        auto it = weld.find(key);
        int idx;
        if (it == weld.end()) {
          idx = static_cast<int>(mesh.positions.size());
          weld.emplace(key, idx);
          mesh.positions.push_back({P[3 * vi], P[3 * vi + 1], P[3 * vi + 2], 0.0f});
          if (ni >= 0) {
            mesh.normals.push_back({N[3 * ni], N[3 * ni + 1], N[3 * ni + 2], 0.0f});
            anyNormal = true;
          } else {
            mesh.normals.push_back({0, 0, 0, 0});
          }
          if (ti >= 0) {
            mesh.uvs.insert(mesh.uvs.end(), {T[2 * ti], T[2 * ti + 1]});
            anyUv = true;
          } else {
            mesh.uvs.insert(mesh.uvs.end(), {0.0f, 0.0f});
          }
        } else {
          idx = it->second;
        }
        poly.push_back(idx);
      }
      for (size_t i = 1; i + 1 < poly.size(); ++i) {
        cur->indices.push_back(poly[0]);
        cur->indices.push_back(poly[i]);
        cur->indices.push_back(poly[i + 1]);
      }
    }
  }
  std::fclose(fp);
  (void)anyNormal;
  if (!anyUv) mesh.uvs.clear();
  // drop empty submeshes (a usemtl with no faces)
  std::vector<Submesh> kept;
  for (auto &sm : mesh.submeshes)
    if (!sm.indices.empty()) kept.push_back(std::move(sm));
  mesh.submeshes.swap(kept);
  if (mesh.submeshes.empty() || mesh.positions.empty()) {
    err = "no faces in " + path;
    return -1;
  }
  // textures named by the MTL (base colour map forces baseColor = 1, SubMesh.swift:119-125; a normal map
  // is also bound as the opacity map, SubMesh.swift:127-134 — reproduced as written)
  for (auto &sm : mesh.submeshes) {
    auto it = mtl.find(sm.name);
    if (it == mtl.end()) continue;
    if (!it->second.mapKd.empty()) {
      Texture t;
      if (decodePng(dirOf(path) + "/" + it->second.mapKd, t)) {
        t.srgb = true;
        sm.texture[RT_SLOT_BASECOLOR] = s.addTexture(std::move(t));
        sm.material.textureFlags |= RT_MATERIAL_TEXTURE_BASECOLOR;
        sm.material.baseColor = {1, 1, 1, 0};
      }
    }
    if (!it->second.mapBump.empty()) {
      Texture t;
      if (decodePng(dirOf(path) + "/" + it->second.mapBump, t)) {
        t.srgb = false;
        int id = s.addTexture(std::move(t));
        sm.texture[RT_SLOT_NORMAL] = id;
        sm.texture[RT_SLOT_OPACITY] = id;
        sm.material.textureFlags |= RT_MATERIAL_TEXTURE_NORMAL | RT_MATERIAL_TEXTURE_OPACITY;
      }
    }
  }
  if (glass) { // Model.swift:22-26 + SubMesh.swift:275-289
    for (auto &sm : mesh.submeshes) {
      sm.material.baseColor = {0.95f, 0.98f, 1.0f, 0.0f};
      sm.material.refractionIndex = 1.52f;
      sm.material.opacity = 0.08f;
    }
  }
  s.meshes.push_back(std::move(mesh));
  return static_cast<int>(s.meshes.size()) - 1;
}

} // namespace rts
