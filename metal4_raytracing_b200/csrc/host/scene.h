// scene.h — host-side scene model (C++), the data the hot path consumes.
// Mirrors the roles of the reference's Scene / Model / Mesh / Submesh / Skeleton / AnimationClip classes
// (MetalRaytracing/Scene.swift, Model.swift, Mesh.swift, SubMesh.swift) without ModelIO: everything is
// plain std::vector storage laid out exactly as the kernels read it (float3 at stride 16, int32 indices).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/rt_scene.h"
#include "hostmath.h"

namespace rts {

struct Texture {
  int width = 1, height = 1;
  bool srgb = false;
  std::vector<uint8_t> rgba;
};

struct Submesh {
  std::string name;
  std::vector<int32_t> indices;
  rt_material material{};
  int texture[RT_SLOT_COUNT];
};

// Skeleton + one looping clip (Model.swift:346-414). Parents always precede children.
struct Skeleton {
  std::vector<int> parent;
  std::vector<rth::M4> rest;        // local rest transforms
  std::vector<rth::M4> inverseBind; // inverse of the global bind transform
  // clip: per joint rotation about `axis` by amplitude*sin(2*pi*freq*t + phase); translation = rest offset
  std::vector<rth::V3> restOffset;
  std::vector<rth::V3> axis;
  std::vector<float> amplitude, freq, phase;
  double duration = 0.0;
  // keyed clip (rts_set_animation_keys): when keyTimes is non-empty it replaces the procedural one. keyTRS holds
  // keyCount x jointCount x 10 floats (translation xyz, quaternion xyzw, scale xyz); restTRS is the pose without a clip.
  std::vector<float> keyTimes, keyTRS, restTRS;
};

struct Mesh {
  std::string name;
  std::vector<rt_float3> positions, normals;
  std::vector<float> uvs;              // empty when absent
  std::vector<uint16_t> jointIndices;  // 4 per vertex, empty when static
  std::vector<float> jointWeights;     // 4 per vertex
  std::vector<Submesh> submeshes;
  Skeleton skeleton;                   // used when jointIndices is non-empty
  std::vector<float> jointMatrices;    // current palette, 16 floats per joint
  std::vector<float> jointLocalTRS;    // this frame's local transforms: 10 floats per joint (T, quaternion, S)
  std::vector<float> jointInverseBind; // 16 floats per joint (flat copy of skeleton.inverseBind)
  bool skinned() const { return !jointIndices.empty(); }
};

struct Instance {
  int mesh = 0;
  rth::M4 transform = rth::identity();
  rth::M4 previous = rth::identity();
};

struct Scene {
  std::vector<Mesh> meshes;
  std::vector<Texture> textures;
  std::vector<Instance> instances;
  std::vector<rt_light> lights;
  int texWhite = -1, texNeutralNormal = -1, texBlack = -1;

  // flat view storage
  std::vector<rt_scene_mesh> flatMeshes;
  std::vector<std::vector<rt_scene_submesh>> flatSubmeshes;
  std::vector<rt_scene_texture> flatTextures;
  std::vector<rt_scene_instance> flatInstances;

  Scene();
  int addTexture(Texture t);
  void initSubmeshDefaults(Submesh &sm) const;
  int maxSubmeshes() const;
  void animate(double t);
  void animateMesh(Mesh &m, double t);
  void flatten(rt_scene_desc *out);
};

// loaders / generators (obj_loader.cpp, procedural.cpp)
int loadObj(Scene &s, const std::string &path, bool glass, std::string &err);
int addPlane(Scene &s);
int addUvSphere(Scene &s, int rings, int sectors);
int addBumpyIcosphere(Scene &s, int subdiv, int seed);
int addTorusKnot(Scene &s, int nu, int nv, int seed);
int addHumanoid(Scene &s, int vertexBudget, int joints);
Texture makeProceduralTexture(const std::string &kind, int w, int h, int seed, bool srgb);
void computeSmoothNormals(Mesh &m);
uint32_t hash32(uint32_t x, uint32_t seed);
float valueNoise3(float x, float y, float z, uint32_t seed);
float fbm3(float x, float y, float z, int octaves, uint32_t seed);

}  // namespace rts

struct rts_scene {
  rts::Scene s;
};
