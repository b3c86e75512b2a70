/*
 * rt_scene.h — C-ABI of the host-side scene library (librt_scene.so, pure CPU, no CUDA).
 *
 * It produces every input of the hot path in host memory, the way the reference's Scene / AppScene / Model /
 * Mesh / Submesh classes do through ModelIO (MetalRaytracing/Scene.swift:73-169, AppScene.swift:11-28,
 * Model.swift:45-261, Mesh.swift:41-68, SubMesh.swift:38-323): SoA vertex streams (float3 at stride 16),
 * 32-bit indices per submesh, one Material per submesh, RGBA8 textures, instance transforms
 * (T * Rx*Ry*Rz * S, Utilities.swift:302-355), lights, the orbit camera, skeleton + animation and the
 * per-mesh joint palette (SkinningPass.swift:123-157), plus the deterministic stand-ins for the assets the
 * reference mount does not contain.
 *
 * The flat `rt_scene_desc` view is what both the GPU host renderer (include/rt_b200.h, rtr_* calls) and the
 * CPU oracle (oracle/) consume, so the two sides see byte-identical inputs.
 */
#ifndef RT_SCENE_H
#define RT_SCENE_H

#include "rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* order of the 7 texture slots of a submesh == Resource ids 6..12 (Raytracing.metal:176-182) */
enum rt_texture_slot {
  RT_SLOT_BASECOLOR = 0,
  RT_SLOT_NORMAL = 1,
  RT_SLOT_ROUGHNESS = 2,
  RT_SLOT_METALLIC = 3,
  RT_SLOT_AO = 4,
  RT_SLOT_OPACITY = 5,
  RT_SLOT_EMISSION = 6,
  RT_SLOT_COUNT = 7
};

typedef struct rt_scene_submesh {
  const int32_t *indices; /* 3 * triangleCount, indexes the mesh's vertex streams */
  uint32_t triangleCount;
  uint32_t _pad;
  rt_material material;
  int32_t textureIndex[RT_SLOT_COUNT]; /* index into rt_scene_desc.textures; always valid (1x1 fallbacks) */
  int32_t _pad2;
} rt_scene_submesh;

typedef struct rt_scene_mesh {
  uint32_t vertexCount;
  uint32_t submeshCount;
  const rt_float3 *positions;   /* rest pose, stride 16 */
  const rt_float3 *normals;     /* rest pose, stride 16; all-zero when the asset has none */
  const float *uvs;             /* float2 stride 8, or NULL when the asset has none */
  const uint16_t *jointIndices; /* ushort4 per vertex, or NULL for a static mesh */
  const float *jointWeights;    /* float4 per vertex, or NULL */
  uint32_t jointCount;          /* 0 for a static mesh */
  uint32_t _pad;
  const float *jointMatrices;   /* jointCount x float4x4 column-major: current palette (A22) */
  const rt_scene_submesh *submeshes;
  /* skeleton inputs of that palette, for evaluating it on the device (rt_joint_palette); NULL for a static mesh */
  const int32_t *jointParents;     /* jointCount; -1 = root; parents precede children */
  const float *jointInverseBind;   /* jointCount x float4x4 column-major */
  const float *jointLocalTRS;      /* jointCount x 10: translation xyz, quaternion xyzw, scale xyz (this frame) */
} rt_scene_mesh;

typedef struct rt_scene_texture {
  const uint8_t *texels; /* RGBA8 */
  int32_t width, height, srgb, _pad;
} rt_scene_texture;

typedef struct rt_scene_instance {
  uint32_t meshIndex;
  uint32_t _pad;
  float transform[16];         /* column-major 4x4, object -> world, this frame */
  float previousTransform[16]; /* last frame (motion vectors) */
} rt_scene_instance;

typedef struct rt_scene_desc {
  uint32_t meshCount, textureCount, instanceCount, lightCount;
  uint32_t maxSubmeshes; /* function constant 1 of the reference kernel (Renderer.swift:346-361) */
  uint32_t _pad;
  const rt_scene_mesh *meshes;
  const rt_scene_texture *textures;
  const rt_scene_instance *instances;
  const rt_light *lights;
} rt_scene_desc;

typedef struct rts_scene rts_scene;

/* All functions returning int return 0 on success; rts_last_error() describes the last failure. */
const char *rts_last_error(void);

rts_scene *rts_scene_create(void);
void rts_scene_destroy(rts_scene *s);

/* --- meshes ------------------------------------------------------------------------------------------ */
/* OBJ + MTL loader. Rules (the reference delegates these to ModelIO, so they are this library's contract):
 * one vertex per distinct v/vt/vn tuple in first-appearance order; polygons fan-triangulated (0,i,i+1);
 * one submesh per `usemtl` run in file order; absent normals / uvs are zero-filled / omitted.
 * glass != 0 applies Model.swift:22-26's override (tint .95,.98,1; ior 1.52; opacity 0.08).
 * Returns the mesh index or -1. */
int rts_add_mesh_obj(rts_scene *s, const char *objPath, int glass);
/* Procedural stand-ins. kind: "plane" (2 tris, uv+n), "uvsphere" (rings, sectors), "icosphere_bumpy"
 * (subdiv, seed: bunny stand-in), "torusknot" (nu, nv, seed: dragon stand-in), "humanoid" (vertex budget,
 * joints: skinned robot stand-in, carries a skeleton + animation). p0..p3 are kind-specific. */
int rts_add_mesh_procedural(rts_scene *s, const char *kind, int p0, int p1, int p2, int p3);
/* Raw triangle soup entry point (tests). normals/uvs may be NULL. One submesh. */
int rts_add_mesh_raw(rts_scene *s, const float *positions3, const float *normals3, const float *uvs2,
                     uint32_t vertexCount, const int32_t *indices, uint32_t triangleCount);
/* Skinned mesh from caller data — the path for animated assets without USD tooling (the reference
 * gets the same streams from ModelIO, Model.swift:135-192, 304-341). One submesh. jointIndices4: ushort4 per vertex,
 * jointWeights4: float4 per vertex, used as authored (Skinning.metal:26-31). parents: -1 = root, parents precede
 * children (Skeleton.computeGlobalTransforms, Model.swift:379-387). restTRS: jointCount x 10 floats (translation xyz,
 * rotation quaternion xyzw, scale xyz), the pose without a clip; inverseBind: jointCount x float4x4 column-major.
 * Returns the mesh index or -1. */
int rts_add_mesh_skinned(rts_scene *s, const float *positions3, const float *normals3, const float *uvs2,
                         const uint16_t *jointIndices4, const float *jointWeights4, uint32_t vertexCount,
                         const int32_t *indices, uint32_t triangleCount, uint32_t jointCount, const int32_t *parents,
                         const float *restTRS, const float *inverseBind);
/* Keyed clip of a skinned mesh: keyCount ascending times (seconds) and keyCount x jointCount x 10 floats laid out as
 * restTRS. rts_animate(t) samples it at times[0] + fmod(t, times[last] - times[0]) (Model.update, Model.swift:207-215):
 * translation and scale linearly, rotation by normalised linear interpolation along the shorter arc (ModelIO's own
 * interpolation is not visible in the reference); then the quaternion is renormalised and local = T * R * S,
 * global = parent * local, palette = global * inverseBind as in Model.swift:226-261. keyCount 0 removes the clip. */
int rts_set_animation_keys(rts_scene *s, int mesh, uint32_t keyCount, const float *times, const float *trs);
/* The same data as one little-endian file ("RTSK1", layout in csrc/host/scene.cpp). load returns the mesh index. */
int rts_save_skinned_mesh(const rts_scene *s, int mesh, const char *path);
int rts_load_skinned_mesh(rts_scene *s, const char *path);
int rts_set_material(rts_scene *s, int mesh, int submesh, const rt_material *m);
int rts_get_material(const rts_scene *s, int mesh, int submesh, rt_material *m);
/* RGBA8 texture; returns texture index. */
int rts_add_texture_rgba8(rts_scene *s, const uint8_t *texels, int width, int height, int srgb);
/* kind: "checker", "valuenoise", "bump" (normal map from value noise), "uvgrid". */
int rts_add_texture_procedural(rts_scene *s, const char *kind, int width, int height, int seed, int srgb);
/* bind texture to a slot and set the matching MATERIAL_TEXTURE_* flag (SubMesh.swift:119-131 semantics:
 * a base-colour map forces baseColor = 1). */
int rts_bind_texture(rts_scene *s, int mesh, int submesh, int slot, int texture);

/* --- instances, lights, camera ------------------------------------------------------------------------ */
int rts_add_instance(rts_scene *s, int mesh, const float position[3], const float rotation[3], float scale);
int rts_set_instance_transform(rts_scene *s, int instance, const float position[3], const float rotation[3],
                               float scale); /* keeps the old transform as previousTransform */
int rts_add_light(rts_scene *s, const rt_light *l);
void rts_clear_lights(rts_scene *s);
void rts_default_lights(rts_scene *s); /* Scene.swift:82-93: area light + spot light */
/* Scene.swift:126-159 orbit camera. */
void rts_make_orbit_camera(float width, float height, const float target[3], float azimuth, float elevation,
                           float distance, float fovDegrees, rt_camera *out);
void rts_default_camera(float width, float height, rt_camera *out); /* (0,1,5.38) -> origin, 45 deg */

/* --- animation ----------------------------------------------------------------------------------------- */
/* Model.update + SkinningPass.updateSkinningJointMatrices at absolute time t (seconds): samples the clip,
 * local TRS -> global -> x inverseBind -> per-mesh palette. */
int rts_animate(rts_scene *s, double timeSeconds);

/* --- named benchmark scenes (SURVEY.md §8d / BASELINE.md §3): "K1".."K5", "K2tex", "K3glass", "appscene".
 * assetDir may be NULL when the config needs no OBJ files. Fills defaults for uniforms (without camera
 * aspect surprises: width/height are taken from the arguments). */
rts_scene *rts_scene_create_named(const char *name, const char *assetDir, int width, int height,
                                  rt_uniforms *uniformsOut, uint32_t *seedOut);

/* --- views ---------------------------------------------------------------------------------------------- */
/* Pointers stay valid until the scene is mutated or destroyed. */
int rts_scene_get_desc(rts_scene *s, rt_scene_desc *out);
/* seed[y*W+x] = hash32(y*W+x, seed) & 0xFFFFF — range of Renderer.swift:719-726 (arc4random % 2^20). */
/* Image writer for the post chain: 8-bit RGBA PNG, rows top to bottom. */
int rts_write_png(const char *path, const uint8_t *rgba8, int width, int height);
/* Radiance RGBE (.hdr) reader for the environment extension (rt_b200.h rt_environment; BASELINE's K3 names
 * vulture_hide_4k.hdr, which the reference ships and never loads): "#?RADIANCE" / "#?RGBE" header,
 * FORMAT=32-bit_rle_rgbe, "-Y h +X w", flat or run-length-encoded scanlines; (r, g, b, e) -> (r, g, b) * 2^(e - 136).
 * *rgbaOut receives width * height RGBA32F texels (alpha 1, row 0 = top = straight up), to be released with rts_free. */
int rts_load_hdr(const char *path, int *width, int *height, float **rgbaOut);
void rts_free(void *p);
/* The HDR end of the image writers: width * height RGBA32F texels (e.g. a downloaded fp32
 * accumulation image; alpha ignored, negative values clamped to 0) as a flat Radiance RGBE picture, rows as given. */
int rts_write_hdr(const char *path, const float *rgba32f, int width, int height);
void rts_fill_seed_image(uint32_t *dst, int width, int height, uint32_t seed);
/* Fills the uniform defaults of Renderer.swift:117-192 for a width x height target (frameIndex 0). */
void rts_default_uniforms(int width, int height, rt_uniforms *u);

#ifdef __cplusplus
}
#endif
#endif /* RT_SCENE_H */
