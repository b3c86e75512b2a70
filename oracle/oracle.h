/*
 * oracle.h — C entry points of the CPU oracle (liboracle_rt.so). TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; nothing under metal4_raytracing_b200/ does. PARITY UNPINNED: the reference (Swift + Metal) cannot
 * be built or run on Linux and ships no tests or golden vectors, so this restatement of
 * MetalRaytracing/Raytracing.metal and MetalRaytracing/Skinning.metal is itself the specification.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include "../include/rt_scene.h"
#include "../include/rt_b200.h"
#include "../include/rt_types.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_ctx oracle_ctx;

/* Builds resources, skins frame 0 with the scene's current palettes, builds one SAH BLAS per mesh and the TLAS
 * (the restatement of Renderer.createBuffers / createMTL4AccelerationStructures). The scene memory must outlive
 * the context. threads <= 0 means all cores. */
oracle_ctx *oracle_create(const rt_scene_desc *scene, int threads);
void oracle_destroy(oracle_ctx *c);
/* Per-frame update (Renderer.updateSkinningAndBLAS): previous positions <- current, re-skin with the palettes now
 * in `scene`, rebuild skinned BLAS, take instance transforms (current + previous) from `scene`, rebuild TLAS. */
int oracle_update(oracle_ctx *c, const rt_scene_desc *scene);
/* One dispatch of raytracingKernel over the pixels of 16x16 tiles with tileIndex % tileModulo == tileRemainder
 * (1, 0 = whole frame). textures[9] are host images indexed by TextureIndex. primaryIds (optional) receives
 * 4 x u32 per pixel. stats (optional) receives {closest rays, any-hit rays, closest hits}. */
int oracle_render(oracle_ctx *c, const rt_uniforms *uniforms, const rt_image textures[9], uint32_t *primaryIds,
                  uint64_t stats[3], int tileModulo, int tileRemainder);
/* The reference's compile-time ENABLE_AO (ShaderTypes.h:155-157, Raytracing.metal:405-409,442-446,475-479; default 0):
 * when on, a material with MATERIAL_TEXTURE_AO samples its ambient-occlusion map (x channel), which scales the
 * throughput of the next bounce (Raytracing.metal:672,748), and debug view 5 shows it. */
void oracle_set_enable_ao(oracle_ctx *c, int enable);
/* Environment extension (include/rt_b200.h rt_environment; texels in HOST memory here). NULL switches it off. */
int oracle_set_environment(oracle_ctx *c, const rt_environment *env);
void oracle_sample_environment(const rt_environment *env, const float dir[3], float out_rgb[3]); /* KAT probe */
/* RT_ENV_IMPORTANCE's table, (height + 1) + height * (width + 1) floats (the oracle's own restatement of
 * rt_environment_cdf); bind it through rt_environment.cdfDev (a HOST pointer here). */
int oracle_environment_cdf(const float *texels, int width, int height, float *out);
/* Current skinned streams of a mesh (float4 per vertex). */
int oracle_get_mesh_streams(oracle_ctx *c, int mesh, float *positions4, float *normals4, float *prevPositions4);

/* kernel-level mirrors and known-answer probes */
void oracle_skin(const void *const buffers[18], uint32_t vertexCount); /* skinningKernel over all vertices */
float oracle_halton(int i, int d);
int oracle_intersect_triangle(const float origin[3], const float dir[3], const float v0[3], const float v1[3],
                              const float v2[3], float tmin, float tmax, float out_tuv[3]);
void oracle_sample_texture(const rt_texture2d *t, float u, float v, float out[4]);
void oracle_invert_affine(const float m4x3[12], float inv[12]);
uint16_t oracle_float_to_half(float f);
float oracle_half_to_float(uint16_t h);
/* closest hit of one world ray against the context's TLAS: out = {valid, instance, geometry, primitive}, tuv */
int oracle_trace_ray(oracle_ctx *c, const float origin[3], const float dir[3], float tmin, float tmax,
                     uint32_t out_ids[4], float out_tuv[3]);
int oracle_thread_count(oracle_ctx *c);

#ifdef __cplusplus
}
#endif
#endif
