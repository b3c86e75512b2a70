/*
 * rt_renderer.h — C-ABI of the host-side frame orchestrator that sits on top of rt_b200.h.
 *
 * It is the C++ mirror of the hot-path half of the reference's Renderer (MetalRaytracing/Renderer.swift):
 *   rtr_create   = createBuffers (:342-420) + createTextures (:676-799) + createMTL4AccelerationStructures (:464-606)
 *   rtr_update   = updateSkinningAndBLAS (:1280-1326): descriptors cur->prev, skinned cur->prev, skinning
 *                  dispatch, BLAS refit, TLAS refit (or rebuild: RTR_FLAG_TLAS_REBUILD; :1084-1202)
 *   rtr_draw     = the binding block + dispatch + accumulation swap of draw(in:) (:1445-1494)
 * It consumes the flat host scene of rt_scene.h and only ever calls the rt_* entry points; the GUI, presenter
 * and MetalFX parts of Renderer.swift are out of scope. Lives in librt_b200.so.
 */
#ifndef RT_RENDERER_H
#define RT_RENDERER_H

#include "rt_b200.h"
#include "rt_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtr_renderer rtr_renderer;

const char *rtr_last_error(void);

#define RTR_FLAG_FP32_IMAGES 1u /* accumulation / motion / G-buffer in fp32 instead of the reference's fp16 formats */
#define RTR_FLAG_REBUILD_SKINNED 2u /* full BLAS rebuild instead of refit after skinning */
#define RTR_FLAG_GPU_SKELETON 4u    /* joint palettes evaluated on the device from local TRS (rt_joint_palette) instead of
                                       uploaded from the host scene (SkinningPass.swift:123-157 runs them on the CPU) */

#define RTR_FLAG_ENABLE_AO 8u       /* the reference's compile-time ENABLE_AO (ShaderTypes.h:155-157): materials with
                                       MATERIAL_TEXTURE_AO sample their ambient-occlusion map (RT_TRACE_ENABLE_AO) */
#define RTR_FLAG_TLAS_REBUILD 16u   /* rtr_update rebuilds the TLAS from scratch every frame instead of refitting it
                                       (the reference refits when the device supports it, Renderer.swift:1084-1202) */

int rtr_create(rt_context *ctx, const rt_scene_desc *scene, int width, int height, uint32_t flags,
               rtr_renderer **out);
int rtr_destroy(rtr_renderer *r);
/* r32uint per-pixel Halton offsets, width*height values from host memory (Renderer.swift:712-735). */
int rtr_set_seeds(rtr_renderer *r, const uint32_t *seedsHost);
/* Per-frame scene update from the (mutated) host scene: instance transforms, joint palettes, lights. */
int rtr_update(rtr_renderer *r, const rt_scene_desc *scene);
/* One frame: binds the argument table, dispatches the kernel, swaps the accumulation targets.
 * options may be NULL. After the call TextureIndexAccumulation (0) holds the frame just rendered. */
int rtr_draw(rtr_renderer *r, const rt_uniforms *uniforms, const rt_trace_options *options);
/* Copies the image bound at `textureIndex` to host memory (size = width*height*bytes-per-pixel of its format). */
int rtr_read_image(rtr_renderer *r, int textureIndex, void *dstHost, size_t bytes);
int rtr_image_info(rtr_renderer *r, int textureIndex, rt_image *out); /* device pointer + format */
/* Clears history/motion images (what a resize does in the reference, Renderer.swift:1417-1425). */
int rtr_reset_accumulation(rtr_renderer *r);
/* Test probes. */
int rtr_read_mesh_streams(rtr_renderer *r, int mesh, float *positions4, float *normals4, float *prevPositions4);
int rtr_get_blas_id(rtr_renderer *r, int mesh, uint64_t *id);
int rtr_get_tlas_id(rtr_renderer *r, uint64_t *id);
int rtr_mesh_count(rtr_renderer *r);

#ifdef __cplusplus
}
#endif
#endif
