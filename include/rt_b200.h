/*
 * rt_b200.h — C-ABI of the B200 (sm_100a) ray-tracing hot path (librt_b200.so).
 *
 * Every entry point replaces one Metal interface the reference's host code drives; the reference-side call it
 * stands in for is cited per function. Plain pointers and sizes only. All functions return 0 on success and a
 * non-zero code on failure (the reference traps or silently skips instead — SURVEY.md §5); rt_last_error()
 * returns the message for the calling thread. Nothing here falls back to the CPU: without a CUDA device
 * rt_create fails.
 *
 * Memory model: the host owns every buffer. "dev" pointers are CUDA device pointers from any
 * allocator (rt_malloc, cudaMalloc, a torch tensor's data_ptr). Work is enqueued on the context's stream in
 * call order, so skin -> refit -> TLAS update -> trace need no explicit barriers (the reference's missing
 * skinning->refit barrier, Renderer.swift:1312-1317, cannot occur).
 */
#ifndef RT_B200_H
#define RT_B200_H

#include "rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_context rt_context;

/* ---- context / stream (Renderer.init: device, MTL4CommandQueue; Renderer.swift:228-262) ----------------- */
int rt_create(int device, rt_context **out);
int rt_destroy(rt_context *ctx);
const char *rt_last_error(void);
/* Use an external CUDA stream (cudaStream_t as void*), e.g. torch's current stream; NULL = the context's own. */
int rt_set_stream(rt_context *ctx, void *cudaStream);
/* The stream the context currently enqueues on (cudaStream_t as void*): lets a caller order its own work — an NCCL
 * collective, a torch op — against the library's without moving the library onto the caller's stream. */
int rt_get_stream(rt_context *ctx, void **cudaStreamOut);
/* commitAndWait (Utilities.swift:122-128,240-246) */
int rt_sync(rt_context *ctx);
/* CUDA-event timer on the context's stream: begin/end bracket enqueued work; end synchronises. */
int rt_timer_begin(rt_context *ctx);
int rt_timer_end(rt_context *ctx, float *milliseconds);

/* ---- buffers (device.makeBuffer / contents().copyMemory; Renderer.swift:342-420) ------------------------- */
int rt_malloc(rt_context *ctx, size_t bytes, void **dev);
int rt_free(rt_context *ctx, void *dev);
int rt_malloc_host(rt_context *ctx, size_t bytes, void **pinnedHost); /* pinned staging memory */
int rt_free_host(rt_context *ctx, void *pinnedHost);
int rt_upload(rt_context *ctx, void *dstDev, const void *srcHost, size_t bytes);   /* stream-ordered */
int rt_download(rt_context *ctx, void *dstHost, const void *srcDev, size_t bytes); /* stream-ordered + sync */
int rt_copy(rt_context *ctx, void *dstDev, const void *srcDev, size_t bytes);      /* blit, Renderer.swift:1290-1303 */
/* Asynchronous read-back for frames in flight (the reference keeps three, Renderer.swift:207,1406-1409): the copy is
 * ordered after everything enqueued so far but runs on the context's copy stream, so work enqueued afterwards
 * overlaps it. dstHost should be pinned (rt_malloc_host). rt_download_wait blocks until the copy with that ticket —
 * and every earlier one — has landed. Up to 8 copies may be outstanding. */
/* Stream fences for host-side staging rings (the reference triple-buffers its per-frame host data, Renderer.swift:208-212):
 * rt_fence marks "everything enqueued so far", rt_fence_wait blocks until that point has executed. The ring holds 16
 * fences; waiting on an older ticket waits for the newer fence that took its slot (never returns early). */
int rt_fence(rt_context *ctx, uint64_t *ticket);
int rt_fence_wait(rt_context *ctx, uint64_t ticket);
int rt_download_async(rt_context *ctx, void *dstHost, const void *srcDev, size_t bytes, uint64_t *ticket);
int rt_download_wait(rt_context *ctx, uint64_t ticket);
int rt_memset(rt_context *ctx, void *dstDev, int value, size_t bytes);

/* ---- acceleration structures ------------------------------------------------------------------------------
 * rt_blas_build: MTLAccelerationStructure build (+ compaction) of one primitive AS with one triangle geometry
 * per submesh (Utilities.swift:100-290, Renderer.swift:509-529, Mesh.swift:84-101). geoms is a HOST array whose
 * vertex/index pointers are DEVICE pointers. flags: RT_AS_FLAG_COMPACT | RT_AS_FLAG_REFITTABLE. The returned id
 * is what instance descriptors carry in accelerationStructureID.
 * rt_blas_refit: refit of a skinned mesh's BLAS in place (Renderer.swift:1084-1202): same topology, boxes and
 * triangle records recomputed from the current vertex buffer contents.
 * rt_tlas_build / rt_tlas_update: instance AS over `count` 72-byte descriptors in DEVICE memory
 * (Renderer.swift:547-606, 937-973); update = rebuild from the descriptors' current contents.
 * rt_tlas_refit: the reference's per-frame path when the device supports refitting (Renderer.swift:1084-1202): same
 * instance count as the last build, tree topology kept, instance records (transforms, BLAS pointers) and every node
 * box recomputed from the descriptors' current contents. Results never depend on which of the two is used; a tree
 * refitted over instances that moved far apart just traverses slower until it is rebuilt.
 * Neither call blocks the host for up to 65,536 instances: hierarchy, wide nodes and parent links of a rebuild come
 * from a single resident CTA (<= 8 instances: one node, one thread), node counts stay on the device, and a tree too
 * deep for the traversal stack is reported by the next call on that TLAS instead of being waited for. */
int rt_blas_build(rt_context *ctx, const rt_triangle_geometry *geoms, uint32_t geometryCount, uint32_t flags,
                  uint64_t *outId);
int rt_blas_refit(rt_context *ctx, uint64_t id, const rt_triangle_geometry *geoms, uint32_t geometryCount);
int rt_blas_destroy(rt_context *ctx, uint64_t id);
int rt_tlas_build(rt_context *ctx, const rt_instance_descriptor *descriptorsDev, uint32_t count, uint64_t *outId);
int rt_tlas_update(rt_context *ctx, uint64_t id, const rt_instance_descriptor *descriptorsDev, uint32_t count);
int rt_tlas_refit(rt_context *ctx, uint64_t id, const rt_instance_descriptor *descriptorsDev, uint32_t count);
int rt_tlas_destroy(rt_context *ctx, uint64_t id);

typedef struct rt_as_info {
  uint32_t primitiveCount; /* triangles (BLAS) or instances (TLAS) */
  uint32_t wideNodeCount;  /* 80-byte 8-wide nodes */
  uint32_t levelCount;
  uint32_t _pad;
  uint64_t bytes;          /* resident device bytes */
  float boundsMin[3], boundsMax[3];
  float sahCost;           /* surface-area-heuristic cost of the wide tree (node + leaf terms) */
  float _pad2;
} rt_as_info;
int rt_as_get_info(rt_context *ctx, uint64_t id, rt_as_info *out);

/* rt_intersect: the reference's intersector call on its own — intersector<triangle_data, instancing>::intersect()
 * for a closest hit (Raytracing.metal:301-318) or, with RT_INTERSECT_ANY, accept_any_intersection(true) as for its shadow
 * rays (:664-665, :730-737) — over `count` caller-supplied rays in DEVICE memory against the TLAS `tlasId`. hitsDev[i]
 * answers raysDev[i]: closest hit = smallest t with tmin < t < tmax, ties -> smallest (instance, geometry, primitive);
 * (u, v) are the weights of the triangle's second and third vertex (Metal's triangle_barycentric_coord); a miss has
 * t = +inf and ids 0xFFFFFFFF. RT_INTERSECT_ANY: t = 0 when anything lies in (tmin, tmax), +inf otherwise; the other
 * fields are not meaningful. The direction need not be normalised (t is in units of its length). Stream-ordered, no
 * host synchronisation; it runs the same traversal iteration as rt_trace. */
typedef struct rt_ray {
  float origin[3], tmin;
  float direction[3], tmax;
} rt_ray; /* 32 bytes, 16-byte aligned in device memory */
typedef struct rt_ray_hit {
  float t, u, v;
  uint32_t instance, geometry, primitive;
} rt_ray_hit; /* 24 bytes */
#define RT_INTERSECT_ANY 1u
int rt_intersect(rt_context *ctx, uint64_t tlasId, const rt_ray *raysDev, uint32_t count, uint32_t flags, rt_ray_hit *hitsDev);

/* ---- kernels ------------------------------------------------------------------------------------------------
 * rt_skin: skinningKernel dispatch (Skinning.metal:7-49; SkinningPass.swift:160-211). buffers[] is the argument
 * table indexed by BufferIndex: 10 rest positions, 11 rest normals, 12 joint indices (ushort4), 13 joint weights
 * (float4), 14 joint matrices (float4x4 column-major), 15 skinned positions (out), 16 skinned normals (out); all
 * device pointers. Index 0 (vertexCount in the reference) is passed by value. */
int rt_skin(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], uint32_t vertexCount);

/* Environment light — an EXTENSION, off unless bound. The reference's kernel ends a path that misses with no
 * contribution (Raytracing.metal:320-322) and never loads the vulture_hide_4k.hdr it ships; BASELINE's north_star
 * asks for an "HDR environment lookup", so it is specified here (and in the oracle): a closest-hit ray that
 * misses adds throughput * intensity * bilinear(texels, u, v) with the equirectangular mapping
 *   u = atan2(d.z, d.x) / (2 pi) + 0.5,  v = acos(d.y) / pi   (row 0 = straight up, columns wrap, rows clamp),
 * texel centres at (i + 0.5) / size, then the path ends as before.
 *
 * RT_ENV_IMPORTANCE additionally samples the environment as a light. With the flag set and a
 * table from rt_environment_cdf bound, the uniform light pick of Raytracing.metal:587-590 chooses among
 * lightCount + 1 lights, the last one being the environment: the two Halton dimensions of the area-light sample
 * (2 + 6 step + 1, + 2) pick a row from the marginal and a column from that row's conditional distribution
 * (piecewise constant over texels, proportional to luminance * sin(theta)), a point inside the texel, and with it a
 * direction L; its solid-angle density is p_env = p_row * p_col * width * height / (2 pi^2 sin(theta)). The
 * environment's radiance along L (the bilinear lookup above) is combined with the cosine-hemisphere bounce by the
 * balance heuristic: the light sample carries radiance / (p_env / n + p_bsdf) with p_bsdf = max(N.L, 0) / pi and
 * n = lightCount + 1 (this replaces the `* lightCount` of the other lights), its shadow ray has no far end, and a
 * bounce ray that misses picks the environment up with weight p_bsdf / (p_bsdf + p_env / n). Camera rays and glass
 * reflections / refractions are not sampled by the light and keep weight 1. All of it is evaluated in fp32 with
 * the operation order of csrc/shade.cuh, which the oracle repeats. */
#define RT_ENV_IMPORTANCE 1u
/* cdfDev is a table of this library's rt_environment_cdf, which appends guide tables to the running sums: for the
 * marginal and for every row, 65 entries g[k] = the largest cell index whose running sum is <= k / 64. The device search
 * for xi then starts inside xi's 1/64 bucket instead of at [0, n) — the same cell comes out (the search invariant is
 * unchanged) after 3 - 5 dependent loads instead of 11 - 12. Without the flag the table is searched from scratch, so a
 * table built elsewhere (the oracle builds its own) still works. */
#define RT_ENV_GUIDED 2u
#define RT_ENV_GUIDE_CELLS 64
typedef struct rt_environment {
  const float *texelsDev; /* RGBA32F, width * height texels in device memory (host memory for the oracle) */
  int32_t width, height;
  float intensity;
  uint32_t flags;         /* RT_ENV_* */
  const float *cdfDev;    /* rt_environment_cdf's table in device memory (host memory for the oracle), or NULL */
} rt_environment;
/* Builds the sampling table of RT_ENV_IMPORTANCE on the host (+ the guide tables of RT_ENV_GUIDED behind it, see
 * rt_environment_cdf_floats): (height + 1) marginal values followed by height rows
 * of (width + 1) conditional values, each a running sum normalised to [0, 1] (accumulated in double in row / column
 * order, stored as float; weight = (0.2126 r + 0.7152 g + 0.0722 b) * sin(pi (row + 0.5) / height); a row or a map
 * without weight becomes uniform). cdfOut: rt_environment_cdf_floats(width, height) floats. Needs no context. */
size_t rt_environment_cdf_floats(int32_t width, int32_t height);
int rt_environment_cdf(const float *texelsHost, int32_t width, int32_t height, float *cdfOutHost);

/* rt_joint_palette: the joint-palette computation the reference does on the host every animated frame, as one
 * kernel: local[j] = T * R * S with the rotation quaternion renormalised (Model.update,
 * Model.swift:207-261, matrix4x4_trs :497-506); global[j] = global[parent[j]] * local[j], parents precede children
 * (Skeleton hierarchy, Model.swift:379-387); palette[j] = global[j] * inverseBind[j] (SkinningPass.swift:123-157).
 * localTRS: jointCount x 10 floats {translation xyz, quaternion xyzw (any length), scale xyz}; parents: int32, a
 * negative or forward index = root; inverseBind / paletteOut: jointCount x float4x4 column-major. All device
 * pointers; jointCount <= 1024. The products are evaluated in the reference's order, so the palette equals the
 * host-computed one bit for bit. */
int rt_joint_palette(rt_context *ctx, const float *localTRSDev, const int32_t *parentsDev, const float *inverseBindDev,
                     uint32_t jointCount, float *paletteOutDev);

/* Optional extras of rt_trace that have no counterpart in the reference's binding table. Zero-initialise. */
typedef struct rt_trace_options {
  int32_t tileModulo;    /* multi-GPU ownership: this call renders 16x16 tiles with tile % tileModulo == */
  int32_t tileRemainder; /*   tileRemainder (0/1 and 0 = every tile) */
  uint32_t *primaryIdsDev;   /* probe: 4 x u32 per pixel (instance, geometry, primitive, t bits) of sample 0's
                                first intersect call; 0xFFFFFFFF x4 on a miss. NULL = off */
  uint64_t *rayCountersDev;  /* probe: 24 x u64, accumulated: {closest-hit rays, any-hit rays, closest hits} and, filled
                                only by the counter build of the library (-DRT_COUNT_WORK), {node steps, triangle
                                tests, instance entries} of the closest-hit rays, the same three of the any-hit rays,
                                a histogram of iterations per ray and the warps' iterations after their queue ran dry */
  void *const *peerAccumulation; /* multi-GPU: tileModulo device pointers to every rank's destination
                                    accumulation image (same format/size); owned tiles are also stored there
                                    through NVLink peer mappings. NULL = local only */
  const rt_environment *environment; /* HOST pointer; NULL = the reference's behaviour (a miss is black) */
  uint32_t hints;                    /* RT_TRACE_HINT_* promises of the caller; 0 = none */
  uint32_t _pad;
  int32_t sampleModulo;  /* multi-GPU sample partition (the alternative to tiles for multi-spp frames): */
  int32_t sampleRemainder; /* this call traces the samples s of every pixel with s % sampleModulo == sampleRemainder
                            (0/1 and 0 = every sample) and writes its share of the frame,
                              (sum of its samples) / samplesPerPixel * (1 - w)  [+ w * history on the call with remainder 0],
                            w = the EMA history weight (0 on frame 0), so that the SUM of the destination images of all
                            sampleModulo calls is the frame (an NCCL all-reduce when the calls run on different GPUs;
                            metal4_raytracing_b200/parallel.py mode "samples"). Needs an rgba32f destination, tileModulo
                            <= 1 and both motion-adaptive features off. Depth, motion, G-buffer and primary ids come
                            from the call that owns sample 0. Float sums are reassociated, so the frame equals the
                            single-call frame to rounding (relative 1e-6), not bit for bit as the tile partition does. */
} rt_trace_options;
/* The caller promises that no Material bound in this dispatch has a textureFlags bit set. The shading kernel is then
 * a build without the texture paths (about a tenth faster); a material that breaks the promise is shaded as if its
 * maps were absent. rtr_draw sets the hint by itself from the scene it was given. */
#define RT_TRACE_HINT_UNTEXTURED 1u
/* The caller promises that no Material bound in this dispatch takes the glass branch (opacity >= 0.999 and
 * refractionIndex <= 1.01, Raytracing.metal:517-519). A path then has at most maxBounces segments instead of
 * maxBounces (maxBounces + 1), and the dispatch launches that many segment passes (without the hint the extra passes
 * are launched too and end at once when their queue is empty — no device->host read-back either way); a material
 * that breaks the promise has its paths cut after maxBounces segments. rtr_draw sets it from the scene. */
#define RT_TRACE_HINT_NO_GLASS 2u
/* Not a promise but a switch: the reference's compile-time ENABLE_AO (ShaderTypes.h:155-157; 0 in the shipping build).
 * With it a material whose textureFlags has MATERIAL_TEXTURE_AO samples its ambient-occlusion map (x channel) and the
 * value scales the throughput of the next bounce (Raytracing.metal:405-409,442-446,672,748); debug view 5 shows it
 * (:475-479). Ignored under RT_TRACE_HINT_UNTEXTURED. rtr_draw passes it when the renderer was created with
 * RTR_FLAG_ENABLE_AO. */
#define RT_TRACE_ENABLE_AO 4u

/* rt_trace: raytracingKernel dispatch (Raytracing.metal:220-831; binding block Renderer.swift:1453-1490).
 * buffers[]: 0 Uniforms (HOST pointer; copied into the launch like a `constant` argument), 5 Resource rows (dev),
 * 6 lights (dev), 8 TLAS id from rt_tlas_build (cast to pointer), 9 instance descriptors (dev), 17 previous
 * instance descriptors (dev). textures[]: the nine images of TextureIndex (device memory): 2 random (read),
 * 0 history (read), 1 destination (write), 3 depth, 4 motion (read-write), 5..8 G-buffer.
 * Function constants: 0 resourcesStride (accepted, unused — as in the reference), 1 maxSubmeshes. */
int rt_trace(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], const rt_image textures[RT_TEXTURE_COUNT],
             int resourcesStride, int maxSubmeshes, const rt_trace_options *options);

/* Upload an RGBA8 texture and return the device address of its rt_texture2d record (a Resource texture slot). */
int rt_texture_create(rt_context *ctx, const uint8_t *rgba8Host, int width, int height, int srgb,
                      const rt_texture2d **outRecordDev);
int rt_texture_destroy(rt_context *ctx, const rt_texture2d *recordDev);

/* ---- display transform ---------------------------------------------------------------------
 * rt_tonemap: the reference's presentation shader, color / (1 + color) (Shaders.metal:38-52), from an rgba16f /
 * rgba32f image into width*height RGBA8 pixels in device memory. RT_TONEMAP_SRGB applies the sRGB transfer an
 * *_srgb drawable applies on store; RT_TONEMAP_FLIP_Y writes row 0 = top of the picture (the kernel's row 0 is the
 * bottom of the view and the reference's quad un-flips it on screen, Shaders.metal:30-35). */
#define RT_TONEMAP_SRGB 1u
#define RT_TONEMAP_FLIP_Y 2u
int rt_tonemap(rt_context *ctx, const rt_image *srcDev, uint8_t *dstRGBA8Dev, uint32_t flags);

/* rt_temporal_filter: temporal reprojection of the accumulation image with the kernel's own depth (texture 3), motion
 * (texture 4) and normal G-buffer (texture 7, needs enableDenoiseGBuffer) — the consumer of those outputs, which the
 * reference hands to MetalFX's temporal denoiser (FramePresenter.swift:435-521). Per pixel: previous position =
 * (x - motion.x, y + motion.y); the bilinear history sample is accepted when the nearest history texel agrees in
 * depth (|dz| <= depthTolerance * z) and normal (dot >= normalThreshold), clamped to the current 3x3 colour range
 * and blended with historyWeight. history may be NULL (first frame: output = colour). All images device memory. */
typedef struct rt_denoise_frame {
  rt_image color;  /* rgba16f / rgba32f */
  rt_image motion; /* rg16f / rg32f, pixels, +y down (as the kernel writes it) */
  rt_image depth;  /* r32f, view depth; 1e8 = no primary hit (passed through) */
  rt_image normal; /* rgba16f / rgba32f, n * 0.5 + 0.5 */
} rt_denoise_frame;
int rt_temporal_filter(rt_context *ctx, const rt_denoise_frame *current, const rt_denoise_frame *history,
                       const rt_image *outColorDev, float historyWeight, float depthTolerance, float normalThreshold);
/* rt_spatial_filter: one pass of an edge-avoiding a-trous wavelet filter (Dammertz et al. 2010) over frame->color,
 * guided by the kernel's depth (texture 3) and normal G-buffer (texture 7) — the spatial half of what the reference
 * leaves to MetalFX's denoiser (FramePresenter.swift:435-521); frame->motion is not read and may be unbound. 5 x 5 taps
 * spaced `step` pixels apart (call it with step 1, 2, 4, ... ping-ponging two images) with the B3-spline weights
 * (1, 4, 6, 4, 1) / 16 per axis, each multiplied by
 *   max(0, n_p . n_q) squared normalSquarings times,   1 / (1 + (|z_p - z_q| / (depthSigma * z_p))^2)   and, when
 *   colorSigma > 0,   1 / (1 + |c_p - c_q|^2 / colorSigma^2);
 * taps outside the image or without a primary hit (depth >= 1e7) are skipped, a pixel without a primary hit is passed
 * through. out = sum(w * c_q) / sum(w), accumulated row by row in fp32. */
int rt_spatial_filter(rt_context *ctx, const rt_denoise_frame *frame, const rt_image *outColorDev, int step,
                      float depthSigma, int normalSquarings, float colorSigma);

/* ---- multi-GPU frame exchange (no counterpart in the single-device reference) ---------------
 * Rank g of N owns 16x16 tiles with tile % N == g. Two ways to assemble the frame:
 *  (a) rt_trace with options->peerAccumulation: the kernel stores owned pixels straight into every rank's frame
 *      through NVLink peer mappings obtained with rt_ipc_export / rt_ipc_import (one kernel = compute + exchange);
 *  (b) rt_pack_tiles -> NCCL all-gather of the slabs (done by the caller) -> rt_unpack_tiles.
 * Slab layout: [ceil(tileCount / N) tiles][256 pixels][bytes per pixel]; rank-major after the gather. */
int rt_pack_tiles(rt_context *ctx, const rt_image *imageDev, void *slabDev, int tileModulo, int tileRemainder);
int rt_unpack_tiles(rt_context *ctx, const void *slabsDev, const rt_image *imageDev, int tileModulo);
int rt_ipc_export(rt_context *ctx, void *dev, unsigned char handle[64]);
int rt_ipc_import(rt_context *ctx, const unsigned char handle[64], void **outDev);
int rt_ipc_close(rt_context *ctx, void *importedDev);

/* Number of kernels this library has launched on the context since creation (bench.py's gpu_launches). */
uint64_t rt_launch_count(rt_context *ctx);
/* Number of times the library itself blocked the host on the device outside the calls that exist to wait (rt_sync,
 * rt_download, rt_timer_end, rt_fence_wait, rt_download_wait, rt_kernel_timing_read): build-time read-backs, scratch
 * growth, pageable uploads. The per-frame paths (rt_skin, rt_blas_refit, rt_tlas_refit / _update, rt_trace) add none. */
uint64_t rt_host_sync_count(rt_context *ctx);
/* Select the trace kernel layout: 0 = megakernel, 1 = wavefront (default). */
int rt_set_trace_mode(rt_context *ctx, int mode);
/* Tuning knobs that never change results: "trace_mode" (0/1), "traversal_variant" (0..2, traverse.cuh),
 * "blocks_per_sm" (persistent grid size of the traversal kernels; default 0 = as many CTAs as are resident: 8 per SM
 * for dispatches of at least 6 M paths per pipeline lane, otherwise 7), "sample_batch" (1..64, samples of a pixel the
 * wavefront layout keeps in flight at once; default 16), "pipeline_lanes" (1..4 independent tile subsets of a
 * dispatch whose kernel sequences run on separate streams so that launch tails overlap; default 0 = two lanes for
 * dispatches of at least 16 M paths, otherwise one), "classify_rays" (0 / 1, default 1: with a TLAS of at most 8 instances
 * new rays are queued by class — likely to walk a BVH / cheap — so that warps hold rays of one kind and the long rays start first), "ploc_radius" (builder: PLOC neighbour search radius for
 * acceleration structures built after the call, default 16; 0 = plain LBVH), "tlas_ploc_radius" (the same for instance
 * acceleration structures; default 0 = automatic, up to 256 for small instance counts), "leaf_size" / "tlas_leaf_size" (1..3
 * triangles / instances per leaf slot of a wide node, defaults 3 / 1). */
int rt_set_option(rt_context *ctx, const char *key, int value);
/* Per-kernel-class device timing (bench.py's roofline of the dominant kernel). While enabled, the library records
 * a CUDA event on the context's stream after each of its launches; rt_kernel_timing_read synchronises, returns the
 * summed milliseconds and launch counts per class since the last read and resets them. Classes: */
enum {
  RT_KERNEL_GENERATE = 0, /* camera rays / sample bookkeeping */
  RT_KERNEL_TRACE = 1,    /* closest-hit traversal (dominant) */
  RT_KERNEL_SHADE = 2,    /* material + light sampling + next ray */
  RT_KERNEL_SHADOW = 3,   /* any-hit traversal */
  RT_KERNEL_RESOLVE = 4,  /* sample mean + EMA + image writes */
  RT_KERNEL_SKIN = 5,
  RT_KERNEL_REFIT = 6,
  RT_KERNEL_BUILD = 7,    /* BLAS / TLAS builds */
  RT_KERNEL_MEGAKERNEL = 8,
  RT_KERNEL_OTHER = 9,
  RT_KERNEL_CLASS_COUNT = 10
};
int rt_kernel_timing_enable(rt_context *ctx, int enable);
int rt_kernel_timing_read(rt_context *ctx, float msByClass[RT_KERNEL_CLASS_COUNT],
                          uint32_t launchesByClass[RT_KERNEL_CLASS_COUNT]);
/* Device self-test: for raysPerNode pseudo-random rays per wide node of an acceleration structure, the fast child-box
 * test must report every child the plain-conversion form reports. out[0] = missed children (must be 0), out[1] = extra
 * (allowed: the fast form is slightly more conservative), out[2] = tests run, out[3..10] = first failure details. */
int rt_selftest_child_boxes(rt_context *ctx, uint64_t id, uint32_t raysPerNode, uint32_t seed, uint64_t out[11]);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
