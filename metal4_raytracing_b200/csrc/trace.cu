// trace.cu — path-tracing kernel for sm_100a, megakernel layout (trace mode 0), plus the launch front end.
//
// B200-native counterpart of the reference's raytracingKernel (MetalRaytracing/Raytracing.metal:220-831) behind
// the same argument table (buffers by BufferIndex, images by TextureIndex, Uniforms, function constant
// maxSubmeshes). In this layout one thread owns one pixel for all of its samples and path segments; a CTA is one
// 16x16 screen tile (the reference's threadgroup, Renderer.swift:1445-1451) whose eight warps each cover an 8x4
// pixel block so primary rays of a warp stay coherent. Tiles come from the list this rank owns
// (tile % tileModulo == tileRemainder), so the same kernel serves 1..8 GPUs.
// ncu (profiles/r1_megakernel_*.md) shows this layout is divergence-bound — 11.4 of 32 threads active per
// instruction, 128 registers, instruction-cache stalls — which is why trace_wavefront.cu is the default.
// Compile with -fmad=false (numeric contract).
#include <algorithm>
#include <cstring>

#include "path_step.cuh"

namespace rtb {

__global__ void __launch_bounds__(256) k_trace_megakernel(const __grid_constant__ TraceParams P) {
  const rt_uniforms &U = P.uniforms;
  int px, py;
  if (!ownedPixel(P, blockIdx.x, threadIdx.x, px, py)) return;
  const size_t pixelIndex = size_t(py) * size_t(U.width) + size_t(px);
  const int lane = threadIdx.x & 31;

  const uint32_t offset = reinterpret_cast<const uint32_t *>(P.images[RT_TEXTURE_RANDOM].data)[pixelIndex];
  const f4 pm = readImage(P.images[RT_TEXTURE_MOTION], px, py);
  const f2 prevMotion = mk2(pm.x, pm.y);
  PrimaryOutputs prim = emptyPrimaryOutputs();
  f3 totalColor = mk3(0.0f);
  unsigned long long nClosest = 0, nAny = 0, nHits = 0;

  const int baseSamples = max(U.samplesPerPixel, 1);
  const int maxExtraSamples = (U.enableMotionAdaptiveSampling != 0) ? max(U.motionSamplingMaxExtraSamples, 0) : 0;
  const int sampleStride = baseSamples + maxExtraSamples;
  int totalSamples = baseSamples;

  for (int sampleIndex = 0; sampleIndex < totalSamples; ++sampleIndex) {
    if (sampleIndex % P.sampleModulo != P.sampleRemainder) continue; // sample partition: another dispatch owns it
    const int hIndex = haltonIndex(U, offset, sampleStride, sampleIndex);
    PathState s;
    startPath(U, px, py, hIndex, s);
    bool alive = s.bounce < U.maxBounces;
    while (alive) {
      RayHit hit;
      ++nClosest;
      const bool found = traverseScene<false>(P.tlas, s.origin.x, s.origin.y, s.origin.z, s.dir.x, s.dir.y, s.dir.z,
                                              0.0f, INFINITY, hit);
      if (P.primaryIds != nullptr && sampleIndex == 0 && s.step == 0) {
        const uint4 id = found ? make_uint4(hit.instance, hit.geometry, hit.primitive, __float_as_uint(hit.t))
                               : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        reinterpret_cast<uint4 *>(P.primaryIds)[pixelIndex] = id;
      }
      if (!found) {
        shadeMiss(P, s);
        break;
      }
      ++nHits;
      ShadowRequest shadow;
      alive = shadeSegment(P, s, hit, hIndex, sampleIndex, prevMotion, prim, shadow);
      if (shadow.valid) {
        ++nAny;
        RayHit sh;
        if (!traverseScene<true>(P.tlas, shadow.origin.x, shadow.origin.y, shadow.origin.z, shadow.dir.x, shadow.dir.y,
                                 shadow.dir.z, 0.0f, shadow.tmax, sh))
          s.radiance += shadow.contribution;
      }
    }
    totalColor += s.radiance;
    if (sampleIndex == 0 && maxExtraSamples > 0)
      totalSamples = adaptiveSampleCount(U, baseSamples, maxExtraSamples, prim.motion, prevMotion);
  }

  resolvePixel(P, px, py, totalColor, totalSamples, prevMotion, prim, true);

  if (P.rayCounters != nullptr) { // probe: warp-aggregated counters
    const unsigned mask = __activemask();
    for (int o = 16; o > 0; o >>= 1) {
      nClosest += __shfl_down_sync(mask, nClosest, o);
      nAny += __shfl_down_sync(mask, nAny, o);
      nHits += __shfl_down_sync(mask, nHits, o);
    }
    if (lane == __ffs(int(mask)) - 1) {
      atomicAdd(P.rayCounters + 0, nClosest);
      atomicAdd(P.rayCounters + 1, nAny);
      atomicAdd(P.rayCounters + 2, nHits);
    }
  }
}

// normalize(direction) and cos(coneAngle) of every light, evaluated once per dispatch with the same functions the
// per-hit code used (Raytracing.metal:620-636 evaluates them per hit), so results are unchanged.
__global__ void k_prepare_lights(const rt_light *lights, int count, float4 *derived) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const f3 d = normalize(mk3(lights[i].direction));
  derived[i] = make_float4(d.x, d.y, d.z, cosDet(lights[i].coneAngle));
}

// Flat TLAS: bounding sphere (centre, radius) of the union of the world boxes of the instances that have nodes to traverse
// (instanceBox lo.w == 0), once per dispatch; a far-away point when there is none. Feeds rayReachesNodes (traverse.cuh).
__global__ void k_prepare_classes(const float4 *instanceBox, uint32_t count, float4 *sphere) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (uint32_t k = 0; k < count; ++k) {
    const float4 l = instanceBox[2 * k], h = instanceBox[2 * k + 1];
    if (l.w != 0.0f) continue;
    lo[0] = fminf(lo[0], l.x), lo[1] = fminf(lo[1], l.y), lo[2] = fminf(lo[2], l.z);
    hi[0] = fmaxf(hi[0], h.x), hi[1] = fmaxf(hi[1], h.y), hi[2] = fmaxf(hi[2], h.z);
  }
  if (lo[0] > hi[0]) { // no instance with nodes: every ray is cheap
    sphere[0] = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.0f);
    return;
  }
  const float ex = 0.5f * (hi[0] - lo[0]), ey = 0.5f * (hi[1] - lo[1]), ez = 0.5f * (hi[2] - lo[2]);
  sphere[0] = make_float4(0.5f * (lo[0] + hi[0]), 0.5f * (lo[1] + hi[1]), 0.5f * (lo[2] + hi[2]), sqrtf(ex * ex + ey * ey + ez * ez));
}

int fillTraceParams(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], const rt_image textures[RT_TEXTURE_COUNT],
                    int maxSubmeshes, const rt_trace_options *opt, TraceParams &P) {
  RT_CHECK(buffers != nullptr && textures != nullptr, "rt_trace: null argument table");
  RT_CHECK(buffers[RT_BUFFER_UNIFORMS] != nullptr, "rt_trace: buffer 0 (Uniforms) is not bound");
  RT_CHECK(buffers[RT_BUFFER_ACCELERATION_STRUCTURE] != nullptr, "rt_trace: buffer 8 (acceleration structure) is not bound");
  RT_CHECK(buffers[RT_BUFFER_RESOURCES] != nullptr, "rt_trace: buffer 5 (resources) is not bound");
  RT_CHECK(buffers[RT_BUFFER_INSTANCE_DESCRIPTORS] != nullptr, "rt_trace: buffer 9 (instance descriptors) is not bound");
  RT_CHECK(buffers[RT_BUFFER_PREVIOUS_INSTANCE_DESCRIPTORS] != nullptr, "rt_trace: buffer 17 (previous instance descriptors) is not bound");
  RT_CHECK(buffers[RT_BUFFER_LIGHTS] != nullptr, "rt_trace: buffer 6 (lights) is not bound");
  RT_CHECK(maxSubmeshes >= 1, "rt_trace: function constant maxSubmeshes must be >= 1");
  std::memset(&P, 0, sizeof P);
  std::memcpy(&P.uniforms, buffers[RT_BUFFER_UNIFORMS], sizeof(rt_uniforms));
  RT_CHECK(P.uniforms.width > 0 && P.uniforms.height > 0, "rt_trace: empty render target");
  RT_CHECK(P.uniforms.lightCount >= 1, "rt_trace: lightCount must be >= 1");
  uint64_t tlasId = uint64_t(reinterpret_cast<uintptr_t>(buffers[RT_BUFFER_ACCELERATION_STRUCTURE]));
  auto it = ctx->accels.find(tlasId);
  RT_CHECK(it != ctx->accels.end() && it->second->isTlas, "rt_trace: buffer 8 is not a TLAS id from rt_tlas_build");
  const AccelObject *tl = it->second;
  P.tlas.nodes = tl->nodes;
  P.tlas.instances = tl->instances;
  P.tlas.leafInstance = tl->leafPrim;
  P.tlas.instanceBox = tl->instanceBox;
  P.tlas.instanceCount = tl->primCount;
  P.tlas.nodeCount = tl->primCount ? tl->nodeCount : 0u;
  P.tlasRootBox = tl->nodeBox;
  P.resources = static_cast<const rt_resource *>(buffers[RT_BUFFER_RESOURCES]);
  P.instances = static_cast<const rt_instance_descriptor *>(buffers[RT_BUFFER_INSTANCE_DESCRIPTORS]);
  P.prevInstances = static_cast<const rt_instance_descriptor *>(buffers[RT_BUFFER_PREVIOUS_INSTANCE_DESCRIPTORS]);
  P.lights = static_cast<const rt_light *>(buffers[RT_BUFFER_LIGHTS]);
  for (int i = 0; i < RT_TEXTURE_COUNT; ++i) P.images[i] = textures[i];
  const rt_image &rnd = textures[RT_TEXTURE_RANDOM], &dst = textures[RT_TEXTURE_PREVIOUS_ACCUMULATION];
  RT_CHECK(rnd.data && rnd.format == RT_FORMAT_R32_UINT && rnd.width == P.uniforms.width && rnd.height == P.uniforms.height,
           "rt_trace: texture 2 (random) must be an r32uint image of the render size");
  RT_CHECK(dst.data && dst.width == P.uniforms.width && dst.height == P.uniforms.height,
           "rt_trace: texture 1 (destination accumulation) is not bound at the render size");
  RT_CHECK(textures[RT_TEXTURE_MOTION].data && textures[RT_TEXTURE_DEPTH].data, "rt_trace: depth/motion images are not bound");
  RT_CHECK(P.uniforms.frameIndex == 0 || textures[RT_TEXTURE_ACCUMULATION].data, "rt_trace: texture 0 (history) is not bound");
  P.srgbLut = ctx->srgbLutDev;
  if (ctx->lightDerivedCap < P.uniforms.lightCount) {
    RT_CUDA(RT_SYNC_STREAM(ctx, ctx->stream));
    if (ctx->lightDerivedDev) cudaFree(ctx->lightDerivedDev);
    ctx->lightDerivedDev = nullptr;
    ctx->lightDerivedCap = 0;
    const int cap = std::max(16, P.uniforms.lightCount);
    RT_CUDA(cudaMalloc(&ctx->lightDerivedDev, size_t(cap + 2) * sizeof(float4))); // + the union box of k_prepare_classes
    ctx->lightDerivedCap = cap;
  }
  P.lightDerived = ctx->lightDerivedDev;
  P.nodeUnionBox = ctx->lightDerivedDev + ctx->lightDerivedCap;
  P.maxSubmeshes = maxSubmeshes;
  P.tileModulo = (opt && opt->tileModulo > 1) ? opt->tileModulo : 1;
  P.tileRemainder = (opt && opt->tileModulo > 1) ? opt->tileRemainder : 0;
  RT_CHECK(P.tileRemainder >= 0 && P.tileRemainder < P.tileModulo, "rt_trace: tileRemainder out of range");
  P.sampleModulo = (opt && opt->sampleModulo > 1) ? opt->sampleModulo : 1;
  P.sampleRemainder = (opt && opt->sampleModulo > 1) ? opt->sampleRemainder : 0;
  RT_CHECK(P.sampleRemainder >= 0 && P.sampleRemainder < P.sampleModulo, "rt_trace: sampleRemainder out of range");
  if (P.sampleModulo > 1) {
    RT_CHECK(P.tileModulo == 1, "rt_trace: the sample partition and the tile partition cannot be combined");
    RT_CHECK(dst.format == RT_FORMAT_RGBA32_FLOAT, "rt_trace: the sample partition needs an rgba32f destination (shares are summed)");
    RT_CHECK(P.uniforms.enableMotionAdaptiveSampling == 0 && P.uniforms.enableMotionAdaptiveAccumulation == 0,
             "rt_trace: the sample partition needs both motion-adaptive features off (they depend on sample 0's motion)");
  }
  P.tilesX = (P.uniforms.width + 15) / 16;
  P.tilesY = (P.uniforms.height + 15) / 16;
  P.primaryIds = opt ? opt->primaryIdsDev : nullptr;
  P.rayCounters = opt ? reinterpret_cast<unsigned long long *>(opt->rayCountersDev) : nullptr;
  if (opt && opt->environment && opt->environment->texelsDev) {
    P.env = *opt->environment;
    RT_CHECK(P.env.width > 0 && P.env.height > 0, "rt_trace: environment has no texels");
    RT_CHECK((P.env.flags & RT_ENV_IMPORTANCE) == 0u || P.env.cdfDev != nullptr,
             "rt_trace: RT_ENV_IMPORTANCE needs the table of rt_environment_cdf in cdfDev");
  }
  P.hints = opt ? opt->hints : 0u;
  P.peerCount = 0;
  if (opt && opt->peerAccumulation) {
    RT_CHECK(P.tileModulo <= 8, "rt_trace: at most 8 peers");
    P.peerCount = P.tileModulo;
    for (int p = 0; p < P.peerCount; ++p) P.peerAccumulation[p] = opt->peerAccumulation[p];
  }
  return 0;
}

int launchTrace(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], const rt_image textures[RT_TEXTURE_COUNT],
                int maxSubmeshes, const rt_trace_options *opt) {
  TraceParams P;
  RT_TRY(fillTraceParams(ctx, buffers, textures, maxSubmeshes, opt, P));
  const int tileCount = P.tilesX * P.tilesY;
  const int owned = (tileCount - P.tileRemainder + P.tileModulo - 1) / P.tileModulo;
  if (owned <= 0) return 0;
  k_prepare_lights<<<(P.uniforms.lightCount + 63) / 64, 64, 0, ctx->stream>>>(P.lights, P.uniforms.lightCount,
                                                                               ctx->lightDerivedDev);
  ++ctx->launches;
  if (P.tlas.instanceCount <= kFlatTlasMax && P.tlas.instanceBox != nullptr) {
    k_prepare_classes<<<1, 32, 0, ctx->stream>>>(P.tlas.instanceBox, P.tlas.instanceCount, ctx->lightDerivedDev + ctx->lightDerivedCap);
    ++ctx->launches;
  }
  // the wavefront layout packs (bounce, step, transparency passes) into 10 bits each; deeper paths than that
  // (maxBounces > 31, far beyond anything the reference's UI offers) run in the megakernel
  if (ctx->traceMode == 1 && P.uniforms.maxBounces <= 31) return launchTraceWavefront(ctx, P);
  ctx->mark(-1);
  k_trace_megakernel<<<owned, 256, 0, ctx->stream>>>(P);
  ctx->mark(RT_KERNEL_MEGAKERNEL);
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

} // namespace rtb
