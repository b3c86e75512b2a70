// trace_wavefront.cu — path-tracing kernels for sm_100a, wavefront layout (trace mode 1, the default).
//
// Same contract as trace.cu (the reference's raytracingKernel, MetalRaytracing/Raytracing.metal:220-831, behind
// its argument table) but split by phase so that every kernel runs converged and small:
//
//   for each sample index s (the reference's sample loop, :269):
//     generate      finish sample s-1 of every owned pixel, start sample s: camera ray -> path state, queue
//     repeat for each path segment (the reference's bounce loop, :311):
//       trace       closest hit for every queued path                       (traverse.cuh, ~64 registers)
//       shade       one shadeSegment() per hit: emission, light sample, BRDF, next ray; emits a shadow
//                   request and re-queues surviving paths (warp-aggregated queue appends = ray compaction)
//       shadow      any-hit for every shadow request; unoccluded ones add their contribution
//   resolve         sample mean, EMA with history, image writes (+ NVLink peer stores for multi-GPU)
//
// Paths that miss or die leave the queue, so later segments run on dense warps instead of the megakernel's
// 11-of-32 active threads (profiles/). Per-pixel results do not depend on queue order, and the arithmetic is the
// shared path_step.cuh, so this layout is bit-identical to the megakernel and to the CPU oracle.
// All kernels are persistent-style: a fixed grid of (SM count x resident CTAs) strides over the queue, whose
// length is read from device memory, so no host round trip is needed between phases.
#include <cstring>

#include "path_step.cuh"

namespace rtb {

namespace {

constexpr int kBlock = 256;

struct WfState {
  uint32_t capacity; // slots = owned tiles * 256
  float4 *rayO, *rayD;
  float4 *thr;  // throughput.xyz, w = halton index bits
  float4 *rad;  // radiance.xyz
  int4 *ctr;    // bounce, step, transparencyPasses, unused
  float4 *tot;  // totalColor.xyz, w = totalSamples bits
  float4 *mot;  // motion.xy, prevMotion.xy
  float4 *misc; // depth, flags bits (1 = hadPrimaryHit, 2 = wroteGBuffer), seed offset bits, unused
  float4 *hitA; // t, u, v, valid
  uint4 *hitB;  // instance, geometry, primitive, 0
  float4 *shO, *shD, *shC; // shadow origin + tmax, direction, contribution
  uint32_t *queue[2], *shadowQueue;
  uint32_t *counts; // [0], [1] path queues, [2] shadow queue
};

__device__ __forceinline__ void slotPixel(const TraceParams &P, uint32_t slot, int &px, int &py, bool &valid) {
  valid = ownedPixel(P, int(slot >> 8), int(slot & 255u), px, py);
}

// Appends `slot` for every lane with `push` set; one atomic per warp.
__device__ __forceinline__ void queuePush(uint32_t *queue, uint32_t *count, bool push, uint32_t slot) {
  const unsigned active = __activemask();
  const unsigned votes = __ballot_sync(active, push);
  if (votes == 0u) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(int(votes)) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(count, uint32_t(__popc(votes)));
  base = __shfl_sync(active, base, leader);
  if (push) queue[base + __popc(votes & ((1u << lane) - 1u))] = slot;
}

__global__ void __launch_bounds__(kBlock) k_wf_generate(const __grid_constant__ TraceParams P, const WfState W,
                                                        int sampleIndex, int baseSamples, int maxExtraSamples) {
  const rt_uniforms &U = P.uniforms;
  const int sampleStride = baseSamples + maxExtraSamples;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < W.capacity; slot += gridDim.x * blockDim.x) {
    int px, py;
    bool valid;
    slotPixel(P, slot, px, py, valid);
    bool push = false;
    if (valid) {
      const size_t pixelIndex = size_t(py) * size_t(U.width) + size_t(px);
      f3 total;
      int totalSamples;
      uint32_t offset;
      if (sampleIndex == 0) {
        offset = reinterpret_cast<const uint32_t *>(P.images[RT_TEXTURE_RANDOM].data)[pixelIndex];
        const f4 pm = readImage(P.images[RT_TEXTURE_MOTION], px, py);
        total = mk3(0.0f);
        totalSamples = baseSamples;
        W.mot[slot] = make_float4(0.0f, 0.0f, pm.x, pm.y);
        W.misc[slot] = make_float4(1.0e8f, __uint_as_float(0u), __uint_as_float(offset), 0.0f);
        if (U.enableDenoiseGBuffer != 0) { // a pixel whose first segment misses keeps zeros (Raytracing.metal:257-260)
          const f4 z = {0.0f, 0.0f, 0.0f, 0.0f};
          writeImage(P.images[RT_TEXTURE_DIFFUSE_ALBEDO], px, py, z);
          writeImage(P.images[RT_TEXTURE_SPECULAR_ALBEDO], px, py, z);
          writeImage(P.images[RT_TEXTURE_NORMAL], px, py, z);
          writeImage(P.images[RT_TEXTURE_ROUGHNESS], px, py, z);
        }
      } else {
        const float4 t4 = W.tot[slot];
        total = mk3(t4.x, t4.y, t4.z);
        totalSamples = int(__float_as_uint(t4.w));
        offset = __float_as_uint(W.misc[slot].z);
        if (sampleIndex - 1 < totalSamples) { // finish the previous sample in sample order
          const float4 r4 = W.rad[slot];
          total += mk3(r4.x, r4.y, r4.z);
        }
        if (sampleIndex == 1 && maxExtraSamples > 0) {
          const float4 m4 = W.mot[slot];
          totalSamples = adaptiveSampleCount(U, baseSamples, maxExtraSamples, mk2(m4.x, m4.y), mk2(m4.z, m4.w));
        }
      }
      W.tot[slot] = make_float4(total.x, total.y, total.z, __uint_as_float(uint32_t(totalSamples)));
      if (sampleIndex < totalSamples) {
        const int hIndex = haltonIndex(U, offset, sampleStride, sampleIndex);
        PathState s;
        startPath(U, px, py, hIndex, s);
        W.rayO[slot] = make_float4(s.origin.x, s.origin.y, s.origin.z, 0.0f);
        W.rayD[slot] = make_float4(s.dir.x, s.dir.y, s.dir.z, 0.0f);
        W.thr[slot] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(hIndex));
        W.rad[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        W.ctr[slot] = make_int4(0, 0, 0, 0);
        push = U.maxBounces > 0;
      }
    }
    queuePush(W.queue[0], W.counts + 0, push, slot);
  }
}

template <int kVariant>
__global__ void __launch_bounds__(kBlock) k_wf_trace(const __grid_constant__ TraceParams P, const WfState W, int qin,
                                                     int firstSegment) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { // the queues the next two phases append to start empty
    W.counts[qin ^ 1] = 0u;
    W.counts[2] = 0u;
  }
  const uint32_t count = W.counts[qin];
  const uint32_t *queue = W.queue[qin];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
    const uint32_t slot = queue[j];
    const float4 o = W.rayO[slot], d = W.rayD[slot];
    RayHit hit;
    const bool found = traverseScene<false, kVariant>(P.tlas, o.x, o.y, o.z, d.x, d.y, d.z, 0.0f, INFINITY, hit);
    W.hitA[slot] = make_float4(hit.t, hit.u, hit.v, found ? 1.0f : 0.0f);
    W.hitB[slot] = make_uint4(hit.instance, hit.geometry, hit.primitive, 0u);
    if (firstSegment && P.primaryIds != nullptr) {
      int px, py;
      bool valid;
      slotPixel(P, slot, px, py, valid);
      const uint4 id = found ? make_uint4(hit.instance, hit.geometry, hit.primitive, __float_as_uint(hit.t))
                             : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      reinterpret_cast<uint4 *>(P.primaryIds)[size_t(py) * size_t(P.uniforms.width) + size_t(px)] = id;
    }
  }
  if (P.rayCounters != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.rayCounters + 0, (unsigned long long)count);
}

__global__ void __launch_bounds__(kBlock) k_wf_shade(const __grid_constant__ TraceParams P, const WfState W, int qin,
                                                     int sampleIndex) {
  const uint32_t count = W.counts[qin];
  const uint32_t *queue = W.queue[qin];
  const uint32_t rounds = (count + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint32_t round = 0; round < rounds; ++round, j += gridDim.x * blockDim.x) { // whole warps stay in the loop
    bool pushPath = false, pushShadow = false, isHit = false;
    uint32_t slot = 0;
    if (j < count) {
      slot = queue[j];
      const float4 ha = W.hitA[slot];
      if (ha.w != 0.0f) {
        isHit = true;
        const uint4 hb = W.hitB[slot];
        RayHit hit;
        hit.t = ha.x, hit.u = ha.y, hit.v = ha.z;
        hit.instance = hb.x, hit.geometry = hb.y, hit.primitive = hb.z;
        const float4 o = W.rayO[slot], d = W.rayD[slot], th = W.thr[slot], ra = W.rad[slot];
        const int4 c = W.ctr[slot];
        const float4 m4 = W.mot[slot];
        const float4 mi = W.misc[slot];
        PathState s;
        s.origin = mk3(o.x, o.y, o.z);
        s.dir = mk3(d.x, d.y, d.z);
        s.throughput = mk3(th.x, th.y, th.z);
        s.radiance = mk3(ra.x, ra.y, ra.z);
        s.bounce = c.x, s.step = c.y, s.transparencyPasses = c.z;
        const int hIndex = __float_as_int(th.w);
        PrimaryOutputs prim = emptyPrimaryOutputs();
        const uint32_t flags = __float_as_uint(mi.y);
        prim.depth = mi.x;
        prim.motion = mk2(m4.x, m4.y);
        prim.hadPrimaryHit = (flags & 1u) != 0u;
        prim.wroteGBuffer = (flags & 2u) != 0u;
        const bool primarySegment = (s.bounce == 0 && sampleIndex == 0);
        const bool hadGBuffer = prim.wroteGBuffer;
        ShadowRequest shadow;
        pushPath = shadeSegment(P, s, hit, hIndex, sampleIndex, mk2(m4.z, m4.w), prim, shadow);
        W.rayO[slot] = make_float4(s.origin.x, s.origin.y, s.origin.z, 0.0f);
        W.rayD[slot] = make_float4(s.dir.x, s.dir.y, s.dir.z, 0.0f);
        W.thr[slot] = make_float4(s.throughput.x, s.throughput.y, s.throughput.z, th.w);
        W.rad[slot] = make_float4(s.radiance.x, s.radiance.y, s.radiance.z, 0.0f);
        W.ctr[slot] = make_int4(s.bounce, s.step, s.transparencyPasses, 0);
        if (primarySegment || (prim.wroteGBuffer && !hadGBuffer)) {
          const uint32_t nf = (prim.hadPrimaryHit ? 1u : 0u) | (prim.wroteGBuffer ? 2u : 0u);
          W.mot[slot] = make_float4(prim.motion.x, prim.motion.y, m4.z, m4.w);
          W.misc[slot] = make_float4(prim.depth, __uint_as_float(nf), mi.z, 0.0f);
        }
        if (prim.wroteGBuffer && !hadGBuffer) {
          int px, py;
          bool valid;
          slotPixel(P, slot, px, py, valid);
          writeImage(P.images[RT_TEXTURE_DIFFUSE_ALBEDO], px, py, prim.gDiffuse);
          writeImage(P.images[RT_TEXTURE_SPECULAR_ALBEDO], px, py, prim.gSpecular);
          writeImage(P.images[RT_TEXTURE_NORMAL], px, py, prim.gNormal);
          writeImage(P.images[RT_TEXTURE_ROUGHNESS], px, py, prim.gRoughness);
        }
        if (shadow.valid) {
          pushShadow = true;
          W.shO[slot] = make_float4(shadow.origin.x, shadow.origin.y, shadow.origin.z, shadow.tmax);
          W.shD[slot] = make_float4(shadow.dir.x, shadow.dir.y, shadow.dir.z, 0.0f);
          W.shC[slot] = make_float4(shadow.contribution.x, shadow.contribution.y, shadow.contribution.z, 0.0f);
        }
      }
    }
    queuePush(W.shadowQueue, W.counts + 2, pushShadow, slot);
    queuePush(W.queue[qin ^ 1], W.counts + (qin ^ 1), pushPath, slot);
    if (P.rayCounters != nullptr) {
      const unsigned active = __activemask();
      const unsigned hits = __ballot_sync(active, isHit);
      if ((threadIdx.x & 31) == __ffs(int(active)) - 1 && hits) atomicAdd(P.rayCounters + 2, (unsigned long long)__popc(hits));
    }
  }
}

template <int kVariant>
__global__ void __launch_bounds__(kBlock) k_wf_shadow(const __grid_constant__ TraceParams P, const WfState W) {
  const uint32_t count = W.counts[2];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
    const uint32_t slot = W.shadowQueue[j];
    const float4 o = W.shO[slot], d = W.shD[slot];
    RayHit hit;
    if (!traverseScene<true, kVariant>(P.tlas, o.x, o.y, o.z, d.x, d.y, d.z, 0.0f, o.w, hit)) {
      const float4 c = W.shC[slot];
      float4 r = W.rad[slot];
      r.x = r.x + c.x, r.y = r.y + c.y, r.z = r.z + c.z;
      W.rad[slot] = r;
    }
  }
  if (P.rayCounters != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.rayCounters + 1, (unsigned long long)count);
}

__global__ void __launch_bounds__(kBlock) k_wf_resolve(const __grid_constant__ TraceParams P, const WfState W,
                                                       int sampleLoopBound) {
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < W.capacity; slot += gridDim.x * blockDim.x) {
    int px, py;
    bool valid;
    slotPixel(P, slot, px, py, valid);
    if (!valid) continue;
    const float4 t4 = W.tot[slot];
    f3 total = mk3(t4.x, t4.y, t4.z);
    const int totalSamples = int(__float_as_uint(t4.w));
    if (sampleLoopBound - 1 < totalSamples) {
      const float4 r4 = W.rad[slot];
      total += mk3(r4.x, r4.y, r4.z);
    }
    const float4 m4 = W.mot[slot];
    const float4 mi = W.misc[slot];
    PrimaryOutputs prim = emptyPrimaryOutputs();
    prim.depth = mi.x;
    prim.motion = mk2(m4.x, m4.y);
    // G-buffer images were written by the shade phase; resolvePixel must not overwrite them
    resolvePixel(P, px, py, total, totalSamples, mk2(m4.z, m4.w), prim, false);
  }
}

int ensureState(rt_context *ctx, uint32_t capacity, WfState &out) {
  const size_t need = 256 + 13 * (size_t(capacity) * 16 + 256) + 3 * (size_t(capacity) * 4 + 256);
  if (ctx->wfState == nullptr || ctx->wfBytes < need) {
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->wfState) cudaFree(ctx->wfState);
    ctx->wfState = nullptr;
    RT_CUDA(cudaMalloc(&ctx->wfState, need));
    ctx->wfBytes = need;
  }
  uint8_t *p = static_cast<uint8_t *>(ctx->wfState);
  auto take = [&](size_t bytes) {
    void *r = p;
    p += (bytes + 255) & ~size_t(255);
    return r;
  };
  WfState s{};
  s.capacity = capacity;
  const size_t v = size_t(capacity) * 16;
  s.counts = static_cast<uint32_t *>(take(256));
  s.rayO = static_cast<float4 *>(take(v));
  s.rayD = static_cast<float4 *>(take(v));
  s.thr = static_cast<float4 *>(take(v));
  s.rad = static_cast<float4 *>(take(v));
  s.ctr = static_cast<int4 *>(take(v));
  s.tot = static_cast<float4 *>(take(v));
  s.mot = static_cast<float4 *>(take(v));
  s.misc = static_cast<float4 *>(take(v));
  s.hitA = static_cast<float4 *>(take(v));
  s.hitB = static_cast<uint4 *>(take(v));
  s.shO = static_cast<float4 *>(take(v));
  s.shD = static_cast<float4 *>(take(v));
  s.shC = static_cast<float4 *>(take(v));
  s.queue[0] = static_cast<uint32_t *>(take(size_t(capacity) * 4));
  s.queue[1] = static_cast<uint32_t *>(take(size_t(capacity) * 4));
  s.shadowQueue = static_cast<uint32_t *>(take(size_t(capacity) * 4));
  RT_CHECK(size_t(p - static_cast<uint8_t *>(ctx->wfState)) <= ctx->wfBytes, "internal: wavefront state overflow");
  out = s;
  return 0;
}

} // namespace

int launchTraceWavefront(rt_context *ctx, const TraceParams &P) {
  const rt_uniforms &U = P.uniforms;
  const int tileCount = P.tilesX * P.tilesY;
  const int owned = (tileCount - P.tileRemainder + P.tileModulo - 1) / P.tileModulo;
  WfState W;
  RT_TRY(ensureState(ctx, uint32_t(owned) * 256u, W));
  cudaStream_t st = ctx->stream;
  const int baseSamples = std::max(U.samplesPerPixel, 1);
  const int maxExtraSamples = (U.enableMotionAdaptiveSampling != 0) ? std::max(U.motionSamplingMaxExtraSamples, 0) : 0;
  const int sampleLoopBound = baseSamples + maxExtraSamples;
  const int maxBounces = std::max(U.maxBounces, 0);
  // a refraction does not consume a bounce until transparencyPasses > maxBounces (Raytracing.metal:563-575)
  const int maxSegments = maxBounces * (maxBounces + 1);
  const int slotBlocks = (int(W.capacity) + kBlock - 1) / kBlock;
  const int persistent = std::min(slotBlocks, ctx->smCount * std::max(1, ctx->blocksPerSm));
  for (int s = 0; s < sampleLoopBound; ++s) {
    RT_CUDA(cudaMemsetAsync(W.counts, 0, 16, st));
    k_wf_generate<<<persistent, kBlock, 0, st>>>(P, W, s, baseSamples, maxExtraSamples);
    ++ctx->launches;
    int qin = 0;
    for (int segment = 0; segment < maxSegments; ++segment) {
      if (segment >= maxBounces) { // only glass paths get here: ask the device whether any are left
        uint32_t remaining = 0;
        RT_CUDA(cudaMemcpyAsync(&remaining, W.counts + qin, 4, cudaMemcpyDeviceToHost, st));
        RT_CUDA(cudaStreamSynchronize(st));
        if (remaining == 0) break;
      }
      const int first = (s == 0 && segment == 0) ? 1 : 0;
      switch (ctx->traversalVariant) {
        case 1: k_wf_trace<1><<<persistent, kBlock, 0, st>>>(P, W, qin, first); break;
        case 2: k_wf_trace<2><<<persistent, kBlock, 0, st>>>(P, W, qin, first); break;
        default: k_wf_trace<0><<<persistent, kBlock, 0, st>>>(P, W, qin, first); break;
      }
      k_wf_shade<<<persistent, kBlock, 0, st>>>(P, W, qin, s);
      switch (ctx->traversalVariant) {
        case 1: k_wf_shadow<1><<<persistent, kBlock, 0, st>>>(P, W); break;
        case 2: k_wf_shadow<2><<<persistent, kBlock, 0, st>>>(P, W); break;
        default: k_wf_shadow<0><<<persistent, kBlock, 0, st>>>(P, W); break;
      }
      ctx->launches += 3;
      qin ^= 1;
    }
  }
  k_wf_resolve<<<persistent, kBlock, 0, st>>>(P, W, sampleLoopBound);
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

} // namespace rtb
