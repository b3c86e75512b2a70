// trace_wavefront.cu — path-tracing kernels for sm_100a, wavefront layout (trace mode 1, the default).
//
// Same contract as trace.cu (the reference's raytracingKernel, MetalRaytracing/Raytracing.metal:220-831, behind
// its argument table) but split by phase so that every kernel runs converged and small:
//
//   for each batch of sample indices [s0, s0 + n) (the reference's sample loop, :269; up to 16 samples of every
//   owned pixel are in flight at once so that each launch has enough rays to amortise its tail):
//     generate      fold the previous batch into the pixel sums in sample order, start the batch: camera rays ->
//                   path state, queue
//     repeat for each path segment k (the reference's bounce loop, :311):
//       traverse    persistent software traversal (traverse.cuh, 80 registers): closest hit for every queued path
//                   of segment k, then — in the same launch — any-hit for the shadow requests of segment k - 1;
//                   unoccluded shadow rays add their contribution
//       shade       one shadeSegment() per hit: emission, light sample, BRDF, next ray; emits a shadow
//                   request and re-queues surviving paths (warp-aggregated queue appends = ray compaction)
//     traverse      the shadow requests of the last segment
//   resolve         fold the last batch, sample mean, EMA with history, image writes (+ NVLink peer stores)
//
// Paths that miss or die leave the queue, so later segments run on dense warps instead of the megakernel's
// 11-of-32 active threads (profiles/). Per-pixel results do not depend on queue order, and the arithmetic is the
// shared path_step.cuh, so this layout is bit-identical to the megakernel and to the CPU oracle.
// All kernels are persistent-style: a fixed grid of (SM count x resident CTAs) strides over the queue, whose
// length is read from device memory, so no host round trip is needed between phases.
#include <cub/device/device_radix_sort.cuh>

#include <cstring>

#include "path_step.cuh"

// This file is compiled twice (csrc/Makefile): with -DRT_TU_TRAVERSE into the traversal kernel + the host launcher, and
// with -DRT_TU_SHADE into the generate / shade / resolve kernels. The split lets the two halves take different compiler
// flags: the traversal always keeps -fmad=false (ids are part of the bit-exact contract), the shading half can be built
// with FMA contraction and float sincosf / atan2f / acosf (-DRT_FAST_SHADE, `make fast`, librt_b200_fast.so) to measure
// what the numeric contract costs. Without either macro everything lands in one object (tools/build_variants.sh).
#if !defined(RT_TU_TRAVERSE) && !defined(RT_TU_SHADE)
#define RT_TU_TRAVERSE 1
#define RT_TU_SHADE 1
#endif

namespace rtb {

struct WfState;
// launch wrappers of the shading half (defined under RT_TU_SHADE)
void wfLaunchGenerate(int grid, cudaStream_t st, const TraceParams &P, const WfState &W, int s0, int n, int prevS0, int prevN,
                      int baseSamples, int maxExtraSamples);
void wfLaunchShade(bool textures, bool plain, int grid, cudaStream_t st, const TraceParams &P, const WfState &W, int qin,
                   int s0, int cameraRays, int parity);
void wfLaunchResolve(int grid, cudaStream_t st, const TraceParams &P, const WfState &W, int lastS0, int lastN);
int wfShadeIsFast(); // 1 when the shading half was built with RT_FAST_SHADE

constexpr int kBlock = 256;
// traversal kernels: small CTAs free their registers as soon as their rays finish; min-blocks caps registers
#ifndef RT_TRACE_BLOCK
#define RT_TRACE_BLOCK 128
#endif
// Two builds of every traversal kernel: 7 resident CTAs per SM (72 registers, no spills) and 8 (64 registers, 36 bytes of
// spills). Eight are 1 - 2 % faster on launches of tens of millions of rays, seven on small ones (more resident warps
// lengthen the tail of a persistent launch): launchTraceWavefront picks by the size of the dispatch. RT_TRACE_MINBLOCKS
// forces one value (tools/build_variants.sh). With the ray and the hit ids behind the stack (traverse.cuh, RT_SLIM_LANE)
// the kernel needs 72 registers where it needed 80 + spills for 6 CTAs (profiles/r2_experiments.md section 9).
#ifdef RT_TRACE_MINBLOCKS
constexpr int kMinBlocksSmall = RT_TRACE_MINBLOCKS, kMinBlocksLarge = RT_TRACE_MINBLOCKS;
#else
constexpr int kMinBlocksSmall = 7, kMinBlocksLarge = 8;
#endif
// RT_PREFETCH > 0 (an experiment, measured slower: profiles/r2_experiments.md section 10): entries of a per-warp ring in
// shared memory into which the next rays are copied by cp.async ahead of time (traceQueuePrefetch); RT_PREFETCH_LARGE /
// RT_PREFETCH_SMALL say which of the two builds get it
#ifndef RT_PREFETCH
#define RT_PREFETCH 0
#endif
#ifndef RT_PREFETCH_SMALL
#define RT_PREFETCH_SMALL 0
#endif
#ifndef RT_PREFETCH_LARGE
#define RT_PREFETCH_LARGE 1
#endif
constexpr bool kPrefetchLarge = RT_PREFETCH > 0 && RT_PREFETCH_LARGE != 0, kPrefetchSmall = RT_PREFETCH > 0 && RT_PREFETCH_SMALL != 0;
constexpr size_t kLargeDispatchPaths = size_t(6) << 20; // path slots per pipeline lane from which the 8-CTA build is used
constexpr int kTraceBlock = RT_TRACE_BLOCK;
// streaming (evict-first) access for the per-path state so it does not push the BVH out of L2
#ifdef RT_NO_STREAMING_HINTS
#define RT_LDS(ptr) (*(ptr))
#define RT_STS(ptr, v) (*(ptr) = (v))
#else
#define RT_LDS(ptr) __ldcs(ptr)
#define RT_STS(ptr, v) __stcs(ptr, v)
#endif

// Path slot = b * capacity + pixelSlot for sample b of the batch in flight: sample-major, so a warp (32 pixels of
// one tile, same sample) touches consecutive state records.
struct WfState { // same definition in both translation units
  uint32_t capacity; // pixel slots = owned tiles * 256
  uint32_t batch;    // samples of a pixel in flight at once; path slots = capacity * batch
  uint32_t queueCapacity; // entries of each queue array (= path slots): class B entries fill it from the back
  uint32_t classify;      // 1: rays are queued by class (flat TLAS): A = reaches a BLAS with nodes, from the front of the
                          // queue; B = cheap (misses, single-leaf instances only), from the back. 0: everything is class A
  // per path slot
  float4 *rayO, *rayD; // rayD.w: sample (halton) index bits on camera rays, packed (bounce, step, transparency
                       // passes) afterwards; rayO doubles as the origin of the segment's shadow ray
  float4 *thr;  // throughput.xyz, w = halton index bits
  float4 *rad;  // radiance.xyz
  float4 *hitA; // t (+inf = miss), u, v, primitive bits
  uint2 *hitB;  // instance, geometry (written for hits only)
  float4 *shD, *shC;   // shadow ray direction + tmax; the path's radiance + the light sample's contribution = what rad
                       // becomes when the shadow ray is unoccluded (its origin is rayO)
  // per pixel slot
  float4 *tot;  // totalColor.xyz, w = totalSamples bits
  float4 *mot;  // motion.xy, prevMotion.xy
  float4 *misc; // depth, flags bits (1 = hadPrimaryHit, 2 = wroteGBuffer), seed offset bits, unused
  uint32_t *queue[2], *shadowQueue;
  uint32_t *shadeQueue[2]; // classify: the paths of queue[q] once more in plain append order — what the shade kernel walks
                           // (consecutive entries = neighbouring path slots, so its state accesses stay coalesced; walking
                           // the class-ordered queue cost it 33 %, profiles/r2_experiments.md section 6)
  uint32_t *counts; // [pathCount(q)] path queues, [3] trace cursor, [shadowCount(p)] shadow queue length and [10 + p] shadow cursor of
                    // segment parity p (double-buffered: the shadow rays of segment k are traced in the same launch
                    // as the closest-hit rays of segment k + 1)
  // ray reordering (option sort_rays): key buffers, the sorted copy of a queue and CUB's workspace
  uint32_t *sortKeys[2], *sortedQueue;
  void *sortTemp;
  size_t sortTempBytes;
};

namespace {

__device__ __forceinline__ void slotPixel(const TraceParams &P, uint32_t slot, int &px, int &py, bool &valid) {
  valid = ownedPixel(P, int(slot >> 8), int(slot & 255u), px, py);
}

// Where the queue lengths live in WfState::counts. The shade kernel of a segment appends to path queue qin ^ 1 and to
// the shadow queue of parity qin, so those two lengths share an aligned 64-bit word and one atomic serves both.
__host__ __device__ __forceinline__ int pathCount(int q) { return q == 1 ? 16 : 18; }
__host__ __device__ __forceinline__ int shadowCount(int p) { return p == 0 ? 17 : 19; }
// the same for the class B ends of the queues (entries stored from the back of the array), eight words further on
__host__ __device__ __forceinline__ int pathCountB(int q) { return pathCount(q) + 8; }
__host__ __device__ __forceinline__ int plainCount(int q) { return q == 1 ? 20 : 21; } // length of shadeQueue[q]
__host__ __device__ __forceinline__ int shadowCountB(int p) { return shadowCount(p) + 8; }
// entry j of a two-ended queue holding countA class A entries at the front and the class B entries at the back
__device__ __forceinline__ uint32_t queueEntry(const uint32_t *__restrict__ queue, uint32_t capacity, uint32_t countA, uint32_t j) {
  return j < countA ? queue[j] : queue[capacity - 1u - (j - countA)];
}

// Appends `slot` to the path queue (lanes with pushPath) and to the shadow queue (lanes with pushShadow) with one
// 64-bit atomic per warp: low word = path queue length, high word = shadow queue length (`pair` points at both).
__device__ __forceinline__ void queuePushBoth(uint32_t *pathQueue, uint32_t *shadowQueue, uint32_t *pair, bool pushPath,
                                              bool pushShadow, uint32_t slot) {
  const unsigned active = __activemask();
  const unsigned vp = __ballot_sync(active, pushPath), vs = __ballot_sync(active, pushShadow);
  if ((vp | vs) == 0u) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(int(active)) - 1;
  unsigned long long base = 0ull;
  if (lane == leader)
    base = atomicAdd(reinterpret_cast<unsigned long long *>(pair),
                     (unsigned long long)__popc(vp) | ((unsigned long long)__popc(vs) << 32));
  base = __shfl_sync(active, base, leader);
  const unsigned below = (1u << lane) - 1u;
  if (pushPath) pathQueue[uint32_t(base) + __popc(vp & below)] = slot;
  if (pushShadow) shadowQueue[uint32_t(base >> 32) + __popc(vs & below)] = slot;
}

// The same with classes: class A entries go to the front of the queues (lengths in pairA), class B entries to the back
// (lengths in pairB); one 64-bit atomic per class that has entries.
__device__ __forceinline__ void queuePushClassified(uint32_t *pathQueue, uint32_t *shadowQueue, uint32_t *pairA, uint32_t *pairB,
                                                    uint32_t capacity, bool pushPath, bool pathIsB, bool pushShadow,
                                                    bool shadowIsB, uint32_t slot, uint32_t *plainQueue, uint32_t *plainLength) {
  // called by whole warps. The three reservations (class A pair, class B pair, plain path order) are issued by three
  // different lanes so that they are in flight together: the shade kernel is latency-bound and three atomics in a row
  // cost it a fifth of its time.
  const unsigned full = 0xFFFFFFFFu;
  const unsigned vp = __ballot_sync(full, pushPath), vs = __ballot_sync(full, pushShadow);
  if ((vp | vs) == 0u) return;
  const unsigned vpB = __ballot_sync(full, pushPath && pathIsB), vsB = __ballot_sync(full, pushShadow && shadowIsB);
  const unsigned vpA = vp & ~vpB, vsA = vs & ~vsB;
  const int lane = threadIdx.x & 31;
  unsigned long long got = 0ull;
  if (lane == 0 && (vpA | vsA) != 0u)
    got = atomicAdd(reinterpret_cast<unsigned long long *>(pairA),
                    (unsigned long long)__popc(vpA) | ((unsigned long long)__popc(vsA) << 32));
  if (lane == 1 && (vpB | vsB) != 0u)
    got = atomicAdd(reinterpret_cast<unsigned long long *>(pairB),
                    (unsigned long long)__popc(vpB) | ((unsigned long long)__popc(vsB) << 32));
  if (lane == 2 && vp != 0u) got = (unsigned long long)atomicAdd(plainLength, uint32_t(__popc(vp)));
  const unsigned long long baseA = __shfl_sync(full, got, 0), baseB = __shfl_sync(full, got, 1);
  const uint32_t basePlain = uint32_t(__shfl_sync(full, got, 2));
  const unsigned below = (1u << lane) - 1u;
  if (pushPath) {
    plainQueue[basePlain + __popc(vp & below)] = slot; // the next shade pass walks this one
    if (!pathIsB) pathQueue[uint32_t(baseA) + __popc(vpA & below)] = slot;
    else pathQueue[capacity - 1u - (uint32_t(baseB) + __popc(vpB & below))] = slot;
  }
  if (pushShadow) {
    if (!shadowIsB) shadowQueue[uint32_t(baseA >> 32) + __popc(vsA & below)] = slot;
    else shadowQueue[capacity - 1u - (uint32_t(baseB >> 32) + __popc(vsB & below))] = slot;
  }
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

// Folds the radiances of the batch [s0, s0 + n) into the pixel sum, in sample order (float addition order is part of
// the result), exactly where the reference does `totalColor += accumulatedColor` (Raytracing.metal:776-777).
__device__ __forceinline__ void foldBatch(const WfState &W, uint32_t pixelSlot, int s0, int n, int totalSamples, f3 &total) {
  for (int b = 0; b < n; ++b) {
    if (s0 + b < totalSamples) {
      const float4 r4 = RT_LDS(W.rad + size_t(b) * W.capacity + pixelSlot);
      total += mk3(r4.x, r4.y, r4.z);
    }
  }
}

#ifdef RT_TU_SHADE
// Starts the samples [s0, s0 + n) of every owned pixel after folding the previous batch [prevS0, prevS0 + prevN).
// The first batch never extends past baseSamples, so whether one of its samples exists does not depend on the
// motion-adaptive count, which is evaluated right after sample 0 has been folded (Raytracing.metal:779-789).
template <bool kClassify>
__global__ void __launch_bounds__(kBlock) k_wf_generate(const __grid_constant__ TraceParams P, const WfState W, int s0,
                                                        int n, int prevS0, int prevN, int baseSamples,
                                                        int maxExtraSamples) {
  const rt_uniforms &U = P.uniforms;
  const int sampleStride = baseSamples + maxExtraSamples;
  for (uint32_t pixelSlot = blockIdx.x * blockDim.x + threadIdx.x; pixelSlot < W.capacity;
       pixelSlot += gridDim.x * blockDim.x) { // whole warps stay in the loop (capacity is a multiple of 256)
    int px, py;
    bool valid;
    slotPixel(P, pixelSlot, px, py, valid);
    int totalSamples = 0;
    uint32_t offset = 0;
    if (valid) {
      const size_t pixelIndex = size_t(py) * size_t(U.width) + size_t(px);
      f3 total;
      if (s0 == 0) {
        offset = reinterpret_cast<const uint32_t *>(P.images[RT_TEXTURE_RANDOM].data)[pixelIndex];
        const f4 pm = readImage(P.images[RT_TEXTURE_MOTION], px, py);
        total = mk3(0.0f);
        totalSamples = baseSamples;
        RT_STS(W.mot + pixelSlot, make_float4(0.0f, 0.0f, pm.x, pm.y));
        RT_STS(W.misc + pixelSlot, make_float4(1.0e8f, __uint_as_float(0u), __uint_as_float(offset), 0.0f));
        if (U.enableDenoiseGBuffer != 0) { // a pixel whose first segment misses keeps zeros (Raytracing.metal:257-260)
          const f4 z = {0.0f, 0.0f, 0.0f, 0.0f};
          writeImage(P.images[RT_TEXTURE_DIFFUSE_ALBEDO], px, py, z);
          writeImage(P.images[RT_TEXTURE_SPECULAR_ALBEDO], px, py, z);
          writeImage(P.images[RT_TEXTURE_NORMAL], px, py, z);
          writeImage(P.images[RT_TEXTURE_ROUGHNESS], px, py, z);
        }
      } else {
        const float4 t4 = RT_LDS(W.tot + pixelSlot);
        total = mk3(t4.x, t4.y, t4.z);
        totalSamples = int(__float_as_uint(t4.w));
        offset = __float_as_uint(RT_LDS(W.misc + pixelSlot).z);
        foldBatch(W, pixelSlot, prevS0, prevN, totalSamples, total);
        if (prevS0 == 0 && maxExtraSamples > 0) {
          const float4 m4 = RT_LDS(W.mot + pixelSlot);
          totalSamples = adaptiveSampleCount(U, baseSamples, maxExtraSamples, mk2(m4.x, m4.y), mk2(m4.z, m4.w));
        }
      }
      RT_STS(W.tot + pixelSlot, make_float4(total.x, total.y, total.z, __uint_as_float(uint32_t(totalSamples))));
    }
    unsigned long long pushed = 0ull; // bit b: sample b of this pixel has a camera ray to trace (batch <= 64)
    unsigned long long cheap = 0ull;  // bit b: that ray is class B (W.classify: it reaches no BLAS with nodes)
    for (int b = 0; b < n; ++b) {
      const int sampleIndex = s0 + b;
      const uint32_t slot = uint32_t(b) * W.capacity + pixelSlot;
      if (valid && sampleIndex < totalSamples && sampleIndex % P.sampleModulo != P.sampleRemainder) {
        RT_STS(W.rad + slot, make_float4(0.0f, 0.0f, 0.0f, 0.0f)); // sample partition: another dispatch owns it; folds as 0
      } else if (valid && sampleIndex < totalSamples) {
        const int hIndex = haltonIndex(U, offset, sampleStride, sampleIndex);
        PathState s;
        startPath(U, px, py, hIndex, s);
        // a camera ray's other state is implied (origin = camera, throughput 1, radiance 0, counters 0): the first
        // segment's trace and shade kernels supply it themselves, so only 16 of the 80 bytes are written here
        RT_STS(W.rayD + slot, make_float4(s.dir.x, s.dir.y, s.dir.z, __int_as_float(hIndex)));
        if (U.maxBounces > 0) {
          pushed |= 1ull << b;
          if (kClassify && !rayReachesNodes(P.nodeUnionBox, s.origin.x, s.origin.y, s.origin.z, s.dir.x, s.dir.y, s.dir.z, INFINITY))
            cheap |= 1ull << b;
        } else {
          RT_STS(W.rad + slot, make_float4(0.0f, 0.0f, 0.0f, 0.0f)); // never traced: folds as black
        }
      }
    }
    // one reservation per warp and class for all of its samples (instead of n atomics); entries stay sample-major, so 32
    // consecutive queue entries of a class are still up to 32 consecutive path slots
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    // per warp: how many rays of each kind, one reservation each — issued by different lanes so they overlap
    uint32_t totalA = 0, totalB = 0;
    for (int b = 0; b < n; ++b) {
      totalA += uint32_t(__popc(__ballot_sync(full, ((pushed & ~cheap) >> b) & 1ull)));
      totalB += uint32_t(__popc(__ballot_sync(full, ((pushed & cheap) >> b) & 1ull)));
    }
    uint32_t got = 0;
    if (lane == 0 && totalA != 0u) got = atomicAdd(W.counts + pathCount(0), totalA);
    if (kClassify && lane == 1 && totalB != 0u) got = atomicAdd(W.counts + pathCountB(0), totalB);
    if (kClassify && lane == 2 && totalA + totalB != 0u) got = atomicAdd(W.counts + plainCount(0), totalA + totalB);
    uint32_t runA = __shfl_sync(full, got, 0), runB = __shfl_sync(full, got, 1), runPlain = __shfl_sync(full, got, 2);
    if (totalA + totalB != 0u) {
      for (int b = 0; b < n; ++b) { // entries stay sample-major: consecutive entries of a kind are up to 32 consecutive slots
        const bool push = (pushed >> b) & 1ull, isB = (cheap >> b) & 1ull;
        const unsigned votes = __ballot_sync(full, push), votesB = __ballot_sync(full, push && isB);
        const unsigned votesA = votes & ~votesB, below = (1u << lane) - 1u;
        const uint32_t slot = uint32_t(b) * W.capacity + pixelSlot;
        if (push && !isB) W.queue[0][runA + uint32_t(__popc(votesA & below))] = slot;
        if (push && isB) W.queue[0][W.queueCapacity - 1u - (runB + uint32_t(__popc(votesB & below)))] = slot;
        if (kClassify && push) W.shadeQueue[0][runPlain + uint32_t(__popc(votes & below))] = slot;
        runA += uint32_t(__popc(votesA)), runB += uint32_t(__popc(votesB)), runPlain += uint32_t(__popc(votes));
      }
    }
  }
}
#endif // RT_TU_SHADE

#ifdef RT_TU_TRAVERSE
// Ray reordering (option sort_rays, off by default): key = direction octant (3 bits) | Morton code of the origin in
// a 128^3 grid over the TLAS bounds (21 bits). Rays of one warp then start in the same region and walk the tree in
// the same child order. Results do not depend on queue order, so this only moves time between kernels.
__device__ __forceinline__ uint32_t spread7(uint32_t v) { // 7 bits -> every third bit
  v &= 0x7Fu;
  v = (v | (v << 8)) & 0x0000700Fu;
  v = (v | (v << 4)) & 0x000430C3u;
  v = (v | (v << 2)) & 0x00049249u;
  return v;
}
__global__ void __launch_bounds__(kBlock) k_ray_keys(const uint32_t *__restrict__ queue, uint32_t count,
                                                     const float4 *__restrict__ rayO, const float4 *__restrict__ rayD,
                                                     const float4 *__restrict__ rootBox, uint32_t *__restrict__ keys) {
  const float4 lo = rootBox[0], hi = rootBox[1];
  const float sx = 128.0f / fmaxf(hi.x - lo.x, 1e-20f), sy = 128.0f / fmaxf(hi.y - lo.y, 1e-20f),
              sz = 128.0f / fmaxf(hi.z - lo.z, 1e-20f);
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
    const uint32_t slot = queue[j];
    const float4 o = RT_LDS(rayO + slot), d = RT_LDS(rayD + slot);
    const uint32_t cx = uint32_t(fminf(fmaxf((o.x - lo.x) * sx, 0.0f), 127.0f));
    const uint32_t cy = uint32_t(fminf(fmaxf((o.y - lo.y) * sy, 0.0f), 127.0f));
    const uint32_t cz = uint32_t(fminf(fmaxf((o.z - lo.z) * sz, 0.0f), 127.0f));
    const uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    keys[j] = (oct << 21) | (spread7(cx) << 2) | (spread7(cy) << 1) | spread7(cz);
  }
}

// The two traversal kernels are persistent: a fixed grid of resident CTAs pulls rays from a device-side cursor.
// A warp claims rays for its idle lanes whenever at least kRefill of them are idle (or all are), so lanes whose
// rays ended early are refilled instead of idling until the warp's longest ray finishes; ballot + popc give each
// idle lane its rank in the claimed range (warp-level ray compaction). kRefill == 0 claims 32 rays at a time only
// when the whole warp is idle.
#ifndef RT_STEPS_PER_CHECK
#define RT_STEPS_PER_CHECK 2
#endif
constexpr int kStepsPerCheck = RT_STEPS_PER_CHECK;
// RT_FUSED_PRIMS > 0: fused pop -> node -> primitive iterations (LaneTraversal::stepFused) with that many primitive
// tests per iteration; 0: one unit of work per iteration (LaneTraversal::step)
#ifndef RT_FUSED_PRIMS
#define RT_FUSED_PRIMS 1
#endif
// RT_CONVERGED: LaneTraversal::stepConverged; bit 0 = an entry stage before the node stage, bit 2 = one after it,
// bit 1 = early finish, bit 3 = entry and triangle share one stage as in stepFused, bits 4-5 = extra triangle stages
// Default 22 + per-instantiation entry order: early finish, two triangle stages, and the entry stage after the node stage
// for a real TLAS (the best fixed order, tune34-38) but before it in the kernels instantiated for a flat TLAS
// (<= 8 instances, traverse.cuh begin()): profiles/r2_experiments.md section 5.
#ifndef RT_CONVERGED
#define RT_CONVERGED 22
#endif
// RT_COOP_TRIS: 1 = the triangle stages of stepConverged are done by the warp together, 32 pending (lane, triangle) pairs
// per pass (LaneTraversal::triangleStageCoop); 2 = only when some lane has more than one triangle pending; 0: every lane
// tests its own next triangle
#ifndef RT_COOP_TRIS
#define RT_COOP_TRIS 0
#endif
// the closest-hit switch of the kernels built for a real TLAS (more than 8 instances); follows RT_COOP_TRIS unless set
#ifndef RT_COOP_TRIS_REAL
#define RT_COOP_TRIS_REAL RT_COOP_TRIS
#endif
// The same switch for the any-hit (shadow) rays, where the cooperative pass only has to hand back one bit per owner and
// testing every pending triangle at once finds an occluder sooner. Default 2: K3 -0.6 %, K4 -1.3 %, the 1-spp frame and a
// rank's slice of an 8-GPU frame -2 ... -2.7 %; for closest hits the same stage costs K3 2 % (experiment log 9a, 9c).
#ifndef RT_COOP_ANY
#define RT_COOP_ANY 2
#endif
// entries of each lane's traversal stack kept in shared memory (0 = all in local memory), traverse.cuh SplitStack
#ifndef RT_SHARED_STACK
#define RT_SHARED_STACK 0
#endif

#ifdef RT_COUNT_WORK
// counter build: rayCounters[3..5] (closest-hit rays) and [6..8] (any-hit rays) += node steps, triangle tests, entries
template <bool kAny>
__device__ __forceinline__ void countWork(const TraceParams &P, const LaneTraversal<kAny> &t) {
  if (P.rayCounters == nullptr) return;
  unsigned long long *c = P.rayCounters + (kAny ? 6 : 3);
  atomicAdd(c + 0, (unsigned long long)t.nNodes);
  atomicAdd(c + 1, (unsigned long long)t.nTris);
  atomicAdd(c + 2, (unsigned long long)t.nEntries);
  // iterations this ray lived for: histogram over <= 4, 8, 16, 32, 64, 128, 256, more (rayCounters[9..16]) and the maximum
  const uint32_t it = t.nIters;
  const int bin = it <= 4 ? 0 : it <= 8 ? 1 : it <= 16 ? 2 : it <= 32 ? 3 : it <= 64 ? 4 : it <= 128 ? 5 : it <= 256 ? 6 : 7;
  atomicAdd(P.rayCounters + 9 + bin, 1ull);
  atomicMax(P.rayCounters + 17, (unsigned long long)it);
  atomicAdd(P.rayCounters + 18, (unsigned long long)it);
}
#endif

template <bool kAny, int kRefill, bool kFlat, typename Finish>
__device__ __forceinline__ void traceQueue(const TraceParams &P, const uint32_t *__restrict__ queue, uint32_t countA,
                                           uint32_t countB, uint32_t queueCapacity, uint32_t *cursor,
                                           const float4 *__restrict__ rayO,
                                           const float4 *__restrict__ rayD, bool cameraRays, uint2 *sharedStack,
                                           uint32_t *warpPairs, Finish finish) {
  const unsigned full = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  LaneTraversal<kAny> t;
#if RT_SHARED_STACK > 0
  SplitStack<RT_SHARED_STACK, kTraceBlock> stack;
  stack.shared = sharedStack + threadIdx.x;
#else
  LocalStack stack;
  (void)sharedStack;
#endif
  bool active = false, exhausted = false;
#if !RT_SLIM_LANE
  uint32_t slot = 0;
#endif
  const uint32_t count = countA + countB; // the cursor walks the class A entries (front) first, then class B (back)
#ifdef RT_COUNT_WORK
  uint32_t tailIters = 0; // warp iterations after the queue ran dry
#endif
  while (true) {
    const unsigned idle = __ballot_sync(full, !active);
    if (idle == full && exhausted) break;
    if (!exhausted && (idle == full || (kRefill > 0 && __popc(idle) >= kRefill))) {
      const uint32_t want = uint32_t(__popc(idle));
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(cursor, want);
      base = __shfl_sync(full, base, 0);
      if (base + want >= count) exhausted = true;
      if (!active) {
        const uint32_t j = base + uint32_t(__popc(idle & ((1u << lane) - 1u)));
        if (j < count) {
          const uint32_t newSlot = queueEntry(queue, queueCapacity, countA, j);
#if RT_SLIM_LANE
          stack.set(kSlotPath, make_uint2(newSlot, 0u)); // read again when the ray has finished
#else
          slot = newSlot;
#endif
          const float4 d = RT_LDS(rayD + newSlot);
          float4 o;
          if (!kAny && cameraRays) // first segment: every ray starts at the camera (k_wf_generate)
            o = make_float4(P.uniforms.camera.position.x, P.uniforms.camera.position.y, P.uniforms.camera.position.z, 0.0f);
          else
            o = RT_LDS(rayO + newSlot);
          t.template begin<kFlat>(P.tlas, stack, o.x, o.y, o.z, d.x, d.y, d.z, 0.0f, kAny ? d.w : INFINITY);
          active = true;
        }
      }
      if (base >= count && idle == full) break;
    }
#ifdef RT_COUNT_WORK
    if (exhausted) tailIters += uint32_t(kStepsPerCheck);
#endif
#pragma unroll 1
    for (int k = 0; k < kStepsPerCheck; ++k) {
#if RT_CONVERGED > 0
      // flat TLAS: entry stage before the node stage only; otherwise as RT_CONVERGED says
      if (t.template stepConverged<kFlat || (RT_CONVERGED & 1) != 0, !kFlat && (RT_CONVERGED & 4) != 0, (RT_CONVERGED & 2) != 0, (RT_CONVERGED & 8) != 0, 1 + ((RT_CONVERGED >> 4) & 3), (kAny ? RT_COOP_ANY : (kFlat ? RT_COOP_TRIS : RT_COOP_TRIS_REAL))>(P.tlas, stack, active, warpPairs)) {
#elif RT_FUSED_PRIMS > 0
      if (active && !t.template stepFused<RT_FUSED_PRIMS>(P.tlas, stack)) {
#else
      if (active && !t.step(P.tlas, stack)) {
#endif
#if RT_SLIM_LANE
        const uint32_t slot = stack.get(kSlotPath).x;
#endif
        if constexpr (kAny) finish(slot, t.found, RayHit{}, nullptr);
        else finish(slot, t.found, t.result(stack), nullptr);
#ifdef RT_COUNT_WORK
        countWork<kAny>(P, t);
#endif
        active = false;
      }
    }
  }
#ifdef RT_COUNT_WORK
  if (P.rayCounters != nullptr && lane == 0) { // [19] sum, [20] max of warp iterations spent after the queue ran dry, [21] warps
    atomicAdd(P.rayCounters + 19, (unsigned long long)tailIters);
    atomicMax(P.rayCounters + 20, (unsigned long long)tailIters);
    atomicAdd(P.rayCounters + 21, 1ull);
  }
#endif
}

// RT_PREFETCH > 0 — the experiment of profiles/r2_experiments.md section 10, measured 15 - 44 % slower and off by default.
// The classic loop above fetches rays when lanes go idle: an atomic on the cursor, the queue entry, then the ray's records
// — three dependent trips to L2 / DRAM during which the whole (converged) warp waits; ncu's source view charges a sixth
// of the traversal kernel's stall samples to them. Here the trips are taken ahead of time, one per iteration of the
// warp, and without registers: the cursor is advanced by a chunk whose result is picked up an iteration later, the
// chunk's queue entries are copied into the ring half a ring at a time, and the rays they name — origin, direction and,
// for shadow rays, the radiance record to write when unoccluded — follow, global -> shared by cp.async (LDGSTS, L2
// evict-first), while the warp traverses. A lane that goes idle takes its next ray from shared memory. What it costs:
// 6.6 KB of shared memory per CTA taken from the L1 that caches the BVH, a supply of half a ring per two iterations
// (rays that live for two or three iterations drain it faster), and bookkeeping in every iteration.
#if RT_PREFETCH > 0
struct __align__(16) PrefetchRing {
  float4 o[RT_PREFETCH], d[RT_PREFETCH], c[RT_PREFETCH];
  uint32_t slot[RT_PREFETCH];
};
__device__ __forceinline__ void cpAsync16(void *sharedDst, const void *globalSrc) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(uint32_t(__cvta_generic_to_shared(sharedDst))),
               "l"(globalSrc), "l"(policy)
               : "memory");
}

template <bool kAny, int kRefill, bool kFlat, typename Finish>
__device__ __forceinline__ void traceQueuePrefetch(const TraceParams &P, const uint32_t *__restrict__ queue, uint32_t countA,
                                                   uint32_t countB, uint32_t queueCapacity, uint32_t *cursor,
                                                   const float4 *__restrict__ rayO, const float4 *__restrict__ rayD,
                                                   const float4 *__restrict__ rayC, bool cameraRays, uint2 *sharedStack,
                                                   uint32_t *warpPairs, PrefetchRing *ring, Finish finish) {
  constexpr uint32_t kRing = RT_PREFETCH, kHalf = kRing / 2, kChunk = kRing;
  static_assert((kRing & (kRing - 1)) == 0 && kRing >= 16 && kRing <= 64, "RT_PREFETCH: 16, 32 or 64");
  const unsigned full = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31u;
  LaneTraversal<kAny> t;
#if RT_SHARED_STACK > 0
  SplitStack<RT_SHARED_STACK, kTraceBlock> stack;
  stack.shared = sharedStack + threadIdx.x;
#else
  LocalStack stack;
  (void)sharedStack;
#endif
  bool active = false;
  const uint32_t count = countA + countB; // the cursor walks the class A entries (front) first, then class B (back)
  // The state of the prefetcher is warp-uniform and packed into two registers (as nine variables it cost the kernel 250
  // bytes of spills at 64 registers). rangeNext: the next queue entry to fetch; the cursor advances by kChunk from 0, so
  // the chunk it belongs to ends at the next multiple of kChunk. Lane 0's rangeNext also receives the cursor's old
  // value while an advance is in flight. pf: bits 0-5 head (first ready ring position), 6-12 ready rays, 13-18 rays of
  // the fetch in flight, 19-20 its stage (1: queue entries on their way into ring->slot, 2: rays on their way),
  // 21 the ring half the next fetch fills, 22 cursor advance in flight, 23 cursor past the end, 24 chunk used up.
  uint32_t rangeNext = 0;
  constexpr uint32_t kHeadMask = 63u, kAvailShift = 6, kAvailMask = 127u << 6, kCountShift = 13, kCountMask = 63u << 13,
                     kStageShift = 19, kStageMask = 3u << 19, kTailHalf = 1u << 21, kAtomicPending = 1u << 22, kNoMore = 1u << 23,
                     kRangeEmpty = 1u << 24;
  uint32_t pf = kRangeEmpty | (count == 0u ? kNoMore : 0u);
#ifdef RT_COUNT_WORK
  uint32_t tailIters = 0;
#endif
  while (true) {
    // the fetch pipeline advances one stage per iteration
    const uint32_t stage = (pf & kStageMask) >> kStageShift;
    if (stage != 0u) {
      const uint32_t fetchCount = (pf & kCountMask) >> kCountShift;
      const uint32_t e = ((pf & kTailHalf) ? 0u : kHalf) + lane; // the half filled now is the one before the tail half
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (stage == 1u) { // the queue entries have arrived in ring->slot: fetch the rays they name
        if (lane < fetchCount) {
          const uint32_t slot = ring->slot[e];
          if (kAny || !cameraRays) cpAsync16(&ring->o[e], rayO + slot);
          cpAsync16(&ring->d[e], rayD + slot);
          if (kAny) cpAsync16(&ring->c[e], rayC + slot);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        pf += 1u << kStageShift;
      } else { // the rays have arrived
        __syncwarp();
        if ((pf & kAvailMask) == 0u) pf = (pf & ~kHeadMask) | ((pf & kTailHalf) ? 0u : kHalf);
        pf = (pf + (fetchCount << kAvailShift)) & ~(kStageMask | kCountMask);
      }
    }
    if (pf & kAtomicPending) { // the cursor's value from the last iteration
      rangeNext = __shfl_sync(full, rangeNext, 0);
      pf &= ~(kAtomicPending | kRangeEmpty);
      if (rangeNext >= count) pf |= kNoMore | kRangeEmpty;
    }
    if ((pf & (kStageMask | kRangeEmpty)) == 0u &&
        ((pf & kAvailMask) == 0u || (((pf & kHeadMask) / kHalf) != ((pf & kTailHalf) ? 1u : 0u)))) {
      // start a fetch into the tail half: its queue entries -> ring->slot
      const uint32_t chunkEnd = min((rangeNext | (kChunk - 1u)) + 1u, count);
      const uint32_t fetchCount = min(kHalf, chunkEnd - rangeNext);
      if (lane < fetchCount) {
        const uint32_t j = rangeNext + lane;
        const uint32_t *src = j < countA ? queue + j : queue + (queueCapacity - 1u - (j - countA));
        const uint32_t dst = uint32_t(__cvta_generic_to_shared(&ring->slot[((pf & kTailHalf) ? kHalf : 0u) + lane]));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      rangeNext += fetchCount;
      if (rangeNext == chunkEnd) pf |= kRangeEmpty;
      pf = (pf ^ kTailHalf) | (fetchCount << kCountShift) | (1u << kStageShift);
    }
    if ((pf & (kRangeEmpty | kNoMore | kAtomicPending)) == kRangeEmpty) { // reserve the next chunk; picked up an iteration later
      if (lane == 0u) rangeNext = atomicAdd(cursor, kChunk);
      pf |= kAtomicPending;
    }
    const unsigned idle = __ballot_sync(full, !active);
    const bool drained = (pf & (kNoMore | kAtomicPending | kStageMask | kAvailMask)) == kNoMore;
    if (idle == full && drained) break;
    if ((pf & kAvailMask) != 0u && (idle == full || (kRefill > 0 && __popc(idle) >= kRefill))) {
      const uint32_t avail = (pf & kAvailMask) >> kAvailShift, headPos = pf & kHeadMask;
      const uint32_t take = min(uint32_t(__popc(idle)), avail);
      const uint32_t rank = uint32_t(__popc(idle & ((1u << lane) - 1u)));
      if (!active && rank < take) {
        const uint32_t e = (headPos + rank) & (kRing - 1u);
        const uint32_t newSlot = ring->slot[e];
        stack.set(kSlotPath, make_uint2(newSlot, 0u));
        const float4 d = ring->d[e];
        float4 o;
        if (!kAny && cameraRays)
          o = make_float4(P.uniforms.camera.position.x, P.uniforms.camera.position.y, P.uniforms.camera.position.z, 0.0f);
        else
          o = ring->o[e];
        if (kAny) {
          const float4 c = ring->c[e];
          stack.set(kSlotCarry0, make_uint2(__float_as_uint(c.x), __float_as_uint(c.y)));
          stack.set(kSlotCarry1, make_uint2(__float_as_uint(c.z), __float_as_uint(c.w)));
        }
        t.template begin<kFlat>(P.tlas, stack, o.x, o.y, o.z, d.x, d.y, d.z, 0.0f, kAny ? d.w : INFINITY);
        active = true;
      }
      pf = (pf & ~(kHeadMask | kAvailMask)) | ((headPos + take) & (kRing - 1u)) | ((avail - take) << kAvailShift);
      __syncwarp(); // the ring entries just read may be refilled from the next iteration on
    }
#ifdef RT_COUNT_WORK
    if (drained) tailIters += uint32_t(kStepsPerCheck);
#endif
#pragma unroll 1
    for (int k = 0; k < kStepsPerCheck; ++k) {
      if (t.template stepConverged<kFlat || (RT_CONVERGED & 1) != 0, !kFlat && (RT_CONVERGED & 4) != 0, (RT_CONVERGED & 2) != 0, (RT_CONVERGED & 8) != 0, 1 + ((RT_CONVERGED >> 4) & 3), (kAny ? RT_COOP_ANY : (kFlat ? RT_COOP_TRIS : RT_COOP_TRIS_REAL))>(P.tlas, stack, active, warpPairs)) {
        const uint32_t slot = stack.get(kSlotPath).x;
        if constexpr (kAny) {
          const uint2 c0 = stack.get(kSlotCarry0), c1 = stack.get(kSlotCarry1);
          const float4 carried = make_float4(__uint_as_float(c0.x), __uint_as_float(c0.y), __uint_as_float(c1.x), __uint_as_float(c1.y));
          finish(slot, t.found, RayHit{}, &carried);
        } else {
          finish(slot, t.found, t.result(stack), nullptr);
        }
#ifdef RT_COUNT_WORK
        countWork<kAny>(P, t);
#endif
        active = false;
      }
    }
  }
#ifdef RT_COUNT_WORK
  if (P.rayCounters != nullptr && lane == 0u) {
    atomicAdd(P.rayCounters + 19, (unsigned long long)tailIters);
    atomicMax(P.rayCounters + 20, (unsigned long long)tailIters);
    atomicAdd(P.rayCounters + 21, 1ull);
  }
#endif
}
#endif // RT_PREFETCH > 0

// One launch of the persistent traversal kernel does up to two jobs: the closest-hit rays of segment k (queue qin)
// and then the any-hit shadow rays the shade kernel of segment k - 1 emitted (parity shadowParity). Both only depend on
// that shade kernel, so putting them in one launch lets warps that run out of closest-hit rays go straight on to
// shadow rays instead of idling through the tail of a separate launch (a persistent launch has a ~50 us tail, which
// matters once a GPU holds only a slice of the frame). The closest-hit rays go first: they are the longer ones.
template <int kRefill, bool kFlat, int kMinBlocks, bool kPrefetch>
__global__ void __launch_bounds__(kTraceBlock, kMinBlocks) k_wf_traverse(const __grid_constant__ TraceParams P, const WfState W,
                                                                                 int qin, int firstSegment, int cameraRays,
                                                                                 int doClosest, int doShadow, int shadowParity) {
#if RT_SHARED_STACK > 0
  __shared__ uint2 s_stack[RT_SHARED_STACK * kTraceBlock]; // [entry][thread], used by both phases in turn
#else
  uint2 *s_stack = nullptr;
#endif
#if RT_COOP_TRIS || RT_COOP_ANY || RT_COOP_TRIS_REAL
  __shared__ uint32_t s_pairs[kTraceBlock]; // 32 words per warp: the pair list of the cooperative triangle stage
  uint32_t *warpPairs = s_pairs + (threadIdx.x & ~31u);
#else
  uint32_t *warpPairs = nullptr; // no static shared memory: the whole L1 stays a cache for the BVH
#endif
#if RT_PREFETCH > 0
  __shared__ PrefetchRing s_ring[kPrefetch ? kTraceBlock / 32 : 1];
  PrefetchRing *ring = s_ring + (kPrefetch ? (threadIdx.x >> 5) : 0);
#endif
  if (doClosest) {
    const int nextParity = doShadow ? (shadowParity ^ 1) : shadowParity; // parity of the segment traced here
    if (blockIdx.x == 0 && threadIdx.x == 0) { // the queues this segment's shade kernel appends to start empty
      W.counts[pathCount(qin ^ 1)] = 0u;
      W.counts[pathCountB(qin ^ 1)] = 0u;
      W.counts[plainCount(qin ^ 1)] = 0u;
      W.counts[shadowCount(nextParity)] = 0u;
      W.counts[shadowCountB(nextParity)] = 0u;
      W.counts[10 + nextParity] = 0u;
    }
    const uint32_t countA = W.counts[pathCount(qin)], countB = W.counts[pathCountB(qin)];
    auto finishClosest = [&](uint32_t slot, bool found, const RayHit &hit, const float4 *) {
                                 RT_STS(W.hitA + slot, make_float4(found ? hit.t : INFINITY, hit.u, hit.v,
                                                                   __uint_as_float(hit.primitive)));
                                 if (found)
                                   RT_STS(W.hitB + slot, make_uint2(hit.instance, hit.geometry));
                                 if (firstSegment && P.primaryIds != nullptr && slot < W.capacity) { // sample 0
                                   int px, py;
                                   bool valid;
                                   slotPixel(P, slot, px, py, valid);
                                   const uint4 id = found ? make_uint4(hit.instance, hit.geometry, hit.primitive, __float_as_uint(hit.t))
                                                            : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                                   reinterpret_cast<uint4 *>(P.primaryIds)[size_t(py) * size_t(P.uniforms.width) + size_t(px)] = id;
                                 }
                               };
#if RT_PREFETCH > 0
    if constexpr (kPrefetch)
      traceQueuePrefetch<false, kRefill, kFlat>(P, W.queue[qin], countA, countB, W.queueCapacity, W.counts + 3, W.rayO, W.rayD,
                                                nullptr, cameraRays != 0, s_stack, warpPairs, ring, finishClosest);
    else
#endif
      traceQueue<false, kRefill, kFlat>(P, W.queue[qin], countA, countB, W.queueCapacity, W.counts + 3, W.rayO, W.rayD,
                                        cameraRays != 0, s_stack, warpPairs, finishClosest);
    if (P.rayCounters != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
      atomicAdd(P.rayCounters + 0, (unsigned long long)countA + (unsigned long long)countB);
  }
  if (doShadow) {
    const uint32_t countA = W.counts[shadowCount(shadowParity)], countB = W.counts[shadowCountB(shadowParity)];
    auto finishShadow = [&](uint32_t slot, bool found, const RayHit &, const float4 *carried) {
                                if (!found) // unoccluded: the light sample contributes (shC = radiance + contribution)
                                  RT_STS(W.rad + slot, carried != nullptr ? *carried : RT_LDS(W.shC + slot));
                              };
#if RT_PREFETCH > 0
    if constexpr (kPrefetch)
      traceQueuePrefetch<true, kRefill, kFlat>(P, W.shadowQueue, countA, countB, W.queueCapacity, W.counts + 10 + shadowParity,
                                               W.rayO, W.shD, W.shC, false, s_stack, warpPairs, ring, finishShadow);
    else
#endif
      traceQueue<true, kRefill, kFlat>(P, W.shadowQueue, countA, countB, W.queueCapacity, W.counts + 10 + shadowParity, W.rayO,
                                       W.shD, false, s_stack, warpPairs, finishShadow);
    if (P.rayCounters != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
      atomicAdd(P.rayCounters + 1, (unsigned long long)countA + (unsigned long long)countB);
  }
}
#endif // RT_TU_TRAVERSE

#ifdef RT_TU_SHADE
// RT_SHADE_COMPACT: hits compacted per warp before shading (see k_wf_shade)
#ifndef RT_SHADE_COMPACT
#define RT_SHADE_COMPACT 1
#endif
// threads per CTA and resident CTAs per SM of the shade kernel: together they set its register budget
// (65536 / (block x minblocks)); 256 x 4 = 64 registers
#ifndef RT_SHADE_BLOCK
#define RT_SHADE_BLOCK 256
#endif
#ifndef RT_SHADE_MINBLOCKS
#define RT_SHADE_MINBLOCKS 4 // 64 registers: measured faster than 128 (the kernel is bound by gather latency; more warps hide it)
#endif
constexpr int kShadeBlock = RT_SHADE_BLOCK;
template <bool kTextures, bool kPlain, bool kClassify>
__global__ void __launch_bounds__(kShadeBlock, RT_SHADE_MINBLOCKS) k_wf_shade(const __grid_constant__ TraceParams P,
                                                                         const WfState W, int qin, int s0,
                                                                         int cameraRays, int shadowParity) {
  if (blockIdx.x == 0 && threadIdx.x == 0) W.counts[3] = 0u; // cursor of the next trace kernel
  // the paths of this segment: with classes the plain-order copy, otherwise the queue the traversal walked
  constexpr bool classify = kClassify;
  const uint32_t count = classify ? W.counts[plainCount(qin)] : W.counts[pathCount(qin)];
  const uint32_t *queue = classify ? W.shadeQueue[qin] : W.queue[qin];
  uint32_t *plainOut = classify ? W.shadeQueue[qin ^ 1] : nullptr;
  uint32_t hitCount = 0; // this thread's closest hits, added to the probe counter once per warp at the end
  // one hit: the path state in, shadeSegment, the state of the next segment and of the shadow ray out; pathIsB / shadowIsB:
  // the class of the ray it emits (WfState::classify)
  auto shadeHit = [&](uint32_t slot, const float4 &ha, bool &pushPath, bool &pushShadow, bool &pathIsB, bool &shadowIsB) {
    const uint32_t b = slot / W.capacity; // sample of the batch, pixel slot
    const uint32_t pixelSlot = slot - b * W.capacity;
    const int sampleIndex = s0 + int(b);
    const uint2 hb = RT_LDS(W.hitB + slot);
    RayHit hit;
    hit.t = ha.x, hit.u = ha.y, hit.v = ha.z;
    hit.instance = hb.x, hit.geometry = hb.y, hit.primitive = __float_as_uint(ha.w);
    const float4 d = RT_LDS(W.rayD + slot);
    float4 o, th, ra;
    int4 c;
    if (cameraRays) { // first segment: the state k_wf_generate did not write
      o = make_float4(P.uniforms.camera.position.x, P.uniforms.camera.position.y, P.uniforms.camera.position.z, 0.0f);
      th = make_float4(1.0f, 1.0f, 1.0f, d.w); // d.w = halton index bits
      ra = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      c = make_int4(0, 0, 0, 0);
    } else {
      o = RT_LDS(W.rayO + slot), th = RT_LDS(W.thr + slot), ra = RT_LDS(W.rad + slot);
      const uint32_t packed = __float_as_uint(d.w);
      c = make_int4(int(packed & 1023u), int((packed >> 10) & 1023u), int(packed >> 20), 0);
    }
    // per-pixel primary outputs are only touched by sample 0 (first segment, or until the G-buffer is written)
    const bool primarySegment = (c.x == 0 && sampleIndex == 0);
    const bool needPrimary = primarySegment || (sampleIndex == 0 && P.uniforms.enableDenoiseGBuffer != 0) ||
                             P.uniforms.debugTextureMode == RT_DEBUG_MOTION;
    float4 m4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), mi = m4;
    if (needPrimary) {
      m4 = RT_LDS(W.mot + pixelSlot);
      mi = RT_LDS(W.misc + pixelSlot);
    }
    PathState s;
    s.origin = mk3(o.x, o.y, o.z);
    s.dir = mk3(d.x, d.y, d.z);
    s.throughput = mk3(th.x, th.y, th.z);
    s.radiance = mk3(ra.x, ra.y, ra.z);
    s.bounce = c.x, s.step = c.y, s.transparencyPasses = c.z;
    s.bsdfPdf = o.w;
    const int hIndex = __float_as_int(th.w);
    PrimaryOutputs prim = emptyPrimaryOutputs();
    const uint32_t flags = __float_as_uint(mi.y);
    prim.depth = mi.x;
    prim.motion = mk2(m4.x, m4.y);
    prim.hadPrimaryHit = (flags & 1u) != 0u;
    prim.wroteGBuffer = (flags & 2u) != 0u;
    const bool hadGBuffer = prim.wroteGBuffer;
    ShadowRequest shadow;
    pushPath = shadeSegment<kTextures, kPlain>(P, s, hit, hIndex, sampleIndex, mk2(m4.z, m4.w), prim, shadow);
    // the radiance record changes only where the surface emits (or a debug view / the environment writes it): most hits
    // leave it as it is, and 16 of the ~180 bytes a hit moves need not be written back. A camera ray's record does not
    // exist yet (k_wf_generate wrote only the direction).
    if (cameraRays || s.radiance.x != ra.x || s.radiance.y != ra.y || s.radiance.z != ra.z)
      RT_STS(W.rad + slot, make_float4(s.radiance.x, s.radiance.y, s.radiance.z, 0.0f));
    if (pushPath) { // a path that ends here (every path of the last segment) leaves only its radiance behind
      const uint32_t packed = uint32_t(s.bounce) | (uint32_t(s.step) << 10) | (uint32_t(s.transparencyPasses) << 20);
      RT_STS(W.rayD + slot, make_float4(s.dir.x, s.dir.y, s.dir.z, __uint_as_float(packed)));
      RT_STS(W.thr + slot, make_float4(s.throughput.x, s.throughput.y, s.throughput.z, th.w));
      if (classify)
        pathIsB = !rayReachesNodes(P.nodeUnionBox, s.origin.x, s.origin.y, s.origin.z, s.dir.x, s.dir.y, s.dir.z, INFINITY);
    }
    // the shadow ray starts where the next segment starts (shadeSegment: both are hit + N * 1e-3), so one origin
    // record serves both; a path that ends but still has a shadow ray to trace stores it for that alone
    if (pushPath || shadow.valid) {
      const f3 org = pushPath ? s.origin : shadow.origin;
      RT_STS(W.rayO + slot, make_float4(org.x, org.y, org.z, pushPath ? s.bsdfPdf : 0.0f));
    }
    if (primarySegment || (prim.wroteGBuffer && !hadGBuffer)) {
      const uint32_t nf = (prim.hadPrimaryHit ? 1u : 0u) | (prim.wroteGBuffer ? 2u : 0u);
      RT_STS(W.mot + pixelSlot, make_float4(prim.motion.x, prim.motion.y, m4.z, m4.w));
      RT_STS(W.misc + pixelSlot, make_float4(prim.depth, __uint_as_float(nf), mi.z, 0.0f));
    }
    if (prim.wroteGBuffer && !hadGBuffer) {
      int px, py;
      bool valid;
      slotPixel(P, pixelSlot, px, py, valid);
      writeImage(P.images[RT_TEXTURE_DIFFUSE_ALBEDO], px, py, prim.gDiffuse);
      writeImage(P.images[RT_TEXTURE_SPECULAR_ALBEDO], px, py, prim.gSpecular);
      writeImage(P.images[RT_TEXTURE_NORMAL], px, py, prim.gNormal);
      writeImage(P.images[RT_TEXTURE_ROUGHNESS], px, py, prim.gRoughness);
    }
    if (shadow.valid) {
      pushShadow = true;
      RT_STS(W.shD + slot, make_float4(shadow.dir.x, shadow.dir.y, shadow.dir.z, shadow.tmax));
      // what the path's radiance becomes if the shadow ray gets through: the traversal kernel only has to copy it
      RT_STS(W.shC + slot, make_float4(s.radiance.x + shadow.contribution.x, s.radiance.y + shadow.contribution.y,
                                       s.radiance.z + shadow.contribution.z, 0.0f));
      // its origin is the record written above (the next segment's origin, or its own when the path ends here)
      const f3 so = pushPath ? s.origin : shadow.origin;
#ifndef RT_CLASSIFY_SHADOW
#define RT_CLASSIFY_SHADOW 1 // 0: only path rays are queued by class; shadow rays all go to the front of their queue
#endif
      if (classify && RT_CLASSIFY_SHADOW)
        shadowIsB = !rayReachesNodes(P.nodeUnionBox, so.x, so.y, so.z, shadow.dir.x, shadow.dir.y, shadow.dir.z, shadow.tmax);
    }
  };
  // a miss: a camera ray still has to leave radiance 0 behind for the fold (k_wf_generate did not write it); with the
  // environment extension bound the path picks it up first
  auto shadeMissed = [&](uint32_t slot) {
    PathState s;
    s.radiance = mk3(0.0f);
    s.throughput = mk3(1.0f);
    const float4 d = RT_LDS(W.rayD + slot);
    s.dir = mk3(d.x, d.y, d.z);
    s.bsdfPdf = 0.0f;
    if (!cameraRays) {
      const float4 th = RT_LDS(W.thr + slot), ra = RT_LDS(W.rad + slot);
      s.throughput = mk3(th.x, th.y, th.z);
      s.radiance = mk3(ra.x, ra.y, ra.z);
      if (environmentIsLight(P)) s.bsdfPdf = RT_LDS(W.rayO + slot).w;
    }
    shadeMiss(P, s);
    RT_STS(W.rad + slot, make_float4(s.radiance.x, s.radiance.y, s.radiance.z, 0.0f));
  };
  const bool missesMatter = cameraRays || P.env.texelsDev != nullptr;
#if RT_SHADE_COMPACT
  // Hits are compacted per warp before shading: a warp takes 64 queue entries per round (two per lane), appends the
  // slots of the hits to a list in shared memory and shades 32 of them whenever the list holds that many; what is
  // left after the last round is the only partial pass. On bounce segments 40 % of the entries are hits: shading
  // them where they sit keeps ~13 lanes of a warp busy, this keeps 32 — and the number of resident warps, which is
  // what hides the gather latency, stays what it was (compacting per CTA took warps away and was slower).
  __shared__ uint32_t s_slots[kShadeBlock / 32][96];
  uint32_t *mySlots = s_slots[threadIdx.x >> 5];
  const uint32_t lane = threadIdx.x & 31u;
  const unsigned below = (1u << lane) - 1u;
  const uint32_t span = gridDim.x * blockDim.x * 2u;
  const uint32_t rounds = (count + span - 1u) / span;
  uint32_t jw = (blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) * 2u;
  uint32_t listed = 0; // warp-uniform: hits waiting in mySlots (< 32 between rounds)
  for (uint32_t round = 0; round <= rounds; ++round, jw += span) { // whole warps stay in the loop; the extra trip drains
    if (round < rounds) {
      bool hit0 = false, hit1 = false;
      uint32_t slot0 = 0, slot1 = 0;
      if (jw + lane < count) {
        slot0 = queue[jw + lane];
        hit0 = RT_LDS(reinterpret_cast<const float *>(W.hitA + slot0)) < INFINITY;
        if (!hit0 && missesMatter) shadeMissed(slot0);
      }
      if (jw + 32u + lane < count) {
        slot1 = queue[jw + 32u + lane];
        hit1 = RT_LDS(reinterpret_cast<const float *>(W.hitA + slot1)) < INFINITY;
        if (!hit1 && missesMatter) shadeMissed(slot1);
      }
      const unsigned votes0 = __ballot_sync(0xFFFFFFFFu, hit0), votes1 = __ballot_sync(0xFFFFFFFFu, hit1);
      const uint32_t n0 = uint32_t(__popc(votes0));
      if (hit0) mySlots[listed + uint32_t(__popc(votes0 & below))] = slot0;
      if (hit1) mySlots[listed + n0 + uint32_t(__popc(votes1 & below))] = slot1;
      listed += n0 + uint32_t(__popc(votes1));
      __syncwarp();
    }
    while (listed >= 32u || (round == rounds && listed != 0u)) {
      const uint32_t take = min(listed, 32u);
      listed -= take;
      bool pushPath = false, pushShadow = false, pathIsB = false, shadowIsB = false;
      uint32_t slot = 0;
      if (lane < take) {
        slot = mySlots[listed + lane];
        shadeHit(slot, RT_LDS(W.hitA + slot), pushPath, pushShadow, pathIsB, shadowIsB);
        ++hitCount;
      }
      // shadowParity == qin (both are the segment's parity), so the two lengths of a class are the halves of one 64-bit word
      if (classify)
        queuePushClassified(W.queue[qin ^ 1], W.shadowQueue, W.counts + pathCount(qin ^ 1), W.counts + pathCountB(qin ^ 1),
                            W.queueCapacity, pushPath, pathIsB, pushShadow, shadowIsB, slot, plainOut,
                            W.counts + plainCount(qin ^ 1));
      else
        queuePushBoth(W.queue[qin ^ 1], W.shadowQueue, W.counts + pathCount(qin ^ 1), pushPath, pushShadow, slot);
    }
    __syncwarp();
  }
#else
  const uint32_t rounds = (count + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint32_t round = 0; round < rounds; ++round, j += gridDim.x * blockDim.x) { // whole warps stay in the loop
    bool pushPath = false, pushShadow = false, pathIsB = false, shadowIsB = false;
    uint32_t slot = 0;
    float4 ha = make_float4(INFINITY, 0.0f, 0.0f, 0.0f);
    if (j < count) {
      slot = queue[j];
      ha = RT_LDS(W.hitA + slot);
    }
    if (ha.x < INFINITY) {
      shadeHit(slot, ha, pushPath, pushShadow, pathIsB, shadowIsB);
      ++hitCount;
    } else if (j < count && missesMatter) {
      shadeMissed(slot);
    }
    if (classify)
      queuePushClassified(W.queue[qin ^ 1], W.shadowQueue, W.counts + pathCount(qin ^ 1), W.counts + pathCountB(qin ^ 1),
                          W.queueCapacity, pushPath, pathIsB, pushShadow, shadowIsB, slot, plainOut,
                          W.counts + plainCount(qin ^ 1));
    else
      queuePushBoth(W.queue[qin ^ 1], W.shadowQueue, W.counts + pathCount(qin ^ 1), pushPath, pushShadow, slot);
  }
#endif
  if (P.rayCounters != nullptr) {
    const uint32_t warpHits = __reduce_add_sync(0xFFFFFFFFu, hitCount);
    if ((threadIdx.x & 31) == 0 && warpHits) atomicAdd(P.rayCounters + 2, (unsigned long long)warpHits);
  }
}

__global__ void __launch_bounds__(kBlock) k_wf_resolve(const __grid_constant__ TraceParams P, const WfState W,
                                                       int lastS0, int lastN) {
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < W.capacity; slot += gridDim.x * blockDim.x) {
    int px, py;
    bool valid;
    slotPixel(P, slot, px, py, valid);
    if (!valid) continue;
    const float4 t4 = RT_LDS(W.tot + slot);
    f3 total = mk3(t4.x, t4.y, t4.z);
    const int totalSamples = int(__float_as_uint(t4.w));
    foldBatch(W, slot, lastS0, lastN, totalSamples, total);
    const float4 m4 = RT_LDS(W.mot + slot);
    const float4 mi = RT_LDS(W.misc + slot);
    PrimaryOutputs prim = emptyPrimaryOutputs();
    prim.depth = mi.x;
    prim.motion = mk2(m4.x, m4.y);
    // G-buffer images were written by the shade phase; resolvePixel must not overwrite them
    resolvePixel(P, px, py, total, totalSamples, mk2(m4.z, m4.w), prim, false);
  }
}

#endif // RT_TU_SHADE

#ifdef RT_TU_TRAVERSE
int ensureState(rt_context *ctx, int lane, uint32_t capacity, uint32_t batch, WfState &out) {
  const size_t paths = size_t(capacity) * batch;
  size_t sortTempBytes = 0;
  if (ctx->sortRays > 0)
    cub::DeviceRadixSort::SortPairs(nullptr, sortTempBytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (uint32_t *)nullptr, int(paths), 0, 24, ctx->stream);
  const size_t need = 256 + 8 * (paths * 16 + 256) + 3 * (size_t(capacity) * 16 + 256) + 5 * (paths * 4 + 256) +
                      (ctx->sortRays > 0 ? 3 * (paths * 4 + 256) + sortTempBytes + 256 : 0);
  if (ctx->wfState[lane] == nullptr || ctx->wfBytes[lane] < need) {
    RT_CUDA(RT_SYNC_STREAM(ctx, ctx->stream)); // every lane joined the context's stream at the end of its last dispatch
    if (ctx->wfState[lane]) cudaFree(ctx->wfState[lane]);
    ctx->wfState[lane] = nullptr;
    ctx->wfBytes[lane] = 0;
    RT_CUDA(cudaMalloc(&ctx->wfState[lane], need));
    ctx->wfBytes[lane] = need;
  }
  uint8_t *p = static_cast<uint8_t *>(ctx->wfState[lane]);
  auto take = [&](size_t bytes) {
    void *r = p;
    p += (bytes + 255) & ~size_t(255);
    return r;
  };
  WfState s{};
  s.capacity = capacity;
  s.batch = batch;
  s.queueCapacity = uint32_t(paths);
  s.classify = 0u;
  const size_t v = paths * 16, pv = size_t(capacity) * 16;
  s.counts = static_cast<uint32_t *>(take(256));
  s.rayO = static_cast<float4 *>(take(v));
  s.rayD = static_cast<float4 *>(take(v));
  s.thr = static_cast<float4 *>(take(v));
  s.rad = static_cast<float4 *>(take(v));
  s.hitA = static_cast<float4 *>(take(v));
  s.hitB = static_cast<uint2 *>(take(v / 2));
  s.shD = static_cast<float4 *>(take(v));
  s.shC = static_cast<float4 *>(take(v));
  s.tot = static_cast<float4 *>(take(pv));
  s.mot = static_cast<float4 *>(take(pv));
  s.misc = static_cast<float4 *>(take(pv));
  s.queue[0] = static_cast<uint32_t *>(take(paths * 4));
  s.queue[1] = static_cast<uint32_t *>(take(paths * 4));
  s.shadowQueue = static_cast<uint32_t *>(take(paths * 4));
  s.shadeQueue[0] = static_cast<uint32_t *>(take(paths * 4));
  s.shadeQueue[1] = static_cast<uint32_t *>(take(paths * 4));
  if (ctx->sortRays > 0) {
    s.sortKeys[0] = static_cast<uint32_t *>(take(paths * 4));
    s.sortKeys[1] = static_cast<uint32_t *>(take(paths * 4));
    s.sortedQueue = static_cast<uint32_t *>(take(paths * 4));
    s.sortTemp = take(sortTempBytes);
    s.sortTempBytes = sortTempBytes;
  }
  RT_CHECK(size_t(p - static_cast<uint8_t *>(ctx->wfState[lane])) <= ctx->wfBytes[lane], "internal: wavefront state overflow");
  out = s;
  return 0;
}

// Sorts `*queue` (count read back from the device) by ray key; the sorted copy takes the queue's place.
int sortQueue(rt_context *ctx, const TraceParams &P, WfState &W, uint32_t **queue, const uint32_t *countDev,
              const float4 *rayO, const float4 *rayD) {
  cudaStream_t st = ctx->stream;
  uint32_t count = 0;
  RT_CUDA(cudaMemcpyAsync(&count, countDev, 4, cudaMemcpyDeviceToHost, st));
  RT_CUDA(RT_SYNC_STREAM(ctx, st));
  if (count < 65536u) return 0; // not worth two more launches
  ctx->mark(-1);
  const int grid = int(std::min<uint32_t>((count + kBlock - 1) / kBlock, uint32_t(ctx->smCount) * 8u));
  k_ray_keys<<<grid, kBlock, 0, st>>>(*queue, count, rayO, rayD, P.tlasRootBox, W.sortKeys[0]);
  size_t bytes = W.sortTempBytes;
  RT_CUDA(cub::DeviceRadixSort::SortPairs(W.sortTemp, bytes, W.sortKeys[0], W.sortKeys[1], *queue, W.sortedQueue,
                                          int(count), 0, 24, st));
  ctx->launches += 4;
  ctx->mark(RT_KERNEL_OTHER);
  std::swap(*queue, W.sortedQueue);
  return 0;
}

#endif // RT_TU_TRAVERSE

} // namespace

#ifdef RT_TU_SHADE
void wfLaunchGenerate(int grid, cudaStream_t st, const TraceParams &P, const WfState &W, int s0, int n, int prevS0, int prevN,
                      int baseSamples, int maxExtraSamples) {
  if (W.classify != 0u) k_wf_generate<true><<<grid, kBlock, 0, st>>>(P, W, s0, n, prevS0, prevN, baseSamples, maxExtraSamples);
  else k_wf_generate<false><<<grid, kBlock, 0, st>>>(P, W, s0, n, prevS0, prevN, baseSamples, maxExtraSamples);
}
// the shade kernel is specialised on what the host knows: texture hint, plain PBR without debug views
void wfLaunchShade(bool textures, bool plain, int grid, cudaStream_t st, const TraceParams &P, const WfState &W, int qin,
                   int s0, int cameraRays, int parity) {
#define RT_SHADE(T, PL)                                                                                       \
  do {                                                                                                         \
    if (W.classify != 0u) k_wf_shade<T, PL, true><<<grid * (kBlock / kShadeBlock), kShadeBlock, 0, st>>>(P, W, qin, s0, cameraRays, parity); \
    else k_wf_shade<T, PL, false><<<grid * (kBlock / kShadeBlock), kShadeBlock, 0, st>>>(P, W, qin, s0, cameraRays, parity);                 \
  } while (0)
  if (textures && plain) RT_SHADE(true, true);
  else if (textures) RT_SHADE(true, false);
  else if (plain) RT_SHADE(false, true);
  else RT_SHADE(false, false);
#undef RT_SHADE
}
void wfLaunchResolve(int grid, cudaStream_t st, const TraceParams &P, const WfState &W, int lastS0, int lastN) {
  k_wf_resolve<<<grid, kBlock, 0, st>>>(P, W, lastS0, lastN);
}
int wfShadeIsFast() {
#ifdef RT_FAST_SHADE
  return 1;
#else
  return 0;
#endif
}
#endif // RT_TU_SHADE

#ifdef RT_TU_TRAVERSE

// One dispatch = `lanes` independent pipelines over interleaved tile subsets (lane l of L on rank r of N renders the
// tiles with tile % (N L) == r + l N — the multi-GPU tile partition applied once more), each with its own path state,
// queues and stream. A persistent launch ends with a tail in which ever fewer warps finish the longest rays (50-100 us
// however small the launch); with two lanes the other lane's next launch moves onto the SMs the tail frees, so the
// frame pays the tails of its last launches only instead of one per launch. Results do not depend on the partition
// (tests: tile partition == full frame). Zero device->host read-backs: glass paths, which can take up to
// maxBounces (maxBounces + 1) segments, are handled by launching that many segments — a launch whose queue turns out
// to be empty ends after one load.
int launchTraceWavefront(rt_context *ctx, const TraceParams &P0) {
  const rt_uniforms &U = P0.uniforms;
  const int baseSamples = std::max(U.samplesPerPixel, 1);
  const int maxExtraSamples = (U.enableMotionAdaptiveSampling != 0) ? std::max(U.motionSamplingMaxExtraSamples, 0) : 0;
  const int sampleLoopBound = baseSamples + maxExtraSamples;
  const int maxBounces = std::max(U.maxBounces, 0);
  // a refraction does not consume a bounce until transparencyPasses > maxBounces (Raytracing.metal:563-575)
  // without glass (RT_TRACE_HINT_NO_GLASS) every segment consumes a bounce: exactly maxBounces segments
  const int maxSegments = (P0.hints & RT_TRACE_HINT_NO_GLASS) ? maxBounces : maxBounces * (maxBounces + 1);
  const int tileCount = P0.tilesX * P0.tilesY;
  const int ownedAll = (tileCount - P0.tileRemainder + P0.tileModulo - 1) / P0.tileModulo;
  // ray sorting reads queue lengths back (an experiment, off by default): one lane, on the context's stream
  // pipeline_lanes 0 = auto: two lanes once the dispatch is large enough for the overlap to outweigh the doubled launch
  // count (measured, profiles/r2_lanes.md: +2.4 % on the 33 M-path benchmark frame, -1 ... -3 % on frames of 2 - 4 M paths)
  int wanted = ctx->pipelineLanes;
  if (wanted <= 0) {
    const size_t paths = size_t(ownedAll) * 256u * size_t(std::max(1, std::min(ctx->sampleBatch, sampleLoopBound)));
    wanted = paths >= (size_t(16) << 20) ? 2 : 1;
  }
  int lanes = ctx->sortRays > 0 ? 1 : std::max(1, std::min(std::min(wanted, kMaxLanes), ownedAll));
  struct Lane {
    TraceParams P;
    WfState W;
    cudaStream_t st;
    int id; // -1: the context's own stream (single lane)
    int persistent;
    int qin = 0;
    bool shadowPending = false;
    int pendingParity = 0;
    bool first = true; // no launch of this lane has been timed yet
  };
  std::vector<Lane> L;
  L.resize(size_t(lanes));
  for (int l = 0; l < lanes; ++l) {
    Lane &ln = L[size_t(l)];
    ln.P = P0;
    ln.P.tileModulo = P0.tileModulo * lanes;
    ln.P.tileRemainder = P0.tileRemainder + l * P0.tileModulo;
    const int owned = (tileCount - ln.P.tileRemainder + ln.P.tileModulo - 1) / ln.P.tileModulo;
    const uint32_t capacity = uint32_t(std::max(owned, 0)) * 256u;
    // samples in flight per pixel: as many as the option allows while the path state stays under ~48 M paths
    // (10.6 GB) over all lanes; the motion debug view reads what sample 0 wrote for its pixel, so it keeps one sample
    int batch = std::max(1, std::min(ctx->sampleBatch, sampleLoopBound));
    batch = std::min<int>(batch, std::max<uint32_t>(1u, ((48u << 20) / uint32_t(lanes)) / std::max(capacity, 1u)));
    if (U.debugTextureMode == RT_DEBUG_MOTION) batch = 1;
    RT_TRY(ensureState(ctx, l, std::max(capacity, 256u), uint32_t(batch), ln.W));
    ln.W.capacity = capacity;
    // rays queued by class (A: reaches a BLAS with nodes, B: cheap) when the TLAS is flat; the ray-sorting experiment
    // works on plain queues
    ln.W.classify = (ctx->classifyRays != 0 && ctx->sortRays == 0 && P0.tlas.instanceCount <= kFlatTlasMax &&
                     P0.tlas.instanceBox != nullptr) ? 1u : 0u;
    ln.id = lanes > 1 ? l : -1;
    ln.st = lanes > 1 ? ctx->laneStream[l] : ctx->stream;
    const int slotBlocks = (int(capacity) + kBlock - 1) / kBlock;
    ln.persistent = std::max(1, std::min(slotBlocks, ctx->smCount * 8));
  }
  const int batch = int(L[0].W.batch); // lane 0 owns the most tiles, so its batch is the smallest; all lanes use it
  for (Lane &ln : L) ln.W.batch = uint32_t(batch);
  // traversal kernels: exactly the resident CTA count (they pull work from a cursor): 8 per SM for large dispatches, 7 for
  // small ones (see kMinBlocksSmall); blocks_per_sm > 0 overrides the grid size
  const bool largeDispatch = size_t(L[0].W.capacity) * size_t(batch) >= kLargeDispatchPaths;
  const int traceGrid = ctx->smCount * (ctx->blocksPerSm > 0 ? ctx->blocksPerSm : (largeDispatch ? kMinBlocksLarge : kMinBlocksSmall));
  if (lanes > 1) { // fork: the lanes start after everything enqueued on the context's stream so far
    ctx->mark(-1);
    RT_CUDA(cudaEventRecord(ctx->evFork, ctx->stream));
    for (Lane &ln : L) RT_CUDA(cudaStreamWaitEvent(ln.st, ctx->evFork, 0));
  }
  auto timed = [&](Lane &ln, int klass) { // the launch just enqueued on this lane belongs to `klass`
    ctx->mark(klass, ln.id, ln.id >= 0 && ln.first);
    ln.first = false;
    ++ctx->launches;
  };
  auto traverse = [&](Lane &ln, int first, int cameraRays, int doClosest, int doShadow, int shadowParity) {
    const TraceParams &P = ln.P;
    const WfState &W = ln.W;
    cudaStream_t st = ln.st;
    const int qin = ln.qin;
    // instantiations: lane refill threshold (traversal_variant 0 / 1 / 2 = never / 8 / 16 idle lanes) x TLAS kind
    const bool flat = P.tlas.instanceCount <= kFlatTlasMax && P.tlas.instanceBox != nullptr;
#define RT_LAUNCH_TRAVERSE(R, F)                                                                                              \
  do {                                                                                                                       \
    if (largeDispatch)                                                                                                       \
      k_wf_traverse<R, F, kMinBlocksLarge, kPrefetchLarge><<<traceGrid, kTraceBlock, 0, st>>>(P, W, qin, first, cameraRays, doClosest, doShadow, shadowParity); \
    else                                                                                                                     \
      k_wf_traverse<R, F, kMinBlocksSmall, kPrefetchSmall><<<traceGrid, kTraceBlock, 0, st>>>(P, W, qin, first, cameraRays, doClosest, doShadow, shadowParity); \
  } while (0)
    switch (ctx->traversalVariant * 2 + (flat ? 1 : 0)) {
      case 0: RT_LAUNCH_TRAVERSE(0, false); break;
      case 1: RT_LAUNCH_TRAVERSE(0, true); break;
      case 4: RT_LAUNCH_TRAVERSE(16, false); break;
      case 5: RT_LAUNCH_TRAVERSE(16, true); break;
      case 3: RT_LAUNCH_TRAVERSE(8, true); break;
      default: RT_LAUNCH_TRAVERSE(8, false); break;
    }
#undef RT_LAUNCH_TRAVERSE
  };
  int prevS0 = 0, prevN = 0;
  for (int s0 = 0; s0 < sampleLoopBound;) {
    // the first batch stays within the base samples (see k_wf_generate)
    const int n = std::min(batch, (s0 < baseSamples ? baseSamples : sampleLoopBound) - s0);
    for (Lane &ln : L) {
      if (ln.W.capacity == 0) continue;
      RT_CUDA(cudaMemsetAsync(ln.W.counts, 0, 128, ln.st));
      if (ln.id < 0) ctx->mark(-1);
      wfLaunchGenerate(ln.persistent, ln.st, ln.P, ln.W, s0, n, prevS0, prevN, baseSamples, maxExtraSamples);
      timed(ln, RT_KERNEL_GENERATE);
      ln.qin = 0;
      ln.shadowPending = false;
      ln.pendingParity = 0;
    }
    for (int segment = 0; segment < maxSegments; ++segment) {
      const int parity = segment & 1;
      const int first = (s0 == 0 && segment == 0) ? 1 : 0;
      for (Lane &ln : L) {
        if (ln.W.capacity == 0) continue;
        const TraceParams &P = ln.P;
        WfState &W = ln.W;
        if (ctx->sortRays > 0 && segment > 0) RT_TRY(sortQueue(ctx, P, W, &W.queue[ln.qin], W.counts + pathCount(ln.qin), W.rayO, W.rayD));
        if (ctx->fuseTraversal && ln.shadowPending && ctx->sortRays < 2) {
          traverse(ln, first, segment == 0, 1, 1, ln.pendingParity); // closest hits of this segment + shadow rays of the last
          ln.shadowPending = false;
          timed(ln, RT_KERNEL_TRACE);
        } else {
          if (ln.shadowPending) {
            if (ctx->sortRays > 1) RT_TRY(sortQueue(ctx, P, W, &W.shadowQueue, W.counts + shadowCount(ln.pendingParity), W.rayO, W.shD));
            traverse(ln, 0, 0, 0, 1, ln.pendingParity);
            ln.shadowPending = false;
            timed(ln, RT_KERNEL_SHADOW);
          }
          traverse(ln, first, segment == 0, 1, 0, parity);
          timed(ln, RT_KERNEL_TRACE);
        }
        {
          const bool textures = (P.hints & RT_TRACE_HINT_UNTEXTURED) == 0u;
          const bool plain = U.debugTextureMode == RT_DEBUG_NONE && U.shadingMode != RT_SHADING_LEGACY && !environmentIsLight(P);
          wfLaunchShade(textures, plain, ln.persistent, ln.st, P, W, ln.qin, s0, segment == 0, parity);
        }
        timed(ln, RT_KERNEL_SHADE);
        ln.shadowPending = true;
        ln.pendingParity = parity;
        ln.qin ^= 1;
      }
    }
    for (Lane &ln : L) {
      if (ln.W.capacity == 0 || !ln.shadowPending) continue;
      if (ctx->sortRays > 1) RT_TRY(sortQueue(ctx, ln.P, ln.W, &ln.W.shadowQueue, ln.W.counts + shadowCount(ln.pendingParity), ln.W.rayO, ln.W.shD));
      traverse(ln, 0, 0, 0, 1, ln.pendingParity);
      timed(ln, RT_KERNEL_SHADOW);
      ln.shadowPending = false;
    }
    prevS0 = s0;
    prevN = n;
    s0 += n;
  }
  for (Lane &ln : L) {
    if (ln.W.capacity == 0) continue;
    wfLaunchResolve(ln.persistent, ln.st, ln.P, ln.W, prevS0, prevN);
    timed(ln, RT_KERNEL_RESOLVE);
  }
  if (lanes > 1) { // join: the context's stream continues after every lane
    for (int l = 0; l < lanes; ++l) {
      RT_CUDA(cudaEventRecord(ctx->evLaneDone[l], L[size_t(l)].st));
      RT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->evLaneDone[l], 0));
    }
    ctx->mark(-1);
  }
  RT_CUDA(cudaGetLastError());
  return 0;
}

#endif // RT_TU_TRAVERSE

} // namespace rtb
