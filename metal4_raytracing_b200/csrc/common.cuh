// common.cuh — shared declarations of the sm_100a library: context, error plumbing, acceleration-structure
// records. Device layouts are sized for 128-bit loads (LDG.128): an 8-wide node is 5 x 16 B, a triangle 3 x 16 B,
// an instance record 4 x 16 B.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rt_b200.h"

namespace rtb {

void setError(const std::string &msg);
#define RT_CUDA(expr)                                                                                         \
  do {                                                                                                         \
    cudaError_t _e = (expr);                                                                                   \
    if (_e != cudaSuccess) {                                                                                   \
      rtb::setError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" +        \
                    std::to_string(__LINE__) + ")");                                                           \
      return 1;                                                                                                \
    }                                                                                                          \
  } while (0)
// every blocking wait of the library goes through these, so tests can assert that a per-frame path has none
#define RT_SYNC_STREAM(ctx, s) (++(ctx)->hostSyncs, cudaStreamSynchronize(s))
#define RT_SYNC_EVENT(ctx, e) (++(ctx)->hostSyncs, cudaEventSynchronize(e))
#define RT_CHECK(cond, msg)          \
  do {                               \
    if (!(cond)) {                   \
      rtb::setError(msg);            \
      return 2;                      \
    }                                \
  } while (0)
#define RT_TRY(expr)       \
  do {                     \
    int _r = (expr);       \
    if (_r != 0) return _r; \
  } while (0)

// ---- 8-wide quantised node, 80 bytes (after Ylitie, Karras, Laine 2017 "compressed wide BVH") ---------------
//  w0: origin.x, origin.y, origin.z (float bits), {ex, ey, ez, imask} bytes
//  w1: childBase, primBase, meta[0..3], meta[4..7]
//  w2: qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7]
//  w3: qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7]
//  w4: qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7]
// child box = origin + q * 2^(e-127) per axis. meta[s]: 0 = empty slot; internal child = 0b001'11sss
// (low five bits 24 + slot); leaf child = unary primitive count in bits 5..7 (1 -> 001, 2 -> 011, 3 -> 111) and the
// primitive's offset from primBase (0..23) in the low five bits.
struct alignas(16) WideNode {
  uint4 w[5];
};
static_assert(sizeof(WideNode) == 80, "wide node is 80 bytes");

// Triangle record, 48 bytes: raw vertices (the watertight test needs them untransformed) + ids in the w lanes.
struct alignas(16) TriRecord {
  float4 v0; // w = primitive index within its geometry (bits)
  float4 v1; // w = geometry index (bits)
  float4 v2; // w = unused
};
static_assert(sizeof(TriRecord) == 48, "triangle record is 48 bytes");

// Device-resident header of a BLAS; its address is the accelerationStructureID the host writes into descriptors.
struct BlasHeader {
  const WideNode *nodes;
  const TriRecord *tris;
  float boundsLo[3];
  uint32_t triCount;
  float boundsHi[3];
  uint32_t nodeCount;
};

// Per-instance traversal record, 64 bytes: world->object 3x4 (rows), then the BLAS arrays.
struct alignas(16) InstanceRecord {
  float4 row0, row1, row2; // object = row_r . (world, 1)
  const WideNode *nodes;
  const TriRecord *tris; // low 5 bits: triangle count of a single-leaf-node BLAS (tested directly), else 0
};
static_assert(sizeof(InstanceRecord) == 64, "instance record is 64 bytes");

struct TlasHeader {
  const WideNode *nodes;
  const InstanceRecord *instances; // indexed by instance id (descriptor order)
  const uint32_t *leafInstance;    // leaf slot (primBase + offset) -> instance id
  const float4 *instanceBox;       // world box of instance i: [2 i] = lo, [2 i + 1] = hi (exact floats, k_instance_bounds)
  uint32_t instanceCount;
  uint32_t nodeCount;
};
// A TLAS of at most this many instances (the reference's scenes have 2 - 8, AppScene.swift:14-28) is one wide node whose
// leaf slot k holds instance k; the traversal then tests a new ray against the instances' world boxes directly instead of
// stepping through that node (traverse.cuh LaneTraversal::begin).
constexpr uint32_t kFlatTlasMax = 8;

// Result of a device-side TLAS build (bvh_build.cu k_tlas_build_cta): lives in device memory, is copied to pinned host
// memory after every build without waiting, and is looked at one library call later (status != 0 = tree too deep for
// the traversal stack: reported as an error then, never silently).
struct TlasBuildInfo {
  uint32_t nodeCount, levelCount, status, builds;
};

struct Aabb {
  float lo[3], hi[3];
};

// Host-side bookkeeping of one acceleration structure (BLAS or TLAS).
struct AccelObject {
  bool isTlas = false;
  uint32_t flags = 0;
  uint32_t primCount = 0;      // triangles or instances
  uint32_t primCapacity = 0;
  uint32_t nodeCount = 0;
  uint32_t nodeCapacity = 0;
  std::vector<uint32_t> levelStart; // wide-node index where each BFS level starts (+ end sentinel)
  void *headerDev = nullptr;   // BlasHeader / TlasHeader
  WideNode *nodes = nullptr;
  float4 *nodeBox = nullptr;   // exact float box of each wide node: 2 x float4 per node
  TriRecord *tris = nullptr;   // BLAS
  InstanceRecord *instances = nullptr; // TLAS (in descriptor order)
  float4 *instanceBox = nullptr;       // TLAS: world boxes in descriptor order (lo, hi)
  uint32_t *leafPrim = nullptr; // TLAS: leaf slot -> instance index (the node's primBase + offset indexes this)
  uint2 *triSource = nullptr;   // BLAS: per triangle slot (geometry, primitive) — refit source mapping
  uint32_t *nodeParent = nullptr, *nodePending = nullptr; // refittable BLAS / TLAS: parent links, per-refit child counters
  // TLAS built on the device (no host round trip): node / level counts live there
  TlasBuildInfo *infoDev = nullptr, *infoHost = nullptr; // infoHost: pinned
  cudaEvent_t infoEvent = nullptr;                        // the copy of infoDev into infoHost has landed
  bool infoPending = false;                               // a build's info has not been checked yet
  bool deviceBuilt = false;                               // the current tree came from k_tlas_build_cta
  bool treeValid = false;                                 // a tree over primCount instances exists (refit possible)
  // geometry table for refit: device copy of per-geometry (vertex ptr, stride, index ptr, index stride)
  void *geomTableDev = nullptr;
  std::vector<uint8_t> geomTableHost; // what geomTableDev holds: a refit with the same buffers uploads nothing
  uint32_t geomCount = 0;
  uint64_t bytes = 0;
  float sahCost = 0.0f;
  Aabb bounds{};
};

} // namespace rtb

namespace rtb {
// Per-kernel-class device timing (rt_kernel_timing_*): when enabled, an event is recorded after every launch and
// the interval since the previous event is attributed to that launch's class. Off by default (no events at all).
// With pipeline lanes (trace_wavefront.cu) launches of different streams overlap at their edges, so an interval is not
// simply "since the previous event": record i ends a launch of class klass[i] whose stream predecessor is record
// prev[i]; rt_kernel_timing_read charges it the time from max(predecessor's end, latest end of any other launch before
// it) to its own end — the machine's busy time is attributed once, to the launch that ended each stretch.
struct KernelTimer {
  bool enabled = false;
  std::vector<cudaEvent_t> pool;
  std::vector<int> klass; // class of the launch that ends at event i; -1 = not attributed (sequence start)
  std::vector<int> prev;  // record of the previous event on the same stream, -1 = none
  size_t used = 0;
  int last[1 + 4] = {-1, -1, -1, -1, -1}; // latest record per stream: [0] the context stream, [1 + lane] the lanes
};
constexpr int kMaxLanes = 4;
} // namespace rtb

struct rt_context {
  int device = 0;
  cudaStream_t ownStream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t evBegin = nullptr, evEnd = nullptr;
  // asynchronous read-back (rt_download_async): copy stream, "work so far is done" event, ring of completion events
  cudaStream_t copyStream = nullptr;
  cudaEvent_t evReady = nullptr;
  cudaEvent_t evCopied[8] = {};
  uint64_t copiesIssued = 0;
  cudaEvent_t evFence[16] = {}; // rt_fence ring
  uint64_t fencesIssued = 0;
  uint64_t launches = 0;
  uint64_t hostSyncs = 0; // times the library blocked the host on the device (rt_host_sync_count): per-frame paths keep it 0
  int traceMode = 1;        // 0 megakernel, 1 wavefront
  int traversalVariant = 1; // lane refill threshold of the traversal kernels: 0 none, 1 = 8, 2 = 16, 3 = 24 idle lanes
  int fuseTraversal = 1;    // wavefront: shadow rays of segment k traced in the launch of segment k + 1's closest hits
  int sortRays = 0;         // wavefront: 0 = queues in arrival order, 1 = bounce rays sorted by octant + origin cell
                            // before tracing, 2 = shadow rays too
  int leafSize = 3;         // BVH builder: primitives per leaf slot of a wide node (1..3)
  int tlasLeafSize = 1;     // the same for TLAS builds: instances per leaf slot. 1: entering an instance costs far more than a
                            // triangle test, so no instance is entered because it shares a leaf box (4096-instance scene: -10 % frame)
  int plocRadius = 16;      // BVH builder: PLOC search radius; 0 = plain LBVH (Karras) hierarchy
  int tlasPlocRadius = 0;   // the same for TLAS builds; 0 = automatic (wide for small TLASes, bvh_build.cu buildTlas)
  int sampleBatch = 16;     // samples of a pixel in flight at once in the wavefront layout (1 = one sample per pass)
  int blocksPerSm = 0;      // persistent grid of the traversal kernels = smCount * blocksPerSm; 0 = the resident CTA count
                            // of the build the dispatch uses (7 or 8, trace_wavefront.cu)
  std::unordered_map<uint64_t, rtb::AccelObject *> accels;
  // reusable build scratch
  void *scratch = nullptr;
  size_t scratchBytes = 0;
  float *srgbLutDev = nullptr;
  int smCount = 148;
  // wavefront state (trace_wavefront.cu), one per pipeline lane
  void *wfState[rtb::kMaxLanes] = {};
  size_t wfBytes[rtb::kMaxLanes] = {};
  // pipeline lanes: a dispatch is split into `pipelineLanes` interleaved tile subsets whose kernel sequences run on
  // their own streams, so that the tail of one lane's persistent launch is filled by the other lane's next launch
  int classifyRays = 1;  // flat TLAS: queue rays by class (reaches a BLAS with nodes / cheap), long rays first
  int pipelineLanes = 0; // 0 = auto (two lanes for dispatches of >= 16 M paths, else one)
  cudaStream_t laneStream[rtb::kMaxLanes] = {};
  cudaEvent_t evFork = nullptr, evLaneDone[rtb::kMaxLanes] = {};
  // per-light constants derived once per rt_trace (trace.cu k_prepare_lights)
  float4 *lightDerivedDev = nullptr;
  int lightDerivedCap = 0;
  rtb::KernelTimer timer;
  // records an event (only when timing is enabled) on the context's stream (lane < 0) or on a lane's stream;
  // klass < 0 starts a new sequence on that stream. forkFrom >= 0: the stream predecessor is that stream's latest
  // record (a lane's first launch follows the fork point on the context's stream)
  void mark(int klass, int lane = -1, bool afterFork = false);
};

namespace rtb {
int ensureScratch(rt_context *ctx, size_t bytes);
int buildBlas(rt_context *ctx, const rt_triangle_geometry *geoms, uint32_t n, uint32_t flags, AccelObject **out);
int refitBlas(rt_context *ctx, AccelObject *as, const rt_triangle_geometry *geoms, uint32_t n);
int buildTlas(rt_context *ctx, AccelObject *as, const rt_instance_descriptor *descDev, uint32_t count, bool refit);
int checkTlasInfo(rt_context *ctx, AccelObject *as, bool wait);
void destroyAccel(AccelObject *as);
int launchSkin(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], uint32_t vertexCount);
int launchJointPalette(rt_context *ctx, const float *trs, const int32_t *parents, const float *inverseBind,
                       uint32_t jointCount, float *palette);
int launchSpatialFilter(rt_context *ctx, const rt_denoise_frame *fr, const rt_image *out, int step, float depthSigma,
                        int normalSquarings, float colorSigma);
int launchTemporalFilter(rt_context *ctx, const rt_denoise_frame *cur, const rt_denoise_frame *hist, const rt_image *out,
                         float historyWeight, float depthTolerance, float normalThreshold);
int launchTonemap(rt_context *ctx, const rt_image *src, uint8_t *dst, uint32_t flags);
int packTiles(rt_context *ctx, const rt_image *image, void *slab, int modulo, int remainder);
int unpackTiles(rt_context *ctx, const void *slabs, const rt_image *image, int modulo);
int selftestChildBoxes(rt_context *ctx, AccelObject *as, uint32_t raysPerNode, uint32_t seed, unsigned long long outHost[11]);
int launchIntersect(rt_context *ctx, const AccelObject *tl, const rt_ray *rays, uint32_t count, uint32_t flags, rt_ray_hit *hits);
int launchTrace(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], const rt_image textures[RT_TEXTURE_COUNT],
                int maxSubmeshes, const rt_trace_options *opt);
} // namespace rtb
