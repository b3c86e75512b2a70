// rt_api.cu — extern "C" entry points of librt_b200.so (include/rt_b200.h).
// Thin: argument validation, handle bookkeeping, stream plumbing. No CPU fallback anywhere: every compute entry
// point enqueues CUDA kernels or fails with an error code.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace rtb {

static thread_local std::string g_lastError;
void setError(const std::string &msg) { g_lastError = msg; }

int ensureScratch(rt_context *ctx, size_t bytes) {
  if (bytes <= ctx->scratchBytes) return 0;
  RT_CUDA(RT_SYNC_STREAM(ctx, ctx->stream));
  if (ctx->scratch) cudaFree(ctx->scratch);
  ctx->scratch = nullptr;
  ctx->scratchBytes = 0;
  size_t want = bytes + bytes / 4;
  RT_CUDA(cudaMalloc(&ctx->scratch, want));
  ctx->scratchBytes = want;
  return 0;
}

} // namespace rtb

void rt_context::mark(int klass, int lane, bool afterFork) {
  if (!timer.enabled) return;
  if (timer.used == timer.pool.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    timer.pool.push_back(e);
    timer.klass.push_back(-1);
    timer.prev.push_back(-1);
  }
  const int slot = lane < 0 ? 0 : 1 + lane;
  timer.klass[timer.used] = klass;
  timer.prev[timer.used] = klass < 0 ? -1 : (afterFork ? timer.last[0] : timer.last[slot]);
  cudaEventRecord(timer.pool[timer.used], lane < 0 ? stream : laneStream[lane]);
  timer.last[slot] = int(timer.used);
  ++timer.used;
}

using namespace rtb;

#define RT_CTX(ctx)                             \
  do {                                          \
    if (!(ctx)) {                               \
      rtb::setError("null rt_context");         \
      return 3;                                 \
    }                                           \
    cudaError_t _e = cudaSetDevice((ctx)->device); \
    if (_e != cudaSuccess) {                    \
      rtb::setError(std::string("cudaSetDevice failed: ") + cudaGetErrorString(_e)); \
      return 1;                                 \
    }                                           \
  } while (0)

extern "C" {

const char *rt_last_error(void) { return g_lastError.c_str(); }

int rt_create(int device, rt_context **out) {
  RT_CHECK(out != nullptr, "rt_create: null out pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    setError(std::string("rt_create: no CUDA device available (") + cudaGetErrorString(e) +
             "); this library has no CPU path");
    return 1;
  }
  RT_CHECK(device >= 0 && device < count, "rt_create: device index out of range");
  RT_CUDA(cudaSetDevice(device));
  rt_context *ctx = new rt_context();
  ctx->device = device;
  cudaDeviceProp prop{};
  RT_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->smCount = prop.multiProcessorCount;
  RT_CUDA(cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking));
  ctx->stream = ctx->ownStream;
  RT_CUDA(cudaEventCreate(&ctx->evBegin));
  RT_CUDA(cudaEventCreate(&ctx->evEnd));
  RT_CUDA(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
  RT_CUDA(cudaEventCreateWithFlags(&ctx->evReady, cudaEventDisableTiming));
  for (cudaEvent_t &e : ctx->evCopied) RT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (cudaEvent_t &e : ctx->evFence) RT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (cudaStream_t &s : ctx->laneStream) RT_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  RT_CUDA(cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
  for (cudaEvent_t &e : ctx->evLaneDone) RT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  // sRGB decode table, evaluated in double and rounded once (same table as the oracle's)
  float lut[256];
  for (int i = 0; i < 256; ++i) {
    double c = double(i) / 255.0;
    lut[i] = float(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
  }
  RT_CUDA(cudaMalloc(&ctx->srgbLutDev, sizeof lut));
  RT_CUDA(cudaMemcpy(ctx->srgbLutDev, lut, sizeof lut, cudaMemcpyHostToDevice));
  *out = ctx;
  // tuning overrides for experiments: RT_B200_OPTIONS="key=value,key=value" (same keys as rt_set_option)
  if (const char *env = std::getenv("RT_B200_OPTIONS")) {
    std::string all(env);
    size_t pos = 0;
    while (pos < all.size()) {
      size_t end = all.find(',', pos);
      if (end == std::string::npos) end = all.size();
      const std::string item = all.substr(pos, end - pos);
      const size_t eq = item.find('=');
      if (eq != std::string::npos && rt_set_option(ctx, item.substr(0, eq).c_str(), std::atoi(item.c_str() + eq + 1)) != 0) {
        rt_destroy(ctx);
        *out = nullptr;
        return 2;
      }
      pos = end + 1;
    }
  }
  return 0;
}

int rt_destroy(rt_context *ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &kv : ctx->accels) destroyAccel(kv.second);
  ctx->accels.clear();
  cudaFree(ctx->scratch);
  cudaFree(ctx->srgbLutDev);
  for (cudaStream_t s : ctx->laneStream) if (s) cudaStreamSynchronize(s);
  for (void *p : ctx->wfState) cudaFree(p);
  for (cudaStream_t s : ctx->laneStream) if (s) cudaStreamDestroy(s);
  if (ctx->evFork) cudaEventDestroy(ctx->evFork);
  for (cudaEvent_t e : ctx->evLaneDone) if (e) cudaEventDestroy(e);
  cudaFree(ctx->lightDerivedDev);
  cudaEventDestroy(ctx->evBegin);
  cudaEventDestroy(ctx->evEnd);
  cudaStreamSynchronize(ctx->copyStream);
  cudaEventDestroy(ctx->evReady);
  for (cudaEvent_t e : ctx->evCopied) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->evFence) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->copyStream);
  for (cudaEvent_t e : ctx->timer.pool) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->ownStream);
  delete ctx;
  return 0;
}

int rt_set_stream(rt_context *ctx, void *cudaStream) {
  RT_CTX(ctx);
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cudaStream ? static_cast<cudaStream_t>(cudaStream) : ctx->ownStream;
  return 0;
}

int rt_get_stream(rt_context *ctx, void **cudaStreamOut) {
  RT_CTX(ctx);
  RT_CHECK(cudaStreamOut != nullptr, "rt_get_stream: null out pointer");
  *cudaStreamOut = static_cast<void *>(ctx->stream);
  return 0;
}

int rt_sync(rt_context *ctx) {
  RT_CTX(ctx);
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int rt_timer_begin(rt_context *ctx) {
  RT_CTX(ctx);
  RT_CUDA(cudaEventRecord(ctx->evBegin, ctx->stream));
  return 0;
}

int rt_timer_end(rt_context *ctx, float *milliseconds) {
  RT_CTX(ctx);
  RT_CUDA(cudaEventRecord(ctx->evEnd, ctx->stream));
  RT_CUDA(cudaEventSynchronize(ctx->evEnd));
  float ms = 0.0f;
  RT_CUDA(cudaEventElapsedTime(&ms, ctx->evBegin, ctx->evEnd));
  if (milliseconds) *milliseconds = ms;
  return 0;
}

int rt_malloc(rt_context *ctx, size_t bytes, void **dev) {
  RT_CTX(ctx);
  RT_CHECK(dev != nullptr, "rt_malloc: null out pointer");
  RT_CUDA(cudaMalloc(dev, bytes ? bytes : 16));
  return 0;
}

int rt_free(rt_context *ctx, void *dev) {
  RT_CTX(ctx);
  if (dev) {
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    RT_CUDA(cudaFree(dev));
  }
  return 0;
}

int rt_malloc_host(rt_context *ctx, size_t bytes, void **pinnedHost) {
  RT_CTX(ctx);
  RT_CHECK(pinnedHost != nullptr, "rt_malloc_host: null out pointer");
  RT_CUDA(cudaMallocHost(pinnedHost, bytes ? bytes : 16));
  return 0;
}

int rt_free_host(rt_context *ctx, void *pinnedHost) {
  RT_CTX(ctx);
  if (pinnedHost) RT_CUDA(cudaFreeHost(pinnedHost));
  return 0;
}

int rt_upload(rt_context *ctx, void *dstDev, const void *srcHost, size_t bytes) {
  RT_CTX(ctx);
  if (!bytes) return 0;
  RT_CHECK(dstDev && srcHost, "rt_upload: null pointer");
  RT_CUDA(cudaMemcpyAsync(dstDev, srcHost, bytes, cudaMemcpyHostToDevice, ctx->stream));
  cudaPointerAttributes attr{};
  bool pinned = cudaPointerGetAttributes(&attr, srcHost) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (!pinned) RT_CUDA(RT_SYNC_STREAM(ctx, ctx->stream)); // pageable source: caller may reuse it right away
  return 0;
}

int rt_download(rt_context *ctx, void *dstHost, const void *srcDev, size_t bytes) {
  RT_CTX(ctx);
  if (!bytes) return 0;
  RT_CHECK(dstHost && srcDev, "rt_download: null pointer");
  RT_CUDA(cudaMemcpyAsync(dstHost, srcDev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int rt_fence(rt_context *ctx, uint64_t *ticket) {
  RT_CTX(ctx);
  RT_CHECK(ticket != nullptr, "rt_fence: null ticket");
  const uint64_t id = ctx->fencesIssued;
  RT_CUDA(cudaEventRecord(ctx->evFence[id % 16], ctx->stream));
  ctx->fencesIssued = id + 1;
  *ticket = id + 1;
  return 0;
}

int rt_fence_wait(rt_context *ctx, uint64_t ticket) {
  RT_CTX(ctx);
  RT_CHECK(ticket >= 1 && ticket <= ctx->fencesIssued, "rt_fence_wait: unknown ticket");
  // a ticket older than the ring had its slot re-recorded by a later fence of the same stream: waiting for that later
  // point covers the older one, so the wait below is right in both cases (it never returns before the work is done)
  RT_CUDA(cudaEventSynchronize(ctx->evFence[(ticket - 1) % 16]));
  return 0;
}

int rt_download_async(rt_context *ctx, void *dstHost, const void *srcDev, size_t bytes, uint64_t *ticket) {
  RT_CTX(ctx);
  RT_CHECK(dstHost && srcDev && ticket && bytes, "rt_download_async: null pointer or empty copy");
  const uint64_t id = ctx->copiesIssued;
  if (id >= 8) RT_CUDA(RT_SYNC_EVENT(ctx, ctx->evCopied[id % 8])); // the ring slot's previous copy must be done
  RT_CUDA(cudaEventRecord(ctx->evReady, ctx->stream));
  RT_CUDA(cudaStreamWaitEvent(ctx->copyStream, ctx->evReady, 0));
  RT_CUDA(cudaMemcpyAsync(dstHost, srcDev, bytes, cudaMemcpyDeviceToHost, ctx->copyStream));
  RT_CUDA(cudaEventRecord(ctx->evCopied[id % 8], ctx->copyStream));
  ctx->copiesIssued = id + 1;
  *ticket = id + 1;
  return 0;
}

int rt_download_wait(rt_context *ctx, uint64_t ticket) {
  RT_CTX(ctx);
  RT_CHECK(ticket >= 1 && ticket <= ctx->copiesIssued, "rt_download_wait: unknown ticket");
  if (ctx->copiesIssued - ticket >= 8) return 0; // older than the ring: finished when its slot was reused
  // copies complete in issue order on the copy stream, so waiting for this one covers the earlier ones
  RT_CUDA(cudaEventSynchronize(ctx->evCopied[(ticket - 1) % 8]));
  return 0;
}

int rt_copy(rt_context *ctx, void *dstDev, const void *srcDev, size_t bytes) {
  RT_CTX(ctx);
  if (!bytes) return 0;
  RT_CHECK(dstDev && srcDev, "rt_copy: null pointer");
  RT_CUDA(cudaMemcpyAsync(dstDev, srcDev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

int rt_memset(rt_context *ctx, void *dstDev, int value, size_t bytes) {
  RT_CTX(ctx);
  if (!bytes) return 0;
  RT_CHECK(dstDev != nullptr, "rt_memset: null pointer");
  RT_CUDA(cudaMemsetAsync(dstDev, value, bytes, ctx->stream));
  return 0;
}

int rt_blas_build(rt_context *ctx, const rt_triangle_geometry *geoms, uint32_t geometryCount, uint32_t flags,
                  uint64_t *outId) {
  RT_CTX(ctx);
  RT_CHECK(outId != nullptr, "rt_blas_build: null out pointer");
  RT_CHECK(geometryCount == 0 || geoms != nullptr, "rt_blas_build: null geometry array");
  AccelObject *as = nullptr;
  ctx->mark(-1);
  RT_TRY(buildBlas(ctx, geoms, geometryCount, flags, &as));
  ctx->mark(RT_KERNEL_BUILD);
  uint64_t id = uint64_t(reinterpret_cast<uintptr_t>(as->headerDev));
  ctx->accels[id] = as;
  *outId = id;
  return 0;
}

static int findAccel(rt_context *ctx, uint64_t id, bool tlas, AccelObject **out) {
  auto it = ctx->accels.find(id);
  RT_CHECK(it != ctx->accels.end(), "unknown acceleration structure id");
  RT_CHECK(it->second->isTlas == tlas, tlas ? "id is not a TLAS" : "id is not a BLAS");
  *out = it->second;
  return 0;
}

int rt_blas_refit(rt_context *ctx, uint64_t id, const rt_triangle_geometry *geoms, uint32_t geometryCount) {
  RT_CTX(ctx);
  AccelObject *as = nullptr;
  RT_TRY(findAccel(ctx, id, false, &as));
  ctx->mark(-1);
  const int rc = refitBlas(ctx, as, geoms, geometryCount);
  ctx->mark(RT_KERNEL_REFIT);
  return rc;
}

int rt_blas_destroy(rt_context *ctx, uint64_t id) {
  RT_CTX(ctx);
  AccelObject *as = nullptr;
  RT_TRY(findAccel(ctx, id, false, &as));
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->accels.erase(id);
  destroyAccel(as);
  return 0;
}

int rt_tlas_build(rt_context *ctx, const rt_instance_descriptor *descriptorsDev, uint32_t count, uint64_t *outId) {
  RT_CTX(ctx);
  RT_CHECK(outId != nullptr, "rt_tlas_build: null out pointer");
  AccelObject *as = new AccelObject();
  as->isTlas = true;
  ctx->mark(-1);
  int rc = buildTlas(ctx, as, descriptorsDev, count, false);
  ctx->mark(RT_KERNEL_BUILD);
  if (rc) {
    destroyAccel(as);
    return rc;
  }
  uint64_t id = uint64_t(reinterpret_cast<uintptr_t>(as->headerDev));
  ctx->accels[id] = as;
  *outId = id;
  return 0;
}

int rt_tlas_update(rt_context *ctx, uint64_t id, const rt_instance_descriptor *descriptorsDev, uint32_t count) {
  RT_CTX(ctx);
  AccelObject *as = nullptr;
  RT_TRY(findAccel(ctx, id, true, &as));
  ctx->mark(-1);
  const int rc = buildTlas(ctx, as, descriptorsDev, count, false);
  ctx->mark(RT_KERNEL_BUILD);
  return rc;
}

int rt_tlas_refit(rt_context *ctx, uint64_t id, const rt_instance_descriptor *descriptorsDev, uint32_t count) {
  RT_CTX(ctx);
  AccelObject *as = nullptr;
  RT_TRY(findAccel(ctx, id, true, &as));
  RT_CHECK(as->treeValid && count == as->primCount,
           "rt_tlas_refit: instance count differs from the last build (rebuild with rt_tlas_update)");
  ctx->mark(-1);
  const int rc = buildTlas(ctx, as, descriptorsDev, count, true);
  ctx->mark(RT_KERNEL_REFIT);
  return rc;
}

int rt_tlas_destroy(rt_context *ctx, uint64_t id) {
  RT_CTX(ctx);
  AccelObject *as = nullptr;
  RT_TRY(findAccel(ctx, id, true, &as));
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->accels.erase(id);
  destroyAccel(as);
  return 0;
}

int rt_as_get_info(rt_context *ctx, uint64_t id, rt_as_info *out) {
  RT_CTX(ctx);
  RT_CHECK(out != nullptr, "rt_as_get_info: null out pointer");
  auto it = ctx->accels.find(id);
  RT_CHECK(it != ctx->accels.end(), "unknown acceleration structure id");
  AccelObject *as = it->second;
  std::memset(out, 0, sizeof *out);
  out->primitiveCount = as->primCount;
  out->wideNodeCount = as->nodeCount;
  out->levelCount = as->levelStart.empty() ? 0 : uint32_t(as->levelStart.size() - 1);
  if (as->isTlas && as->deviceBuilt) { // counts of a device-side build live on the device: wait for their copy
    RT_TRY(checkTlasInfo(ctx, as, true));
    out->wideNodeCount = as->infoHost->nodeCount;
    out->levelCount = as->infoHost->levelCount;
  }
  out->bytes = as->bytes;
  for (int a = 0; a < 3; ++a) {
    out->boundsMin[a] = as->bounds.lo[a];
    out->boundsMax[a] = as->bounds.hi[a];
  }
  out->sahCost = as->sahCost;
  return 0;
}

int rt_skin(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], uint32_t vertexCount) {
  RT_CTX(ctx);
  ctx->mark(-1);
  const int rc = launchSkin(ctx, buffers, vertexCount);
  ctx->mark(RT_KERNEL_SKIN);
  return rc;
}

int rt_joint_palette(rt_context *ctx, const float *localTRSDev, const int32_t *parentsDev, const float *inverseBindDev,
                     uint32_t jointCount, float *paletteOutDev) {
  RT_CTX(ctx);
  ctx->mark(-1);
  const int rc = launchJointPalette(ctx, localTRSDev, parentsDev, inverseBindDev, jointCount, paletteOutDev);
  ctx->mark(RT_KERNEL_SKIN);
  return rc;
}

int rt_trace(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], const rt_image textures[RT_TEXTURE_COUNT],
             int resourcesStride, int maxSubmeshes, const rt_trace_options *options) {
  RT_CTX(ctx);
  (void)resourcesStride; // function constant 0 is declared but unused by the reference kernel as well
  return launchTrace(ctx, buffers, textures, maxSubmeshes, options);
}

int rt_texture_create(rt_context *ctx, const uint8_t *rgba8Host, int width, int height, int srgb,
                      const rt_texture2d **outRecordDev) {
  RT_CTX(ctx);
  RT_CHECK(rgba8Host && outRecordDev && width > 0 && height > 0, "rt_texture_create: bad arguments");
  size_t bytes = size_t(width) * size_t(height) * 4;
  // one allocation: [record (32 B, padded)] [texels]
  uint8_t *block = nullptr;
  RT_CUDA(cudaMalloc(&block, 32 + bytes));
  rt_texture2d rec{};
  rec.texels = block + 32;
  rec.width = width;
  rec.height = height;
  rec.srgb = srgb ? 1 : 0;
  RT_CUDA(cudaMemcpyAsync(block, &rec, sizeof rec, cudaMemcpyHostToDevice, ctx->stream));
  RT_CUDA(cudaMemcpyAsync(block + 32, rgba8Host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  *outRecordDev = reinterpret_cast<const rt_texture2d *>(block);
  return 0;
}

int rt_texture_destroy(rt_context *ctx, const rt_texture2d *recordDev) {
  RT_CTX(ctx);
  if (recordDev) {
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    RT_CUDA(cudaFree(const_cast<rt_texture2d *>(recordDev)));
  }
  return 0;
}

int rt_tonemap(rt_context *ctx, const rt_image *srcDev, uint8_t *dstRGBA8Dev, uint32_t flags) {
  RT_CTX(ctx);
  ctx->mark(-1);
  const int rc = launchTonemap(ctx, srcDev, dstRGBA8Dev, flags);
  ctx->mark(RT_KERNEL_OTHER);
  return rc;
}

int rt_temporal_filter(rt_context *ctx, const rt_denoise_frame *current, const rt_denoise_frame *history,
                       const rt_image *outColorDev, float historyWeight, float depthTolerance, float normalThreshold) {
  RT_CTX(ctx);
  ctx->mark(-1);
  const int rc = launchTemporalFilter(ctx, current, history, outColorDev, historyWeight, depthTolerance, normalThreshold);
  ctx->mark(RT_KERNEL_OTHER);
  return rc;
}

int rt_spatial_filter(rt_context *ctx, const rt_denoise_frame *frame, const rt_image *outColorDev, int step,
                      float depthSigma, int normalSquarings, float colorSigma) {
  RT_CTX(ctx);
  ctx->mark(-1);
  const int rc = launchSpatialFilter(ctx, frame, outColorDev, step, depthSigma, normalSquarings, colorSigma);
  ctx->mark(RT_KERNEL_OTHER);
  return rc;
}

int rt_pack_tiles(rt_context *ctx, const rt_image *imageDev, void *slabDev, int tileModulo, int tileRemainder) {
  RT_CTX(ctx);
  return packTiles(ctx, imageDev, slabDev, tileModulo, tileRemainder);
}

int rt_unpack_tiles(rt_context *ctx, const void *slabsDev, const rt_image *imageDev, int tileModulo) {
  RT_CTX(ctx);
  return unpackTiles(ctx, slabsDev, imageDev, tileModulo);
}

int rt_ipc_export(rt_context *ctx, void *dev, unsigned char handle[64]) {
  RT_CTX(ctx);
  RT_CHECK(dev && handle, "rt_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  RT_CUDA(cudaIpcGetMemHandle(&h, dev));
  std::memcpy(handle, &h, 64);
  return 0;
}

int rt_ipc_import(rt_context *ctx, const unsigned char handle[64], void **outDev) {
  RT_CTX(ctx);
  RT_CHECK(handle && outDev, "rt_ipc_import: null pointer");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  RT_CUDA(cudaIpcOpenMemHandle(outDev, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int rt_ipc_close(rt_context *ctx, void *importedDev) {
  RT_CTX(ctx);
  if (importedDev) RT_CUDA(cudaIpcCloseMemHandle(importedDev));
  return 0;
}

uint64_t rt_launch_count(rt_context *ctx) { return ctx ? ctx->launches : 0; }
uint64_t rt_host_sync_count(rt_context *ctx) { return ctx ? ctx->hostSyncs : 0; }

int rt_set_trace_mode(rt_context *ctx, int mode) {
  RT_CTX(ctx);
  RT_CHECK(mode == 0 || mode == 1, "rt_set_trace_mode: 0 = megakernel, 1 = wavefront");
  ctx->traceMode = mode;
  return 0;
}

int rt_kernel_timing_enable(rt_context *ctx, int enable) {
  RT_CTX(ctx);
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->timer.enabled = enable != 0;
  ctx->timer.used = 0;
  for (int &l : ctx->timer.last) l = -1;
  return 0;
}

int rt_kernel_timing_read(rt_context *ctx, float msByClass[RT_KERNEL_CLASS_COUNT],
                          uint32_t launchesByClass[RT_KERNEL_CLASS_COUNT]) {
  RT_CTX(ctx);
  RT_CHECK(msByClass != nullptr, "rt_kernel_timing_read: null output");
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < RT_KERNEL_CLASS_COUNT; ++k) {
    msByClass[k] = 0.0f;
    if (launchesByClass) launchesByClass[k] = 0;
  }
  for (cudaStream_t s : ctx->laneStream) RT_CUDA(cudaStreamSynchronize(s));
  KernelTimer &T = ctx->timer;
  if (T.used > 0) {
    // completion time of every record relative to the first one, then the records in completion order
    std::vector<double> at(T.used, 0.0);
    for (size_t i = 1; i < T.used; ++i) {
      float ms = 0.0f;
      RT_CUDA(cudaEventElapsedTime(&ms, T.pool[0], T.pool[i]));
      at[i] = double(ms);
    }
    std::vector<uint32_t> order(T.used);
    for (size_t i = 0; i < T.used; ++i) order[i] = uint32_t(i);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return at[a] < at[b]; });
    double latestEnd = 0.0; // latest completion seen so far among launches (class >= 0)
    for (size_t r = 0; r < T.used; ++r) {
      const uint32_t i = order[r];
      const int k = T.klass[i];
      if (k >= 0 && k < RT_KERNEL_CLASS_COUNT && T.prev[i] >= 0) {
        const double begin = std::max(at[size_t(T.prev[i])], latestEnd);
        if (at[i] > begin) msByClass[k] += float(at[i] - begin);
        if (launchesByClass) ++launchesByClass[k];
      }
      if (k >= 0) latestEnd = std::max(latestEnd, at[i]);
    }
  }
  T.used = 0;
  for (int &l : T.last) l = -1;
  return 0;
}

int rt_intersect(rt_context *ctx, uint64_t tlasId, const rt_ray *raysDev, uint32_t count, uint32_t flags, rt_ray_hit *hitsDev) {
  RT_CTX(ctx);
  RT_CHECK(count == 0 || (raysDev != nullptr && hitsDev != nullptr), "rt_intersect: null ray or hit array");
  RT_CHECK((reinterpret_cast<uintptr_t>(raysDev) & 15u) == 0, "rt_intersect: raysDev must be 16-byte aligned");
  RT_CHECK((flags & ~RT_INTERSECT_ANY) == 0u, "rt_intersect: unknown flag");
  auto it = ctx->accels.find(tlasId);
  RT_CHECK(it != ctx->accels.end() && it->second->isTlas, "rt_intersect: not a TLAS id from rt_tlas_build");
  RT_TRY(checkTlasInfo(ctx, it->second, false));
  return launchIntersect(ctx, it->second, raysDev, count, flags, hitsDev);
}

int rt_selftest_child_boxes(rt_context *ctx, uint64_t id, uint32_t raysPerNode, uint32_t seed, uint64_t out[11]) {
  RT_CTX(ctx);
  RT_CHECK(out != nullptr && raysPerNode > 0, "rt_selftest_child_boxes: bad arguments");
  auto it = ctx->accels.find(id);
  RT_CHECK(it != ctx->accels.end(), "unknown acceleration structure id");
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "u64");
  return selftestChildBoxes(ctx, it->second, raysPerNode, seed, reinterpret_cast<unsigned long long *>(out));
}

size_t rt_environment_cdf_floats(int32_t width, int32_t height) {
  if (width <= 0 || height <= 0) return 0;
  // marginal + conditional rows, then the guide tables (RT_ENV_GUIDED): 65 entries for the marginal and for every row
  return size_t(height + 1) + size_t(height) * size_t(width + 1) + size_t(RT_ENV_GUIDE_CELLS + 1) * size_t(height + 1);
}

// guide[k] = largest i in [0, n] with c[i] <= k / RT_ENV_GUIDE_CELLS (float comparison, as the device search does)
static void buildGuide(const float *c, int n, float *guideOut) {
  int i = 0;
  for (int k = 0; k <= RT_ENV_GUIDE_CELLS; ++k) {
    const float bound = float(k) / float(RT_ENV_GUIDE_CELLS);
    while (i + 1 <= n && c[i + 1] <= bound) ++i;
    const uint32_t v = uint32_t(i);
    std::memcpy(guideOut + k, &v, 4);
  }
}

// rt_b200.h RT_ENV_IMPORTANCE: marginal over rows + one conditional per row, running sums in double
int rt_environment_cdf(const float *texelsHost, int32_t width, int32_t height, float *cdfOutHost) {
  if (!texelsHost || !cdfOutHost || width <= 0 || height <= 0) return 1; // no context to hold an error string
  const double pi = 3.14159265358979323846;
  float *marginal = cdfOutHost;
  float *rows = cdfOutHost + (height + 1);
  std::vector<double> rowSum(size_t(height), 0.0), weight(size_t(width), 0.0);
  double total = 0.0;
  for (int y = 0; y < height; ++y) {
    const double sinTheta = std::sin(pi * (double(y) + 0.5) / double(height));
    double sum = 0.0;
    for (int x = 0; x < width; ++x) {
      const float *t = texelsHost + (size_t(y) * size_t(width) + size_t(x)) * 4;
      const double lum = 0.2126 * double(t[0]) + 0.7152 * double(t[1]) + 0.0722 * double(t[2]);
      weight[size_t(x)] = (lum > 0.0 ? lum : 0.0) * sinTheta;
      sum += weight[size_t(x)];
    }
    rowSum[size_t(y)] = sum;
    total += sum;
    float *row = rows + size_t(y) * size_t(width + 1);
    row[0] = 0.0f;
    double run = 0.0;
    for (int x = 0; x < width; ++x) {
      run += weight[size_t(x)];
      row[x + 1] = sum > 0.0 ? float(run / sum) : float(double(x + 1) / double(width));
    }
    row[width] = 1.0f;
  }
  marginal[0] = 0.0f;
  double run = 0.0;
  for (int y = 0; y < height; ++y) {
    run += rowSum[size_t(y)];
    marginal[y + 1] = total > 0.0 ? float(run / total) : float(double(y + 1) / double(height));
  }
  marginal[height] = 1.0f;
  // guide tables: where a search for xi may start (the device narrows [lo, hi) to the cells of xi's 1/64 bucket first)
  float *guides = cdfOutHost + (height + 1) + size_t(height) * size_t(width + 1);
  buildGuide(marginal, height, guides);
  for (int y = 0; y < height; ++y)
    buildGuide(rows + size_t(y) * size_t(width + 1), width, guides + size_t(y + 1) * size_t(RT_ENV_GUIDE_CELLS + 1));
  return 0;
}

int rt_set_option(rt_context *ctx, const char *key, int value) {
  RT_CTX(ctx);
  RT_CHECK(key != nullptr, "rt_set_option: null key");
  const std::string k(key);
  if (k == "trace_mode") return rt_set_trace_mode(ctx, value);
  if (k == "traversal_variant") {
    RT_CHECK(value >= 0 && value <= 2, "rt_set_option: traversal_variant is 0..2 (idle lanes refilled never / at 8 / at 16)");
    ctx->traversalVariant = value;
    return 0;
  }
  if (k == "fuse_traversal") {
    RT_CHECK(value == 0 || value == 1, "rt_set_option: fuse_traversal is 0 or 1");
    ctx->fuseTraversal = value;
    return 0;
  }
  if (k == "sort_rays") {
    RT_CHECK(value >= 0 && value <= 2, "rt_set_option: sort_rays is 0 (off), 1 (bounce rays) or 2 (bounce + shadow rays)");
    ctx->sortRays = value;
    return 0;
  }
  if (k == "leaf_size") {
    RT_CHECK(value >= 1 && value <= 3, "rt_set_option: leaf_size is 1..3 primitives per leaf slot");
    ctx->leafSize = value;
    return 0;
  }
  if (k == "tlas_leaf_size") {
    RT_CHECK(value >= 1 && value <= 3, "rt_set_option: tlas_leaf_size is 1..3 instances per leaf slot");
    ctx->tlasLeafSize = value;
    return 0;
  }
  if (k == "ploc_radius") {
    RT_CHECK(value >= 0 && value <= 256, "rt_set_option: ploc_radius is 0 (LBVH) .. 256");
    ctx->plocRadius = value;
    return 0;
  }
  if (k == "tlas_ploc_radius") {
    RT_CHECK(value >= 0 && value <= 65536, "rt_set_option: tlas_ploc_radius is 0 (automatic) .. 65536");
    ctx->tlasPlocRadius = value;
    return 0;
  }
  if (k == "sample_batch") {
    RT_CHECK(value >= 1 && value <= 64, "rt_set_option: sample_batch is 1..64");
    ctx->sampleBatch = value;
    return 0;
  }
  if (k == "pipeline_lanes") {
    RT_CHECK(value >= 0 && value <= rtb::kMaxLanes, "rt_set_option: pipeline_lanes is 0 (auto) or 1..4");
    ctx->pipelineLanes = value;
    return 0;
  }
  if (k == "classify_rays") {
    RT_CHECK(value == 0 || value == 1, "rt_set_option: classify_rays is 0 or 1");
    ctx->classifyRays = value;
    return 0;
  }
  if (k == "blocks_per_sm") {
    RT_CHECK(value >= 0 && value <= 32, "rt_set_option: blocks_per_sm is 0 (automatic) or 1..32");
    ctx->blocksPerSm = value;
    return 0;
  }
  setError("rt_set_option: unknown key " + k);
  return 2;
}

} // extern "C"
