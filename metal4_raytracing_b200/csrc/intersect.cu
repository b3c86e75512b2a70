// intersect.cu — rt_intersect: a batch of caller-supplied rays against an instance acceleration structure.
//
// The reference calls intersector<triangle_data, instancing>::intersect() per path segment (closest hit,
// MetalRaytracing/Raytracing.metal:301-318) and per shadow ray (any hit, :664-665 and :730-737); this entry point is that
// call on its own, without the path tracer around it: rays in, (t, u, v, instance, geometry, primitive) out. It runs the
// iteration the wavefront traversal kernel runs (LaneTraversal::stepConverged, flat or real TLAS), so a host program —
// and tests/test_gpu_parity.py::test_intersect_against_brute_force, which checks it against a float64 brute force over
// triangles it built itself, with neither the scene library nor the oracle involved — sees exactly the shipped
// traversal.
#include "traverse.cuh"

namespace rtb {
namespace {

constexpr int kIntersectBlock = 128;

template <bool kAny, bool kFlat>
__global__ void __launch_bounds__(kIntersectBlock, 7) k_intersect(const TlasHeader tlas, const rt_ray *__restrict__ rays, uint32_t count,
                                                                  rt_ray_hit *__restrict__ hits) {
  const unsigned full = 0xFFFFFFFFu;
  const uint32_t warpsPerGrid = gridDim.x * (kIntersectBlock / 32);
  const uint32_t lane = threadIdx.x & 31u;
  LaneTraversal<kAny> t;
  LocalStack stack;
  // whole warps stay in the loop: 32 consecutive rays per warp and round
  for (uint32_t base = (blockIdx.x * (kIntersectBlock / 32) + (threadIdx.x >> 5)) * 32u; base < count; base += warpsPerGrid * 32u) {
    const uint32_t i = base + lane;
    bool active = i < count;
    if (active) {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(rays + i)), b = __ldg(reinterpret_cast<const float4 *>(rays + i) + 1);
      t.template begin<kFlat>(tlas, stack, a.x, a.y, a.z, b.x, b.y, b.z, a.w, b.w); // origin, tmin | direction, tmax
    }
    while (__ballot_sync(full, active) != 0u) {
      if (t.template stepConverged<kFlat, !kFlat, true, false, 2, 0>(tlas, stack, active, nullptr)) {
        const RayHit h = t.result(stack);
        rt_ray_hit out;
        out.t = t.found ? (kAny ? 0.0f : h.t) : INFINITY;
        out.u = h.u, out.v = h.v;
        out.instance = t.found && !kAny ? h.instance : 0xFFFFFFFFu;
        out.geometry = t.found && !kAny ? h.geometry : 0xFFFFFFFFu;
        out.primitive = t.found && !kAny ? h.primitive : 0xFFFFFFFFu;
        hits[i] = out;
        active = false;
      }
    }
  }
}

} // namespace

int launchIntersect(rt_context *ctx, const AccelObject *tl, const rt_ray *rays, uint32_t count, uint32_t flags, rt_ray_hit *hits) {
  if (count == 0) return 0;
  TlasHeader H{};
  H.nodes = tl->nodes;
  H.instances = tl->instances;
  H.leafInstance = tl->leafPrim;
  H.instanceBox = tl->instanceBox;
  H.instanceCount = tl->primCount;
  H.nodeCount = tl->primCount ? tl->nodeCount : 0u;
  const bool flat = H.instanceCount <= kFlatTlasMax && H.instanceBox != nullptr;
  const bool any = (flags & RT_INTERSECT_ANY) != 0u;
  const uint32_t warps = (count + 31u) / 32u;
  const int grid = int(std::min<uint32_t>((warps + kIntersectBlock / 32 - 1) / (kIntersectBlock / 32), uint32_t(ctx->smCount) * 7u));
  ctx->mark(-1);
  if (any && flat) k_intersect<true, true><<<grid, kIntersectBlock, 0, ctx->stream>>>(H, rays, count, hits);
  else if (any) k_intersect<true, false><<<grid, kIntersectBlock, 0, ctx->stream>>>(H, rays, count, hits);
  else if (flat) k_intersect<false, true><<<grid, kIntersectBlock, 0, ctx->stream>>>(H, rays, count, hits);
  else k_intersect<false, false><<<grid, kIntersectBlock, 0, ctx->stream>>>(H, rays, count, hits);
  ctx->mark(RT_KERNEL_OTHER);
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

} // namespace rtb
