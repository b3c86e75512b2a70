// selftest.cu — device-side invariants that tests/ can call through the C-ABI.
// rt_selftest_child_boxes: for pseudo-random rays against every wide node of an acceleration structure, the fast
// child-box test (PRMT form, traverse.cuh) must report every child the plain-conversion form reports.
#include "traverse.cuh"

namespace rtb {

__device__ __forceinline__ uint32_t lcg(uint32_t &s) {
  s = s * 1664525u + 1013904223u;
  return s;
}
__device__ __forceinline__ float unit(uint32_t &s) { return float(lcg(s) >> 8) * (1.0f / 16777216.0f); }

__global__ void k_selftest_child_boxes(const WideNode *nodes, const float4 *nodeBox, uint32_t nodeCount, uint32_t raysPerNode,
                                       uint32_t seed, unsigned long long *out /* [0] missed children, [1] extra, [2] tests, [3..10] first failure */) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nodeCount * raysPerNode) return;
  const uint32_t n = i / raysPerNode;
  uint32_t s = seed ^ (i * 2654435761u);
  const float4 lo = nodeBox[2 * n], hi = nodeBox[2 * n + 1];
  // origin around the node, direction toward a random point inside it (plus axis-aligned special cases)
  const float ex = fmaxf(hi.x - lo.x, 1e-3f), ey = fmaxf(hi.y - lo.y, 1e-3f), ez = fmaxf(hi.z - lo.z, 1e-3f);
  const float ox = lo.x + (unit(s) * 6.0f - 2.5f) * ex, oy = lo.y + (unit(s) * 6.0f - 2.5f) * ey, oz = lo.z + (unit(s) * 6.0f - 2.5f) * ez;
  const float tx = lo.x + unit(s) * ex, ty = lo.y + unit(s) * ey, tz = lo.z + unit(s) * ez;
  float dx = tx - ox, dy = ty - oy, dz = tz - oz;
  const uint32_t special = lcg(s) & 15u;
  if (special == 0) dx = 0.0f;
  if (special == 1) dy = 0.0f;
  if (special == 2) dz = -0.0f;
  const BoxSetup b = makeBoxSetup(ox, oy, oz, dx, dy, dz);
  const uint4 *np = reinterpret_cast<const uint4 *>(nodes) + size_t(n) * 5;
  const uint4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3], n4 = np[4];
  const float tmax = (lcg(s) & 1u) ? INFINITY : unit(s) * 4.0f;
  const uint32_t fast = intersectChildren(n0, n1, n2, n3, n4, b, 0.0f, tmax);
  const uint32_t ref = intersectChildrenLegacy(n0, n1, n2, n3, n4, b, 0.0f, tmax);
  const uint32_t missed = ref & ~fast, extra = fast & ~ref;
  atomicAdd(out + 2, 1ull);
  if (missed) {
    if (atomicAdd(out + 0, (unsigned long long)__popc(missed)) == 0ull) {
      out[3] = n, out[4] = fast, out[5] = ref, out[6] = __float_as_uint(ox), out[7] = __float_as_uint(dx);
      out[8] = n0.w, out[9] = n1.z, out[10] = n1.w;
    }
  }
  if (extra) atomicAdd(out + 1, (unsigned long long)__popc(extra));
}

int selftestChildBoxes(rt_context *ctx, AccelObject *as, uint32_t raysPerNode, uint32_t seed, unsigned long long outHost[11]) {
  RT_CHECK(as->nodeCount > 0, "rt_selftest_child_boxes: empty acceleration structure");
  unsigned long long *out = nullptr;
  RT_CUDA(cudaMalloc(&out, 11 * sizeof(unsigned long long)));
  RT_CUDA(cudaMemsetAsync(out, 0, 11 * sizeof(unsigned long long), ctx->stream));
  const uint32_t total = as->nodeCount * raysPerNode;
  k_selftest_child_boxes<<<(total + 255) / 256, 256, 0, ctx->stream>>>(as->nodes, as->nodeBox, as->nodeCount, raysPerNode, seed, out);
  ++ctx->launches;
  RT_CUDA(cudaMemcpyAsync(outHost, out, 11 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(out);
  return 0;
}

} // namespace rtb
