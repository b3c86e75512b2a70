// skin.cu — linear-blend skinning for sm_100a.
//
// Counterpart of the reference's skinningKernel (MetalRaytracing/Skinning.metal:7-49) behind the same argument
// table (SkinningPass.swift:160-211): 4 influences per vertex, weights used as authored (only an all-zero weight
// vector falls back to joint 0), normals transformed by the same matrices and left unnormalised.
// The pass is pure streaming — 56 B read + 32 B written per vertex — so each thread moves its vertex with
// 128-bit loads/stores; the joint palette (64 B x joints) is staged once per CTA in shared memory when it fits.
// Compile with -fmad=false: skinned positions feed the BVH and must round like the oracle's.
#include "common.cuh"

namespace rtb {

constexpr int kSkinBlock = 256;
constexpr int kMaxSharedJoints = 512; // 32 KB of palette

struct SkinParams {
  const float4 *restPositions;
  const float4 *restNormals;
  const uint2 *jointIndices; // ushort4
  const float4 *jointWeights;
  const float4 *jointMatrices; // 4 x float4 columns per joint
  float4 *skinnedPositions;
  float4 *skinnedNormals;
  uint32_t vertexCount;
  uint32_t jointCount; // 0 = unknown: read the palette from global memory
};

__device__ __forceinline__ void accumulate(const float4 *m, float w, float px, float py, float pz, float nx, float ny,
                                           float nz, float3 &pos, float3 &nrm) {
  const float4 c0 = m[0], c1 = m[1], c2 = m[2], c3 = m[3];
  // (M * (p, 1)).xyz = ((c0*p.x + c1*p.y) + c2*p.z) + c3 ; (M * (n, 0)).xyz = (c0*n.x + c1*n.y) + c2*n.z
  const float tx = ((c0.x * px + c1.x * py) + c2.x * pz) + c3.x;
  const float ty = ((c0.y * px + c1.y * py) + c2.y * pz) + c3.y;
  const float tz = ((c0.z * px + c1.z * py) + c2.z * pz) + c3.z;
  pos.x = pos.x + w * tx;
  pos.y = pos.y + w * ty;
  pos.z = pos.z + w * tz;
  const float ux = (c0.x * nx + c1.x * ny) + c2.x * nz;
  const float uy = (c0.y * nx + c1.y * ny) + c2.y * nz;
  const float uz = (c0.z * nx + c1.z * ny) + c2.z * nz;
  nrm.x = nrm.x + w * ux;
  nrm.y = nrm.y + w * uy;
  nrm.z = nrm.z + w * uz;
}

template <bool kSharedPalette>
__global__ void __launch_bounds__(kSkinBlock) k_skin(const SkinParams P) {
  extern __shared__ float4 s_palette[];
  const float4 *palette = P.jointMatrices;
  if (kSharedPalette) {
    for (uint32_t i = threadIdx.x; i < P.jointCount * 4; i += blockDim.x) s_palette[i] = __ldg(P.jointMatrices + i);
    __syncthreads();
    palette = s_palette;
  }
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < P.vertexCount; v += gridDim.x * blockDim.x) {
    const float4 p = __ldg(P.restPositions + v);
    const float4 n = __ldg(P.restNormals + v);
    const uint2 packed = __ldg(P.jointIndices + v);
    float4 w = __ldg(P.jointWeights + v);
    const uint32_t j0 = packed.x & 0xFFFFu, j1 = packed.x >> 16, j2 = packed.y & 0xFFFFu, j3 = packed.y >> 16;
    const float weightSum = w.x + w.y + w.z + w.w;
    if (weightSum < 0.0001f) w = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
    float3 pos = make_float3(0.0f, 0.0f, 0.0f), nrm = make_float3(0.0f, 0.0f, 0.0f);
    // influence order x, y, z, w: the partial sums must associate like the reference's
    accumulate(palette + 4 * j0, w.x, p.x, p.y, p.z, n.x, n.y, n.z, pos, nrm);
    accumulate(palette + 4 * j1, w.y, p.x, p.y, p.z, n.x, n.y, n.z, pos, nrm);
    accumulate(palette + 4 * j2, w.z, p.x, p.y, p.z, n.x, n.y, n.z, pos, nrm);
    accumulate(palette + 4 * j3, w.w, p.x, p.y, p.z, n.x, n.y, n.z, pos, nrm);
    P.skinnedPositions[v] = make_float4(pos.x, pos.y, pos.z, 0.0f);
    P.skinnedNormals[v] = make_float4(nrm.x, nrm.y, nrm.z, 0.0f);
  }
}

int launchSkinImpl(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], uint32_t vertexCount,
                   uint32_t jointCount) {
  RT_CHECK(buffers != nullptr, "rt_skin: null argument table");
  static const int need[] = {RT_BUFFER_REST_POSITIONS, RT_BUFFER_REST_NORMALS, RT_BUFFER_JOINT_INDICES,
                             RT_BUFFER_JOINT_WEIGHTS,  RT_BUFFER_JOINT_MATRICES, RT_BUFFER_SKINNED_POSITIONS,
                             RT_BUFFER_SKINNED_NORMALS};
  for (int idx : need) RT_CHECK(buffers[idx] != nullptr, "rt_skin: buffer " + std::to_string(idx) + " is not bound");
  if (vertexCount == 0) return 0;
  SkinParams P;
  P.restPositions = static_cast<const float4 *>(buffers[RT_BUFFER_REST_POSITIONS]);
  P.restNormals = static_cast<const float4 *>(buffers[RT_BUFFER_REST_NORMALS]);
  P.jointIndices = static_cast<const uint2 *>(buffers[RT_BUFFER_JOINT_INDICES]);
  P.jointWeights = static_cast<const float4 *>(buffers[RT_BUFFER_JOINT_WEIGHTS]);
  P.jointMatrices = static_cast<const float4 *>(buffers[RT_BUFFER_JOINT_MATRICES]);
  P.skinnedPositions = static_cast<float4 *>(const_cast<void *>(buffers[RT_BUFFER_SKINNED_POSITIONS]));
  P.skinnedNormals = static_cast<float4 *>(const_cast<void *>(buffers[RT_BUFFER_SKINNED_NORMALS]));
  P.vertexCount = vertexCount;
  P.jointCount = jointCount;
  const uint32_t blocksNeeded = (vertexCount + kSkinBlock - 1) / kSkinBlock;
  // grid: whole waves of the SM count (8 resident CTAs of 256 threads per SM), grid-stride beyond that
  const uint32_t wave = uint32_t(ctx->smCount) * 8u;
  const uint32_t grid = blocksNeeded <= wave ? blocksNeeded : wave * ((blocksNeeded + wave - 1) / wave > 4 ? 4 : (blocksNeeded + wave - 1) / wave);
  if (jointCount > 0 && jointCount <= kMaxSharedJoints) {
    k_skin<true><<<grid, kSkinBlock, size_t(jointCount) * 64, ctx->stream>>>(P);
  } else {
    k_skin<false><<<grid, kSkinBlock, 0, ctx->stream>>>(P);
  }
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

int launchSkin(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], uint32_t vertexCount) {
  return launchSkinImpl(ctx, buffers, vertexCount, 0);
}

} // namespace rtb
