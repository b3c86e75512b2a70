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

// ---- joint palette on the device (rt_joint_palette) -------------------------------------------------------------
struct M4 {
  float m[16]; // column-major, m[col * 4 + row]
};
// r = a * b with the accumulation order of the host helper (csrc/host/hostmath.h mul): s = 0; s += a(k,row) * b(c,k)
__device__ __forceinline__ void mul44(const M4 &a, const M4 &b, M4 &r) {
  for (int c = 0; c < 4; ++c)
    for (int row = 0; row < 4; ++row) {
      float s = 0.0f;
      for (int k = 0; k < 4; ++k) s = s + a.m[k * 4 + row] * b.m[c * 4 + k];
      r.m[c * 4 + row] = s;
    }
}

// One CTA, one thread per joint. The hierarchy is resolved level by level (a joint's level = number of ancestors),
// which applies exactly the products the reference's sequential parents-first loop applies.
__global__ void __launch_bounds__(1024) k_joint_palette(const float *__restrict__ trs, const int32_t *__restrict__ parents,
                                                        const float *__restrict__ inverseBind, uint32_t jointCount,
                                                        float *__restrict__ palette) {
  extern __shared__ float s_global[]; // jointCount x 16
  __shared__ int s_maxLevel;
  const uint32_t j = threadIdx.x;
  if (j == 0) s_maxLevel = 0;
  __syncthreads();
  M4 local{};
  int parent = -1, level = 0;
  if (j < jointCount) {
    const float *t = trs + size_t(j) * 10;
    float qx = t[3], qy = t[4], qz = t[5], qw = t[6];
    const float ql = sqrtf(((qw * qw + qx * qx) + qy * qy) + qz * qz);
    if (ql > 0.0001f) {
      qx = qx / ql, qy = qy / ql, qz = qz / ql, qw = qw / ql;
    } else {
      qx = qy = qz = 0.0f, qw = 1.0f;
    }
    M4 T{}, R{}, S{}, TR;
    T.m[0] = T.m[5] = T.m[10] = T.m[15] = 1.0f;
    T.m[12] = t[0], T.m[13] = t[1], T.m[14] = t[2];
    const float xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
    const float wx = qw * qx, wy = qw * qy, wz = qw * qz;
    R.m[15] = 1.0f;
    R.m[0] = 1 - 2 * (yy + zz), R.m[1] = 2 * (xy + wz), R.m[2] = 2 * (xz - wy);
    R.m[4] = 2 * (xy - wz), R.m[5] = 1 - 2 * (xx + zz), R.m[6] = 2 * (yz + wx);
    R.m[8] = 2 * (xz + wy), R.m[9] = 2 * (yz - wx), R.m[10] = 1 - 2 * (xx + yy);
    S.m[0] = t[7], S.m[5] = t[8], S.m[10] = t[9], S.m[15] = 1.0f;
    mul44(T, R, TR);
    mul44(TR, S, local);
    parent = parents[j];
    if (parent < 0 || uint32_t(parent) >= j) parent = -1; // the reference composes only with an earlier joint
    for (int p = parent; p >= 0;) {
      ++level;
      const int pp = parents[p];
      p = (pp >= 0 && pp < p) ? pp : -1;
    }
    atomicMax(&s_maxLevel, level);
    for (int i = 0; i < 16; ++i) s_global[j * 16 + i] = local.m[i];
  }
  __syncthreads();
  const int maxLevel = s_maxLevel;
  for (int l = 1; l <= maxLevel; ++l) {
    if (j < jointCount && level == l) {
      M4 pg, g;
      for (int i = 0; i < 16; ++i) pg.m[i] = s_global[size_t(parent) * 16 + i];
      mul44(pg, local, g);
      for (int i = 0; i < 16; ++i) s_global[j * 16 + i] = g.m[i];
    }
    __syncthreads();
  }
  if (j < jointCount) {
    M4 g, ib, skin;
    for (int i = 0; i < 16; ++i) g.m[i] = s_global[j * 16 + i], ib.m[i] = inverseBind[size_t(j) * 16 + i];
    mul44(g, ib, skin);
    for (int i = 0; i < 16; ++i) palette[size_t(j) * 16 + i] = skin.m[i];
  }
}

int launchJointPalette(rt_context *ctx, const float *trs, const int32_t *parents, const float *inverseBind,
                       uint32_t jointCount, float *palette) {
  RT_CHECK(trs && parents && inverseBind && palette, "rt_joint_palette: null pointer");
  RT_CHECK(jointCount >= 1 && jointCount <= 1024, "rt_joint_palette: jointCount must be 1..1024");
  const uint32_t threads = (jointCount + 31u) & ~31u;
  const size_t smem = size_t(jointCount) * 64;
  if (smem > 48 * 1024)
    RT_CUDA(cudaFuncSetAttribute(k_joint_palette, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  k_joint_palette<<<1, threads, smem, ctx->stream>>>(trs, parents, inverseBind, jointCount, palette);
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  return 0;
}

int launchSkin(rt_context *ctx, const void *const buffers[RT_BUFFER_COUNT], uint32_t vertexCount) {
  return launchSkinImpl(ctx, buffers, vertexCount, 0);
}

} // namespace rtb
