// traverse.cuh — software ray traversal for sm_100a (B200 has no RT cores).
//
// Stands in for Metal's intersector<triangle_data, instancing>::intersect as the reference calls it
// (MetalRaytracing/Raytracing.metal:301-318 closest hit, :664-665 and :730-737 any hit): two-level traversal of
// the 8-wide quantised BVH built by bvh_build.cu, closest hit with the deterministic tie rule, or any hit.
//
// Contract shared with the CPU oracle (oracle/oracle_bvh.cpp) so primary-hit ids agree bit for bit:
//   * the ray enters an instance through the float world->object matrix, direction not renormalised;
//   * ray/triangle = watertight test of Woop, Benthin, Wald 2013, same operation order, every float op rounded
//     on its own (this file must be compiled with -fmad=false; FMAs appear only where written as fmaf, i.e. in
//     the box tests, which only have to be conservative);
//   * hit iff tmin < t < tmax; closest = smallest t, ties -> smallest (instance, geometry, primitive).
// Traversal order follows Ylitie et al. 2017: children sit in slots matched to octants, a node's hit mask is
// permuted by the ray octant so the highest set bit is the nearest child, and triangle hits of one node form a
// 24-bit mask over a contiguous primitive range.
#pragma once
#include "common.cuh"

namespace rtb {

struct RayHit {
  float t, u, v;
  uint32_t instance, geometry, primitive;
};

constexpr int kStackSize = 48;
// RT_SLIM_LANE: the state of a ray that is touched once or twice in its life — the world-space ray (needed again only
// when an instance is entered or left), the (u, v) and ids of the best hit (written on an accepted hit, read at the
// end) and the path slot — lives in kStackExtra entries behind the traversal stack (local memory, L1) instead of 12
// registers per lane. The traversal kernel then fits 72 registers with hardly a spill: 7 CTAs per SM instead of 6
// (profiles/r2_experiments.md section 9).
#ifndef RT_SLIM_LANE
#define RT_SLIM_LANE 1
#endif
constexpr int kStackExtra = 9;
constexpr int kSlotHitUV = kStackSize, kSlotHitIds = kStackSize + 1, kSlotHitPrim = kStackSize + 2, kSlotRay0 = kStackSize + 3,
              kSlotRay1 = kStackSize + 4, kSlotRay2 = kStackSize + 5, kSlotPath = kStackSize + 6,
              kSlotCarry0 = kStackSize + 7, kSlotCarry1 = kStackSize + 8; // what the kernel carries along for the ray's end

// Traversal stack of one lane. LocalStack: a plain array (local memory, served by L1). SplitStack: the first
// kShared entries — all a ray normally needs — live in shared memory, laid out [entry][thread] so a warp's accesses
// are conflict-free; deeper entries spill to a local array. Pops then cost a shared-memory load instead of a local
// load that competes with the BVH nodes for L1 lines.
struct LocalStack {
  uint2 e[kStackSize + kStackExtra];
  __device__ __forceinline__ void set(int i, uint2 v) { e[i] = v; }
  __device__ __forceinline__ uint2 get(int i) const { return e[i]; }
};
template <int kShared, int kThreads>
struct SplitStack {
  uint2 *shared; // &smem[threadIdx.x]; entry i at shared[i * kThreads]
  uint2 spill[kStackSize - kShared + kStackExtra];
  __device__ __forceinline__ void set(int i, uint2 v) {
    if (i < kShared) shared[i * kThreads] = v;
    else spill[i - kShared] = v;
  }
  __device__ __forceinline__ uint2 get(int i) const { return i < kShared ? shared[i * kThreads] : spill[i - kShared]; }
};

__device__ __forceinline__ float pick3(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

// Component selection without branches or predicates: m0 / m1 are all-ones when the wanted axis is 0 / 1. Two LOP3
// per pick (ncu showed the `?:` form compiled to divergent branches, 10 % of the traversal kernel's instructions).
struct AxisMask {
  uint32_t m0, m1;
};
__device__ __forceinline__ AxisMask axisMask(int k) { return {k == 0 ? 0xFFFFFFFFu : 0u, k == 1 ? 0xFFFFFFFFu : 0u}; }
__device__ __forceinline__ float pickMasked(float x, float y, float z, const AxisMask &m) {
  const uint32_t yz = (__float_as_uint(y) & m.m1) | (__float_as_uint(z) & ~m.m1);
  return __uint_as_float((__float_as_uint(x) & m.m0) | (yz & ~m.m0));
}

struct TriSetup { // Woop et al. per-ray constants in the current space
  uint32_t axes;    // kx | ky << 2 | kz << 4: the permutation as three axis numbers (one register instead of six masks;
                    // a lane keeps this for as long as it is inside an instance, whoever tests a triangle expands it)
  float Sx, Sy, Sz;
  float ox, oy, oz; // origin permuted to (kx, ky, kz)
};
struct TriAxes { // the permutation of a TriSetup as select masks (pickMasked)
  AxisMask kx, ky, kz;
};
__device__ __forceinline__ TriAxes expandAxes(uint32_t axes) {
  return {axisMask(int(axes & 3u)), axisMask(int((axes >> 2) & 3u)), axisMask(int((axes >> 4) & 3u))};
}

__device__ __forceinline__ TriSetup makeTriSetup(float ox, float oy, float oz, float dx, float dy, float dz) {
  TriSetup s;
  float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
  const int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
  const int k1 = kz == 2 ? 0 : kz + 1;
  const int k2 = k1 == 2 ? 0 : k1 + 1;
  const AxisMask mz = axisMask(kz);
  const float dkz = pickMasked(dx, dy, dz, mz);
  const bool swap = dkz < 0.0f; // keeps the winding
  const int kx = swap ? k2 : k1, ky = swap ? k1 : k2;
  const AxisMask mx = axisMask(kx), my = axisMask(ky);
  s.axes = uint32_t(kx) | (uint32_t(ky) << 2) | (uint32_t(kz) << 4);
  s.Sx = pickMasked(dx, dy, dz, mx) / dkz;
  s.Sy = pickMasked(dx, dy, dz, my) / dkz;
  s.Sz = 1.0f / dkz;
  s.ox = pickMasked(ox, oy, oz, mx);
  s.oy = pickMasked(ox, oy, oz, my);
  s.oz = pickMasked(ox, oy, oz, mz);
  return s;
}

// Returns true when tmin < t < tmax; same arithmetic as oracle intersectTriangle().
__device__ __forceinline__ bool intersectTriangle(const TriSetup &s, const float4 &v0, const float4 &v1,
                                                  const float4 &v2, float tmin, float tmax, float &tOut, float &uOut,
                                                  float &vOut) {
  const TriAxes a = expandAxes(s.axes);
  const float Akx = pickMasked(v0.x, v0.y, v0.z, a.kx) - s.ox, Aky = pickMasked(v0.x, v0.y, v0.z, a.ky) - s.oy,
              Akz = pickMasked(v0.x, v0.y, v0.z, a.kz) - s.oz;
  const float Bkx = pickMasked(v1.x, v1.y, v1.z, a.kx) - s.ox, Bky = pickMasked(v1.x, v1.y, v1.z, a.ky) - s.oy,
              Bkz = pickMasked(v1.x, v1.y, v1.z, a.kz) - s.oz;
  const float Ckx = pickMasked(v2.x, v2.y, v2.z, a.kx) - s.ox, Cky = pickMasked(v2.x, v2.y, v2.z, a.ky) - s.oy,
              Ckz = pickMasked(v2.x, v2.y, v2.z, a.kz) - s.oz;
  const float Ax = Akx - s.Sx * Akz, Ay = Aky - s.Sy * Akz;
  const float Bx = Bkx - s.Sx * Bkz, By = Bky - s.Sy * Bkz;
  const float Cx = Ckx - s.Sx * Ckz, Cy = Cky - s.Sy * Ckz;
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    double CxBy = double(Cx) * double(By), CyBx = double(Cy) * double(Bx);
    U = float(CxBy - CyBx);
    double AxCy = double(Ax) * double(Cy), AyCx = double(Ay) * double(Cx);
    V = float(AxCy - AyCx);
    double BxAy = double(Bx) * double(Ay), ByAx = double(By) * double(Ax);
    W = float(BxAy - ByAx);
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  const float det = (U + V) + W;
  if (det == 0.0f) return false;
  const float Az = s.Sz * Akz, Bz = s.Sz * Bkz, Cz = s.Sz * Ckz;
  const float T = (U * Az + V * Bz) + W * Cz;
  const float invDet = 1.0f / det;
  const float t = T * invDet;
  if (!(t > tmin && t < tmax)) return false;
  tOut = t;
  uOut = V * invDet;
  vOut = W * invDet;
  return true;
}

struct BoxSetup { // per-ray constants for the quantised child-box tests in the current space
  float idx, idy, idz; // 1 / direction (zero components replaced by a tiny value of the same sign)
  float ox, oy, oz;
  uint32_t octinv;     // 7 ^ (sign bits of the direction): permutes slots into front-to-back priority
};
// The bits of 1.0f as a value the compiler cannot fold into an immediate (see byteAsUnitFloat): derived from a kernel
// parameter (a node count is always < 2^31), so it is uniform, costs no per-lane state and can be rematerialised.
__device__ __forceinline__ uint32_t unfoldableOne(uint32_t anyCountBelow2G) { return 0x3F800000u | (anyCountBelow2G >> 31); }

// 1 / d for the box tests only: they just have to be conservative, so the single-instruction reciprocal
// (MUFU.RCP, <= 1 ulp) is enough — its error is covered by the widening `eps` in intersectChildren.
__device__ __forceinline__ float safeInverse(float d) {
  const float tiny = 1.0e-20f;
  float a = fabsf(d) < tiny ? copysignf(tiny, d) : d;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}

__device__ __forceinline__ BoxSetup makeBoxSetup(float ox, float oy, float oz, float dx, float dy, float dz) {
  BoxSetup b;
  b.idx = safeInverse(dx);
  b.idy = safeInverse(dy);
  b.idz = safeInverse(dz);
  b.ox = ox, b.oy = oy, b.oz = oz;
  uint32_t neg = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u);
  b.octinv = 7u ^ neg;
  return b;
}

// Plain-conversion form of the child test: the specification the fast PRMT form below is checked against
// (rt_selftest_child_boxes) and a drop-in fallback when RT_LEGACY_CHILD_TEST is defined.
__device__ __forceinline__ float byteToFloat(uint32_t word, int byteIndex) {
  return float((word >> (8 * byteIndex)) & 0xFFu);
}

// Tests the eight children of one node; returns the hit mask: bits 24..31 internal children in traversal
// priority, bits 0..23 leaf primitives relative to primBase.
__device__ __forceinline__ uint32_t intersectChildrenLegacy(const uint4 &n0, const uint4 &n1, const uint4 &n2,
                                                      const uint4 &n3, const uint4 &n4, const BoxSetup &b, float tmin,
                                                      float tmax) {
  const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
              sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
  const float aix = sx * b.idx, aiy = sy * b.idy, aiz = sz * b.idz;
  const float aox = (__uint_as_float(n0.x) - b.ox) * b.idx, aoy = (__uint_as_float(n0.y) - b.oy) * b.idy,
              aoz = (__uint_as_float(n0.z) - b.oz) * b.idz;
  // conservative widening: |t| <= |ao| + 255 |ai| along each axis; a few ulp of that covers the rounding of the
  // two products and the fma below, so a box is never missed because of float error
  const float eps = 6.0e-7f;
  const float wx = eps * (fabsf(aox) + 255.0f * fabsf(aix)), wy = eps * (fabsf(aoy) + 255.0f * fabsf(aiy)),
              wz = eps * (fabsf(aoz) + 255.0f * fabsf(aiz));
  const float nox = aox - wx, noy = aoy - wy, noz = aoz - wz; // near-plane offsets
  const float fox = aox + wx, foy = aoy + wy, foz = aoz + wz; // far-plane offsets
  const bool negx = b.idx < 0.0f, negy = b.idy < 0.0f, negz = b.idz < 0.0f;
  // word pairs: lo_x = n2.xy, lo_y = n2.zw, lo_z = n3.xy, hi_x = n3.zw, hi_y = n4.xy, hi_z = n4.zw
  uint32_t hitmask = 0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
    const uint32_t hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
    const uint32_t nearx = negx ? hix : lox, farx = negx ? lox : hix;
    const uint32_t neary = negy ? hiy : loy, fary = negy ? loy : hiy;
    const uint32_t nearz = negz ? hiz : loz, farz = negz ? loz : hiz;
    const uint32_t meta4 = half ? n1.w : n1.z;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t m = (meta4 >> (8 * k)) & 0xFFu;
      const float tnx = fmaf(byteToFloat(nearx, k), aix, nox), tfx = fmaf(byteToFloat(farx, k), aix, fox);
      const float tny = fmaf(byteToFloat(neary, k), aiy, noy), tfy = fmaf(byteToFloat(fary, k), aiy, foy);
      const float tnz = fmaf(byteToFloat(nearz, k), aiz, noz), tfz = fmaf(byteToFloat(farz, k), aiz, foz);
      const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
      const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
      if (m != 0u && tn <= tf) {
        const bool inner = (m & 0x18u) == 0x18u;
        const uint32_t bitIndex = (m & 31u) ^ (inner ? b.octinv : 0u);
        hitmask |= (m >> 5) << bitIndex;
      }
    }
  }
  return hitmask;
}


// Packed single precision (sm_100a): two floats in a 64-bit register pair, one instruction for both. Each half is the
// IEEE round-to-nearest result of the scalar operation, so results are bit-identical to the unpacked code.
// RT_PACKED_FMA=1 uses it for the child-box test (24 FFMA2 instead of 48 FFMA per node). Measured and left off: K3 19.70
// against 19.03 ms, K4 70.0 against 66.9 — the register pairs cost 100 bytes of spills at 64 registers and the issue
// slots saved do not make up for it (profiles/r2_experiments.md section 11).
#ifndef RT_PACKED_FMA
#define RT_PACKED_FMA 0
#endif
__device__ __forceinline__ unsigned long long packF2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpackF2(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, float b, unsigned long long c) { // a * (b, b) + c
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(packF2(b, b)), "l"(c));
  return r;
}

// Quantised byte -> float without an integer conversion: PRMT drops the byte into mantissa bits 8..15 of 1.0f,
// giving 1 + q * 2^-15 exactly; the 2^15 is folded into the per-axis scale and the 1 into the offset.
// `one` must live in a register: PRMT takes a single immediate, and with a literal 1.0f the compiler spends it on
// that and moves the selector into a register before every PRMT (43 extra instructions per node in ncu's SASS view).
template <int k>
__device__ __forceinline__ float byteAsUnitFloat(uint32_t word, uint32_t one) {
  return __uint_as_float(__byte_perm(word, one, 0x7604u | (uint32_t(k) << 4)));
}

// Tests the eight children of one node; returns the hit mask: bits 24..31 internal children in traversal
// priority, bits 0..23 leaf primitives relative to primBase.
__device__ __forceinline__ uint32_t intersectChildren(const uint4 &n0, const uint4 &n1, const uint4 &n2,
                                                      const uint4 &n3, const uint4 &n4, const BoxSetup &b, float tmin,
                                                      float tmax, uint32_t one = 0x3F800000u) {
  // per-axis scale 2^(e-127) times 2^15 (the builder keeps e small enough for the sum not to overflow)
  const float sx = __uint_as_float(((n0.w & 0xFFu) + 15u) << 23), sy = __uint_as_float((((n0.w >> 8) & 0xFFu) + 15u) << 23),
              sz = __uint_as_float((((n0.w >> 16) & 0xFFu) + 15u) << 23);
  const float aix = sx * b.idx, aiy = sy * b.idy, aiz = sz * b.idz;
  const float aox = (__uint_as_float(n0.x) - b.ox) * b.idx, aoy = (__uint_as_float(n0.y) - b.oy) * b.idy,
              aoz = (__uint_as_float(n0.z) - b.oz) * b.idz;
  // plane distance t = (1 + q 2^-15) * ai + (ao - ai). Conservative widening: a few ulp of |ao| + |ai| covers
  // the rounding of both products, the difference and the fma, so no box is missed because of float error
  // (in units of one quantisation step this is < 0.02, far below the outward rounding of the boxes themselves).
  const float eps = 6.0e-7f; // 5 ulp: products, difference, fma and the approximate reciprocal of the direction
  const float wx = eps * (fabsf(aox) + fabsf(aix)), wy = eps * (fabsf(aoy) + fabsf(aiy)), wz = eps * (fabsf(aoz) + fabsf(aiz));
  const float cx = aox - aix, cy = aoy - aiy, cz = aoz - aiz;
  const float nox = cx - wx, noy = cy - wy, noz = cz - wz; // near-plane offsets
  const float fox = cx + wx, foy = cy + wy, foz = cz + wz; // far-plane offsets
  const bool negx = b.idx < 0.0f, negy = b.idy < 0.0f, negz = b.idz < 0.0f;
  const uint32_t octinv4 = b.octinv * 0x01010101u;
  // word pairs: lo_x = n2.xy, lo_y = n2.zw, lo_z = n3.xy, hi_x = n3.zw, hi_y = n4.xy, hi_z = n4.zw
  uint32_t hitmask = 0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
    const uint32_t hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
    const uint32_t nearx = negx ? hix : lox, farx = negx ? lox : hix;
    const uint32_t neary = negy ? hiy : loy, fary = negy ? loy : hiy;
    const uint32_t nearz = negz ? hiz : loz, farz = negz ? loz : hiz;
    const uint32_t meta4 = half ? n1.w : n1.z;
    // four children at once: internal children have both of bits 3,4 set in their low five bits (24..31)
    const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    // 0xFF in every internal child's byte: PTX prmt replicates a byte's sign bit when the selector nibble's msb is
    // set (the __byte_perm intrinsic only honours three selector bits, so this one is spelled in PTX)
    uint32_t innerMask4;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(innerMask4) : "r"(isInner4 << 3), "r"(0u), "r"(0xBA98u));
    const uint32_t bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1F1F1F1Fu;
    const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u; // empty slots contribute no bits
#if RT_PACKED_FMA
  // sm_100a's packed single-precision FMA (FFMA2: two independent round-to-nearest FMAs on a register pair, the scale
  // broadcast from one register): the near and the far plane of an axis in one instruction, 24 instead of 48 per node
  const unsigned long long offx = packF2(nox, fox), offy = packF2(noy, foy), offz = packF2(noz, foz);
#define RT_CHILD(k)                                                                                                  \
  {                                                                                                                  \
    float tnx, tfx, tny, tfy, tnz, tfz;                                                                              \
    unpackF2(fma2(packF2(byteAsUnitFloat<k>(nearx, one), byteAsUnitFloat<k>(farx, one)), aix, offx), tnx, tfx);      \
    unpackF2(fma2(packF2(byteAsUnitFloat<k>(neary, one), byteAsUnitFloat<k>(fary, one)), aiy, offy), tny, tfy);      \
    unpackF2(fma2(packF2(byteAsUnitFloat<k>(nearz, one), byteAsUnitFloat<k>(farz, one)), aiz, offz), tnz, tfz);      \
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));                                                       \
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));                                                       \
    if (tn <= tf) {                                                                                                  \
      const uint32_t bits = __byte_perm(childBits4, 0u, 0x4440u | uint32_t(k));                                      \
      const uint32_t index = __byte_perm(bitIndex4, 0u, 0x4440u | uint32_t(k));                                      \
      hitmask |= bits << index;                                                                                      \
    }                                                                                                                \
  }
#else
#define RT_CHILD(k)                                                                                                  \
  {                                                                                                                  \
    const float tnx = fmaf(byteAsUnitFloat<k>(nearx, one), aix, nox), tfx = fmaf(byteAsUnitFloat<k>(farx, one), aix, fox); \
    const float tny = fmaf(byteAsUnitFloat<k>(neary, one), aiy, noy), tfy = fmaf(byteAsUnitFloat<k>(fary, one), aiy, foy); \
    const float tnz = fmaf(byteAsUnitFloat<k>(nearz, one), aiz, noz), tfz = fmaf(byteAsUnitFloat<k>(farz, one), aiz, foz); \
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));                                                       \
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));                                                       \
    if (tn <= tf) {                                                                                                  \
      const uint32_t bits = __byte_perm(childBits4, 0u, 0x4440u | uint32_t(k));                                      \
      const uint32_t index = __byte_perm(bitIndex4, 0u, 0x4440u | uint32_t(k));                                      \
      hitmask |= bits << index;                                                                                      \
    }                                                                                                                \
  }
#endif
    RT_CHILD(0) RT_CHILD(1) RT_CHILD(2) RT_CHILD(3)
#undef RT_CHILD
  }
  return hitmask;
}

// Flat TLAS only (<= kFlatTlasMax instances): is the ray likely to reach an instance whose BLAS has nodes to traverse?
// Rays that do not — they miss everything, or reach only single-leaf instances such as the reference's planes — live
// for two or three iterations; the others walk a BVH. The wavefront kernels queue the two classes at opposite ends of a
// queue so that warps hold rays of one kind and the long rays start first. Only the ORDER of the work depends on the
// answer, never a result, so the test is the cheapest one that separates the classes well rather than a conservative
// one: the ray (unit direction) against the bounding sphere of the union box of the instances with nodes (`sphere` =
// centre xyz, radius; written once per dispatch by k_prepare_classes). About 25 instructions, no reciprocals — the
// slab test it replaced cost the shade kernel twice that for the same split.
__device__ __forceinline__ bool rayReachesNodes(const float4 *__restrict__ sphere, float ox, float oy, float oz, float dx,
                                                float dy, float dz, float tmax) {
  const float4 s = __ldg(sphere);
  const float px = s.x - ox, py = s.y - oy, pz = s.z - oz;
  const float along = px * dx + py * dy + pz * dz;       // distance along the ray to the point closest to the centre
  const float c = px * px + py * py + pz * pz - s.w * s.w; // <= 0: the origin is inside the sphere
  return c <= 0.0f || (along > 0.0f && along * along >= c && along - s.w <= tmax);
}

// Two-level traversal of one ray as an explicit state machine, so a kernel can either run it to completion
// (traverseScene) or interleave it with fetching new rays into idle lanes (trace_wavefront.cu).
// Each step() does exactly one of: one primitive (a triangle test, or entering an instance from a TLAS leaf), one
// node step (take the nearest pending child, test its eight children), or a pop. ncu showed this "one unit of
// work per iteration" shape keeps more lanes of a warp busy than draining all of a node's triangles in place.
// kAny: the traversal ends at the first accepted triangle (found == occluded).
template <bool kAny>
struct LaneTraversal {
#if !RT_SLIM_LANE
  float ox, oy, oz, dx, dy, dz; // world-space ray
#endif
  float tmin, tmax;
  // the TLAS arrays are not kept here: every step takes the TlasHeader that sits in the kernel's parameter space, so
  // those uniform pointers cost constant-bank operands instead of six registers per lane
  const uint4 *nodes;
  const float4 *tris;
  BoxSetup box;
  TriSetup tri;
  uint2 ngroup, tgroup;
  int sp, instanceSp; // instanceSp: stack depth at which the current instance was entered; -1 = world space
  uint32_t instance;
  RayHit hit; // RT_SLIM_LANE: only hit.t is kept here, the rest of the best hit is in the stack's extra entries
  bool found;
#ifdef RT_COUNT_WORK
  // counter build (tools/count_work.py): work done for this ray — node steps, triangle tests, instance entries.
  // bench.py turns them into counted bytes per ray (80 B per node, 48 B per triangle, 64 B per instance record)
  uint32_t nNodes, nTris, nEntries, nIters;
#define RT_COUNT(x) (++(x))
#else
#define RT_COUNT(x) ((void)0)
#endif
  // the traversal stack lives outside (a plain local array passed to every step) so that the compiler keeps the
  // scalar members above in registers instead of placing the whole object in local memory

  template <typename Stack>
  __device__ __forceinline__ void worldRay(const Stack &stack, float &ox_, float &oy_, float &oz_, float &dx_, float &dy_,
                                           float &dz_) const {
#if RT_SLIM_LANE
    const uint2 a = stack.get(kSlotRay0), b = stack.get(kSlotRay1), c = stack.get(kSlotRay2);
    ox_ = __uint_as_float(a.x), oy_ = __uint_as_float(a.y), oz_ = __uint_as_float(b.x);
    dx_ = __uint_as_float(b.y), dy_ = __uint_as_float(c.x), dz_ = __uint_as_float(c.y);
#else
    (void)stack;
    ox_ = ox, oy_ = oy, oz_ = oz, dx_ = dx, dy_ = dy, dz_ = dz;
#endif
  }

  // the result of a finished closest-hit traversal (hit.t == tmax and !found when nothing was hit)
  template <typename Stack>
  __device__ __forceinline__ RayHit result(const Stack &stack) const {
#if RT_SLIM_LANE
    RayHit h;
    h.t = hit.t;
    h.u = h.v = 0.0f;
    h.instance = h.geometry = h.primitive = 0u;
    if (found) {
      const uint2 uv = stack.get(kSlotHitUV), ids = stack.get(kSlotHitIds);
      h.u = __uint_as_float(uv.x), h.v = __uint_as_float(uv.y);
      h.instance = ids.x, h.geometry = ids.y;
      h.primitive = stack.get(kSlotHitPrim).x;
    }
    return h;
#else
    (void)stack;
    return hit;
#endif
  }

  // a triangle of the current instance was hit at t in (tmin, tmax): keep it if it is the best so far
  template <typename Stack>
  __device__ __forceinline__ void acceptHit(Stack &stack, float t, float u, float v, uint32_t geom, uint32_t prim) {
    bool better = t < hit.t;
    if (!better && found && t == hit.t) {
#if RT_SLIM_LANE
      const uint2 ids = stack.get(kSlotHitIds);
      const uint32_t bestPrim = stack.get(kSlotHitPrim).x;
      better = instance < ids.x || (instance == ids.x && (geom < ids.y || (geom == ids.y && prim < bestPrim)));
#else
      better = instance < hit.instance ||
               (instance == hit.instance && (geom < hit.geometry || (geom == hit.geometry && prim < hit.primitive)));
#endif
    }
    if (better) {
      found = true;
      hit.t = t;
#if RT_SLIM_LANE
      stack.set(kSlotHitUV, make_uint2(__float_as_uint(u), __float_as_uint(v)));
      stack.set(kSlotHitIds, make_uint2(instance, geom));
      stack.set(kSlotHitPrim, make_uint2(prim, 0u));
#else
      hit.u = u;
      hit.v = v;
      hit.instance = instance;
      hit.geometry = geom;
      hit.primitive = prim;
#endif
    }
  }

  template <bool kFlatAllowed = true, typename Stack>
  __device__ __forceinline__ void begin(const TlasHeader &tlas, Stack &stack, float ox_, float oy_, float oz_, float dx_,
                                        float dy_, float dz_, float tmin_, float tmax_) {
    const float ox = ox_, oy = oy_, oz = oz_, dx = dx_, dy = dy_, dz = dz_;
#if RT_SLIM_LANE
    stack.set(kSlotRay0, make_uint2(__float_as_uint(ox), __float_as_uint(oy)));
    stack.set(kSlotRay1, make_uint2(__float_as_uint(oz), __float_as_uint(dx)));
    stack.set(kSlotRay2, make_uint2(__float_as_uint(dy), __float_as_uint(dz)));
#else
    this->ox = ox_, this->oy = oy_, this->oz = oz_, this->dx = dx_, this->dy = dy_, this->dz = dz_;
#endif
    tmin = tmin_, tmax = tmax_;
    hit.t = tmax_;
    hit.u = hit.v = 0.0f;
    hit.instance = hit.geometry = hit.primitive = 0u;
    found = false;
#ifdef RT_COUNT_WORK
    nNodes = nTris = nEntries = nIters = 0u;
#endif
    sp = 0;
    instanceSp = -1;
    instance = 0;
    nodes = reinterpret_cast<const uint4 *>(tlas.nodes);
    tris = nullptr;
    // an empty TLAS has nothing pending: the first step() pops an empty stack and finishes
    const uint32_t nodeCount = tlas.nodeCount;
    box = makeBoxSetup(ox, oy, oz, dx, dy, dz);
    tri = TriSetup{};
    ngroup = make_uint2(0u, nodeCount != 0 ? 0x80000000u : 0u);
    tgroup = make_uint2(0u, 0u);
#ifndef RT_NO_FLAT_TLAS
    // Small scenes (<= kFlatTlasMax instances, a uniform property of the launch): the TLAS is one wide node whose leaf
    // slot k is instance k, and testing the ray against the k exact world boxes right here — a dozen instructions per
    // instance, executed by every lane that just received a ray — replaces that node's step (268 instructions at ~21
    // of 32 lanes, and a third to a half of all node steps of such scenes). The test is conservative (widened by a
    // few ulp of the magnitudes involved, as in intersectChildren), so exactly the instances the node test would have
    // reported, or fewer, are queued: results are unchanged.
    if (kFlatAllowed && tlas.instanceCount <= kFlatTlasMax && tlas.instanceBox != nullptr) {
      const float cx = ox * box.idx, cy = oy * box.idy, cz = oz * box.idz;
      uint32_t mask = 0u;
      for (uint32_t k = 0; k < tlas.instanceCount; ++k) {
        const float4 lo = __ldg(tlas.instanceBox + 2 * k), hi = __ldg(tlas.instanceBox + 2 * k + 1);
        const float ax = lo.x * box.idx, bx = hi.x * box.idx, ay = lo.y * box.idy, by = hi.y * box.idy,
                    az = lo.z * box.idz, bz = hi.z * box.idz;
        const float eps = 6.0e-7f;
        const float wx = eps * (fmaxf(fabsf(ax), fabsf(bx)) + fabsf(cx)), wy = eps * (fmaxf(fabsf(ay), fabsf(by)) + fabsf(cy)),
                    wz = eps * (fmaxf(fabsf(az), fabsf(bz)) + fabsf(cz));
        const float tn = fmaxf(fmaxf(fminf(ax - cx, bx - cx) - wx, fminf(ay - cy, by - cy) - wy),
                               fmaxf(fminf(az - cz, bz - cz) - wz, tmin_));
        const float tf = fminf(fminf(fmaxf(ax - cx, bx - cx) + wx, fmaxf(ay - cy, by - cy) + wy),
                               fminf(fmaxf(az - cz, bz - cz) + wz, tmax_));
        if (tn <= tf) mask |= 1u << k;
      }
      ngroup = make_uint2(0u, 0u);
      tgroup = make_uint2(0u, mask);
    }
#endif
  }

  // take the nearest pending child of ngroup, test its eight children
  template <typename Stack>
  __device__ __forceinline__ void nodeStep(const TlasHeader &tlas, Stack &stack) {
    RT_COUNT(nNodes);
    const uint32_t hits = ngroup.y;
    const uint32_t bit = 31u - uint32_t(__clz(int(hits)));
    ngroup.y &= ~(1u << bit);
    if (ngroup.y > 0x00FFFFFFu && sp < kStackSize) stack.set(sp++, ngroup);
    const uint32_t slot = (bit - 24u) ^ (box.octinv & 7u);
    const uint32_t rel = __popc(hits & 0xFFu & ~(0xFFFFFFFFu << slot));
    const uint4 *np = nodes + size_t(ngroup.x + rel) * 5;
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
#ifdef RT_LEGACY_CHILD_TEST
    const uint32_t hitmask = intersectChildrenLegacy(n0, n1, n2, n3, n4, box, tmin, hit.t);
#else
    const uint32_t hitmask = intersectChildren(n0, n1, n2, n3, n4, box, tmin, hit.t, unfoldableOne(tlas.nodeCount));
#endif
    ngroup = make_uint2(n1.x, (hitmask & 0xFF000000u) | (n0.w >> 24));
    tgroup = make_uint2(n1.y, hitmask & 0x00FFFFFFu);
  }

  // one primitive of tgroup. Returns true when an any-hit query is satisfied.
  template <typename Stack>
  __device__ __forceinline__ bool primitiveStep(const TlasHeader &tlas, Stack &stack) {
    if (instanceSp < 0) {
      enterInstance(tlas, stack);
      return false;
    }
    return triangleStep(stack);
  }

  // tgroup holds TLAS leaf entries (instanceSp < 0): enter the next instance
  template <typename Stack>
  __device__ __forceinline__ void enterInstance(const TlasHeader &tlas, Stack &stack) {
    RT_COUNT(nEntries);
    const uint32_t bit = uint32_t(__ffs(int(tgroup.y))) - 1u;
    tgroup.y &= ~(1u << bit);
    {
      // TLAS leaf: enter the instance. Pending world-space work goes to the stack first.
      if (tgroup.y != 0u && sp < kStackSize) stack.set(sp++, tgroup);
      if (ngroup.y > 0x00FFFFFFu && sp < kStackSize) stack.set(sp++, ngroup);
      instance = __ldg(tlas.leafInstance + tgroup.x + bit);
      const InstanceRecord *rec = tlas.instances + instance;
      const float4 r0 = __ldg(&rec->row0), r1 = __ldg(&rec->row1), r2 = __ldg(&rec->row2);
      const WideNode *bn = rec->nodes;
      tgroup.y = 0u;
      ngroup = make_uint2(0u, 0u);
      if (bn != nullptr) {
        float ox, oy, oz, dx, dy, dz;
        worldRay(stack, ox, oy, oz, dx, dy, dz);
        const float lox = ((r0.x * ox + r0.y * oy) + r0.z * oz) + r0.w;
        const float loy = ((r1.x * ox + r1.y * oy) + r1.z * oz) + r1.w;
        const float loz = ((r2.x * ox + r2.y * oy) + r2.z * oz) + r2.w;
        const float ldx = (r0.x * dx + r0.y * dy) + r0.z * dz;
        const float ldy = (r1.x * dx + r1.y * dy) + r1.z * dz;
        const float ldz = (r2.x * dx + r2.y * dy) + r2.z * dz;
        tri = makeTriSetup(lox, loy, loz, ldx, ldy, ldz);
        const uintptr_t tagged = reinterpret_cast<uintptr_t>(rec->tris);
#ifdef RT_NO_DIRECT_TRIS
        const uint32_t direct = 0u;
#else
        const uint32_t direct = uint32_t(tagged) & 31u; // single-leaf-node BLAS: its triangle count (bvh_build.cu)
#endif
        tris = reinterpret_cast<const float4 *>(tagged & ~uintptr_t(31));
        instanceSp = sp;
        if (direct != 0u) {
          // no node of this BLAS is ever tested: the world-space box setup (and `nodes`) stay as they are, and
          // popStep sees from `nodes` that there is nothing to restore when the instance is left
          tgroup = make_uint2(0u, 0xFFFFFFFFu >> (32u - direct));
        } else {
          box = makeBoxSetup(lox, loy, loz, ldx, ldy, ldz);
          nodes = reinterpret_cast<const uint4 *>(bn);
          ngroup = make_uint2(0u, 0x80000000u);
        }
      }
    }
  }

  // tgroup holds triangles of the current BLAS (instanceSp >= 0): test the next one.
  // Returns true when an any-hit query is satisfied.
  template <typename Stack>
  __device__ __forceinline__ bool triangleStep(Stack &stack) {
    RT_COUNT(nTris);
    const uint32_t bit = uint32_t(__ffs(int(tgroup.y))) - 1u;
    tgroup.y &= ~(1u << bit);
    const float4 *tp = tris + size_t(tgroup.x + bit) * 3;
    const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
    float t, u, v;
    if (intersectTriangle(tri, v0, v1, v2, tmin, kAny ? hit.t : tmax, t, u, v)) { // any-hit: hit.t stays tmax
      if (kAny) return true;
      acceptHit(stack, t, u, v, __float_as_uint(v1.w), __float_as_uint(v0.w));
    }
    return false;
  }

  // The triangle stage of stepConverged, done by the warp together (kCoop). Every lane calls it; `want` = this lane is
  // inside an instance with triangles pending. In a warp of incoherent rays about a quarter of the lanes have triangles
  // to test in any one iteration, and testing them where they sit ran the ~110 instructions of the watertight test at 7 - 9
  // of 32 threads, twice per iteration (profiles/r2_experiments.md section 9). Here the pending (lane, triangle) pairs of
  // the whole warp are numbered by a prefix sum over the lanes' triangle counts, the first 32 of them are listed in
  // shared memory, lane p tests pair p — it fetches the owner's Woop constants, triangle base and t range by shuffle —
  // and a ballot hands the outcome back: an any-hit owner only needs to know whether one of its pairs hit, a closest-hit
  // owner fetches the (t, u, v, ids) of each of its hits from the lane that found it. One pass retires up to 32 triangles
  // however they are spread over the lanes. The arithmetic of a test does not depend on the lane that runs it, and the
  // best hit under (t, instance, geometry, primitive) does not depend on the order the candidates arrive in, so results
  // are unchanged. Returns true when this lane's any-hit query has just been satisfied.
  template <typename Stack>
  __device__ __forceinline__ bool triangleStageCoop(Stack &stack, bool want, uint32_t *pairs) {
    const unsigned full = 0xFFFFFFFFu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = want ? uint32_t(__popc(tgroup.y)) : 0u;
    uint32_t incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(full, incl, d);
      if (lane >= uint32_t(d)) incl += up;
    }
    const uint32_t excl = incl - n;
    const uint32_t total = min(__shfl_sync(full, incl, 31), 32u);
    const uint32_t base = tgroup.x;
    uint32_t consumed = 0u;
    if (want && excl < 32u) { // list my triangles: pair number -> (owner lane, bit of the owner's leaf mask)
      uint32_t m = tgroup.y, p = excl;
      do {
        const uint32_t b = uint32_t(__ffs(int(m))) - 1u;
        m &= m - 1u;
        pairs[p++] = lane | (b << 8);
      } while (m != 0u && p < 32u);
      consumed = p - excl;
      tgroup.y = m; // what did not fit waits for the next pass
#ifdef RT_COUNT_WORK
      nTris += consumed;
#endif
    }
    __syncwarp();
    const bool testing = lane < total;
    uint32_t owner = lane, bit = 0u;
    if (testing) {
      const uint32_t pr = pairs[lane];
      owner = pr & 31u;
      bit = pr >> 8;
    }
    TriSetup s;
    s.axes = __shfl_sync(full, tri.axes, owner);
    s.Sx = __shfl_sync(full, tri.Sx, owner), s.Sy = __shfl_sync(full, tri.Sy, owner), s.Sz = __shfl_sync(full, tri.Sz, owner);
    s.ox = __shfl_sync(full, tri.ox, owner), s.oy = __shfl_sync(full, tri.oy, owner), s.oz = __shfl_sync(full, tri.oz, owner);
    const uintptr_t trisBits = reinterpret_cast<uintptr_t>(tris);
    const uint32_t tlo = __shfl_sync(full, uint32_t(trisBits), owner), thi = __shfl_sync(full, uint32_t(trisBits >> 32), owner);
    const uint32_t obase = __shfl_sync(full, base, owner);
    const float otmin = __shfl_sync(full, tmin, owner), otmax = __shfl_sync(full, tmax, owner);
    bool hitHere = false;
    float t = 0.0f, u = 0.0f, v = 0.0f;
    uint32_t prim = 0u, geom = 0u;
    if (testing) {
      const float4 *tp = reinterpret_cast<const float4 *>(uintptr_t(tlo) | (uintptr_t(thi) << 32)) + size_t(obase + bit) * 3;
      const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
      hitHere = intersectTriangle(s, v0, v1, v2, otmin, otmax, t, u, v);
      prim = __float_as_uint(v0.w), geom = __float_as_uint(v1.w);
    }
    const unsigned hits = __ballot_sync(full, hitHere);
    if (hits == 0u) return false;
    unsigned mine = consumed != 0u ? ((hits >> excl) & (0xFFFFFFFFu >> (32u - consumed))) : 0u;
    if (kAny) return mine != 0u;
    while (__ballot_sync(full, mine != 0u) != 0u) { // warp-uniform: one round per hit of the owner with the most hits
      const bool have = mine != 0u;
      const uint32_t src = have ? excl + uint32_t(__ffs(int(mine))) - 1u : lane;
      mine &= mine - 1u;
      const float ht = __shfl_sync(full, t, src), hu = __shfl_sync(full, u, src), hv = __shfl_sync(full, v, src);
      const uint32_t hp = __shfl_sync(full, prim, src), hg = __shfl_sync(full, geom, src);
      if (have) acceptHit(stack, ht, hu, hv, hg, hp);
    }
    return false;
  }

  // nothing pending in registers: leave the instance if its subtree is exhausted, then pop.
  // Returns false when the traversal is complete.
  template <typename Stack>
  __device__ __forceinline__ bool popStep(const TlasHeader &tlas, Stack &stack) {
    if (sp == instanceSp) {
      instanceSp = -1;
      if (nodes != reinterpret_cast<const uint4 *>(tlas.nodes)) { // a directly-tested BLAS never replaced them
        float ox, oy, oz, dx, dy, dz;
        worldRay(stack, ox, oy, oz, dx, dy, dz);
        box = makeBoxSetup(ox, oy, oz, dx, dy, dz);
        nodes = reinterpret_cast<const uint4 *>(tlas.nodes);
      }
    }
    if (sp == 0) return false;
    const uint2 e = stack.get(--sp);
    if (e.y > 0x00FFFFFFu) {
      ngroup = e;
    } else {
      tgroup = e;
      ngroup = make_uint2(0u, 0u);
    }
    return true;
  }

  // One fused iteration: a lane with nothing pending pops, a lane with a pending child (and no pending primitives)
  // tests that node, and a lane with pending primitives tests up to kPrims of them — each stage falls through
  // into the next, so a lane does up to three units of work per iteration while the warp still executes each stage
  // once. Primitives of a node are still all tested before any of its children is entered (hit.t shrinks first).
  // Returns false when the traversal has finished.
  template <int kPrims, typename Stack>
  __device__ __forceinline__ bool stepFused(const TlasHeader &tlas, Stack &stack) {
    if (tgroup.y == 0u && ngroup.y <= 0x00FFFFFFu) {
      if (!popStep(tlas, stack)) return false;
    }
    if (tgroup.y == 0u && ngroup.y > 0x00FFFFFFu) nodeStep(tlas, stack);
#pragma unroll
    for (int k = 0; k < kPrims; ++k) {
      if (tgroup.y != 0u) {
        if (primitiveStep(tlas, stack)) {
          found = true;
          return false;
        }
      }
    }
    return true;
  }

  // The iteration the wavefront kernel runs (trace_wavefront.cu, RT_CONVERGED): every lane of the warp calls it
  // (`active` = the lane holds a ray) and the stages pop -> [entry] -> node -> [entry] -> triangle x kTriangles are
  // separated by __syncwarp(). With the instance entry as a stage of its own a lane can test a node, enter the
  // instance a TLAS leaf names and test that instance's first triangles in one iteration; kEarlyFinish ends a ray at
  // the end of the iteration that emptied it instead of at the pop of the next one. The __syncwarp()s are what makes
  // this pay: written as plain consecutive `if`s (and with the extra exit) the compiler left the lanes that took an
  // earlier stage running apart from those that did not — ncu showed the same thread instructions in 1.3-1.5x the
  // warp instructions (profiles/r1_traversal.md, step 12). Per lane the order of node, entry and triangle tests is
  // the one stepFused has, so results are unchanged. Returns true when this lane's ray has just finished.
  template <bool kEntryBefore, bool kEntryAfter, bool kEarlyFinish, bool kCombined, int kTriangles, int kCoop, typename Stack>
  __device__ __forceinline__ bool stepConverged(const TlasHeader &tlas, Stack &stack, bool active, uint32_t *pairs) {
    bool alive = active;
#ifdef RT_COUNT_WORK
    if (active) ++nIters;
#endif
    // Which entry stage exists is chosen per kernel instantiation (trace_wavefront.cu): a flat TLAS (begin()) hands every
    // new ray its instances directly, so the entry comes first and the BLAS root is tested in the same iteration; with a
    // real TLAS the node step produces the instances, so the entry follows it.
    if (alive && tgroup.y == 0u && ngroup.y <= 0x00FFFFFFu) alive = popStep(tlas, stack);
    __syncwarp();
    if (kEntryBefore && !kCombined) {
      if (alive && tgroup.y != 0u && instanceSp < 0) enterInstance(tlas, stack);
      __syncwarp();
    }
    if (alive && tgroup.y == 0u && ngroup.y > 0x00FFFFFFu) nodeStep(tlas, stack);
    __syncwarp();
    if (kEntryAfter && !kCombined) {
      if (alive && tgroup.y != 0u && instanceSp < 0) enterInstance(tlas, stack);
      __syncwarp();
    }
    bool satisfied = false;
    if (kCombined) {
      if (alive && tgroup.y != 0u) satisfied = primitiveStep(tlas, stack);
      __syncwarp();
    } else if (kCoop) {
      // the warp tests its pending triangles together, 32 per pass (triangleStageCoop); a second pass only when more than
      // 32 were pending
#pragma unroll
      for (int k = 0; k < kTriangles; ++k) {
        const bool want = alive && !satisfied && tgroup.y != 0u && instanceSp >= 0;
        if (__ballot_sync(0xFFFFFFFFu, want) != 0u) {
          // kCoop == 2: when no lane has more than one triangle pending there is nothing to redistribute and every lane
          // tests its own in place (the cheaper stage); the cooperative pass runs when it replaces two or more of those
          if (kCoop == 2 && __ballot_sync(0xFFFFFFFFu, want && (tgroup.y & (tgroup.y - 1u)) != 0u) == 0u) {
            if (want) satisfied = triangleStep(stack);
          } else if (triangleStageCoop(stack, want, pairs)) {
            satisfied = true;
          }
        }
        __syncwarp();
      }
    } else {
#pragma unroll
      for (int k = 0; k < kTriangles; ++k) {
        if (alive && !satisfied && tgroup.y != 0u && instanceSp >= 0) satisfied = triangleStep(stack);
        __syncwarp();
      }
    }
    if (satisfied) {
      found = true;
      alive = false;
    }
    if (kEarlyFinish && sp == 0 && tgroup.y == 0u && ngroup.y <= 0x00FFFFFFu) alive = false;
    return active && !alive;
  }

  // One unit of work. Returns false when the traversal has finished (result in `hit` / `found`).
  template <typename Stack>
  __device__ __forceinline__ bool step(const TlasHeader &tlas, Stack &stack) {
    if (tgroup.y != 0u) {
      if (primitiveStep(tlas, stack)) {
        found = true;
        return false;
      }
      return true;
    }
    if (ngroup.y > 0x00FFFFFFu) {
      nodeStep(tlas, stack);
      return true;
    }
    return popStep(tlas, stack);
  }
};

// Runs one ray to completion. Closest hit: `hit` holds the result (hit.t == tmax and false when nothing was
// hit). kAny: true at the first accepted triangle.
template <bool kAny>
__device__ __forceinline__ bool traverseScene(const TlasHeader &tlas, float ox, float oy, float oz,
                                              float dx, float dy, float dz, float tmin, float tmax, RayHit &hit) {
  LaneTraversal<kAny> t;
  LocalStack stack;
  t.begin(tlas, stack, ox, oy, oz, dx, dy, dz, tmin, tmax);
  while (t.step(tlas, stack)) {
  }
  hit = t.result(stack);
  return t.found;
}

} // namespace rtb
