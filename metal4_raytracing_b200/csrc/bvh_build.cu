// bvh_build.cu — GPU acceleration-structure builder for sm_100a.
//
// Replaces MTLAccelerationStructure build / compact / refit, which the reference drives from
// MetalRaytracing/Utilities.swift:100-290 and MetalRaytracing/Renderer.swift:464-606,1084-1202 and whose
// implementation is a closed Apple driver. Pipeline (all on the device, one stream):
//   primitive boxes -> 63-bit Morton keys -> radix sort (CUB) -> LBVH hierarchy (Karras 2012) -> bottom-up
//   boxes + subtree sizes -> level-synchronous collapse into 80-byte 8-wide quantised nodes (common.cuh)
//   -> triangle / instance leaf records.
// A refit keeps the wide topology and recomputes triangle records, exact node boxes and quantised child boxes
// level by level from the leaves up; the BVH2 scratch is not needed for it.
// Compiled with -fmad=false: the instance-matrix inverse must round exactly like the host double arithmetic the
// oracle uses (DESIGN.md "numeric contract").
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace rtb {

namespace {

constexpr uint32_t kLeafBit = 0x80000000u;
constexpr int kMaxLeafPrims = 3;

struct GeomEntry {
  const uint8_t *vertices;
  const uint8_t *indices;
  uint32_t vertexStride;
  uint32_t indexStride;
  uint32_t firstTriangle; // prefix offset into the BLAS-wide triangle numbering
  uint32_t triangleCount;
};

// ---- ordered-uint encoding so float min/max can use integer atomics --------------------------------------
__device__ __forceinline__ uint32_t orderedFromFloat(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float floatFromOrdered(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

struct BoundsAtomics {
  uint32_t lo[3], hi[3];
};

__global__ void k_init_bounds(BoundsAtomics *b) {
  if (threadIdx.x == 0) {
    for (int a = 0; a < 3; ++a) {
      b->lo[a] = 0xFFFFFFFFu;
      b->hi[a] = 0u;
    }
  }
}

__device__ __forceinline__ void reduceBounds(BoundsAtomics *b, float3 lo, float3 hi, bool valid) {
  // warp reduce, then one atomic per warp
  const unsigned full = 0xFFFFFFFFu;
  if (!valid) {
    lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX);
    hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo.x = fminf(lo.x, __shfl_xor_sync(full, lo.x, o));
    lo.y = fminf(lo.y, __shfl_xor_sync(full, lo.y, o));
    lo.z = fminf(lo.z, __shfl_xor_sync(full, lo.z, o));
    hi.x = fmaxf(hi.x, __shfl_xor_sync(full, hi.x, o));
    hi.y = fmaxf(hi.y, __shfl_xor_sync(full, hi.y, o));
    hi.z = fmaxf(hi.z, __shfl_xor_sync(full, hi.z, o));
  }
  if ((threadIdx.x & 31) == 0 && lo.x <= hi.x) {
    atomicMin(&b->lo[0], orderedFromFloat(lo.x));
    atomicMin(&b->lo[1], orderedFromFloat(lo.y));
    atomicMin(&b->lo[2], orderedFromFloat(lo.z));
    atomicMax(&b->hi[0], orderedFromFloat(hi.x));
    atomicMax(&b->hi[1], orderedFromFloat(hi.y));
    atomicMax(&b->hi[2], orderedFromFloat(hi.z));
  }
}

__device__ __forceinline__ void loadTriangle(const GeomEntry *geoms, uint32_t geomCount, uint32_t tri, float3 &a,
                                             float3 &b, float3 &c, uint32_t &geom, uint32_t &prim) {
  uint32_t g = 0;
  while (g + 1 < geomCount && tri >= geoms[g + 1].firstTriangle) ++g;
  const GeomEntry ge = geoms[g];
  uint32_t p = tri - ge.firstTriangle;
  uint32_t i0, i1, i2;
  if (ge.indexStride == 2) {
    const uint16_t *ix = reinterpret_cast<const uint16_t *>(ge.indices);
    i0 = ix[3 * p], i1 = ix[3 * p + 1], i2 = ix[3 * p + 2];
  } else {
    const uint32_t *ix = reinterpret_cast<const uint32_t *>(ge.indices);
    i0 = ix[3 * p], i1 = ix[3 * p + 1], i2 = ix[3 * p + 2];
  }
  const float *pa = reinterpret_cast<const float *>(ge.vertices + size_t(i0) * ge.vertexStride);
  const float *pb = reinterpret_cast<const float *>(ge.vertices + size_t(i1) * ge.vertexStride);
  const float *pc = reinterpret_cast<const float *>(ge.vertices + size_t(i2) * ge.vertexStride);
  a = make_float3(pa[0], pa[1], pa[2]);
  b = make_float3(pb[0], pb[1], pb[2]);
  c = make_float3(pc[0], pc[1], pc[2]);
  geom = g;
  prim = p;
}

// One thread per triangle: box + centroid bounds.
__global__ void k_triangle_bounds(const GeomEntry *geoms, uint32_t geomCount, uint32_t n, float4 *primLo,
                                  float4 *primHi, BoundsAtomics *bounds) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = i < n;
  float3 lo = make_float3(0, 0, 0), hi = lo;
  if (valid) {
    float3 a, b, c;
    uint32_t g, p;
    loadTriangle(geoms, geomCount, i, a, b, c, g, p);
    lo = make_float3(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z));
    hi = make_float3(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z));
    // a triangle with a NaN or infinite coordinate can never be hit (every comparison of the watertight test fails),
    // but its box must not poison the scene bounds, the Morton keys or PLOC's distances: it becomes a point at the
    // origin and stays out of the global bounds
    const float big = 1.0e30f;
    const bool finite = fabsf(a.x) < big && fabsf(a.y) < big && fabsf(a.z) < big && fabsf(b.x) < big && fabsf(b.y) < big &&
                        fabsf(b.z) < big && fabsf(c.x) < big && fabsf(c.y) < big && fabsf(c.z) < big;
    if (!finite) {
      lo = hi = make_float3(0.0f, 0.0f, 0.0f);
      valid = false;
    }
    primLo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
    primHi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
  }
  reduceBounds(bounds, lo, hi, valid);
}

// One thread per instance: world box of the BLAS bounds, world->object matrix (double cofactor inverse, the
// exact op order of oracle/oracle_bvh.cpp invertAffine4x3), traversal record.
__global__ void k_instance_bounds(const rt_instance_descriptor *desc, uint32_t n, InstanceRecord *records,
                                  float4 *primLo, float4 *primHi, float4 *instanceBox, BoundsAtomics *bounds) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = i < n;
  float3 lo = make_float3(0, 0, 0), hi = lo;
  if (valid) {
    const rt_instance_descriptor d = desc[i];
    const BlasHeader *blas = reinterpret_cast<const BlasHeader *>(static_cast<uintptr_t>(d.accelerationStructureID));
    const float(*m)[3] = d.transformationMatrix;
    double a00 = m[0][0], a10 = m[0][1], a20 = m[0][2];
    double a01 = m[1][0], a11 = m[1][1], a21 = m[1][2];
    double a02 = m[2][0], a12 = m[2][1], a22 = m[2][2];
    double t0 = m[3][0], t1 = m[3][1], t2 = m[3][2];
    double c00 = a11 * a22 - a12 * a21;
    double c01 = a12 * a20 - a10 * a22;
    double c02 = a10 * a21 - a11 * a20;
    double det = (a00 * c00 + a01 * c01) + a02 * c02;
    double id = 1.0 / det;
    double i00 = c00 * id, i01 = (a02 * a21 - a01 * a22) * id, i02 = (a01 * a12 - a02 * a11) * id;
    double i10 = c01 * id, i11 = (a00 * a22 - a02 * a20) * id, i12 = (a02 * a10 - a00 * a12) * id;
    double i20 = c02 * id, i21 = (a01 * a20 - a00 * a21) * id, i22 = (a00 * a11 - a01 * a10) * id;
    double it0 = -((i00 * t0 + i01 * t1) + i02 * t2);
    double it1 = -((i10 * t0 + i11 * t1) + i12 * t2);
    double it2 = -((i20 * t0 + i21 * t1) + i22 * t2);
    InstanceRecord r;
    r.row0 = make_float4(float(i00), float(i01), float(i02), float(it0));
    r.row1 = make_float4(float(i10), float(i11), float(i12), float(it1));
    r.row2 = make_float4(float(i20), float(i21), float(i22), float(it2));
    bool empty = blas == nullptr || blas->triCount == 0;
    r.nodes = empty ? nullptr : blas->nodes;
    r.tris = empty ? nullptr : blas->tris;
    // a BLAS that is a single leaf-only node (a quad, a small prop): the traversal tests its <= 24 triangles
    // directly after entering the instance instead of testing the node's child boxes first; the count travels in
    // the low bits of the (256-byte aligned) triangle pointer
    if (!empty && blas->nodeCount == 1 && blas->triCount <= 24)
      r.tris = reinterpret_cast<const TriRecord *>(reinterpret_cast<uintptr_t>(blas->tris) | uintptr_t(blas->triCount));
    records[i] = r;
    if (!empty) {
      lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX);
      hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
      for (int corner = 0; corner < 8; ++corner) {
        float px = (corner & 1) ? blas->boundsHi[0] : blas->boundsLo[0];
        float py = (corner & 2) ? blas->boundsHi[1] : blas->boundsLo[1];
        float pz = (corner & 4) ? blas->boundsHi[2] : blas->boundsLo[2];
        float wx = ((m[0][0] * px + m[1][0] * py) + m[2][0] * pz) + m[3][0];
        float wy = ((m[0][1] * px + m[1][1] * py) + m[2][1] * pz) + m[3][1];
        float wz = ((m[0][2] * px + m[1][2] * py) + m[2][2] * pz) + m[3][2];
        lo = make_float3(fminf(lo.x, wx), fminf(lo.y, wy), fminf(lo.z, wz));
        hi = make_float3(fmaxf(hi.x, wx), fmaxf(hi.y, wy), fmaxf(hi.z, wz));
      }
      // pad for the world->object->world round trip (same rule as the oracle's TLAS)
      float px = 1.0e-5f * fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), 1.0e-3f);
      float py = 1.0e-5f * fmaxf(fmaxf(fabsf(lo.y), fabsf(hi.y)), 1.0e-3f);
      float pz = 1.0e-5f * fmaxf(fmaxf(fabsf(lo.z), fabsf(hi.z)), 1.0e-3f);
      lo = make_float3(lo.x - px, lo.y - py, lo.z - pz);
      hi = make_float3(hi.x + px, hi.y + py, hi.z + pz);
    } else {
      valid = false; // keeps the global bounds clean; the box below is a point at the origin of the instance
      lo = hi = make_float3(float(t0), float(t1), float(t2));
    }
    primLo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
    primHi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    // kept with the TLAS for the flat traversal of small scenes (an empty instance is a point; entering it is a no-op);
    // lo.w = 1 marks an instance without nodes to traverse (empty, or a single-leaf BLAS tested directly): a ray whose box
    // test reaches only such instances is a cheap ray (trace_wavefront.cu queues those separately)
    const bool nodeless = empty || (blas->nodeCount == 1 && blas->triCount <= 24);
    instanceBox[2 * i] = make_float4(lo.x, lo.y, lo.z, nodeless ? 1.0f : 0.0f);
    instanceBox[2 * i + 1] = primHi[i];
  }
  reduceBounds(bounds, lo, hi, valid);
}

__device__ __forceinline__ uint64_t expandBits21(uint32_t v) {
  uint64_t x = v & 0x1FFFFFu;
  x = (x | x << 32) & 0x1F00000000FFFFull;
  x = (x | x << 16) & 0x1F0000FF0000FFull;
  x = (x | x << 8) & 0x100F00F00F00F00Full;
  x = (x | x << 4) & 0x10C30C30C30C30C3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__global__ void k_morton(const float4 *primLo, const float4 *primHi, uint32_t n, const BoundsAtomics *bounds,
                         uint64_t *keys, uint32_t *vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float lo[3], ext[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = floatFromOrdered(bounds->lo[a]);
    float h = floatFromOrdered(bounds->hi[a]);
    ext[a] = h > lo[a] ? h - lo[a] : 1.0f;
  }
  float4 l = primLo[i], h = primHi[i];
  float c[3] = {0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z)};
  uint32_t q[3];
  for (int a = 0; a < 3; ++a) {
    float f = (c[a] - lo[a]) / ext[a];
    f = fminf(fmaxf(f, 0.0f), 1.0f);
    q[a] = min(uint32_t(f * 2097152.0f), 2097151u);
  }
  keys[i] = (expandBits21(q[0]) << 2) | (expandBits21(q[1]) << 1) | expandBits21(q[2]);
  vals[i] = i;
}

// ---- LBVH hierarchy (Karras 2012) ---------------------------------------------------------------------------
__device__ __forceinline__ int deltaKey(const uint64_t *keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(uint32_t(i) ^ uint32_t(j));
  return __clzll(static_cast<long long>(a ^ b));
}

struct Bvh2 {
  uint32_t n;
  uint32_t *left, *right; // [n-1]
  uint32_t *parent;       // [2n-1]: internal i -> i, leaf j -> (n-1) + j
  float4 *lo, *hi;        // [n-1]
  uint32_t *count;        // [n-1]
  uint32_t *flag;         // [n-1]
};

__global__ void k_karras(const uint64_t *keys, Bvh2 t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = int(t.n);
  if (i >= n - 1) return;
  int d = (deltaKey(keys, n, i, i + 1) - deltaKey(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = deltaKey(keys, n, i, i - d);
  int lmax = 2;
  while (deltaKey(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int s = lmax >> 1; s >= 1; s >>= 1)
    if (deltaKey(keys, n, i, i + (l + s) * d) > dmin) l += s;
  int j = i + l * d;
  int dnode = deltaKey(keys, n, i, j);
  int s = 0, step = l;
  do {
    step = (step + 1) >> 1;
    if (deltaKey(keys, n, i, i + (s + step) * d) > dnode) s += step;
  } while (step > 1);
  int gamma = i + s * d + min(d, 0);
  uint32_t L = (min(i, j) == gamma) ? (uint32_t(gamma) | kLeafBit) : uint32_t(gamma);
  uint32_t R = (max(i, j) == gamma + 1) ? (uint32_t(gamma + 1) | kLeafBit) : uint32_t(gamma + 1);
  t.left[i] = L;
  t.right[i] = R;
  t.parent[(L & kLeafBit) ? (n - 1) + int(L & ~kLeafBit) : int(L)] = uint32_t(i);
  t.parent[(R & kLeafBit) ? (n - 1) + int(R & ~kLeafBit) : int(R)] = uint32_t(i);
  if (i == 0) t.parent[0] = 0xFFFFFFFFu;
  t.flag[i] = 0;
}

// One thread per leaf walks up; the second visitor of a node merges its children.
__global__ void k_bvh2_bounds(const float4 *primLo, const float4 *primHi, const uint32_t *sorted, Bvh2 t) {
  uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= t.n || t.n < 2) return;
  uint32_t node = t.parent[(t.n - 1) + leaf];
  while (node != 0xFFFFFFFFu) {
    if (atomicAdd(&t.flag[node], 1u) == 0u) return; // first visitor leaves
    __threadfence();
    uint32_t L = t.left[node], R = t.right[node];
    float4 llo, lhi, rlo, rhi;
    uint32_t lc, rc;
    if (L & kLeafBit) {
      uint32_t p = sorted[L & ~kLeafBit];
      llo = primLo[p], lhi = primHi[p], lc = 1;
    } else {
      llo = t.lo[L], lhi = t.hi[L], lc = t.count[L];
    }
    if (R & kLeafBit) {
      uint32_t p = sorted[R & ~kLeafBit];
      rlo = primLo[p], rhi = primHi[p], rc = 1;
    } else {
      rlo = t.lo[R], rhi = t.hi[R], rc = t.count[R];
    }
    t.lo[node] = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
    t.hi[node] = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
    t.count[node] = lc + rc;
    __threadfence();
    node = t.parent[node];
  }
}

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner 2018) --------------------------------------
// Builds the binary hierarchy bottom-up over the Morton-sorted primitives: every cluster looks `radius` places to
// either side for the neighbour whose union with it has the smallest surface area; mutual nearest neighbours merge;
// the survivors are compacted in order and the round repeats until one cluster is left. The tree has SAH quality
// close to a top-down binned builder at a fraction of the cost, and replaces the Karras hierarchy for everything
// that is not rebuilt every frame. Node ids come from a prefix sum, so the tree is deterministic.
struct PlocClusters {
  uint32_t *ref;  // leaf (sorted index | kLeafBit) or internal node id
  float4 *lo, *hi;
};

__global__ void k_ploc_init(const float4 *primLo, const float4 *primHi, const uint32_t *sorted, uint32_t n,
                            PlocClusters c) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t p = sorted[i];
  c.ref[i] = i | kLeafBit;
  c.lo[i] = primLo[p];
  c.hi[i] = primHi[p];
}

__global__ void k_ploc_nearest(PlocClusters c, uint32_t count, int radius, uint32_t *nearest) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float4 lo = c.lo[i], hi = c.hi[i];
  const int first = max(0, int(i) - radius), last = min(int(count) - 1, int(i) + radius);
  float best = FLT_MAX;
  uint32_t bestJ = i;
  for (int j = first; j <= last; ++j) {
    if (j == int(i)) continue;
    const float4 l = c.lo[j], h = c.hi[j];
    const float dx = fmaxf(hi.x, h.x) - fminf(lo.x, l.x), dy = fmaxf(hi.y, h.y) - fminf(lo.y, l.y),
                dz = fmaxf(hi.z, h.z) - fminf(lo.z, l.z);
    const float area = dx * dy + dy * dz + dz * dx;
    if (area < best) { // ties keep the smaller index: the relation stays symmetric enough for mutual pairs to exist
      best = area;
      bestJ = uint32_t(j);
    }
  }
  nearest[i] = bestJ;
}

// flags[i]: low word 1 when cluster i survives the round (alone or as the merged pair), high word 1 when it is the
// lower index of a mutual pair (it creates a node)
__global__ void k_ploc_flags(const uint32_t *nearest, uint32_t count, unsigned long long *flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t j = nearest[i];
  const bool mutual = j != i && nearest[j] == i;
  const bool merges = mutual && i < j, dies = mutual && i > j;
  flags[i] = (dies ? 0ull : 1ull) | (merges ? (1ull << 32) : 0ull);
}

__global__ void k_ploc_merge(PlocClusters in, PlocClusters out, const uint32_t *nearest, const unsigned long long *scan,
                             uint32_t count, uint32_t nodeBase, Bvh2 t) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t j = nearest[i];
  const bool mutual = j != i && nearest[j] == i;
  if (mutual && i > j) return;
  // inclusive scan: position = (survivors up to and including i) - 1
  const unsigned long long inc = scan[i];
  const uint32_t pos = uint32_t(inc & 0xFFFFFFFFull) - 1u;
  float4 lo = in.lo[i], hi = in.hi[i];
  uint32_t ref = in.ref[i];
  if (mutual) {
    const uint32_t node = nodeBase + uint32_t(inc >> 32) - 1u;
    const float4 l = in.lo[j], h = in.hi[j];
    const uint32_t rj = in.ref[j];
    lo = make_float4(fminf(lo.x, l.x), fminf(lo.y, l.y), fminf(lo.z, l.z), 0.0f);
    hi = make_float4(fmaxf(hi.x, h.x), fmaxf(hi.y, h.y), fmaxf(hi.z, h.z), 0.0f);
    t.left[node] = ref;
    t.right[node] = rj;
    t.lo[node] = lo;
    t.hi[node] = hi;
    t.count[node] = ((ref & kLeafBit) ? 1u : t.count[ref]) + ((rj & kLeafBit) ? 1u : t.count[rj]);
    ref = node;
  }
  out.ref[pos] = ref;
  out.lo[pos] = lo;
  out.hi[pos] = hi;
}

// ---- quantisation of one wide node ---------------------------------------------------------------------------
struct ChildBox {
  float lo[3], hi[3];
};

// Chooses origin/exponents for `box` and quantises the child boxes conservatively (checked in double, where
// origin + q * 2^e is exact). Writes w0 (keeping imask), w2..w4.
__device__ void quantiseNode(WideNode &node, const float lo[3], const float hi[3], const ChildBox child[8],
                             const uint8_t present[8], uint8_t imask) {
  uint32_t ebyte[3];
  float scale[3];
  for (int a = 0; a < 3; ++a) {
    float ext = hi[a] - lo[a];
    int e;
    if (!(ext > 0.0f)) {
      e = -126;
    } else {
      // smallest e with ext / 2^e <= 255
      float m = frexpf(ext / 255.0f, &e); // ext/255 = m * 2^e, m in [0.5, 1)
      if (m == 0.5f) e -= 1;
      e = max(e, -126);
      while (double(ext) / ldexp(1.0, e) > 255.0) ++e;
    }
    e = min(e, 100); // traversal adds 15 to the exponent byte (traverse.cuh); extents beyond 2^108 are not scenes
    ebyte[a] = uint32_t(e + 127);
    scale[a] = __uint_as_float(ebyte[a] << 23);
  }
  uint32_t q[6][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}}; // [lo xyz, hi xyz][word]
  for (int s = 0; s < 8; ++s) {
    if (!present[s]) continue;
    for (int a = 0; a < 3; ++a) {
      double sc = double(scale[a]), org = double(lo[a]);
      int ql = int(floor((double(child[s].lo[a]) - org) / sc));
      int qh = int(ceil((double(child[s].hi[a]) - org) / sc));
      ql = max(0, min(255, ql));
      qh = max(0, min(255, qh));
      while (ql > 0 && org + double(ql) * sc > double(child[s].lo[a])) --ql;
      while (qh < 255 && org + double(qh) * sc < double(child[s].hi[a])) ++qh;
      if (qh < ql) qh = ql;
      q[a][s >> 2] |= uint32_t(ql) << (8 * (s & 3));
      q[3 + a][s >> 2] |= uint32_t(qh) << (8 * (s & 3));
    }
  }
  node.w[0] = make_uint4(__float_as_uint(lo[0]), __float_as_uint(lo[1]), __float_as_uint(lo[2]),
                         ebyte[0] | (ebyte[1] << 8) | (ebyte[2] << 16) | (uint32_t(imask) << 24));
  node.w[2] = make_uint4(q[0][0], q[0][1], q[1][0], q[1][1]);
  node.w[3] = make_uint4(q[2][0], q[2][1], q[3][0], q[3][1]);
  node.w[4] = make_uint4(q[4][0], q[4][1], q[5][0], q[5][1]);
}

__device__ __forceinline__ float halfArea(const float lo[3], const float hi[3]) {
  float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  return dx * dy + dy * dz + dz * dx;
}

struct CollapseCounters {
  uint32_t nodeCount; // next free wide-node index
  uint32_t primCount; // next free leaf-primitive slot
};

// Wide node levelStart + i of the current level. queueIn[i] = BVH2 reference of that node.
__device__ void collapseOne(uint32_t i, const Bvh2 &t, const float4 *primLo, const float4 *primHi, const uint32_t *sorted,
                            const uint32_t *queueIn, uint32_t *queueOut, uint32_t levelStart, uint32_t nextLevelStart,
                            CollapseCounters *counters, WideNode *nodes, float4 *nodeBox, uint32_t *leafPrim,
                            uint32_t maxLeafPrims) {
  const uint32_t self = queueIn[i];
  auto boxOf = [&](uint32_t ref, float lo[3], float hi[3]) {
    float4 l, h;
    if (ref & kLeafBit) {
      uint32_t p = sorted[ref & ~kLeafBit];
      l = primLo[p], h = primHi[p];
    } else {
      l = t.lo[ref], h = t.hi[ref];
    }
    lo[0] = l.x, lo[1] = l.y, lo[2] = l.z, hi[0] = h.x, hi[1] = h.y, hi[2] = h.z;
  };
  auto countOf = [&](uint32_t ref) { return (ref & kLeafBit) ? 1u : t.count[ref]; };

  uint32_t list[8];
  float area[8];
  int n = 0;
  if (self & kLeafBit) {
    list[n++] = self;
  } else {
    list[n++] = t.left[self];
    list[n++] = t.right[self];
  }
  for (int k = 0; k < n; ++k) {
    float lo[3], hi[3];
    boxOf(list[k], lo, hi);
    area[k] = halfArea(lo, hi);
  }
  // open the largest multi-primitive entry until eight children exist
  while (n < 8) {
    int best = -1;
    float bestArea = -1.0f;
    for (int k = 0; k < n; ++k)
      if (!(list[k] & kLeafBit) && area[k] > bestArea) {
        bestArea = area[k];
        best = k;
      }
    if (best < 0) break;
    uint32_t ref = list[best];
    uint32_t L = t.left[ref], R = t.right[ref];
    float lo[3], hi[3];
    list[best] = L;
    boxOf(L, lo, hi);
    area[best] = halfArea(lo, hi);
    list[n] = R;
    boxOf(R, lo, hi);
    area[n] = halfArea(lo, hi);
    ++n;
  }
  // node box = union of children (exact floats)
  float nlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, nhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  ChildBox cb[8];
  for (int k = 0; k < n; ++k) {
    boxOf(list[k], cb[k].lo, cb[k].hi);
    for (int a = 0; a < 3; ++a) {
      nlo[a] = fminf(nlo[a], cb[k].lo[a]);
      nhi[a] = fmaxf(nhi[a], cb[k].hi[a]);
    }
  }
  // slot assignment: child k goes to the free slot whose octant diagonal best matches its offset from the centre
  float centre[3] = {0.5f * (nlo[0] + nhi[0]), 0.5f * (nlo[1] + nhi[1]), 0.5f * (nlo[2] + nhi[2])};
  int slotOf[8];
  bool childDone[8] = {false, false, false, false, false, false, false, false};
  bool slotUsed[8] = {false, false, false, false, false, false, false, false};
  for (int round = 0; round < n; ++round) {
    float bestScore = -FLT_MAX;
    int bk = 0, bs = 0;
    for (int k = 0; k < n; ++k) {
      if (childDone[k]) continue;
      float dx = 0.5f * (cb[k].lo[0] + cb[k].hi[0]) - centre[0];
      float dy = 0.5f * (cb[k].lo[1] + cb[k].hi[1]) - centre[1];
      float dz = 0.5f * (cb[k].lo[2] + cb[k].hi[2]) - centre[2];
      for (int s = 0; s < 8; ++s) {
        if (slotUsed[s]) continue;
        float score = ((s & 1) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 4) ? dz : -dz);
        if (score > bestScore) {
          bestScore = score;
          bk = k;
          bs = s;
        }
      }
    }
    childDone[bk] = true;
    slotUsed[bs] = true;
    slotOf[bk] = bs;
  }
  // classify + allocate
  uint32_t refAt[8];
  uint8_t present[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  ChildBox slotBox[8];
  for (int k = 0; k < n; ++k) {
    refAt[slotOf[k]] = list[k];
    present[slotOf[k]] = 1;
    slotBox[slotOf[k]] = cb[k];
  }
  uint32_t internalCount = 0, primTotal = 0;
  uint8_t imask = 0;
  for (int s = 0; s < 8; ++s) {
    if (!present[s]) continue;
    uint32_t c = countOf(refAt[s]);
    if (c > maxLeafPrims) {
      imask |= uint8_t(1u << s);
      ++internalCount;
    } else {
      primTotal += c;
    }
  }
  uint32_t childBase = internalCount ? atomicAdd(&counters->nodeCount, internalCount) : 0u;
  uint32_t primBase = primTotal ? atomicAdd(&counters->primCount, primTotal) : 0u;
  uint32_t meta[2] = {0, 0};
  uint32_t nextChild = childBase, primOffset = 0;
  for (int s = 0; s < 8; ++s) {
    if (!present[s]) continue;
    uint32_t ref = refAt[s];
    uint32_t m;
    if (imask & (1u << s)) {
      m = 0x20u | (24u + uint32_t(s));
      queueOut[nextChild - nextLevelStart] = ref;
      ++nextChild;
    } else {
      uint32_t c = countOf(ref);
      m = (c == 1 ? 0x20u : c == 2 ? 0x60u : 0xE0u) | primOffset;
      // gather the (<= 3) primitives of this subtree
      uint32_t stack[4];
      int sp = 0;
      stack[sp++] = ref;
      while (sp) {
        uint32_t r = stack[--sp];
        if (r & kLeafBit) {
          leafPrim[primBase + primOffset] = sorted[r & ~kLeafBit];
          ++primOffset;
        } else {
          stack[sp++] = t.right[r];
          stack[sp++] = t.left[r];
        }
      }
    }
    meta[s >> 2] |= m << (8 * (s & 3));
  }
  WideNode node;
  quantiseNode(node, nlo, nhi, slotBox, present, imask);
  node.w[1] = make_uint4(childBase, primBase, meta[0], meta[1]);
  uint32_t idx = levelStart + i;
  nodes[idx] = node;
  nodeBox[2 * idx] = make_float4(nlo[0], nlo[1], nlo[2], 0.0f);
  nodeBox[2 * idx + 1] = make_float4(nhi[0], nhi[1], nhi[2], 0.0f);
}

// One thread per wide node of the current level (BLAS builds, and TLAS builds too large for the one-CTA builder).
__global__ void k_collapse_level(Bvh2 t, const float4 *primLo, const float4 *primHi, const uint32_t *sorted,
                                 const uint32_t *queueIn, uint32_t *queueOut, uint32_t levelStart,
                                 uint32_t levelCount, uint32_t nextLevelStart, CollapseCounters *counters,
                                 WideNode *nodes, float4 *nodeBox, uint32_t *leafPrim, uint32_t maxLeafPrims) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= levelCount) return;
  collapseOne(i, t, primLo, primHi, sorted, queueIn, queueOut, levelStart, nextLevelStart, counters, nodes, nodeBox,
              leafPrim, maxLeafPrims);
}

// BLAS leaves: write 48-byte triangle records in leaf order + remember where each came from (for refits).
__global__ void k_emit_triangles(const GeomEntry *geoms, uint32_t geomCount, const uint32_t *leafPrim, uint32_t n,
                                 TriRecord *tris, uint2 *triSource) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float3 a, b, c;
  uint32_t g, p;
  loadTriangle(geoms, geomCount, leafPrim[i], a, b, c, g, p);
  TriRecord r;
  r.v0 = make_float4(a.x, a.y, a.z, __uint_as_float(p));
  r.v1 = make_float4(b.x, b.y, b.z, __uint_as_float(g));
  r.v2 = make_float4(c.x, c.y, c.z, 0.0f);
  tris[i] = r;
  triSource[i] = make_uint2(g, p);
}

// Refit step 1: re-read the vertices of every triangle slot.
__global__ void k_refresh_triangles(const GeomEntry *geoms, uint32_t n, const uint2 *triSource, TriRecord *tris,
                                    const WideNode *nodes, uint32_t nodeCount, uint32_t *pending) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  // the per-refit child counters of the bottom-up pass ride along (there are fewer nodes than triangles)
  if (i < nodeCount) pending[i] = uint32_t(__popc(nodes[i].w[0].w >> 24));
  if (i >= n) return;
  uint2 src = triSource[i];
  const GeomEntry ge = geoms[src.x];
  uint32_t i0, i1, i2;
  if (ge.indexStride == 2) {
    const uint16_t *ix = reinterpret_cast<const uint16_t *>(ge.indices);
    i0 = ix[3 * src.y], i1 = ix[3 * src.y + 1], i2 = ix[3 * src.y + 2];
  } else {
    const uint32_t *ix = reinterpret_cast<const uint32_t *>(ge.indices);
    i0 = ix[3 * src.y], i1 = ix[3 * src.y + 1], i2 = ix[3 * src.y + 2];
  }
  const float *pa = reinterpret_cast<const float *>(ge.vertices + size_t(i0) * ge.vertexStride);
  const float *pb = reinterpret_cast<const float *>(ge.vertices + size_t(i1) * ge.vertexStride);
  const float *pc = reinterpret_cast<const float *>(ge.vertices + size_t(i2) * ge.vertexStride);
  TriRecord r;
  r.v0 = make_float4(pa[0], pa[1], pa[2], __uint_as_float(src.y));
  r.v1 = make_float4(pb[0], pb[1], pb[2], __uint_as_float(src.x));
  r.v2 = make_float4(pc[0], pc[1], pc[2], 0.0f);
  tris[i] = r;
}

// Refit of one wide node: recompute child boxes (leaf primitives, or child nodes that are already refitted), the exact
// node box and the quantised boxes. Topology words (w1, imask) are untouched. Child node boxes are read with ld.cg: they
// were written by other threads of the same launch. Leaf primitives are triangles (BLAS: `tris`) or instances (TLAS:
// the world boxes k_instance_bounds just wrote, through the leaf -> instance table).
struct TriangleLeaves {
  const TriRecord *tris;
  __device__ void grow(uint32_t slot, float lo[3], float hi[3]) const {
    const TriRecord tr = tris[slot];
    const float v[3][3] = {{tr.v0.x, tr.v0.y, tr.v0.z}, {tr.v1.x, tr.v1.y, tr.v1.z}, {tr.v2.x, tr.v2.y, tr.v2.z}};
    bool finite = true; // a non-finite triangle cannot be hit and must not blow up the box (see k_triangle_bounds)
    for (int q = 0; q < 3; ++q)
      for (int a = 0; a < 3; ++a) finite = finite && fabsf(v[q][a]) < 1.0e30f;
    if (!finite) return;
    for (int q = 0; q < 3; ++q)
      for (int a = 0; a < 3; ++a) {
        lo[a] = fminf(lo[a], v[q][a]);
        hi[a] = fmaxf(hi[a], v[q][a]);
      }
  }
};
struct InstanceLeaves {
  const float4 *primLo, *primHi;
  const uint32_t *leafPrim;
  __device__ void grow(uint32_t slot, float lo[3], float hi[3]) const {
    const uint32_t p = leafPrim[slot];
    const float4 l = primLo[p], h = primHi[p];
    lo[0] = fminf(lo[0], l.x), lo[1] = fminf(lo[1], l.y), lo[2] = fminf(lo[2], l.z);
    hi[0] = fmaxf(hi[0], h.x), hi[1] = fmaxf(hi[1], h.y), hi[2] = fmaxf(hi[2], h.z);
  }
};

template <typename Leaves>
__device__ void refitNode(WideNode *nodes, float4 *nodeBox, const Leaves &leaves, uint32_t idx) {
  WideNode node = nodes[idx];
  uint8_t imask = uint8_t(node.w[0].w >> 24);
  uint32_t childBase = node.w[1].x, primBase = node.w[1].y;
  ChildBox cb[8];
  uint8_t present[8];
  float nlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, nhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  uint32_t rank = 0;
  for (int s = 0; s < 8; ++s) {
    uint32_t m = ((s < 4 ? node.w[1].z : node.w[1].w) >> (8 * (s & 3))) & 0xFFu;
    present[s] = m != 0;
    if (!m) continue;
    if (imask & (1u << s)) {
      uint32_t c = childBase + rank++;
      float4 l = __ldcg(nodeBox + 2 * c), h = __ldcg(nodeBox + 2 * c + 1);
      cb[s].lo[0] = l.x, cb[s].lo[1] = l.y, cb[s].lo[2] = l.z;
      cb[s].hi[0] = h.x, cb[s].hi[1] = h.y, cb[s].hi[2] = h.z;
    } else {
      uint32_t cnt = (m >> 5) == 1 ? 1 : ((m >> 5) == 3 ? 2 : 3);
      uint32_t first = primBase + (m & 31u);
      for (int a = 0; a < 3; ++a) cb[s].lo[a] = FLT_MAX, cb[s].hi[a] = -FLT_MAX;
      for (uint32_t k = 0; k < cnt; ++k) leaves.grow(first + k, cb[s].lo, cb[s].hi);
    }
    if (cb[s].lo[0] > cb[s].hi[0]) // only non-finite triangles in this slot: an empty box at the origin
      for (int a = 0; a < 3; ++a) cb[s].lo[a] = cb[s].hi[a] = 0.0f;
    for (int a = 0; a < 3; ++a) {
      nlo[a] = fminf(nlo[a], cb[s].lo[a]);
      nhi[a] = fmaxf(nhi[a], cb[s].hi[a]);
    }
  }
  quantiseNode(node, nlo, nhi, cb, present, imask);
  nodes[idx] = node;
  nodeBox[2 * idx] = make_float4(nlo[0], nlo[1], nlo[2], 0.0f);
  nodeBox[2 * idx + 1] = make_float4(nhi[0], nhi[1], nhi[2], 0.0f);
}

// parent[] of every wide node (built once for a refittable BLAS) and the per-refit count of internal children
__global__ void k_node_parents(const WideNode *nodes, uint32_t nodeCount, uint32_t *parent) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nodeCount) return;
  if (i == 0) parent[0] = 0xFFFFFFFFu;
  const uint32_t internal = uint32_t(__popc(nodes[i].w[0].w >> 24)), childBase = nodes[i].w[1].x;
  for (uint32_t k = 0; k < internal; ++k) parent[childBase + k] = i;
}
// Refit step 2 in one launch: the walk starts at the nodes that have no internal children and goes up; whoever
// delivers a node's last child refits that node (the other deliverers stop). Replaces one launch per tree level —
// on a 100 k-vertex mesh the level launches were latency-bound and cost 0.15 ms per frame.
// One WARP per node: lane s < 8 owns child slot s — its box from the
// child node or from its <= 3 triangles, then its six quantised bytes — and the node box is a warp reduction. The
// arithmetic per child is that of refitNode / quantiseNode, so the nodes come out bit-identical; what changes is the
// length of the dependent chain: a node costs one child's work instead of eight children's, and the walk from the leaves
// to the root is ~7 such steps (one thread per node made the refit of a 100 k-vertex mesh a 0.14 ms latency chain).
__device__ __forceinline__ void quantiseAxis(float lo, float hi, uint32_t &ebyte, float &scale) {
  float ext = hi - lo;
  int e;
  if (!(ext > 0.0f)) {
    e = -126;
  } else {
    float m = frexpf(ext / 255.0f, &e); // ext/255 = m * 2^e, m in [0.5, 1)
    if (m == 0.5f) e -= 1;
    e = max(e, -126);
    while (double(ext) / ldexp(1.0, e) > 255.0) ++e;
  }
  e = min(e, 100);
  ebyte = uint32_t(e + 127);
  scale = __uint_as_float(ebyte << 23);
}

__device__ void refitNodeWarp(WideNode *nodes, float4 *nodeBox, const TriRecord *tris, uint32_t idx) {
  const unsigned full = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const uint4 w0 = nodes[idx].w[0], w1 = nodes[idx].w[1]; // same address for the whole warp: one transaction each
  const uint32_t imask = w0.w >> 24, childBase = w1.x, primBase = w1.y;
  const int s = lane & 7; // lanes 8..31 mirror lanes 0..7 (keeps the shuffles below full-warp)
  const uint32_t m = ((s < 4 ? w1.z : w1.w) >> (8 * (s & 3))) & 0xFFu;
  const bool present = m != 0u;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (present) {
    if (imask & (1u << s)) {
      const uint32_t c = childBase + uint32_t(__popc(imask & ((1u << s) - 1u)));
      const float4 l = __ldcg(nodeBox + 2 * c), h = __ldcg(nodeBox + 2 * c + 1);
      lo[0] = l.x, lo[1] = l.y, lo[2] = l.z, hi[0] = h.x, hi[1] = h.y, hi[2] = h.z;
    } else {
      const uint32_t cnt = (m >> 5) == 1 ? 1 : ((m >> 5) == 3 ? 2 : 3);
      const TriangleLeaves leaves{tris};
      for (uint32_t k = 0; k < cnt; ++k) leaves.grow(primBase + (m & 31u) + k, lo, hi);
    }
    if (lo[0] > hi[0]) // only non-finite triangles in this slot: an empty box at the origin
      for (int a = 0; a < 3; ++a) lo[a] = hi[a] = 0.0f;
  }
  float nlo[3], nhi[3];
  for (int a = 0; a < 3; ++a) {
    nlo[a] = present ? lo[a] : FLT_MAX;
    nhi[a] = present ? hi[a] : -FLT_MAX;
    for (int o = 4; o > 0; o >>= 1) { // over the eight slots (the three mirror groups reduce the same values)
      nlo[a] = fminf(nlo[a], __shfl_xor_sync(full, nlo[a], o));
      nhi[a] = fmaxf(nhi[a], __shfl_xor_sync(full, nhi[a], o));
    }
  }
  uint32_t ebyte[3];
  float scale[3];
  for (int a = 0; a < 3; ++a) quantiseAxis(nlo[a], nhi[a], ebyte[a], scale[a]);
  uint32_t q[6] = {0, 0, 0, 0, 0, 0}; // this slot's bytes: lo xyz, hi xyz
  if (present) {
    for (int a = 0; a < 3; ++a) {
      const double sc = double(scale[a]), org = double(nlo[a]);
      int ql = int(floor((double(lo[a]) - org) / sc));
      int qh = int(ceil((double(hi[a]) - org) / sc));
      ql = max(0, min(255, ql));
      qh = max(0, min(255, qh));
      while (ql > 0 && org + double(ql) * sc > double(lo[a])) --ql;
      while (qh < 255 && org + double(qh) * sc < double(hi[a])) ++qh;
      if (qh < ql) qh = ql;
      q[a] = uint32_t(ql), q[3 + a] = uint32_t(qh);
    }
  }
  // words of the node: array a = two words, byte (s & 3) of word (s >> 2) belongs to slot s
  uint32_t word[6][2];
  for (int a = 0; a < 6; ++a) {
    const uint32_t b0 = __shfl_sync(full, q[a], 0), b1 = __shfl_sync(full, q[a], 1), b2 = __shfl_sync(full, q[a], 2),
                   b3 = __shfl_sync(full, q[a], 3), b4 = __shfl_sync(full, q[a], 4), b5 = __shfl_sync(full, q[a], 5),
                   b6 = __shfl_sync(full, q[a], 6), b7 = __shfl_sync(full, q[a], 7);
    word[a][0] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    word[a][1] = b4 | (b5 << 8) | (b6 << 16) | (b7 << 24);
  }
  if (lane == 0) {
    nodes[idx].w[0] = make_uint4(__float_as_uint(nlo[0]), __float_as_uint(nlo[1]), __float_as_uint(nlo[2]),
                                 ebyte[0] | (ebyte[1] << 8) | (ebyte[2] << 16) | (imask << 24));
    nodes[idx].w[2] = make_uint4(word[0][0], word[0][1], word[1][0], word[1][1]);
    nodes[idx].w[3] = make_uint4(word[2][0], word[2][1], word[3][0], word[3][1]);
    nodes[idx].w[4] = make_uint4(word[4][0], word[4][1], word[5][0], word[5][1]);
    nodeBox[2 * idx] = make_float4(nlo[0], nlo[1], nlo[2], 0.0f);
    nodeBox[2 * idx + 1] = make_float4(nhi[0], nhi[1], nhi[2], 0.0f);
  }
}

__global__ void k_refit_bottom_up_warp(WideNode *nodes, float4 *nodeBox, const TriRecord *tris, uint32_t nodeCount,
                                       const uint32_t *parent, uint32_t *pending, BlasHeader *header, uint32_t triCount) {
  const unsigned full = 0xFFFFFFFFu;
  uint32_t node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (node >= nodeCount || (nodes[node].w[0].w >> 24) != 0u) return; // whole warps leave together
  while (true) {
    refitNodeWarp(nodes, nodeBox, tris, node);
    uint32_t next = 0xFFFFFFFFu;
    if ((threadIdx.x & 31) == 0) {
      __threadfence();
      const uint32_t p = parent[node];
      if (p != 0xFFFFFFFFu && atomicSub(pending + p, 1u) == 1u) next = p; // delivered the last child: refit the parent
      if (p == 0xFFFFFFFFu) { // the root is done: the header's bounds are its box (what k_write_blas_header writes)
        const float4 lo = nodeBox[0], hi = nodeBox[1];
        header->nodes = nodes;
        header->tris = tris;
        header->triCount = triCount;
        header->nodeCount = nodeCount;
        header->boundsLo[0] = lo.x, header->boundsLo[1] = lo.y, header->boundsLo[2] = lo.z;
        header->boundsHi[0] = hi.x, header->boundsHi[1] = hi.y, header->boundsHi[2] = hi.z;
      }
    }
    next = __shfl_sync(full, next, 0);
    if (next == 0xFFFFFFFFu) return;
    node = next;
  }
}

__global__ void k_write_blas_header(BlasHeader *h, const WideNode *nodes, const TriRecord *tris,
                                    const float4 *nodeBox, uint32_t triCount, uint32_t nodeCount) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  h->nodes = nodes;
  h->tris = tris;
  h->triCount = triCount;
  h->nodeCount = nodeCount;
  float4 lo = nodeBox[0], hi = nodeBox[1];
  h->boundsLo[0] = lo.x, h->boundsLo[1] = lo.y, h->boundsLo[2] = lo.z;
  h->boundsHi[0] = hi.x, h->boundsHi[1] = hi.y, h->boundsHi[2] = hi.z;
}

__global__ void k_write_tlas_header(TlasHeader *h, const WideNode *nodes, const InstanceRecord *instances,
                                    const uint32_t *leafInstance, const float4 *instanceBox, uint32_t instanceCount,
                                    uint32_t nodeCount) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  h->nodes = nodes;
  h->instances = instances;
  h->leafInstance = leafInstance;
  h->instanceBox = instanceBox;
  h->instanceCount = instanceCount;
  h->nodeCount = nodeCount;
}

// TLAS over at most eight instances (the reference's scenes have 2 - 8): the whole tree is one wide node whose
// slots are the instances, written by one thread. No sort, no hierarchy, no host round trip — the general builder
// costs ~0.1 ms per frame in launches and level read-backs, which is 5 - 10 % of a small frame.
__global__ void k_tlas_small(const float4 *primLo, const float4 *primHi, uint32_t count, WideNode *nodes, float4 *nodeBox,
                             uint32_t *leafPrim) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ChildBox cb[8];
  uint8_t present[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float nlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, nhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  uint32_t meta[2] = {0, 0};
  for (uint32_t s = 0; s < count; ++s) {
    const float4 l = primLo[s], h = primHi[s];
    cb[s].lo[0] = l.x, cb[s].lo[1] = l.y, cb[s].lo[2] = l.z;
    cb[s].hi[0] = h.x, cb[s].hi[1] = h.y, cb[s].hi[2] = h.z;
    for (int a = 0; a < 3; ++a) {
      nlo[a] = fminf(nlo[a], cb[s].lo[a]);
      nhi[a] = fmaxf(nhi[a], cb[s].hi[a]);
    }
    present[s] = 1;
    leafPrim[s] = s;
    meta[s >> 2] |= (0x20u | s) << (8 * (s & 3)); // leaf slot with one primitive at offset s
  }
  WideNode node;
  quantiseNode(node, nlo, nhi, cb, present, 0);
  node.w[1] = make_uint4(0u, 0u, meta[0], meta[1]);
  nodes[0] = node;
  nodeBox[0] = make_float4(nlo[0], nlo[1], nlo[2], 0.0f);
  nodeBox[1] = make_float4(nhi[0], nhi[1], nhi[2], 0.0f);
}

// ---- TLAS built by one CTA, no host round trips ---------------------------------------------------------------
// The general builder above reads a counter back after every PLOC round and every collapse level (fine for a BLAS,
// which is built once); a TLAS is updated every frame (Renderer.swift:937-973,1084-1202), so for up to kTlasCtaMax
// instances the hierarchy (PLOC over the Morton-sorted instance boxes), its collapse into wide nodes and the parent
// links a later refit needs are produced by a single 1024-thread CTA that iterates on the device: rounds and levels
// are separated by __syncthreads() instead of launches. 4096 instances take ~25 rounds of 128 box pairs per thread.
constexpr uint32_t kTlasCtaMax = 1u << 16;
constexpr int kTlasCtaThreads = 1024;
constexpr uint32_t kMaxTlasLevels = 20, kMaxBlasLevels = 24; // + 2 for an instance entry must fit kStackSize = 48 (traverse.cuh)

// inclusive scan of one 64-bit value per thread over the CTA (warp shuffles + one shared array); every thread calls it
__device__ unsigned long long blockInclusiveScan(unsigned long long v, unsigned long long *warpTotals, unsigned long long &total) {
  const unsigned full = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(full, v, o);
    if (lane >= o) v += up;
  }
  if (lane == 31) warpTotals[warp] = v;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = lane < warps ? warpTotals[lane] : 0ull;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(full, w, o);
      if (lane >= o) w += up;
    }
    warpTotals[lane] = w; // inclusive totals of warps 0..lane
  }
  __syncthreads();
  if (warp > 0) v += warpTotals[warp - 1];
  total = warpTotals[warps - 1];
  __syncthreads(); // warpTotals may be reused by the next call
  return v;
}

__global__ void __launch_bounds__(kTlasCtaThreads) k_tlas_build_cta(uint32_t n, const float4 *primLo, const float4 *primHi,
                                                                     const uint32_t *sorted, Bvh2 t, PlocClusters ca,
                                                                     PlocClusters cb, uint32_t *nearest, int radius,
                                                                     uint32_t *queueA, uint32_t *queueB,
                                                                     CollapseCounters *counters, WideNode *nodes,
                                                                     float4 *nodeBox, uint32_t *leafPrim,
                                                                     uint32_t maxLeafPrims, uint32_t *parent,
                                                                     TlasBuildInfo *info) {
  __shared__ unsigned long long s_warpTotals[32];
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  // PLOC (same rounds as k_ploc_nearest / k_ploc_flags / scan / k_ploc_merge, hence the same tree)
  for (uint32_t i = tid; i < n; i += T) {
    const uint32_t p = sorted[i];
    ca.ref[i] = i | kLeafBit;
    ca.lo[i] = primLo[p];
    ca.hi[i] = primHi[p];
  }
  __syncthreads();
  uint32_t count = n, nodeBase = 0;
  PlocClusters cur = ca, nxt = cb;
  while (count > 1) {
    for (uint32_t i = tid; i < count; i += T) {
      const float4 lo = cur.lo[i], hi = cur.hi[i];
      const int first = max(0, int(i) - radius), last = min(int(count) - 1, int(i) + radius);
      float best = FLT_MAX;
      uint32_t bestJ = i;
      for (int j = first; j <= last; ++j) {
        if (j == int(i)) continue;
        const float4 l = cur.lo[j], h = cur.hi[j];
        const float dx = fmaxf(hi.x, h.x) - fminf(lo.x, l.x), dy = fmaxf(hi.y, h.y) - fminf(lo.y, l.y),
                    dz = fmaxf(hi.z, h.z) - fminf(lo.z, l.z);
        const float area = dx * dy + dy * dz + dz * dx;
        if (area < best) {
          best = area;
          bestJ = uint32_t(j);
        }
      }
      nearest[i] = bestJ;
    }
    __syncthreads();
    unsigned long long carry = 0ull; // survivors (low word) and merges (high word) of the tiles before this one
    for (uint32_t base = 0; base < count; base += T) {
      const uint32_t i = base + tid;
      bool mutual = false;
      uint32_t j = i;
      unsigned long long flag = 0ull;
      if (i < count) {
        j = nearest[i];
        mutual = j != i && nearest[j] == i;
        const bool merges = mutual && i < j, dies = mutual && i > j;
        flag = (dies ? 0ull : 1ull) | (merges ? (1ull << 32) : 0ull);
      }
      unsigned long long tileTotal;
      const unsigned long long inc = carry + blockInclusiveScan(flag, s_warpTotals, tileTotal);
      carry += tileTotal;
      if (i < count && !(mutual && i > j)) {
        const uint32_t pos = uint32_t(inc & 0xFFFFFFFFull) - 1u;
        float4 lo = cur.lo[i], hi = cur.hi[i];
        uint32_t ref = cur.ref[i];
        if (mutual) {
          const uint32_t node = nodeBase + uint32_t(inc >> 32) - 1u;
          const float4 l = cur.lo[j], h = cur.hi[j];
          const uint32_t rj = cur.ref[j];
          lo = make_float4(fminf(lo.x, l.x), fminf(lo.y, l.y), fminf(lo.z, l.z), 0.0f);
          hi = make_float4(fmaxf(hi.x, h.x), fmaxf(hi.y, h.y), fmaxf(hi.z, h.z), 0.0f);
          t.left[node] = ref;
          t.right[node] = rj;
          t.lo[node] = lo;
          t.hi[node] = hi;
          t.count[node] = ((ref & kLeafBit) ? 1u : t.count[ref]) + ((rj & kLeafBit) ? 1u : t.count[rj]);
          ref = node;
        }
        nxt.ref[pos] = ref;
        nxt.lo[pos] = lo;
        nxt.hi[pos] = hi;
      }
    }
    __syncthreads();
    const uint32_t survivors = uint32_t(carry & 0xFFFFFFFFull), merges = uint32_t(carry >> 32);
    if (merges == 0u) break; // cannot happen (the globally closest pair is always mutual); never spin
    nodeBase += merges;
    count = survivors;
    const PlocClusters tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  // collapse, one level per trip
  if (tid == 0) {
    counters->nodeCount = 1u;
    counters->primCount = 0u;
    queueA[0] = n > 1 ? n - 2 : kLeafBit; // the last node PLOC created is the root
  }
  __syncthreads();
  uint32_t levelStart = 0, levelCount = 1, levels = 0;
  uint32_t *qin = queueA, *qout = queueB;
  while (levelCount) {
    const uint32_t nextStart = levelStart + levelCount;
    for (uint32_t i = tid; i < levelCount; i += T)
      collapseOne(i, t, primLo, primHi, sorted, qin, qout, levelStart, nextStart, counters, nodes, nodeBox, leafPrim,
                  maxLeafPrims);
    __syncthreads();
    const uint32_t now = *reinterpret_cast<volatile uint32_t *>(&counters->nodeCount);
    __syncthreads();
    levelStart = nextStart;
    levelCount = now - nextStart;
    ++levels;
    uint32_t *tmp = qin;
    qin = qout;
    qout = tmp;
  }
  const uint32_t nodeCount = levelStart;
  // parent links for refits
  for (uint32_t i = tid; i < nodeCount; i += T) {
    if (i == 0) parent[0] = 0xFFFFFFFFu;
    const uint32_t internal = uint32_t(__popc(nodes[i].w[0].w >> 24)), childBase = nodes[i].w[1].x;
    for (uint32_t k = 0; k < internal; ++k) parent[childBase + k] = i;
  }
  if (tid == 0) {
    info->nodeCount = nodeCount;
    info->levelCount = levels;
    info->status = levels > kMaxTlasLevels ? 1u : 0u;
    info->builds += 1u;
  }
}

// TLAS refit (Renderer.swift:1084-1202 refits the instance AS when the device supports it): topology, leaf order and
// node count stay; instance records and world boxes come fresh from the descriptors (k_instance_bounds), node boxes are
// recomputed bottom-up exactly like a BLAS refit. The node count lives on the device (TlasBuildInfo).
__global__ void k_tlas_refit_pending(const WideNode *nodes, const TlasBuildInfo *info, uint32_t *pending) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < info->nodeCount) pending[i] = uint32_t(__popc(nodes[i].w[0].w >> 24));
}
__global__ void k_tlas_refit_bottom_up(WideNode *nodes, float4 *nodeBox, const float4 *primLo, const float4 *primHi,
                                       const uint32_t *leafPrim, const TlasBuildInfo *info, const uint32_t *parent,
                                       uint32_t *pending) {
  uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= info->nodeCount || (nodes[node].w[0].w >> 24) != 0u) return;
  const InstanceLeaves leaves{primLo, primHi, leafPrim};
  while (true) {
    refitNode(nodes, nodeBox, leaves, node);
    __threadfence();
    const uint32_t p = parent[node];
    if (p == 0xFFFFFFFFu) return;
    if (atomicSub(pending + p, 1u) != 1u) return;
    node = p;
  }
}

struct Bump {
  uint8_t *base;
  size_t offset = 0, capacity;
  template <typename T>
  T *take(size_t count) {
    offset = (offset + 255) & ~size_t(255);
    T *p = reinterpret_cast<T *>(base + offset);
    offset += count * sizeof(T);
    return p;
  }
};

size_t scratchNeed(uint32_t n, size_t cubBytes) {
  size_t per = 16 + 16 + 8 + 8 + 4 + 4 + (4 + 4 + 8 + 16 + 16 + 4 + 4) + 4 + 4 + 4;
  per += 2 * (4 + 16 + 16) + 4 + 8 + 8; // PLOC: two cluster arrays, nearest, flags, scan
  return size_t(n) * per + 2 * cubBytes + 128 * 1024;
}

inline uint32_t gridFor(uint32_t n, uint32_t block) { return (n + block - 1) / block; }

} // namespace

// Builds the wide tree over `n` primitives whose boxes are already in primLo/primHi (scratch). Fills as->nodes,
// as->nodeBox, as->levelStart, as->nodeCount and leafPrim (device array of n primitive ids in leaf order).
static int buildWideTree(rt_context *ctx, AccelObject *as, uint32_t n, const float4 *primLo, const float4 *primHi,
                         BoundsAtomics *bounds, Bump &bump, uint32_t *leafPrim, int plocRadius) {
  cudaStream_t st = ctx->stream;
  const uint32_t B = 256;
  uint64_t *keysA = bump.take<uint64_t>(n), *keysB = bump.take<uint64_t>(n);
  uint32_t *valsA = bump.take<uint32_t>(n), *valsB = bump.take<uint32_t>(n);
  k_morton<<<gridFor(n, B), B, 0, st>>>(primLo, primHi, n, bounds, keysA, valsA);
  ++ctx->launches;
  size_t cubBytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, keysA, keysB, valsA, valsB, int(n), 0, 63, st);
  void *cubTemp = bump.take<uint8_t>(cubBytes);
  RT_CUDA(cub::DeviceRadixSort::SortPairs(cubTemp, cubBytes, keysA, keysB, valsA, valsB, int(n), 0, 63, st));
  ctx->launches += 4; // CUB's histogram + onesweep passes (library kernels)
  const uint64_t *keys = keysB;
  const uint32_t *sorted = valsB;

  Bvh2 t{};
  t.n = n;
  uint32_t ni = n > 1 ? n - 1 : 1;
  t.left = bump.take<uint32_t>(ni);
  t.right = bump.take<uint32_t>(ni);
  t.parent = bump.take<uint32_t>(size_t(2) * n);
  t.lo = bump.take<float4>(ni);
  t.hi = bump.take<float4>(ni);
  t.count = bump.take<uint32_t>(ni);
  t.flag = bump.take<uint32_t>(ni);
  uint32_t rootRef = n > 1 ? 0u : kLeafBit;
  if (n > 1 && plocRadius > 0) {
    PlocClusters ca{bump.take<uint32_t>(n), bump.take<float4>(n), bump.take<float4>(n)};
    PlocClusters cb{bump.take<uint32_t>(n), bump.take<float4>(n), bump.take<float4>(n)};
    uint32_t *nearest = bump.take<uint32_t>(n);
    unsigned long long *flags = bump.take<unsigned long long>(n), *scan = bump.take<unsigned long long>(n);
    size_t scanBytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, scanBytes, flags, scan, int(n), st);
    void *scanTemp = bump.take<uint8_t>(scanBytes);
    RT_CHECK(bump.offset <= bump.capacity, "internal: build scratch overflow (PLOC)");
    k_ploc_init<<<gridFor(n, B), B, 0, st>>>(primLo, primHi, sorted, n, ca);
    ++ctx->launches;
    uint32_t count = n, nodeBase = 0;
    PlocClusters *cur = &ca, *nxt = &cb;
    while (count > 1) {
      k_ploc_nearest<<<gridFor(count, B), B, 0, st>>>(*cur, count, plocRadius, nearest);
      k_ploc_flags<<<gridFor(count, B), B, 0, st>>>(nearest, count, flags);
      RT_CUDA(cub::DeviceScan::InclusiveSum(scanTemp, scanBytes, flags, scan, int(count), st));
      k_ploc_merge<<<gridFor(count, B), B, 0, st>>>(*cur, *nxt, nearest, scan, count, nodeBase, t);
      ctx->launches += 5;
      unsigned long long totals = 0;
      RT_CUDA(cudaMemcpyAsync(&totals, scan + (count - 1), sizeof totals, cudaMemcpyDeviceToHost, st));
      RT_CUDA(RT_SYNC_STREAM(ctx, st));
      const uint32_t survivors = uint32_t(totals & 0xFFFFFFFFull), merges = uint32_t(totals >> 32);
      RT_CHECK(merges > 0 && survivors == count - merges, "internal: PLOC round made no progress");
      nodeBase += merges;
      count = survivors;
      std::swap(cur, nxt);
    }
    RT_CHECK(nodeBase == n - 1, "internal: PLOC node count");
    rootRef = n - 2; // the last node created
  } else if (n > 1) {
    k_karras<<<gridFor(n - 1, B), B, 0, st>>>(keys, t);
    k_bvh2_bounds<<<gridFor(n, B), B, 0, st>>>(primLo, primHi, sorted, t);
    ctx->launches += 2;
  }
  uint32_t *queueA = bump.take<uint32_t>(n), *queueB = bump.take<uint32_t>(n);
  CollapseCounters *counters = bump.take<CollapseCounters>(1);
  RT_CHECK(bump.offset <= bump.capacity, "internal: build scratch overflow");

  CollapseCounters init{1u, 0u};
  RT_CUDA(cudaMemcpyAsync(counters, &init, sizeof init, cudaMemcpyHostToDevice, st));
  RT_CUDA(cudaMemcpyAsync(queueA, &rootRef, 4, cudaMemcpyHostToDevice, st));
  RT_CUDA(RT_SYNC_STREAM(ctx, st)); // rootRef / init are stack variables
  as->levelStart.clear();
  uint32_t levelStart = 0, levelCount = 1;
  uint32_t *qin = queueA, *qout = queueB;
  while (levelCount) {
    as->levelStart.push_back(levelStart);
    uint32_t nextStart = levelStart + levelCount;
    RT_CHECK(nextStart <= as->nodeCapacity, "internal: wide node capacity exceeded");
    k_collapse_level<<<gridFor(levelCount, 128), 128, 0, st>>>(t, primLo, primHi, sorted, qin, qout, levelStart,
                                                                levelCount, nextStart, counters, as->nodes,
                                                                as->nodeBox, leafPrim,
                                                                uint32_t(std::min(std::max(as->isTlas ? ctx->tlasLeafSize : ctx->leafSize, 1), kMaxLeafPrims)));
    ++ctx->launches;
    CollapseCounters now;
    RT_CUDA(cudaMemcpyAsync(&now, counters, sizeof now, cudaMemcpyDeviceToHost, st));
    RT_CUDA(RT_SYNC_STREAM(ctx, st));
    levelStart = nextStart;
    levelCount = now.nodeCount - nextStart;
    std::swap(qin, qout);
  }
  as->levelStart.push_back(levelStart);
  as->nodeCount = levelStart;
  return 0;
}

static int uploadGeomTable(rt_context *ctx, AccelObject *as, const rt_triangle_geometry *geoms, uint32_t n,
                           uint32_t *totalOut) {
  std::vector<GeomEntry> table(n);
  uint32_t total = 0;
  for (uint32_t g = 0; g < n; ++g) {
    RT_CHECK(geoms[g].indexStride == 2 || geoms[g].indexStride == 4, "rt_blas: indexStride must be 2 or 4");
    RT_CHECK(geoms[g].vertexStride >= 12 && geoms[g].vertexStride % 4 == 0, "rt_blas: vertexStride must be >= 12 and a multiple of 4");
    RT_CHECK(geoms[g].triangleCount == 0 || (geoms[g].vertexBuffer && geoms[g].indexBuffer), "rt_blas: null geometry buffer");
    table[g] = {static_cast<const uint8_t *>(geoms[g].vertexBuffer), static_cast<const uint8_t *>(geoms[g].indexBuffer),
                geoms[g].vertexStride, geoms[g].indexStride, total, geoms[g].triangleCount};
    total += geoms[g].triangleCount;
  }
  *totalOut = total;
  const size_t bytes = size_t(n) * sizeof(GeomEntry);
  if (as->geomTableDev && as->geomCount == n && as->geomTableHost.size() == bytes &&
      (bytes == 0 || std::memcmp(as->geomTableHost.data(), table.data(), bytes) == 0))
    return 0; // same buffers as last time (the per-frame refit of a skinned mesh): nothing to upload, no sync
  if (!as->geomTableDev || as->geomCount != n) {
    if (as->geomTableDev) cudaFree(as->geomTableDev);
    as->geomTableDev = nullptr;
    RT_CUDA(cudaMalloc(&as->geomTableDev, std::max<size_t>(1, n) * sizeof(GeomEntry)));
    as->geomCount = n;
  }
  if (n) RT_CUDA(cudaMemcpyAsync(as->geomTableDev, table.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
  RT_CUDA(RT_SYNC_STREAM(ctx, ctx->stream)); // `table` is a stack-lifetime staging buffer
  as->geomTableHost.assign(reinterpret_cast<const uint8_t *>(table.data()), reinterpret_cast<const uint8_t *>(table.data()) + bytes);
  return 0;
}

static int shrinkNodes(rt_context *ctx, AccelObject *as) {
  if (as->nodeCount == as->nodeCapacity) return 0;
  uint32_t cap = std::max(1u, as->nodeCount);
  WideNode *nodes = nullptr;
  float4 *boxes = nullptr;
  RT_CUDA(cudaMalloc(&nodes, size_t(cap) * sizeof(WideNode)));
  RT_CUDA(cudaMalloc(&boxes, size_t(cap) * 2 * sizeof(float4)));
  RT_CUDA(cudaMemcpyAsync(nodes, as->nodes, size_t(cap) * sizeof(WideNode), cudaMemcpyDeviceToDevice, ctx->stream));
  RT_CUDA(cudaMemcpyAsync(boxes, as->nodeBox, size_t(cap) * 2 * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  RT_CUDA(RT_SYNC_STREAM(ctx, ctx->stream));
  cudaFree(as->nodes);
  cudaFree(as->nodeBox);
  as->nodes = nodes;
  as->nodeBox = boxes;
  as->nodeCapacity = cap;
  return 0;
}

static int finishInfo(rt_context *ctx, AccelObject *as) {
  // bounds + SAH cost from the exact node boxes (host side; build-time only)
  std::vector<float4> boxes(size_t(as->nodeCount) * 2);
  std::vector<WideNode> nodes(as->nodeCount);
  if (as->nodeCount) {
    RT_CUDA(cudaMemcpyAsync(boxes.data(), as->nodeBox, boxes.size() * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(cudaMemcpyAsync(nodes.data(), as->nodes, nodes.size() * sizeof(WideNode), cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(RT_SYNC_STREAM(ctx, ctx->stream));
  }
  auto area = [&](uint32_t i) {
    float dx = boxes[2 * i + 1].x - boxes[2 * i].x, dy = boxes[2 * i + 1].y - boxes[2 * i].y,
          dz = boxes[2 * i + 1].z - boxes[2 * i].z;
    return std::max(0.0f, dx * dy + dy * dz + dz * dx);
  };
  double cost = 0.0;
  if (as->nodeCount) {
    as->bounds = {{boxes[0].x, boxes[0].y, boxes[0].z}, {boxes[1].x, boxes[1].y, boxes[1].z}};
    double rootArea = std::max(1e-30f, area(0));
    for (uint32_t i = 0; i < as->nodeCount; ++i) {
      int prims = 0;
      for (int s = 0; s < 8; ++s) {
        uint32_t m = ((s < 4 ? nodes[i].w[1].z : nodes[i].w[1].w) >> (8 * (s & 3))) & 0xFFu;
        bool internal = (nodes[i].w[0].w >> 24) & (1u << s);
        if (m && !internal) prims += (m >> 5) == 1 ? 1 : ((m >> 5) == 3 ? 2 : 3);
      }
      cost += double(area(i)) / rootArea * (1.0 + 0.3 * prims); // node visit = 1, triangle test = 0.3
    }
  }
  as->sahCost = float(cost);
  return 0;
}

int buildBlas(rt_context *ctx, const rt_triangle_geometry *geoms, uint32_t geomCount, uint32_t flags,
              AccelObject **out) {
  AccelObject *as = new AccelObject();
  as->isTlas = false;
  as->flags = flags;
  uint32_t n = 0;
  int rc = uploadGeomTable(ctx, as, geoms, geomCount, &n);
  if (rc) {
    destroyAccel(as);
    return rc;
  }
  cudaStream_t st = ctx->stream;
  as->primCount = as->primCapacity = n;
  auto fail = [&](int code) {
    destroyAccel(as);
    return code;
  };
#define RT_TRYF(expr)             \
  do {                            \
    int _r = (expr);              \
    if (_r != 0) return fail(_r); \
  } while (0)
#define RT_CUDAF(expr)                                                                               \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      setError(std::string(#expr) + " failed: " + cudaGetErrorString(_e));                           \
      return fail(1);                                                                                \
    }                                                                                                \
  } while (0)
  RT_CUDAF(cudaMalloc(&as->headerDev, sizeof(BlasHeader)));
  RT_CUDAF(cudaMemsetAsync(as->headerDev, 0, sizeof(BlasHeader), st));
  if (n == 0) { // empty mesh: header with triCount 0, instances of it are skipped
    *out = as;
    as->bytes = sizeof(BlasHeader);
    return 0;
  }
  as->nodeCapacity = n;
  RT_CUDAF(cudaMalloc(&as->nodes, size_t(as->nodeCapacity) * sizeof(WideNode)));
  RT_CUDAF(cudaMalloc(&as->nodeBox, size_t(as->nodeCapacity) * 2 * sizeof(float4)));
  RT_CUDAF(cudaMalloc(&as->tris, size_t(n) * sizeof(TriRecord)));
  RT_CUDAF(cudaMalloc(&as->triSource, size_t(n) * sizeof(uint2)));
  size_t cubBytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr,
                                  (uint32_t *)nullptr, int(n), 0, 63, st);
  RT_TRYF(ensureScratch(ctx, scratchNeed(n, cubBytes)));
  Bump bump{static_cast<uint8_t *>(ctx->scratch), 0, ctx->scratchBytes};
  float4 *primLo = bump.take<float4>(n), *primHi = bump.take<float4>(n);
  BoundsAtomics *bounds = bump.take<BoundsAtomics>(1);
  uint32_t *leafPrim = bump.take<uint32_t>(n);
  const GeomEntry *table = static_cast<const GeomEntry *>(as->geomTableDev);
  k_init_bounds<<<1, 32, 0, st>>>(bounds);
  k_triangle_bounds<<<gridFor(n, 256), 256, 0, st>>>(table, geomCount, n, primLo, primHi, bounds);
  ctx->launches += 2;
  {
    // the traversal stack (traverse.cuh kStackSize) holds TLAS levels + 2 + BLAS levels entries: a hierarchy deeper than
    // kMaxBlasLevels wide levels (PLOC can chain on adversarial input) is rebuilt as a plain LBVH, whose depth is bounded
    // by the Morton key length; failing that the build is refused instead of dropping subtrees during traversal
    const Bump saved = bump;
    RT_TRYF(buildWideTree(ctx, as, n, primLo, primHi, bounds, bump, leafPrim, ctx->plocRadius));
    if (as->levelStart.size() - 1 > kMaxBlasLevels && ctx->plocRadius > 0) {
      bump = saved;
      RT_TRYF(buildWideTree(ctx, as, n, primLo, primHi, bounds, bump, leafPrim, 0));
    }
    if (as->levelStart.size() - 1 > kMaxBlasLevels) {
      setError("rt_blas_build: hierarchy is " + std::to_string(as->levelStart.size() - 1) + " levels deep; the traversal stack holds " +
               std::to_string(kMaxBlasLevels));
      return fail(2);
    }
  }
  k_emit_triangles<<<gridFor(n, 256), 256, 0, st>>>(table, geomCount, leafPrim, n, as->tris, as->triSource);
  ++ctx->launches;
  RT_CUDAF(RT_SYNC_STREAM(ctx, st));
  RT_TRYF(shrinkNodes(ctx, as));
  k_write_blas_header<<<1, 32, 0, st>>>(static_cast<BlasHeader *>(as->headerDev), as->nodes, as->tris, as->nodeBox, n,
                                        as->nodeCount);
  ++ctx->launches;
  RT_CUDAF(cudaGetLastError());
  RT_TRYF(finishInfo(ctx, as));
  if (!(flags & RT_AS_FLAG_REFITTABLE)) { // compaction: drop what only a refit needs
    cudaFree(as->triSource);
    as->triSource = nullptr;
  } else { // refit walks the tree bottom-up in one launch: parent links + a per-refit counter per node
    RT_CUDAF(cudaMalloc(&as->nodeParent, size_t(as->nodeCount) * sizeof(uint32_t)));
    RT_CUDAF(cudaMalloc(&as->nodePending, size_t(as->nodeCount) * sizeof(uint32_t)));
    k_node_parents<<<gridFor(as->nodeCount, 256), 256, 0, st>>>(as->nodes, as->nodeCount, as->nodeParent);
    ++ctx->launches;
    RT_CUDAF(cudaGetLastError());
  }
  as->bytes = sizeof(BlasHeader) + size_t(as->nodeCapacity) * (sizeof(WideNode) + 2 * sizeof(float4)) +
              size_t(n) * sizeof(TriRecord) + (as->triSource ? size_t(n) * sizeof(uint2) : 0);
  *out = as;
  return 0;
#undef RT_TRYF
#undef RT_CUDAF
}

int refitBlas(rt_context *ctx, AccelObject *as, const rt_triangle_geometry *geoms, uint32_t geomCount) {
  RT_CHECK(!as->isTlas, "rt_blas_refit: id is a TLAS");
  RT_CHECK(as->flags & RT_AS_FLAG_REFITTABLE, "rt_blas_refit: BLAS was not built with RT_AS_FLAG_REFITTABLE");
  RT_CHECK(geomCount == as->geomCount, "rt_blas_refit: geometry count differs from the build");
  uint32_t n = 0;
  RT_TRY(uploadGeomTable(ctx, as, geoms, geomCount, &n));
  RT_CHECK(n == as->primCount, "rt_blas_refit: triangle count differs from the build");
  if (n == 0) return 0;
  cudaStream_t st = ctx->stream;
  // two launches: triangle records + child counters, then the bottom-up pass (whose last warp also writes the header)
  k_refresh_triangles<<<gridFor(std::max(n, as->nodeCount), 256), 256, 0, st>>>(static_cast<const GeomEntry *>(as->geomTableDev), n,
                                                                               as->triSource, as->tris, as->nodes, as->nodeCount,
                                                                               as->nodePending);
  k_refit_bottom_up_warp<<<gridFor(as->nodeCount * 32u, 128), 128, 0, st>>>(as->nodes, as->nodeBox, as->tris, as->nodeCount,
                                                                            as->nodeParent, as->nodePending,
                                                                            static_cast<BlasHeader *>(as->headerDev), n);
  ctx->launches += 2;
  RT_CUDA(cudaGetLastError());
  return 0;
}

// Looks at the result of the last device-side TLAS build once its copy has landed (wait = block for it). A tree deeper
// than the traversal stack allows is an error reported here — one library call after the build, never silently.
int checkTlasInfo(rt_context *ctx, AccelObject *as, bool wait) {
  if (!as->infoPending) return 0;
  if (wait) {
    RT_CUDA(RT_SYNC_EVENT(ctx, as->infoEvent));
  } else {
    const cudaError_t q = cudaEventQuery(as->infoEvent);
    if (q == cudaErrorNotReady) return 0;
    RT_CUDA(q);
  }
  as->infoPending = false;
  as->nodeCount = as->infoHost->nodeCount;
  RT_CHECK(as->infoHost->status == 0u,
           "TLAS is " + std::to_string(as->infoHost->levelCount) + " levels deep; the traversal stack holds " +
               std::to_string(kMaxTlasLevels) + " (rebuild it after rt_set_option(ctx, \"ploc_radius\", 0))");
  return 0;
}

// refit = true: keep the topology of the last build (rt_tlas_refit). The per-frame paths — refit, the one-node TLAS of
// small scenes, the one-CTA build — enqueue kernels only: no device->host read-back, no synchronisation.
int buildTlas(rt_context *ctx, AccelObject *as, const rt_instance_descriptor *descDev, uint32_t count, bool refit) {
  cudaStream_t st = ctx->stream;
  as->isTlas = true;
  RT_TRY(checkTlasInfo(ctx, as, false));
  if (!as->headerDev) RT_CUDA(cudaMalloc(&as->headerDev, sizeof(TlasHeader)));
  if (!as->infoDev) {
    RT_CUDA(cudaMalloc(&as->infoDev, sizeof(TlasBuildInfo)));
    RT_CUDA(cudaMemsetAsync(as->infoDev, 0, sizeof(TlasBuildInfo), st));
    RT_CUDA(cudaMallocHost(&as->infoHost, sizeof(TlasBuildInfo)));
    std::memset(as->infoHost, 0, sizeof(TlasBuildInfo));
    RT_CUDA(cudaEventCreateWithFlags(&as->infoEvent, cudaEventDisableTiming));
  }
  if (count > as->primCapacity) {
    if (as->primCapacity) RT_CUDA(RT_SYNC_STREAM(ctx, st)); // the old arrays may still be in use
    if (as->instances) cudaFree(as->instances);
    if (as->instanceBox) cudaFree(as->instanceBox);
    if (as->leafPrim) cudaFree(as->leafPrim);
    if (as->nodes) cudaFree(as->nodes);
    if (as->nodeBox) cudaFree(as->nodeBox);
    if (as->nodeParent) cudaFree(as->nodeParent);
    if (as->nodePending) cudaFree(as->nodePending);
    as->instances = nullptr, as->leafPrim = nullptr, as->nodes = nullptr, as->nodeBox = nullptr;
    as->nodeParent = nullptr, as->nodePending = nullptr, as->instanceBox = nullptr;
    RT_CUDA(cudaMalloc(&as->instances, size_t(count) * sizeof(InstanceRecord)));
    RT_CUDA(cudaMalloc(&as->instanceBox, size_t(count) * 2 * sizeof(float4)));
    RT_CUDA(cudaMalloc(&as->leafPrim, size_t(count) * sizeof(uint32_t)));
    RT_CUDA(cudaMalloc(&as->nodes, size_t(count) * sizeof(WideNode)));
    RT_CUDA(cudaMalloc(&as->nodeBox, size_t(count) * 2 * sizeof(float4)));
    RT_CUDA(cudaMalloc(&as->nodeParent, size_t(count) * sizeof(uint32_t)));
    RT_CUDA(cudaMalloc(&as->nodePending, size_t(count) * sizeof(uint32_t)));
    as->primCapacity = count;
    as->nodeCapacity = count;
    as->treeValid = false;
  }
  if (count != as->primCount) as->treeValid = false;
  as->primCount = count;
  if (count == 0) {
    as->nodeCount = 0;
    as->deviceBuilt = false;
    k_write_tlas_header<<<1, 32, 0, st>>>(static_cast<TlasHeader *>(as->headerDev), nullptr, nullptr, nullptr, nullptr, 0, 0);
    ++ctx->launches;
    return 0;
  }
  RT_CHECK(descDev != nullptr, "rt_tlas: null descriptor buffer");
  size_t cubBytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr,
                                  (uint32_t *)nullptr, int(count), 0, 63, st);
  RT_TRY(ensureScratch(ctx, scratchNeed(count, cubBytes)));
  Bump bump{static_cast<uint8_t *>(ctx->scratch), 0, ctx->scratchBytes};
  float4 *primLo = bump.take<float4>(count), *primHi = bump.take<float4>(count);
  BoundsAtomics *bounds = bump.take<BoundsAtomics>(1);
  k_init_bounds<<<1, 32, 0, st>>>(bounds);
  k_instance_bounds<<<gridFor(count, 128), 128, 0, st>>>(descDev, count, as->instances, primLo, primHi, as->instanceBox, bounds);
  ctx->launches += 2;
  if (count <= 8) { // one wide node written by one thread: build and refit are the same thing
    k_tlas_small<<<1, 32, 0, st>>>(primLo, primHi, count, as->nodes, as->nodeBox, as->leafPrim);
    ++ctx->launches;
    as->levelStart = {0u, 1u};
    as->nodeCount = 1;
    as->deviceBuilt = false;
  } else if (refit && as->treeValid) {
    // nodes without internal children start, whoever delivers a node's last child refits it (as for a BLAS); the
    // grids cover the node capacity because the node count of a device-side build is only known on the device
    TlasBuildInfo *info = as->infoDev;
    if (!as->deviceBuilt) { // host-driven build: publish its node count where the kernels look for it
      const TlasBuildInfo known{as->nodeCount, uint32_t(as->levelStart.size() ? as->levelStart.size() - 1 : 0), 0u, 0u};
      *as->infoHost = known;
      RT_CUDA(cudaMemcpyAsync(info, as->infoHost, sizeof known, cudaMemcpyHostToDevice, st));
    }
    const uint32_t cap = as->deviceBuilt ? as->nodeCapacity : as->nodeCount;
    k_tlas_refit_pending<<<gridFor(cap, 256), 256, 0, st>>>(as->nodes, info, as->nodePending);
    k_tlas_refit_bottom_up<<<gridFor(cap, 128), 128, 0, st>>>(as->nodes, as->nodeBox, primLo, primHi, as->leafPrim, info,
                                                            as->nodeParent, as->nodePending);
    ctx->launches += 2;
    RT_CUDA(cudaGetLastError());
    return 0; // header unchanged: same arrays, same counts
  } else if (count <= kTlasCtaMax) {
    // sort by Morton code, then the whole tree in one CTA (k_tlas_build_cta)
    const uint32_t B = 256;
    uint64_t *keysA = bump.take<uint64_t>(count), *keysB = bump.take<uint64_t>(count);
    uint32_t *valsA = bump.take<uint32_t>(count), *valsB = bump.take<uint32_t>(count);
    k_morton<<<gridFor(count, B), B, 0, st>>>(primLo, primHi, count, bounds, keysA, valsA);
    void *cubTemp = bump.take<uint8_t>(cubBytes);
    RT_CUDA(cub::DeviceRadixSort::SortPairs(cubTemp, cubBytes, keysA, keysB, valsA, valsB, int(count), 0, 63, st));
    ctx->launches += 5;
    Bvh2 t{};
    t.n = count;
    const uint32_t ni = count - 1;
    t.left = bump.take<uint32_t>(ni);
    t.right = bump.take<uint32_t>(ni);
    t.parent = nullptr; // Karras only
    t.lo = bump.take<float4>(ni);
    t.hi = bump.take<float4>(ni);
    t.count = bump.take<uint32_t>(ni);
    t.flag = nullptr;
    PlocClusters ca{bump.take<uint32_t>(count), bump.take<float4>(count), bump.take<float4>(count)};
    PlocClusters cb{bump.take<uint32_t>(count), bump.take<float4>(count), bump.take<float4>(count)};
    uint32_t *nearest = bump.take<uint32_t>(count);
    uint32_t *queueA = bump.take<uint32_t>(count), *queueB = bump.take<uint32_t>(count);
    CollapseCounters *counters = bump.take<CollapseCounters>(1);
    RT_CHECK(bump.offset <= bump.capacity, "internal: build scratch overflow (TLAS)");
    // The TLAS is small next to the rays that walk it, so its search window is wide: on the 4,096-instance scene a window of
    // 256 instead of 16 takes 8 % off the frame (profiles/r2_experiments.md 8) and saturates there. The one-CTA build costs
    // count * 2 * radius distance evaluations per round, hence the cap by count; the per-frame path is the refit anyway.
    const int automatic = std::min(256, std::max(std::max(ctx->plocRadius, 16), int((1u << 20) / count)));
    const int radius = ctx->tlasPlocRadius > 0 ? ctx->tlasPlocRadius : automatic;
    k_tlas_build_cta<<<1, kTlasCtaThreads, 0, st>>>(count, primLo, primHi, valsB, t, ca, cb, nearest, radius, queueA, queueB,
                                                    counters, as->nodes, as->nodeBox, as->leafPrim,
                                                    uint32_t(std::min(std::max(ctx->tlasLeafSize, 1), kMaxLeafPrims)),
                                                    as->nodeParent, as->infoDev);
    ++ctx->launches;
    RT_CUDA(cudaMemcpyAsync(as->infoHost, as->infoDev, sizeof(TlasBuildInfo), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaEventRecord(as->infoEvent, st));
    as->infoPending = true;
    as->deviceBuilt = true;
    as->levelStart.clear();
    as->nodeCount = 1; // a lower bound until the info arrives; the traversal only needs "not empty"
  } else {
    // beyond the one-CTA builder: the general builder, which reads counters back between rounds and levels
    RT_TRY(buildWideTree(ctx, as, count, primLo, primHi, bounds, bump, as->leafPrim,
                         ctx->tlasPlocRadius > 0 ? ctx->tlasPlocRadius : ctx->plocRadius));
    RT_CHECK(as->levelStart.size() - 1 <= kMaxTlasLevels, "TLAS too deep for the traversal stack");
    k_node_parents<<<gridFor(as->nodeCount, 256), 256, 0, st>>>(as->nodes, as->nodeCount, as->nodeParent);
    ++ctx->launches;
    as->deviceBuilt = false;
  }
  as->treeValid = true;
  k_write_tlas_header<<<1, 32, 0, st>>>(static_cast<TlasHeader *>(as->headerDev), as->nodes, as->instances,
                                        as->leafPrim, as->instanceBox, count, as->nodeCount);
  ++ctx->launches;
  RT_CUDA(cudaGetLastError());
  as->bytes = sizeof(TlasHeader) + size_t(as->primCapacity) * (sizeof(InstanceRecord) + 32 + 4 + sizeof(WideNode) + 32 + 8);
  return 0;
}

void destroyAccel(AccelObject *as) {
  if (!as) return;
  cudaFree(as->headerDev);
  cudaFree(as->nodes);
  cudaFree(as->nodeBox);
  cudaFree(as->tris);
  cudaFree(as->instances);
  cudaFree(as->instanceBox);
  cudaFree(as->leafPrim);
  cudaFree(as->triSource);
  cudaFree(as->geomTableDev);
  cudaFree(as->nodeParent);
  cudaFree(as->nodePending);
  cudaFree(as->infoDev);
  if (as->infoHost) cudaFreeHost(as->infoHost);
  if (as->infoEvent) cudaEventDestroy(as->infoEvent);
  delete as;
}

} // namespace rtb
