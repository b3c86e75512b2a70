// oracle_bvh.h — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Stands in for the closed Metal ray-tracing stack the reference calls: MTLAccelerationStructure build
// (MetalRaytracing/Utilities.swift:100-290, Renderer.swift:464-606) and
// intersector<triangle_data, instancing>::intersect (MetalRaytracing/Raytracing.metal:301-318,665,737).
// Parity is UNPINNED at this boundary (no source, no reference tests), so this file is the specification:
//
//  * plain binned-SAH BVH2 per BLAS over all geometries' triangles, BVH2 TLAS over instance world boxes;
//  * ray -> object space with the float inverse of the 4x3 instance matrix, the inverse evaluated in double
//    by cofactors (op order below), direction NOT renormalised so t is shared between spaces;
//  * watertight ray/triangle test (Woop, Benthin, Wald 2013) with a double-precision fallback when an edge
//    function is exactly zero; two-sided; accepted iff tmin < t < tmax;
//  * closest hit = smallest t, ties broken by smallest (instance, geometry, primitive);
//  * barycentrics (u, v) weight vertices 1 and 2 (Metal convention, Raytracing.metal:61-74).
// Node-box tests are conservative (2-ulp widening), so the accepted triangle set is BVH independent.
#pragma once
#include <cstdint>
#include <vector>

#include "../include/rt_types.h"
#include "oracle_math.h"

namespace orc {

struct Hit {
  float t;
  float u, v; // weights of vertex 1 and vertex 2
  uint32_t instance, geometry, primitive;
  bool valid;
};

struct Aabb {
  float lo[3], hi[3];
  void reset() {
    for (int a = 0; a < 3; ++a) {
      lo[a] = 3.0e38f;
      hi[a] = -3.0e38f;
    }
  }
  void grow(const float p[3]) {
    for (int a = 0; a < 3; ++a) {
      lo[a] = p[a] < lo[a] ? p[a] : lo[a];
      hi[a] = p[a] > hi[a] ? p[a] : hi[a];
    }
  }
  void grow(const Aabb &b) {
    grow(b.lo);
    grow(b.hi);
  }
  float area() const {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
  }
};

struct Node {
  Aabb box;
  uint32_t left;  // internal: index of left child (right = left + 1); leaf: first primitive slot
  uint32_t count; // 0 = internal, else primitive count
};

struct Tri {
  float v0[3], v1[3], v2[3];
  uint32_t geometry, primitive;
};

struct Blas {
  std::vector<Node> nodes;
  std::vector<Tri> tris; // in leaf order
  Aabb bounds;
  void build(const rt_triangle_geometry *geoms, int count);
};

struct TlasInstance {
  float inv[12]; // inverse 4x3, column-major [col*3+row]
  const Blas *blas;
  Aabb worldBox;
};

struct Tlas {
  std::vector<Node> nodes;
  std::vector<uint32_t> order; // leaf slot -> instance index
  std::vector<TlasInstance> instances;
  void build(const rt_instance_descriptor *desc, uint32_t count);
};

// The inverse of a 4x3 affine instance matrix, defined op by op (double, no contraction).
void invertAffine4x3(const float m[4][3], float inv[12]);

struct RayPrecalc {
  int kx, ky, kz;
  float Sx, Sy, Sz;
};
RayPrecalc precalcRay(const float d[3]);
// Returns true and fills t,u,v when the ray hits the triangle with tmin < t < tmax.
bool intersectTriangle(const float o[3], const RayPrecalc &rp, const float v0[3], const float v1[3],
                       const float v2[3], float tmin, float tmax, float &t, float &u, float &v);

Hit traceClosest(const Tlas &tlas, float3 origin, float3 dir, float tmin, float tmax);
bool traceAny(const Tlas &tlas, float3 origin, float3 dir, float tmin, float tmax);

} // namespace orc
