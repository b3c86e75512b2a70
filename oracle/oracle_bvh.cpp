// oracle_bvh.cpp — TEST INFRASTRUCTURE (CPU oracle). See oracle_bvh.h for what this specifies.
#include "oracle_bvh.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace orc {

// ---------------------------------------------------------------------------------------------------------
// inverse of the instance matrix (double, cofactors; every product/sum is a separately rounded double op)
// ---------------------------------------------------------------------------------------------------------
void invertAffine4x3(const float m[4][3], float inv[12]) {
  // a_rc = row r, column c of the 3x3 part; m is [column][row]
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
  // column-major [col*3+row]
  inv[0] = float(i00), inv[1] = float(i10), inv[2] = float(i20);
  inv[3] = float(i01), inv[4] = float(i11), inv[5] = float(i21);
  inv[6] = float(i02), inv[7] = float(i12), inv[8] = float(i22);
  inv[9] = float(it0), inv[10] = float(it1), inv[11] = float(it2);
}

// ---------------------------------------------------------------------------------------------------------
// watertight ray/triangle
// ---------------------------------------------------------------------------------------------------------
RayPrecalc precalcRay(const float d[3]) {
  RayPrecalc r;
  float ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]);
  r.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
  r.kx = r.kz + 1 == 3 ? 0 : r.kz + 1;
  r.ky = r.kx + 1 == 3 ? 0 : r.kx + 1;
  if (d[r.kz] < 0.0f) std::swap(r.kx, r.ky);
  r.Sx = d[r.kx] / d[r.kz];
  r.Sy = d[r.ky] / d[r.kz];
  r.Sz = 1.0f / d[r.kz];
  return r;
}

bool intersectTriangle(const float o[3], const RayPrecalc &rp, const float v0[3], const float v1[3],
                       const float v2[3], float tmin, float tmax, float &tOut, float &uOut, float &vOut) {
  const float Akx = v0[rp.kx] - o[rp.kx], Aky = v0[rp.ky] - o[rp.ky], Akz = v0[rp.kz] - o[rp.kz];
  const float Bkx = v1[rp.kx] - o[rp.kx], Bky = v1[rp.ky] - o[rp.ky], Bkz = v1[rp.kz] - o[rp.kz];
  const float Ckx = v2[rp.kx] - o[rp.kx], Cky = v2[rp.ky] - o[rp.ky], Ckz = v2[rp.kz] - o[rp.kz];
  const float Ax = Akx - rp.Sx * Akz, Ay = Aky - rp.Sy * Akz;
  const float Bx = Bkx - rp.Sx * Bkz, By = Bky - rp.Sy * Bkz;
  const float Cx = Ckx - rp.Sx * Ckz, Cy = Cky - rp.Sy * Ckz;
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
  const float Az = rp.Sz * Akz, Bz = rp.Sz * Bkz, Cz = rp.Sz * Ckz;
  const float T = (U * Az + V * Bz) + W * Cz;
  const float invDet = 1.0f / det;
  const float t = T * invDet;
  if (!(t > tmin && t < tmax)) return false;
  tOut = t;
  uOut = V * invDet;
  vOut = W * invDet;
  return true;
}

// ---------------------------------------------------------------------------------------------------------
// binned-SAH BVH2
// ---------------------------------------------------------------------------------------------------------
namespace {

struct PrimRef {
  Aabb box;
  float c[3];
  uint32_t index;
};

struct BuildTask {
  uint32_t node, first, count;
};

void buildBvh2(std::vector<PrimRef> &prims, std::vector<Node> &nodes, uint32_t maxLeaf) {
  nodes.clear();
  nodes.reserve(prims.size() * 2 + 1);
  nodes.push_back({});
  if (prims.empty()) {
    nodes[0].box.reset();
    nodes[0].left = 0;
    nodes[0].count = 0;
    return;
  }
  std::vector<BuildTask> stack;
  stack.push_back({0, 0, uint32_t(prims.size())});
  constexpr int kBins = 16;
  while (!stack.empty()) {
    BuildTask task = stack.back();
    stack.pop_back();
    Aabb box, cbox;
    box.reset();
    cbox.reset();
    for (uint32_t i = task.first; i < task.first + task.count; ++i) {
      box.grow(prims[i].box);
      cbox.grow(prims[i].c);
    }
    nodes[task.node].box = box;
    if (task.count <= 1) {
      nodes[task.node].left = task.first;
      nodes[task.node].count = task.count;
      continue;
    }
    int bestAxis = -1, bestSplit = 0;
    float bestCost = 3.0e38f;
    for (int axis = 0; axis < 3; ++axis) {
      float lo = cbox.lo[axis], ext = cbox.hi[axis] - lo;
      if (!(ext > 0.0f)) continue;
      Aabb binBox[kBins];
      uint32_t binCount[kBins] = {0};
      for (auto &b : binBox) b.reset();
      float scale = float(kBins) / ext;
      for (uint32_t i = task.first; i < task.first + task.count; ++i) {
        int b = std::min(kBins - 1, std::max(0, int((prims[i].c[axis] - lo) * scale)));
        binBox[b].grow(prims[i].box);
        ++binCount[b];
      }
      float rightArea[kBins];
      uint32_t rightCount[kBins];
      Aabb acc;
      acc.reset();
      uint32_t cnt = 0;
      for (int b = kBins - 1; b > 0; --b) {
        if (binCount[b]) acc.grow(binBox[b]);
        cnt += binCount[b];
        rightArea[b] = cnt ? acc.area() : 0.0f;
        rightCount[b] = cnt;
      }
      acc.reset();
      cnt = 0;
      for (int b = 0; b < kBins - 1; ++b) {
        if (binCount[b]) acc.grow(binBox[b]);
        cnt += binCount[b];
        if (cnt == 0 || rightCount[b + 1] == 0) continue;
        float cost = acc.area() * float(cnt) + rightArea[b + 1] * float(rightCount[b + 1]);
        if (cost < bestCost) {
          bestCost = cost;
          bestAxis = axis;
          bestSplit = b;
        }
      }
    }
    float leafCost = box.area() * float(task.count);
    uint32_t mid;
    if (bestAxis < 0 || (task.count <= maxLeaf && !(bestCost + box.area() < leafCost))) {
      if (task.count <= maxLeaf) {
        nodes[task.node].left = task.first;
        nodes[task.node].count = task.count;
        continue;
      }
      // degenerate centroids: split the range in half
      mid = task.first + task.count / 2;
    } else {
      float lo = cbox.lo[bestAxis], scale = float(kBins) / (cbox.hi[bestAxis] - lo);
      auto it = std::partition(prims.begin() + task.first, prims.begin() + task.first + task.count,
                               [&](const PrimRef &p) {
                                 int b = std::min(kBins - 1, std::max(0, int((p.c[bestAxis] - lo) * scale)));
                                 return b <= bestSplit;
                               });
      mid = uint32_t(it - prims.begin());
      if (mid == task.first || mid == task.first + task.count) mid = task.first + task.count / 2;
    }
    uint32_t left = uint32_t(nodes.size());
    nodes.push_back({});
    nodes.push_back({});
    nodes[task.node].left = left;
    nodes[task.node].count = 0;
    stack.push_back({left, task.first, mid - task.first});
    stack.push_back({left + 1, mid, task.first + task.count - mid});
  }
}

// 1/d for the slab tests only (the triangle test uses the true direction). A (near-)zero component returns 0 as
// a marker: the ray is parallel to that slab, which then constrains nothing when the origin lies inside it
// (faces included, so a ray running exactly in a box face still enters the box) and rejects the box otherwise.
inline float safeInverse(float d) { return fabsf(d) < 1.0e-20f ? 0.0f : 1.0f / d; }

inline bool slab(const Aabb &b, const float o[3], const float invd[3], float tmin, float tmax, float &tnear) {
  float tn = -3.0e38f, tf = 3.0e38f;
  for (int a = 0; a < 3; ++a) {
    if (invd[a] == 0.0f) {
      if (o[a] < b.lo[a] || o[a] > b.hi[a]) return false;
      continue;
    }
    float t0 = (b.lo[a] - o[a]) * invd[a];
    float t1 = (b.hi[a] - o[a]) * invd[a];
    tn = fmaxf(fminf(t0, t1), tn);
    tf = fminf(fmaxf(t0, t1), tf);
  }
  // conservative widening (a few ulp) so the accepted triangle set never depends on box rounding
  tn -= fabsf(tn) * 4.0e-7f;
  tf += fabsf(tf) * 4.0e-7f;
  tnear = tn;
  return tn <= tf && tf >= tmin && tn <= tmax;
}

inline bool better(float t, uint32_t inst, uint32_t geom, uint32_t prim, const Hit &h) {
  if (!h.valid) return true;
  if (t < h.t) return true;
  if (t > h.t) return false;
  if (inst != h.instance) return inst < h.instance;
  if (geom != h.geometry) return geom < h.geometry;
  return prim < h.primitive;
}

// closest: updates `best`; any: returns on first accepted triangle
template <bool kAny>
bool traverseBlas(const Blas &blas, const float o[3], const float d[3], float tmin, float tmax, uint32_t instance,
                  Hit &best) {
  if (blas.tris.empty()) return false;
  float invd[3] = {safeInverse(d[0]), safeInverse(d[1]), safeInverse(d[2])};
  RayPrecalc rp = precalcRay(d);
  uint32_t stack[128];
  int sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const Node &n = blas.nodes[stack[--sp]];
    float tn;
    float limit = kAny ? tmax : (best.valid ? best.t : tmax);
    if (!slab(n.box, o, invd, tmin, limit, tn)) continue;
    if (n.count) {
      for (uint32_t i = n.left; i < n.left + n.count; ++i) {
        const Tri &tr = blas.tris[i];
        float t, u, v;
        if (!intersectTriangle(o, rp, tr.v0, tr.v1, tr.v2, tmin, tmax, t, u, v)) continue;
        if (kAny) return true;
        if (better(t, instance, tr.geometry, tr.primitive, best)) {
          best.valid = true;
          best.t = t;
          best.u = u;
          best.v = v;
          best.instance = instance;
          best.geometry = tr.geometry;
          best.primitive = tr.primitive;
        }
      }
    } else {
      float tl, tr2;
      bool hl = slab(blas.nodes[n.left].box, o, invd, tmin, limit, tl);
      bool hr = slab(blas.nodes[n.left + 1].box, o, invd, tmin, limit, tr2);
      if (hl && hr) {
        if (tl <= tr2) {
          stack[sp++] = n.left + 1;
          stack[sp++] = n.left;
        } else {
          stack[sp++] = n.left;
          stack[sp++] = n.left + 1;
        }
      } else if (hl) {
        stack[sp++] = n.left;
      } else if (hr) {
        stack[sp++] = n.left + 1;
      }
    }
  }
  return false;
}

template <bool kAny>
bool traverseTlas(const Tlas &tlas, float3 origin, float3 dir, float tmin, float tmax, Hit &best) {
  best.valid = false;
  best.t = tmax;
  best.u = best.v = 0.0f;
  best.instance = best.geometry = best.primitive = 0;
  if (tlas.instances.empty()) return false;
  const float o[3] = {origin.x, origin.y, origin.z}, d[3] = {dir.x, dir.y, dir.z};
  float invd[3] = {safeInverse(d[0]), safeInverse(d[1]), safeInverse(d[2])};
  uint32_t stack[128];
  int sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const Node &n = tlas.nodes[stack[--sp]];
    float tn;
    float limit = kAny ? tmax : (best.valid ? best.t : tmax);
    if (!slab(n.box, o, invd, tmin, limit, tn)) continue;
    if (n.count) {
      for (uint32_t i = n.left; i < n.left + n.count; ++i) {
        uint32_t instance = tlas.order[i];
        const TlasInstance &in = tlas.instances[instance];
        if (!in.blas) continue;
        const float *m = in.inv;
        // object-space ray: o' = ((I0*o.x + I1*o.y) + I2*o.z) + It ; d' = (I0*d.x + I1*d.y) + I2*d.z
        float oo[3], dd[3];
        for (int r = 0; r < 3; ++r) {
          oo[r] = ((m[0 + r] * o[0] + m[3 + r] * o[1]) + m[6 + r] * o[2]) + m[9 + r];
          dd[r] = (m[0 + r] * d[0] + m[3 + r] * d[1]) + m[6 + r] * d[2];
        }
        if (traverseBlas<kAny>(*in.blas, oo, dd, tmin, tmax, instance, best)) return true;
      }
    } else {
      float tl, tr2;
      bool hl = slab(tlas.nodes[n.left].box, o, invd, tmin, limit, tl);
      bool hr = slab(tlas.nodes[n.left + 1].box, o, invd, tmin, limit, tr2);
      if (hl && hr) {
        if (tl <= tr2) {
          stack[sp++] = n.left + 1;
          stack[sp++] = n.left;
        } else {
          stack[sp++] = n.left;
          stack[sp++] = n.left + 1;
        }
      } else if (hl) {
        stack[sp++] = n.left;
      } else if (hr) {
        stack[sp++] = n.left + 1;
      }
    }
  }
  return false;
}

} // namespace

void Blas::build(const rt_triangle_geometry *geoms, int count) {
  std::vector<Tri> all;
  for (int g = 0; g < count; ++g) {
    const rt_triangle_geometry &ge = geoms[g];
    const uint8_t *vb = static_cast<const uint8_t *>(ge.vertexBuffer);
    for (uint32_t t = 0; t < ge.triangleCount; ++t) {
      uint32_t idx[3];
      for (int k = 0; k < 3; ++k) {
        if (ge.indexStride == 2)
          idx[k] = static_cast<const uint16_t *>(ge.indexBuffer)[3 * t + k];
        else
          idx[k] = static_cast<const uint32_t *>(ge.indexBuffer)[3 * t + k];
      }
      Tri tr;
      std::memcpy(tr.v0, vb + size_t(idx[0]) * ge.vertexStride, 12);
      std::memcpy(tr.v1, vb + size_t(idx[1]) * ge.vertexStride, 12);
      std::memcpy(tr.v2, vb + size_t(idx[2]) * ge.vertexStride, 12);
      tr.geometry = uint32_t(g);
      tr.primitive = t;
      all.push_back(tr);
    }
  }
  std::vector<PrimRef> prims(all.size());
  bounds.reset();
  for (size_t i = 0; i < all.size(); ++i) {
    prims[i].box.reset();
    prims[i].box.grow(all[i].v0);
    prims[i].box.grow(all[i].v1);
    prims[i].box.grow(all[i].v2);
    for (int a = 0; a < 3; ++a) prims[i].c[a] = 0.5f * (prims[i].box.lo[a] + prims[i].box.hi[a]);
    prims[i].index = uint32_t(i);
    bounds.grow(prims[i].box);
  }
  buildBvh2(prims, nodes, 4);
  tris.resize(all.size());
  for (size_t i = 0; i < all.size(); ++i) tris[i] = all[prims[i].index];
}

void Tlas::build(const rt_instance_descriptor *desc, uint32_t count) {
  instances.resize(count);
  std::vector<PrimRef> prims;
  for (uint32_t i = 0; i < count; ++i) {
    TlasInstance &in = instances[i];
    in.blas = reinterpret_cast<const Blas *>(static_cast<uintptr_t>(desc[i].accelerationStructureID));
    invertAffine4x3(desc[i].transformationMatrix, in.inv);
    in.worldBox.reset();
    if (!in.blas || in.blas->tris.empty()) continue;
    const float(*m)[3] = desc[i].transformationMatrix;
    for (int corner = 0; corner < 8; ++corner) {
      float p[3] = {corner & 1 ? in.blas->bounds.hi[0] : in.blas->bounds.lo[0],
                    corner & 2 ? in.blas->bounds.hi[1] : in.blas->bounds.lo[1],
                    corner & 4 ? in.blas->bounds.hi[2] : in.blas->bounds.lo[2]};
      float w[3];
      for (int r = 0; r < 3; ++r) w[r] = ((m[0][r] * p[0] + m[1][r] * p[1]) + m[2][r] * p[2]) + m[3][r];
      in.worldBox.grow(w);
    }
    // pad: object->world->object round trips lose a few ulp of the box extent
    for (int a = 0; a < 3; ++a) {
      float pad = 1.0e-5f * fmaxf(fmaxf(fabsf(in.worldBox.lo[a]), fabsf(in.worldBox.hi[a])), 1.0e-3f);
      in.worldBox.lo[a] -= pad;
      in.worldBox.hi[a] += pad;
    }
    PrimRef pr;
    pr.box = in.worldBox;
    for (int a = 0; a < 3; ++a) pr.c[a] = 0.5f * (pr.box.lo[a] + pr.box.hi[a]);
    pr.index = i;
    prims.push_back(pr);
  }
  buildBvh2(prims, nodes, 1);
  order.resize(prims.size());
  for (size_t i = 0; i < prims.size(); ++i) order[i] = prims[i].index;
}

Hit traceClosest(const Tlas &tlas, float3 origin, float3 dir, float tmin, float tmax) {
  Hit h;
  traverseTlas<false>(tlas, origin, dir, tmin, tmax, h);
  return h;
}

bool traceAny(const Tlas &tlas, float3 origin, float3 dir, float tmin, float tmax) {
  Hit h;
  return traverseTlas<true>(tlas, origin, dir, tmin, tmax, h);
}

} // namespace orc
