// procedural.cpp — deterministic stand-ins for the assets the reference mount lacks (bunny.obj, dragon.obj,
// robot.usdz, the coatball maps; /root/reference/.MISSING_LARGE_BLOBS) plus small analytic meshes for tests.
// Everything is a pure function of its integer parameters, so the oracle and the GPU path see the same bytes
// on every machine. Triangle counts follow SURVEY.md §8(d): icosphere subdiv 6 -> 81,920 tris,
// torus knot 1320 x 330 -> 871,200 tris, humanoid -> 100,000 vertices / 64 joints.
#include <algorithm>
#include <array>
#include <cmath>
#include <map>

#include "scene.h"

namespace rts {

uint32_t hash32(uint32_t x, uint32_t seed) {
  uint32_t h = x * 0x9E3779B1u + seed * 0x85EBCA77u + 0x165667B1u;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  h *= 0x297A2D39u;
  h ^= h >> 15;
  return h;
}

static float lattice(int x, int y, int z, uint32_t seed) {
  uint32_t h = hash32(uint32_t(x) * 73856093u ^ uint32_t(y) * 19349663u ^ uint32_t(z) * 83492791u, seed);
  return float(h >> 8) * (1.0f / 16777216.0f);
}

float valueNoise3(float x, float y, float z, uint32_t seed) {
  float fx = std::floor(x), fy = std::floor(y), fz = std::floor(z);
  int ix = int(fx), iy = int(fy), iz = int(fz);
  float tx = x - fx, ty = y - fy, tz = z - fz;
  tx = tx * tx * (3 - 2 * tx);
  ty = ty * ty * (3 - 2 * ty);
  tz = tz * tz * (3 - 2 * tz);
  float c[2][2][2];
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      for (int d = 0; d < 2; ++d) c[a][b][d] = lattice(ix + a, iy + b, iz + d, seed);
  auto lerp = [](float a, float b, float t) { return a + (b - a) * t; };
  float x00 = lerp(c[0][0][0], c[1][0][0], tx), x10 = lerp(c[0][1][0], c[1][1][0], tx);
  float x01 = lerp(c[0][0][1], c[1][0][1], tx), x11 = lerp(c[0][1][1], c[1][1][1], tx);
  return lerp(lerp(x00, x10, ty), lerp(x01, x11, ty), tz);
}

float fbm3(float x, float y, float z, int octaves, uint32_t seed) {
  float a = 0.5f, s = 0.0f, f = 1.0f;
  for (int o = 0; o < octaves; ++o) {
    s += a * (valueNoise3(x * f, y * f, z * f, seed + o) - 0.5f);
    a *= 0.5f;
    f *= 2.03f;
  }
  return s; // roughly [-0.5, 0.5]
}

void computeSmoothNormals(Mesh &m) {
  std::vector<double> acc(m.positions.size() * 3, 0.0);
  for (auto &sm : m.submeshes)
    for (size_t t = 0; t + 2 < sm.indices.size(); t += 3) {
      int i0 = sm.indices[t], i1 = sm.indices[t + 1], i2 = sm.indices[t + 2];
      rth::V3 a{m.positions[i0].x, m.positions[i0].y, m.positions[i0].z};
      rth::V3 b{m.positions[i1].x, m.positions[i1].y, m.positions[i1].z};
      rth::V3 c{m.positions[i2].x, m.positions[i2].y, m.positions[i2].z};
      rth::V3 n = rth::cross(b - a, c - a); // area-weighted
      for (int i : {i0, i1, i2}) {
        acc[3 * i] += n.x;
        acc[3 * i + 1] += n.y;
        acc[3 * i + 2] += n.z;
      }
    }
  m.normals.resize(m.positions.size());
  for (size_t i = 0; i < m.positions.size(); ++i) {
    double l = std::sqrt(acc[3 * i] * acc[3 * i] + acc[3 * i + 1] * acc[3 * i + 1] + acc[3 * i + 2] * acc[3 * i + 2]);
    if (l > 0)
      m.normals[i] = {float(acc[3 * i] / l), float(acc[3 * i + 1] / l), float(acc[3 * i + 2] / l), 0.0f};
    else
      m.normals[i] = {0, 1, 0, 0};
  }
}

static rt_material plainMaterial(float r, float g, float b) {
  rt_material m{};
  m.baseColor = {r, g, b, 0};
  m.refractionIndex = 1.0f;
  m.opacity = 1.0f;
  return m;
}

static Submesh &singleSubmesh(Scene &s, Mesh &m, const char *name, rt_material mat) {
  m.submeshes.emplace_back();
  Submesh &sm = m.submeshes.back();
  sm.name = name;
  s.initSubmeshDefaults(sm);
  sm.material = mat;
  return sm;
}

// Same data as AssetResources/plane.obj after loading (quad -> fan, uv + up normal).
int addPlane(Scene &s) {
  Mesh m;
  m.name = "plane(procedural)";
  const float P[4][3] = {{-1, 0, 1}, {1, 0, 1}, {1, 0, -1}, {-1, 0, -1}};
  const float T[4][2] = {{0.0001f, 0.0001f}, {0.9999f, 0.0001f}, {0.9999f, 0.9999f}, {0.0001f, 0.9999f}};
  for (int i = 0; i < 4; ++i) {
    m.positions.push_back({P[i][0], P[i][1], P[i][2], 0});
    m.normals.push_back({0, 1, 0, 0});
    m.uvs.insert(m.uvs.end(), {T[i][0], T[i][1]});
  }
  Submesh &sm = singleSubmesh(s, m, "None", plainMaterial(0.5f, 0.5f, 0.5f));
  sm.indices = {0, 1, 2, 0, 2, 3};
  s.meshes.push_back(std::move(m));
  return int(s.meshes.size()) - 1;
}

int addUvSphere(Scene &s, int rings, int sectors) {
  rings = std::max(rings, 2);
  sectors = std::max(sectors, 3);
  Mesh m;
  m.name = "uvsphere";
  const float pi = 3.14159265358979323846f;
  for (int r = 0; r <= rings; ++r)
    for (int c = 0; c <= sectors; ++c) {
      float th = pi * float(r) / float(rings), ph = 2 * pi * float(c) / float(sectors);
      float x = std::sin(th) * std::cos(ph), y = std::cos(th), z = std::sin(th) * std::sin(ph);
      m.positions.push_back({x, y, z, 0});
      m.normals.push_back({x, y, z, 0});
      m.uvs.insert(m.uvs.end(), {float(c) / float(sectors), float(r) / float(rings)});
    }
  Submesh &sm = singleSubmesh(s, m, "sphere", plainMaterial(1.0f, 1.0f, 0.5f));
  for (int r = 0; r < rings; ++r)
    for (int c = 0; c < sectors; ++c) {
      int a = r * (sectors + 1) + c, b = a + 1, d = a + sectors + 1, e = d + 1;
      if (r != 0) sm.indices.insert(sm.indices.end(), {a, b, d});
      if (r != rings - 1) sm.indices.insert(sm.indices.end(), {b, e, d});
    }
  s.meshes.push_back(std::move(m));
  return int(s.meshes.size()) - 1;
}

// Bunny stand-in: icosahedron subdivided `subdiv` times, radially displaced by fBm, smooth normals,
// spherical uvs. subdiv 6 -> 40,962 vertices / 81,920 triangles.
int addBumpyIcosphere(Scene &s, int subdiv, int seed) {
  Mesh m;
  m.name = "icosphere_bumpy";
  const double t = (1.0 + std::sqrt(5.0)) / 2.0;
  std::vector<std::array<double, 3>> V = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                                         {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
  for (auto &v : V) {
    double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (double &c : v) c /= l;
  }
  std::vector<std::array<int, 3>> F = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                                      {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                                      {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
  for (int it = 0; it < subdiv; ++it) {
    std::map<std::pair<int, int>, int> mid;
    auto midpoint = [&](int a, int b) {
      auto key = std::make_pair(std::min(a, b), std::max(a, b));
      auto f = mid.find(key);
      if (f != mid.end()) return f->second;
      std::array<double, 3> p = {(V[a][0] + V[b][0]) / 2, (V[a][1] + V[b][1]) / 2, (V[a][2] + V[b][2]) / 2};
      double l = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
      for (double &c : p) c /= l;
      V.push_back(p);
      return mid[key] = int(V.size()) - 1;
    };
    std::vector<std::array<int, 3>> G;
    G.reserve(F.size() * 4);
    for (auto &f : F) {
      int a = midpoint(f[0], f[1]), b = midpoint(f[1], f[2]), c = midpoint(f[2], f[0]);
      G.push_back({f[0], a, c});
      G.push_back({f[1], b, a});
      G.push_back({f[2], c, b});
      G.push_back({a, b, c});
    }
    F.swap(G);
  }
  const float pi = 3.14159265358979323846f;
  for (auto &v : V) {
    float x = float(v[0]), y = float(v[1]), z = float(v[2]);
    float d = 1.0f + 0.35f * fbm3(x * 2.5f + 7.0f, y * 2.5f + 3.0f, z * 2.5f + 11.0f, 5, uint32_t(seed));
    // a coarse "ears + body" lobe so the silhouette is not a plain ball
    d *= 1.0f + 0.25f * std::max(0.0f, y) * std::fabs(x);
    m.positions.push_back({x * d, y * d, z * d, 0});
    m.uvs.insert(m.uvs.end(), {0.5f + std::atan2(z, x) / (2 * pi), std::acos(std::max(-1.0f, std::min(1.0f, y))) / pi});
  }
  Submesh &sm = singleSubmesh(s, m, "bunny", plainMaterial(0.8f, 0.75f, 0.7f));
  for (auto &f : F) sm.indices.insert(sm.indices.end(), {f[0], f[1], f[2]});
  computeSmoothNormals(m);
  s.meshes.push_back(std::move(m));
  return int(s.meshes.size()) - 1;
}

// Dragon stand-in: (2,3) torus knot swept tube with fBm displacement on a nu x nv wrapped grid.
// 1320 x 330 -> 435,600 vertices / 871,200 triangles (Stanford dragon: 871,414). Fits a ~1 unit box,
// lowest point at y = -0.3 so AppScene's placement (y 0.38, scale 1.2) rests it just above the plane.
int addTorusKnot(Scene &s, int nu, int nv, int seed) {
  nu = std::max(nu, 3);
  nv = std::max(nv, 3);
  Mesh m;
  m.name = "torusknot";
  const double pi = 3.14159265358979323846;
  const double sc = 1.0 / 7.0, tube = 0.62 * sc;
  auto curve = [&](double t, double out[3]) {
    double r = 2.0 + std::cos(3.0 * t);
    out[0] = r * std::cos(2.0 * t) * sc;
    out[1] = r * std::sin(2.0 * t) * sc;
    out[2] = -std::sin(3.0 * t) * sc;
  };
  m.positions.resize(size_t(nu) * nv);
  double minY = 1e30;
  for (int i = 0; i < nu; ++i) {
    double t = 2.0 * pi * double(i) / double(nu);
    double c[3], c1[3], c0[3];
    curve(t, c);
    const double h = 1e-4;
    curve(t + h, c1);
    curve(t - h, c0);
    double T[3] = {c1[0] - c0[0], c1[1] - c0[1], c1[2] - c0[2]};
    double A[3] = {c1[0] - 2 * c[0] + c0[0], c1[1] - 2 * c[1] + c0[1], c1[2] - 2 * c[2] + c0[2]};
    double tl = std::sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]);
    for (double &x : T) x /= tl;
    double B[3] = {T[1] * A[2] - T[2] * A[1], T[2] * A[0] - T[0] * A[2], T[0] * A[1] - T[1] * A[0]};
    double bl = std::sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
    for (double &x : B) x /= bl;
    double Nn[3] = {B[1] * T[2] - B[2] * T[1], B[2] * T[0] - B[0] * T[2], B[0] * T[1] - B[1] * T[0]};
    for (int j = 0; j < nv; ++j) {
      double a = 2.0 * pi * double(j) / double(nv);
      double dir[3] = {std::cos(a) * Nn[0] + std::sin(a) * B[0], std::cos(a) * Nn[1] + std::sin(a) * B[1],
                       std::cos(a) * Nn[2] + std::sin(a) * B[2]};
      double px = c[0] + tube * dir[0], py = c[1] + tube * dir[1], pz = c[2] + tube * dir[2];
      float n = fbm3(float(px * 22.0), float(py * 22.0), float(pz * 22.0), 5, uint32_t(seed));
      float ridge = 0.10f * float(std::sin(16.0 * t) * std::cos(3.0 * a)); // scale-like ridges
      double r = tube * (1.0 + 0.45 * n + ridge);
      double x = c[0] + r * dir[0], y = c[1] + r * dir[1], z = c[2] + r * dir[2];
      m.positions[size_t(i) * nv + j] = {float(x), float(y), float(z), 0};
      minY = std::min(minY, y);
      m.uvs.insert(m.uvs.end(), {float(i) / float(nu), float(j) / float(nv)});
    }
  }
  float shift = float(-0.3 - minY);
  for (auto &p : m.positions) p.y += shift;
  rt_material red = plainMaterial(1.0f, 0.0f, 0.0f); // AssetResources/dragon.mtl: Kd 1 0 0
  red.specular = {0.2f, 0.2f, 0.2f, 0};
  Submesh &sm = singleSubmesh(s, m, "Dragon", red);
  sm.indices.reserve(size_t(nu) * nv * 6);
  for (int i = 0; i < nu; ++i)
    for (int j = 0; j < nv; ++j) {
      int i1 = (i + 1) % nu, j1 = (j + 1) % nv;
      int a = i * nv + j, b = i1 * nv + j, c = i1 * nv + j1, d = i * nv + j1;
      // counter-clockwise seen from outside the tube, so the smooth normals below point outward
      sm.indices.insert(sm.indices.end(), {a, c, b, a, d, c});
    }
  computeSmoothNormals(m);
  s.meshes.push_back(std::move(m));
  return int(s.meshes.size()) - 1;
}

// Skinned robot stand-in: limb tubes around a 64-joint skeleton (root, spine, head, two arms, two legs, tail),
// 4 influences per vertex with weights that sum to 1, a looping 2 s clip of per-joint sinusoidal rotations.
// Units are centimetres (AppScene places the robot at scale 0.01). Exactly `vertexBudget` vertices.
int addHumanoid(Scene &s, int vertexBudget, int joints) {
  joints = std::max(joints, 8);
  vertexBudget = std::max(vertexBudget, 2000);
  Mesh m;
  m.name = "humanoid";
  Skeleton &sk = m.skeleton;
  struct Chain {
    int firstJoint, count; // consecutive joints
    float radius;
  };
  std::vector<Chain> chains;
  std::vector<rth::V3> globalPos;
  auto addJoint = [&](int parent, rth::V3 offset, rth::V3 axis, float amp, float freq, float phase) {
    sk.parent.push_back(parent);
    sk.restOffset.push_back(offset);
    sk.axis.push_back(axis);
    sk.amplitude.push_back(amp);
    sk.freq.push_back(freq);
    sk.phase.push_back(phase);
    rth::V3 gp = parent >= 0 ? globalPos[parent] + offset : offset;
    globalPos.push_back(gp);
    return int(sk.parent.size()) - 1;
  };
  // distribute joints: root 1, spine 7, head 4, arms 2x, legs 2x, tail the rest
  int perLimb = std::max(2, (joints - 12) / 5);
  int tailCount = std::max(2, joints - 12 - 4 * perLimb);
  auto addChain = [&](int parent, int count, rth::V3 start, rth::V3 step, rth::V3 axis, float amp, float freq,
                      float phase0, float radius) {
    int first = -1, prev = parent;
    for (int k = 0; k < count; ++k) {
      rth::V3 off = k == 0 ? start : step;
      int j = addJoint(prev, off, axis, amp * (k == 0 ? 1.0f : 1.5f / float(count)), freq, phase0 + 0.4f * float(k));
      if (k == 0) first = j;
      prev = j;
    }
    chains.push_back({first, count, radius});
    return first;
  };
  int root = addJoint(-1, {0, 95, 0}, {0, 1, 0}, 0.15f, 0.5f, 0.0f);
  int spine = addChain(root, 7, {0, 6, 0}, {0, 8, 0}, {0, 0, 1}, 0.08f, 0.5f, 0.3f, 14.0f);
  int spineTop = spine + 6;
  addChain(spineTop, 4, {0, 8, 0}, {0, 6, 0}, {0, 1, 0}, 0.25f, 1.0f, 1.1f, 9.0f);                       // head
  addChain(spineTop, perLimb, {-16, 0, 0}, {-62.0f / perLimb, 0, 0}, {0, 0, 1}, 0.55f, 1.0f, 0.0f, 5.0f); // left arm
  addChain(spineTop, perLimb, {16, 0, 0}, {62.0f / perLimb, 0, 0}, {0, 0, 1}, 0.55f, 1.0f, 3.14159f, 5.0f);
  addChain(root, perLimb, {-9, -4, 0}, {0, -91.0f / float(perLimb - 1), 0}, {1, 0, 0}, 0.45f, 1.0f, 0.0f, 7.5f);     // left leg
  addChain(root, perLimb, {9, -4, 0}, {0, -91.0f / float(perLimb - 1), 0}, {1, 0, 0}, 0.45f, 1.0f, 3.14159f, 7.5f);
  addChain(root, tailCount, {0, -2, -10}, {0, -1.5f, -70.0f / tailCount}, {0, 1, 0}, 0.35f, 1.5f, 0.7f, 4.0f);
  sk.duration = 2.0;
  int J = int(sk.parent.size());
  sk.rest.resize(J);
  sk.inverseBind.resize(J);
  for (int j = 0; j < J; ++j) {
    sk.rest[j] = rth::translate(sk.restOffset[j]);
    sk.inverseBind[j] = rth::inverse(rth::translate(globalPos[j]));
  }
  // tubes: `around` vertices per ring, rings distributed over chains by length until the budget is met
  const int around = 100;
  int totalRings = vertexBudget / around;
  int leftover = vertexBudget - totalRings * around; // extra vertices appended to the last ring set
  std::vector<float> chainLen(chains.size());
  float sumLen = 0;
  for (size_t c = 0; c < chains.size(); ++c) {
    float L = 0;
    for (int k = 1; k < chains[c].count; ++k) L += rth::length(globalPos[chains[c].firstJoint + k] - globalPos[chains[c].firstJoint + k - 1]);
    chainLen[c] = L;
    sumLen += L;
  }
  std::vector<int> ringsPer(chains.size());
  int assigned = 0;
  for (size_t c = 0; c < chains.size(); ++c) {
    ringsPer[c] = std::max(2, int(std::floor(totalRings * chainLen[c] / sumLen)));
    assigned += ringsPer[c];
  }
  ringsPer[0] += totalRings - assigned;
  Submesh &sm = singleSubmesh(s, m, "robot", plainMaterial(0.55f, 0.6f, 0.7f));
  const float pi = 3.14159265358979323846f;
  for (size_t c = 0; c < chains.size(); ++c) {
    const Chain &ch = chains[c];
    int R = ringsPer[c];
    int base = int(m.positions.size());
    for (int r = 0; r < R; ++r) {
      float u = float(r) / float(R - 1) * float(ch.count - 1); // position along the joint chain
      int k0 = std::min(int(u), ch.count - 2);
      float f = u - float(k0);
      rth::V3 a = globalPos[ch.firstJoint + k0], b = globalPos[ch.firstJoint + k0 + 1];
      rth::V3 center = a + (b - a) * f;
      rth::V3 dir = rth::normalize(b - a);
      rth::V3 ref = std::fabs(dir.y) < 0.9f ? rth::V3{0, 1, 0} : rth::V3{1, 0, 0};
      rth::V3 e1 = rth::normalize(rth::cross(dir, ref)), e2 = rth::cross(dir, e1);
      float taper = 1.0f - 0.35f * (u / float(ch.count - 1));
      float bulge = 1.0f + 0.12f * std::sin(u * pi); // soft joint bulges
      float rad = ch.radius * taper * bulge;
      // 4 influences: joints k0-1 .. k0+2 (clamped), smooth tent weights, normalised to sum 1
      int jn[4];
      float w[4];
      float wsum = 0;
      for (int q = 0; q < 4; ++q) {
        int kk = std::min(std::max(k0 - 1 + q, 0), ch.count - 1);
        jn[q] = ch.firstJoint + kk;
        float d = std::fabs(u - float(k0 - 1 + q));
        w[q] = std::max(0.0f, 1.0f - d / 1.5f);
        w[q] *= w[q];
        wsum += w[q];
      }
      for (int q = 0; q < 4; ++q) w[q] /= wsum;
      for (int a2 = 0; a2 < around; ++a2) {
        float ang = 2 * pi * float(a2) / float(around);
        rth::V3 p = center + e1 * (rad * std::cos(ang)) + e2 * (rad * std::sin(ang));
        m.positions.push_back({p.x, p.y, p.z, 0});
        m.uvs.insert(m.uvs.end(), {float(a2) / float(around), float(r) / float(R - 1)});
        for (int q = 0; q < 4; ++q) {
          m.jointIndices.push_back(uint16_t(jn[q]));
          m.jointWeights.push_back(w[q]);
        }
      }
    }
    for (int r = 0; r + 1 < R; ++r)
      for (int a2 = 0; a2 < around; ++a2) {
        int a3 = (a2 + 1) % around;
        int p0 = base + r * around + a2, p1 = base + r * around + a3, p2 = base + (r + 1) * around + a3,
            p3 = base + (r + 1) * around + a2;
        sm.indices.insert(sm.indices.end(), {p0, p1, p2, p0, p2, p3});
      }
  }
  // pad to the exact vertex budget with copies of vertex 0 (unreferenced by any triangle)
  for (int k = 0; k < leftover; ++k) {
    m.positions.push_back(m.positions[0]);
    m.uvs.insert(m.uvs.end(), {0.0f, 0.0f});
    for (int q = 0; q < 4; ++q) {
      m.jointIndices.push_back(m.jointIndices[q]);
      m.jointWeights.push_back(m.jointWeights[q]);
    }
  }
  computeSmoothNormals(m);
  m.jointMatrices.assign(size_t(J) * 16, 0.0f);
  for (int j = 0; j < J; ++j) {
    rth::M4 id = rth::identity();
    std::copy(id.m, id.m + 16, &m.jointMatrices[size_t(j) * 16]);
  }
  s.meshes.push_back(std::move(m));
  return int(s.meshes.size()) - 1;
}

Texture makeProceduralTexture(const std::string &kind, int w, int h, int seed, bool srgb) {
  Texture t;
  t.width = std::max(w, 1);
  t.height = std::max(h, 1);
  t.srgb = srgb;
  t.rgba.resize(size_t(t.width) * t.height * 4);
  auto noise = [&](float u, float v) {
    // tileable-ish value noise on a 16-cell grid, 4 octaves
    return 0.5f + fbm3(u * 16.0f, v * 16.0f, 0.5f, 4, uint32_t(seed));
  };
  for (int y = 0; y < t.height; ++y)
    for (int x = 0; x < t.width; ++x) {
      float u = (float(x) + 0.5f) / float(t.width), v = (float(y) + 0.5f) / float(t.height);
      uint8_t *p = &t.rgba[(size_t(y) * t.width + x) * 4];
      auto q = [](float f) { return uint8_t(std::max(0.0f, std::min(255.0f, std::floor(f * 255.0f + 0.5f)))); };
      if (kind == "checker") {
        bool on = ((x * 8 / t.width) + (y * 8 / t.height)) & 1;
        p[0] = on ? 230 : 40;
        p[1] = on ? 230 : 40;
        p[2] = on ? 230 : 60;
        p[3] = 255;
      } else if (kind == "uvgrid") {
        bool line = (x % std::max(1, t.width / 16) == 0) || (y % std::max(1, t.height / 16) == 0);
        p[0] = line ? 255 : q(u);
        p[1] = line ? 255 : q(v);
        p[2] = line ? 255 : 64;
        p[3] = 255;
      } else if (kind == "bump") {
        float e = 1.0f / float(t.width);
        float hx = noise(u + e, v) - noise(u - e, v), hy = noise(u, v + e) - noise(u, v - e);
        rth::V3 n = rth::normalize({-hx * 6.0f, -hy * 6.0f, 1.0f});
        p[0] = q(n.x * 0.5f + 0.5f);
        p[1] = q(n.y * 0.5f + 0.5f);
        p[2] = q(n.z * 0.5f + 0.5f);
        p[3] = 255;
      } else { // "valuenoise"
        uint8_t g = q(noise(u, v));
        p[0] = p[1] = p[2] = g;
        p[3] = 255;
      }
    }
  return t;
}

} // namespace rts
