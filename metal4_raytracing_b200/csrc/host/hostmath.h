// hostmath.h — small column-major float4x4 / float3 helpers for the host-side scene code.
// Conventions follow the reference's matrix helpers (MetalRaytracing/Utilities.swift:302-355):
// m[col*4 + row], rotate(r) = Rx(r.x) * Ry(r.y) * Rz(r.z), model transform = T * R * S (Mesh.swift:61-68).
#pragma once
#include <cmath>
#include <cstring>

namespace rth {

struct V3 {
  float x = 0, y = 0, z = 0;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) {
  float l = length(a);
  return l > 0 ? a * (1.0f / l) : a;
}

struct M4 {
  float m[16];
  float &at(int col, int row) { return m[col * 4 + row]; }
  float at(int col, int row) const { return m[col * 4 + row]; }
};

inline M4 identity() {
  M4 r{};
  r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
  return r;
}

inline M4 mul(const M4 &a, const M4 &b) {
  M4 r{};
  for (int c = 0; c < 4; ++c)
    for (int row = 0; row < 4; ++row) {
      float s = 0.0f;
      for (int k = 0; k < 4; ++k) s += a.at(k, row) * b.at(c, k);
      r.at(c, row) = s;
    }
  return r;
}

inline M4 translate(V3 t) {
  M4 r = identity();
  r.at(3, 0) = t.x;
  r.at(3, 1) = t.y;
  r.at(3, 2) = t.z;
  return r;
}

inline M4 scale(V3 s) {
  M4 r = identity();
  r.at(0, 0) = s.x;
  r.at(1, 1) = s.y;
  r.at(2, 2) = s.z;
  return r;
}

// Axis-angle rotation, column layout of Utilities.swift:313-326.
inline M4 rotateAxis(float radians, V3 axis) {
  axis = normalize(axis);
  float ct = std::cos(radians), st = std::sin(radians), ci = 1.0f - ct;
  float x = axis.x, y = axis.y, z = axis.z;
  M4 r = identity();
  r.at(0, 0) = ct + x * x * ci;
  r.at(0, 1) = y * x * ci + z * st;
  r.at(0, 2) = z * x * ci - y * st;
  r.at(1, 0) = x * y * ci - z * st;
  r.at(1, 1) = ct + y * y * ci;
  r.at(1, 2) = z * y * ci + x * st;
  r.at(2, 0) = x * z * ci + y * st;
  r.at(2, 1) = y * z * ci - x * st;
  r.at(2, 2) = ct + z * z * ci;
  return r;
}

inline M4 rotateXYZ(V3 r) {
  return mul(mul(rotateAxis(r.x, {1, 0, 0}), rotateAxis(r.y, {0, 1, 0})), rotateAxis(r.z, {0, 0, 1}));
}

inline M4 trs(V3 position, V3 rotation, float s) {
  return mul(mul(translate(position), rotateXYZ(rotation)), scale({s, s, s}));
}

// Unit quaternion (ix, iy, iz, r) -> rotation matrix (what simd's matrix_float4x4(simd_quatf) yields).
inline M4 fromQuat(float qx, float qy, float qz, float qw) {
  M4 r = identity();
  float xx = qx * qx, yy = qy * qy, zz = qz * qz;
  float xy = qx * qy, xz = qx * qz, yz = qy * qz;
  float wx = qw * qx, wy = qw * qy, wz = qw * qz;
  r.at(0, 0) = 1 - 2 * (yy + zz);
  r.at(0, 1) = 2 * (xy + wz);
  r.at(0, 2) = 2 * (xz - wy);
  r.at(1, 0) = 2 * (xy - wz);
  r.at(1, 1) = 1 - 2 * (xx + zz);
  r.at(1, 2) = 2 * (yz + wx);
  r.at(2, 0) = 2 * (xz + wy);
  r.at(2, 1) = 2 * (yz - wx);
  r.at(2, 2) = 1 - 2 * (xx + yy);
  return r;
}

// General 4x4 inverse evaluated in double (stand-in for simd_inverse, Model.swift:157, SkinningPass.swift:150).
inline M4 inverse(const M4 &a) {
  double m[16], inv[16];
  for (int i = 0; i < 16; ++i) m[i] = a.m[i];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] +
           m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] -
           m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] +
           m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] -
            m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] -
           m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] +
           m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] -
           m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] +
            m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] +
           m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] -
           m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] +
            m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] -
            m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] -
           m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] +
           m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] -
            m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] +
            m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  M4 r = identity();
  if (det == 0.0) return r;
  det = 1.0 / det;
  for (int i = 0; i < 16; ++i) r.m[i] = static_cast<float>(inv[i] * det);
  return r;
}

}  // namespace rth
