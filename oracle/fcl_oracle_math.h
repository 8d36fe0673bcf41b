// oracle/fcl_oracle_math.h -- TEST INFRASTRUCTURE (CPU restatement, "port" oracle).
//
// Tiny 3-vector / 3x3 / pose types for the CPU restatement of the reference's
// hot path.  Arithmetic conventions are those of oracle/eigen_shim (which
// defines the "reference" oracle's floating point): left-to-right
// accumulation, coefficient-wise division, normalize() guarded by z > 0.
// Compiled with -ffp-contract=off and without -mfma.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace orc {

template <typename T>
struct Vec3 {
  T v[3];
  Vec3() : v{T(0), T(0), T(0)} {}
  Vec3(T x, T y, T z) : v{x, y, z} {}
  T& operator[](int i) { return v[i]; }
  T operator[](int i) const { return v[i]; }
  Vec3 operator+(const Vec3& o) const { return Vec3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  Vec3 operator-(const Vec3& o) const { return Vec3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  Vec3 operator-() const { return Vec3(-v[0], -v[1], -v[2]); }
  Vec3 operator*(T s) const { return Vec3(v[0] * s, v[1] * s, v[2] * s); }
  Vec3 operator/(T s) const { return Vec3(v[0] / s, v[1] / s, v[2] / s); }
  T dot(const Vec3& o) const { return (v[0] * o.v[0] + v[1] * o.v[1]) + v[2] * o.v[2]; }
  T squaredNorm() const { return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]; }
  T norm() const { return std::sqrt(squaredNorm()); }
  Vec3 cross(const Vec3& o) const {
    return Vec3(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
  }
  void normalize() {
    const T z = squaredNorm();
    if (z > T(0)) *this = *this / std::sqrt(z);
  }
  Vec3 normalized() const {
    Vec3 r = *this;
    r.normalize();
    return r;
  }
};
template <typename T>
Vec3<T> operator*(T s, const Vec3<T>& a) {
  return Vec3<T>(s * a.v[0], s * a.v[1], s * a.v[2]);
}

template <typename T>
struct Mat3 {
  T m[3][3];
  Vec3<T> operator*(const Vec3<T>& x) const {
    return Vec3<T>((m[0][0] * x[0] + m[0][1] * x[1]) + m[0][2] * x[2], (m[1][0] * x[0] + m[1][1] * x[1]) + m[1][2] * x[2],
                   (m[2][0] * x[0] + m[2][1] * x[1]) + m[2][2] * x[2]);
  }
  Mat3 operator*(const Mat3& b) const {
    Mat3 r;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r.m[i][j] = (m[i][0] * b.m[0][j] + m[i][1] * b.m[1][j]) + m[i][2] * b.m[2][j];
    return r;
  }
  Mat3 transpose() const {
    Mat3 r;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r.m[i][j] = m[j][i];
    return r;
  }
  Vec3<T> col(int j) const { return Vec3<T>(m[0][j], m[1][j], m[2][j]); }
};

template <typename T>
struct Xform {
  Mat3<T> R;
  Vec3<T> t;
  Vec3<T> operator*(const Vec3<T>& x) const { return R * x + t; }
  Xform operator*(const Xform& o) const {
    Xform r;
    r.R = R * o.R;
    r.t = R * o.t + t;
    return r;
  }
  Xform inverse() const {  // Isometry: (R^T, -(R^T t))
    Xform r;
    r.R = R.transpose();
    r.t = -(r.R * t);
    return r;
  }
};

template <typename T>
Xform<T> loadPose(const T* p) {
  Xform<T> x;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) x.R.m[i][j] = p[3 * i + j];
  x.t = Vec3<T>(p[9], p[10], p[11]);
  return x;
}

}  // namespace orc
