// fclb_math.cuh -- 3-vector / 3x3 / rigid-pose arithmetic for the device kernels.
//
// Parity note.  Every boolean the narrowphase produces is a strict comparison of
// sums of products, so the ORDER of floating-point operations is part of the
// contract with the reference (mind-fcl evaluates everything through Eigen
// fixed-size expressions in scalar type S).  The helpers below pin that order:
//   dot / squaredNorm / matrix products accumulate left to right
//       ((a0*b0 + a1*b1) + a2*b2)
//   cross(a,b) = (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)
//   v / s divides every coefficient (no reciprocal-multiply)
//   normalize divides by sqrt(squaredNorm) when squaredNorm > 0
// and the translation unit is compiled with --fmad=false and IEEE div/sqrt, so
// no product-sum is contracted into an FMA.  (Reference build: no -mfma,
// CMakeLists.txt:76; math/math_simd_details.h:63-65.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FCLB_DI __device__ __forceinline__

namespace fclb {

template <typename S>
struct V3 {
  S x, y, z;
};

template <typename S>
FCLB_DI V3<S> mk(S x, S y, S z) {
  V3<S> r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
template <typename S>
FCLB_DI V3<S> zero3() {
  return mk<S>(S(0), S(0), S(0));
}
template <typename S>
FCLB_DI V3<S> operator+(const V3<S>& a, const V3<S>& b) {
  return mk<S>(a.x + b.x, a.y + b.y, a.z + b.z);
}
template <typename S>
FCLB_DI V3<S> operator-(const V3<S>& a, const V3<S>& b) {
  return mk<S>(a.x - b.x, a.y - b.y, a.z - b.z);
}
template <typename S>
FCLB_DI V3<S> operator-(const V3<S>& a) {
  return mk<S>(-a.x, -a.y, -a.z);
}
template <typename S>
FCLB_DI V3<S> operator*(const V3<S>& a, S s) {
  return mk<S>(a.x * s, a.y * s, a.z * s);
}
template <typename S>
FCLB_DI V3<S> operator*(S s, const V3<S>& a) {
  return mk<S>(s * a.x, s * a.y, s * a.z);
}
// IEEE division expands to ~10 SASS instructions per quotient plus a slow-path
// call; the GJK kernels divide vectors at a dozen sites, so the three quotients
// live in ONE out-of-line copy (instruction-cache footprint).
template <typename S>
__device__ __noinline__ V3<S> div3(V3<S> a, S s) {
  return mk<S>(a.x / s, a.y / s, a.z / s);
}
template <typename S>
FCLB_DI V3<S> operator/(const V3<S>& a, S s) {
  return div3<S>(a, s);
}
template <typename S>
FCLB_DI S dot(const V3<S>& a, const V3<S>& b) {
  return (a.x * b.x + a.y * b.y) + a.z * b.z;
}
template <typename S>
FCLB_DI S sqnorm(const V3<S>& a) {
  return (a.x * a.x + a.y * a.y) + a.z * a.z;
}
__device__ __noinline__ inline float fsqrt(float v) { return sqrtf(v); }
__device__ __noinline__ inline double fsqrt(double v) { return sqrt(v); }
FCLB_DI float fabs_(float v) { return fabsf(v); }
FCLB_DI double fabs_(double v) { return fabs(v); }
template <typename S>
FCLB_DI S fmin_(S a, S b) {  // std::min semantics: (b < a) ? b : a
  return (b < a) ? b : a;
}
template <typename S>
FCLB_DI S fmax_(S a, S b) {  // std::max semantics: (a < b) ? b : a
  return (a < b) ? b : a;
}
template <typename S>
FCLB_DI S norm(const V3<S>& a) {
  return fsqrt(sqnorm(a));
}
template <typename S>
FCLB_DI V3<S> cross(const V3<S>& a, const V3<S>& b) {
  return mk<S>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Eigen normalize(): z = squaredNorm; if (z > 0) v /= sqrt(z)
template <typename S>
FCLB_DI V3<S> normalized(const V3<S>& a) {
  const S z = sqnorm(a);
  if (z > S(0)) return a / fsqrt(z);
  return a;
}
template <typename S>
FCLB_DI S comp(const V3<S>& a, int i) {
  return i == 0 ? a.x : (i == 1 ? a.y : a.z);
}
template <typename S>
FCLB_DI void setComp(V3<S>& a, int i, S v) {
  if (i == 0)
    a.x = v;
  else if (i == 1)
    a.y = v;
  else
    a.z = v;
}

// Row-major 3x3.
template <typename S>
struct M3 {
  S m[9];
  FCLB_DI S operator()(int i, int j) const { return m[3 * i + j]; }
  FCLB_DI S& operator()(int i, int j) { return m[3 * i + j]; }
};
template <typename S>
FCLB_DI V3<S> mulMV(const M3<S>& a, const V3<S>& v) {
  return mk<S>((a.m[0] * v.x + a.m[1] * v.y) + a.m[2] * v.z, (a.m[3] * v.x + a.m[4] * v.y) + a.m[5] * v.z,
               (a.m[6] * v.x + a.m[7] * v.y) + a.m[8] * v.z);
}
// a^T * v
template <typename S>
FCLB_DI V3<S> mulMtV(const M3<S>& a, const V3<S>& v) {
  return mk<S>((a.m[0] * v.x + a.m[3] * v.y) + a.m[6] * v.z, (a.m[1] * v.x + a.m[4] * v.y) + a.m[7] * v.z,
               (a.m[2] * v.x + a.m[5] * v.y) + a.m[8] * v.z);
}
template <typename S>
FCLB_DI M3<S> mulMM(const M3<S>& a, const M3<S>& b) {
  M3<S> r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = (a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j]) + a.m[3 * i + 2] * b.m[6 + j];
  return r;
}
// a^T * b
template <typename S>
FCLB_DI M3<S> mulMtM(const M3<S>& a, const M3<S>& b) {
  M3<S> r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = (a.m[i] * b.m[j] + a.m[3 + i] * b.m[3 + j]) + a.m[6 + i] * b.m[6 + j];
  return r;
}
template <typename S>
FCLB_DI M3<S> transpose(const M3<S>& a) {
  M3<S> r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * j + i];
  return r;
}
template <typename S>
FCLB_DI V3<S> col(const M3<S>& a, int j) {
  return mk<S>(a.m[j], a.m[3 + j], a.m[6 + j]);
}
template <typename S>
FCLB_DI V3<S> row(const M3<S>& a, int i) {
  return mk<S>(a.m[3 * i], a.m[3 * i + 1], a.m[3 * i + 2]);
}

// Rigid pose: x -> R x + t  (Eigen Transform<S,3,Isometry>)
template <typename S>
struct Pose {
  M3<S> R;
  V3<S> t;
};
template <typename S>
FCLB_DI V3<S> apply(const Pose<S>& p, const V3<S>& v) {
  return mulMV(p.R, v) + p.t;
}
// Transform::inverse(Isometry): (R^T, -(R^T t))
template <typename S>
FCLB_DI Pose<S> inverse(const Pose<S>& p) {
  Pose<S> r;
  r.R = transpose(p.R);
  r.t = -mulMV(r.R, p.t);
  return r;
}
// Transform * Transform: (Ra Rb, Ra tb + ta)
template <typename S>
FCLB_DI Pose<S> compose(const Pose<S>& a, const Pose<S>& b) {
  Pose<S> r;
  r.R = mulMM(a.R, b.R);
  r.t = mulMV(a.R, b.t) + a.t;
  return r;
}

// ---- pose I/O: 12 S per pose (R row-major, then t), 16-byte vector loads ----
FCLB_DI Pose<float> loadPose(const float* __restrict__ base, size_t q) {
  const float4* p = reinterpret_cast<const float4*>(base + 12 * q);
  const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  Pose<float> r;
  r.R.m[0] = a.x; r.R.m[1] = a.y; r.R.m[2] = a.z; r.R.m[3] = a.w;
  r.R.m[4] = b.x; r.R.m[5] = b.y; r.R.m[6] = b.z; r.R.m[7] = b.w;
  r.R.m[8] = c.x; r.t.x = c.y; r.t.y = c.z; r.t.z = c.w;
  return r;
}
FCLB_DI Pose<double> loadPose(const double* __restrict__ base, size_t q) {
  const double2* p = reinterpret_cast<const double2*>(base + 12 * q);
  Pose<double> r;
  double2 v;
  v = __ldg(p);     r.R.m[0] = v.x; r.R.m[1] = v.y;
  v = __ldg(p + 1); r.R.m[2] = v.x; r.R.m[3] = v.y;
  v = __ldg(p + 2); r.R.m[4] = v.x; r.R.m[5] = v.y;
  v = __ldg(p + 3); r.R.m[6] = v.x; r.R.m[7] = v.y;
  v = __ldg(p + 4); r.R.m[8] = v.x; r.t.x = v.y;
  v = __ldg(p + 5); r.t.y = v.x; r.t.z = v.y;
  return r;
}

template <typename S>
FCLB_DI void store3(S* __restrict__ base, size_t q, const V3<S>& v) {
  base[3 * q] = v.x;
  base[3 * q + 1] = v.y;
  base[3 * q + 2] = v.z;
}

}  // namespace fclb
