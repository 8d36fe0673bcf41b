// fclb_mpr.cuh -- per-thread Minkowski Portal Refinement, boolean intersection.
//
// Behavioural contract: include/fcl/cvx_collide/mpr.hpp RunIntersect (:34-186),
// findPortal (:188-333), updatePortal (:436-493).  This is what fcl::collide
// runs first whenever no contact is requested (gjk_solver-inl.h:88-98).
// The support directions are kept only where the algorithm reads them back.
#pragma once
#include "fclb_shapes.cuh"

namespace fclb {

enum MprStatus : int { MPR_INTERSECT = 0, MPR_SEPARATED = 1, MPR_FAILED = 2 };  // mpr.h:22

template <typename S>
FCLB_DI S absNorm(const V3<S>& v) {  // mpr.hpp:17-20
  return fabs_(v.x) + fabs_(v.y) + fabs_(v.z);
}
// mpr.hpp:11-15: normalises the direction IN PLACE, then supports along it
// One out-of-line copy per Minkowski-difference type: the seven call sites of MPR (ten with the penetration
// queries) would otherwise each inline both support mappings (a convex hill climb is ~400 SASS instructions)
// and push the traversal kernels past the instruction cache (profiles/r01_mesh_shape.summary.txt: 4.5 issue
// slots stalled on instruction fetch per issued instruction before this change).
#ifndef FCLB_MPR_SUPPORT_INLINE
#define FCLB_MPR_SUPPORT_ATTR __device__ __noinline__
#else
#define FCLB_MPR_SUPPORT_ATTR FCLB_DI
#endif
template <typename S, typename MD>
FCLB_MPR_SUPPORT_ATTR V3<S> mprSupport(const MD& shape, V3<S>& dir, uint32_t* n_support) {
  dir = normalized(dir);
  if (n_support) *n_support += 2;
  return shape.support(dir);
}
template <typename S>
FCLB_DI void swap3(V3<S>& a, V3<S>& b) {
  const V3<S> t = a;
  a = b;
  b = t;
}

template <typename S>
struct numeric_eps;
template <>
struct numeric_eps<float> {
  static FCLB_DI float value() { return 1.1920928955078125e-07f; }
};
template <>
struct numeric_eps<double> {
  static FCLB_DI double value() { return 2.220446049250313e-16; }
};

// 0 IterationLimit, 1 DetectSeparated, 2 PortalFound      (mpr.hpp:188-333)
template <typename S, typename MD>
FCLB_DI int mprFindPortal(const MD& shape, const V3<S>& v0, V3<S>& v1, V3<S>& v2, V3<S>& v3, int max_iterations,
                          uint32_t* n_support) {
  const S dot_eps_ratio = numeric_eps<S>::value();  // mpr.h:103
  const S v0_abs = absNorm(v0);
  int it = 0;
  while (true) {
    if (it >= max_iterations) return 0;
    it += 1;
    V3<S> v0v1 = v1 - v0;
    V3<S> v0v2 = v2 - v0;
    V3<S> v0v3 = v3 - v0;
    V3<S> n031 = cross(v0v3, v0v1);
    V3<S> n012 = cross(v0v1, v0v2);
    const S signed_volume = dot(v0v2, n031);
    if (signed_volume < 0) {
      swap3(v2, v3);
      swap3(v0v2, v0v3);
      swap3(n012, n031);
      n031 = n031 * S(-1);
      n012 = n012 * S(-1);
    }
    if (dot(v0, n031) > dot_eps_ratio * v0_abs * absNorm(n031)) {
      V3<S> d = n031 * S(-1);
      v2 = mprSupport(shape, d, n_support);
      if (dot(v2, d) < 0) return 1;
      continue;
    }
    if (dot(v0, n012) > dot_eps_ratio * v0_abs * absNorm(n012)) {
      V3<S> d = n012 * S(-1);
      v3 = mprSupport(shape, d, n_support);
      if (dot(v3, d) < 0) return 1;
      continue;
    }
    const V3<S> n023 = cross(v0v2, v0v3);
    if (dot(v0, n023) > dot_eps_ratio * v0_abs * absNorm(n023)) {
      V3<S> d = n023 * S(-1);
      v1 = mprSupport(shape, d, n_support);
      if (dot(v1, d) < 0) return 1;
      continue;
    }
    return 2;
  }
}

// mpr.hpp:436-493 (vertex part)
template <typename S>
FCLB_DI void mprUpdatePortal(const V3<S>& v0, const V3<S>& v4, V3<S>& v1, V3<S>& v2, V3<S>& v3) {
  const V3<S> n = cross(v4, v0);
  S d = dot(v1, n);
  if (d > 0) {
    d = dot(v2, n);
    if (d > 0)
      v1 = v4;
    else
      v3 = v4;
  } else {
    d = dot(v3, n);
    if (d > 0)
      v2 = v4;
    else
      v1 = v4;
  }
}

// mpr.hpp:34-186
template <typename S, typename MD>
FCLB_DI int mprIntersect(const MD& shape, int max_iterations, S tolerance, uint32_t* n_support) {
  const V3<S> v0 = shape.interior();
  if (sqnorm(v0) <= tolerance * tolerance) return MPR_INTERSECT;
  V3<S> d1 = -v0;
  V3<S> v1 = mprSupport(shape, d1, n_support);
  if (dot(d1, v1) < 0) return MPR_SEPARATED;
  V3<S> d2 = cross(v0, v1);
  if (absNorm(d2) <= absNorm(v0) * absNorm(v1) * tolerance) return MPR_INTERSECT;
  V3<S> v2 = mprSupport(shape, d2, n_support);
  if (dot(d2, v2) < 0) return MPR_SEPARATED;
  V3<S> d3 = cross(v1 - v0, v2 - v0);
  if (dot(d3, v0) > 0) {
    swap3(v1, v2);
    d3 = d3 * S(-1);
  }
  V3<S> v3 = mprSupport(shape, d3, n_support);
  if (dot(d3, v3) < 0) return MPR_SEPARATED;

  const int fp = mprFindPortal(shape, v0, v1, v2, v3, max_iterations, n_support);
  if (fp == 0) return MPR_FAILED;
  if (fp == 1) return MPR_SEPARATED;

  int it = 0;
  while (it < max_iterations) {
    it += 1;
    V3<S> n123 = cross(v2 - v1, v3 - v1);
    if (dot(n123, v0) > 0) {
      swap3(v2, v3);
      n123 = n123 * S(-1);
    }
    if (!(dot(v1, n123) < 0)) return MPR_INTERSECT;
    const V3<S> v4 = mprSupport(shape, n123, n_support);  // n123 is unit from here on
    if (dot(v4, n123) < 0) return MPR_SEPARATED;
    const V3<S> v1v4 = v4 - v1;
    if (fabs_(dot(v1v4, n123)) < tolerance * absNorm(n123)) return MPR_SEPARATED;
    mprUpdatePortal(v0, v4, v1, v2, v3);
  }
  return MPR_FAILED;
}

}  // namespace fclb
