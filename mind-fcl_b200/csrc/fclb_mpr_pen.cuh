// fclb_mpr_pen.cuh -- per-thread MPR penetration queries.
//
// Behavioural contract (reference include/fcl/cvx_collide):
//   MPR::RunDirectedPenetration            mpr.hpp:497-618, finalize :629-704
//   MPR::findPortal / updatePortal         mpr.hpp:188-333 / :436-493 (with the support directions tracked)
//   MPR::RunIncrementalMinimumPenetrationDistance, incrementalMinimumDistanceExploreDirection,
//   computeExploredDistanceLowerUpperBound, finalizeIncrementalPenetrationResult
//                                           mpr_incremental_penetration.hpp:10-337
// driven per contact by detail::computePenetrationMPR (narrowphase/collision_penetration-inl.h:107-186),
// which fcl::collide runs for the DirectedPenetration / IncrementalMinimumPenetration request modes
// (collision_interface-inl.h:22-30 -> collisionPenetrationMPR, collision_penetration-inl.h:189-252).
#pragma once
#include "fclb_mpr.cuh"

namespace fclb {

// the portal (v1..v3) together with the directions that produced each vertex
template <typename S>
struct Portal {
  V3<S> v1, v2, v3;
  V3<S> d1, d2, d3;
};

enum MprPenStatus : int { PEN_OK = 0, PEN_ITERATION_LIMIT = 1, PEN_NO_INTERSECT = 2, PEN_FAILED = 3 };

// findPortal with direction tracking: 0 IterationLimit, 1 DetectSeparated, 2 PortalFound
template <typename S, typename MD>
FCLB_DI int mprFindPortalDirs(const MD& shape, const V3<S>& v0, Portal<S>& p, int max_iterations) {
  const S dot_eps_ratio = numeric_eps<S>::value();
  const S v0_abs = absNorm(v0);
  int it = 0;
  while (true) {
    if (it >= max_iterations) return 0;
    it += 1;
    V3<S> v0v1 = p.v1 - v0;
    V3<S> v0v2 = p.v2 - v0;
    V3<S> v0v3 = p.v3 - v0;
    V3<S> n031 = cross(v0v3, v0v1);
    V3<S> n012 = cross(v0v1, v0v2);
    const S signed_volume = dot(v0v2, n031);
    if (signed_volume < 0) {
      swap3(p.v2, p.v3);
      swap3(p.d2, p.d3);
      swap3(v0v2, v0v3);
      swap3(n012, n031);
      n031 = n031 * S(-1);
      n012 = n012 * S(-1);
    }
    if (dot(v0, n031) > dot_eps_ratio * v0_abs * absNorm(n031)) {
      V3<S> d = n031 * S(-1);
      p.v2 = mprSupport(shape, d, nullptr);
      p.d2 = d;
      if (dot(p.v2, d) < 0) return 1;
      continue;
    }
    if (dot(v0, n012) > dot_eps_ratio * v0_abs * absNorm(n012)) {
      V3<S> d = n012 * S(-1);
      p.v3 = mprSupport(shape, d, nullptr);
      p.d3 = d;
      if (dot(p.v3, d) < 0) return 1;
      continue;
    }
    const V3<S> n023 = cross(v0v2, v0v3);
    if (dot(v0, n023) > dot_eps_ratio * v0_abs * absNorm(n023)) {
      V3<S> d = n023 * S(-1);
      p.v1 = mprSupport(shape, d, nullptr);
      p.d1 = d;
      if (dot(p.v1, d) < 0) return 1;
      continue;
    }
    return 2;
  }
}

// updatePortal with direction tracking (mpr.hpp:436-493)
template <typename S>
FCLB_DI void mprUpdatePortalDirs(const V3<S>& v0, const V3<S>& v4, const V3<S>& n123, Portal<S>& p) {
  const V3<S> n = cross(v4, v0);
  S d = dot(p.v1, n);
  if (d > 0) {
    d = dot(p.v2, n);
    if (d > 0) {
      p.v1 = v4;
      p.d1 = n123;
    } else {
      p.v3 = v4;
      p.d3 = n123;
    }
  } else {
    d = dot(p.v3, n);
    if (d > 0) {
      p.v2 = v4;
      p.d2 = n123;
    } else {
      p.v1 = v4;
      p.d1 = n123;
    }
  }
}

// v0 scaled to the largest portal vertex norm (mpr.hpp:556-565)
template <typename S>
FCLB_DI V3<S> scaledV0(const V3<S>& v0, const Portal<S>& p) {
  const S n1 = sqnorm(p.v1), n2 = sqnorm(p.v2), n3 = sqnorm(p.v3);
  const S mx = fmax_(n1, fmax_(n2, n3));
  return v0 * fsqrt(mx);
}

// barycentric witness points of a point on the ray inside portal triangle v1 v2 v3
template <typename S, typename MD>
FCLB_DI void portalWitness(const MD& shape, const Portal<S>& p, const V3<S>& o_projected, V3<S>& p0, V3<S>& p1) {
  const V3<S> v1v2 = p.v2 - p.v1;
  const V3<S> v1v3 = p.v3 - p.v1;
  const S area = norm(cross(v1v2, v1v3));
  const S s2 = norm(cross(v1v3, p.v1 - o_projected)) / area;
  const S s3 = norm(cross(v1v2, p.v1 - o_projected)) / area;
  const S s1 = S(1.0) - s2 - s3;
  p0 = (shape.support0(p.d1) * s1 + shape.support0(p.d2) * s2) + shape.support0(p.d3) * s3;
  p1 = (shape.support1(-p.d1) * s1 + shape.support1(-p.d2) * s2) + shape.support1(-p.d3) * s3;
}

template <typename S>
struct DirectedPenOut {
  S distance;
  V3<S> p0, p1;
};

// finalizeDirectedPenetrationResult (mpr.hpp:629-704)
template <typename S, typename MD>
FCLB_DI void finalizeDirected(const MD& shape, const V3<S>& d, const V3<S>& n123, const Portal<S>& p, const V3<S>* v4,
                              DirectedPenOut<S>& out) {
  const S n_dot_d = dot(d, n123);
  if (fabs_(n_dot_d) <= 0) {
    const S dd[3] = {dot(p.v1, d), dot(p.v2, d), dot(p.v3, d)};
    S max_distance = -(sizeof(S) == 4 ? S(__int_as_float(0x7f800000)) : S(__longlong_as_double(0x7ff0000000000000ll)));
    int mi = 0;
    for (int i = 0; i < 3; i++)
      if (dd[i] > max_distance) {
        max_distance = dd[i];
        mi = i;
      }
    out.distance = max_distance;
    const V3<S> dm = mi == 0 ? p.d1 : (mi == 1 ? p.d2 : p.d3);
    out.p0 = shape.support0(dm);
    out.p1 = shape.support1(-dm);
    return;
  }
  const S v1_dot_n = dot(p.v1, n123);
  const S distance_to_v123 = v1_dot_n / n_dot_d;
  out.distance = v4 ? (dot(*v4, n123) / n_dot_d) : distance_to_v123;
  portalWitness(shape, p, distance_to_v123 * d, out.p0, out.p1);
}

// MPR::RunDirectedPenetration (mpr.hpp:497-618)
template <typename S, typename MD>
FCLB_DI int mprDirectedPenetration(const MD& shape, const V3<S>& d, int max_iterations, S tolerance, DirectedPenOut<S>& out) {
  const V3<S> v0 = -d;
  Portal<S> p;
  p.d1 = d;
  p.v1 = mprSupport(shape, p.d1, nullptr);
  if (dot(p.d1, p.v1) < 0) return PEN_NO_INTERSECT;
  p.d2 = cross(v0, p.v1);
  if (absNorm(p.d2) <= absNorm(p.v1) * tolerance) {
    out.distance = dot(p.v1, d);
    out.p0 = shape.support0(p.d1);
    out.p1 = shape.support1(-p.d1);
    return PEN_OK;
  }
  p.v2 = mprSupport(shape, p.d2, nullptr);
  if (dot(p.d2, p.v2) < 0) return PEN_NO_INTERSECT;
  p.d3 = cross(p.v1, p.v2);
  if (dot(p.d3, v0) > 0) {
    swap3(p.v1, p.v2);
    swap3(p.d1, p.d2);
    p.d3 = p.d3 * S(-1);
  }
  p.v3 = mprSupport(shape, p.d3, nullptr);
  if (dot(p.d3, p.v3) < 0) return PEN_NO_INTERSECT;
  const V3<S> v0s = scaledV0(v0, p);
  const int fp = mprFindPortalDirs(shape, v0s, p, max_iterations);
  if (fp == 0) return PEN_FAILED;
  if (fp == 1) return PEN_NO_INTERSECT;
  int it = 0;
  while (it < max_iterations) {
    it += 1;
    V3<S> n123 = cross(p.v2 - p.v1, p.v3 - p.v1);
    if (dot(n123, d) < 0) {
      swap3(p.v2, p.v3);
      swap3(p.d2, p.d3);
      n123 = n123 * S(-1);
    }
    const V3<S> v4 = mprSupport(shape, n123, nullptr);
    if (dot(v4, n123) < 0) return PEN_NO_INTERSECT;
    const V3<S> v1v4 = v4 - p.v1;
    if (fabs_(dot(v1v4, n123)) < tolerance * absNorm(n123)) {
      finalizeDirected(shape, d, n123, p, &v4, out);
      return PEN_OK;
    }
    mprUpdatePortalDirs(v0s, v4, n123, p);
  }
  return PEN_ITERATION_LIMIT;  // FailedRefinementIterationLimit: not OK for the caller
}

enum ExploreStatus : int { EX_NEW_DIRECTION = 0, EX_CONVERGE = 1, EX_DEGENERATED = 2, EX_NO_INTERSECT = 3, EX_FAILED = 4 };

// incrementalMinimumDistanceExploreDirection (mpr_incremental_penetration.hpp:40-188)
template <typename S, typename MD>
FCLB_DI int mprExploreDirection(const MD& shape, const V3<S>& d, Portal<S>& p, S& lb, S& ub, V3<S>& new_direction,
                                bool v123_valid, int max_iterations, S tolerance) {
  const V3<S> v0 = -d;
  V3<S> init_dir = d;
  const V3<S> init_support = mprSupport(shape, init_dir, nullptr);
  if (dot(d, init_support) < 0) return EX_NO_INTERSECT;
  if (!v123_valid) {
    p.d1 = d;
    p.v1 = init_support;
    p.d2 = cross(v0, p.v1);
    if (absNorm(p.d2) <= absNorm(p.v1) * tolerance) {
      lb = dot(p.v1, d);
      ub = lb;
      new_direction = d;
      return EX_DEGENERATED;
    }
    p.v2 = mprSupport(shape, p.d2, nullptr);
    if (dot(p.d2, p.v2) < 0) return EX_NO_INTERSECT;
    p.d3 = cross(p.v1, p.v2);
    if (dot(p.d3, v0) > 0) {
      swap3(p.v1, p.v2);
      swap3(p.d1, p.d2);
      p.d3 = p.d3 * S(-1);
    }
    p.v3 = mprSupport(shape, p.d3, nullptr);
    if (dot(p.d3, p.v3) < 0) return EX_NO_INTERSECT;
  }
  const V3<S> v0s = scaledV0(v0, p);
  const int fp = mprFindPortalDirs(shape, v0s, p, max_iterations);
  if (fp == 0) return EX_FAILED;
  if (fp == 1) return EX_NO_INTERSECT;
  int it = 0;
  while (it < max_iterations) {
    it += 1;
    V3<S> n123 = cross(p.v2 - p.v1, p.v3 - p.v1);
    if (dot(n123, d) < 0) {
      swap3(p.v2, p.v3);
      swap3(p.d2, p.d3);
      n123 = n123 * S(-1);
    }
    const V3<S> v4 = mprSupport(shape, n123, nullptr);  // n123 is unit from here on
    const S ub_plane = dot(v4, n123);
    if (ub_plane < 0) return EX_NO_INTERSECT;
    const S v1_dot_n = dot(p.v1, n123);
    const S v4_dot_n = dot(v4, n123);
    const S v1v4_on_n = v4_dot_n - v1_dot_n;
    {  // computeExploredDistanceLowerUpperBound (:10-36)
      const S n_dot_d = dot(d, n123);
      if (fabs_(n_dot_d) <= 0) {
        lb = fmax_(dot(p.v1, d), fmax_(dot(p.v2, d), dot(p.v3, d)));
        ub = lb;
      } else {
        lb = v1_dot_n / n_dot_d;
        ub = v4_dot_n / n_dot_d;
      }
    }
    if (lb > ub_plane + tolerance) {
      new_direction = n123;
      return EX_NEW_DIRECTION;
    }
    if (fabs_(v1v4_on_n) <= tolerance) {
      new_direction = n123;
      return EX_CONVERGE;
    }
    mprUpdatePortalDirs(v0s, v4, n123, p);
  }
  return EX_FAILED;
}

template <typename S>
struct IncrementalPenOut {
  S minimum_penetration;
  V3<S> direction;
  V3<S> p0, p1;
};

// finalizeIncrementalPenetrationResult (mpr_incremental_penetration.hpp:304-334)
template <typename S, typename MD>
FCLB_DI void finalizeIncremental(const MD& shape, const V3<S>& d, const Portal<S>& p, S lb, S ub, IncrementalPenOut<S>& out) {
  out.minimum_penetration = ub;
  out.direction = d;
  portalWitness(shape, p, lb * d, out.p0, out.p1);
}

// MPR::RunIncrementalMinimumPenetrationDistance (:191-290).  return_on_subroutine_converge defaults to
// true (mpr.h:93-97) and computePenetrationMPR does not override it.
template <typename S, typename MD>
FCLB_DI int mprIncrementalPenetration(const MD& shape, const V3<S>& init_direction, int max_iteration, S tolerance,
                                      IncrementalPenOut<S>& out, bool return_on_subroutine_converge = true) {
  V3<S> d = init_direction;
  Portal<S> p;
  p.v1 = p.v2 = p.v3 = p.d1 = p.d2 = p.d3 = zero3<S>();
  V3<S> prev_direction = init_direction;
  S prev_lb = S(-1), prev_ub = S(-1);
  #pragma unroll 1
  for (int outer = 0; outer < max_iteration; outer++) {
    V3<S> new_d = zero3<S>();
    S lb = S(0), ub = S(0);
    const int st = mprExploreDirection(shape, d, p, lb, ub, new_d, outer >= 1, max_iteration, tolerance);
    if (st == EX_NEW_DIRECTION) {
      d = new_d;
    } else if (st == EX_CONVERGE) {
      if (return_on_subroutine_converge) {
        finalizeIncremental(shape, d, p, lb, ub, out);
        return PEN_OK;
      }
      if (outer >= 1 && fabs_(ub - prev_ub) <= tolerance) {
        finalizeIncremental(shape, d, p, lb, ub, out);
        return PEN_OK;
      }
      if (norm(new_d - prev_direction) < S(1e-3)) {
        finalizeIncremental(shape, d, p, lb, ub, out);
        return PEN_OK;
      }
      d = new_d;
    } else if (st == EX_DEGENERATED) {
      out.minimum_penetration = ub;
      out.direction = new_d;
      out.p0 = shape.support0(new_d);
      out.p1 = shape.support1(-new_d);
      return PEN_OK;
    } else if (st == EX_NO_INTERSECT) {
      return PEN_NO_INTERSECT;
    } else {
      return PEN_FAILED;
    }
    prev_direction = new_d;
    prev_lb = lb;
    prev_ub = ub;
  }
  finalizeIncremental(shape, d, p, prev_lb, prev_ub, out);
  return PEN_ITERATION_LIMIT;
}

// detail::computePenetrationMPR (collision_penetration-inl.h:107-186): contact {pos, normal, depth}
template <typename S, typename MD>
FCLB_DI void computePenetrationMpr(const MD& md, const Pose<S>& tf1, const V3<S>& dir_world, bool incremental, int max_iter,
                                   S tolerance, V3<S>& pos, V3<S>& normal, S& depth) {
  const V3<S> dir = mulMtV(tf1.R, dir_world);
  if (incremental) {
    IncrementalPenOut<S> o;
    const int st = mprIncrementalPenetration(md, dir, max_iter, tolerance, o);
    if (st == PEN_OK || st == PEN_ITERATION_LIMIT) {
      depth = o.minimum_penetration;
      pos = apply(tf1, S(0.5) * (o.p0 + o.p1));
      normal = mulMV(tf1.R, o.direction);
      return;
    }
  } else {
    DirectedPenOut<S> o;
    const int st = mprDirectedPenetration(md, dir, max_iter, tolerance, o);
    if (st == PEN_OK) {
      depth = o.distance;
      pos = apply(tf1, S(0.5) * (o.p0 + o.p1));
      normal = dir_world;
      return;
    }
  }
  pos = zero3<S>();
  depth = S(-1);
  normal = dir_world;
}

}  // namespace fclb
