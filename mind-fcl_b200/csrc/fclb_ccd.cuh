// fclb_ccd.cuh -- translational continuous collision of two shapes: shape 1 sweeps along a straight segment.
//
// Behavioural contract (reference include/fcl/narrowphase/detail/ccd):
//   ShapePairTranslationalCollisionImpl::RunIntersect       shape_pair_ccd-inl.h:57-131 (Box-Box :119-137)
//   TranslationalCollisionGJK::CheckSweptVolumeCollision*   gjk_ccd-inl.h:21-114  (MPR on the Minkowski difference
//                                                           of the SWEPT shape 1 and shape 2)
//   computeOneTocSampleForIntersection                      gjk_ccd-inl.h:116-186
//   swept-volume support / interior                         cvx_collide/gjk_shape.hpp:90-141
//   BoxPairTranslationalCCD::IsDisjoint / isDisjointArray   box_pair_ccd-inl.h:101-214,374-392
//   MPR::RunIntersect with IntersectData                    cvx_collide/mpr.hpp:34-186
// reached through fcl::translational_ccd (narrowphase/continuous_collision-inl.h:21-36) and the shape-shape entries
// of TranslationalCollisionFunctionMatrix (ccd/translational_collision_func_matrix-inl.h:358-426).
#pragma once
#include "fclb_bound.h"
#include "fclb_internal.h"
#include "fclb_mpr_pen.cuh"

namespace fclb {

enum CcdRequestType : int { CCD_NOT_REQUESTED = 0, CCD_BOX_APPROXIMATE = 1, CCD_ONE_TOC_SAMPLE = 2 };  // ccd_request.h:10-16

template <typename S>
struct TocInterval {  // ccd_typedef.h:19-33
  S lo, hi;
  FCLB_DI void intersect(const TocInterval& r) {
    lo = fmax_(lo, r.lo);
    hi = fmin_(hi, r.hi);
  }
  FCLB_DI bool empty(S tol) const { return lo > hi + tol; }
};

// Minkowski difference of (shape 0 swept by `disp`, shape 1): gjk_shape.hpp:93-103,133-139
template <typename S, int T0, int T1>
struct SweptMinkDiff {
  MinkDiff<S, T0, T1> md;
  V3<S> disp;  // displacement of shape 0 in its own frame
  FCLB_DI V3<S> support0(const V3<S>& d) const {
    const V3<S> p = md.support0(d);
    return (dot(d, disp) > 0) ? (p + disp) : p;
  }
  FCLB_DI V3<S> support1(const V3<S>& d) const { return md.support1(d); }
  FCLB_DI V3<S> support(const V3<S>& d) const { return support0(d) - support1(-d); }
  FCLB_DI V3<S> interior() const {
    const V3<S> i0 = interiorOf(md.s0) + S(0.5) * disp;
    return i0 - apply(md.toshape0, interiorOf(md.s1));
  }
};

template <typename S>
struct MprIntersectData {  // mpr.h:23-33; unset vertices are NaN
  V3<S> v0, v1, v2, v3, d1, d2, d3;
};

template <typename S>
FCLB_DI S qnan() {
  return sizeof(S) == 4 ? S(__int_as_float(0x7fc00000)) : S(__longlong_as_double(0x7ff8000000000000ll));
}

// MPR::RunIntersect with the portal and its support directions kept (mpr.hpp:34-186)
template <typename S, typename MD>
FCLB_DI int mprIntersectData(const MD& shape, int max_iterations, S tolerance, MprIntersectData<S>& out) {
  const S nan = qnan<S>();
  Portal<S> p;
  p.v1 = p.v2 = p.v3 = mk<S>(nan, nan, nan);
  p.d1 = p.d2 = p.d3 = zero3<S>();
  const V3<S> v0 = shape.interior();
  out.v0 = v0;
  auto done = [&](int status) {
    out.v1 = p.v1; out.v2 = p.v2; out.v3 = p.v3;
    out.d1 = p.d1; out.d2 = p.d2; out.d3 = p.d3;
    return status;
  };
  if (sqnorm(v0) <= tolerance * tolerance) return done(MPR_INTERSECT);
  p.d1 = -v0;
  p.v1 = mprSupport(shape, p.d1, nullptr);
  if (dot(p.d1, p.v1) < 0) return done(MPR_SEPARATED);
  p.d2 = cross(v0, p.v1);
  if (absNorm(p.d2) <= absNorm(v0) * absNorm(p.v1) * tolerance) return done(MPR_INTERSECT);
  p.v2 = mprSupport(shape, p.d2, nullptr);
  if (dot(p.d2, p.v2) < 0) return done(MPR_SEPARATED);
  p.d3 = cross(p.v1 - v0, p.v2 - v0);
  if (dot(p.d3, v0) > 0) {
    swap3(p.v1, p.v2);
    swap3(p.d1, p.d2);
    p.d3 = p.d3 * S(-1);
  }
  p.v3 = mprSupport(shape, p.d3, nullptr);
  if (dot(p.d3, p.v3) < 0) return done(MPR_SEPARATED);
  const int fp = mprFindPortalDirs(shape, v0, p, max_iterations);
  if (fp == 0) return done(MPR_FAILED);
  if (fp == 1) return done(MPR_SEPARATED);
  int it = 0;
  while (it < max_iterations) {
    it += 1;
    V3<S> n123 = cross(p.v2 - p.v1, p.v3 - p.v1);
    if (dot(n123, v0) > 0) {
      swap3(p.v2, p.v3);
      swap3(p.d2, p.d3);
      n123 = n123 * S(-1);
    }
    if (!(dot(p.v1, n123) < 0)) return done(MPR_INTERSECT);
    const V3<S> v4 = mprSupport(shape, n123, nullptr);
    if (dot(v4, n123) < 0) return done(MPR_SEPARATED);
    const V3<S> v1v4 = v4 - p.v1;
    if (fabs_(dot(v1v4, n123)) < tolerance * absNorm(n123)) return done(MPR_SEPARATED);
    mprUpdatePortalDirs(v0, v4, n123, p);
  }
  return done(MPR_FAILED);
}

template <typename S>
FCLB_DI bool anyNan(const V3<S>& v) {
  return v.x != v.x || v.y != v.y || v.z != v.z;
}

// computeOneTocSampleForIntersection (gjk_ccd-inl.h:116-186)
template <typename S>
FCLB_DI S oneTocSample(const V3<S>& disp, const MprIntersectData<S>& d) {
  const S v0_phase = S(0.5);
  if (anyNan(d.v1)) return v0_phase;
  if (anyNan(d.v2)) {
    const S n0 = norm(d.v0), n1 = norm(d.v1);
    const S total = n0 + n1;
    if (total <= S(0.0)) return v0_phase;
    const S inv = S(1.0) / total;
    const S w0 = n1 * inv, w1 = n0 * inv;
    const S v1_phase = (dot(d.d1, disp) > 0) ? S(1.0) : S(0.0);
    return w0 * v0_phase + w1 * v1_phase;
  }
  const V3<S> v01 = d.v1 - d.v0, v02 = d.v2 - d.v0, v03 = d.v3 - d.v0;
  const S vol012 = fabs_(dot(cross(v01, v02), d.v0));
  const S vol013 = fabs_(dot(cross(v01, v03), d.v0));
  const S vol023 = fabs_(dot(cross(v02, v03), d.v0));
  const S vol123 = fabs_(dot(cross(d.v2 - d.v1, d.v3 - d.v1), d.v1));
  const S total = vol012 + vol013 + vol023 + vol123;
  if (total <= S(0.0)) return S(0.0);
  const S inv = S(1.0) / total;
  const S w0 = vol123 * inv, w1 = vol023 * inv, w2 = vol013 * inv, w3 = vol012 * inv;
  const S ph1 = (dot(d.d1, disp) > 0) ? S(1.0) : S(0.0);
  const S ph2 = (dot(d.d2, disp) > 0) ? S(1.0) : S(0.0);
  const S ph3 = (dot(d.d3, disp) > 0) ? S(1.0) : S(0.0);
  return w0 * v0_phase + w1 * ph1 + w2 * ph2 + w3 * ph3;
}

// ---- swept box pair (box_pair_ccd-inl.h) -----------------------------------------------------------------------
// CheckEmptyTocIntervalOnAxis, Array3 form (:45-82): three axes at once, scaled to [0, 1]
template <typename S>
FCLB_DI bool tocOnAxes3(const V3<S>& h1, const V3<S>& disp, const V3<S>& h2, const V3<S>& off, TocInterval<S>& toc, S zero_tol) {
  const S h1a[3] = {h1.x, h1.y, h1.z}, da[3] = {disp.x, disp.y, disp.z}, h2a[3] = {h2.x, h2.y, h2.z}, oa[3] = {off.x, off.y, off.z};
  S lower[3], upper[3];
  bool disjoint = false;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const bool pos = da[k] > S(0.0);
    const S b1_lb = pos ? -h1a[k] : (-h1a[k] + da[k]);
    const S b1_ub = pos ? (h1a[k] + da[k]) : h1a[k];
    const S b2_lb = oa[k] - h2a[k], b2_ub = oa[k] + h2a[k];
    if (b1_lb > b2_ub || b2_lb > b1_ub) disjoint = true;
    const S lo_raw = pos ? (-h1a[k] + b2_lb) : (-h1a[k] - b2_ub);
    const S hi_raw = pos ? (h1a[k] + b2_ub) : (h1a[k] - b2_lb);
    const S inv_scale = S(1.0) / fmax_(fabs_(da[k]), zero_tol);
    lower[k] = fmax_(lo_raw * inv_scale, S(0.0));
    upper[k] = fmin_(hi_raw * inv_scale, S(1.0));
  }
  if (disjoint) return true;
  toc.lo = fmax_(fmax_(lower[0], lower[1]), lower[2]);
  toc.hi = fmin_(fmin_(upper[0], upper[1]), upper[2]);
  return toc.empty(S(0.0));
}
// scalar form (:10-43)
template <typename S>
FCLB_DI bool tocOnAxis(S h1, S abs_disp, bool pos, S h2, S off, TocInterval<S>& toc) {
  const S b1_lb = pos ? (-h1) : (-h1 - abs_disp);
  const S b1_ub = pos ? (h1 + abs_disp) : h1;
  const S b2_lb = off - h2, b2_ub = off + h2;
  if (b1_lb > b2_ub || b2_lb > b1_ub) return true;
  if (pos) {
    toc.lo = fmax_(S(0.0), -h1 + b2_lb);
    toc.hi = fmin_(abs_disp, h1 + b2_ub);
  } else {
    toc.lo = fmax_(S(0.0), -h1 - b2_ub);
    toc.hi = fmin_(abs_disp, h1 - b2_lb);
  }
  return false;
}

// BoxPairTranslationalCCD::IsDisjoint -> isDisjointArray (:374-392, :101-214).  axis1 / axis2: box rotations (columns =
// box directions), unit_axis: displacement direction in box 1's frame.
template <typename S>
FCLB_DI bool boxPairCcdDisjoint(const M3<S>& axis1, const V3<S>& To1, const V3<S>& ext1, const V3<S>& unit_axis, S scalar_disp,
                                const M3<S>& axis2, const V3<S>& To2, const V3<S>& ext2, TocInterval<S>& interval, S zero_tol,
                                bool keep_init = false) {  // keep_init: `interval` holds init_toc_interval_bound (:378-392)
  const V3<S> t_world = To2 - To1;
  const V3<S> t21 = mulMtV(axis1, t_world);
  const M3<S> R = mulMtM(axis1, axis2);  // rotation_2in1
  M3<S> Rabs;
#pragma unroll
  for (int i = 0; i < 9; i++) Rabs.m[i] = fabs_(R.m[i]);
  if (!keep_init) {
    interval.lo = S(0.0);
    interval.hi = S(1.0);
  }
  {  // box-1 axes
    const V3<S> disp1 = unit_axis * scalar_disp;
    const V3<S> h2 = mulMV(Rabs, ext2);
    TocInterval<S> t;
    if (tocOnAxes3(ext1, disp1, h2, t21, t, zero_tol)) return true;
    interval.intersect(t);
    if (interval.empty(S(0.0))) return true;
  }
  {  // box-2 axes
    const V3<S> h1 = mulMtV(Rabs, ext1);
    const V3<S> unit2 = mulMtV(R, unit_axis);
    const V3<S> disp2 = unit2 * scalar_disp;
    const V3<S> off = mulMtV(R, t21);
    TocInterval<S> t;
    if (tocOnAxes3(h1, disp2, ext2, off, t, zero_tol)) return true;
    interval.intersect(t);
    if (interval.empty(S(0.0))) return true;
  }
  for (int k = 0; k < 3; k++) {
    const V3<S> ek = mk<S>(k == 0 ? S(1) : S(0), k == 1 ? S(1) : S(0), k == 2 ? S(1) : S(0));
    for (int i = 0; i < 3; i++) {
      const V3<S> a1 = cross(ek, col(R, i));
      const V3<S> a1abs = mk<S>(fabs_(a1.x), fabs_(a1.y), fabs_(a1.z));
      const S h1 = dot(a1abs, ext1);
      const S off = dot(a1, t21);
      const V3<S> a2 = mulMtV(R, a1);
      const V3<S> a2abs = mk<S>(fabs_(a2.x), fabs_(a2.y), fabs_(a2.z));
      const S h2 = dot(a2abs, ext2);
      const S proj = dot(a1, unit_axis);
      const bool pos = proj > 0;
      const S abs_disp = pos ? proj * scalar_disp : -proj * scalar_disp;
      TocInterval<S> t;
      if (tocOnAxis(h1, abs_disp, pos, h2, off, t)) return true;
      if (abs_disp < zero_tol) {  // ScaleIntervalBoxDisjoint (:92-99)
        t.lo = S(0);
        t.hi = S(1);
      } else {
        t.lo /= abs_disp;
        t.hi /= abs_disp;
      }
      interval.intersect(t);
      if (interval.empty(S(0.0))) return true;
    }
  }
  return false;
}

struct CcdArgs {
  const void* shapes;   // ShapeD<S>[]
  const void* convex;   // ConvexD<S>[]
  const void* local;    // LocalAabbD<S>[] (aabb_local of every table entry)
  const fclb_pair* pairs;
  const void* poses1;
  const void* poses2;
  const void* disp;     // 4 S per query: unit axis in shape 1's frame, scalar displacement
  size_t n;
  int request_type;
  double zero_tol, gjk_tol;
  int max_iter;
  uint8_t* hit;
  void* toc;            // 2 S per query (lower, upper) or nullptr
};

// ShapePairTranslationalCollisionImpl<S, Shape1, Shape2>::RunIntersect (shape_pair_ccd-inl.h:55-137) for one pair: the
// Box-Box specialisation, or the generic swept-volume MPR with the request type's extras.  l1 / l2: aabb_local of the
// two shapes (read by kBoxApproximate only).  Returns the hit flag; toc as RunIntersect leaves it.
template <typename S>
FCLB_DI bool ccdShapePairEval(const ShapeInst<S>& s1, const LocalAabbD<S>& l1, const Pose<S>& tf1, const ShapeInst<S>& s2,
                              const LocalAabbD<S>& l2, const Pose<S>& tf2, const V3<S>& unit_axis, S scalar_disp, int request_type,
                              S zero_tol, S tol, int max_iter, TocInterval<S>& toc) {
  toc.lo = toc.hi = S(-1.0);
  if (s1.type == ST_BOX && s2.type == ST_BOX) {  // computeBV<OBB, Box> + the swept box test, whatever the request type
    const V3<S> e1 = mk<S>(s1.p0, s1.p1, s1.p2) * S(0.5), e2 = mk<S>(s2.p0, s2.p1, s2.p2) * S(0.5);
    return !boxPairCcdDisjoint(tf1.R, tf1.t, e1, unit_axis, scalar_disp, tf2.R, tf2.t, e2, toc, zero_tol);
  }
  SweptMinkDiff<S, ST_DYNAMIC, ST_DYNAMIC> sm;
  sm.md.s0 = s1;
  sm.md.s1 = s2;
  sm.md.setPoses(tf1, tf2);
  sm.disp = unit_axis * scalar_disp;
  if (request_type == CCD_BOX_APPROXIMATE) {  // convertBV(aabb_local, tf) boxes first (:98-110)
    auto obbOf = [](const LocalAabbD<S>& l, const Pose<S>& tf, V3<S>& To, V3<S>& ext) {
      const V3<S> c = mk<S>(l.center[0], l.center[1], l.center[2]);
      To = mk<S>(((tf.R.m[0] * c.x + tf.R.m[1] * c.y) + tf.R.m[2] * c.z) + tf.t.x,
                 ((tf.R.m[3] * c.x + tf.R.m[4] * c.y) + tf.R.m[5] * c.z) + tf.t.y,
                 ((tf.R.m[6] * c.x + tf.R.m[7] * c.y) + tf.R.m[8] * c.z) + tf.t.z);
      ext = mk<S>((l.mx[0] - l.mn[0]) * S(0.5), (l.mx[1] - l.mn[1]) * S(0.5), (l.mx[2] - l.mn[2]) * S(0.5));
    };
    V3<S> To1, e1, To2, e2;
    obbOf(l1, tf1, To1, e1);
    obbOf(l2, tf2, To2, e2);
    if (boxPairCcdDisjoint(tf1.R, To1, e1, unit_axis, scalar_disp, tf2.R, To2, e2, toc, zero_tol)) return false;
  }
  if (request_type == CCD_ONE_TOC_SAMPLE) {
    MprIntersectData<S> data;
    const bool hit = mprIntersectData<S>(sm, max_iter, tol, data) == MPR_INTERSECT;
    if (hit) toc.lo = toc.hi = oneTocSample(sm.disp, data);
    return hit;
  }
  const bool hit = mprIntersect<S>(sm, max_iter, tol, nullptr) == MPR_INTERSECT;
  if (request_type == CCD_NOT_REQUESTED) toc.lo = toc.hi = S(-1.0);
  return hit;
}

// ShapePairTranslationalCollisionSolver::RunShapePair for one query (shape_pair_ccd-inl.h:139-170)
template <typename S>
__global__ void __launch_bounds__(kBlock) translationalCcdKernel(CcdArgs a) {
  const ShapeD<S>* __restrict__ shapes = static_cast<const ShapeD<S>*>(a.shapes);
  const ConvexD<S>* __restrict__ cvx = static_cast<const ConvexD<S>*>(a.convex);
  const LocalAabbD<S>* __restrict__ local = static_cast<const LocalAabbD<S>*>(a.local);
  const S* __restrict__ disp = static_cast<const S*>(a.disp);
  const S zero_tol = S(a.zero_tol), tol = S(a.gjk_tol);
  #pragma unroll 1
  for (size_t q = blockIdx.x * size_t(blockDim.x) + threadIdx.x; q < a.n; q += size_t(gridDim.x) * blockDim.x) {
    const fclb_pair pr = a.pairs[q];
    const Pose<S> tf1 = loadPose(static_cast<const S*>(a.poses1), q);
    const Pose<S> tf2 = loadPose(static_cast<const S*>(a.poses2), q);
    const V3<S> unit_axis = mk<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
    const S scalar_disp = disp[4 * q + 3];
    TocInterval<S> toc;
    const bool hit = ccdShapePairEval<S>(bindShape(shapes, cvx, pr.shape1), local[pr.shape1], tf1, bindShape(shapes, cvx, pr.shape2),
                                         local[pr.shape2], tf2, unit_axis, scalar_disp, a.request_type, zero_tol, tol, a.max_iter, toc);
    a.hit[q] = hit ? 1 : 0;
    if (a.toc) {
      S* o = static_cast<S*>(a.toc) + 2 * q;
      // ContinuousContactMeta::writeToContact(contact, toc): the interval is kept only when both bounds are >= 0
      const bool valid = hit && toc.lo >= 0 && toc.hi >= 0;
      o[0] = valid ? toc.lo : S(-1.0);
      o[1] = valid ? toc.hi : S(-1.0);
    }
  }
}

// FixedOrientationBoxPairTranslationalCCD (box_pair_ccd_fixed_orientation-inl.h): box 1 = an AABB in frame 1 that moves
// along `unit` (frame 1), box 2 = an AABB in frame 2; R / t = frame 2 in frame 1.  Same 15 axes as above, but every face
// axis goes through the scalar interval routine with its own early exit (the OBB form evaluates three at once).
template <typename S>
struct FixedCcd {
  M3<S> R;
  V3<S> t, unit;
  S scalar;
};
template <typename S>
FCLB_DI bool scaleAndIntersect(TocInterval<S>& interval, TocInterval<S> t, S abs_disp, S zero_tol) {
  if (abs_disp < zero_tol) {  // ScaleIntervalBoxDisjoint (box_pair_ccd-inl.h:92-104)
    t.lo = S(0);
    t.hi = S(1);
  } else {
    t.lo /= abs_disp;
    t.hi /= abs_disp;
  }
  interval.intersect(t);
  return interval.empty(S(0.0));
}
// aabb1 / aabb2 given as min / max; `interval` holds the parent's interval on entry (:57-86)
template <typename S>
FCLB_DI bool fixedCcdDisjoint(const FixedCcd<S>& f, const V3<S>& mn1, const V3<S>& mx1, const V3<S>& mn2, const V3<S>& mx2,
                              TocInterval<S>& interval, S zero_tol) {
  const V3<S> c1 = (mn1 + mx1) * S(0.5), h1 = S(0.5) * (mx1 - mn1);
  const V3<S> c2 = (mn2 + mx2) * S(0.5), h2 = S(0.5) * (mx2 - mn2);
  const V3<S> trans = (mulMV(f.R, c2) + f.t) - c1;
  M3<S> Rabs;
#pragma unroll
  for (int i = 0; i < 9; i++) Rabs.m[i] = fabs_(f.R.m[i]);
  {  // box-1 axes
    const V3<S> r2 = mulMV(Rabs, h2);
    const S h1a[3] = {h1.x, h1.y, h1.z}, r2a[3] = {r2.x, r2.y, r2.z}, oa[3] = {trans.x, trans.y, trans.z},
            ua[3] = {f.unit.x, f.unit.y, f.unit.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const bool pos = ua[i] > 0;
      const S abs_disp = pos ? ua[i] * f.scalar : -ua[i] * f.scalar;
      TocInterval<S> t;
      if (tocOnAxis(h1a[i], abs_disp, pos, r2a[i], oa[i], t)) return true;
      if (scaleAndIntersect(interval, t, abs_disp, zero_tol)) return true;
    }
  }
  {  // box-2 axes
    const V3<S> r1 = mulMtV(Rabs, h1), u2 = mulMtV(f.R, f.unit), off = mulMtV(f.R, trans);
    const S r1a[3] = {r1.x, r1.y, r1.z}, h2a[3] = {h2.x, h2.y, h2.z}, oa[3] = {off.x, off.y, off.z}, ua[3] = {u2.x, u2.y, u2.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const bool pos = ua[i] > 0;
      const S abs_disp = pos ? ua[i] * f.scalar : -ua[i] * f.scalar;
      TocInterval<S> t;
      if (tocOnAxis(r1a[i], abs_disp, pos, h2a[i], oa[i], t)) return true;
      if (scaleAndIntersect(interval, t, abs_disp, zero_tol)) return true;
    }
  }
  for (int k = 0; k < 3; k++) {
    const V3<S> ek = mk<S>(k == 0 ? S(1) : S(0), k == 1 ? S(1) : S(0), k == 2 ? S(1) : S(0));
    for (int i = 0; i < 3; i++) {
      const V3<S> a1 = cross(ek, col(f.R, i));
      const V3<S> a1abs = mk<S>(fabs_(a1.x), fabs_(a1.y), fabs_(a1.z));
      const V3<S> a2 = mulMtV(f.R, a1);
      const V3<S> a2abs = mk<S>(fabs_(a2.x), fabs_(a2.y), fabs_(a2.z));
      const S r1 = dot(a1abs, h1), off = dot(a1, trans), r2 = dot(a2abs, h2);
      const S proj = dot(a1, f.unit);
      const bool pos = proj > 0;
      const S abs_disp = pos ? proj * f.scalar : -proj * f.scalar;
      TocInterval<S> t;
      if (tocOnAxis(r1, abs_disp, pos, r2, off, t)) return true;
      if (scaleAndIntersect(interval, t, abs_disp, zero_tol)) return true;
    }
  }
  return false;
}

}  // namespace fclb
