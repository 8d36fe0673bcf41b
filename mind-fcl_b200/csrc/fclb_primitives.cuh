// fclb_primitives.cuh -- closed-form pair routines (distance side).
//
// Behavioural contract: include/fcl/narrowphase/detail/primitive_shape_algorithm/
//   sphere_box-inl.h:59-81,167-205      sphere_capsule-inl.h:51-67,105-147
//   sphere_cylinder-inl.h:62-93,206-244 sphere_sphere-inl.h:72-89
//   capsule_capsule-inl.h:49-250
// which GJKSolver<S>::shapeDistance selects instead of GJK for these pairs
// (gjk_solver-inl.h:902-988).  Operation order follows the reference
// expression by expression (see fclb_math.cuh).
#pragma once
#include "fclb_math.cuh"

namespace fclb {

// X_B^-1 * X_A, translation part only:  (R_B^T t_A) + (-(R_B^T t_B))
template <typename S>
FCLB_DI V3<S> originInFrame(const Pose<S>& X_FB, const Pose<S>& X_FA) {
  const M3<S> Rt = transpose(X_FB.R);
  const V3<S> tinv = -mulMV(Rt, X_FB.t);
  return mulMV(Rt, X_FA.t) + tinv;
}

// sphere_box-inl.h:167-205
template <typename S>
FCLB_DI bool sphereBoxDistance(S r, const Pose<S>& X_FS, const V3<S>& side, const Pose<S>& X_FB, S& dist, V3<S>& p_FSb,
                               V3<S>& p_FBs) {
  const V3<S> p_BC = originInFrame(X_FB, X_FS);
  const V3<S> half = side / S(2);
  V3<S> p_BN = p_BC;
  bool clamped = false;
  if (p_BC.x < -half.x) { clamped = true; p_BN.x = -half.x; }
  if (p_BC.x > half.x) { clamped = true; p_BN.x = half.x; }
  if (p_BC.y < -half.y) { clamped = true; p_BN.y = -half.y; }
  if (p_BC.y > half.y) { clamped = true; p_BN.y = half.y; }
  if (p_BC.z < -half.z) { clamped = true; p_BN.z = -half.z; }
  if (p_BC.z > half.z) { clamped = true; p_BN.z = half.z; }
  if (clamped) {
    const V3<S> p_NC_B = p_BC - p_BN;
    const S sq = sqnorm(p_NC_B);
    if (sq > r * r) {
      const S d = fsqrt(sq);
      dist = d - r;
      p_FBs = apply(X_FB, p_BN);
      const V3<S> p_BSb = (p_NC_B / d) * (d - r) + p_BN;
      p_FSb = apply(X_FB, p_BSb);
      return true;
    }
  }
  dist = S(-1);
  return false;
}

// sphere_capsule-inl.h:51-67
template <typename S>
FCLB_DI V3<S> segmentPointClosestTo(const V3<S>& p, const V3<S>& s1, const V3<S>& s2) {
  const V3<S> v = s2 - s1;
  const V3<S> w = p - s1;
  const S c1 = dot(w, v);
  const S c2 = dot(v, v);
  if (c1 <= 0) return s1;
  if (c2 <= c1) return s2;
  const S b = c1 / c2;
  return s1 + v * b;
}

// sphere_capsule-inl.h:105-147
template <typename S>
FCLB_DI bool sphereCapsuleDistance(S r1, const Pose<S>& tf1, S r2, S lz, const Pose<S>& tf2, S& dist, V3<S>& p1,
                                   V3<S>& p2) {
  const V3<S> pos1 = mk<S>(S(0), S(0), S(0.5) * lz);
  const V3<S> pos2 = mk<S>(S(0), S(0), S(-0.5) * lz);
  const V3<S> s_c = apply(inverse(tf2), tf1.t);
  const V3<S> seg = segmentPointClosestTo(s_c, pos1, pos2);
  V3<S> diff = s_c - seg;
  const S distance = norm(diff) - r1 - r2;
  if (distance <= 0) {
    dist = S(-1);
    return false;
  }
  dist = distance;
  diff = normalized(diff);
  p1 = apply(tf2, s_c - diff * r1);
  p2 = apply(tf2, seg + diff * r2);
  return true;
}

// sphere_cylinder-inl.h:206-244 with nearestPointInCylinder :62-93
template <typename S>
FCLB_DI bool sphereCylinderDistance(S r_s, const Pose<S>& X_FS, S radius, S height, const Pose<S>& X_FC, S& dist,
                                    V3<S>& p_FSc, V3<S>& p_FCs) {
  const V3<S> p_CS = originInFrame(X_FC, X_FS);
  V3<S> p_CN = p_CS;
  bool clamped = false;
  const S half_h = height / S(2);
  if (p_CS.z > half_h) {
    clamped = true;
    p_CN.z = half_h;
  } else if (p_CS.z < -half_h) {
    clamped = true;
    p_CN.z = -half_h;
  }
  const S sq_xy = p_CS.x * p_CS.x + p_CS.y * p_CS.y;
  if (sq_xy > radius * radius) {
    clamped = true;
    // The reference calls an unqualified sqrt() here (sphere_cylinder-inl.h:85), which
    // for S = float resolves to the C double overload: the quotient is formed in
    // double and rounded once to S by Eigen's operator*=(Scalar).
    const S k = S(double(radius) / sqrt(double(sq_xy)));
    p_CN.x = p_CS.x * k;
    p_CN.y = p_CS.y * k;
  }
  if (clamped) {
    const V3<S> p_NS_C = p_CS - p_CN;
    const S sq = sqnorm(p_NS_C);
    if (sq > r_s * r_s) {
      const S d = fsqrt(sq);
      dist = d - r_s;
      p_FCs = apply(X_FC, p_CN);
      const V3<S> p_CSc = p_CS - ((p_NS_C * r_s) / d);
      p_FSc = apply(X_FC, p_CSc);
      return true;
    }
  }
  dist = S(-1);
  return false;
}

// sphere_sphere-inl.h:72-89
template <typename S>
FCLB_DI bool sphereSphereDistance(S r1, const Pose<S>& tf1, S r2, const Pose<S>& tf2, S& dist, V3<S>& p1, V3<S>& p2) {
  const V3<S> o1 = tf1.t, o2 = tf2.t;
  const V3<S> diff = o1 - o2;
  const S len = norm(diff);
  if (len > r1 + r2) {
    dist = len - (r1 + r2);
    p1 = o1 - diff * (r1 / len);
    p2 = o2 + diff * (r2 / len);
    return true;
  }
  dist = S(-1);
  return false;
}

template <typename S>
FCLB_DI S clampS(S n, S lo, S hi) {
  if (n < lo) return lo;
  if (n > hi) return hi;
  return n;
}

// capsule_capsule-inl.h:60-139 ; eps78 = constants<S>::eps_78()
template <typename S>
FCLB_DI S closestPtSegmentSegment(const V3<S>& P1, const V3<S>& Q1, const V3<S>& P2, const V3<S>& Q2, S eps78, V3<S>& C1,
                                  V3<S>& C2) {
  const S eps_sq = eps78 * eps78;
  const V3<S> d1 = Q1 - P1;
  const V3<S> d2 = Q2 - P2;
  const V3<S> r = P1 - P2;
  const S a = dot(d1, d1);
  const S e = dot(d2, d2);
  const S f = dot(d2, r);
  S s, t;
  if (a <= eps_sq && e <= eps_sq) {
    C1 = P1;
    C2 = P2;
    return sqnorm(C1 - C2);
  }
  if (a <= eps_sq) {
    s = S(0);
    t = clampS(f / e, S(0), S(1));
  } else {
    const S c = dot(d1, r);
    if (e <= eps_sq) {
      t = S(0);
      s = clampS(-c / a, S(0), S(1));
    } else {
      const S b = dot(d1, d2);
      const S denom = fmax_(S(0), a * e - b * b);
      if (denom > eps_sq) {
        s = clampS((b * f - c * e) / denom, S(0), S(1));
      } else {
        s = S(0);
      }
      t = (b * s + f) / e;
      if (t < S(0)) {
        t = S(0);
        s = clampS(-c / a, S(0), S(1));
      } else if (t > S(1)) {
        t = S(1);
        s = clampS((b - c) / a, S(0), S(1));
      }
    }
  }
  C1 = P1 + d1 * s;
  C2 = P2 + d2 * t;
  return sqnorm(C1 - C2);
}

// capsule_capsule-inl.h:141-246 (always "true": dist may be negative)
template <typename S>
FCLB_DI bool capsuleCapsuleDistance(S r1, S lz1, const Pose<S>& X1, S r2, S lz2, const Pose<S>& X2, S eps78, S& dist,
                                    V3<S>& pW1, V3<S>& pW2) {
  const V3<S> o1 = X1.t, o2 = X2.t;
  const V3<S> z1 = col(X1.R, 2), z2 = col(X2.R, 2);
  const V3<S> arm1 = (lz1 / S(2)) * z1;
  const V3<S> a1 = o1 + arm1, b1 = o1 - arm1;
  const V3<S> arm2 = (lz2 / S(2)) * z2;
  const V3<S> a2 = o2 + arm2, b2 = o2 - arm2;
  V3<S> N1, N2;
  const S sq = closestPtSegmentSegment(a1, b1, a2, b2, eps78, N1, N2);
  const S seg_dist = fsqrt(sq);
  dist = seg_dist - r1 - r2;
  V3<S> vhat;
  if (seg_dist > eps78) {
    vhat = (N2 - N1) / seg_dist;
  } else {
    if (fabs_(dot(z1, z2)) < S(1) - eps78) {
      vhat = normalized(cross(z1, z2));
    } else {
      vhat = col(X1.R, 0);
    }
  }
  pW1 = N1 + vhat * r1;
  pW2 = N2 - vhat * r2;
  return true;
}

}  // namespace fclb
