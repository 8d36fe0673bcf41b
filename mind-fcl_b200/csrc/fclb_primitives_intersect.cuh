// fclb_primitives_intersect.cuh -- closed-form pair routines (collide side):
// boolean + one contact {normal, pos, depth} in the world frame.
//
// Behavioural contract: include/fcl/narrowphase/detail/primitive_shape_algorithm/
//   sphere_sphere-inl.h:49-69   sphere_capsule-inl.h:69-101
//   sphere_box-inl.h:85-163     sphere_cylinder-inl.h:96-202
// selected by GJKSolver<S>::shapeIntersect instead of GJK/EPA
// (gjk_solver-inl.h:152-243).  The (Shape2, Shape1) argument order runs the same
// routine with swapped arguments and flips the normal (:186-197).
#pragma once
#include "fclb_primitives.cuh"

namespace fclb {

template <typename S>
struct ContactPt {  // narrowphase/contact_point.h:44-91
  V3<S> normal;
  V3<S> pos;
  S depth;
};

template <typename S>
struct eps16 {  // 16 * constants<S>::eps()
  static FCLB_DI S value();
};
template <>
FCLB_DI float eps16<float>::value() { return 16.f * 1.1920928955078125e-07f; }
template <>
FCLB_DI double eps16<double>::value() { return 16. * 2.220446049250313e-16; }

// sphere_sphere-inl.h:49-69
template <typename S>
FCLB_DI bool sphereSphereIntersect(S r1, const Pose<S>& tf1, S r2, const Pose<S>& tf2, bool want, ContactPt<S>& c) {
  const V3<S> diff = tf2.t - tf1.t;
  const S len = norm(diff);
  if (len > r1 + r2) return false;
  if (want) {
    c.normal = (len > 0) ? (diff / len) : diff;
    c.pos = tf1.t + (diff * r1) / (r1 + r2);
    c.depth = r1 + r2 - len;
  }
  return true;
}

// sphere_capsule-inl.h:69-101
template <typename S>
FCLB_DI bool sphereCapsuleIntersect(S r1, const Pose<S>& tf1, S r2, S lz, const Pose<S>& tf2, bool want,
                                    ContactPt<S>& c) {
  const V3<S> pos1 = mk<S>(S(0), S(0), S(0.5) * lz);
  const V3<S> pos2 = mk<S>(S(0), S(0), S(-0.5) * lz);
  const V3<S> s_c = apply(inverse(tf2), tf1.t);
  const V3<S> seg = segmentPointClosestTo(s_c, pos1, pos2);
  const V3<S> diff = s_c - seg;
  const S distance = norm(diff) - r1 - r2;
  if (distance > 0) return false;
  if (want) {
    const V3<S> local_normal = -normalized(diff);
    c.normal = mulMV(tf2.R, local_normal);
    c.pos = apply(tf2, seg + local_normal * distance);
    c.depth = -distance;
  }
  return true;
}

// sphere_box-inl.h:85-163
template <typename S>
FCLB_DI bool sphereBoxIntersect(S r, const Pose<S>& X_FS, const V3<S>& side, const Pose<S>& X_FB, bool want,
                                ContactPt<S>& c) {
  const V3<S> p_BC = originInFrame(X_FB, X_FS);
  const V3<S> half = side / S(2);
  V3<S> p_BN = p_BC;
  bool N_is_not_C = false;
  if (p_BC.x < -half.x) { N_is_not_C = true; p_BN.x = -half.x; }
  if (p_BC.x > half.x) { N_is_not_C = true; p_BN.x = half.x; }
  if (p_BC.y < -half.y) { N_is_not_C = true; p_BN.y = -half.y; }
  if (p_BC.y > half.y) { N_is_not_C = true; p_BN.y = half.y; }
  if (p_BC.z < -half.z) { N_is_not_C = true; p_BN.z = -half.z; }
  if (p_BC.z > half.z) { N_is_not_C = true; p_BN.z = half.z; }
  const V3<S> p_CN_B = p_BN - p_BC;
  const S sq = sqnorm(p_CN_B);
  if (sq > r * r) return false;
  if (want) {
    const S eps = eps16<S>::value();
    S depth;
    V3<S> n_SB_B, p_BP;
    if (N_is_not_C && sq > eps * eps) {
      const S distance = fsqrt(sq);
      n_SB_B = p_CN_B / distance;
      depth = r - distance;
      p_BP = p_BN + n_SB_B * (depth * S(0.5));
    } else {
      S min_distance = S(INFINITY);
      int min_axis = -1;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const S pc = comp(p_BC, i), hs = comp(half, i);
        const S dist = (pc >= 0) ? (hs - pc) : (pc + hs);
        if (dist + eps < min_distance) {
          min_distance = dist;
          min_axis = i;
        }
      }
      n_SB_B = zero3<S>();
      if (min_axis >= 0) setComp(n_SB_B, min_axis, (comp(p_BC, min_axis) >= 0) ? S(-1) : S(1));
      depth = min_distance + r;
      p_BP = p_BC + n_SB_B * ((r - min_distance) / S(2));
    }
    c.normal = mulMV(X_FB.R, n_SB_B);
    c.pos = apply(X_FB, p_BP);
    c.depth = depth;
  }
  return true;
}

// sphere_cylinder-inl.h:96-202
template <typename S>
FCLB_DI bool sphereCylinderIntersect(S r_s, const Pose<S>& X_FS, S radius, S height, const Pose<S>& X_FC, bool want,
                                     ContactPt<S>& c) {
  const V3<S> p_CS = originInFrame(X_FC, X_FS);
  V3<S> p_CN = p_CS;
  bool S_is_outside = false;
  const S half_h = height / S(2);
  if (p_CS.z > half_h) {
    S_is_outside = true;
    p_CN.z = half_h;
  } else if (p_CS.z < -half_h) {
    S_is_outside = true;
    p_CN.z = -half_h;
  }
  const S sq_xy = p_CS.x * p_CS.x + p_CS.y * p_CS.y;
  if (sq_xy > radius * radius) {
    S_is_outside = true;
    const S k = S(double(radius) / sqrt(double(sq_xy)));  // unqualified sqrt(): see fclb_primitives.cuh
    p_CN.x = p_CS.x * k;
    p_CN.y = p_CS.y * k;
  }
  const V3<S> p_SN_C = p_CN - p_CS;
  const S sq = sqnorm(p_SN_C);
  if (sq > r_s * r_s) return false;
  if (want) {
    const S eps = eps16<S>::value();
    S depth;
    V3<S> n_SC_C, p_CP;
    if (S_is_outside && sq > eps * eps) {
      const S d_NS = fsqrt(sq);
      n_SC_C = p_SN_C / d_NS;
      depth = r_s - d_NS;
      p_CP = p_CN + n_SC_C * (depth * S(0.5));
    } else {
      const S h = height;
      const S face_distance = (p_CS.z >= 0) ? (h / S(2) - p_CS.z) : (p_CS.z + h / S(2));
      const S d_CS_xy = fsqrt(p_CS.x * p_CS.x + p_CS.y * p_CS.y);
      const S barrel_distance = radius - d_CS_xy;
      if (barrel_distance < face_distance - eps) {
        if (d_CS_xy > eps) {
          n_SC_C = mk<S>(-p_CS.x / d_CS_xy, -p_CS.y / d_CS_xy, S(0));
          depth = r_s + barrel_distance;
          p_CP = p_CS + n_SC_C * ((r_s - barrel_distance) / S(2));
        } else {
          n_SC_C = mk<S>(S(-1), S(0), S(0));
          depth = r_s + radius;
          p_CP = p_CS + n_SC_C * ((r_s - barrel_distance) / S(2));
        }
      } else {
        n_SC_C = mk<S>(S(0), S(0), (p_CS.z >= 0) ? S(-1) : S(1));
        depth = face_distance + r_s;
        p_CP = p_CS + n_SC_C * ((r_s - face_distance) / S(2));
      }
    }
    c.normal = mulMV(X_FC.R, n_SC_C);
    c.pos = apply(X_FC, p_CP);
    c.depth = depth;
  }
  return true;
}

}  // namespace fclb
