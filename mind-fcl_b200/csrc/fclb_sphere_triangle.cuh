// fclb_sphere_triangle.cuh -- sphereTriangleIntersect (narrowphase/detail/primitive_shape_algorithm/
// sphere_triangle-inl.h:50-186): the boolean used by the mesh-shape leaf stage and the contact-generating form used by
// the leaf batch of DefaultGJK_EPA requests (ShapeTransformedTriangleIntersectIndepImpl<S, Sphere<S>>,
// gjk_solver-inl.h:570-581).  The triangle is given in the world frame.
#pragma once
#include "fclb_mpr.cuh"  // numeric_eps
#include "fclb_primitives_intersect.cuh"

namespace fclb {

// ---- sphere_triangle-inl.h:50-186 (boolean part) ----
template <typename S>
FCLB_DI S segmentSqrDistance(const V3<S>& from, const V3<S>& to, const V3<S>& p, V3<S>& nearest) {
  V3<S> diff = p - from;
  const V3<S> v = to - from;
  S t = dot(v, diff);
  if (t > 0) {
    const S dotVV = dot(v, v);
    if (t < dotVV) {
      t /= dotVV;
      diff = diff - v * t;
    } else {
      t = 1;
      diff = diff - v;
    }
  } else {
    t = 0;
  }
  nearest = from + v * t;
  return dot(diff, diff);
}
template <typename S>
FCLB_DI bool projectInTriangle(const V3<S>& p1, const V3<S>& p2, const V3<S>& p3, const V3<S>& normal, const V3<S>& p) {
  const V3<S> edge1 = p2 - p1, edge2 = p3 - p2, edge3 = p1 - p3;
  const V3<S> p1_to_p = p - p1, p2_to_p = p - p2, p3_to_p = p - p3;
  const S r1 = dot(cross(edge1, normal), p1_to_p);
  const S r2 = dot(cross(edge2, normal), p2_to_p);
  const S r3 = dot(cross(edge3, normal), p3_to_p);
  return (r1 > 0 && r2 > 0 && r3 > 0) || (r1 <= 0 && r2 <= 0 && r3 <= 0);
}
template <typename S>
FCLB_DI bool sphereTriangleIntersect(S radius, const V3<S>& center, const V3<S>& P1, const V3<S>& P2, const V3<S>& P3) {
  V3<S> normal = normalized(cross(P2 - P1, P3 - P1));
  const S radius_with_threshold = radius + numeric_eps<S>::value();
  const V3<S> p1_to_center = center - P1;
  S distance_from_plane = dot(p1_to_center, normal);
  if (distance_from_plane < 0) {
    distance_from_plane *= -1;
    normal = normal * S(-1);
  }
  const bool is_inside_contact_plane = (distance_from_plane < radius_with_threshold);
  bool has_contact = false;
  V3<S> contact_point = zero3<S>();
  if (is_inside_contact_plane) {
    if (projectInTriangle(P1, P2, P3, normal, center)) {
      has_contact = true;
      contact_point = center - normal * distance_from_plane;
    } else {
      const S contact_capsule_radius_sqr = radius_with_threshold * radius_with_threshold;
      V3<S> nearest_on_edge;
      S distance_sqr = segmentSqrDistance(P1, P2, center, nearest_on_edge);
      if (distance_sqr < contact_capsule_radius_sqr) {
        has_contact = true;
        contact_point = nearest_on_edge;
      }
      distance_sqr = segmentSqrDistance(P2, P3, center, nearest_on_edge);
      if (distance_sqr < contact_capsule_radius_sqr) {
        has_contact = true;
        contact_point = nearest_on_edge;
      }
      distance_sqr = segmentSqrDistance(P3, P1, center, nearest_on_edge);
      if (distance_sqr < contact_capsule_radius_sqr) {
        has_contact = true;
        contact_point = nearest_on_edge;
      }
    }
  }
  if (has_contact) {
    const V3<S> contact_to_center = contact_point - center;
    const S distance_sqr = sqnorm(contact_to_center);
    if (distance_sqr < radius_with_threshold * radius_with_threshold) return true;
  }
  return false;
}

// the same decision with the reference's contact outputs (:169-184): normal = (contact - center) normalised (or the
// flipped plane normal when the contact is the centre), position = the contact point, depth = -(radius - distance)
// -- negative, as the reference writes it.
template <typename S>
FCLB_DI bool sphereTriangleContact(S radius, const V3<S>& center, const V3<S>& P1, const V3<S>& P2, const V3<S>& P3,
                                   ContactPt<S>& cp) {
  V3<S> normal = normalized(cross(P2 - P1, P3 - P1));
  const S radius_with_threshold = radius + numeric_eps<S>::value();
  const V3<S> p1_to_center = center - P1;
  S distance_from_plane = dot(p1_to_center, normal);
  if (distance_from_plane < 0) {
    distance_from_plane *= -1;
    normal = normal * S(-1);
  }
  const bool is_inside_contact_plane = (distance_from_plane < radius_with_threshold);
  bool has_contact = false;
  V3<S> contact_point = zero3<S>();
  if (is_inside_contact_plane) {
    if (projectInTriangle(P1, P2, P3, normal, center)) {
      has_contact = true;
      contact_point = center - normal * distance_from_plane;
    } else {
      const S contact_capsule_radius_sqr = radius_with_threshold * radius_with_threshold;
      V3<S> nearest_on_edge;
      S distance_sqr = segmentSqrDistance(P1, P2, center, nearest_on_edge);
      if (distance_sqr < contact_capsule_radius_sqr) {
        has_contact = true;
        contact_point = nearest_on_edge;
      }
      distance_sqr = segmentSqrDistance(P2, P3, center, nearest_on_edge);
      if (distance_sqr < contact_capsule_radius_sqr) {
        has_contact = true;
        contact_point = nearest_on_edge;
      }
      distance_sqr = segmentSqrDistance(P3, P1, center, nearest_on_edge);
      if (distance_sqr < contact_capsule_radius_sqr) {
        has_contact = true;
        contact_point = nearest_on_edge;
      }
    }
  }
  if (!has_contact) return false;
  const V3<S> contact_to_center = contact_point - center;
  const S distance_sqr = sqnorm(contact_to_center);
  if (!(distance_sqr < radius_with_threshold * radius_with_threshold)) return false;
  cp.pos = contact_point;
  if (distance_sqr > 0) {
    const S distance = fsqrt(distance_sqr);
    cp.normal = normalized(contact_to_center);
    cp.depth = -(radius - distance);
  } else {
    cp.normal = -normal;
    cp.depth = -radius;
  }
  return true;
}

}  // namespace fclb
