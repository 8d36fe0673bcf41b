// fclb_bound.h -- the pose-independent part of ShapeBase<S>::getBoundVertices(tf):
// the bounding polytope of a primitive shape in its own frame, in scalar type S.
// The mesh-shape and heightmap-shape traversals fit an OBB to tf * these points
// per query (computeBV<OBBRSS, Shape> -> generic ComputeBVImpl,
// geometry/shape/utility-inl.h:62-69).  Expression and literal types follow the
// reference (a float instantiation forms several of them in double):
//   box-inl.h:91-108, sphere-inl.h:78-103, ellipsoid-inl.h:82-119,
//   capsule-inl.h:99-153, cone-inl.h:74-94, cylinder-inl.h:75-100.
// Convex uses its own vertices (convex-inl.h:97-107) and is not tabulated here.
#pragma once
#include <cmath>
#include <limits>
#include <vector>

#include "../../include/fclb200.h"

namespace fclb {

constexpr int kMaxBound = 36;

template <typename S>
struct BoundD {
  int n;  // 0 => take the Convex's vertex array
  S v[3 * kMaxBound];
};

template <typename S>
inline void boundVertices(uint32_t type, const S p[3], BoundD<S>& out) {
  int n = 0;
  auto put = [&](double x, double y, double z) {  // Vector3<S>(x, y, z): every coefficient is cast to S
    out.v[3 * n] = S(x);
    out.v[3 * n + 1] = S(y);
    out.v[3 * n + 2] = S(z);
    n++;
  };
  switch (type) {
    case FCLB_BOX: {
      const S a = p[0] / 2, b = p[1] / 2, c = p[2] / 2;
      put(a, b, c); put(a, b, -c); put(a, -b, c); put(a, -b, -c);
      put(-a, b, c); put(-a, b, -c); put(-a, -b, c); put(-a, -b, -c);
      break;
    }
    case FCLB_SPHERE: {
      const S radius = p[0];
      const auto m = (1 + std::sqrt(5.0)) / 2.0;
      auto edge_size = radius * 6 / (std::sqrt(27.0) + std::sqrt(15.0));
      auto a = edge_size;
      auto b = m * edge_size;
      put(0, a, b); put(0, -a, b); put(0, a, -b); put(0, -a, -b);
      put(a, b, 0); put(-a, b, 0); put(a, -b, 0); put(-a, -b, 0);
      put(b, 0, a); put(b, 0, -a); put(-b, 0, a); put(-b, 0, -a);
      break;
    }
    case FCLB_ELLIPSOID: {
      const auto phi = (1.0 + std::sqrt(5.0)) / 2.0;
      const auto a = std::sqrt(3.0) / (phi * phi);
      const auto b = phi * a;
      const S A = p[0], B = p[1], C = p[2];
      const auto Aa = A * a, Ab = A * b, Ba = B * a, Bb = B * b, Ca = C * a, Cb = C * b;
      put(0, Ba, Cb); put(0, -Ba, Cb); put(0, Ba, -Cb); put(0, -Ba, -Cb);
      put(Aa, Bb, 0); put(-Aa, Bb, 0); put(Aa, -Bb, 0); put(-Aa, -Bb, 0);
      put(Ab, 0, Ca); put(Ab, 0, -Ca); put(-Ab, 0, Ca); put(-Ab, 0, -Ca);
      break;
    }
    case FCLB_CAPSULE: {
      const S radius = p[0], lz = p[1];
      const auto m = (1 + std::sqrt(5.0)) / 2.0;
      auto hl = lz * 0.5;
      auto edge_size = radius * 6 / (std::sqrt(27.0) + std::sqrt(15.0));
      auto a = edge_size;
      auto b = m * edge_size;
      auto r2 = radius * 2 / std::sqrt(3.0);
      put(0, a, b + hl); put(0, -a, b + hl); put(0, a, -b + hl); put(0, -a, -b + hl);
      put(a, b, hl); put(-a, b, hl); put(a, -b, hl); put(-a, -b, hl);
      put(b, 0, a + hl); put(b, 0, -a + hl); put(-b, 0, a + hl); put(-b, 0, -a + hl);
      put(0, a, b - hl); put(0, -a, b - hl); put(0, a, -b - hl); put(0, -a, -b - hl);
      put(a, b, -hl); put(-a, b, -hl); put(a, -b, -hl); put(-a, -b, -hl);
      put(b, 0, a - hl); put(b, 0, -a - hl); put(-b, 0, a - hl); put(-b, 0, -a - hl);
      auto c = 0.5 * r2;
      auto d = radius;
      put(r2, 0, hl); put(c, d, hl); put(-c, d, hl); put(-r2, 0, hl); put(-c, -d, hl); put(c, -d, hl);
      put(r2, 0, -hl); put(c, d, -hl); put(-c, d, -hl); put(-r2, 0, -hl); put(-c, -d, -hl); put(c, -d, -hl);
      break;
    }
    case FCLB_CONE: {
      const S radius = p[0], lz = p[1];
      auto hl = lz * 0.5;
      auto r2 = radius * 2 / std::sqrt(3.0);
      auto a = 0.5 * r2;
      auto b = radius;
      put(r2, 0, -hl); put(a, b, -hl); put(-a, b, -hl); put(-r2, 0, -hl); put(-a, -b, -hl); put(a, -b, -hl);
      put(0, 0, hl);
      break;
    }
    case FCLB_CYLINDER: {
      const S radius = p[0], lz = p[1];
      auto hl = lz * 0.5;
      auto r2 = radius * 2 / std::sqrt(3.0);
      auto a = 0.5 * r2;
      auto b = radius;
      put(r2, 0, -hl); put(a, b, -hl); put(-a, b, -hl); put(-r2, 0, -hl); put(-a, -b, -hl); put(a, -b, -hl);
      put(r2, 0, hl); put(a, b, hl); put(-a, b, hl); put(-r2, 0, hl); put(-a, -b, hl); put(a, -b, hl);
      break;
    }
    default:
      break;
  }
  out.n = n;
}

// CollisionGeometry::aabb_local / aabb_center / aabb_radius as each shape's
// computeLocalAABB() leaves them (box-inl.h:72-80, sphere-inl.h:54-61, ellipsoid-inl.h:64-71,
// capsule-inl.h:56-64, cone-inl.h:55-63, cylinder-inl.h:56-64, convex-inl.h:77-87);
// read by CollisionObject::computeAABB (narrowphase/collision_object-inl.h:141-154).
template <typename S>
struct LocalAabbD {
  S mn[3], mx[3], center[3], radius;
};

template <typename S>
inline void localAabb(uint32_t type, const S p[3], const double* convex_verts, int n_verts, LocalAabbD<S>& o) {
  S d[3] = {0, 0, 0};
  bool symmetric = true;
  switch (type) {
    case FCLB_BOX:
      for (int k = 0; k < 3; k++) d[k] = S(0.5) * p[k];
      break;
    case FCLB_SPHERE:
      d[0] = d[1] = d[2] = p[0];
      break;
    case FCLB_ELLIPSOID:
      for (int k = 0; k < 3; k++) d[k] = p[k];
      break;
    case FCLB_CAPSULE:
      d[0] = d[1] = p[0];
      d[2] = S(0.5 * p[1] + p[0]);
      break;
    case FCLB_CONE:
    case FCLB_CYLINDER:
      d[0] = d[1] = p[0];
      d[2] = S(0.5 * p[1]);
      break;
    default:
      symmetric = false;
      break;
  }
  if (symmetric) {
    for (int k = 0; k < 3; k++) {
      o.mx[k] = d[k];
      o.mn[k] = -d[k];
    }
  } else {
    const S big = std::numeric_limits<S>::max();
    for (int k = 0; k < 3; k++) {
      o.mn[k] = big;
      o.mx[k] = -big;
    }
    for (int i = 0; i < n_verts; i++)
      for (int k = 0; k < 3; k++) {
        const S v = S(convex_verts[3 * size_t(i) + k]);
        if (v < o.mn[k]) o.mn[k] = v;
        if (v > o.mx[k]) o.mx[k] = v;
      }
  }
  for (int k = 0; k < 3; k++) o.center[k] = (o.mn[k] + o.mx[k]) * S(0.5);
  if (type == FCLB_SPHERE) {
    o.radius = p[0];
  } else {
    const S e0 = o.mn[0] - o.center[0], e1 = o.mn[1] - o.center[1], e2 = o.mn[2] - o.center[2];
    o.radius = std::sqrt((e0 * e0 + e1 * e1) + e2 * e2);
  }
}

}  // namespace fclb
