// fclb_shapes.cuh -- device-side shape records and support mappings.
//
// Mirrors (behaviour, not code) the reference's support functions:
//   analytic shapes      include/fcl/cvx_collide/gjk_shape.hpp:36-179
//   Convex               include/fcl/geometry/shape/convex-inl.h:111-281
//   Triangle             include/fcl/geometry/shape/shape_gjk_interface-inl.h:19-42
//   dispatch             include/fcl/narrowphase/detail/gjk_solver_cvx-inl.h:63-108
// Support functions are specialised per shape type at compile time
// (template<int T>) so a (type0,type1)-bucketed kernel carries no switch in its
// inner loop; T = ST_DYNAMIC keeps the run-time switch for mixed buckets.
#pragma once
#include "fclb_math.cuh"

namespace fclb {

// Type codes == FCLB_* in include/fclb200.h == cvx_collide::GJKShapeType order
// for the analytic shapes (gjk_shape.h:12-30).
enum ShapeType : int {
  ST_BOX = 0,
  ST_SPHERE = 1,
  ST_ELLIPSOID = 2,
  ST_CAPSULE = 3,
  ST_CONE = 4,
  ST_CYLINDER = 5,
  ST_CONVEX = 6,
  ST_TRIANGLE = 7,
  ST_COUNT = 8,
  ST_DYNAMIC = -1
};

// Convex<S> as the device sees it: vertices (3 S each), the reference's CSR
// neighbour encoding verbatim (convex.h:219-242), the six axis-extreme seeds
// (convex-inl.h:153-200) and the mean-vertex interior point (convex-inl.h:64-71).
template <typename S>
struct ConvexD {
  const S* verts;     // 4 S per vertex: x, y, z, pad
  const int* nbr;     // the reference's CSR: nbr[v] = offset of {count, neighbours...}
  const int2* vinfo;  // per vertex: (offset of its first neighbour in nbr, neighbour count)
  int n_verts;
  int walk;  // find_extreme_via_neighbors_
  int seed[6];
  S interior[3];
};

template <typename S>
struct ShapeD {
  int type;
  int geom;  // index into the ConvexD table for ST_CONVEX
  S p[3];    // Box: side xyz; Sphere: r; Ellipsoid: radii; Capsule/Cone/Cylinder: r, lz
};

// A shape instance bound for one query (registers only).
template <typename S>
struct ShapeInst {
  int type;
  S p0, p1, p2;
  const ConvexD<S>* cvx;  // ST_CONVEX
  V3<S> tri[3];           // ST_TRIANGLE (vertices in the shape's own frame)
};

#ifndef FCLB_CVX_UNROLL
#define FCLB_CVX_UNROLL 1
#endif
#ifndef FCLB_CVX_OLDWALK
#define FCLB_CVX_OLDWALK 0
#endif
#ifndef FCLB_CVX_EAGER
#define FCLB_CVX_EAGER 1
#endif
// plain 3-S-per-point arrays (bounding vertices of the primitive shapes)
template <typename S>
FCLB_DI V3<S> loadVert3(const S* __restrict__ v, int i) {
  return mk<S>(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
}
// Convex vertices are stored 4 S apart (x, y, z, pad) so that one vertex is one 128-bit load (two for double)
#ifndef FCLB_CVX_LOAD
#define FCLB_CVX_LOAD 1  // 0 one 128-bit __ldg | 1 one plain 128-bit load | 2 three scalar loads
#endif
FCLB_DI V3<float> loadVert(const float* __restrict__ v, int i) {
#if FCLB_CVX_LOAD == 2
  return mk<float>(v[4 * i], v[4 * i + 1], v[4 * i + 2]);
#elif FCLB_CVX_LOAD == 1
  const float4 t = *(reinterpret_cast<const float4*>(v) + i);
  return mk<float>(t.x, t.y, t.z);
#else
  const float4 t = __ldg(reinterpret_cast<const float4*>(v) + i);
  return mk<float>(t.x, t.y, t.z);
#endif
}
FCLB_DI V3<double> loadVert(const double* __restrict__ v, int i) {
#if FCLB_CVX_LOAD == 2
  return mk<double>(v[4 * i], v[4 * i + 1], v[4 * i + 2]);
#elif FCLB_CVX_LOAD == 1
  const double2 a = *(reinterpret_cast<const double2*>(v) + 2 * i);
  const double2 b = *(reinterpret_cast<const double2*>(v) + 2 * i + 1);
  return mk<double>(a.x, a.y, b.x);
#else
  const double2 a = __ldg(reinterpret_cast<const double2*>(v) + 2 * i);
  const double2 b = __ldg(reinterpret_cast<const double2*>(v) + 2 * i + 1);
  return mk<double>(a.x, a.y, b.x);
#endif
}

// convex-inl.h:133-150 : linear scan, first maximum wins (strict >)
template <typename S>
FCLB_DI int convexExtremeNaive(const ConvexD<S>& c, const V3<S>& d) {
  int best = 0;
  S best_v = dot(d, loadVert(c.verts, 0));
#if FCLB_CVX_UNROLL
#pragma unroll 4
#endif
  #pragma unroll 1
  for (int i = 1; i < c.n_verts; i++) {
    const S v = dot(d, loadVert(c.verts, i));
    if (v > best_v) {
      best = i;
      best_v = v;
    }
  }
  return best;
}

// convex-inl.h:202-281 : hill-climb from the best of six cached axis extremes,
// skipping the two most recently visited parents.
template <typename S>
FCLB_DI int convexExtremeWalk(const ConvexD<S>& c, const V3<S>& d) {
  int init = -1;
  S max_dot = S(0);
  // cached directions are +x,-x,+y,-y,+z,-z; this_direction.dot(v_C) with a unit
  // axis evaluates (1*dx + 0*dy) + 0*dz etc.; adding the zero products is exact
  // except for the sign of zero, which the strict > below cannot observe.
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const S comp_k = (k < 2) ? d.x : ((k < 4) ? d.y : d.z);
    const S dv = (k & 1) ? -comp_k : comp_k;
    if (init < 0 || dv > max_dot) {
      init = c.seed[k];
      max_dot = dv;
    }
  }
  int ext = init;
  S ext_v = dot(d, loadVert(c.verts, ext));
  int parent0 = init, parent1 = init;
  bool keep = true;
  while (keep) {
    keep = false;
#if FCLB_CVX_OLDWALK
    int2 span;
    span.x = c.nbr[ext] + 1;
    span.y = c.nbr[span.x - 1];
#else
    const int2 span = __ldg(c.vinfo + ext);  // (offset of the first neighbour in nbr, neighbour count)
#endif
    const int old_ext = ext;
#if FCLB_CVX_UNROLL
#pragma unroll 4
#endif
    #pragma unroll 1
    for (int k = 0; k < span.y; k++) {
#if FCLB_CVX_OLDWALK
      const int nb = c.nbr[span.x + k];
#else
      const int nb = __ldg(c.nbr + span.x + k);
#endif
#if FCLB_CVX_EAGER
      // the neighbour's vertex is fetched whether or not it is one of the two parents: the loads of one vertex's
      // neighbours then do not depend on the walk's decisions and overlap
      const S nv = dot(d, loadVert(c.verts, nb));
      if (nb == parent0 || nb == parent1) continue;
#else
      if (nb == parent0 || nb == parent1) continue;
      const S nv = dot(d, loadVert(c.verts, nb));
#endif
      if (nv > ext_v) {
        parent1 = ext;
        keep = true;
        ext = nb;
        ext_v = nv;
      }
    }
    parent0 = old_ext;
  }
  return ext;
}

template <int T, typename S>
FCLB_DI V3<S> supportT(const ShapeInst<S>& s, const V3<S>& dir) {
  if (T == ST_BOX) {  // gjk_shape.hpp:44-52
    return mk<S>((dir.x > 0) ? (s.p0 / 2) : (-s.p0 / 2), (dir.y > 0) ? (s.p1 / 2) : (-s.p1 / 2),
                 (dir.z > 0) ? (s.p2 / 2) : (-s.p2 / 2));
  } else if (T == ST_SPHERE) {  // gjk_shape.hpp:65-72 (dir must be unit)
    return dir * s.p0;
  } else if (T == ST_ELLIPSOID) {  // gjk_shape.hpp:80-91
    const S a2 = s.p0 * s.p0, b2 = s.p1 * s.p1, c2 = s.p2 * s.p2;
    const V3<S> v = mk<S>(a2 * dir.x, b2 * dir.y, c2 * dir.z);
    const S d = fsqrt(dot(v, dir));
    return v / d;
  } else if (T == ST_CAPSULE) {  // gjk_shape.hpp:104-121
    const S half_h = s.p1 * S(0.5);
    V3<S> pos1 = mk<S>(S(0), S(0), half_h);
    V3<S> pos2 = mk<S>(S(0), S(0), -half_h);
    const V3<S> v = dir * s.p0;
    pos1 = pos1 + v;
    pos2 = pos2 + v;
    return (dot(dir, pos1) > dot(dir, pos2)) ? pos1 : pos2;
  } else if (T == ST_CONE) {  // gjk_shape.hpp:134-158
    const S radius = s.p0, lz = s.p1;
    S zdist = dir.x * dir.x + dir.y * dir.y;
    S len = zdist + dir.z * dir.z;
    zdist = fsqrt(zdist);
    len = fsqrt(len);
    const S half_h = lz * S(0.5);
    const S sin_a = radius / fsqrt(radius * radius + S(4) * half_h * half_h);
    if (dir.z > len * sin_a) {
      return mk<S>(S(0), S(0), half_h);
    } else if (zdist > 0) {
      const S rad = radius / zdist;
      return mk<S>(rad * dir.x, rad * dir.y, -half_h);
    } else {
      return mk<S>(S(0), S(0), -half_h);
    }
  } else if (T == ST_CYLINDER) {  // gjk_shape.hpp:171-185
    const S radius = s.p0, lz = s.p1;
    const S zdist = fsqrt(dir.x * dir.x + dir.y * dir.y);
    const S half_h = lz * S(0.5);
    if (zdist == S(0)) {
      return mk<S>(S(0), S(0), (dir.z > 0) ? half_h : -half_h);
    } else {
      const S d = radius / zdist;
      return mk<S>(d * dir.x, d * dir.y, (dir.z > 0) ? half_h : -half_h);
    }
  } else if (T == ST_CONVEX) {  // convex-inl.h:111-121
    const ConvexD<S>& c = *s.cvx;
    const int i = c.walk ? convexExtremeWalk(c, dir) : convexExtremeNaive(c, dir);
    return loadVert(c.verts, i);
  } else if (T == ST_TRIANGLE) {  // shape_gjk_interface-inl.h:19-42
    const S da = dot(dir, s.tri[0]), db = dot(dir, s.tri[1]), dc = dot(dir, s.tri[2]);
    if (da > db) {
      return (dc > da) ? s.tri[2] : s.tri[0];
    } else {
      return (dc > db) ? s.tri[2] : s.tri[1];
    }
  } else {  // ST_DYNAMIC: gjk_solver_cvx-inl.h:63-88
    switch (s.type) {
      case ST_BOX:
        return supportT<ST_BOX>(s, dir);
      case ST_SPHERE:
        return supportT<ST_SPHERE>(s, dir);
      case ST_ELLIPSOID:
        return supportT<ST_ELLIPSOID>(s, dir);
      case ST_CAPSULE:
        return supportT<ST_CAPSULE>(s, dir);
      case ST_CONE:
        return supportT<ST_CONE>(s, dir);
      case ST_CYLINDER:
        return supportT<ST_CYLINDER>(s, dir);
      case ST_CONVEX:
        return supportT<ST_CONVEX>(s, dir);
      case ST_TRIANGLE:
        return supportT<ST_TRIANGLE>(s, dir);
      default:
        return zero3<S>();
    }
  }
}

// gjk_solver_cvx-inl.h:110-128 : interior point of a shape (its own frame)
template <typename S>
FCLB_DI V3<S> interiorOf(const ShapeInst<S>& s) {
  if (s.type == ST_CONVEX) return mk<S>(s.cvx->interior[0], s.cvx->interior[1], s.cvx->interior[2]);
  if (s.type == ST_TRIANGLE) return ((s.tri[0] + s.tri[1]) + s.tri[2]) / S(3.0);
  return zero3<S>();
}

// tris: the device BVH triangle array (12 S per triangle: 3 x {x, y, z, pad}) behind ST_TRIANGLE entries of a leaf-batch
// table (geom = triangle id), or nullptr when the table holds no triangles.
template <typename S>
FCLB_DI ShapeInst<S> bindShape(const ShapeD<S>* __restrict__ tab, const ConvexD<S>* __restrict__ cvx, uint32_t idx,
                               const S* __restrict__ tris = nullptr) {
  ShapeInst<S> s;
  const ShapeD<S> r = tab[idx];
  s.type = r.type;
  s.p0 = r.p[0];
  s.p1 = r.p[1];
  s.p2 = r.p[2];
  s.cvx = (r.type == ST_CONVEX) ? (cvx + r.geom) : nullptr;
  if (r.type == ST_TRIANGLE && tris) {
    const S* t = tris + size_t(12) * size_t(r.geom);
#pragma unroll
    for (int v = 0; v < 3; v++) s.tri[v] = mk<S>(t[4 * v], t[4 * v + 1], t[4 * v + 2]);
  }
  return s;
}

// Minkowski difference A (-) B expressed in A's frame
// (include/fcl/cvx_collide/minkowski_diff.hpp:17-84).
//   toshape1 = R2^T R1                  rotates a direction from frame 0 into frame 1
//   toshape0 = tf1^-1 * tf2             maps a point of shape 1 into frame 0
// built exactly as gjk_solver-inl.h:79-85 does.
template <typename S, int T0, int T1>
struct MinkDiff {
  ShapeInst<S> s0, s1;
  M3<S> toshape1;
  Pose<S> toshape0;

  FCLB_DI void setPoses(const Pose<S>& tf1, const Pose<S>& tf2) {
    toshape1 = mulMtM(tf2.R, tf1.R);
    toshape0 = compose(inverse(tf1), tf2);
  }
  FCLB_DI V3<S> support0(const V3<S>& d) const { return supportT<T0>(s0, d); }
  FCLB_DI V3<S> support1(const V3<S>& d) const {
    const V3<S> d1 = mulMV(toshape1, d);
    return apply(toshape0, supportT<T1>(s1, d1));
  }
  FCLB_DI V3<S> support(const V3<S>& d) const { return support0(d) - support1(-d); }
  FCLB_DI V3<S> interior() const { return interiorOf(s0) - apply(toshape0, interiorOf(s1)); }
};

}  // namespace fclb
