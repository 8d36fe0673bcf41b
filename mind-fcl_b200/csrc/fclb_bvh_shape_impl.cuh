// fclb_bvh_shape_impl.cuh -- batched mesh-shape collide: BVHModel<OBBRSS> vs one
// convex shape per query, ONE WARP PER QUERY.
//
// Reference path (results contract):
//   fcl::collide(BVH, tf1, Shape, tf2) -> BVHShapeCollider<OBBRSS, Shape>
//     -> orientedBVHShapeCollide (collision_func_matrix-inl.h:390-408)
//     -> OrientedNodeBVHSolver::MeshShapeIntersect (traversal/collision/bvh_solver-inl.h:8-72)
//   shape BV: computeBV<OBBRSS, Shape> = fit(getBoundVertices(tf2)) (geometry/shape/utility-inl.h:62-69
//     -> math/bv/utility-inl.h:133-146 fitn: covariance, eigen_old, axisFromEigen, extent/centre)
//   node test: overlap(tf1.R, tf1.t, shape_bv, node_bv) -> OBB only (math/bv/OBB-inl.h:305-436)
//   leaf: ShapeSimplexIntersect (shape_pair_intersect-inl.h:133-198) -> GJKSolver::shapeTriangleIntersect
//     (gjk_solver-inl.h:540-600): Sphere -> sphereTriangleIntersect (sphere_triangle-inl.h:108-186),
//     Box -> boxTriangleIntersect (box_triangle-inl.h:8-150), every other shape -> MPR on the
//     Minkowski difference (shape, triangle) (gjk_solver-inl.h:452-477); MPR "Failed" counts as no hit.
// The reference walks one node at a time from a std::stack; the number of
// intersecting triangles does not depend on the visiting order, so the warp pops
// up to 32 nodes per step (one per lane), compacts children / leaves with ballots
// and runs the leaf routine on batches of 32 triangles.
#pragma once
#include "fclb_bound.h"
#include "fclb_bvh.cuh"
#include "fclb_bvh_build.h"
#include "fclb_leafcand.cuh"
#include "fclb_mpr.cuh"
#include "fclb_sphere_triangle.cuh"

// 3 CTAs per SM (80 registers, no spills): 26.1 -> 25.4 ms on C4
#ifndef FCLB_SCENE_MIN_BLOCKS
#define FCLB_SCENE_MIN_BLOCKS 3
#endif

namespace fclb {

// ---- fitn on the transformed bound vertices (math/bv/utility-inl.h:133-146) ----
template <typename S, typename PointFn>
FCLB_DI NodeD<S> fitObbPoints(int n, PointFn pt) {
  S S1[3] = {0, 0, 0};
  S c00 = 0, c11 = 0, c22 = 0, c01 = 0, c02 = 0, c12 = 0;
  #pragma unroll 1
  for (int i = 0; i < n; i++) {  // getCovariance, point-cloud branch (math/geometry-inl.h:757-768)
    const V3<S> p = pt(i);
    S1[0] += p.x;
    S1[1] += p.y;
    S1[2] += p.z;
    c00 += (p.x * p.x);
    c11 += (p.y * p.y);
    c22 += (p.z * p.z);
    c01 += (p.x * p.y);
    c02 += (p.x * p.z);
    c12 += (p.y * p.z);
  }
  const int n_points = n;
  S M[3][3];
  M[0][0] = c00 - S1[0] * S1[0] / n_points;
  M[1][1] = c11 - S1[1] * S1[1] / n_points;
  M[2][2] = c22 - S1[2] * S1[2] / n_points;
  M[0][1] = c01 - S1[0] * S1[1] / n_points;
  M[1][2] = c12 - S1[1] * S1[2] / n_points;
  M[0][2] = c02 - S1[0] * S1[2] / n_points;
  M[1][0] = M[0][1];
  M[2][0] = M[0][2];
  M[2][1] = M[1][2];
  S d[3] = {0, 0, 0}, vec[3][3];
  if (!hostbuild::jacobi3<S>(M, d, vec)) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) vec[i][j] = (i == j) ? S(1) : S(0);
  }
  NodeD<S> bv;
  hostbuild::axesFromEigen<S>(vec, d, bv.axis.m);
  // getExtentAndCenter_pointcloud (math/geometry-inl.h:116-144)
  const S big = sizeof(S) == 4 ? S(3.402823466e+38f) : S(1.7976931348623157e+308);
  V3<S> mn = mk<S>(big, big, big), mx = mk<S>(-big, -big, -big);
  const V3<S> a0 = col(bv.axis, 0), a1 = col(bv.axis, 1), a2 = col(bv.axis, 2);
  #pragma unroll 1
  for (int i = 0; i < n; i++) {
    const V3<S> p = pt(i);
    const V3<S> proj = mk<S>(dot(a0, p), dot(a1, p), dot(a2, p));
    if (proj.x > mx.x) mx.x = proj.x;
    if (proj.x < mn.x) mn.x = proj.x;
    if (proj.y > mx.y) mx.y = proj.y;
    if (proj.y < mn.y) mn.y = proj.y;
    if (proj.z > mx.z) mx.z = proj.z;
    if (proj.z < mn.z) mn.z = proj.z;
  }
  const V3<S> o = mk<S>((mx.x + mn.x) / 2, (mx.y + mn.y) / 2, (mx.z + mn.z) / 2);
  bv.To = mulMV(bv.axis, o);
  bv.extent = mk<S>((mx.x - mn.x) * S(0.5), (mx.y - mn.y) * S(0.5), (mx.z - mn.z) * S(0.5));
  bv.first_child = -1;
  return bv;
}

// The same fit, warp-cooperative and bit-identical: the points are staged in shared memory (each lane
// transforms every 32nd point), the nine covariance sums are accumulated by nine lanes, each in the
// reference's sequential order, the Jacobi solve runs redundantly on every lane, and the extents -- pure
// min / max, hence order-free -- are reduced across the lanes.  `pts`: 3 * n S of this warp's scratch.
constexpr int kFitMaxPoints = 256;
template <typename S, typename PointFn>
FCLB_DI NodeD<S> fitObbPointsWarp(int n, PointFn pt, S* pts, int lane) {
  #pragma unroll 1
  for (int i = lane; i < n; i += 32) {
    const V3<S> p = pt(i);
    pts[3 * i] = p.x;
    pts[3 * i + 1] = p.y;
    pts[3 * i + 2] = p.z;
  }
  __syncwarp();
  // lane k accumulates sum k: 0..2 = S1[x,y,z]; 3 = xx, 4 = yy, 5 = zz, 6 = xy, 7 = xz, 8 = yz
  S acc = S(0);
  if (lane < 9) {
    const int a = lane < 3 ? lane : (lane == 3 ? 0 : (lane == 4 ? 1 : (lane == 5 ? 2 : (lane == 8 ? 1 : 0))));
    const int b = lane == 3 ? 0 : (lane == 4 ? 1 : (lane == 5 ? 2 : (lane == 6 ? 1 : 2)));
    if (lane < 3) {
      #pragma unroll 1
      for (int i = 0; i < n; i++) acc += pts[3 * i + a];
    } else {
      #pragma unroll 1
      for (int i = 0; i < n; i++) acc += (pts[3 * i + a] * pts[3 * i + b]);
    }
  }
  S sums[9];
#pragma unroll
  for (int k = 0; k < 9; k++) sums[k] = __shfl_sync(0xffffffffu, acc, k);
  const int n_points = n;
  S M[3][3];
  M[0][0] = sums[3] - sums[0] * sums[0] / n_points;
  M[1][1] = sums[4] - sums[1] * sums[1] / n_points;
  M[2][2] = sums[5] - sums[2] * sums[2] / n_points;
  M[0][1] = sums[6] - sums[0] * sums[1] / n_points;
  M[1][2] = sums[8] - sums[1] * sums[2] / n_points;
  M[0][2] = sums[7] - sums[0] * sums[2] / n_points;
  M[1][0] = M[0][1];
  M[2][0] = M[0][2];
  M[2][1] = M[1][2];
  S d[3] = {0, 0, 0}, vec[3][3];
  if (!hostbuild::jacobi3<S>(M, d, vec)) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) vec[i][j] = (i == j) ? S(1) : S(0);
  }
  NodeD<S> bv;
  hostbuild::axesFromEigen<S>(vec, d, bv.axis.m);
  const S big = sizeof(S) == 4 ? S(3.402823466e+38f) : S(1.7976931348623157e+308);
  V3<S> mn = mk<S>(big, big, big), mx = mk<S>(-big, -big, -big);
  const V3<S> a0 = col(bv.axis, 0), a1 = col(bv.axis, 1), a2 = col(bv.axis, 2);
  #pragma unroll 1
  for (int i = lane; i < n; i += 32) {
    const V3<S> p = mk<S>(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    const V3<S> proj = mk<S>(dot(a0, p), dot(a1, p), dot(a2, p));
    if (proj.x > mx.x) mx.x = proj.x;
    if (proj.x < mn.x) mn.x = proj.x;
    if (proj.y > mx.y) mx.y = proj.y;
    if (proj.y < mn.y) mn.y = proj.y;
    if (proj.z > mx.z) mx.z = proj.z;
    if (proj.z < mn.z) mn.z = proj.z;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    mx.x = fmax_(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, off));
    mx.y = fmax_(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, off));
    mx.z = fmax_(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, off));
    mn.x = fmin_(mn.x, __shfl_xor_sync(0xffffffffu, mn.x, off));
    mn.y = fmin_(mn.y, __shfl_xor_sync(0xffffffffu, mn.y, off));
    mn.z = fmin_(mn.z, __shfl_xor_sync(0xffffffffu, mn.z, off));
  }
  __syncwarp();
  const V3<S> o = mk<S>((mx.x + mn.x) / 2, (mx.y + mn.y) / 2, (mx.z + mn.z) / 2);
  bv.To = mulMV(bv.axis, o);
  bv.extent = mk<S>((mx.x - mn.x) * S(0.5), (mx.y - mn.y) * S(0.5), (mx.z - mn.z) * S(0.5));
  bv.first_child = -1;
  return bv;
}

// computeBV<OBBRSS<S>, Shape>(shape, tf, bv), OBB half
template <typename S>
FCLB_DI NodeD<S> shapeWorldObb(const ShapeInst<S>& sh, const BoundD<S>* __restrict__ bound, const Pose<S>& tf, S* pts,
                               int lane) {
  if (sh.type == ST_CONVEX) {
    const ConvexD<S>& c = *sh.cvx;
    if (c.n_verts <= kFitMaxPoints)
      return fitObbPointsWarp<S>(c.n_verts, [&](int i) { return apply(tf, loadVert(c.verts, i)); }, pts, lane);
    return fitObbPoints<S>(c.n_verts, [&](int i) { return apply(tf, loadVert(c.verts, i)); });
  }
  return fitObbPointsWarp<S>(bound->n, [&](int i) { return apply(tf, loadVert3(bound->v, i)); }, pts, lane);
}

// ---- box_triangle-inl.h:8-150 ----
template <typename S>
FCLB_DI bool boxTriSeparatedByEdgeAxes(const V3<S>& h, const V3<S>& e0, const V3<S>& v0, const V3<S>& v1) {
  const S ax = fabs_(e0.x), ay = fabs_(e0.y), az = fabs_(e0.z);
#define FCLB_BT_AXIS(RAD, D0, D1)       \
  {                                     \
    const S r = (RAD);                  \
    const S d0 = (D0), d1 = (D1);       \
    if (d0 < d1) {                      \
      if (d0 > r) return true;          \
      if (d1 < -r) return true;         \
    } else {                            \
      if (d1 > r) return true;          \
      if (d0 < -r) return true;         \
    }                                   \
  }
  FCLB_BT_AXIS(h.y * az + h.z * ay, v0.y * e0.z - v0.z * e0.y, v1.y * e0.z - v1.z * e0.y)
  FCLB_BT_AXIS(h.x * az + h.z * ax, v0.x * e0.z - v0.z * e0.x, v1.x * e0.z - v1.z * e0.x)
  FCLB_BT_AXIS(h.x * ay + h.y * ax, v0.x * e0.y - v0.y * e0.x, v1.x * e0.y - v1.y * e0.x)
#undef FCLB_BT_AXIS
  return false;
}
template <typename S>
FCLB_DI void findMinMax3(S a, S b, S c, S& mn, S& mx) {
  mn = mx = a;
  if (b < mn) mn = b;
  if (b > mx) mx = b;
  if (c < mn) mn = c;
  if (c > mx) mx = c;
}
template <typename S>
FCLB_DI bool boxTriangleOverlap(const V3<S>& h, const V3<S>& v0, const V3<S>& v1, const V3<S>& v2) {
  S mn, mx;
  findMinMax3(v0.x, v1.x, v2.x, mn, mx);
  if (mx < -h.x || mn > h.x) return false;
  findMinMax3(v0.y, v1.y, v2.y, mn, mx);
  if (mx < -h.y || mn > h.y) return false;
  findMinMax3(v0.z, v1.z, v2.z, mn, mx);
  if (mx < -h.z || mn > h.z) return false;
  const V3<S> e0 = v1 - v0;
  if (boxTriSeparatedByEdgeAxes(h, e0, v0, v2)) return false;
  const V3<S> e1 = v2 - v1;
  if (boxTriSeparatedByEdgeAxes(h, e1, v1, v0)) return false;
  const V3<S> e2 = v0 - v2;
  if (boxTriSeparatedByEdgeAxes(h, e2, v0, v1)) return false;
  const V3<S> nrm = cross(e0, e1);
  V3<S> p_min, p_max;
#define FCLB_BT_FACE(C)             \
  if (nrm.C > S(0.0)) {             \
    p_min.C = -h.C - v0.C;          \
    p_max.C = h.C - v0.C;           \
  } else {                          \
    p_min.C = h.C - v0.C;           \
    p_max.C = -h.C - v0.C;          \
  }
  FCLB_BT_FACE(x)
  FCLB_BT_FACE(y)
  FCLB_BT_FACE(z)
#undef FCLB_BT_FACE
  return dot(nrm, p_min) <= 0 && dot(nrm, p_max) >= 0;
}

// Per-query constants of the leaf routine.
template <typename S>
struct LeafCtx {
  ShapeInst<S> shape;
  Pose<S> tf_shape, tf_mesh;
  M3<S> toshape1;    // tf_mesh.R^T * tf_shape.R        (gjk_solver-inl.h:466)
  Pose<S> toshape0;  // tf_shape^-1 * tf_mesh           (gjk_solver-inl.h:467; box_triangle-inl.h:138)
  S tol;
  int max_iter;
};

// GJKSolver::shapeTriangleIntersect(s, tf1, P1, P2, P3, tf2, nullptr), boolean
template <typename S, int T0>
FCLB_DI bool shapeTriangleHit(const LeafCtx<S>& c, const V3<S> P[3]) {
  const int type = (T0 == ST_DYNAMIC) ? c.shape.type : T0;
  if (type == ST_SPHERE) {
    return sphereTriangleIntersect(c.shape.p0, c.tf_shape.t, apply(c.tf_mesh, P[0]), apply(c.tf_mesh, P[1]),
                                   apply(c.tf_mesh, P[2]));
  } else if (type == ST_BOX) {
    const V3<S> h = mk<S>(S(0.5) * c.shape.p0, S(0.5) * c.shape.p1, S(0.5) * c.shape.p2);
    return boxTriangleOverlap(h, apply(c.toshape0, P[0]), apply(c.toshape0, P[1]), apply(c.toshape0, P[2]));
  } else {
    MinkDiff<S, T0, ST_TRIANGLE> md;
    md.s0 = c.shape;
    md.s1.type = ST_TRIANGLE;
    md.s1.cvx = nullptr;
    md.s1.tri[0] = P[0];
    md.s1.tri[1] = P[1];
    md.s1.tri[2] = P[2];
    md.toshape1 = c.toshape1;
    md.toshape0 = c.toshape0;
    return mprIntersect<S>(md, c.max_iter, c.tol, nullptr) == MPR_INTERSECT;
  }
}

// pop width of a query that wants only a few contacts; measured on C3 / C4 (B200): 32 -> 4.76 / 23.4 ms,
// 16 -> 5.04 / 25.1 ms, 8 -> 5.85 / 28.7 ms (most of the work is proving the non-colliding queries separate)
// the eager leaf stage (few contacts wanted) runs as soon as this many triangles are queued
#ifndef FCLB_EAGER_LEAF_MIN
#define FCLB_EAGER_LEAF_MIN 1
#endif
#ifndef FCLB_EAGER_WIDTH
#define FCLB_EAGER_WIDTH 32
#endif
constexpr int kBsWarps = kBvhShapeWarps;
constexpr int kBsStackCap = 1024;
constexpr int kBsLeafCap = 64;

template <typename S, int T0>
__global__ void __launch_bounds__(kBsWarps * 32, FCLB_SCENE_MIN_BLOCKS) bvhShapeCollideKernel(BvhShapeArgs a) {
  extern __shared__ __align__(16) int s_bs[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* stack = s_bs + size_t(warp) * (kBsStackCap + kBsLeafCap);
  int* leafq = stack + kBsStackCap;
  S* fit_pts = reinterpret_cast<S*>(s_bs + size_t(kBsWarps) * (kBsStackCap + kBsLeafCap)) + size_t(warp) * 3 * kFitMaxPoints;
  const S* __restrict__ nodes = static_cast<const S*>(a.nodes);
  const S* __restrict__ tris = static_cast<const S*>(a.tris);
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned long long st_bv = 0, st_leaf = 0;

  while (true) {
    unsigned long long q64 = 0;
    if (lane == 0) q64 = atomicAdd(a.work_counter, 1ull);
    q64 = __shfl_sync(0xffffffffu, q64, 0);
    if (q64 >= a.n) break;
    const size_t q = size_t(q64);
    LeafCtx<S> ctx;
    const uint32_t sid = a.shape_ids[q];
    ctx.shape = bindShape(static_cast<const ShapeD<S>*>(a.shapes), static_cast<const ConvexD<S>*>(a.convex), sid);
    ctx.tf_mesh = loadPose(static_cast<const S*>(a.poses_mesh), q);
    ctx.tf_shape = loadPose(static_cast<const S*>(a.poses_shape), q);
    ctx.toshape1 = mulMtM(ctx.tf_mesh.R, ctx.tf_shape.R);
    ctx.toshape0 = compose(inverse(ctx.tf_shape), ctx.tf_mesh);
    ctx.tol = S(a.tol);
    ctx.max_iter = a.max_iter;
    // every lane fits the same OBB (uniform control flow, no shuffles needed)
    const NodeD<S> shape_bv =
        shapeWorldObb(ctx.shape, static_cast<const BoundD<S>*>(a.bound) + sid, ctx.tf_shape, fit_pts, lane);

    uint32_t count = 0;
    int first = -1;
    int sp = 1, nleaf = 0;
    if (lane == 0) stack[0] = 0;
    __syncwarp();
    bool done = (a.max_contacts == 0);

    // few contacts wanted (boolean collide): narrow pops and an eager leaf stage, see fclb_bvh.cu
    const bool eager = a.max_contacts <= 8;
    const int width = eager ? FCLB_EAGER_WIDTH : 32;
    while (!done && (sp > 0 || nleaf > 0)) {
      if (sp > 0 && nleaf < 32) {
        int take = sp < width ? sp : width;
        if (sp + take > kBsStackCap - 64) take = 1;
        int id = -1;
        if (lane < take) id = stack[sp - 1 - lane];
        sp -= take;
        __syncwarp();
        bool expand = false, leaf = false;
        int c0 = 0;
        if (lane < take) {
          const NodeD<S> nd = loadNode(nodes, id);
          st_bv++;
          if (obbOverlap(ctx.tf_mesh.R, ctx.tf_mesh.t, shape_bv, nd)) {
            if (nd.first_child < 0) {
              leaf = true;
              c0 = -(nd.first_child + 1);
            } else {
              expand = true;
              c0 = nd.first_child;
            }
          }
        }
        const unsigned em = __ballot_sync(0xffffffffu, expand);
        const unsigned lm = __ballot_sync(0xffffffffu, leaf);
        if (sp + 2 * __popc(em) > kBsStackCap) {  // deeper than the depth-first head room: report, never corrupt
          if (lane == 0) atomicAdd(&a.stats[2], 1ull);
          done = true;
          expand = false;
        }
        if (expand) {
          const int pos = sp + 2 * __popc(em & lt_mask);
          stack[pos] = c0;
          stack[pos + 1] = c0 + 1;
        }
        if (leaf) leafq[nleaf + __popc(lm & lt_mask)] = c0;
        if (!done) sp += 2 * __popc(em);
        nleaf += __popc(lm);
        __syncwarp();
      }
      if (nleaf >= 32 || (sp == 0 && nleaf > 0) || (eager && nleaf >= FCLB_EAGER_LEAF_MIN)) {
        const int batch = nleaf < 32 ? nleaf : 32;
        bool hit = false;
        int tri_id = -1;
        if (lane < batch) tri_id = leafq[nleaf - 1 - lane];
        if (a.cand.count) {  // candidate mode: the leaf batch decides (ShapeSimplexIntersect with contacts)
          candAppend<S>(a.cand, lane < batch, uint32_t(q), tri_id, -1, nullptr, nullptr);
        } else if (lane < batch) {
          V3<S> P[3];
          loadTri(tris, tri_id, P);
          st_leaf++;
          hit = shapeTriangleHit<S, T0>(ctx, P);
        }
        nleaf -= batch;
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm) {
          if (first < 0) first = __shfl_sync(0xffffffffu, tri_id, __ffs(hm) - 1);
          if (a.out_b1 && hit) {
            const uint32_t slot = count + uint32_t(__popc(hm & lt_mask));
            if (slot < a.max_keep && slot < a.max_contacts) a.out_b1[q * a.max_keep + slot] = tri_id;
          }
          count += uint32_t(__popc(hm));
          if (count >= a.max_contacts) {
            count = a.max_contacts;
            done = true;
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0) {
      a.counts[q] = count;
      if (a.first_tri) a.first_tri[q] = first;
    }
    __syncwarp();
  }
  if (a.stats) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      st_bv += __shfl_xor_sync(0xffffffffu, st_bv, off);
      st_leaf += __shfl_xor_sync(0xffffffffu, st_leaf, off);
    }
    if (lane == 0) {
      atomicAdd(&a.stats[0], st_bv);
      atomicAdd(&a.stats[1], st_leaf);
    }
  }
}

#ifndef FCLB_MESH_SHAPE_PACKED_KERNEL
#define FCLB_MESH_SHAPE_PACKED_KERNEL 0
#endif
#if FCLB_MESH_SHAPE_PACKED_KERNEL
// ---- boolean queries, four per warp ----------------------------------------------------------------
// A boolean mesh-shape query (max_contacts == 1, no contact sink) keeps 8.5 of 32 lanes busy in its node tests and 2.5 in
// its leaf tests (ncu, C4): the frontier of a depth-first walk that stops at the first hit is narrow.  Here a warp runs
// FOUR queries, one per group of 8 lanes, through ONE warp-uniform loop: every trip each group pops up to 8 nodes of its
// own stack, tests them, pushes the children, and tests the triangles it found.  A group that finishes takes the next
// query from the work counter (its OBB fit runs on all 32 lanes, as in the kernel above).  Same node / leaf tests and
// the same answer per query; `first_tri` is a colliding triangle, as above not necessarily the depth-first one.
// MEASURED (B200, C4 mesh-shape launch, same box): 51.9 ms against 24.8 ms for the one-query-per-warp kernel -- a trip now
// lasts as long as its slowest group (an MPR leaf test in one group stalls the node tests of the other three, and every
// refill runs its OBB fit in front of all four), which costs more than the fuller lanes give back.  Compiled out
// (-DFCLB_MESH_SHAPE_PACKED_KERNEL=1 builds it, FCLB_MESH_SHAPE_PACKED=1 selects it); kept as the record of the experiment.
constexpr int kBpGroups = 4, kBpLanes = 8;
constexpr int kBpStackCap = (kBsStackCap + kBsLeafCap) / kBpGroups;  // 272 per group; near the cap a group pops one node per trip

template <typename S, int T0>
__global__ void __launch_bounds__(kBsWarps * 32, FCLB_SCENE_MIN_BLOCKS) bvhShapePackedKernel(BvhShapeArgs a) {
  extern __shared__ __align__(16) int s_bs[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / kBpLanes, li = lane % kBpLanes;
  int* stack = s_bs + size_t(warp) * (kBsStackCap + kBsLeafCap) + grp * kBpStackCap;
  S* fit_pts = reinterpret_cast<S*>(s_bs + size_t(kBsWarps) * (kBsStackCap + kBsLeafCap)) + size_t(warp) * 3 * kFitMaxPoints;
  const S* __restrict__ nodes = static_cast<const S*>(a.nodes);
  const S* __restrict__ tris = static_cast<const S*>(a.tris);
  const unsigned gshift = unsigned(grp * kBpLanes);
  const unsigned lt_in_group = (1u << li) - 1u;
  unsigned long long st_bv = 0, st_leaf = 0;
  bool more = true;

  LeafCtx<S> ctx;
  NodeD<S> shape_bv;
  bool active = false;
  size_t q = 0;
  int sp = 0;
  uint32_t count = 0;
  int first = -1;

  while (true) {
    // ---- refill: every idle group takes one query; the fits run one after the other on the whole warp
    const unsigned idle = __ballot_sync(0xffffffffu, !active);
    if (more && idle) {
      unsigned need = 0;
#pragma unroll
      for (int g = 0; g < kBpGroups; g++)
        if ((idle >> (g * kBpLanes)) & 1u) need |= 1u << g;
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.work_counter, (unsigned long long)__popc(need));
      base = __shfl_sync(0xffffffffu, base, 0);
      int k = 0;
#pragma unroll
      for (int g = 0; g < kBpGroups; g++) {
        if (!((need >> g) & 1u)) continue;
        const unsigned long long q64 = base + k;
        k++;
        if (q64 >= a.n) {
          more = false;
          continue;
        }
        LeafCtx<S> c;
        const uint32_t sid = a.shape_ids[q64];
        c.shape = bindShape(static_cast<const ShapeD<S>*>(a.shapes), static_cast<const ConvexD<S>*>(a.convex), sid);
        c.tf_mesh = loadPose(static_cast<const S*>(a.poses_mesh), size_t(q64));
        c.tf_shape = loadPose(static_cast<const S*>(a.poses_shape), size_t(q64));
        c.toshape1 = mulMtM(c.tf_mesh.R, c.tf_shape.R);
        c.toshape0 = compose(inverse(c.tf_shape), c.tf_mesh);
        c.tol = S(a.tol);
        c.max_iter = a.max_iter;
        const NodeD<S> bv = shapeWorldObb(c.shape, static_cast<const BoundD<S>*>(a.bound) + sid, c.tf_shape, fit_pts, lane);
        __syncwarp();
        if (grp == g) {
          ctx = c;
          shape_bv = bv;
          q = size_t(q64);
          active = true;
          sp = 1;
          count = 0;
          first = -1;
          if (li == 0) stack[0] = 0;
        }
      }
      __syncwarp();
    }
    if (__ballot_sync(0xffffffffu, active) == 0) break;

    // ---- one step of every active group: pop <= 8 nodes, test, push children, collect triangles
    int take = 0;
    if (active) {
      take = sp < kBpLanes ? sp : kBpLanes;
      if (sp + take > kBpStackCap - 2 * kBpLanes) take = 1;
    }
    int id = -1;
    if (li < take) id = stack[sp - 1 - li];
    sp -= take;
    __syncwarp();
    bool expand = false, leaf = false;
    int c0 = 0;
    if (li < take) {
      const NodeD<S> nd = loadNode(nodes, id);
      st_bv++;
      if (obbOverlap(ctx.tf_mesh.R, ctx.tf_mesh.t, shape_bv, nd)) {
        if (nd.first_child < 0) {
          leaf = true;
          c0 = -(nd.first_child + 1);
        } else {
          expand = true;
          c0 = nd.first_child;
        }
      }
    }
    const unsigned em = (__ballot_sync(0xffffffffu, expand) >> gshift) & 0xffu;
    bool overflow = false;
    if (active && sp + 2 * __popc(em) > kBpStackCap) {  // deeper than the head room: report, never corrupt
      overflow = true;
      expand = false;
    }
    if (expand) {
      const int pos = sp + 2 * __popc(em & lt_in_group);
      stack[pos] = c0;
      stack[pos + 1] = c0 + 1;
    }
    if (active && !overflow) sp += 2 * __popc(em);
    // the triangles found in this step are tested right away, by the lanes that found them
    bool hit = false;
    if (leaf) {
      V3<S> P[3];
      loadTri(tris, c0, P);
      st_leaf++;
      hit = shapeTriangleHit<S, T0>(ctx, P);
    }
    const unsigned hm = (__ballot_sync(0xffffffffu, hit) >> gshift) & 0xffu;
    const int hit_tri = __shfl_sync(0xffffffffu, c0, hm ? int(gshift) + __ffs(hm) - 1 : lane);
    bool finished = false;
    if (active) {
      if (hm) {
        first = hit_tri;
        count = 1;
        finished = true;
      } else if (sp == 0 || overflow) {
        finished = true;
      }
      if (overflow && li == 0) atomicAdd(&a.stats[2], 1ull);
    }
    if (finished) {
      if (li == 0) {
        a.counts[q] = count;
        if (a.first_tri) a.first_tri[q] = first;
      }
      active = false;
    }
    __syncwarp();
  }
  if (a.stats) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      st_bv += __shfl_xor_sync(0xffffffffu, st_bv, off);
      st_leaf += __shfl_xor_sync(0xffffffffu, st_leaf, off);
    }
    if (lane == 0) {
      atomicAdd(&a.stats[0], st_bv);
      atomicAdd(&a.stats[1], st_leaf);
    }
  }
}

// FCLB_MESH_SHAPE_PACKED=0 keeps boolean queries on the one-query-per-warp kernel
inline bool meshShapePackedEnabled() {
  static int v = [] {
    const char* e = getenv("FCLB_MESH_SHAPE_PACKED");
    return e ? atoi(e) : 0;
  }();
  return v != 0;
}

#endif

template <typename S>
cudaError_t launchBvhShape(int type0, const BvhShapeArgs& a, int grid, cudaStream_t st) {
  const size_t smem = size_t(kBsWarps) * ((kBsStackCap + kBsLeafCap) * sizeof(int) + 3 * kFitMaxPoints * sizeof(S));
#if FCLB_MESH_SHAPE_PACKED_KERNEL
  const bool packed = meshShapePackedEnabled() && a.max_contacts == 1 && !a.out_b1 && !a.out_box && !a.cand.count && a.stats;
  if (packed) {
#define FCLB_BP_CASE(T)                                                                                          \
  case T: {                                                                                                      \
    cudaError_t e_ = cudaFuncSetAttribute(bvhShapePackedKernel<S, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); \
    if (e_ != cudaSuccess) return e_;                                                                            \
    bvhShapePackedKernel<S, T><<<grid, kBsWarps * 32, smem, st>>>(a);                                            \
    return cudaGetLastError();                                                                                   \
  }
    switch (type0) {
      FCLB_BP_CASE(ST_BOX)
      FCLB_BP_CASE(ST_SPHERE)
      FCLB_BP_CASE(ST_CONVEX)
      default: {
        cudaError_t e_ = cudaFuncSetAttribute(bvhShapePackedKernel<S, ST_DYNAMIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e_ != cudaSuccess) return e_;
        bvhShapePackedKernel<S, ST_DYNAMIC><<<grid, kBsWarps * 32, smem, st>>>(a);
        return cudaGetLastError();
      }
    }
#undef FCLB_BP_CASE
  }
#endif
#define FCLB_BS_CASE(T)                                                                                          \
  case T: {                                                                                                      \
    cudaError_t e_ = cudaFuncSetAttribute(bvhShapeCollideKernel<S, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); \
    if (e_ != cudaSuccess) return e_;                                                                            \
    bvhShapeCollideKernel<S, T><<<grid, kBsWarps * 32, smem, st>>>(a);                                           \
    break;                                                                                                       \
  }
  switch (type0) {
    FCLB_BS_CASE(ST_BOX)
    FCLB_BS_CASE(ST_SPHERE)
    FCLB_BS_CASE(ST_ELLIPSOID)
    FCLB_BS_CASE(ST_CAPSULE)
    FCLB_BS_CASE(ST_CONE)
    FCLB_BS_CASE(ST_CYLINDER)
    FCLB_BS_CASE(ST_CONVEX)
    default: {
      cudaError_t e_ = cudaFuncSetAttribute(bvhShapeCollideKernel<S, ST_DYNAMIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e_ != cudaSuccess) return e_;
      bvhShapeCollideKernel<S, ST_DYNAMIC><<<grid, kBsWarps * 32, smem, st>>>(a);
      break;
    }
  }
#undef FCLB_BS_CASE
  return cudaGetLastError();
}

}  // namespace fclb
