// fclb_ccd_mesh.cuh -- translational continuous collision, shape vs mesh.
//
// Reference: fcl::translational_ccd(shape, tf1, displacement, BVHModel<OBB<S>>, tf2, request, result)
//   -> TranslationalDisplacementShapeBVHSolverImpl<S, Shape, OBB<S>>::RunSweptBV (detail/ccd/bvh_ccd_solver-inl.h:120-218):
//      a depth-first walk over the mesh's OBB tree with a stack of (node, parent time-of-collision interval); a node is
//      culled when BoxPairTranslationalCCD::IsDisjoint(shape OBB, its displacement, node OBB in the world, parent
//      interval) says so, otherwise its children inherit the node's interval; a leaf runs
//      ShapePairTranslationalCollisionSolver::RunShapeSimplex (shape_pair_ccd-inl.h:196-212) with the leaf's interval as
//      `external_box_toc`, which adds one contact (b2 = triangle id, toc) per colliding triangle until
//      num_max_contacts is reached.
//   The mesh-first matrix entry (RunMeshShape, :572-584) is the same walk with the displacement moved into the
//   shape's frame and negated.
//
// Here: (A) one warp per query walks the tree 32 nodes at a time and appends every surviving leaf to a candidate list
// (a leaf's interval depends on its ancestors only, not on the visiting order); (B) one thread per candidate runs the
// swept-volume MPR; (C) the hits of a query are ordered by the leaf's rank in the reference's walk (children are
// pushed left then right, so the right subtree is visited first) and the first max_contacts are reported -- exactly
// the contacts the reference's early-terminating walk returns.
//
// The CCD matrix knows OBB and AABB trees only (translational_collision_func_matrix-inl.h:449-490).  A
// BVHModel<OBB<S>> has the same hierarchy and the same internal boxes as the OBB half of our BVHModel<OBBRSS<S>>; its
// LEAF boxes come from the 3-point fit (BV_fitter-inl.h:174-194 -> OBB_fit_functions::fit3, math/bv/utility-inl.h:85-109)
// instead of the covariance fit, so leaf boxes are re-derived from the triangle here.
#pragma once
#include "fclb_bvh_shape_impl.cuh"
#include "fclb_ccd.cuh"

namespace fclb {

struct CcdMeshArgs {
  const void* nodes;
  const void* tris;
  const void* shapes;   // ShapeD<S>[]
  const void* convex;   // ConvexD<S>[]
  const uint32_t* shape_ids;
  const void* poses_shape;
  const void* poses_mesh;
  const void* disp;     // 4 S per query: unit axis (frame of the moving object), scalar displacement
  size_t n;
  int mesh_moves;       // 1: the displacement is the mesh's (matrix entry [BV_OBB][GEOM_x])
  int request_type;
  double zero_tol, gjk_tol;
  int max_iter;
  // candidate list
  uint32_t* cand_count;  // [0] candidates appended (may exceed the capacity: the caller retries), [1] stack overflows
  uint32_t cand_cap;
  uint32_t* cand_q;
  int32_t* cand_tri;
  void* cand_iv;        // 2 S per candidate: the leaf's interval
  unsigned long long* work_counter;
  // leaf stage
  const int* dfs_rank;  // per triangle: its position in the reference's walk
  unsigned long long* keys;  // per candidate: query << 32 | rank, ~0 when the triangle is not hit
  void* cand_toc;       // 2 S per candidate
};

// OBB_fit_functions::fit3 + getExtentAndCenter_pointcloud on the three vertices
template <typename S>
FCLB_DI void fitObb3(const V3<S> p[3], M3<S>& axis, V3<S>& To, V3<S>& ext) {
  const V3<S> e0 = p[0] - p[1], e1 = p[1] - p[2], e2 = p[2] - p[0];
  const S l0 = sqnorm(e0), l1 = sqnorm(e1), l2 = sqnorm(e2);
  int imax = 0;
  if (l1 > l0) imax = 1;
  if (l2 > (imax == 0 ? l0 : l1)) imax = 2;
  V3<S> c2 = cross(e0, e1);
  {
    const S z = sqnorm(c2);
    if (z > S(0)) {
      const S s = fsqrt(z);
      c2 = mk<S>(c2.x / s, c2.y / s, c2.z / s);
    }
  }
  V3<S> c0 = imax == 0 ? e0 : (imax == 1 ? e1 : e2);
  {
    const S z = sqnorm(c0);
    if (z > S(0)) {
      const S s = fsqrt(z);
      c0 = mk<S>(c0.x / s, c0.y / s, c0.z / s);
    }
  }
  const V3<S> c1 = cross(c2, c0);
  axis.m[0] = c0.x; axis.m[3] = c0.y; axis.m[6] = c0.z;
  axis.m[1] = c1.x; axis.m[4] = c1.y; axis.m[7] = c1.z;
  axis.m[2] = c2.x; axis.m[5] = c2.y; axis.m[8] = c2.z;
  const S big = sizeof(S) == 4 ? S(3.402823466e+38F) : S(1.7976931348623157e+308);
  V3<S> mn = mk<S>(big, big, big), mx = mk<S>(-big, -big, -big);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const V3<S> pr = mk<S>(dot(c0, p[i]), dot(c1, p[i]), dot(c2, p[i]));
    if (pr.x > mx.x) mx.x = pr.x;
    if (pr.x < mn.x) mn.x = pr.x;
    if (pr.y > mx.y) mx.y = pr.y;
    if (pr.y < mn.y) mn.y = pr.y;
    if (pr.z > mx.z) mx.z = pr.z;
    if (pr.z < mn.z) mn.z = pr.z;
  }
  const V3<S> o = mk<S>((mx.x + mn.x) / 2, (mx.y + mn.y) / 2, (mx.z + mn.z) / 2);
  To = mulMV(axis, o);
  ext = mk<S>((mx.x - mn.x) * S(0.5), (mx.y - mn.y) * S(0.5), (mx.z - mn.z) * S(0.5));
}

// ConvertBVImpl<S, OBB<S>, OBB<S>>::run (math/bv/utility-inl.h:612-630): the node box in the world
template <typename S>
FCLB_DI void obbToWorld(const Pose<S>& tf, const M3<S>& axis, const V3<S>& To, M3<S>& axis_w, V3<S>& To_w) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++)
      axis_w.m[3 * i + j] = (tf.R.m[3 * i] * axis.m[j] + tf.R.m[3 * i + 1] * axis.m[3 + j]) + tf.R.m[3 * i + 2] * axis.m[6 + j];
  }
  To_w.x = ((tf.R.m[0] * To.x + tf.R.m[1] * To.y) + tf.R.m[2] * To.z) + tf.t.x;
  To_w.y = ((tf.R.m[3] * To.x + tf.R.m[4] * To.y) + tf.R.m[5] * To.z) + tf.t.y;
  To_w.z = ((tf.R.m[6] * To.x + tf.R.m[7] * To.y) + tf.R.m[8] * To.z) + tf.t.z;
}

// computeBV<OBB<S>, Shape>(shape, tf, bv) (geometry/shape/utility-inl.h:93-246, 780-786); uniform over the warp
template <typename S>
FCLB_DI void shapeObbForCcd(const ShapeInst<S>& sh, const Pose<S>& tf, S* fit_pts, int lane, M3<S>& axis, V3<S>& To, V3<S>& ext) {
  axis = tf.R;
  To = tf.t;
  switch (sh.type) {
    case ST_BOX:
      ext = mk<S>(sh.p0 * S(0.5), sh.p1 * S(0.5), sh.p2 * S(0.5));
      break;
    case ST_SPHERE:
      axis.m[0] = axis.m[4] = axis.m[8] = S(1);
      axis.m[1] = axis.m[2] = axis.m[3] = axis.m[5] = axis.m[6] = axis.m[7] = S(0);
      ext = mk<S>(sh.p0, sh.p0, sh.p0);
      break;
    case ST_ELLIPSOID:
      ext = mk<S>(sh.p0, sh.p1, sh.p2);
      break;
    case ST_CAPSULE:
      ext = mk<S>(sh.p0, sh.p0, sh.p1 / 2 + sh.p0);
      break;
    case ST_CONVEX: {  // fit(vertices) in the shape's frame, then rotated / moved by tf
      const ConvexD<S>& c = *sh.cvx;
      const NodeD<S> f = c.n_verts <= kFitMaxPoints
                             ? fitObbPointsWarp<S>(c.n_verts, [&](int i) { return loadVert(c.verts, i); }, fit_pts, lane)
                             : fitObbPoints<S>(c.n_verts, [&](int i) { return loadVert(c.verts, i); });
      obbToWorld(tf, f.axis, f.To, axis, To);
      ext = f.extent;
      break;
    }
    default:  // ST_CONE, ST_CYLINDER
      ext = mk<S>(sh.p0, sh.p0, sh.p1 / 2);
      break;
  }
}

constexpr int kCmWarps = 4;
constexpr int kCmStackCap = 768;

// (A) the walk: every leaf whose swept-box test passes becomes a candidate with its interval
template <typename S>
__global__ void __launch_bounds__(kCmWarps * 32) ccdMeshTraverseKernel(CcdMeshArgs a) {
  extern __shared__ __align__(16) unsigned char s_cm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* st_id = reinterpret_cast<int*>(s_cm) + size_t(warp) * kCmStackCap;
  S* st_iv = reinterpret_cast<S*>(s_cm + size_t(kCmWarps) * kCmStackCap * sizeof(int)) + size_t(warp) * kCmStackCap * 2;
  S* fit_pts = reinterpret_cast<S*>(s_cm + size_t(kCmWarps) * kCmStackCap * (sizeof(int) + 2 * sizeof(S))) +
               size_t(warp) * 3 * kFitMaxPoints;
  const S* __restrict__ nodes = static_cast<const S*>(a.nodes);
  const S* __restrict__ tris = static_cast<const S*>(a.tris);
  const S* __restrict__ disp = static_cast<const S*>(a.disp);
  const unsigned lt_mask = (1u << lane) - 1u;
  const S zero_tol = S(a.zero_tol);
  while (true) {
    unsigned long long q64 = 0;
    if (lane == 0) q64 = atomicAdd(a.work_counter, 1ull);
    q64 = __shfl_sync(0xffffffffu, q64, 0);
    if (q64 >= a.n) break;
    const size_t q = size_t(q64);
    const uint32_t sid = a.shape_ids[q];
    const ShapeInst<S> sh = bindShape(static_cast<const ShapeD<S>*>(a.shapes), static_cast<const ConvexD<S>*>(a.convex), sid);
    const Pose<S> tf_s = loadPose(static_cast<const S*>(a.poses_shape), q);
    const Pose<S> tf_m = loadPose(static_cast<const S*>(a.poses_mesh), q);
    V3<S> unit = mk<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
    const S scalar = disp[4 * q + 3];
    if (a.mesh_moves) unit = mulMtV(tf_s.R, mulMV(tf_m.R, -unit));  // RunMeshShape (:572-584)
    M3<S> b_axis;
    V3<S> b_To, b_ext;
    shapeObbForCcd(sh, tf_s, fit_pts, lane, b_axis, b_To, b_ext);
    const V3<S> b_unit = mulMtV(b_axis, mulMV(tf_s.R, unit));  // shape_bv_displacement (:139-146)

    int sp = 1;
    if (lane == 0) {
      st_id[0] = 0;
      st_iv[0] = S(0.0);
      st_iv[1] = S(1.0);
    }
    __syncwarp();
    bool overflow = false;
    while (sp > 0) {
      int take = sp < 32 ? sp : 32;
      if (sp + take > kCmStackCap - 64) take = 1;
      int id = -1;
      TocInterval<S> parent;
      parent.lo = parent.hi = S(0);
      if (lane < take) {
        id = st_id[sp - 1 - lane];
        parent.lo = st_iv[2 * (sp - 1 - lane)];
        parent.hi = st_iv[2 * (sp - 1 - lane) + 1];
      }
      sp -= take;
      __syncwarp();
      bool expand = false, leaf = false;
      int c0 = 0;
      TocInterval<S> iv;
      iv.lo = iv.hi = S(0);
      if (lane < take) {
        NodeD<S> nd = loadNode(nodes, id);
        if (nd.first_child < 0) {  // a BVHModel<OBB> leaf box is the 3-point fit of its triangle
          V3<S> P[3];
          loadTri(tris, -(nd.first_child + 1), P);
          fitObb3(P, nd.axis, nd.To, nd.extent);
        }
        M3<S> n_axis;
        V3<S> n_To;
        obbToWorld(tf_m, nd.axis, nd.To, n_axis, n_To);
        iv = parent;
        if (!boxPairCcdDisjoint(b_axis, b_To, b_ext, b_unit, scalar, n_axis, n_To, nd.extent, iv, zero_tol, true)) {
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
      if (sp + 2 * __popc(em) > kCmStackCap) {
        overflow = true;
        expand = false;
      }
      if (expand) {
        const int pos = sp + 2 * __popc(em & lt_mask);
        st_id[pos] = c0;
        st_id[pos + 1] = c0 + 1;
        st_iv[2 * pos] = iv.lo;
        st_iv[2 * pos + 1] = iv.hi;
        st_iv[2 * pos + 2] = iv.lo;
        st_iv[2 * pos + 3] = iv.hi;
      }
      if (!overflow) sp += 2 * __popc(em);
      if (lm) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.cand_count, uint32_t(__popc(lm)));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (leaf) {
          const uint32_t slot = base + uint32_t(__popc(lm & lt_mask));
          if (slot < a.cand_cap) {
            a.cand_q[slot] = uint32_t(q);
            a.cand_tri[slot] = c0;
            S* o = static_cast<S*>(a.cand_iv) + 2 * size_t(slot);
            o[0] = iv.lo;
            o[1] = iv.hi;
          }
        }
      }
      __syncwarp();
      if (overflow) break;
    }
    if (overflow && lane == 0) atomicAdd(a.cand_count + 1, 1u);
    __syncwarp();
  }
}

// (B) RunShapePair<Shape, TriangleP> per candidate (shape_pair_ccd-inl.h:139-170 with the generic RunIntersect :55-112):
// the external interval is valid, so kBoxApproximate runs the binary swept-volume test and reports the leaf's interval
template <typename S>
__global__ void __launch_bounds__(kBlock) ccdMeshLeafKernel(CcdMeshArgs a, uint32_t n_cand) {
  const S* __restrict__ disp = static_cast<const S*>(a.disp);
  const S* __restrict__ tris = static_cast<const S*>(a.tris);
  const S tol = S(a.gjk_tol);
  #pragma unroll 1
  for (size_t c = blockIdx.x * size_t(blockDim.x) + threadIdx.x; c < n_cand; c += size_t(gridDim.x) * blockDim.x) {
    const size_t q = a.cand_q[c];
    const int tri = a.cand_tri[c];
    const uint32_t sid = a.shape_ids[q];
    const Pose<S> tf_s = loadPose(static_cast<const S*>(a.poses_shape), q);
    const Pose<S> tf_m = loadPose(static_cast<const S*>(a.poses_mesh), q);
    V3<S> unit = mk<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
    const S scalar = disp[4 * q + 3];
    if (a.mesh_moves) unit = mulMtV(tf_s.R, mulMV(tf_m.R, -unit));
    SweptMinkDiff<S, ST_DYNAMIC, ST_TRIANGLE> sm;
    sm.md.s0 = bindShape(static_cast<const ShapeD<S>*>(a.shapes), static_cast<const ConvexD<S>*>(a.convex), sid);
    sm.md.s1.type = ST_TRIANGLE;
    sm.md.s1.cvx = nullptr;
    loadTri(tris, tri, sm.md.s1.tri);
    sm.md.setPoses(tf_s, tf_m);
    sm.disp = unit * scalar;
    const S* ivp = static_cast<const S*>(a.cand_iv) + 2 * c;
    TocInterval<S> toc;
    toc.lo = ivp[0];  // ContinuousContactMeta::writeToContact: external_box_toc unless the leaf computes a valid one
    toc.hi = ivp[1];
    bool hit;
    if (a.request_type == CCD_ONE_TOC_SAMPLE) {
      MprIntersectData<S> data;
      hit = mprIntersectData<S>(sm, a.max_iter, tol, data) == MPR_INTERSECT;
      if (hit) {
        const S t = oneTocSample(sm.disp, data);
        if (t >= 0) toc.lo = toc.hi = t;
      }
    } else {
      hit = mprIntersect<S>(sm, a.max_iter, tol, nullptr) == MPR_INTERSECT;
    }
    a.keys[c] = hit ? ((unsigned long long)q << 32) | (unsigned long long)uint32_t(a.dfs_rank[tri]) : ~0ull;
    S* o = static_cast<S*>(a.cand_toc) + 2 * c;
    o[0] = toc.lo;
    o[1] = toc.hi;
  }
}

// (C) after the sort by (query, rank): the first max_contacts hits of every query, in the reference's order
template <typename S>
__global__ void ccdMeshSelectKernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ order,
                                    uint32_t n_cand, const int32_t* __restrict__ cand_tri, const S* __restrict__ cand_toc,
                                    uint32_t max_contacts, uint32_t keep, uint32_t* __restrict__ counts,
                                    long long* __restrict__ prim, S* __restrict__ toc) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n_cand) return;
  const unsigned long long key = keys[i];
  if (key == ~0ull) return;
  const unsigned long long q = key >> 32;
  auto lowerBound = [&](unsigned long long v) {  // first index with keys[j] >= v
    size_t lo = 0, hi = n_cand;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const size_t first = lowerBound(q << 32);
  const size_t k = i - first;
  if (k == 0) {
    const size_t end = lowerBound((q + 1) << 32);
    const size_t cnt = end - first;
    counts[q] = uint32_t(cnt < max_contacts ? cnt : max_contacts);
  }
  if (k < max_contacts && k < keep) {
    const uint32_t c = order[i];
    prim[q * keep + k] = cand_tri[c];
    if (toc) {
      toc[(q * keep + k) * 2] = cand_toc[2 * size_t(c)];
      toc[(q * keep + k) * 2 + 1] = cand_toc[2 * size_t(c) + 1];
    }
  }
}

// ---- mesh vs mesh ------------------------------------------------------------------------------------
// TranslationalDisplacementBVH_PairSolverImpl<S, OBB<S>>::Run (bvh_ccd_solver-inl.h:425-551): the same walk over PAIRS of
// nodes; the first tree is descended when the second node is a leaf, or when both are internal and the first box is
// the larger one (extent.squaredNorm()).  A leaf pair runs RunSimplexPair (shape_pair_ccd-inl.h:214-244): swept-volume
// MPR of (triangle 1 swept, triangle 2).  The reference's visiting order is encoded per candidate as its PATH: one bit
// per descent step (0 = the child popped first, i.e. the right one), most significant bit first, so that sorting the
// hits of a query by path reproduces the order in which the reference's early-terminating walk reports them.
struct CcdMeshPairArgs {
  const void* nodes1;
  const void* tris1;
  const void* nodes2;
  const void* tris2;
  const void* poses1;
  const void* poses2;
  const void* disp;  // 4 S per query: unit axis in mesh 1's frame, scalar displacement
  size_t n;
  int request_type;
  double zero_tol, gjk_tol;
  int max_iter;
  uint32_t* cand_count;  // [0] appended, [1] stack overflows, [2] paths longer than 64 steps
  uint32_t cand_cap;
  uint32_t* cand_q;
  int2* cand_tri;
  void* cand_iv;
  unsigned long long* cand_path;
  unsigned long long* work_counter;
  uint32_t* qkey;       // per candidate: the query, 0xffffffff when the triangles are not hit
  void* cand_toc;
};

constexpr int kCpStackCap = 512;

template <typename S>
struct CcdPairStack {  // one warp's slice
  int2* id;
  S* iv;
  unsigned long long* path;
  unsigned char* depth;
};

template <typename S>
__global__ void __launch_bounds__(kCmWarps * 32) ccdMeshPairTraverseKernel(CcdMeshPairArgs a) {
  extern __shared__ __align__(16) unsigned char s_cp[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long* st_path = reinterpret_cast<unsigned long long*>(s_cp) + size_t(warp) * kCpStackCap;
  int2* st_id = reinterpret_cast<int2*>(s_cp + size_t(kCmWarps) * kCpStackCap * 8) + size_t(warp) * kCpStackCap;
  S* st_iv = reinterpret_cast<S*>(s_cp + size_t(kCmWarps) * kCpStackCap * 16) + size_t(warp) * kCpStackCap * 2;
  unsigned char* st_depth = s_cp + size_t(kCmWarps) * kCpStackCap * (16 + 2 * sizeof(S)) + size_t(warp) * kCpStackCap;
  const S* __restrict__ nodes1 = static_cast<const S*>(a.nodes1);
  const S* __restrict__ nodes2 = static_cast<const S*>(a.nodes2);
  const S* __restrict__ tris1 = static_cast<const S*>(a.tris1);
  const S* __restrict__ tris2 = static_cast<const S*>(a.tris2);
  const S* __restrict__ disp = static_cast<const S*>(a.disp);
  const unsigned lt_mask = (1u << lane) - 1u;
  const S zero_tol = S(a.zero_tol);
  while (true) {
    unsigned long long q64 = 0;
    if (lane == 0) q64 = atomicAdd(a.work_counter, 1ull);
    q64 = __shfl_sync(0xffffffffu, q64, 0);
    if (q64 >= a.n) break;
    const size_t q = size_t(q64);
    const Pose<S> tf1 = loadPose(static_cast<const S*>(a.poses1), q);
    const Pose<S> tf2 = loadPose(static_cast<const S*>(a.poses2), q);
    const V3<S> axis_world = mulMV(tf1.R, mk<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]));
    const S scalar = disp[4 * q + 3];
    int sp = 1;
    if (lane == 0) {
      st_id[0] = make_int2(0, 0);
      st_iv[0] = S(0.0);
      st_iv[1] = S(1.0);
      st_path[0] = 0ull;
      st_depth[0] = 0;
    }
    __syncwarp();
    bool overflow = false;
    while (sp > 0) {
      int take = sp < 32 ? sp : 32;
      if (sp + take > kCpStackCap - 64) take = 1;
      int2 id = make_int2(-1, -1);
      TocInterval<S> iv;
      iv.lo = iv.hi = S(0);
      unsigned long long path = 0;
      int depth = 0;
      if (lane < take) {
        const int e = sp - 1 - lane;
        id = st_id[e];
        iv.lo = st_iv[2 * e];
        iv.hi = st_iv[2 * e + 1];
        path = st_path[e];
        depth = st_depth[e];
      }
      sp -= take;
      __syncwarp();
      bool expand = false, leaf = false, too_deep = false;
      int2 c_first = make_int2(0, 0), c_second = make_int2(0, 0);
      int2 tri = make_int2(0, 0);
      if (lane < take) {
        NodeD<S> n1 = loadNode(nodes1, id.x), n2 = loadNode(nodes2, id.y);
        const bool leaf1 = n1.first_child < 0, leaf2 = n2.first_child < 0;
        const S size1 = sqnorm(n1.extent), size2 = sqnorm(n2.extent);  // (of the stored boxes: compared only when both are internal)
        if (leaf1) {
          V3<S> P[3];
          loadTri(tris1, -(n1.first_child + 1), P);
          fitObb3(P, n1.axis, n1.To, n1.extent);
        }
        if (leaf2) {
          V3<S> P[3];
          loadTri(tris2, -(n2.first_child + 1), P);
          fitObb3(P, n2.axis, n2.To, n2.extent);
        }
        M3<S> a1, a2;
        V3<S> t1, t2;
        obbToWorld(tf1, n1.axis, n1.To, a1, t1);
        obbToWorld(tf2, n2.axis, n2.To, a2, t2);
        const V3<S> unit1 = mulMtV(a1, axis_world);
        if (!boxPairCcdDisjoint(a1, t1, n1.extent, unit1, scalar, a2, t2, n2.extent, iv, zero_tol, true)) {
          if (leaf1 && leaf2) {
            leaf = true;
            tri = make_int2(-(n1.first_child + 1), -(n2.first_child + 1));
          } else if (depth >= 64) {
            too_deep = true;
          } else {
            expand = true;
            const bool on1 = leaf2 || (!leaf1 && size1 > size2);
            // pushed left then right: the right child (second) is popped first
            c_first = on1 ? make_int2(n1.first_child, id.y) : make_int2(id.x, n2.first_child);
            c_second = on1 ? make_int2(n1.first_child + 1, id.y) : make_int2(id.x, n2.first_child + 1);
          }
        }
      }
      const unsigned em = __ballot_sync(0xffffffffu, expand);
      const unsigned lm = __ballot_sync(0xffffffffu, leaf);
      if (__any_sync(0xffffffffu, too_deep)) {
        if (lane == 0) atomicAdd(a.cand_count + 2, 1u);
        overflow = true;
      }
      if (sp + 2 * __popc(em) > kCpStackCap) {
        overflow = true;
        expand = false;
      }
      if (expand && !overflow) {
        const int pos = sp + 2 * __popc(em & lt_mask);
        st_id[pos] = c_first;       // left: visited second -> path bit 1
        st_id[pos + 1] = c_second;  // right: visited first -> path bit 0
        st_iv[2 * pos] = iv.lo;
        st_iv[2 * pos + 1] = iv.hi;
        st_iv[2 * pos + 2] = iv.lo;
        st_iv[2 * pos + 3] = iv.hi;
        st_path[pos] = path | (1ull << (63 - depth));
        st_path[pos + 1] = path;
        st_depth[pos] = st_depth[pos + 1] = (unsigned char)(depth + 1);
      }
      if (!overflow) sp += 2 * __popc(em);
      if (lm) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.cand_count, uint32_t(__popc(lm)));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (leaf) {
          const uint32_t slot = base + uint32_t(__popc(lm & lt_mask));
          if (slot < a.cand_cap) {
            a.cand_q[slot] = uint32_t(q);
            a.cand_tri[slot] = tri;
            S* o = static_cast<S*>(a.cand_iv) + 2 * size_t(slot);
            o[0] = iv.lo;
            o[1] = iv.hi;
            a.cand_path[slot] = path;
          }
        }
      }
      __syncwarp();
      if (overflow) break;
    }
    if (overflow && lane == 0) atomicAdd(a.cand_count + 1, 1u);
    __syncwarp();
  }
}

// RunShapePair<TriangleP, TriangleP> per candidate
template <typename S>
__global__ void __launch_bounds__(kBlock) ccdMeshPairLeafKernel(CcdMeshPairArgs a, uint32_t n_cand) {
  const S* __restrict__ disp = static_cast<const S*>(a.disp);
  const S tol = S(a.gjk_tol);
  #pragma unroll 1
  for (size_t c = blockIdx.x * size_t(blockDim.x) + threadIdx.x; c < n_cand; c += size_t(gridDim.x) * blockDim.x) {
    const size_t q = a.cand_q[c];
    const int2 tri = a.cand_tri[c];
    const Pose<S> tf1 = loadPose(static_cast<const S*>(a.poses1), q);
    const Pose<S> tf2 = loadPose(static_cast<const S*>(a.poses2), q);
    SweptMinkDiff<S, ST_TRIANGLE, ST_TRIANGLE> sm;
    sm.md.s0.type = sm.md.s1.type = ST_TRIANGLE;
    sm.md.s0.cvx = sm.md.s1.cvx = nullptr;
    loadTri(static_cast<const S*>(a.tris1), tri.x, sm.md.s0.tri);
    loadTri(static_cast<const S*>(a.tris2), tri.y, sm.md.s1.tri);
    sm.md.setPoses(tf1, tf2);
    sm.disp = mk<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]) * disp[4 * q + 3];
    const S* ivp = static_cast<const S*>(a.cand_iv) + 2 * c;
    TocInterval<S> toc;
    toc.lo = ivp[0];
    toc.hi = ivp[1];
    bool hit;
    if (a.request_type == CCD_ONE_TOC_SAMPLE) {
      MprIntersectData<S> data;
      hit = mprIntersectData<S>(sm, a.max_iter, tol, data) == MPR_INTERSECT;
      if (hit) {
        const S t = oneTocSample(sm.disp, data);
        if (t >= 0) toc.lo = toc.hi = t;
      }
    } else {
      hit = mprIntersect<S>(sm, a.max_iter, tol, nullptr) == MPR_INTERSECT;
    }
    a.qkey[c] = hit ? uint32_t(q) : 0xffffffffu;
    S* o = static_cast<S*>(a.cand_toc) + 2 * c;
    o[0] = toc.lo;
    o[1] = toc.hi;
  }
}

static __global__ void ccdGatherKeyKernel(const uint32_t* __restrict__ qkey, const uint32_t* __restrict__ order, uint32_t n, uint32_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = qkey[order[i]];
}

// after the stable sorts by path, then by query: the first max_contacts hits of every query
template <typename S>
__global__ void ccdMeshPairSelectKernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ order, uint32_t n_cand,
                                        const int2* __restrict__ cand_tri, const S* __restrict__ cand_toc, uint32_t max_contacts,
                                        uint32_t keep, uint32_t* __restrict__ counts, long long* __restrict__ prim,
                                        S* __restrict__ toc) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n_cand) return;
  const uint32_t q = keys[i];
  if (q == 0xffffffffu) return;
  auto lowerBound = [&](unsigned long long v) {
    size_t lo = 0, hi = n_cand;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if ((unsigned long long)keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const size_t first = lowerBound(q);
  const size_t k = i - first;
  if (k == 0) {
    const size_t cnt = lowerBound((unsigned long long)q + 1) - first;
    counts[q] = uint32_t(cnt < max_contacts ? cnt : max_contacts);
  }
  if (k < max_contacts && k < keep) {
    const uint32_t c = order[i];
    prim[(size_t(q) * keep + k) * 2] = cand_tri[c].x;
    prim[(size_t(q) * keep + k) * 2 + 1] = cand_tri[c].y;
    if (toc) {
      toc[(size_t(q) * keep + k) * 2] = cand_toc[2 * size_t(c)];
      toc[(size_t(q) * keep + k) * 2 + 1] = cand_toc[2 * size_t(c) + 1];
    }
  }
}

}  // namespace fclb
