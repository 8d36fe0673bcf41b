// fclb_scene_gjk_impl.cuh -- contact generation (request.useDefaultPenetration()) for queries that touch a scene
// geometry: the "leaf batch".
//
// Reference: with penetration enabled every leaf of a scene traversal ends in the shape-pair leaf stage
//   mesh - shape        ShapeSimplexIntersect<Shape>(shape, tf_shape, triangle, tf_mesh)   bvh_solver-inl.h:52-66
//                       -> GJKSolver::shapeTriangleIntersect with contacts                  gjk_solver-inl.h:479-531
//                          (Sphere: sphereTriangleIntersect :570-581; every other shape, Box included: GJK + EPA)
//   heightmap - shape   ShapeIntersect<Box, Shape>(pixel box, box_tf, shape, tf_shape)      heightmap_solver_leaf-inl.h:10-31
//   octree - shape      ShapeIntersect<Box, Shape>(voxel box, box_tf, shape, tf_shape)      octree2_solver_leaf-inl.h:22-44
//   box hierarchies     ShapeIntersect<Box, Box>(box1, box1_tf, box2, box2_tf) = boxBox2    heightmap_solver_leaf-inl.h:56-66,
//                                                                                           octree2_solver_leaf-inl.h:85-404
//   box - mesh          ShapeSimplexIntersect<Box>(box, box_tf, triangle, tf_mesh)          heightmap_solver_leaf-inl.h:70-88
// i.e. fcl::collide on a (leaf geometry, shape) pair.  The traversal kernels therefore run in candidate mode (every
// leaf that survives the node culls is appended to a list, fclb_leafcand.cuh), the candidates are ordered by
// (query, b1, b2), turned into one shape-pair batch -- a ShapeD table that holds the caller's shapes followed by one
// Box / Triangle entry per leaf, pairs, poses -- and pushed through the shape-pair collide pipeline
// (collideLeafBatch: bucketing by pair kind, closed forms, GJK boolean, tiered EPA).  A last pass walks each query's
// leaves in id order and writes its contacts the way ShapeIntersect does (free-space clipping with the deepest-first
// partial_sort, reverse_normal / reverse_o1_and_o2 of ContactMeta::writeToContact, shape_pair_intersect-inl.h:19-46).
#pragma once
#include "fclb_collide_impl.cuh"

namespace fclb {

template <typename S>
struct LeafBuildArgs {
  int mode;                  // 0: mesh-shape, 1: box scene (heightmap / octree)-shape, 2: box-box pair, 3: box-mesh pair
  int octree_pair;           // mode 2: both sides are octrees (reverse_tree12 cases exist)
  size_t m;                  // candidates
  size_t q_base;             // query index of the chunk's first query (candidates carry chunk-local indices)
  const uint32_t* order;     // candidate index of item j ((query, b1, b2) order)
  const uint32_t* cq;
  const long long* cb1;
  const long long* cb2;
  const S* box1;
  const S* box2;
  const uint32_t* shape_ids;  // modes 0, 1
  const S* poses_a;           // scene (modes 0, 1) / side 1 (modes 2, 3), 12 S per query of the whole batch
  const S* poses_b;           // shape (modes 0, 1) / side 2 (modes 2, 3)
  uint32_t n_user;            // entries of the caller's shape table at the head of `table`
  ShapeD<S>* table;           // [n_user + 2 m]
  fclb_pair* pairs;           // [m]
  S* p1;                      // [m * 12]
  S* p2;
  uint32_t* item_q;           // [m] chunk-local query of item j
  long long* item_b1;         // [m] Contact::b1 / b2 of the item's contacts
  long long* item_b2;
  uint8_t* item_flags;        // [m] bit 0: reverse the normal
  uint32_t* q_count;          // [n_chunk] items per query
};

// constructBox(aabb, tf, box, box_tf) (geometry/shape/utility-inl.h): side = max - min, box_tf = tf with the
// translation advanced by R * centre
template <typename S>
FCLB_DI void leafBox(const S* b, const Pose<S>& tf, ShapeD<S>& rec, S* pose_out) {
  const V3<S> mn = mk<S>(b[0], b[1], b[2]), mx = mk<S>(b[3], b[4], b[5]);
  const V3<S> side = mx - mn;
  const V3<S> center = (mn + mx) * S(0.5);
  rec.type = ST_BOX;
  rec.geom = 0;
  rec.p[0] = side.x;
  rec.p[1] = side.y;
  rec.p[2] = side.z;
  const V3<S> t = tf.t + mulMV(tf.R, center);
#pragma unroll
  for (int k = 0; k < 9; k++) pose_out[k] = tf.R.m[k];
  pose_out[9] = t.x; pose_out[10] = t.y; pose_out[11] = t.z;
}
template <typename S>
FCLB_DI void copyPose(const S* src, S* dst) {
#pragma unroll
  for (int k = 0; k < 12; k++) dst[k] = src[k];
}

template <typename S>
__global__ void __launch_bounds__(256) leafBatchBuildKernel(LeafBuildArgs<S> a) {
  #pragma unroll 1
  for (size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x; j < a.m; j += size_t(gridDim.x) * blockDim.x) {
    const uint32_t c = a.order[j];
    const uint32_t ql = a.cq[c];
    const size_t q = a.q_base + ql;
    const long long b1 = a.cb1[c], b2 = a.cb2 ? a.cb2[c] : -1;
    a.item_q[j] = ql;
    a.item_b1[j] = b1;
    a.item_b2[j] = b2;
    atomicAdd(&a.q_count[ql], 1u);
    uint8_t flags = 0;
    const uint32_t e1 = a.n_user + uint32_t(2 * j), e2 = e1 + 1;
    ShapeD<S> r1, r2;
    r1.type = r2.type = ST_BOX;
    r1.geom = r2.geom = 0;
    r1.p[0] = r1.p[1] = r1.p[2] = r2.p[0] = r2.p[1] = r2.p[2] = S(0);
    S* o1 = a.p1 + j * 12;
    S* o2 = a.p2 + j * 12;
    const S* pa = a.poses_a + q * 12;
    const S* pb = a.poses_b + q * 12;
    fclb_pair pr;
    if (a.mode == 0) {  // (shape, tf_shape) vs (triangle, tf_mesh); the contact is written with reverse_normal
      r1.type = ST_TRIANGLE;
      r1.geom = int(b1);
      pr.shape1 = a.shape_ids[q];
      pr.shape2 = e1;
      copyPose(pb, o1);
      copyPose(pa, o2);
      flags = 1;
    } else if (a.mode == 1) {  // (box, box_tf) vs (shape, tf_shape)
      leafBox<S>(a.box1 + size_t(c) * 6, loadPose(a.poses_a, q), r1, o1);
      pr.shape1 = e1;
      pr.shape2 = a.shape_ids[q];
      copyPose(pb, o2);
    } else if (a.mode == 2) {  // (box1, box1_tf) vs (box2, box2_tf)
      // octreePairIntersect hands a (leaf-layer node of tree 1, fully occupied inner node of tree 2) pair to
      // octreePairInnerNodeWithLeafNode with the trees swapped and reverse_tree12 = true
      // (octree2_solver_traverse-inl.h:374-382): boxBox2 runs on (box of tree 2, box of tree 1), the normal is
      // reversed and o1 / o2, b1 / b2 are swapped back (octree2_solver_leaf-inl.h:296-297)
      const bool rev = a.octree_pair && ((b1 >> 48) & 1) && !((b2 >> 48) & 1);
      S t1[12], t2[12];
      leafBox<S>(a.box1 + size_t(c) * 6, loadPose(a.poses_a, q), r1, t1);
      leafBox<S>(a.box2 + size_t(c) * 6, loadPose(a.poses_b, q), r2, t2);
      if (rev) {
        const ShapeD<S> tmp = r1;
        r1 = r2;
        r2 = tmp;
        copyPose(t2, o1);
        copyPose(t1, o2);
        flags = 1;
      } else {
        copyPose(t1, o1);
        copyPose(t2, o2);
      }
      pr.shape1 = e1;
      pr.shape2 = e2;
    } else {  // (box, box_tf) vs (triangle, tf_mesh)
      leafBox<S>(a.box1 + size_t(c) * 6, loadPose(a.poses_a, q), r1, o1);
      r2.type = ST_TRIANGLE;
      r2.geom = int(b2);
      pr.shape1 = e1;
      pr.shape2 = e2;
      copyPose(pb, o2);
    }
    a.table[e1] = r1;
    a.table[e2] = r2;
    a.pairs[j] = pr;
    a.item_flags[j] = flags;
  }
}

template <typename S>
struct LeafScatterArgs {
  size_t n_chunk;             // queries of the chunk
  size_t q_base;
  const uint32_t* q_off;      // [n_chunk] first item of the query (exclusive scan of q_count)
  const uint32_t* q_count;
  const long long* item_b1;
  const long long* item_b2;
  const uint8_t* item_flags;
  const S* leaf_contacts;     // [m * 4 * 9] {b1, b2, normal, pos, depth} per contact of the leaf pair
  const uint32_t* leaf_counts;
  uint32_t max_contacts;
  uint32_t max_keep;
  uint32_t* counts;           // outputs, indexed by the query of the whole batch
  long long* out_b1;
  long long* out_b2;          // or nullptr
  S* out_contacts;            // [n * max_keep * 7] normal, pos, depth
};

// One thread per query: the leaves of the query in (b1, b2) order, each adding its contacts as ShapeIntersect does
// (shape_pair_intersect-inl.h:88-115): all of them while they fit, else the `free_space` deepest.
template <typename S>
__global__ void __launch_bounds__(128) leafBatchScatterKernel(LeafScatterArgs<S> a) {
  #pragma unroll 1
  for (size_t ql = blockIdx.x * size_t(blockDim.x) + threadIdx.x; ql < a.n_chunk; ql += size_t(gridDim.x) * blockDim.x) {
    const size_t q = a.q_base + ql;
    uint32_t total = 0;
    const uint32_t first = a.q_off[ql], cnt = a.q_count[ql];
    #pragma unroll 1
    for (uint32_t i = 0; i < cnt && total < a.max_contacts; i++) {
      const uint32_t j = first + i;
      const uint32_t c = a.leaf_counts[j];
      if (c == 0) continue;
      const uint32_t free_space = a.max_contacts - total;
      uint32_t adding = c;
      int pick[4] = {0, 1, 2, 3};
      if (free_space < c) {
        S depth[8];
        SmallPartialSort<S> ps;
        #pragma unroll 1
        for (uint32_t k = 0; k < c; k++) {
          depth[k] = a.leaf_contacts[(size_t(j) * 4 + k) * 9 + 8];
          ps.idx[k] = int(k);
        }
        ps.depth = depth;
        ps.run(int(c), int(free_space));
        adding = free_space;
        #pragma unroll 1
        for (uint32_t k = 0; k < adding; k++) pick[k] = ps.idx[k];
      }
      const S sgn = (a.item_flags[j] & 1) ? S(-1) : S(1);
      #pragma unroll 1
      for (uint32_t k = 0; k < adding; k++) {
        const uint32_t slot = total + k;
        if (slot >= a.max_keep) break;
        const S* r = a.leaf_contacts + (size_t(j) * 4 + pick[k]) * 9;
        const size_t o = q * a.max_keep + slot;
        a.out_b1[o] = a.item_b1[j];
        if (a.out_b2) a.out_b2[o] = a.item_b2[j];
        S* w = a.out_contacts + o * 7;
        w[0] = sgn * r[2]; w[1] = sgn * r[3]; w[2] = sgn * r[4];
        w[3] = r[5]; w[4] = r[6]; w[5] = r[7];
        w[6] = r[8];
      }
      total += adding;
    }
    a.counts[q] = total;
    #pragma unroll 1
    for (uint32_t slot = total; slot < a.max_keep; slot++) {
      const size_t o = q * a.max_keep + slot;
      a.out_b1[o] = -1;
      if (a.out_b2) a.out_b2[o] = -1;
      S* w = a.out_contacts + o * 7;
#pragma unroll
      for (int k = 0; k < 7; k++) w[k] = S(0);
    }
  }
}

// gather helpers of the (query, b1, b2) ordering: keys of the next radix pass in the current order
__global__ void gatherI64Kernel(const long long* __restrict__ src, const uint32_t* __restrict__ idx, size_t m,
                                unsigned long long* __restrict__ dst) {
  // ids are >= -1: bias by one so that the unsigned radix order equals the signed order
  #pragma unroll 1
  for (size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x; j < m; j += size_t(gridDim.x) * blockDim.x)
    dst[j] = static_cast<unsigned long long>(src[idx[j]] + 1);
}
__global__ void gatherU32Kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, size_t m,
                                unsigned long long* __restrict__ dst) {
  #pragma unroll 1
  for (size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x; j < m; j += size_t(gridDim.x) * blockDim.x)
    dst[j] = src[idx[j]];
}
__global__ void iotaKernel(uint32_t* dst, size_t m) {
  #pragma unroll 1
  for (size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x; j < m; j += size_t(gridDim.x) * blockDim.x) dst[j] = uint32_t(j);
}

}  // namespace fclb
