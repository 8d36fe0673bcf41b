// fclb_scene_pair_impl.cuh -- batched scene-vs-scene collide (heightmap / octree / mesh pairs),
// ONE WARP PER QUERY.
//
// Reference paths (results contract), all boolean requests (no penetration):
//   fcl::collide(HeightMap, tf1, HeightMap, tf2)  -> heightMapPairIntersect
//        (traversal/heightmap/heightmap_solver_traverse-inl.h:189-296)
//   fcl::collide(HeightMap, tf1, BVHModel<OBBRSS>, tf2) -> heightMapBVHIntersect (:298-404)
//   fcl::collide(HeightMap, tf1, Octree2, tf2)    -> heightMapOctreeIntersect (:406-570)
//   fcl::collide(Octree2, tf1, BVHModel<OBBRSS>, tf2) -> octreeBVHIntersect
//        (traversal/octree2/octree2_solver_traverse-inl.h:138-288)
//   fcl::collide(Octree2, tf1, Octree2, tf2)      -> octreePairIntersect (:290-447)
// with the leaf routines
//   boxToBoxProcessLeafPair        (heightmap_solver_leaf-inl.h:33-68): contact iff
//        !FixedRotationBoxDisjoint::isDisjoint(aabb_1, aabb_2, /*strict=*/true)
//        (math/fixed_rotation_obb_disjoint-inl.h:35-187 generic, :219-357 float / SSE association)
//   octreePairTwoLeafNode / InnerNodeWithLeafNode / InnerNodePairAsLeaf
//        (octree2_solver_leaf-inl.h:85-404): the same strict test per voxel-box pair
//   boxToTriangleProcessLeafPair / boxToSimplexProcessLeafPair (heightmap_solver_leaf-inl.h:70-88,
//        octree2_solver_leaf-inl.h:46-66): constructBox(aabb, tf) + shapeTriangleIntersect<Box>
//        = boxTriangleIntersect (box_triangle-inl.h:73-200).
// Node boxes: FlatHeightMap::pixelToBox of a (layer, pixel) node (flat_heightmap-inl.h:163-193,
// zero-height pixels are empty), computeChildAABB for octree children (octree_util-inl.h:10-37);
// fully occupied inner / leaf nodes of an octree are one box, partial 2x2x2 leaves are their voxels.
//
// What the reference REPORTS is the set of leaf pairs that pass the leaf test; the node-pair
// tests above the leaves (6-axis / 15-axis SAT of the enclosing boxes, the sphere approximation
// of octreePairTwoLeafNode) are conservative culls and the descent rule only fixes the visiting
// order.  So the warp runs its own traversal: it pops up to 32 node pairs per step (4 when only a few contacts
// are wanted), culls them with a slack-guarded 15-axis SAT that never removes a pair the leaf test would accept,
// pushes the children of the node with the larger refinable extent (or queues the pair when both are leaves) at
// prefix-sum offsets, and runs the reference's exact leaf test on the queued pairs, 32 at a time.  Contact order
// follows the device traversal (declared).
#pragma once
#include "fclb_bvh_shape_impl.cuh"  // boxTriangleOverlap
#include "fclb_leafcand.cuh"
#include "fclb_octree_impl.cuh"     // FixedRot, makeFixedRot, rowDotAssoc, childAabb

namespace fclb {

// ---- FixedRotationBoxDisjoint::isDisjoint(aabb1, aabb2, strict) ------------------------------
// radius of a cross axis: the generic routine sums its four products left to right
// (fixed_rotation_obb_disjoint-inl.h:103-178), the float / SSE routine adds (ra) + (rb) (:249-282)
FCLB_DI float crossRad(float a1, float a2, float b1, float b2) { return (a1 + a2) + (b1 + b2); }
FCLB_DI double crossRad(double a1, double a2, double b1, double b2) { return ((a1 + a2) + b1) + b2; }

template <typename S>
FCLB_DI bool fixedRotDisjointBoxes(const FixedRot<S>& f, const S* mn1, const S* mx1, const S* mn2, const S* mx2,
                                   bool strict) {
  const V3<S> c1 = mk<S>((mn1[0] + mx1[0]) * S(0.5), (mn1[1] + mx1[1]) * S(0.5), (mn1[2] + mx1[2]) * S(0.5));
  const V3<S> a = mk<S>(S(0.5) * (mx1[0] - mn1[0]), S(0.5) * (mx1[1] - mn1[1]), S(0.5) * (mx1[2] - mn1[2]));
  const V3<S> c2 = mk<S>((mn2[0] + mx2[0]) * S(0.5), (mn2[1] + mx2[1]) * S(0.5), (mn2[2] + mx2[2]) * S(0.5));
  const V3<S> b = mk<S>(S(0.5) * (mx2[0] - mn2[0]), S(0.5) * (mx2[1] - mn2[1]), S(0.5) * (mx2[2] - mn2[2]));
  // b2_in_b1 = rotation_2in1 * b2_center + translation_2in1 - aabb1.center()  (:40-42, :312-331)
  const V3<S> T = mk<S>((rowDotAssoc(f.R, 0, c2) + f.t.x) - c1.x, (rowDotAssoc(f.R, 1, c2) + f.t.y) - c1.y,
                        (rowDotAssoc(f.R, 2, c2) + f.t.z) - c1.z);
  const M3<S>& B = f.R;
  const M3<S>& Bf = f.A;
  if (fabs_(T.x) > a.x + rowDotAssoc(Bf, 0, b)) return true;
  if (fabs_(T.y) > a.y + rowDotAssoc(Bf, 1, b)) return true;
  if (fabs_(T.z) > a.z + rowDotAssoc(Bf, 2, b)) return true;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const S s = dot(col(B, j), T);
    if (fabs_(s) > comp(b, j) + dot(col(Bf, j), a)) return true;
  }
  if (!strict) return false;
#define FCLB_FR_EDGE(SEXPR, A1, A2, B1, B2)                   \
  if (fabs_(SEXPR) > crossRad(A1, A2, B1, B2)) return true;
  FCLB_FR_EDGE(T.z * B(1, 0) - T.y * B(2, 0), a.y * Bf(2, 0), a.z * Bf(1, 0), b.y * Bf(0, 2), b.z * Bf(0, 1))
  FCLB_FR_EDGE(T.z * B(1, 1) - T.y * B(2, 1), a.y * Bf(2, 1), a.z * Bf(1, 1), b.x * Bf(0, 2), b.z * Bf(0, 0))
  FCLB_FR_EDGE(T.z * B(1, 2) - T.y * B(2, 2), a.y * Bf(2, 2), a.z * Bf(1, 2), b.x * Bf(0, 1), b.y * Bf(0, 0))
  FCLB_FR_EDGE(T.x * B(2, 0) - T.z * B(0, 0), a.x * Bf(2, 0), a.z * Bf(0, 0), b.y * Bf(1, 2), b.z * Bf(1, 1))
  FCLB_FR_EDGE(T.x * B(2, 1) - T.z * B(0, 1), a.x * Bf(2, 1), a.z * Bf(0, 1), b.x * Bf(1, 2), b.z * Bf(1, 0))
  FCLB_FR_EDGE(T.x * B(2, 2) - T.z * B(0, 2), a.x * Bf(2, 2), a.z * Bf(0, 2), b.x * Bf(1, 1), b.y * Bf(1, 0))
  FCLB_FR_EDGE(T.y * B(0, 0) - T.x * B(1, 0), a.x * Bf(1, 0), a.y * Bf(0, 0), b.y * Bf(2, 2), b.z * Bf(2, 1))
  FCLB_FR_EDGE(T.y * B(0, 1) - T.x * B(1, 1), a.x * Bf(1, 1), a.y * Bf(0, 1), b.x * Bf(2, 2), b.z * Bf(2, 0))
  FCLB_FR_EDGE(T.y * B(0, 2) - T.x * B(1, 2), a.x * Bf(1, 2), a.y * Bf(0, 2), b.x * Bf(2, 1), b.y * Bf(2, 0))
#undef FCLB_FR_EDGE
  return false;
}

// Node-pair cull of the device traversal: the 15 axes of the OBB separating-axis test (math/bv/OBB-inl.h:319-436,
// the same axes as FixedRotationBoxDisjoint) with a SLACK added to every radius.  The cull must never remove a node
// pair above a leaf pair that the reference's leaf test accepts -- the leaf test, not this cull, decides what is
// reported -- and the plain test is not safe for that: for nearly parallel boxes a cross axis compares a centre
// distance that is rounding noise (~1e-7 |T| in float) with a radius of ~1e-6 * extent.  The slack, 1e-5 x the
// magnitude of the coordinates involved, is far above that noise and far below a pixel / voxel; on the cross axes it
// simply switches the axis off when the two edges are nearly parallel.
template <typename S>
FCLB_DI bool nodeObbDisjoint(const M3<S>& B, const V3<S>& T, const V3<S>& a, const V3<S>& b, S slack) {
  M3<S> Bf;
#pragma unroll
  for (int i = 0; i < 9; i++) Bf.m[i] = fabs_(B.m[i]) + S(1e-6);
  if (fabs_(T.x) > a.x + dot(row(Bf, 0), b) + slack) return true;
  if (fabs_(T.y) > a.y + dot(row(Bf, 1), b) + slack) return true;
  if (fabs_(T.z) > a.z + dot(row(Bf, 2), b) + slack) return true;
#pragma unroll
  for (int j = 0; j < 3; j++)
    if (fabs_(dot(col(B, j), T)) > comp(b, j) + dot(col(Bf, j), a) + slack) return true;
#define FCLB_NO_EDGE(SEXPR, RAD) \
  if (fabs_(SEXPR) > (RAD) + slack) return true;
  FCLB_NO_EDGE(T.z * B(1, 0) - T.y * B(2, 0), a.y * Bf(2, 0) + a.z * Bf(1, 0) + b.y * Bf(0, 2) + b.z * Bf(0, 1))
  FCLB_NO_EDGE(T.z * B(1, 1) - T.y * B(2, 1), a.y * Bf(2, 1) + a.z * Bf(1, 1) + b.x * Bf(0, 2) + b.z * Bf(0, 0))
  FCLB_NO_EDGE(T.z * B(1, 2) - T.y * B(2, 2), a.y * Bf(2, 2) + a.z * Bf(1, 2) + b.x * Bf(0, 1) + b.y * Bf(0, 0))
  FCLB_NO_EDGE(T.x * B(2, 0) - T.z * B(0, 0), a.x * Bf(2, 0) + a.z * Bf(0, 0) + b.y * Bf(1, 2) + b.z * Bf(1, 1))
  FCLB_NO_EDGE(T.x * B(2, 1) - T.z * B(0, 1), a.x * Bf(2, 1) + a.z * Bf(0, 1) + b.x * Bf(1, 2) + b.z * Bf(1, 0))
  FCLB_NO_EDGE(T.x * B(2, 2) - T.z * B(0, 2), a.x * Bf(2, 2) + a.z * Bf(0, 2) + b.x * Bf(1, 1) + b.y * Bf(1, 0))
  FCLB_NO_EDGE(T.y * B(0, 0) - T.x * B(1, 0), a.x * Bf(1, 0) + a.y * Bf(0, 0) + b.y * Bf(2, 2) + b.z * Bf(2, 1))
  FCLB_NO_EDGE(T.y * B(0, 1) - T.x * B(1, 1), a.x * Bf(1, 1) + a.y * Bf(0, 1) + b.x * Bf(2, 2) + b.z * Bf(2, 0))
  FCLB_NO_EDGE(T.y * B(0, 2) - T.x * B(1, 2), a.x * Bf(1, 2) + a.y * Bf(0, 2) + b.x * Bf(2, 1) + b.y * Bf(2, 0))
#undef FCLB_NO_EDGE
  return false;
}

template <typename S>
FCLB_DI bool nodeBoxesDisjoint(const FixedRot<S>& f, const S* mn1, const S* mx1, const S* mn2, const S* mx2) {
  const V3<S> c1 = mk<S>((mn1[0] + mx1[0]) * S(0.5), (mn1[1] + mx1[1]) * S(0.5), (mn1[2] + mx1[2]) * S(0.5));
  const V3<S> a = mk<S>(S(0.5) * (mx1[0] - mn1[0]), S(0.5) * (mx1[1] - mn1[1]), S(0.5) * (mx1[2] - mn1[2]));
  const V3<S> c2 = mk<S>((mn2[0] + mx2[0]) * S(0.5), (mn2[1] + mx2[1]) * S(0.5), (mn2[2] + mx2[2]) * S(0.5));
  const V3<S> b = mk<S>(S(0.5) * (mx2[0] - mn2[0]), S(0.5) * (mx2[1] - mn2[1]), S(0.5) * (mx2[2] - mn2[2]));
  const V3<S> T = (mulMV(f.R, c2) + f.t) - c1;
  const S slack = S(1e-5) * (S(1) + fabs_(f.t.x) + fabs_(f.t.y) + fabs_(f.t.z) + fabs_(c1.x) + fabs_(c1.y) + fabs_(c1.z) +
                             fabs_(c2.x) + fabs_(c2.y) + fabs_(c2.z));
  return nodeObbDisjoint(f.R, T, a, b, slack);
}

// ---- the two box hierarchies -------------------------------------------------------------------
// One traversal element of either hierarchy: the node's box in the scene frame + what is needed to
// expand it and to name it in a contact.
//   meta bit 0      terminal: the box itself is a contact candidate (bottom-layer pixel; fully occupied
//                   inner / leaf node; voxel of a partial leaf)
//        bit 1      octree: leaf-layer node (OctreeTraverseStackElement::is_leaf_node)
//        bits 4..7  octree: voxel index within a partial leaf, 8 = the whole node
//        bits 8..15 heightmap: layer above the bottom (0 = bottom); octree: depth from the root
//   index           heightmap: encodePixel = x << 16 | y; octree: node_vector_index
template <typename S>
struct BoxElem {
  S mn[3], mx[3];
  uint32_t index, meta;
};

template <typename S>
struct SideHm {
  const HmView& v;
  FCLB_DI explicit SideHm(const HmView& view) : v(view) {}
  // FlatHeightMap::pixelToBox(pixel, aabb) of layer k (flat_heightmap-inl.h:163-193); false: empty pixel
  FCLB_DI bool pixelBox(int k, int x, int y, BoxElem<S>& e) const {
    const uint16_t h = v.layers[size_t(v.off[k]) + size_t(y) * v.fx[k] + x];
    if (h == 0) return false;
    const S res_x = S(v.res_x) * S(1 << k), res_y = S(v.res_y) * S(1 << k);
    const int half_x = int(v.half_x >> k), half_y = int(v.half_y >> k);
    const S ccx = S((x - half_x + 0.5) * res_x);
    const S ccy = S((y - half_y + 0.5) * res_y);
    e.mn[0] = ccx - S(0.5) * res_x;
    e.mx[0] = ccx + S(0.5) * res_x;
    e.mn[1] = ccy - S(0.5) * res_y;
    e.mx[1] = ccy + S(0.5) * res_y;
    e.mn[2] = S(0.0);
    e.mx[2] = S(h) * S(0.001);
    e.index = (uint32_t(x) << 16) | uint32_t(y);
    e.meta = (uint32_t(k) << 8) | (k == 0 ? 1u : 0u);
    return true;
  }
  FCLB_DI int numRoots() const { return int(v.fx[v.n_layers - 1]) * int(v.fy[v.n_layers - 1]); }
  FCLB_DI bool root(int i, BoxElem<S>& e) const {
    const int k = v.n_layers - 1;
    return pixelBox(k, i % int(v.fx[k]), i / int(v.fx[k]), e);
  }
  FCLB_DI unsigned childMask(const BoxElem<S>& e) const {
    const int k = int((e.meta >> 8) & 0xffu) - 1;
    const int x = int(e.index >> 16) * 2, y = int(e.index & 0xffffu) * 2;
    const uint16_t* row0 = v.layers + size_t(v.off[k]) + size_t(y) * v.fx[k] + x;
    const uint16_t* row1 = row0 + v.fx[k];
    return (row0[0] ? 1u : 0u) | (row0[1] ? 2u : 0u) | (row1[0] ? 4u : 0u) | (row1[1] ? 8u : 0u);
  }
  FCLB_DI BoxElem<S> child(const BoxElem<S>& e, int c) const {
    const int k = int((e.meta >> 8) & 0xffu) - 1;
    BoxElem<S> ch;
    pixelBox(k, int(e.index >> 16) * 2 + (c & 1), int(e.index & 0xffffu) * 2 + (c >> 1), ch);
    return ch;
  }
  static FCLB_DI long long code(const BoxElem<S>& e) { return (long long)e.index; }  // heightmap_types.h:53-58
};

template <typename S>
struct SideOct {
  const OctView& v;
  FCLB_DI explicit SideOct(const OctView& view) : v(view) {}
  FCLB_DI int numRoots() const { return v.n_inner ? 1 : 0; }
  FCLB_DI bool root(int, BoxElem<S>& e) const {
    if (v.pruned && v.pruned[0]) return false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      e.mn[k] = S(v.root_box[k]);
      e.mx[k] = S(v.root_box[3 + k]);
    }
    e.index = 0;
    e.meta = (8u << 4) | (v.inner_full[0] ? 1u : 0u);
    return true;
  }
  FCLB_DI unsigned childMask(const BoxElem<S>& e) const {
    if (e.meta & 2u) return v.leaf_bits[e.index];  // partial leaf: its occupied voxels
    const uint32_t depth = e.meta >> 8;
    const bool child_leaf = int(depth) + 3 >= v.num_layers;  // isChildLayerLeafNode(parent.depth), octree-inl.h:161-163
    unsigned m = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const uint32_t ci = v.children[size_t(8) * e.index + c];
      if (ci == 0xffffffffu) continue;
      if (child_leaf) {
        if (v.leaf_bits[ci] == 0) continue;
      } else if (v.pruned && v.pruned[ci]) {
        continue;  // prune_internal_nodes (octree2_solver_traverse-inl.h:178-181)
      }
      m |= 1u << c;
    }
    return m;
  }
  FCLB_DI BoxElem<S> child(const BoxElem<S>& e, int c) const {
    BoxElem<S> ch;
    V3<S> cmn, cmx;
    childAabb(mk<S>(e.mn[0], e.mn[1], e.mn[2]), mk<S>(e.mx[0], e.mx[1], e.mx[2]), c, cmn, cmx);
    ch.mn[0] = cmn.x; ch.mn[1] = cmn.y; ch.mn[2] = cmn.z;
    ch.mx[0] = cmx.x; ch.mx[1] = cmx.y; ch.mx[2] = cmx.z;
    const uint32_t depth = e.meta >> 8;
    if (e.meta & 2u) {  // voxel of a partial leaf
      ch.index = e.index;
      ch.meta = (depth << 8) | (uint32_t(c) << 4) | 2u | 1u;
      return ch;
    }
    const bool child_leaf = int(depth) + 3 >= v.num_layers;
    ch.index = v.children[size_t(8) * e.index + c];
    const bool terminal = child_leaf ? (v.leaf_bits[ch.index] == 0xffu) : (v.inner_full[ch.index] != 0);
    ch.meta = ((depth + 1) << 8) | (8u << 4) | (child_leaf ? 2u : 0u) | (terminal ? 1u : 0u);
    return ch;
  }
  // encodeOctree2Node(index, is_leaf, child) (octree2_solver_leaf-inl.h:10-20)
  static FCLB_DI long long code(const BoxElem<S>& e) {
    return (long long)e.index + ((long long)((e.meta >> 4) & 0xfu) << 32) + ((long long)((e.meta >> 1) & 1u) << 48);
  }
};

template <typename S, int K>
struct SideOf;
template <typename S>
struct SideOf<S, FCLB_SCENE_HEIGHTMAP> {
  using type = SideHm<S>;
  static FCLB_DI type make(const HmView& h, const OctView&) { return type(h); }
};
template <typename S>
struct SideOf<S, FCLB_SCENE_OCTREE> {
  using type = SideOct<S>;
  static FCLB_DI type make(const HmView&, const OctView& o) { return type(o); }
};

template <typename S, bool MESH>
struct PairElem;
template <typename S>
struct PairElem<S, false> {
  BoxElem<S> a, b;
};
template <typename S>
struct PairElem<S, true> {
  BoxElem<S> a;
  int b, pad;
};

// Descent rule (ours; the reference's only fixes its visiting order): split the node whose REFINABLE extent is
// larger -- a heightmap node only shrinks in x / y when it is split (its z range is the column height at every
// layer), an octree node in all three axes, a mesh node is measured by its longest OBB side.
template <typename S, int K>
FCLB_DI S refinableSize(const BoxElem<S>& e) {
  const S dx = e.mx[0] - e.mn[0], dy = e.mx[1] - e.mn[1];
  const S m = dx > dy ? dx : dy;
  if (K == FCLB_SCENE_HEIGHTMAP) return m;
  const S dz = e.mx[2] - e.mn[2];
  return m > dz ? m : dz;
}

constexpr int kPairWarps = kScenePairWarps;
constexpr int kPairStack = 512;           // node pairs per warp
constexpr int kPairDfsReserve = 7 * 34;   // depth-first head room: <= 7 net pushes per level, <= 15 + 16 + 3 levels
constexpr int kPairQueue = 64;            // queued leaf pairs per warp

template <typename S, int KA, int KB>
__global__ void __launch_bounds__(kPairWarps * 32) scenePairKernel(ScenePairArgs a) {
  constexpr bool MESH = (KB == FCLB_SCENE_BVH);
  using Elem = PairElem<S, MESH>;
  using SA = SideOf<S, KA>;
  using SB = SideOf<S, MESH ? FCLB_SCENE_HEIGHTMAP : KB>;  // unused for a mesh
  extern __shared__ __align__(16) unsigned char s_pair_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wbase = s_pair_raw + size_t(warp) * (kPairStack + kPairQueue) * sizeof(Elem);
  Elem* stack = reinterpret_cast<Elem*>(wbase);
  Elem* queue = stack + kPairStack;
  const typename SA::type sideA = SA::make(a.hm1, a.oct1);
  const typename SB::type sideB = SB::make(a.hm2, a.oct2);
  const S* __restrict__ nodes = static_cast<const S*>(a.bvh2.nodes);
  const S* __restrict__ tris = static_cast<const S*>(a.bvh2.tris);
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned long long st_node = 0, st_leaf = 0;

  while (true) {
    unsigned long long q64 = 0;
    if (lane == 0) q64 = atomicAdd(a.work_counter, 1ull);
    q64 = __shfl_sync(0xffffffffu, q64, 0);
    if (q64 >= a.n) break;
    const size_t q = size_t(q64);
    const Pose<S> tf1 = loadPose(static_cast<const S*>(a.poses1), q);
    const Pose<S> tf2 = loadPose(static_cast<const S*>(a.poses2), q);
    const FixedRot<S> fr = makeFixedRot(tf1, tf2);  // initialize(tf1, tf2): tf_2in1 = tf1^-1 * tf2

    uint32_t count = 0;
    int sp = 0, nq = 0;
    bool done = (a.max_contacts == 0);
    // root pairs (the reference seeds its stack with every top-layer pixel x the other root(s))
    if (!done) {
      const int na = sideA.numRoots();
      const int nb = MESH ? (a.bvh2.n_nodes > 0 ? 1 : 0) : sideB.numRoots();
      if (lane == 0) {
        #pragma unroll 1
        for (int i = 0; i < na; i++) {
          Elem el;
          if (!sideA.root(i, el.a)) continue;
          if constexpr (MESH) {
            if (nb) {
              el.b = 0;
              el.pad = 0;
              stack[sp++] = el;
            }
          } else {
            #pragma unroll 1
            for (int j = 0; j < nb; j++)
              if (sideB.root(j, el.b)) stack[sp++] = el;
          }
        }
      }
      sp = __shfl_sync(0xffffffffu, sp, 0);
    }
    __syncwarp();

    // A query that needs only a few more contacts (boolean collide: one) must not wait for a full batch of leaf
    // pairs, nor expand 32 node pairs per level on the way down: the reference's depth-first walk reaches its first
    // leaf pair after ~one node pair per level.  `eager`: flush the leaf queue after every step and pop narrowly.
    const bool eager = a.max_contacts <= 8;
    auto runLeaf = [&](bool flush) {
      while (!done && (nq >= 32 || ((flush || eager) && nq > 0))) {
        const int batch = nq < 32 ? nq : 32;
        bool hit = false;
        long long c1 = -1, c2 = -1;
        Elem el;
        if (lane < batch) {
          el = queue[nq - 1 - lane];
          c1 = SA::type::code(el.a);
          st_leaf++;
          if constexpr (MESH) {
            // constructBox(aabb, tf1) + GJKSolver::shapeTriangleIntersect(Box, ...) = boxTriangleIntersect
            c2 = el.b;
            const V3<S> bmin = mk<S>(el.a.mn[0], el.a.mn[1], el.a.mn[2]), bmax = mk<S>(el.a.mx[0], el.a.mx[1], el.a.mx[2]);
            const V3<S> side = bmax - bmin;
            const V3<S> center = (bmin + bmax) * S(0.5);
            Pose<S> tf_box;
            tf_box.R = tf1.R;
            tf_box.t = mulMV(tf1.R, center) + tf1.t;
            const Pose<S> toshape0 = compose(inverse(tf_box), tf2);
            V3<S> P[3];
            loadTri(tris, el.b, P);
            const V3<S> h = mk<S>(S(0.5) * side.x, S(0.5) * side.y, S(0.5) * side.z);
            if (!a.cand.count) hit = boxTriangleOverlap(h, apply(toshape0, P[0]), apply(toshape0, P[1]), apply(toshape0, P[2]));
          } else {
            // heightMapOctreeIntersect names the octree side by its bare node_vector_index
            // (heightmap_solver_traverse-inl.h:489-512), the octree solvers by encodeOctree2Node
            if constexpr (KA == FCLB_SCENE_HEIGHTMAP && KB == FCLB_SCENE_OCTREE)
              c2 = (long long)el.b.index;
            else
              c2 = SB::type::code(el.b);
            if (!a.cand.count) hit = !fixedRotDisjointBoxes(fr, el.a.mn, el.a.mx, el.b.mn, el.b.mx, true);
          }
        }
        if (a.cand.count) {  // candidate mode: the leaf batch decides (boxBox2 / box-triangle GJK + EPA with contacts)
          S hb1[6], hb2[6];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            hb1[k] = el.a.mn[k];
            hb1[3 + k] = el.a.mx[k];
            if constexpr (!MESH) {
              hb2[k] = el.b.mn[k];
              hb2[3 + k] = el.b.mx[k];
            } else {
              hb2[k] = hb2[3 + k] = S(0);
            }
          }
          candAppend<S>(a.cand, lane < batch, uint32_t(q), c1, c2, hb1, MESH ? nullptr : hb2);
        }
        nq -= batch;
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm) {
          if (a.out_b1 && hit) {
            const uint32_t slot = count + uint32_t(__popc(hm & lt_mask));
            if (slot < a.max_keep && slot < a.max_contacts) {
              a.out_b1[q * a.max_keep + slot] = c1;
              a.out_b2[q * a.max_keep + slot] = c2;
              if (a.out_box1) {  // Contact::o1_bv / o2_bv for the penetration pass
                S* o1 = static_cast<S*>(a.out_box1) + (q * a.max_keep + slot) * 6;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                  o1[k] = el.a.mn[k];
                  o1[3 + k] = el.a.mx[k];
                }
                if constexpr (!MESH) {
                  S* o2 = static_cast<S*>(a.out_box2) + (q * a.max_keep + slot) * 6;
#pragma unroll
                  for (int k = 0; k < 3; k++) {
                    o2[k] = el.b.mn[k];
                    o2[3 + k] = el.b.mx[k];
                  }
                }
              }
            }
          }
          count += uint32_t(__popc(hm));
          if (count >= a.max_contacts) {
            count = a.max_contacts;
            done = true;
          }
        }
        __syncwarp();
      }
    };

    while (!done && sp > 0) {
      // a popped pair pushes <= 8 children and queues <= 1 leaf pair: bound both before popping.  Wide pops
      // stop while kPairDfsReserve slots are free; from there the warp pops one pair at a time from the top
      // (depth-first order, growth <= 7 per remaining level).
      const int width = eager ? 4 : 32;
      int take = sp < width ? sp : width;
      if (sp + 7 * take > kPairStack - kPairDfsReserve) {
        const int fit = (kPairStack - kPairDfsReserve - sp) / 7;
        take = fit < 1 ? 1 : fit;
      }
      if (take > kPairQueue - nq) take = kPairQueue - nq;
      if (take < 1) {
        runLeaf(true);
        continue;
      }
      Elem el;
      bool have = false;
      if (lane < take) {
        el = stack[sp - 1 - lane];
        have = true;
      }
      sp -= take;
      __syncwarp();
      int n_push = 0;
      bool cand = false, on_a = false;
      unsigned mask = 0;
      int bc0 = 0;  // mesh: first child / triangle id
      if (have) {
        st_node++;
        const bool ta = (el.a.meta & 1u) != 0;
        if constexpr (MESH) {
          const NodeD<S> nd = loadNode(nodes, el.b);
          // the box as an OBB of frame 1 (identity axes): R = R0 * axis2, T = R0 * To2 + T0 - centre
          const V3<S> ctr = mk<S>((el.a.mn[0] + el.a.mx[0]) * S(0.5), (el.a.mn[1] + el.a.mx[1]) * S(0.5), (el.a.mn[2] + el.a.mx[2]) * S(0.5));
          const V3<S> half = mk<S>(S(0.5) * (el.a.mx[0] - el.a.mn[0]), S(0.5) * (el.a.mx[1] - el.a.mn[1]), S(0.5) * (el.a.mx[2] - el.a.mn[2]));
          const M3<S> Rn = mulMM(fr.R, nd.axis);
          const V3<S> Tn = (mulMV(fr.R, nd.To) + fr.t) - ctr;
          const S slack = S(1e-5) * (S(1) + fabs_(ctr.x) + fabs_(ctr.y) + fabs_(ctr.z) + fabs_(fr.t.x) + fabs_(fr.t.y) + fabs_(fr.t.z) +
                                     fabs_(nd.To.x) + fabs_(nd.To.y) + fabs_(nd.To.z));
          if (!nodeObbDisjoint(Rn, Tn, half, nd.extent, slack)) {
            const bool tb = nd.first_child < 0;
            if (ta && tb) {
              cand = true;
              bc0 = -(nd.first_child + 1);
            } else {
              const S eb = nd.extent.x > nd.extent.y ? (nd.extent.x > nd.extent.z ? nd.extent.x : nd.extent.z)
                                                      : (nd.extent.y > nd.extent.z ? nd.extent.y : nd.extent.z);
              on_a = tb || (!ta && refinableSize<S, KA>(el.a) > S(2) * eb);
              if (on_a) {
                mask = sideA.childMask(el.a);
                n_push = __popc(mask);
              } else {
                bc0 = nd.first_child;
                n_push = 2;
              }
            }
          }
        } else {
          if (!nodeBoxesDisjoint(fr, el.a.mn, el.a.mx, el.b.mn, el.b.mx)) {
            const bool tb = (el.b.meta & 1u) != 0;
            if (ta && tb) {
              cand = true;
            } else {
              on_a = tb || (!ta && refinableSize<S, KA>(el.a) > refinableSize<S, KB>(el.b));
              mask = on_a ? sideA.childMask(el.a) : sideB.childMask(el.b);
              n_push = __popc(mask);
            }
          }
        }
      }
      int push_off = n_push;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int pv = __shfl_up_sync(0xffffffffu, push_off, off);
        if (lane >= off) push_off += pv;
      }
      const int tot_push = __shfl_sync(0xffffffffu, push_off, 31);
      const unsigned cm = __ballot_sync(0xffffffffu, cand);
      if (sp + tot_push > kPairStack) {  // deeper than the depth-first head room: report, never corrupt
        if (lane == 0) atomicAdd(&a.stats[2], 1ull);
        done = true;
        break;
      }
      push_off -= n_push;
      if (n_push) {
        Elem ch = el;
        int k = 0;
        if constexpr (MESH) {
          if (!on_a) {
            ch.b = bc0;
            stack[sp + push_off] = ch;
            ch.b = bc0 + 1;
            stack[sp + push_off + 1] = ch;
            mask = 0;
          }
        }
        for (int c = 0; c < 8; c++) {
          if (!(mask & (1u << c))) continue;
          if (on_a) {
            ch.a = sideA.child(el.a, c);
          } else {
            if constexpr (!MESH) ch.b = sideB.child(el.b, c);
          }
          stack[sp + push_off + k] = ch;
          k++;
        }
      }
      if (cand) {
        if constexpr (MESH) el.b = bc0;
        queue[nq + __popc(cm & lt_mask)] = el;
      }
      sp += tot_push;
      nq += __popc(cm);
      __syncwarp();
      runLeaf(false);
    }
    runLeaf(true);
    if (lane == 0) a.counts[q] = count;
    __syncwarp();
  }
  if (a.stats) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      st_node += __shfl_xor_sync(0xffffffffu, st_node, off);
      st_leaf += __shfl_xor_sync(0xffffffffu, st_leaf, off);
    }
    if (lane == 0) {
      atomicAdd(&a.stats[0], st_node);
      atomicAdd(&a.stats[1], st_leaf);
    }
  }
}

template <typename S, int KA, int KB>
cudaError_t launchScenePairT(const ScenePairArgs& a, int grid, cudaStream_t st) {
  using Elem = PairElem<S, KB == FCLB_SCENE_BVH>;
  const size_t smem = size_t(kPairWarps) * (kPairStack + kPairQueue) * sizeof(Elem);
  cudaError_t e = cudaFuncSetAttribute(scenePairKernel<S, KA, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  scenePairKernel<S, KA, KB><<<grid, kPairWarps * 32, smem, st>>>(a);
  return cudaGetLastError();
}

template <typename S>
cudaError_t launchScenePair(const ScenePairArgs& a, int grid, cudaStream_t st) {
  const int k1 = a.kind1, k2 = a.kind2;
  if (k1 == FCLB_SCENE_HEIGHTMAP && k2 == FCLB_SCENE_HEIGHTMAP)
    return launchScenePairT<S, FCLB_SCENE_HEIGHTMAP, FCLB_SCENE_HEIGHTMAP>(a, grid, st);
  if (k1 == FCLB_SCENE_HEIGHTMAP && k2 == FCLB_SCENE_BVH)
    return launchScenePairT<S, FCLB_SCENE_HEIGHTMAP, FCLB_SCENE_BVH>(a, grid, st);
  if (k1 == FCLB_SCENE_HEIGHTMAP && k2 == FCLB_SCENE_OCTREE)
    return launchScenePairT<S, FCLB_SCENE_HEIGHTMAP, FCLB_SCENE_OCTREE>(a, grid, st);
  if (k1 == FCLB_SCENE_OCTREE && k2 == FCLB_SCENE_BVH)
    return launchScenePairT<S, FCLB_SCENE_OCTREE, FCLB_SCENE_BVH>(a, grid, st);
  if (k1 == FCLB_SCENE_OCTREE && k2 == FCLB_SCENE_OCTREE)
    return launchScenePairT<S, FCLB_SCENE_OCTREE, FCLB_SCENE_OCTREE>(a, grid, st);
  return cudaErrorInvalidValue;
}

}  // namespace fclb
