// fclb_octree_build.h -- HOST builder for the flat octree2::Octree<S> arrays.
//
// the step before the path.  The octree kernels name a contact by the
// reference's node numbering (encodeOctree2Node, octree2_solver_leaf-inl.h:10-20), so a builder is
// only a drop-in when it numbers nodes exactly as Octree<S>::rebuildTree does: nodes are appended in
// the order the point stream first reaches them.  This mirror keeps that order (a sequential insert
// per point -- the numbering is a property of the stream order, so it is not parallelised) and the
// reference's arithmetic for the voxel coordinate:
//   layers / root box / inverse resolution   geometry/octree2/octree-inl.h:15-100
//   computeVoxelCoordinate                   geometry/octree2/octree-inl.h:118-142
//   computeChildIndex(voxel, layer)          geometry/octree2/octree-inl.h:180-192
//   isChildLayerLeafNode                     geometry/octree2/octree-inl.h:165-168
//   insertVoxelIntoTree                      geometry/octree2/octree_construction-inl.h:10-74
//   fully-occupied flags                     geometry/octree2/octree_construction-inl.h:111-172
// tests/test_octree_build.py compares every array with the tree exported from oracle/_ref.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace fclb {
namespace hostbuild {

constexpr uint32_t kOctInvalid = 0xffffffffu;

struct OctreeHost {
  std::vector<uint32_t> children;  // 8 per inner node, node 0 = root
  std::vector<uint8_t> full;       // inner_nodes_fully_occupied()
  std::vector<uint8_t> leaf_bits;  // OctreeLeafNode::child_occupied
  double root_box[6];
  int num_layers = 0;
  size_t n_inner() const { return full.size(); }
};

inline int octChildIndex(const uint16_t v[3], int num_layers, int parent_layer) {
  const int diff = num_layers - parent_layer - 2;
  int pos = 0;
  if (v[0] & (1 << diff)) pos += 1;
  if (v[1] & (1 << diff)) pos += 2;
  if (v[2] & (1 << diff)) pos += 4;
  return pos;
}

// OctreeInnerNodeAuxiliaryInfoUpdater::updateRecursive without prune info: the leaf-parent layer stops at
// the first missing / partial child, the upper layers visit every child (their flags are outputs too)
inline bool octUpdateFull(OctreeHost& t, uint32_t node, int depth) {
  const uint32_t* ch = &t.children[size_t(8) * node];
  bool all = true;
  if (depth + 3 >= t.num_layers) {
    for (int c = 0; c < 8; c++)
      if (ch[c] == kOctInvalid || t.leaf_bits[ch[c]] != 0xff) {
        all = false;
        break;
      }
  } else {
    for (int c = 0; c < 8; c++) {
      if (ch[c] == kOctInvalid) {
        all = false;
        continue;
      }
      if (!octUpdateFull(t, ch[c], depth + 1)) all = false;
    }
  }
  t.full[node] = all ? 1 : 0;
  return all;
}

// Octree<S>(bottom_resolution, bottom_half_shape) + rebuildTree over n points (x, y, z doubles, rounded once to S)
template <typename S>
void octreeFromPoints(const double* pts, size_t n, S res, uint32_t bottom_half, OctreeHost& t) {
  int log2h = 0;
  while ((1u << log2h) < bottom_half) log2h++;
  t.num_layers = log2h + 2;
  const S inv = S(1.0) / res;
  const S mx = res * S(bottom_half);
  for (int k = 0; k < 3; k++) {
    t.root_box[k] = double(-mx);
    t.root_box[3 + k] = double(mx);
  }
  t.children.assign(8, kOctInvalid);
  t.full.assign(1, 0);
  t.leaf_bits.clear();
  const int half = int(bottom_half), fullshape = 2 * half;
  const int leaf_depth = t.num_layers - 2;
  for (size_t i = 0; i < n; i++) {
    const S p[3] = {S(pts[3 * i]), S(pts[3 * i + 1]), S(pts[3 * i + 2])};
    // floor(S) + int is an S sum, truncated to int on assignment
    const int x = int(std::floor(p[0] * inv) + S(half));
    const int y = int(std::floor(p[1] * inv) + S(half));
    const int z = int(std::floor(p[2] * inv) + S(half));
    if (!(x >= 0 && x < fullshape && y >= 0 && y < fullshape && z >= 0 && z < fullshape)) continue;
    const uint16_t v[3] = {uint16_t(x), uint16_t(y), uint16_t(z)};
    uint32_t node = 0;
    int depth = 0;
    bool inserted = false;
    while (true) {
      const int c = octChildIndex(v, t.num_layers, depth);
      const bool child_leaf = depth + 3 >= t.num_layers;
      uint32_t child = t.children[size_t(8) * node + c];
      if (child == kOctInvalid) {
        if (child_leaf) {
          child = uint32_t(t.leaf_bits.size());
          t.children[size_t(8) * node + c] = child;
          t.leaf_bits.push_back(uint8_t(1u << octChildIndex(v, t.num_layers, depth + 1)));
          inserted = true;
          break;
        }
        child = uint32_t(t.n_inner());
        t.children.insert(t.children.end(), 8, kOctInvalid);
        t.full.push_back(0);
        t.children[size_t(8) * node + c] = child;
      }
      depth += 1;
      node = child;
      if (child_leaf) break;
    }
    if (!inserted) t.leaf_bits[node] |= uint8_t(1u << octChildIndex(v, t.num_layers, leaf_depth));
  }
  octUpdateFull(t, 0, 0);
}

// ---- pruneOctreeByOBB (geometry/octree2/octree_prune-inl.h:10-103) ------------------------------------------
// OBB<S>::overlap(node box as an identity-axis OBB), OBB<S>::contain(voxel centre) and is_contained(obb, node box)
// decide what is cut, so each keeps the reference's arithmetic: 3-vector reductions left to right (the
// association DESIGN.md 3 pins), and for float the SSE evaluation of OBB<float>::overlap
// (math/bv/OBB-inl.h:573-698 with the non-SSE4, non-FMA helpers of math/math_simd_details.h).
template <typename S>
struct PruneObb {
  S axis[3][3];  // axis[r][c], columns are the box directions
  S To[3], extent[3];
};

// obbDisjoint(B, T, a, b), math/bv/OBB-inl.h:319-436
template <typename S>
inline bool obbDisjointGeneric(const S B[3][3], const S T[3], const S a[3], const S b[3]) {
  S t, s;
  const S reps = S(1e-6);
  S Bf[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Bf[i][j] = std::fabs(B[i][j]) + reps;
  auto rowdot = [](const S M[3][3], int r, const S v[3]) { return M[r][0] * v[0] + M[r][1] * v[1] + M[r][2] * v[2]; };
  auto coldot = [](const S M[3][3], int c, const S v[3]) { return M[0][c] * v[0] + M[1][c] * v[1] + M[2][c] * v[2]; };
  t = (T[0] < 0) ? -T[0] : T[0];
  if (t > (a[0] + rowdot(Bf, 0, b))) return true;
  s = coldot(B, 0, T);
  t = (s < 0) ? -s : s;
  if (t > (b[0] + coldot(Bf, 0, a))) return true;
  t = (T[1] < 0) ? -T[1] : T[1];
  if (t > (a[1] + rowdot(Bf, 1, b))) return true;
  t = (T[2] < 0) ? -T[2] : T[2];
  if (t > (a[2] + rowdot(Bf, 2, b))) return true;
  s = coldot(B, 1, T);
  t = (s < 0) ? -s : s;
  if (t > (b[1] + coldot(Bf, 1, a))) return true;
  s = coldot(B, 2, T);
  t = (s < 0) ? -s : s;
  if (t > (b[2] + coldot(Bf, 2, a))) return true;
#define FCLB_OCT_EDGE(SEXPR, RAD) \
  s = (SEXPR);                    \
  t = (s < 0) ? -s : s;           \
  if (t > (RAD)) return true;
  FCLB_OCT_EDGE(T[2] * B[1][0] - T[1] * B[2][0], a[1] * Bf[2][0] + a[2] * Bf[1][0] + b[1] * Bf[0][2] + b[2] * Bf[0][1])
  FCLB_OCT_EDGE(T[2] * B[1][1] - T[1] * B[2][1], a[1] * Bf[2][1] + a[2] * Bf[1][1] + b[0] * Bf[0][2] + b[2] * Bf[0][0])
  FCLB_OCT_EDGE(T[2] * B[1][2] - T[1] * B[2][2], a[1] * Bf[2][2] + a[2] * Bf[1][2] + b[0] * Bf[0][1] + b[1] * Bf[0][0])
  FCLB_OCT_EDGE(T[0] * B[2][0] - T[2] * B[0][0], a[0] * Bf[2][0] + a[2] * Bf[0][0] + b[1] * Bf[1][2] + b[2] * Bf[1][1])
  FCLB_OCT_EDGE(T[0] * B[2][1] - T[2] * B[0][1], a[0] * Bf[2][1] + a[2] * Bf[0][1] + b[0] * Bf[1][2] + b[2] * Bf[1][0])
  FCLB_OCT_EDGE(T[0] * B[2][2] - T[2] * B[0][2], a[0] * Bf[2][2] + a[2] * Bf[0][2] + b[0] * Bf[1][1] + b[1] * Bf[1][0])
  FCLB_OCT_EDGE(T[1] * B[0][0] - T[0] * B[1][0], a[0] * Bf[1][0] + a[1] * Bf[0][0] + b[1] * Bf[2][2] + b[2] * Bf[2][1])
  FCLB_OCT_EDGE(T[1] * B[0][1] - T[0] * B[1][1], a[0] * Bf[1][1] + a[1] * Bf[0][1] + b[0] * Bf[2][2] + b[2] * Bf[2][0])
  FCLB_OCT_EDGE(T[1] * B[0][2] - T[0] * B[1][2], a[0] * Bf[1][2] + a[1] * Bf[0][2] + b[0] * Bf[2][1] + b[1] * Bf[2][0])
#undef FCLB_OCT_EDGE
  return false;
}

// obbDisjointSSEFloatImpl lane by lane: mat3x4_mul_vec4 = (m0 v0 + m2 v2) + m1 v1, transp_mat3x4_mul_vec4 =
// (v0 m0 + v1 m1) + v2 m2, fmadd / fmsub = separate multiply and add
inline bool obbDisjointSseOrder(const float R[3][3], const float t[3], const float r1[3], const float r2[3]) {
  const float reps = 1e-6f;
  float A[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A[i][j] = std::fabs(R[i][j]) + reps;
  for (int i = 0; i < 3; i++)
    if (std::fabs(t[i]) > r1[i] + ((A[i][0] * r2[0] + A[i][2] * r2[2]) + A[i][1] * r2[1])) return true;
  for (int j = 0; j < 3; j++) {
    const float cd = (t[0] * R[0][j] + t[1] * R[1][j]) + t[2] * R[2][j];
    if (std::fabs(cd) > ((r1[0] * A[0][j] + r1[1] * A[1][j]) + r1[2] * A[2][j]) + r2[j]) return true;
  }
  // rows of the "symmetric matrix" times |R| row k: (s0 x0 + s2 x2) + s1 x1 with the zero terms dropped
  auto rb = [&](int k, int j) {
    const float* x = A[k];
    return j == 0 ? r2[1] * x[2] + r2[2] * x[1] : (j == 1 ? r2[2] * x[0] + r2[0] * x[2] : r2[1] * x[0] + r2[0] * x[1]);
  };
  for (int j = 0; j < 3; j++) {
    const float ra = r1[1] * A[2][j] + r1[2] * A[1][j];
    if (std::fabs(t[2] * R[1][j] - t[1] * R[2][j]) > ra + rb(0, j)) return true;
  }
  for (int j = 0; j < 3; j++) {
    const float ra = r1[0] * A[2][j] + r1[2] * A[0][j];
    if (std::fabs(t[0] * R[2][j] - t[2] * R[0][j]) > ra + rb(1, j)) return true;
  }
  for (int j = 0; j < 3; j++) {
    const float ra = r1[0] * A[1][j] + r1[1] * A[0][j];
    if (std::fabs(t[1] * R[0][j] - t[0] * R[1][j]) > ra + rb(2, j)) return true;
  }
  return false;
}

// pruned_obb.overlap(identity-axis OBB(center c, half extent e))
inline bool pruneObbOverlapsBox(const PruneObb<double>& o, const double c[3], const double e[3]) {
  const double t[3] = {c[0] - o.To[0], c[1] - o.To[1], c[2] - o.To[2]};
  double T[3], R[3][3];
  for (int k = 0; k < 3; k++) T[k] = o.axis[0][k] * t[0] + o.axis[1][k] * t[1] + o.axis[2][k] * t[2];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = o.axis[j][i];  // axis^T * I
  return !obbDisjointGeneric<double>(R, T, o.extent, e);
}
inline bool pruneObbOverlapsBox(const PruneObb<float>& o, const float c[3], const float e[3]) {
  const float t[3] = {c[0] - o.To[0], c[1] - o.To[1], c[2] - o.To[2]};
  float T[3], R[3][3];
  // transp_mat3x3_mul_mat3x4: (l0 r0 + l1 r1) + (l2 r2 + 0)
  for (int k = 0; k < 3; k++) T[k] = (o.axis[0][k] * t[0] + o.axis[1][k] * t[1]) + o.axis[2][k] * t[2];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = o.axis[j][i];
  return !obbDisjointSseOrder(R, T, o.extent, e);
}

template <typename S>
inline bool pruneObbContainsPoint(const PruneObb<S>& o, const S p[3]) {
  const S l[3] = {p[0] - o.To[0], p[1] - o.To[1], p[2] - o.To[2]};
  for (int k = 0; k < 3; k++) {
    const S proj = l[0] * o.axis[0][k] + l[1] * o.axis[1][k] + l[2] * o.axis[2][k];
    if (proj > o.extent[k] || proj < -o.extent[k]) return false;
  }
  return true;
}

// is_contained(obb, aabb), geometry/octree2/octree_util-inl.h:79-107
template <typename S>
inline bool pruneObbContainsBox(const PruneObb<S>& o, const S c[3], const S e[3]) {
  const S d[3] = {c[0] - o.To[0], c[1] - o.To[1], c[2] - o.To[2]};
  for (int j = 0; j < 3; j++) {
    const S cj = o.axis[0][j] * d[0] + o.axis[1][j] * d[1] + o.axis[2][j] * d[2];
    S lo = cj, hi = cj;
    for (int i = 0; i < 3; i++) {
      if (o.axis[i][j] > 0) {
        hi += o.axis[i][j] * e[i];
        lo -= o.axis[i][j] * e[i];
      } else {
        hi -= o.axis[i][j] * e[i];
        lo += o.axis[i][j] * e[i];
      }
    }
    if (hi > o.extent[j]) return false;
    if (lo < -o.extent[j]) return false;
  }
  return true;
}

// every child link reachable from the root (pruned subtrees skipped when a mask is given) stays inside its layer's array
inline bool octLinksInRange(const uint32_t* children, uint32_t n_inner, uint32_t n_leaf, int num_layers, const uint8_t* pruned) {
  struct N {
    uint32_t node;
    int depth;
  };
  std::vector<N> stack{{0u, 0}};
  while (!stack.empty()) {
    const N t = stack.back();
    stack.pop_back();
    if (pruned && pruned[t.node]) continue;
    const bool child_leaf = t.depth + 3 >= num_layers;
    for (int c = 0; c < 8; c++) {
      const uint32_t k = children[size_t(8) * t.node + c];
      if (k == kOctInvalid) continue;
      if (k >= (child_leaf ? n_leaf : n_inner)) return false;
      if (!child_leaf) stack.push_back({k, t.depth + 1});
    }
  }
  return true;
}

// updateRecursive with prune info (octree_construction-inl.h:111-172): a pruned node is not full and is not entered
inline bool octUpdateFullPruned(const uint32_t* children, const uint8_t* pruned, const uint8_t* leaf_bits, uint8_t* full,
                                int num_layers, uint32_t node, int depth) {
  if (pruned[node]) {
    full[node] = 0;
    return false;
  }
  const uint32_t* ch = children + size_t(8) * node;
  bool all = true;
  if (depth + 3 >= num_layers) {
    for (int c = 0; c < 8; c++)
      if (ch[c] == kOctInvalid || leaf_bits[ch[c]] != 0xff) {
        all = false;
        break;
      }
  } else {
    for (int c = 0; c < 8; c++) {
      if (ch[c] == kOctInvalid) {
        all = false;
        continue;
      }
      if (!octUpdateFullPruned(children, pruned, leaf_bits, full, num_layers, ch[c], depth + 1)) all = false;
    }
  }
  full[node] = all ? 1 : 0;
  return all;
}

// pruneOctreeByOBB on the flat arrays.  pruned / full / leaf_bits are the OctreePruneInfo being extended
// (prune_internal_nodes, new_inner_nodes_fully_occupied, new_leaf_nodes): zeros + the tree's own flags and masks
// for a first prune, the previous outputs for a further one.
// Returns false (arrays partly updated) when a child link points outside its layer's array.
template <typename S>
bool octreePrune(const uint32_t* children, uint32_t n_inner, uint32_t n_leaf, int num_layers, const double root_box[6],
                 const double obb15[15], uint8_t* pruned, uint8_t* full, uint8_t* leaf_bits) {
  PruneObb<S> o;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) o.axis[i][j] = S(obb15[3 * i + j]);
  for (int k = 0; k < 3; k++) {
    o.To[k] = S(obb15[9 + k]);
    o.extent[k] = S(obb15[12 + k]);
  }
  struct Frame {
    S mn[3], mx[3];
    uint32_t node;
    int depth;
    bool leaf;
  };
  std::vector<Frame> stack;
  Frame root;
  for (int k = 0; k < 3; k++) {
    root.mn[k] = S(root_box[k]);
    root.mx[k] = S(root_box[3 + k]);
  }
  root.node = 0;
  root.depth = 0;
  root.leaf = false;
  stack.push_back(root);
  auto childBox = [](const Frame& f, int c, S mn[3], S mx[3]) {  // computeChildAABB, octree_util-inl.h:10-37
    for (int k = 0; k < 3; k++) {
      const S mid = S((f.mn[k] + f.mx[k]) * 0.5);
      if (c & (1 << k)) {
        mn[k] = mid;
        mx[k] = f.mx[k];
      } else {
        mn[k] = f.mn[k];
        mx[k] = mid;
      }
    }
  };
  while (!stack.empty()) {
    const Frame f = stack.back();
    stack.pop_back();
    if (!f.leaf && pruned[f.node]) continue;
    S c[3], e[3];
    for (int k = 0; k < 3; k++) {
      c[k] = S((f.mn[k] + f.mx[k]) * 0.5);
      e[k] = S(0.5) * (f.mx[k] - f.mn[k]);
    }
    if (!pruneObbOverlapsBox(o, c, e)) continue;
    if (f.leaf) {
      uint8_t& bits = leaf_bits[f.node];
      for (int ci = 0; ci < 8; ci++) {
        if (!(bits & (1u << ci))) continue;
        S mn[3], mx[3], vc[3];
        childBox(f, ci, mn, mx);
        for (int k = 0; k < 3; k++) vc[k] = S((mn[k] + mx[k]) * 0.5);
        if (pruneObbContainsPoint(o, vc)) bits = uint8_t(bits & ~(1u << ci));
      }
      continue;
    }
    if (pruneObbContainsBox(o, c, e)) {
      pruned[f.node] = 1;
      continue;
    }
    const uint32_t* ch = children + size_t(8) * f.node;
    for (int ci = 0; ci < 8; ci++) {
      if (ch[ci] == kOctInvalid) continue;
      Frame g;
      childBox(f, ci, g.mn, g.mx);
      g.node = ch[ci];
      g.depth = f.depth + 1;
      g.leaf = f.depth + 3 >= num_layers;
      if (g.node >= (g.leaf ? n_leaf : n_inner)) return false;
      stack.push_back(g);
    }
  }
  octUpdateFullPruned(children, pruned, leaf_bits, full, num_layers, 0, 0);
  return true;
}

// Octree<S>::rebuildAccordingToPruneInfo (octree_construction-inl.h:247-369): the pruned tree consolidated into a
// fresh, renumbered one.  Numbering follows the reference's LIFO task stack: an inner child gets its new index when
// its parent is expanded, leaf nodes are appended when their parent is popped.
inline bool octreeConsolidate(const uint32_t* children, uint32_t n_inner, uint32_t n_leaf, const uint8_t* pruned,
                              const uint8_t* leaf_bits, int num_layers, OctreeHost& out) {
  if (!octLinksInRange(children, n_inner, n_leaf, num_layers, pruned)) return false;
  out.num_layers = num_layers;
  out.children.assign(8, kOctInvalid);
  out.full.assign(1, 0);
  out.leaf_bits.clear();
  if (pruned[0]) return true;  // clearNodes(): a bare root
  struct Task {
    uint32_t node;
    int depth;
    uint32_t placement;
  };
  std::vector<Task> stack{{0u, 0, 0u}};
  while (!stack.empty()) {
    const Task t = stack.back();
    stack.pop_back();
    const uint32_t* ch = children + size_t(8) * t.node;
    uint32_t fresh[8];
    for (int i = 0; i < 8; i++) fresh[i] = kOctInvalid;
    if (t.depth + 3 >= num_layers) {
      for (int i = 0; i < 8; i++) {
        if (ch[i] == kOctInvalid) continue;
        fresh[i] = uint32_t(out.leaf_bits.size());
        out.leaf_bits.push_back(leaf_bits[ch[i]]);
      }
    } else {
      for (int i = 0; i < 8; i++) {
        if (ch[i] == kOctInvalid || pruned[ch[i]]) continue;
        fresh[i] = uint32_t(out.n_inner());
        out.children.insert(out.children.end(), 8, kOctInvalid);
        out.full.push_back(0);
        stack.push_back({ch[i], t.depth + 1, fresh[i]});
      }
    }
    for (int i = 0; i < 8; i++) out.children[size_t(8) * t.placement + i] = fresh[i];
  }
  octUpdateFull(out, 0, 0);
  return true;
}

}  // namespace hostbuild
}  // namespace fclb
