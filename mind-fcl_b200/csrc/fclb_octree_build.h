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

}  // namespace hostbuild
}  // namespace fclb
