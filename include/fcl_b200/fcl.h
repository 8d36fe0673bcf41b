// fcl_b200/fcl.h -- host-side C++ mirror of the mind-fcl public API for the
// batched narrowphase path, implemented on top of the C ABI (include/fclb200.h).
//
// What is mirrored (same names, argument meaning and error behaviour):
//   fcl::collide(o1, tf1, o2, tf2, request, result)      reference narrowphase/collision.h:55-64
//   fcl::collide(CollisionObject*, CollisionObject*, ..)  reference narrowphase/collision.h:62-64
//   CollisionRequest<S> / CollisionResult<S> / Contact<S>  collision_request.h:51-94, collision_result.h:56-95, contact.h:46-101
//   CollisionObject<S>, Box / Sphere / Ellipsoid / Capsule / Cone / Cylinder / Convex, BVHModel<OBBRSS<S>>
// What is added (absent from mind-fcl):
//   fcl::distance + DistanceRequest / DistanceResult  == detail::GJKSolver<S>::shapeDistance semantics
//   fcl::collideBatch / fcl::distanceBatch            one call for many (pair, pose) queries
//
// mind-fcl's math types come from Eigen, which this repository does not depend
// on: Vector3 / Matrix3 / Transform3 below are minimal stand-ins with the same
// member names for the operations this path needs (linear(), translation()).
// An integration inside mind-fcl uses Eigen's types and calls the C ABI directly
// (INTEGRATION.md).
//
// Error behaviour follows the reference: no exceptions on the query path;
// unsupported pairs / max_contacts == 0 print a warning and return 0
// (collision-inl.h:73-110).  A missing GPU is NOT silently tolerated: the first
// call aborts with the engine's error message (there is no CPU fallback).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <vector>

#include "../fclb200.h"

namespace fcl {

template <typename S>
struct Vector3 {
  S v[3]{0, 0, 0};
  Vector3() = default;
  Vector3(S x, S y, S z) : v{x, y, z} {}
  S& operator[](int i) { return v[i]; }
  S operator[](int i) const { return v[i]; }
  static Vector3 Zero() { return Vector3(); }
  static Vector3 UnitX() { return Vector3(1, 0, 0); }
};
template <typename S>
struct Matrix3 {
  S m[3][3]{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  S& operator()(int i, int j) { return m[i][j]; }
  S operator()(int i, int j) const { return m[i][j]; }
  static Matrix3 Identity() { return Matrix3(); }
};
template <typename S>
struct Transform3 {
  Matrix3<S> R;
  Vector3<S> t;
  static Transform3 Identity() { return Transform3(); }
  Matrix3<S>& linear() { return R; }
  const Matrix3<S>& linear() const { return R; }
  Vector3<S>& translation() { return t; }
  const Vector3<S>& translation() const { return t; }
  void setIdentity() { *this = Transform3(); }
  void toPose12(S* out) const {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) out[3 * i + j] = R.m[i][j];
    for (int i = 0; i < 3; i++) out[9 + i] = t.v[i];
  }
};

namespace detail {
inline void check(int rc, const char* what) {
  if (rc != FCLB_OK) {
    std::fprintf(stderr, "fcl_b200: %s failed (%d): %s\n", what, rc, fclb_last_error());
    std::abort();
  }
}
template <typename S>
constexpr int scalarType() {
  return sizeof(S) == 4 ? FCLB_F32 : FCLB_F64;
}
}  // namespace detail

enum NODE_TYPE {  // subset of geometry/collision_geometry.h:50-54 this path supports
  GEOM_BOX,
  GEOM_SPHERE,
  GEOM_ELLIPSOID,
  GEOM_CAPSULE,
  GEOM_CONE,
  GEOM_CYLINDER,
  GEOM_CONVEX,
  BV_OBBRSS,
  GEOM_HEIGHTMAP,
  GEOM_OCTREE2
};

template <typename S>
class CollisionGeometry {
 public:
  virtual ~CollisionGeometry() = default;
  virtual NODE_TYPE getNodeType() const = 0;
  virtual bool isShape() const { return true; }
  // shape record for the engine's shape table
  virtual fclb_shape shapeRecord() const = 0;
};

template <typename S>
class ShapeBase : public CollisionGeometry<S> {};

#define FCLB_DEFINE_SHAPE(NAME, NODE, CODE, ...)                                   \
  fclb_shape shapeRecord() const override {                                        \
    fclb_shape s{};                                                                \
    s.type = CODE;                                                                 \
    const double p[3] = {__VA_ARGS__};                                             \
    for (int i = 0; i < 3; i++) s.p[i] = p[i];                                     \
    return s;                                                                      \
  }                                                                                \
  NODE_TYPE getNodeType() const override { return NODE; }

template <typename S>
class Box : public ShapeBase<S> {
 public:
  Box(S x, S y, S z) : side(x, y, z) {}
  Vector3<S> side;
  FCLB_DEFINE_SHAPE(Box, GEOM_BOX, FCLB_BOX, double(side[0]), double(side[1]), double(side[2]))
};
template <typename S>
class Sphere : public ShapeBase<S> {
 public:
  explicit Sphere(S r) : radius(r) {}
  S radius;
  FCLB_DEFINE_SHAPE(Sphere, GEOM_SPHERE, FCLB_SPHERE, double(radius), 0, 0)
};
template <typename S>
class Ellipsoid : public ShapeBase<S> {
 public:
  Ellipsoid(S a, S b, S c) : radii(a, b, c) {}
  Vector3<S> radii;
  FCLB_DEFINE_SHAPE(Ellipsoid, GEOM_ELLIPSOID, FCLB_ELLIPSOID, double(radii[0]), double(radii[1]), double(radii[2]))
};
template <typename S>
class Capsule : public ShapeBase<S> {
 public:
  Capsule(S r, S l) : radius(r), lz(l) {}
  S radius, lz;
  FCLB_DEFINE_SHAPE(Capsule, GEOM_CAPSULE, FCLB_CAPSULE, double(radius), double(lz), 0)
};
template <typename S>
class Cone : public ShapeBase<S> {
 public:
  Cone(S r, S l) : radius(r), lz(l) {}
  S radius, lz;
  FCLB_DEFINE_SHAPE(Cone, GEOM_CONE, FCLB_CONE, double(radius), double(lz), 0)
};
template <typename S>
class Cylinder : public ShapeBase<S> {
 public:
  Cylinder(S r, S l) : radius(r), lz(l) {}
  S radius, lz;
  FCLB_DEFINE_SHAPE(Cylinder, GEOM_CYLINDER, FCLB_CYLINDER, double(radius), double(lz), 0)
};
#undef FCLB_DEFINE_SHAPE

// Convex<S>(vertices, num_faces, faces) -- reference geometry/shape/convex.h:106-108;
// faces use the reference encoding (count, v0, v1, ... per face).
template <typename S>
class Convex : public ShapeBase<S> {
 public:
  Convex(const std::shared_ptr<const std::vector<Vector3<S>>>& vertices, int num_faces,
         const std::shared_ptr<const std::vector<int>>& faces)
      : vertices_(vertices), num_faces_(num_faces), faces_(faces) {
    std::vector<double> v;
    for (const auto& p : *vertices_)
      for (int k = 0; k < 3; k++) v.push_back(double(p[k]));
    detail::check(fclb_convex_upload(v.data(), int(vertices_->size()), faces_->data(), int(faces_->size()), num_faces_,
                                     &slot_),
                  "fclb_convex_upload");
  }
  NODE_TYPE getNodeType() const override { return GEOM_CONVEX; }
  fclb_shape shapeRecord() const override {
    fclb_shape s{};
    s.type = FCLB_CONVEX;
    s.geom = slot_;
    return s;
  }
  const std::vector<Vector3<S>>& getVertices() const { return *vertices_; }

 private:
  std::shared_ptr<const std::vector<Vector3<S>>> vertices_;
  int num_faces_;
  std::shared_ptr<const std::vector<int>> faces_;
  uint32_t slot_ = 0;
};

// BVHModel<OBBRSS<S>>: the tree is built by the caller's builder (mind-fcl's
// BVHModel::endModel in an integration) and handed over flattened.
template <typename S>
struct OBBRSS {};
template <typename BV>
class BVHModel;
template <typename S>
class BVHModel<OBBRSS<S>> : public CollisionGeometry<S> {
 public:
  // obb: 15 S per node (axis row-major, To, extent); first_child per node; tri_verts: 9 S per triangle
  BVHModel(const std::vector<S>& obb, const std::vector<int32_t>& first_child, const std::vector<S>& tri_verts) {
    detail::check(fclb_bvh_upload(obb.data(), first_child.data(), int(first_child.size()), tri_verts.data(),
                                  int(tri_verts.size() / 9), detail::scalarType<S>(), &handle_),
                  "fclb_bvh_upload");
  }
  // the reference's own build protocol (geometry/bvh/BVH_model.h:104-124): beginModel, addSubModel, endModel;
  // endModel runs the host mirror of BVHModel::buildTree (fclb_bvh_build) and uploads the tree
  BVHModel() = default;
  int beginModel() {
    verts_.clear();
    tris_.clear();
    return 0;
  }
  int addSubModel(const std::vector<Vector3<S>>& points, const std::vector<std::array<int, 3>>& triangles) {
    const int32_t base = int32_t(verts_.size() / 3);
    for (const auto& p : points)
      for (int k = 0; k < 3; k++) verts_.push_back(double(p[k]));
    for (const auto& t : triangles)
      for (int k = 0; k < 3; k++) tris_.push_back(base + t[k]);
    return 0;
  }
  int endModel() {
    if (handle_) fclb_bvh_release(handle_);
    handle_ = 0;
    detail::check(fclb_bvh_build(verts_.data(), int(verts_.size() / 3), tris_.data(), int(tris_.size() / 3),
                                 detail::scalarType<S>(), &handle_),
                  "fclb_bvh_build");
    return 0;
  }
  // the replace / update protocol (BVH_model.h:127-160): same triangles, moved vertices; the hierarchy keeps its
  // topology and is refitted on the device -- bottom-up by default, as the reference's default arguments ask
  int beginReplaceModel() {
    if (!handle_) return -1;  // BVH_ERR_BUILD_EMPTY_PREVIOUS_FRAME
    new_verts_.clear();
    return 0;
  }
  int replaceSubModel(const std::vector<Vector3<S>>& points) {
    for (const auto& p : points)
      for (int k = 0; k < 3; k++) new_verts_.push_back(double(p[k]));
    return 0;
  }
  int endReplaceModel(bool refit = true, bool bottomup = true) {
    if (new_verts_.size() != verts_.size()) {
      std::cerr << "BVH Error! The replaced model should have the same number of vertices_ as the old model.\n";
      return -4;  // BVH_ERR_INCORRECT_DATA
    }
    verts_ = new_verts_;
    if (!refit) return endModel();  // reconstruct the tree from the current frame
    std::vector<S> tri_verts(tris_.size() * 3);
    for (std::size_t i = 0; i < tris_.size(); i++)
      for (int k = 0; k < 3; k++) tri_verts[3 * i + k] = S(verts_[3 * std::size_t(tris_[i]) + k]);
    const int n_tris = int(tris_.size() / 3);
    detail::check(bottomup ? fclb_bvh_refit_bottomup_host(handle_, tri_verts.data(), n_tris)
                           : fclb_bvh_refit_host(handle_, tri_verts.data(), n_tris),
                  "fclb_bvh_refit");
    return 0;
  }
  int beginUpdateModel() { return beginReplaceModel(); }
  int updateSubModel(const std::vector<Vector3<S>>& points) { return replaceSubModel(points); }
  int endUpdateModel(bool refit = true, bool bottomup = true) { return endReplaceModel(refit, bottomup); }
  int getNumBVs() const {
    int n = 0;
    if (handle_) fclb_bvh_info(handle_, &n, nullptr, nullptr);
    return n;
  }
  ~BVHModel() override {
    if (handle_) fclb_bvh_release(handle_);
  }
  NODE_TYPE getNodeType() const override { return BV_OBBRSS; }
  bool isShape() const override { return false; }
  fclb_shape shapeRecord() const override { return fclb_shape{}; }
  fclb_handle handle() const { return handle_; }

 private:
  fclb_handle handle_ = 0;
  std::vector<double> verts_, new_verts_;
  std::vector<int32_t> tris_;
};

// heightmap::LayeredHeightMap<S> (geometry/heightmap/layered_heightmap.h): the bottom layer lives on the host
// until it is wrapped in a HeightMapCollisionGeometry, which uploads it (coarser layers are rebuilt on upload).
namespace heightmap {
template <typename S>
class LayeredHeightMap {
 public:
  LayeredHeightMap(S bottom_resolution, uint16_t bottom_half_map_shape)
      : resolution_(bottom_resolution), half_(bottom_half_map_shape),
        heights_(std::size_t(2 * bottom_half_map_shape) * 2 * bottom_half_map_shape, 0) {}
  // updateHeightsByPointGenerationFunctor (layered_heightmap-inl.h:134-140): generator(i, x, y, z)
  template <typename PointGenerator>
  void updateHeightsByPointGenerationFunctor(const PointGenerator& gen, int n_points) {
    std::vector<double> pts(std::size_t(3) * n_points);
    for (int i = 0; i < n_points; i++) {
      S x, y, z;
      gen(i, x, y, z);
      pts[3 * std::size_t(i)] = double(x);
      pts[3 * std::size_t(i) + 1] = double(y);
      pts[3 * std::size_t(i) + 2] = double(z);
    }
    detail::check(fclb_heightmap_build_host(pts.data(), std::size_t(n_points), double(resolution_), double(resolution_),
                                            half_, half_, detail::scalarType<S>(), heights_.data()),
                  "fclb_heightmap_build_host");
  }
  void resetHeights() { std::fill(heights_.begin(), heights_.end(), uint16_t(0)); }
  S resolution() const { return resolution_; }
  uint16_t half_shape() const { return half_; }
  const std::vector<uint16_t>& bottom_heights_mm() const { return heights_; }

 private:
  S resolution_;
  uint16_t half_;
  std::vector<uint16_t> heights_;
};
}  // namespace heightmap

// geometry/heightmap/heightmap_collision_geometry.h
template <typename S>
class HeightMapCollisionGeometry : public CollisionGeometry<S> {
 public:
  explicit HeightMapCollisionGeometry(std::shared_ptr<const heightmap::LayeredHeightMap<S>> map) : map_(std::move(map)) {
    detail::check(fclb_heightmap_upload(map_->bottom_heights_mm().data(), 2u * map_->half_shape(), 2u * map_->half_shape(),
                                        double(map_->resolution()), double(map_->resolution()), 0, &handle_),
                  "fclb_heightmap_upload");
  }
  ~HeightMapCollisionGeometry() override { fclb_heightmap_release(handle_); }
  NODE_TYPE getNodeType() const override { return GEOM_HEIGHTMAP; }
  bool isShape() const override { return false; }
  fclb_shape shapeRecord() const override { return fclb_shape{}; }
  fclb_handle handle() const { return handle_; }
  const std::shared_ptr<const heightmap::LayeredHeightMap<S>>& raw_heightmap() const { return map_; }

 private:
  std::shared_ptr<const heightmap::LayeredHeightMap<S>> map_;
  fclb_handle handle_ = 0;
};

// geometry/octree2/octree.h: Octree<S>(bottom_resolution, bottom_half_shape) + rebuildTree(generator, n_points).
// The tree lives on the host as the reference's flat node arrays, numbered as the reference numbers them
// (fclb_octree_build_host), until it is wrapped in an Octree2CollisionGeometry, which uploads it.
namespace octree2 {
template <typename S>
class Octree {
 public:
  using PointGenerationFunc = std::function<void(int index, S& x, S& y, S& z)>;
  Octree(S bottom_resolution_xyz, std::uint16_t bottom_half_shape) : resolution_(bottom_resolution_xyz), half_(bottom_half_shape) {
    rebuildTree([](int, S&, S&, S&) {}, 0);
  }
  // octree_construction-inl.h:181-205
  void rebuildTree(const PointGenerationFunc& point_generator, int n_points) {
    std::vector<double> pts(std::size_t(3) * n_points);
    for (int i = 0; i < n_points; i++) {
      S x, y, z;
      point_generator(i, x, y, z);
      pts[3 * std::size_t(i)] = double(x);
      pts[3 * std::size_t(i) + 1] = double(y);
      pts[3 * std::size_t(i) + 2] = double(z);
    }
    uint32_t n_inner = 0, n_leaf = 0;
    const int st = detail::scalarType<S>();
    int rc = fclb_octree_build_host(pts.data(), std::size_t(n_points), double(resolution_), half_, st, nullptr, nullptr, 0,
                                    &n_inner, nullptr, 0, &n_leaf, root_aabb_.data(), &n_layers_);
    if (rc != FCLB_ERR_CAPACITY) detail::check(rc ? rc : FCLB_ERR_BAD_ARG, "fclb_octree_build_host");
    inner_children_.assign(std::size_t(8) * n_inner, 0);
    inner_full_.assign(n_inner, 0);
    leaf_bits_.assign(n_leaf ? n_leaf : 1, 0);
    detail::check(fclb_octree_build_host(pts.data(), std::size_t(n_points), double(resolution_), half_, st, inner_children_.data(),
                                         inner_full_.data(), n_inner, &n_inner, leaf_bits_.data(), uint32_t(leaf_bits_.size()),
                                         &n_leaf, root_aabb_.data(), &n_layers_),
                  "fclb_octree_build_host");
    leaf_bits_.resize(n_leaf);
  }
  // isPointOccupied / isVoxelOccupied (octree-inl.h:118-142,194-228): voxel coordinate in S, then a descent that stops
  // at a fully occupied inner node
  bool isPointOccupied(const Vector3<S>& point) const {
    const S inv = S(1.0) / resolution_;
    int v[3];
    for (int k = 0; k < 3; k++) {
      v[k] = int(std::floor(point[k] * inv) + S(int(half_)));
      if (v[k] < 0 || v[k] >= 2 * int(half_)) return false;
    }
    std::uint32_t node = 0;
    for (int depth = 0;; depth++) {
      if (inner_full_[node]) return true;
      const int diff = n_layers_ - depth - 2;
      const int c = ((v[0] >> diff) & 1) | (((v[1] >> diff) & 1) << 1) | (((v[2] >> diff) & 1) << 2);
      const std::uint32_t child = inner_children_[std::size_t(8) * node + c];
      if (child == 0xffffffffu) return false;
      node = child;
      if (depth + 3 >= n_layers_) break;
    }
    const int c = (v[0] & 1) | ((v[1] & 1) << 1) | ((v[2] & 1) << 2);
    return (leaf_bits_[node] >> c) & 1;
  }
  std::uint8_t n_layers() const { return std::uint8_t(n_layers_); }
  std::size_t n_inner_nodes() const { return inner_full_.size(); }
  std::size_t n_leaf_nodes() const { return leaf_bits_.size(); }
  const std::vector<uint32_t>& inner_children() const { return inner_children_; }
  const std::vector<uint8_t>& inner_nodes_fully_occupied() const { return inner_full_; }
  const std::vector<uint8_t>& leaf_bits() const { return leaf_bits_; }
  const std::array<double, 6>& root_aabb() const { return root_aabb_; }

 private:
  S resolution_;
  std::uint16_t half_;
  int n_layers_ = 0;
  std::vector<uint32_t> inner_children_;
  std::vector<uint8_t> inner_full_, leaf_bits_;
  std::array<double, 6> root_aabb_{};
};
}  // namespace octree2

// math/bv/OBB.h: the fields pruneBy reads (axis columns = box directions, centre To, half extents)
template <typename S>
struct OBB {
  Matrix3<S> axis;
  Vector3<S> To;
  Vector3<S> extent;
};

// geometry/octree2/octree_collision_geometry.h: wraps an octree2::Octree built here, or the flat node arrays of an
// octree the caller built with mind-fcl (see fclb_octree_upload).  The host arrays are kept so that pruneBy can
// extend the prune info the way OctreePruneInfo does (octree_node.h:56-72).
template <typename S>
class Octree2CollisionGeometry : public CollisionGeometry<S> {
 public:
  using ConstPtr = std::shared_ptr<const Octree2CollisionGeometry<S>>;
  explicit Octree2CollisionGeometry(std::shared_ptr<const octree2::Octree<S>> octree)
      : Octree2CollisionGeometry(octree->inner_children(), octree->inner_nodes_fully_occupied(), octree->leaf_bits(),
                                 octree->root_aabb(), int(octree->n_layers())) {}
  Octree2CollisionGeometry(const std::vector<uint32_t>& inner_children, const std::vector<uint8_t>& inner_full,
                           const std::vector<uint8_t>& leaf_bits, const std::array<double, 6>& root_aabb, int n_layers,
                           const std::vector<uint8_t>& pruned = {})
      : children_(inner_children), full_(inner_full), leaf_(leaf_bits), pruned_(pruned), root_(root_aabb), n_layers_(n_layers) {
    detail::check(fclb_octree_upload(children_.data(), full_.data(), uint32_t(full_.size()), leaf_.data(), uint32_t(leaf_.size()),
                                     pruned_.empty() ? nullptr : pruned_.data(), root_.data(), n_layers_, &handle_),
                  "fclb_octree_upload");
  }
  ~Octree2CollisionGeometry() override { fclb_octree_release(handle_); }
  // pruneBy(obb, rebuild_octree = false), octree_collision_geometry-inl.h:97-130: a new geometry over the same nodes
  // whose prune info is this one's extended by the box (pruneOctreeByOBB); with rebuild_octree the pruned tree is
  // consolidated into a renumbered one without prune info (Octree::rebuildAccordingToPruneInfo).
  ConstPtr pruneBy(const OBB<S>& obb, bool rebuild_octree = false) const {
    double o[15];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) o[3 * i + j] = double(obb.axis(i, j));
    for (int k = 0; k < 3; k++) {
      o[9 + k] = double(obb.To[k]);
      o[12 + k] = double(obb.extent[k]);
    }
    std::vector<uint8_t> pruned = pruned_.empty() ? std::vector<uint8_t>(full_.size(), 0) : pruned_, full = full_, leaf = leaf_;
    detail::check(fclb_octree_prune_host(children_.data(), uint32_t(full.size()), uint32_t(leaf.size()), root_.data(), n_layers_, o,
                                         detail::scalarType<S>(), pruned.data(), full.data(), leaf.data()),
                  "fclb_octree_prune_host");
    if (!rebuild_octree) return std::make_shared<const Octree2CollisionGeometry<S>>(children_, full, leaf, root_, n_layers_, pruned);
    std::vector<uint32_t> n_children(children_.size());
    std::vector<uint8_t> n_full(full.size()), n_leaf(leaf.size() ? leaf.size() : 1);
    uint32_t ni = 0, nl = 0;
    detail::check(fclb_octree_consolidate_host(children_.data(), uint32_t(full.size()), pruned.data(), leaf.data(),
                                               uint32_t(leaf.size()), n_layers_, n_children.data(), n_full.data(), &ni,
                                               n_leaf.data(), &nl),
                  "fclb_octree_consolidate_host");
    n_children.resize(std::size_t(8) * ni);
    n_full.resize(ni);
    n_leaf.resize(nl);
    return std::make_shared<const Octree2CollisionGeometry<S>>(n_children, n_full, n_leaf, root_, n_layers_);
  }
  NODE_TYPE getNodeType() const override { return GEOM_OCTREE2; }
  bool isShape() const override { return false; }
  fclb_shape shapeRecord() const override { return fclb_shape{}; }
  fclb_handle handle() const { return handle_; }
  const std::vector<uint8_t>& inner_nodes_fully_occupied() const { return full_; }
  const std::vector<uint8_t>& leaf_bits() const { return leaf_; }
  const std::vector<uint8_t>* prune_internal_nodes() const { return pruned_.empty() ? nullptr : &pruned_; }

 private:
  std::vector<uint32_t> children_;
  std::vector<uint8_t> full_, leaf_, pruned_;
  std::array<double, 6> root_;
  int n_layers_;
  fclb_handle handle_ = 0;
};

// narrowphase/collision_request.h:51-94
template <typename S>
struct CollisionRequest {
  explicit CollisionRequest(std::size_t n_max_contacts = 1) : num_max_contacts_(n_max_contacts) {}
  void disablePenetration() { penetration_mode_ = FCLB_PEN_DISABLED; }
  void useDefaultPenetration() { penetration_mode_ = FCLB_PEN_DEFAULT_GJK_EPA; }
  // collision_request.h:76-80 -> detail/collision_penetration_mode.h:16-76
  // both normalise the direction in S, a zero vector becomes UnitZ (collision_request-inl.h:82-106)
  void useDirectedPenetration(const Vector3<S>& shape2_escape_direction) {
    penetration_mode_ = FCLB_PEN_DIRECTED;
    direction_ = unitOrZ(shape2_escape_direction);
  }
  void useIncrementalMinimumDistancePenetration(const Vector3<S>& shape2_escape_direction_init) {
    penetration_mode_ = FCLB_PEN_INCREMENTAL_MIN;
    direction_ = unitOrZ(shape2_escape_direction_init);
  }
  static Vector3<S> unitOrZ(const Vector3<S>& d) {
    const S n = std::sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    if (n <= S(0.0)) return Vector3<S>(S(0), S(0), S(1));
    return Vector3<S>(d[0] / n, d[1] / n, d[2] / n);
  }
  bool isPenetrationEnabled() const { return penetration_mode_ != FCLB_PEN_DISABLED; }
  std::size_t maxNumContacts() const { return num_max_contacts_; }
  void setMaxContactCount(std::size_t n) { num_max_contacts_ = n; }
  void setBinaryCollisionTolerance(S t) { binary_collision_tolerance_ = t; }
  void setPenetrationDistanceTolerance(S t) { distance_tolerance_ = t; }
  S binaryCollisionTolerance() const { return binary_collision_tolerance_; }
  S distanceTolerance() const { return distance_tolerance_; }
  fclb_request toAbi() const {
    fclb_request r{};
    r.max_contacts = num_max_contacts_ > 0xffffffffull ? 0xffffffffu : uint32_t(num_max_contacts_);
    r.penetration_mode = penetration_mode_;
    for (int k = 0; k < 3; k++) r.dir[k] = double(direction_[k]);
    r.binary_tol = double(binary_collision_tolerance_);
    r.distance_tol = double(distance_tolerance_);
    return r;
  }

 private:
  std::size_t num_max_contacts_;
  uint32_t penetration_mode_ = FCLB_PEN_DISABLED;
  Vector3<S> direction_;
  S binary_collision_tolerance_{S(1e-6)};
  S distance_tolerance_{S(1e-6)};
};

// narrowphase/contact.h:46-101
template <typename S>
struct Contact {
  const CollisionGeometry<S>* o1 = nullptr;
  const CollisionGeometry<S>* o2 = nullptr;
  intptr_t b1 = -1, b2 = -1;
  Vector3<S> normal, pos;
  S penetration_depth = 0;
  static constexpr int NONE = -1;
};

// narrowphase/collision_result.h:56-95
template <typename S>
struct CollisionResult {
  void addContact(const Contact<S>& c) { contacts_.push_back(c); }
  std::size_t numContacts() const { return contacts_.size(); }
  bool isCollision() const { return !contacts_.empty(); }
  void clear() { contacts_.clear(); }
  const Contact<S>& getContact(std::size_t i) const { return contacts_[i]; }
  const std::vector<Contact<S>>& getContacts() const { return contacts_; }

 private:
  std::vector<Contact<S>> contacts_;
};

// narrowphase/collision_object.h
template <typename S>
class CollisionObject {
 public:
  CollisionObject(const std::shared_ptr<const CollisionGeometry<S>>& g, const Transform3<S>& tf = Transform3<S>())
      : geom_(g), tf_(tf) {}
  const std::shared_ptr<const CollisionGeometry<S>>& collisionGeometry() const { return geom_; }
  const Transform3<S>& getTransform() const { return tf_; }
  void setTransform(const Transform3<S>& tf) { tf_ = tf; }
  // computeAABB (collision_object-inl.h:141-154), shapes only: min xyz, max xyz
  std::array<S, 6> computeAABB() const {
    const fclb_shape rec = geom_->shapeRecord();
    fclb_handle one = 0;
    detail::check(fclb_shapes_upload(&rec, 1, &one), "fclb_shapes_upload");
    S pose[12];
    tf_.toPose12(pose);
    const uint32_t sid = 0;
    std::array<S, 6> box{};
    detail::check(fclb_compute_aabb_batch_host(one, &sid, pose, 1, detail::scalarType<S>(), box.data()),
                  "fclb_compute_aabb_batch_host");
    fclb_release(one);
    return box;
  }

 private:
  std::shared_ptr<const CollisionGeometry<S>> geom_;
  Transform3<S> tf_;
};

// broadphase/broadphase_common.h:13-22 and broadphase_AABB_tree.h: the tree lives on the device; the collision
// functor runs on the host over the returned pair list (callbacks cannot cross the C ABI).
template <typename S>
struct BroadphaseObjectInfo {
  std::array<S, 6> bv{};  // min xyz, max xyz
  std::uint64_t user_id{0};
};
template <typename S>
class BroadphaseAABB_Tree {
 public:
  using CollisionFn = bool (*)(std::uint64_t, std::uint64_t, void*);
  ~BroadphaseAABB_Tree() {
    if (handle_) fclb_broadphase_release(handle_);
  }
  void Rebuild(BroadphaseObjectInfo<S>* objects, std::uint32_t n_objects) {
    if (handle_) fclb_broadphase_release(handle_);
    handle_ = 0;
    if (n_objects == 0) return;
    std::vector<S> boxes(6 * std::size_t(n_objects));
    std::vector<std::uint64_t> ids(n_objects);
    for (std::uint32_t i = 0; i < n_objects; i++) {
      for (int k = 0; k < 6; k++) boxes[6 * std::size_t(i) + k] = objects[i].bv[k];
      ids[i] = objects[i].user_id;
    }
    detail::check(fclb_broadphase_build_host(boxes.data(), ids.data(), n_objects, detail::scalarType<S>(), &handle_),
                  "fclb_broadphase_build_host");
  }
  bool UpdateObjectAABB(std::uint64_t user_id, const std::array<S, 6>& new_aabb) {
    return handle_ && fclb_broadphase_update_host(handle_, &user_id, new_aabb.data(), 1) == FCLB_OK;
  }
  template <typename Fn>
  void SelfCollision(const Fn& fn, void* data) const {
    if (!handle_) return;
    report(fn, data, [&](std::uint64_t* out, std::size_t cap, std::size_t* n) {
      return fclb_broadphase_self_pairs_host(handle_, out, cap, n);
    });
  }
  template <typename Fn>
  void TreeCollision(const BroadphaseAABB_Tree& tree2, const Fn& fn, void* data) const {
    if (!handle_ || !tree2.handle_) return;
    report(fn, data, [&](std::uint64_t* out, std::size_t cap, std::size_t* n) {
      return fclb_broadphase_tree_pairs_host(handle_, tree2.handle_, out, cap, n);
    });
  }
  template <typename Fn>
  void SingleObjectCollision(const BroadphaseObjectInfo<S>& object, const Fn& fn, void* data) const {
    if (!handle_) return;
    report(fn, data, [&](std::uint64_t* out, std::size_t cap, std::size_t* n) {
      return fclb_broadphase_query_pairs_host(handle_, object.bv.data(), &object.user_id, 1, out, cap, n);
    });
  }

 private:
  template <typename Fn, typename Query>
  static void report(const Fn& fn, void* data, const Query& query) {
    std::size_t n = 0;
    detail::check(query(nullptr, 0, &n), "broadphase pair count");
    std::vector<std::uint64_t> pairs(2 * n);
    if (n) detail::check(query(pairs.data(), n, &n), "broadphase pairs");
    for (std::size_t i = 0; i < n; i++)
      if (fn(pairs[2 * i], pairs[2 * i + 1], data)) return;
  }
  fclb_handle handle_ = 0;
};

// ---- batched entry points --------------------------------------------------------
// One query = (o1, tf1, o2, tf2).  Results: one CollisionResult per query.
template <typename S>
struct CollisionQuery {
  const CollisionGeometry<S>* o1;
  Transform3<S> tf1;
  const CollisionGeometry<S>* o2;
  Transform3<S> tf2;
};

namespace detail {
// batched C-ABI calls issued so far (tests: a batch of queries on one geometry pair must cost ONE call)
inline std::size_t& abiBatchCalls() {
  static std::size_t n = 0;
  return n;
}
// rc of a batched call: UNSUPPORTED / BAD_ARG / CAPACITY are reported the reference's way (a warning, no contacts,
// collision-inl.h:91-110); a missing device or a CUDA fault aborts (there is no CPU path to fall back to)
inline bool batchOk(int rc, const char* what) {
  abiBatchCalls() += 1;
  if (rc == FCLB_OK) return true;
  if (rc == FCLB_ERR_NO_DEVICE || rc == FCLB_ERR_CUDA) check(rc, what);
  std::cerr << "Warning: " << what << ": " << fclb_last_error() << std::endl;
  return false;
}
template <typename S>
inline int sceneKind(const CollisionGeometry<S>* g) {
  return g->getNodeType() == BV_OBBRSS ? FCLB_SCENE_BVH : (g->getNodeType() == GEOM_HEIGHTMAP ? FCLB_SCENE_HEIGHTMAP : FCLB_SCENE_OCTREE);
}
template <typename S>
inline fclb_handle sceneHandle(const CollisionGeometry<S>* g) {
  if (g->getNodeType() == BV_OBBRSS) return static_cast<const BVHModel<OBBRSS<S>>*>(g)->handle();
  if (g->getNodeType() == GEOM_HEIGHTMAP) return static_cast<const HeightMapCollisionGeometry<S>*>(g)->handle();
  return static_cast<const Octree2CollisionGeometry<S>*>(g)->handle();
}
// contacts kept per query by one batched call; a query that reports more is served by a second call sized to fit
inline uint32_t firstKeep(uint32_t max_contacts) { return max_contacts < 8 ? max_contacts : 8; }
}  // namespace detail

// Every query of the batch goes to the device in as few C-ABI calls as its geometry allows: all shape-shape queries in
// ONE call (one shape table), every other query grouped by its (geometry, geometry) handle pair with ONE call per
// group.  Dispatch follows the reference's table (collision_func_matrix-inl.h:720-857) incl. its argument swaps:
//   (Shape, BVH)        collide() swaps the arguments (collision-inl.h:91-100): o1 = mesh, the normal is the one
//                       MeshShapeIntersect writes (reverse_normal = true, bvh_solver-inl.h:55-66)
//   (Shape, HeightMap), (Shape, Octree2), (BVH, HeightMap), (BVH, Octree2), (Octree2, HeightMap)
//                       forward to the scene-first solver, which writes o1 = the heightmap / octree with no reversal
//                       (collision_func_matrix-inl.h:60-78,774-792,815-833)
template <typename S>
void collideBatch(const std::vector<CollisionQuery<S>>& queries, const CollisionRequest<S>& request,
                  std::vector<CollisionResult<S>>& results) {
  const std::size_t n = queries.size();
  results.assign(n, CollisionResult<S>());
  if (request.maxNumContacts() == 0) {
    std::cerr << "Warning: should stop early as num_max_contact is " << request.maxNumContacts() << " !" << std::endl;
    return;
  }
  if (n == 0) return;
  const fclb_request req = request.toAbi();
  const int st = detail::scalarType<S>();
  const bool pen = req.penetration_mode != FCLB_PEN_DISABLED;
  const bool mpr_pen = req.penetration_mode == FCLB_PEN_DIRECTED || req.penetration_mode == FCLB_PEN_INCREMENTAL_MIN;

  // ---- classify -----------------------------------------------------------------------------------------------
  struct Group {
    int kind = 0;  // 0 mesh-mesh, 1 scene-shape, 2 scene-scene
    int k1 = 0, k2 = 0;
    fclb_handle h1 = 0, h2 = 0;
    const CollisionGeometry<S>*g1 = nullptr, *g2 = nullptr;
    bool swapped = false;  // scene-scene: the device call has the arguments in the other order
    std::vector<std::size_t> q;
  };
  std::vector<Group> groups;
  auto groupOf = [&](int kind, int k1, fclb_handle h1, int k2, fclb_handle h2, bool swapped, const CollisionGeometry<S>* g1,
                     const CollisionGeometry<S>* g2) -> Group& {
    for (auto& g : groups)
      if (g.kind == kind && g.k1 == k1 && g.h1 == h1 && g.k2 == k2 && g.h2 == h2 && g.swapped == swapped) return g;
    groups.emplace_back();
    Group& g = groups.back();
    g.kind = kind; g.k1 = k1; g.h1 = h1; g.k2 = k2; g.h2 = h2; g.swapped = swapped; g.g1 = g1; g.g2 = g2;
    return g;
  };
  std::vector<std::size_t> shape_q;
  for (std::size_t q = 0; q < n; q++) {
    const auto& Q = queries[q];
    const bool s1 = Q.o1->isShape(), s2 = Q.o2->isShape();
    if (s1 && s2) {
      shape_q.push_back(q);
    } else if (!s1 && !s2) {
      const int k1 = detail::sceneKind(Q.o1), k2 = detail::sceneKind(Q.o2);
      if (k1 == FCLB_SCENE_BVH && k2 == FCLB_SCENE_BVH) {
        groupOf(0, k1, detail::sceneHandle(Q.o1), k2, detail::sceneHandle(Q.o2), false, Q.o1, Q.o2).q.push_back(q);
      } else {
        auto rank = [](int k) { return k == FCLB_SCENE_HEIGHTMAP ? 0 : (k == FCLB_SCENE_OCTREE ? 1 : 2); };
        const bool swap = rank(k1) > rank(k2);
        const CollisionGeometry<S>* a = swap ? Q.o2 : Q.o1;
        const CollisionGeometry<S>* b = swap ? Q.o1 : Q.o2;
        groupOf(2, detail::sceneKind(a), detail::sceneHandle(a), detail::sceneKind(b), detail::sceneHandle(b), swap, a, b).q.push_back(q);
      }
    } else {
      const CollisionGeometry<S>* scene = s1 ? Q.o2 : Q.o1;
      groupOf(1, detail::sceneKind(scene), detail::sceneHandle(scene), -1, 0, false, scene, nullptr).q.push_back(q);
    }
  }

  // ---- shape-shape: one table (one entry per distinct geometry object), one call ---------------------------------
  std::vector<fclb_shape> shapes;
  std::vector<const CollisionGeometry<S>*> shape_objs;
  auto shapeIndex = [&](const CollisionGeometry<S>* g) -> uint32_t {
    for (std::size_t i = shape_objs.size(); i-- > 0 && shape_objs.size() - i <= 64;)  // recent objects first (bounded scan)
      if (shape_objs[i] == g) return uint32_t(i);
    shape_objs.push_back(g);
    shapes.push_back(g->shapeRecord());
    return uint32_t(shapes.size() - 1);
  };
  if (!shape_q.empty()) {
    const std::size_t m = shape_q.size();
    std::vector<fclb_pair> pairs(m);
    std::vector<S> p1(12 * m), p2(12 * m);
    for (std::size_t i = 0; i < m; i++) {
      const auto& Q = queries[shape_q[i]];
      pairs[i] = fclb_pair{shapeIndex(Q.o1), shapeIndex(Q.o2)};
      Q.tf1.toPose12(&p1[12 * i]);
      Q.tf2.toPose12(&p2[12 * i]);
    }
    fclb_handle table = 0;
    detail::check(fclb_shapes_upload(shapes.data(), uint32_t(shapes.size()), &table), "fclb_shapes_upload");
    const uint32_t keep = req.max_contacts < 4 ? req.max_contacts : 4;  // a shape pair has at most four contacts (boxBox2)
    std::vector<S> contacts(m * keep * 9);
    std::vector<uint32_t> counts(m, 0);
    if (detail::batchOk(fclb_collide_batch_host(table, pairs.data(), p1.data(), p2.data(), m, st, &req, keep, contacts.data(),
                                                counts.data()),
                        "fclb_collide_batch_host")) {
      for (std::size_t i = 0; i < m; i++) {
        const std::size_t q = shape_q[i];
        for (uint32_t c = 0; c < counts[i] && c < keep; c++) {
          const S* r = &contacts[(i * keep + c) * 9];
          Contact<S> ct;
          ct.o1 = queries[q].o1;
          ct.o2 = queries[q].o2;
          if (pen) {
            ct.normal = Vector3<S>(r[2], r[3], r[4]);
            ct.pos = Vector3<S>(r[5], r[6], r[7]);
            ct.penetration_depth = r[8];
          }
          results[q].addContact(ct);
        }
      }
    }
    fclb_release(table);
    shapes.clear();
    shape_objs.clear();
  }

  // ---- one call per (geometry, geometry) group --------------------------------------------------------------------
  for (const Group& g : groups) {
    const std::size_t m = g.q.size();
    std::vector<S> pa(12 * m), pb(12 * m);
    std::vector<uint32_t> counts(m, 0);
    uint32_t keep = detail::firstKeep(req.max_contacts);
    if (g.kind == 0) {  // BVHModel<OBBRSS> x BVHModel<OBBRSS> (collision_func_matrix-inl.h:544-572)
      for (std::size_t i = 0; i < m; i++) {
        queries[g.q[i]].tf1.toPose12(&pa[12 * i]);
        queries[g.q[i]].tf2.toPose12(&pb[12 * i]);
      }
      std::vector<int32_t> ids;
      std::vector<S> rec;
      bool ok = true;
      for (int pass = 0; pass < 2 && ok; pass++) {
        ids.assign(m * keep * 2, -1);
        rec.assign(m * keep * 7, S(0));
        ok = detail::batchOk(fclb_bvh_collide_contacts_batch_host(g.h1, g.h2, pa.data(), pb.data(), m, st, &req, keep, counts.data(),
                                                                  ids.data(), rec.data()),
                             "fclb_bvh_collide_contacts_batch_host");
        const uint32_t most = ok && m ? *std::max_element(counts.begin(), counts.end()) : 0;
        if (most <= keep) break;
        keep = most;
      }
      if (!ok) continue;
      for (std::size_t i = 0; i < m; i++)
        for (uint32_t c = 0; c < counts[i] && c < keep; c++) {
          Contact<S> ct;
          ct.o1 = g.g1;
          ct.o2 = g.g2;
          ct.b1 = ids[(i * keep + c) * 2];
          ct.b2 = ids[(i * keep + c) * 2 + 1];
          if (pen || mpr_pen) {
            const S* r = &rec[(i * keep + c) * 7];
            ct.normal = Vector3<S>(r[0], r[1], r[2]);
            ct.pos = Vector3<S>(r[3], r[4], r[5]);
            ct.penetration_depth = r[6];
          }
          results[g.q[i]].addContact(ct);
        }
    } else if (g.kind == 1) {  // scene geometry x shape, either argument order
      std::vector<uint32_t> sid(m);
      for (std::size_t i = 0; i < m; i++) {
        const auto& Q = queries[g.q[i]];
        const bool shape_first = Q.o1->isShape();
        const CollisionGeometry<S>* shape = shape_first ? Q.o1 : Q.o2;
        sid[i] = shapeIndex(shape);
        (shape_first ? Q.tf2 : Q.tf1).toPose12(&pa[12 * i]);
        (shape_first ? Q.tf1 : Q.tf2).toPose12(&pb[12 * i]);
      }
      fclb_handle table = 0;
      detail::check(fclb_shapes_upload(shapes.data(), uint32_t(shapes.size()), &table), "fclb_shapes_upload");
      std::vector<int64_t> b1;
      std::vector<S> rec;
      bool ok = true;
      for (int pass = 0; pass < 2 && ok; pass++) {
        b1.assign(m * keep, -1);
        rec.assign(m * keep * 7, S(0));
        ok = detail::batchOk(fclb_scene_shape_contacts_batch_host(g.k1, g.h1, table, sid.data(), pa.data(), pb.data(), m, st, &req,
                                                                  keep, counts.data(), b1.data(), rec.data()),
                             "fclb_scene_shape_contacts_batch_host");
        const uint32_t most = ok && m ? *std::max_element(counts.begin(), counts.end()) : 0;
        if (most <= keep) break;
        keep = most;
      }
      fclb_release(table);
      shapes.clear();
      shape_objs.clear();
      if (!ok) continue;
      for (std::size_t i = 0; i < m; i++) {
        const auto& Q = queries[g.q[i]];
        const CollisionGeometry<S>* shape = Q.o1->isShape() ? Q.o1 : Q.o2;
        for (uint32_t c = 0; c < counts[i] && c < keep; c++) {
          Contact<S> ct;
          ct.o1 = g.g1;  // the mesh / heightmap / octree, whatever the argument order (see above)
          ct.o2 = shape;
          ct.b1 = intptr_t(b1[i * keep + c]);
          if (pen) {
            const S* r = &rec[(i * keep + c) * 7];
            ct.normal = Vector3<S>(r[0], r[1], r[2]);
            ct.pos = Vector3<S>(r[3], r[4], r[5]);
            ct.penetration_depth = r[6];
          }
          results[g.q[i]].addContact(ct);
        }
      }
    } else {  // heightmap / octree against heightmap / octree / mesh (collision_func_matrix-inl.h:794-812, 835-856)
      for (std::size_t i = 0; i < m; i++) {
        const auto& Q = queries[g.q[i]];
        (g.swapped ? Q.tf2 : Q.tf1).toPose12(&pa[12 * i]);
        (g.swapped ? Q.tf1 : Q.tf2).toPose12(&pb[12 * i]);
      }
      // collisionPenetrationMPR (collision_penetration-inl.h:189-252).  In the canonical argument order the records are
      // the reference's bit for bit.  With swapped arguments the reference runs MPR on (leaf of o1, leaf of o2) for the
      // escape direction of o2; here the canonical pair is evaluated for the opposite direction and the normal negated
      // -- the same penetration up to MPR's argument order, not bit-identical.
      fclb_request r2 = req;
      if (mpr_pen && g.swapped)
        for (int k = 0; k < 3; k++) r2.dir[k] = -req.dir[k];
      std::vector<int64_t> b1, b2;
      std::vector<S> rec;
      bool ok = true;
      for (int pass = 0; pass < 2 && ok; pass++) {
        b1.assign(m * keep, -1);
        b2.assign(m * keep, -1);
        if (pen) {
          rec.assign(m * keep * 7, S(0));
          ok = detail::batchOk(fclb_scene_pair_contacts_batch_host(g.k1, g.h1, g.k2, g.h2, pa.data(), pb.data(), m, st, &r2, keep,
                                                                   counts.data(), b1.data(), b2.data(), rec.data()),
                               "fclb_scene_pair_contacts_batch_host");
        } else {
          ok = detail::batchOk(fclb_scene_pair_collide_batch_host(g.k1, g.h1, g.k2, g.h2, pa.data(), pb.data(), m, st, &req, keep,
                                                                  counts.data(), b1.data(), b2.data()),
                               "fclb_scene_pair_collide_batch_host");
        }
        const uint32_t most = ok && m ? *std::max_element(counts.begin(), counts.end()) : 0;
        if (most <= keep) break;
        keep = most;
      }
      if (!ok) continue;
      for (std::size_t i = 0; i < m; i++)
        for (uint32_t c = 0; c < counts[i] && c < keep; c++) {
          Contact<S> ct;
          ct.o1 = g.g1;
          ct.o2 = g.g2;
          ct.b1 = intptr_t(b1[i * keep + c]);
          ct.b2 = intptr_t(b2[i * keep + c]);
          if (pen) {
            const S* r = &rec[(i * keep + c) * 7];
            const S sgn = (mpr_pen && g.swapped) ? S(-1) : S(1);
            ct.normal = Vector3<S>(sgn * r[0], sgn * r[1], sgn * r[2]);
            ct.pos = Vector3<S>(r[3], r[4], r[5]);
            ct.penetration_depth = r[6];
          }
          results[g.q[i]].addContact(ct);
        }
    }
  }
}

// CollisionResult with a UserContactProcessFunctor (narrowphase/collision_result.h:56-73): callbacks cannot cross the C ABI,
// so the device returns the contacts and the functor runs on the host over them, in order -- keep decides whether the
// contact is stored, stop ends the query.  The functor sees up to `device_contacts` contacts per query.
template <typename S>
using UserContactProcessFunctor = std::function<void(const Contact<S>& contact, bool& keep, bool& stop)>;
template <typename S>
void collideBatch(const std::vector<CollisionQuery<S>>& queries, const CollisionRequest<S>& request,
                  const UserContactProcessFunctor<S>& functor, std::vector<CollisionResult<S>>& results,
                  std::size_t device_contacts = 1024) {
  CollisionRequest<S> wide = request;
  wide.setMaxContactCount(std::max(request.maxNumContacts(), device_contacts));
  std::vector<CollisionResult<S>> all;
  collideBatch(queries, wide, all);
  results.assign(queries.size(), CollisionResult<S>());
  for (std::size_t q = 0; q < queries.size(); q++)
    for (const auto& c : all[q].getContacts()) {
      bool keep = true, stop = false;
      functor(c, keep, stop);
      if (keep) results[q].addContact(c);
      // terminationConditionSatisfied (collision_result-inl.h:84-98)
      if (stop || (results[q].numContacts() > 0 && results[q].numContacts() >= request.maxNumContacts())) break;
    }
}

// fcl::collide, single pair (reference narrowphase/collision_interface-inl.h:13-32)
template <typename S>
std::size_t collide(const CollisionGeometry<S>* o1, const Transform3<S>& tf1, const CollisionGeometry<S>* o2,
                    const Transform3<S>& tf2, const CollisionRequest<S>& request, CollisionResult<S>& result) {
  std::vector<CollisionQuery<S>> q{{o1, tf1, o2, tf2}};
  std::vector<CollisionResult<S>> r;
  collideBatch(q, request, r);
  for (const auto& c : r[0].getContacts()) result.addContact(c);
  return result.numContacts();
}
template <typename S>
std::size_t collide(const CollisionObject<S>* o1, const CollisionObject<S>* o2, const CollisionRequest<S>& request,
                    CollisionResult<S>& result) {
  return collide(o1->collisionGeometry().get(), o1->getTransform(), o2->collisionGeometry().get(), o2->getTransform(),
                 request, result);
}

// ---- translational continuous collision (narrowphase/continuous_collision.h:15-21), shape pairs -------------------------
// detail/ccd/ccd_typedef.h, ccd_request.h, ccd_contact.h, ccd_result.h
template <typename S>
struct TranslationalDisplacement {
  Vector3<S> unit_axis_in_shape1;
  S scalar_displacement{0};
};
template <typename S>
struct Interval {
  S lower_bound{S(-1)};
  S upper_bound{S(-1)};
};
enum class TimeOfCollisionRequestType { kNotRequested, kBoxApproximate, kOneTocSample };
template <typename S>
struct ContinuousCollisionRequest {
  TimeOfCollisionRequestType request_type{TimeOfCollisionRequestType::kNotRequested};
  std::size_t num_max_contacts{1};
  S zero_movement_tolerance{S(1e-4)};
  S gjk_tolerance{S(1e-6)};
  int max_gjk_iterations{128};
};
// math/bv/AABB.h: the two corners, as far as the contacts need them
template <typename S>
struct AABB {
  Vector3<S> min_{S(0), S(0), S(0)};
  Vector3<S> max_{S(0), S(0), S(0)};
};
template <typename S>
struct ContinuousCollisionContact {
  const CollisionGeometry<S>* o1{nullptr};
  const CollisionGeometry<S>* o2{nullptr};
  static constexpr std::int64_t NONE = -1;
  std::int64_t b1{NONE};
  std::int64_t b2{NONE};
  AABB<S> o1_bv{};  // heightmap / octree contacts: the pixel's / node's box (ccd_contact.h:24-30)
  AABB<S> o2_bv{};
  Interval<S> toc{};
};
template <typename S>
struct ContinuousCollisionResult {
  void AddContact(const ContinuousCollisionContact<S>& c) { contacts_.push_back(c); }
  void ClearContact() { contacts_.clear(); }
  std::size_t num_contacts() const { return contacts_.size(); }
  const std::vector<ContinuousCollisionContact<S>>& raw_contacts() const { return contacts_; }

 private:
  std::vector<ContinuousCollisionContact<S>> contacts_;
};
template <typename S>
struct ContinuousCollisionQuery {
  const CollisionGeometry<S>* o1;
  Transform3<S> tf1;
  TranslationalDisplacement<S> o1_displacement;
  const CollisionGeometry<S>* o2;
  Transform3<S> tf2;
};
// one C-ABI call for the shape pairs of the batch, one per mesh for its (shape, mesh) / (mesh, shape) queries and one per
// pair of meshes, one per heightmap / octree for its shape queries
// (the reference's CCD matrix serves OBB trees, translational_collision_func_matrix-inl.h:469-489; our BVHModel<OBBRSS>
// has the same hierarchy, see fclb_translational_ccd_mesh_batch_host); one per (heightmap / octree, mesh) pair and
// argument order and one per pair of heightmaps / octrees
template <typename S>
void translationalCcdBatch(const std::vector<ContinuousCollisionQuery<S>>& queries, const ContinuousCollisionRequest<S>& request,
                           std::vector<ContinuousCollisionResult<S>>& results) {
  const std::size_t n = queries.size();
  results.assign(n, ContinuousCollisionResult<S>());
  if (n == 0 || request.num_max_contacts == 0) return;
  fclb_ccd_request rq{};
  rq.request_type = uint32_t(request.request_type);
  rq.max_contacts = uint32_t(std::min<std::size_t>(request.num_max_contacts, 0xffffffffu));
  rq.zero_movement_tolerance = double(request.zero_movement_tolerance);
  rq.gjk_tolerance = double(request.gjk_tolerance);
  rq.max_gjk_iterations = request.max_gjk_iterations;
  std::vector<fclb_shape> shapes;
  std::vector<fclb_pair> pairs;
  std::vector<S> p1, p2, disp;
  std::vector<std::size_t> idx;
  struct MeshGroup {
    bool mesh_moves;
    int kind = FCLB_SCENE_BVH;
    std::vector<fclb_shape> shapes;
    std::vector<uint32_t> ids;
    std::vector<S> pose_shape, pose_mesh, disp;
    std::vector<std::size_t> idx;
  };
  std::map<std::pair<fclb_handle, bool>, MeshGroup> meshes, scenes;
  struct PairGroup {
    std::vector<S> pose1, pose2, disp;
    std::vector<std::size_t> idx;
  };
  std::map<std::pair<fclb_handle, fclb_handle>, PairGroup> mesh_pairs;
  struct ScenePairGroup : PairGroup {
    int kind1 = 0, kind2 = 0;
    bool mesh_moves = false;  // (scene, mesh) groups only
  };
  std::map<std::pair<std::pair<fclb_handle, fclb_handle>, bool>, ScenePairGroup> scene_meshes;  // ((scene, mesh), mesh moves)
  std::map<std::pair<fclb_handle, fclb_handle>, ScenePairGroup> scene_pairs;
  auto push12 = [](std::vector<S>& v, const Transform3<S>& tf) {
    v.resize(v.size() + 12);
    tf.toPose12(&v[v.size() - 12]);
  };
  for (std::size_t q = 0; q < n; q++) {
    const auto& Q = queries[q];
    const bool s1 = Q.o1->isShape(), s2 = Q.o2->isShape();
    const bool m1 = Q.o1->getNodeType() == BV_OBBRSS, m2 = Q.o2->getNodeType() == BV_OBBRSS;
    const bool g1 = Q.o1->getNodeType() == GEOM_HEIGHTMAP || Q.o1->getNodeType() == GEOM_OCTREE2;
    const bool g2 = Q.o2->getNodeType() == GEOM_HEIGHTMAP || Q.o2->getNodeType() == GEOM_OCTREE2;
    if (s1 && s2) {
      pairs.push_back(fclb_pair{uint32_t(shapes.size()), uint32_t(shapes.size() + 1)});
      shapes.push_back(Q.o1->shapeRecord());
      shapes.push_back(Q.o2->shapeRecord());
      push12(p1, Q.tf1);
      push12(p2, Q.tf2);
      for (int k = 0; k < 3; k++) disp.push_back(Q.o1_displacement.unit_axis_in_shape1[k]);
      disp.push_back(Q.o1_displacement.scalar_displacement);
      idx.push_back(q);
    } else if ((s1 && m2) || (m1 && s2)) {
      const bool mesh_moves = m1;
      const CollisionGeometry<S>* shape = s1 ? Q.o1 : Q.o2;
      const CollisionGeometry<S>* mesh = s1 ? Q.o2 : Q.o1;
      MeshGroup& g = meshes[std::make_pair(detail::sceneHandle(mesh), mesh_moves)];
      g.mesh_moves = mesh_moves;
      g.ids.push_back(uint32_t(g.shapes.size()));
      g.shapes.push_back(shape->shapeRecord());
      push12(g.pose_shape, s1 ? Q.tf1 : Q.tf2);
      push12(g.pose_mesh, s1 ? Q.tf2 : Q.tf1);
      for (int k = 0; k < 3; k++) g.disp.push_back(Q.o1_displacement.unit_axis_in_shape1[k]);
      g.disp.push_back(Q.o1_displacement.scalar_displacement);
      g.idx.push_back(q);
    } else if ((s1 && g2) || (g1 && s2)) {  // shape vs heightmap / octree, either order
      const bool scene_moves = g1;
      const CollisionGeometry<S>* shape = s1 ? Q.o1 : Q.o2;
      const CollisionGeometry<S>* scene = s1 ? Q.o2 : Q.o1;
      MeshGroup& g = scenes[std::make_pair(detail::sceneHandle(scene), scene_moves)];
      g.mesh_moves = scene_moves;
      g.kind = detail::sceneKind(scene);
      g.ids.push_back(uint32_t(g.shapes.size()));
      g.shapes.push_back(shape->shapeRecord());
      push12(g.pose_shape, s1 ? Q.tf1 : Q.tf2);
      push12(g.pose_mesh, s1 ? Q.tf2 : Q.tf1);
      for (int k = 0; k < 3; k++) g.disp.push_back(Q.o1_displacement.unit_axis_in_shape1[k]);
      g.disp.push_back(Q.o1_displacement.scalar_displacement);
      g.idx.push_back(q);
    } else if (m1 && m2) {
      PairGroup& g = mesh_pairs[std::make_pair(detail::sceneHandle(Q.o1), detail::sceneHandle(Q.o2))];
      push12(g.pose1, Q.tf1);
      push12(g.pose2, Q.tf2);
      for (int k = 0; k < 3; k++) g.disp.push_back(Q.o1_displacement.unit_axis_in_shape1[k]);
      g.disp.push_back(Q.o1_displacement.scalar_displacement);
      g.idx.push_back(q);
    } else if ((g1 && m2) || (m1 && g2)) {  // heightmap / octree vs mesh, either order
      const CollisionGeometry<S>* scene = g1 ? Q.o1 : Q.o2;
      const CollisionGeometry<S>* mesh = g1 ? Q.o2 : Q.o1;
      ScenePairGroup& g = scene_meshes[std::make_pair(std::make_pair(detail::sceneHandle(scene), detail::sceneHandle(mesh)), m1)];
      g.kind1 = detail::sceneKind(scene);
      g.mesh_moves = m1;
      push12(g.pose1, g1 ? Q.tf1 : Q.tf2);  // scene
      push12(g.pose2, g1 ? Q.tf2 : Q.tf1);  // mesh
      for (int k = 0; k < 3; k++) g.disp.push_back(Q.o1_displacement.unit_axis_in_shape1[k]);
      g.disp.push_back(Q.o1_displacement.scalar_displacement);
      g.idx.push_back(q);
    } else if (g1 && g2) {
      ScenePairGroup& g = scene_pairs[std::make_pair(detail::sceneHandle(Q.o1), detail::sceneHandle(Q.o2))];
      g.kind1 = detail::sceneKind(Q.o1);
      g.kind2 = detail::sceneKind(Q.o2);
      push12(g.pose1, Q.tf1);
      push12(g.pose2, Q.tf2);
      for (int k = 0; k < 3; k++) g.disp.push_back(Q.o1_displacement.unit_axis_in_shape1[k]);
      g.disp.push_back(Q.o1_displacement.scalar_displacement);
      g.idx.push_back(q);
    } else {
      std::cerr << "Warning: collision function between node type " << Q.o1->getNodeType() << " and node type " << Q.o2->getNodeType()
                << " is not supported" << std::endl;
    }
  }
  for (auto& kv : scene_meshes) {  // contacts: o1 = the heightmap / octree, b1 = pixel / node code, o1_bv its box; o2 = the mesh, b2 = triangle
    ScenePairGroup& g = kv.second;
    const std::size_t m = g.idx.size();
    uint32_t keep = uint32_t(std::min<std::size_t>(request.num_max_contacts, 64));
    std::vector<uint32_t> counts(m);
    std::vector<int64_t> ids;
    std::vector<S> toc, box;
    for (int pass = 0; pass < 2; pass++) {
      ids.assign(m * keep * 2, -1);
      toc.assign(m * keep * 2, S(-1));
      box.assign(m * keep * 6, S(0));
      if (!detail::batchOk(fclb_translational_ccd_scene_mesh_batch_host(g.kind1, kv.first.first.first, kv.first.first.second, g.pose1.data(),
                                                                        g.pose2.data(), g.disp.data(), m, detail::scalarType<S>(), &rq,
                                                                        g.mesh_moves ? 1 : 0, keep, counts.data(), ids.data(), toc.data(),
                                                                        box.data()),
                           "fclb_translational_ccd_scene_mesh_batch_host")) {
        counts.assign(m, 0);
        break;
      }
      const uint32_t most = *std::max_element(counts.begin(), counts.end());
      if (most <= keep) break;
      keep = most;
    }
    for (std::size_t i = 0; i < m; i++)
      for (uint32_t k = 0; k < counts[i] && k < keep; k++) {
        const auto& Q = queries[g.idx[i]];
        ContinuousCollisionContact<S> c;
        c.o1 = g.mesh_moves ? Q.o2 : Q.o1;
        c.o2 = g.mesh_moves ? Q.o1 : Q.o2;
        c.b1 = ids[(i * keep + k) * 2];
        c.b2 = ids[(i * keep + k) * 2 + 1];
        c.toc.lower_bound = toc[(i * keep + k) * 2];
        c.toc.upper_bound = toc[(i * keep + k) * 2 + 1];
        for (int j = 0; j < 3; j++) {
          c.o1_bv.min_[j] = box[(i * keep + k) * 6 + j];
          c.o1_bv.max_[j] = box[(i * keep + k) * 6 + 3 + j];
        }
        results[g.idx[i]].AddContact(c);
      }
  }
  for (auto& kv : scene_pairs) {
    ScenePairGroup& g = kv.second;
    const std::size_t m = g.idx.size();
    uint32_t keep = uint32_t(std::min<std::size_t>(request.num_max_contacts, 64));
    std::vector<uint32_t> counts(m);
    std::vector<int64_t> ids;
    std::vector<S> toc, box;
    for (int pass = 0; pass < 2; pass++) {
      ids.assign(m * keep * 2, -1);
      toc.assign(m * keep * 2, S(-1));
      box.assign(m * keep * 12, S(0));
      if (!detail::batchOk(fclb_translational_ccd_scene_pair_batch_host(g.kind1, kv.first.first, g.kind2, kv.first.second, g.pose1.data(),
                                                                        g.pose2.data(), g.disp.data(), m, detail::scalarType<S>(), &rq,
                                                                        keep, counts.data(), ids.data(), toc.data(), box.data()),
                           "fclb_translational_ccd_scene_pair_batch_host")) {
        counts.assign(m, 0);
        break;
      }
      const uint32_t most = *std::max_element(counts.begin(), counts.end());
      if (most <= keep) break;
      keep = most;
    }
    // the reference names the heightmap o1 when it is handed (octree, heightmap) (RunOctreeHeightMap); every other pair keeps
    // the caller's order
    const bool swap = g.kind1 == FCLB_SCENE_OCTREE && g.kind2 == FCLB_SCENE_HEIGHTMAP;
    for (std::size_t i = 0; i < m; i++)
      for (uint32_t k = 0; k < counts[i] && k < keep; k++) {
        const auto& Q = queries[g.idx[i]];
        const std::size_t o = i * keep + k;
        ContinuousCollisionContact<S> c;
        c.o1 = swap ? Q.o2 : Q.o1;
        c.o2 = swap ? Q.o1 : Q.o2;
        c.b1 = ids[o * 2 + (swap ? 1 : 0)];
        c.b2 = ids[o * 2 + (swap ? 0 : 1)];
        c.toc.lower_bound = toc[o * 2];
        c.toc.upper_bound = toc[o * 2 + 1];
        for (int j = 0; j < 3; j++) {
          c.o1_bv.min_[j] = box[o * 12 + (swap ? 6 : 0) + j];
          c.o1_bv.max_[j] = box[o * 12 + (swap ? 6 : 0) + 3 + j];
          c.o2_bv.min_[j] = box[o * 12 + (swap ? 0 : 6) + j];
          c.o2_bv.max_[j] = box[o * 12 + (swap ? 0 : 6) + 3 + j];
        }
        results[g.idx[i]].AddContact(c);
      }
  }
  for (auto& kv : mesh_pairs) {
    PairGroup& g = kv.second;
    const std::size_t m = g.idx.size();
    uint32_t keep = uint32_t(std::min<std::size_t>(request.num_max_contacts, 64));
    std::vector<uint32_t> counts(m);
    std::vector<int64_t> prim;
    std::vector<S> toc;
    for (int pass = 0; pass < 2; pass++) {
      prim.assign(m * keep * 2, -1);
      toc.assign(m * keep * 2, S(-1));
      if (!detail::batchOk(fclb_translational_ccd_mesh_pair_batch_host(kv.first.first, kv.first.second, g.pose1.data(), g.pose2.data(),
                                                                       g.disp.data(), m, detail::scalarType<S>(), &rq, keep,
                                                                       counts.data(), prim.data(), toc.data()),
                           "fclb_translational_ccd_mesh_pair_batch_host")) {
        counts.assign(m, 0);
        break;
      }
      const uint32_t most = *std::max_element(counts.begin(), counts.end());
      if (most <= keep) break;
      keep = most;
    }
    for (std::size_t i = 0; i < m; i++)
      for (uint32_t k = 0; k < counts[i] && k < keep; k++) {
        const auto& Q = queries[g.idx[i]];
        ContinuousCollisionContact<S> c;
        c.o1 = Q.o1;
        c.o2 = Q.o2;
        c.b1 = prim[(i * keep + k) * 2];
        c.b2 = prim[(i * keep + k) * 2 + 1];
        c.toc.lower_bound = toc[(i * keep + k) * 2];
        c.toc.upper_bound = toc[(i * keep + k) * 2 + 1];
        results[g.idx[i]].AddContact(c);
      }
  }
  if (!pairs.empty()) {
    fclb_handle table = 0;
    detail::check(fclb_shapes_upload(shapes.data(), uint32_t(shapes.size()), &table), "fclb_shapes_upload");
    std::vector<uint8_t> hit(pairs.size());
    std::vector<S> toc(2 * pairs.size());
    if (detail::batchOk(fclb_translational_ccd_batch_host(table, pairs.data(), p1.data(), p2.data(), disp.data(), pairs.size(),
                                                          detail::scalarType<S>(), &rq, hit.data(), toc.data()),
                        "fclb_translational_ccd_batch_host"))
      for (std::size_t i = 0; i < pairs.size(); i++)
        if (hit[i]) {
          ContinuousCollisionContact<S> c;
          c.o1 = queries[idx[i]].o1;
          c.o2 = queries[idx[i]].o2;
          c.toc.lower_bound = toc[2 * i];
          c.toc.upper_bound = toc[2 * i + 1];
          results[idx[i]].AddContact(c);
        }
    fclb_release(table);
  }
  for (auto& kv : scenes) {  // heightmap / octree: the contact names the pixel / node (b2) and carries its box (o2_bv)
    MeshGroup& g = kv.second;
    const std::size_t m = g.ids.size();
    fclb_handle table = 0;
    detail::check(fclb_shapes_upload(g.shapes.data(), uint32_t(g.shapes.size()), &table), "fclb_shapes_upload");
    uint32_t keep = uint32_t(std::min<std::size_t>(request.num_max_contacts, 64));
    std::vector<uint32_t> counts(m);
    std::vector<int64_t> code;
    std::vector<S> toc, box;
    for (int pass = 0; pass < 2; pass++) {
      code.assign(m * keep, -1);
      toc.assign(m * keep * 2, S(-1));
      box.assign(m * keep * 6, S(0));
      if (!detail::batchOk(fclb_translational_ccd_scene_batch_host(g.kind, kv.first.first, table, g.ids.data(), g.pose_shape.data(),
                                                                   g.pose_mesh.data(), g.disp.data(), m, detail::scalarType<S>(), &rq,
                                                                   g.mesh_moves ? 1 : 0, keep, counts.data(), code.data(), toc.data(),
                                                                   box.data()),
                           "fclb_translational_ccd_scene_batch_host")) {
        counts.assign(m, 0);
        break;
      }
      const uint32_t most = *std::max_element(counts.begin(), counts.end());
      if (most <= keep) break;
      keep = most;
    }
    for (std::size_t i = 0; i < m; i++)
      for (uint32_t k = 0; k < counts[i] && k < keep; k++) {
        const auto& Q = queries[g.idx[i]];
        ContinuousCollisionContact<S> c;
        c.o1 = g.mesh_moves ? Q.o2 : Q.o1;  // both entries report o1 = the shape, o2 = the scene geometry
        c.o2 = g.mesh_moves ? Q.o1 : Q.o2;
        c.b2 = code[i * keep + k];
        c.toc.lower_bound = toc[(i * keep + k) * 2];
        c.toc.upper_bound = toc[(i * keep + k) * 2 + 1];
        for (int j = 0; j < 3; j++) {
          c.o2_bv.min_[j] = box[(i * keep + k) * 6 + j];
          c.o2_bv.max_[j] = box[(i * keep + k) * 6 + 3 + j];
        }
        results[g.idx[i]].AddContact(c);
      }
    fclb_release(table);
  }
  for (auto& kv : meshes) {
    MeshGroup& g = kv.second;
    const std::size_t m = g.ids.size();
    fclb_handle table = 0;
    detail::check(fclb_shapes_upload(g.shapes.data(), uint32_t(g.shapes.size()), &table), "fclb_shapes_upload");
    uint32_t keep = uint32_t(std::min<std::size_t>(request.num_max_contacts, 64));
    std::vector<uint32_t> counts(m);
    std::vector<int64_t> prim;
    std::vector<S> toc;
    for (int pass = 0; pass < 2; pass++) {  // a second pass with room for the largest count, when 64 were not enough
      prim.assign(m * keep, -1);
      toc.assign(m * keep * 2, S(-1));
      if (!detail::batchOk(fclb_translational_ccd_mesh_batch_host(kv.first.first, table, g.ids.data(), g.pose_shape.data(),
                                                                  g.pose_mesh.data(), g.disp.data(), m, detail::scalarType<S>(), &rq,
                                                                  g.mesh_moves ? 1 : 0, keep, counts.data(), prim.data(), toc.data()),
                           "fclb_translational_ccd_mesh_batch_host")) {
        counts.assign(m, 0);
        break;
      }
      const uint32_t most = *std::max_element(counts.begin(), counts.end());
      if (most <= keep) break;
      keep = most;
    }
    for (std::size_t i = 0; i < m; i++)
      for (uint32_t k = 0; k < counts[i] && k < keep; k++) {
        // both matrix entries report o1 = the shape, o2 = the mesh, b2 = triangle id (bvh_ccd_solver-inl.h:199-204, :572-584)
        const auto& Q = queries[g.idx[i]];
        ContinuousCollisionContact<S> c;
        c.o1 = g.mesh_moves ? Q.o2 : Q.o1;
        c.o2 = g.mesh_moves ? Q.o1 : Q.o2;
        c.b2 = prim[i * keep + k];
        c.toc.lower_bound = toc[(i * keep + k) * 2];
        c.toc.upper_bound = toc[(i * keep + k) * 2 + 1];
        results[g.idx[i]].AddContact(c);
      }
    fclb_release(table);
  }
}
template <typename S>
void translational_ccd(const CollisionGeometry<S>* o1, const Transform3<S>& tf1, const TranslationalDisplacement<S>& o1_displacement,
                       const CollisionGeometry<S>* o2, const Transform3<S>& tf2, const ContinuousCollisionRequest<S>& request,
                       ContinuousCollisionResult<S>& result) {
  std::vector<ContinuousCollisionQuery<S>> q{{o1, tf1, o1_displacement, o2, tf2}};
  std::vector<ContinuousCollisionResult<S>> r;
  translationalCcdBatch(q, request, r);
  for (const auto& c : r[0].raw_contacts()) result.AddContact(c);
}

// ---- distance (added API; semantics of detail::GJKSolver<S>::shapeDistance) -----------
template <typename S>
struct DistanceRequest {
  S gjk_tolerance = S(0);        // <= 0: constants<S>::gjk_default_tolerance()
  uint32_t gjk_max_iterations = 0;  // 0: 128
  // true: detail::GJKSolver<S>::shapeSignedDistance (gjk_solver-inl.h:810-880) -- generic GJK, and for
  // penetrating pairs min_distance = -(EPA depth) with the EPA witness points
  bool enable_signed_distance = false;
};
template <typename S>
struct DistanceResult {
  bool separated = false;  // shapeDistance's return value
  S min_distance = S(-1);  // -1 when not separated
  std::array<Vector3<S>, 2> nearest_points;
};

template <typename S>
void distanceBatch(const std::vector<CollisionQuery<S>>& queries, const DistanceRequest<S>& request,
                   std::vector<DistanceResult<S>>& results) {
  const std::size_t n = queries.size();
  results.assign(n, DistanceResult<S>());
  if (n == 0) return;
  std::vector<fclb_shape> shapes;
  std::vector<fclb_pair> pairs(n);
  std::vector<S> p1(12 * n), p2(12 * n), dist(n), w1(3 * n), w2(3 * n);
  std::vector<uint8_t> ok(n);
  for (std::size_t q = 0; q < n; q++) {
    pairs[q] = fclb_pair{uint32_t(shapes.size()), uint32_t(shapes.size() + 1)};
    shapes.push_back(queries[q].o1->shapeRecord());
    shapes.push_back(queries[q].o2->shapeRecord());
    queries[q].tf1.toPose12(&p1[12 * q]);
    queries[q].tf2.toPose12(&p2[12 * q]);
  }
  fclb_handle table = 0;
  detail::check(fclb_shapes_upload(shapes.data(), uint32_t(shapes.size()), &table), "fclb_shapes_upload");
  if (request.enable_signed_distance)
    detail::check(fclb_signed_distance_batch_host(table, pairs.data(), p1.data(), p2.data(), n, detail::scalarType<S>(),
                                                  dist.data(), w1.data(), w2.data(), ok.data()),
                  "fclb_signed_distance_batch_host");
  else
    detail::check(fclb_distance_batch_host(table, pairs.data(), p1.data(), p2.data(), n, detail::scalarType<S>(),
                                           double(request.gjk_tolerance), request.gjk_max_iterations, dist.data(),
                                           w1.data(), w2.data(), ok.data()),
                  "fclb_distance_batch_host");
  fclb_release(table);
  for (std::size_t q = 0; q < n; q++) {
    results[q].separated = ok[q] != 0;
    results[q].min_distance = dist[q];
    results[q].nearest_points[0] = Vector3<S>(w1[3 * q], w1[3 * q + 1], w1[3 * q + 2]);
    results[q].nearest_points[1] = Vector3<S>(w2[3 * q], w2[3 * q + 1], w2[3 * q + 2]);
  }
}

template <typename S>
S distance(const CollisionGeometry<S>* o1, const Transform3<S>& tf1, const CollisionGeometry<S>* o2,
           const Transform3<S>& tf2, const DistanceRequest<S>& request, DistanceResult<S>& result) {
  std::vector<CollisionQuery<S>> q{{o1, tf1, o2, tf2}};
  std::vector<DistanceResult<S>> r;
  distanceBatch(q, request, r);
  result = r[0];
  return result.min_distance;
}

using CollisionRequestf = CollisionRequest<float>;
using CollisionRequestd = CollisionRequest<double>;
using CollisionResultf = CollisionResult<float>;
using CollisionResultd = CollisionResult<double>;

}  // namespace fcl
