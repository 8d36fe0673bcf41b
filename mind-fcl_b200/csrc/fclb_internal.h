// fclb_internal.h -- glue between the C ABI (fclb_engine.cu) and the kernel
// translation units.  Not part of the public interface.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/fclb200.h"

namespace fclb {

constexpr int kNumTypes = 8;                      // ShapeType codes 0..7
constexpr int kNumKinds = kNumTypes * kNumTypes;  // (type1,type2) buckets
constexpr int kBlock = 128;                       // threads per CTA for per-query kernels

// One bucket of a batch: queries perm[begin .. begin+count) all have the same
// (type1,type2).  perm == nullptr means the identity permutation.
struct BatchView {
  const void* shapes;   // ShapeD<S>[]
  const void* convex;   // ConvexD<S>[]
  const void* tris;     // device BVH triangle array (12 S per triangle) behind ST_TRIANGLE table entries, or nullptr
  const fclb_pair* pairs;
  const void* poses1;
  const void* poses2;
  const uint32_t* perm;
  size_t begin, count;
  int type1, type2;
};

struct DistanceOut {
  void* dist;
  void* p1;
  void* p2;
  uint8_t* ok;
};

struct SolverParams {
  double gjk_tol;
  int gjk_max_iter;
  double epa_tol;
  int epa_max_faces;
  int epa_max_iter;
  double eps78;  // constants<S>::eps_78()
  int generic_only = 0;  // 1: skip the closed-form distance routines (shapeSignedDistance has no specialisations)
};

// implemented in fclb_distance_f32.cu / fclb_distance_f64.cu
template <typename S>
cudaError_t launchDistance(const BatchView& b, const SolverParams& sp, const DistanceOut& out, cudaStream_t st,
                           int* n_launches);

// Leaf-candidate sink of the scene traversals (DefaultGJK_EPA requests): every leaf / leaf pair that survives the node
// culls is appended here WITHOUT a leaf test; the leaf batch (fclb_scene_gjk.cu) then runs the reference's leaf routine
// (ShapeIntersect / ShapeSimplexIntersect with contacts) on each through the shape-pair collide pipeline.
struct LeafCandSink {
  unsigned long long* count = nullptr;  // device counter (may exceed cap: the call is then repeated with a larger buffer)
  unsigned long long cap = 0;
  uint32_t* q = nullptr;                // query of the candidate
  long long* b1 = nullptr;              // leaf id on side 1 (triangle id, encodePixel, encodeOctree2Node)
  long long* b2 = nullptr;              // leaf id on side 2 (scene pairs) or unused
  void* box1 = nullptr;                 // [cap * 6 S] leaf box of side 1 in its scene frame (heightmap / octree), or nullptr
  void* box2 = nullptr;                 // [cap * 6 S] leaf box of side 2 (scene pairs with a box hierarchy on side 2)
};

constexpr int kBvhShapeWarps = 8;  // warps per CTA of the mesh-shape kernel
// mesh-shape traversal (fclb_bvh_shape_impl.cuh, instantiated in fclb_bvh_shape_f32/f64.cu)
struct BvhShapeArgs {
  const void* nodes;
  const void* tris;
  const void* shapes;       // ShapeD<S>[]
  const void* convex;       // ConvexD<S>[]
  const void* bound;        // BoundD<S>[]
  const uint32_t* shape_ids;  // per query index into the shape table
  const void* poses_mesh;
  const void* poses_shape;
  size_t n;
  uint32_t max_contacts;
  double tol;
  int max_iter;
  uint32_t* counts;
  int32_t* first_tri;
  // optional contact sink (pass 1 of the MPR penetration modes): ids / boxes of the first max_keep hit leaves
  uint32_t max_keep;
  long long* out_b1;   // [n * max_keep]
  void* out_box;       // [n * max_keep * 6 S] leaf box in the scene frame (heightmap / octree), or nullptr
  unsigned long long* work_counter;
  unsigned long long* stats;  // [0] node tests, [1] leaf tests
  LeafCandSink cand;          // candidate mode (cand.count != nullptr): leaves are appended, not tested
};

template <typename S>
cudaError_t launchBvhShape(int type0, const BvhShapeArgs& a, int grid, cudaStream_t st);

// heightmap-shape scan (fclb_heightmap_impl.cuh, instantiated in fclb_heightmap_f32/f64.cu)
constexpr int kHeightmapWarps = 8;
struct HeightmapArgs {
  const uint16_t* bottom;   // bottom layer, index = y * full_x + x (flat_heightmap-inl.h:121-124)
  const uint16_t* coarse;   // layer `coarse_shift` levels above the bottom (max over 2^shift x 2^shift pixels)
  int coarse_shift;
  uint32_t coarse_full_x;
  uint32_t full_x, full_y, half_x, half_y;
  uint32_t upper_mm;        // height_upper_bound_in_mm
  double res_x, res_y;      // bottom resolution (already rounded to S)
  const void* shapes;
  const void* convex;
  const uint32_t* shape_ids;
  const void* poses_hm;
  const void* poses_shape;
  size_t n;
  uint32_t max_contacts;
  double tol;
  int max_iter;
  uint32_t* counts;
  int32_t* first_pixel;     // encodePixel = x << 16 | y (heightmap_types.h:53-58) or -1
  // optional contact sink (pass 1 of the MPR penetration modes): ids / boxes of the first max_keep hit leaves
  uint32_t max_keep;
  long long* out_b1;   // [n * max_keep]
  void* out_box;       // [n * max_keep * 6 S] leaf box in the scene frame (heightmap / octree), or nullptr
  unsigned long long* work_counter;
  unsigned long long* stats;  // [0] pixels read, [1] pixel boxes tested
  LeafCandSink cand;          // candidate mode (cand.count != nullptr): pixel boxes are appended, not tested
};
template <typename S>
cudaError_t launchHeightmapShape(int type1, const HeightmapArgs& a, int grid, cudaStream_t st);

// octree-shape traversal (fclb_octree_impl.cuh, instantiated in fclb_octree_f32/f64.cu)
constexpr int kOctreeWarps = 4;
struct OctreeArgs {
  const uint32_t* inner_children;  // 8 per inner node, 0xffffffff = no child (octree_node.h:35-39)
  const uint8_t* inner_full;       // inner_nodes_fully_occupied
  const uint8_t* leaf_bits;        // OctreeLeafNode::child_occupied, 1 byte per leaf-layer node
  const uint8_t* pruned;           // prune_internal_nodes or nullptr
  uint32_t n_inner, n_leaf;
  int num_layers;                  // Octree::n_layers()
  double root_box[6];              // root_bv: min xyz, max xyz
  const void* shapes;
  const void* convex;
  const uint32_t* shape_ids;
  const void* poses_octree;
  const void* poses_shape;
  size_t n;
  uint32_t max_contacts;
  double tol;
  int max_iter;
  uint32_t* counts;
  long long* first_node;           // encodeOctree2Node of one hit box or -1
  // optional contact sink (pass 1 of the MPR penetration modes): ids / boxes of the first max_keep hit leaves
  uint32_t max_keep;
  long long* out_b1;   // [n * max_keep]
  void* out_box;       // [n * max_keep * 6 S] leaf box in the scene frame (heightmap / octree), or nullptr
  unsigned long long* work_counter;
  unsigned long long* stats;       // [0] node boxes tested, [1] voxel boxes tested
  LeafCandSink cand;               // candidate mode (cand.count != nullptr): voxel boxes are appended, not tested
};
template <typename S>
cudaError_t launchOctreeShape(int type1, const OctreeArgs& a, int grid, cudaStream_t st);

// pass 2 of the MPR penetration modes for scene contacts (fclb_scene_pen_impl.cuh)
struct ScenePenArgs {
  int leaf_is_triangle;      // 1: b1 = triangle id of `tris`; 0: box from out_box
  const void* tris;          // 12 S per triangle (device BVH layout)
  const void* shapes;
  const void* convex;
  const uint32_t* shape_ids;
  const void* poses_scene;
  const void* poses_shape;
  size_t n;
  uint32_t max_keep;
  const uint32_t* counts;
  const long long* b1;
  const void* box;
  int incremental;
  double dir[3];
  double tol;
  void* out_contacts;        // [n * max_keep * 7 S] = normal, pos, depth
};
template <typename S>
cudaError_t launchScenePenetration(const ScenePenArgs& a, cudaStream_t st);

// scene-vs-scene pair traversal (fclb_scene_pair_impl.cuh, instantiated in fclb_scene_pair_f32/f64.cu)
constexpr int kScenePairWarps = 2;
struct HmView {           // LayeredHeightMap<S> on the device; layer 0 = bottom, layer k = k levels above it
  const uint16_t* layers;
  uint32_t off[16];       // element offset of layer k
  uint16_t fx[16], fy[16];
  int n_layers;
  uint32_t half_x, half_y;  // bottom half shape
  double res_x, res_y;      // bottom resolution (already rounded to S)
};
struct OctView {          // octree2::Octree<S> flat arrays (see OctreeArgs)
  const uint32_t* children;
  const uint8_t* inner_full;
  const uint8_t* leaf_bits;
  const uint8_t* pruned;
  uint32_t n_inner, n_leaf;
  int num_layers;
  double root_box[6];
};
int sceneHmView(fclb_handle h, int st, HmView& v);  // fclb_scene_api.cu
int sceneOctView(fclb_handle h, OctView& v);
struct BvhView {
  const void* nodes;
  const void* tris;
  int n_nodes;
};
struct ScenePairArgs {
  int kind1, kind2;       // FCLB_SCENE_*; kind1 is a heightmap or an octree
  HmView hm1, hm2;
  OctView oct1, oct2;
  BvhView bvh2;
  const void* poses1;
  const void* poses2;
  size_t n;
  uint32_t max_contacts;
  uint32_t* counts;
  uint32_t max_keep;
  long long* out_b1;      // [n * max_keep] or nullptr
  long long* out_b2;
  void* out_box1;         // [n * max_keep * 6 S] leaf boxes of side 1 / side 2 (Contact::o1_bv / o2_bv), or nullptr
  void* out_box2;
  unsigned long long* work_counter;
  unsigned long long* stats;  // [0] node pairs tested, [1] leaf pairs tested, [2] stack overflows
  LeafCandSink cand;          // candidate mode (cand.count != nullptr): leaf pairs are appended, not tested
};
template <typename S>
cudaError_t launchScenePair(const ScenePairArgs& a, int grid, cudaStream_t st);

// pass 2 of the MPR penetration modes for scene-pair contacts (fclb_scene_pen_impl.cuh)
struct ScenePairPenArgs {
  int leaf2_is_triangle;     // 1: side 2 is a mesh, b2 = triangle id of `tris`; 0: box from box2
  const void* tris;
  const void* poses1;
  const void* poses2;
  size_t n;
  uint32_t max_keep;
  const uint32_t* counts;
  const long long* b2;
  const void* box1;
  const void* box2;
  int incremental;
  double dir[3];
  double tol;
  void* out_contacts;        // [n * max_keep * 7 S] = normal, pos, depth
};
template <typename S>
cudaError_t launchScenePairPenetration(const ScenePairPenArgs& a, cudaStream_t st);

}  // namespace fclb
