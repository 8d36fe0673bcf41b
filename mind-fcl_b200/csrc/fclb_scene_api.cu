// fclb_scene_api.cu -- C ABI entry points for shape-vs-scene queries:
//   fclb_bvh_shape_collide_batch_{dev,host}   mesh (BVHModel<OBBRSS>) vs convex shape
// Kernels: fclb_bvh_shape_impl.cuh (instantiated in fclb_bvh_shape_f32/f64.cu).
#include <cmath>
#include <cstring>

#include <cub/cub.cuh>

#include "fclb_bvh.cuh"
#include "fclb_octree_build.h"
#include "fclb_octree_build_dev.cuh"
#include "fclb_scene_gjk_impl.cuh"
#include "fclb_shapes.cuh"

namespace fclb {

struct SceneStats {
  unsigned long long* counters = nullptr;  // device: [0] work counter, [1..2] stats
  unsigned long long stats[3] = {0, 0, 0};
  void* d_box = nullptr;       // pass-1 leaf boxes of the MPR penetration path (scene-shape)
  size_t d_box_cap = 0;
  void* d_box_pair = nullptr;  // ... (scene pairs)
  size_t d_box_pair_cap = 0;
};
static PerDevice<SceneStats> g_scene_pd;
#define g_counters (g_scene_pd.get().counters)
#define g_stats (g_scene_pd.get().stats)

struct ContactSink {  // pass 1 of the MPR penetration modes
  uint32_t max_keep = 0;
  long long* b1 = nullptr;
  void* box = nullptr;
  LeafCandSink cand;  // candidate mode of the DefaultGJK_EPA path (fclb_scene_gjk_impl.cuh)
};

static int tableUniformType(const ShapeTable* t) {
  int type = int(t->host[0].type);
  for (uint32_t i = 1; i < t->n; i++)
    if (int(t->host[i].type) != type) return ST_DYNAMIC;
  return type;
}

template <typename S>
static int bvhShapeDev(Engine& e, const BvhDev* m, const ShapeTable* t, const uint32_t* shape_ids, const void* poses_mesh,
                       const void* poses_shape, size_t n, const fclb_request* req, uint32_t* counts, int32_t* first_tri,
                       const ContactSink& sink = ContactSink()) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  if (!g_counters) FCLB_CUDA(cudaMalloc(&g_counters, 4 * sizeof(unsigned long long)));
  FCLB_CUDA(cudaMemsetAsync(g_counters, 0, 4 * sizeof(unsigned long long), e.compute));
  const SolverParams sp = solverParams(st, req->binary_tol, req->gjk_max_iter, req->distance_tol, req->epa_max_faces,
                                       req->epa_max_iter, true);
  BvhShapeArgs a{};
  a.nodes = m->nodes;
  a.tris = m->tris;
  a.shapes = t->d_shapes[st];
  a.convex = e.d_convex_tab[st];
  a.bound = t->d_bound[st];
  a.shape_ids = shape_ids;
  a.poses_mesh = poses_mesh;
  a.poses_shape = poses_shape;
  a.n = n;
  a.max_contacts = req->max_contacts;
  a.tol = sp.gjk_tol;
  a.max_iter = sp.gjk_max_iter;
  a.counts = counts;
  a.first_tri = first_tri;
  a.max_keep = sink.max_keep;
  a.out_b1 = sink.b1;
  a.out_box = sink.box;
  a.cand = sink.cand;
  a.work_counter = g_counters;
  a.stats = g_counters + 1;
  const size_t need = (n + kBvhShapeWarps - 1) / kBvhShapeWarps;
  const size_t cap = size_t(e.sms) * 4;
  const int grid = int(need < cap ? need : cap);
  FCLB_CUDA(cudaEventRecord(e.ev_call0, e.compute));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  FCLB_CUDA(launchBvhShape<S>(tableUniformType(t), a, grid, e.compute));
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  e.launches += 1;
  FCLB_CUDA(cudaMemcpyAsync(g_stats, g_counters + 1, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  if (g_stats[2]) return fail(FCLB_ERR_CAPACITY, "scene traversal: tree deeper than the per-warp stack allows");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = ms;
  e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -2;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

// ---- LayeredHeightMap<S> on the device ------------------------------------------
struct HeightmapDev {
  uint16_t* d_layers = nullptr;        // all layers, bottom first
  std::vector<size_t> off;             // element offset of layer k (k = 0: bottom, k levels above it)
  std::vector<uint32_t> fx, fy;        // full shape per layer
  uint32_t half_x = 0, half_y = 0;     // bottom half shape
  uint32_t upper_mm = 0;
  double res_x = 0, res_y = 0;
};
static std::map<fclb_handle, HeightmapDev*>& hmTable() {
  static std::map<fclb_handle, HeightmapDev*> t[kMaxDevices];
  return t[currentSlot()];
}

template <typename S>
static int heightmapShapeDev(Engine& e, const HeightmapDev* hm, const ShapeTable* t, const uint32_t* shape_ids,
                             const void* poses_hm, const void* poses_shape, size_t n, const fclb_request* req,
                             uint32_t* counts, int32_t* first_pixel, const ContactSink& sink = ContactSink()) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  if (!g_counters) FCLB_CUDA(cudaMalloc(&g_counters, 4 * sizeof(unsigned long long)));
  FCLB_CUDA(cudaMemsetAsync(g_counters, 0, 4 * sizeof(unsigned long long), e.compute));
  const SolverParams sp = solverParams(st, req->binary_tol, req->gjk_max_iter, req->distance_tol, req->epa_max_faces,
                                       req->epa_max_iter, true);
  HeightmapArgs a{};
  const int shift = int(hm->off.size()) - 1 < 3 ? int(hm->off.size()) - 1 : 3;
  a.bottom = hm->d_layers;
  a.coarse = hm->d_layers + hm->off[shift];
  a.coarse_shift = shift;
  a.coarse_full_x = hm->fx[shift];
  a.full_x = hm->fx[0];
  a.full_y = hm->fy[0];
  a.half_x = hm->half_x;
  a.half_y = hm->half_y;
  a.upper_mm = hm->upper_mm;
  a.res_x = double(S(hm->res_x));
  a.res_y = double(S(hm->res_y));
  a.shapes = t->d_shapes[st];
  a.convex = e.d_convex_tab[st];
  a.shape_ids = shape_ids;
  a.poses_hm = poses_hm;
  a.poses_shape = poses_shape;
  a.n = n;
  a.max_contacts = req->max_contacts;
  a.tol = sp.gjk_tol;
  a.max_iter = sp.gjk_max_iter;
  a.counts = counts;
  a.first_pixel = first_pixel;
  a.max_keep = sink.max_keep;
  a.out_b1 = sink.b1;
  a.out_box = sink.box;
  a.cand = sink.cand;
  a.work_counter = g_counters;
  a.stats = g_counters + 1;
  const size_t need = (n + kHeightmapWarps - 1) / kHeightmapWarps;
  const size_t cap = size_t(e.sms) * 4;
  const int grid = int(need < cap ? need : cap);
  FCLB_CUDA(cudaEventRecord(e.ev_call0, e.compute));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  FCLB_CUDA(launchHeightmapShape<S>(tableUniformType(t), a, grid, e.compute));
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  e.launches += 1;
  FCLB_CUDA(cudaMemcpyAsync(g_stats, g_counters + 1, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  if (g_stats[2]) return fail(FCLB_ERR_CAPACITY, "scene traversal: tree deeper than the per-warp stack allows");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = ms;
  e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -3;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

// ---- Octree2 on the device ------------------------------------------------------------
struct OctreeDev {
  uint32_t* children = nullptr;
  uint8_t *inner_full = nullptr, *leaf_bits = nullptr, *pruned = nullptr;
  uint32_t n_inner = 0, n_leaf = 0;
  int num_layers = 0;
  double root_box[6] = {0, 0, 0, 0, 0, 0};
};
static std::map<fclb_handle, OctreeDev*>& octTable() {
  static std::map<fclb_handle, OctreeDev*> t[kMaxDevices];
  return t[currentSlot()];
}

template <typename S>
static int octreeShapeDev(Engine& e, const OctreeDev* o, const ShapeTable* t, const uint32_t* shape_ids,
                          const void* poses_octree, const void* poses_shape, size_t n, const fclb_request* req,
                          uint32_t* counts, long long* first_node, const ContactSink& sink = ContactSink()) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  if (!g_counters) FCLB_CUDA(cudaMalloc(&g_counters, 4 * sizeof(unsigned long long)));
  FCLB_CUDA(cudaMemsetAsync(g_counters, 0, 4 * sizeof(unsigned long long), e.compute));
  const SolverParams sp = solverParams(st, req->binary_tol, req->gjk_max_iter, req->distance_tol, req->epa_max_faces,
                                       req->epa_max_iter, true);
  OctreeArgs a{};
  a.inner_children = o->children;
  a.inner_full = o->inner_full;
  a.leaf_bits = o->leaf_bits;
  a.pruned = o->pruned;
  a.n_inner = o->n_inner;
  a.n_leaf = o->n_leaf;
  a.num_layers = o->num_layers;
  for (int k = 0; k < 6; k++) a.root_box[k] = o->root_box[k];
  a.shapes = t->d_shapes[st];
  a.convex = e.d_convex_tab[st];
  a.shape_ids = shape_ids;
  a.poses_octree = poses_octree;
  a.poses_shape = poses_shape;
  a.n = n;
  a.max_contacts = req->max_contacts;
  a.tol = sp.gjk_tol;
  a.max_iter = sp.gjk_max_iter;
  a.counts = counts;
  a.first_node = first_node;
  a.max_keep = sink.max_keep;
  a.out_b1 = sink.b1;
  a.out_box = sink.box;
  a.cand = sink.cand;
  a.work_counter = g_counters;
  a.stats = g_counters + 1;
  const size_t need = (n + kOctreeWarps - 1) / kOctreeWarps;
  const size_t cap = size_t(e.sms) * 4;
  const int grid = int(need < cap ? need : cap);
  FCLB_CUDA(cudaEventRecord(e.ev_call0, e.compute));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  FCLB_CUDA(launchOctreeShape<S>(tableUniformType(t), a, grid, e.compute));
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  e.launches += 1;
  FCLB_CUDA(cudaMemcpyAsync(g_stats, g_counters + 1, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  if (g_stats[2]) return fail(FCLB_ERR_CAPACITY, "scene traversal: tree deeper than the per-warp stack allows");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = ms;
  e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -4;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

// fcl::collide with a DirectedPenetration / IncrementalMinimumPenetration request against a scene geometry:
// boolean traversal with a contact sink (pass 1), then one MPR penetration per stored contact (pass 2).
template <typename S>
static int sceneContactsDev(Engine& e, int kind, fclb_handle scene, const ShapeTable* t, const uint32_t* shape_ids,
                            const void* poses_scene, const void* poses_shape, size_t n, const fclb_request* req,
                            uint32_t max_keep, uint32_t* counts, long long* b1, void* contacts) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  void*& d_box = g_scene_pd.get().d_box;
  size_t& d_box_cap = g_scene_pd.get().d_box_cap;
  const size_t box_bytes = n * size_t(max_keep) * 6 * sizeof(S);
  if (kind != FCLB_SCENE_BVH && d_box_cap < box_bytes) {
    cudaFree(d_box);
    d_box = nullptr;
    d_box_cap = 0;
    FCLB_CUDA(cudaMalloc(&d_box, box_bytes));
    d_box_cap = box_bytes;
  }
  ContactSink sink;
  sink.max_keep = max_keep;
  sink.b1 = b1;
  sink.box = kind == FCLB_SCENE_BVH ? nullptr : d_box;
  fclb_request boolean_req = *req;
  boolean_req.penetration_mode = FCLB_PEN_DISABLED;
  const void* tris = nullptr;
  int rc = FCLB_OK;
  if (kind == FCLB_SCENE_BVH) {
    auto it = bvhTable().find(scene);
    if (it == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown BVH handle");
    if (it->second->scalar_type != st) return fail(FCLB_ERR_BAD_ARG, "BVH was uploaded for a different scalar type");
    tris = it->second->tris;
    rc = bvhShapeDev<S>(e, it->second, t, shape_ids, poses_scene, poses_shape, n, &boolean_req, counts, nullptr, sink);
  } else if (kind == FCLB_SCENE_HEIGHTMAP) {
    auto it = hmTable().find(scene);
    if (it == hmTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown heightmap handle");
    rc = heightmapShapeDev<S>(e, it->second, t, shape_ids, poses_scene, poses_shape, n, &boolean_req, counts, nullptr, sink);
  } else if (kind == FCLB_SCENE_OCTREE) {
    auto it = octTable().find(scene);
    if (it == octTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown octree handle");
    rc = octreeShapeDev<S>(e, it->second, t, shape_ids, poses_scene, poses_shape, n, &boolean_req, counts, nullptr, sink);
  } else {
    return fail(FCLB_ERR_BAD_ARG, "unknown scene kind");
  }
  if (rc) return rc;
  if (req->penetration_mode == FCLB_PEN_DISABLED) {  // boolean request: ids only, no contact geometry
    FCLB_CUDA(cudaMemsetAsync(contacts, 0, n * size_t(max_keep) * 7 * sizeof(S), e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    return FCLB_OK;
  }
  const SolverParams sp = solverParams(st, req->binary_tol, req->gjk_max_iter, req->distance_tol, req->epa_max_faces,
                                       req->epa_max_iter, true);
  ScenePenArgs p{};
  p.leaf_is_triangle = kind == FCLB_SCENE_BVH ? 1 : 0;
  p.tris = tris;
  p.shapes = t->d_shapes[st];
  p.convex = e.d_convex_tab[st];
  p.shape_ids = shape_ids;
  p.poses_scene = poses_scene;
  p.poses_shape = poses_shape;
  p.n = n;
  p.max_keep = max_keep;
  p.counts = counts;
  p.b1 = b1;
  p.box = sink.box;
  p.incremental = req->penetration_mode == FCLB_PEN_INCREMENTAL_MIN ? 1 : 0;
  for (int k = 0; k < 3; k++) p.dir[k] = req->dir[k];
  p.tol = sp.epa_tol;  // MPR(128, request.distanceTolerance())
  p.out_contacts = contacts;
  FCLB_CUDA(launchScenePenetration<S>(p, e.compute));
  e.launches += 1;
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}

// fcl::collide between two scene geometries (heightmap / octree first, heightmap / octree / mesh second), boolean request
static int fillHmView(const HeightmapDev* hm, int st, HmView& v) {
  if (hm->off.size() > 16) return fail(FCLB_ERR_UNSUPPORTED, "scene pair: heightmap with more than 16 layers");
  v.layers = hm->d_layers;
  v.n_layers = int(hm->off.size());
  for (size_t k = 0; k < hm->off.size(); k++) {
    v.off[k] = uint32_t(hm->off[k]);
    v.fx[k] = uint16_t(hm->fx[k]);
    v.fy[k] = uint16_t(hm->fy[k]);
  }
  v.half_x = hm->half_x;
  v.half_y = hm->half_y;
  v.res_x = st == 0 ? double(float(hm->res_x)) : hm->res_x;
  v.res_y = st == 0 ? double(float(hm->res_y)) : hm->res_y;
  return FCLB_OK;
}
static void fillOctView(const OctreeDev* o, OctView& v) {
  v.children = o->children;
  v.inner_full = o->inner_full;
  v.leaf_bits = o->leaf_bits;
  v.pruned = o->pruned;
  v.n_inner = o->n_inner;
  v.n_leaf = o->n_leaf;
  v.num_layers = o->num_layers;
  for (int k = 0; k < 6; k++) v.root_box[k] = o->root_box[k];
}

// views of an uploaded heightmap / octree for the other translation units (fclb_ccd_scene.cu)
int sceneHmView(fclb_handle h, int st, HmView& v) {
  auto it = hmTable().find(h);
  if (it == hmTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown heightmap handle");
  return fillHmView(it->second, st, v);
}
int sceneOctView(fclb_handle h, OctView& v) {
  auto it = octTable().find(h);
  if (it == octTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown octree handle");
  fillOctView(it->second, v);
  return FCLB_OK;
}

template <typename S>
static int scenePairDev(Engine& e, int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                        const void* poses2, size_t n, const fclb_request* req, uint32_t max_keep, uint32_t* counts,
                        long long* b1, long long* b2, void* box1 = nullptr, void* box2 = nullptr,
                        const void** tris_out = nullptr, const LeafCandSink* cand = nullptr) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  ScenePairArgs a{};
  a.kind1 = kind1;
  a.kind2 = kind2;
  size_t roots1 = 1, roots2 = 1;
  for (int side = 0; side < 2; side++) {
    const int kind = side == 0 ? kind1 : kind2;
    const fclb_handle h = side == 0 ? scene1 : scene2;
    if (kind == FCLB_SCENE_HEIGHTMAP) {
      auto it = hmTable().find(h);
      if (it == hmTable().end()) return fail(FCLB_ERR_BAD_ARG, "scene pair: unknown heightmap handle");
      const int rc = fillHmView(it->second, st, side == 0 ? a.hm1 : a.hm2);
      if (rc) return rc;
      (side == 0 ? roots1 : roots2) = size_t(it->second->fx.back()) * it->second->fy.back();
    } else if (kind == FCLB_SCENE_OCTREE) {
      auto it = octTable().find(h);
      if (it == octTable().end()) return fail(FCLB_ERR_BAD_ARG, "scene pair: unknown octree handle");
      fillOctView(it->second, side == 0 ? a.oct1 : a.oct2);
    } else if (kind == FCLB_SCENE_BVH && side == 1) {
      auto it = bvhTable().find(h);
      if (it == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "scene pair: unknown BVH handle");
      if (it->second->scalar_type != st) return fail(FCLB_ERR_BAD_ARG, "BVH was uploaded for a different scalar type");
      a.bvh2.nodes = it->second->nodes;
      a.bvh2.tris = it->second->tris;
      a.bvh2.n_nodes = it->second->n_nodes;
      if (tris_out) *tris_out = it->second->tris;
    } else {
      return fail(FCLB_ERR_UNSUPPORTED, "scene pair: supported pairs are heightmap-{heightmap, mesh, octree} and "
                                        "octree-{mesh, octree}, in this argument order");
    }
  }
  if (kind1 == FCLB_SCENE_OCTREE && kind2 == FCLB_SCENE_HEIGHTMAP)
    return fail(FCLB_ERR_UNSUPPORTED, "scene pair: pass the heightmap first (heightMapOctreeIntersect)");
  if (roots1 * roots2 > 256)
    return fail(FCLB_ERR_CAPACITY, "scene pair: more than 256 top-layer pixel pairs (strongly non-square heightmaps)");
  if (!g_counters) FCLB_CUDA(cudaMalloc(&g_counters, 4 * sizeof(unsigned long long)));
  FCLB_CUDA(cudaMemsetAsync(g_counters, 0, 4 * sizeof(unsigned long long), e.compute));
  a.poses1 = poses1;
  a.poses2 = poses2;
  a.n = n;
  a.max_contacts = req->max_contacts;
  a.counts = counts;
  a.max_keep = b1 ? max_keep : 0;
  a.out_b1 = b1;
  a.out_b2 = b2;
  a.out_box1 = box1;
  a.out_box2 = box2;
  if (cand) a.cand = *cand;
  a.work_counter = g_counters;
  a.stats = g_counters + 1;
  const size_t need = (n + kScenePairWarps - 1) / kScenePairWarps;
  const size_t cap = size_t(e.sms) * 4;
  const int grid = int(need < cap ? need : cap);
  FCLB_CUDA(cudaEventRecord(e.ev_call0, e.compute));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  FCLB_CUDA(launchScenePair<S>(a, grid, e.compute));
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  e.launches += 1;
  FCLB_CUDA(cudaMemcpyAsync(g_stats, g_counters + 1, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  if (g_stats[2]) return fail(FCLB_ERR_CAPACITY, "scene pair traversal: hierarchy deeper than the per-warp stack allows");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = ms;
  e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -5;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

// fcl::collide between two scene geometries with a DirectedPenetration / IncrementalMinimumPenetration request:
// boolean pair traversal with a sink of ids + leaf boxes (pass 1), one MPR penetration per stored contact (pass 2)
template <typename S>
static int scenePairContactsDev(Engine& e, int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                const void* poses2, size_t n, const fclb_request* req, uint32_t max_keep, uint32_t* counts,
                                long long* b1, long long* b2, void* contacts) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  void*& d_box = g_scene_pd.get().d_box_pair;
  size_t& d_box_cap = g_scene_pd.get().d_box_pair_cap;
  const size_t one = n * size_t(max_keep) * 6 * sizeof(S);
  if (d_box_cap < 2 * one) {
    cudaFree(d_box);
    d_box = nullptr;
    d_box_cap = 0;
    FCLB_CUDA(cudaMalloc(&d_box, 2 * one));
    d_box_cap = 2 * one;
  }
  void* box1 = d_box;
  void* box2 = static_cast<char*>(d_box) + one;
  fclb_request boolean_req = *req;
  boolean_req.penetration_mode = FCLB_PEN_DISABLED;
  const void* tris = nullptr;
  int rc = scenePairDev<S>(e, kind1, scene1, kind2, scene2, poses1, poses2, n, &boolean_req, max_keep, counts, b1, b2, box1,
                           box2, &tris);
  if (rc) return rc;
  const SolverParams sp = solverParams(st, req->binary_tol, req->gjk_max_iter, req->distance_tol, req->epa_max_faces,
                                       req->epa_max_iter, true);
  ScenePairPenArgs p{};
  p.leaf2_is_triangle = kind2 == FCLB_SCENE_BVH ? 1 : 0;
  p.tris = tris;
  p.poses1 = poses1;
  p.poses2 = poses2;
  p.n = n;
  p.max_keep = max_keep;
  p.counts = counts;
  p.b2 = b2;
  p.box1 = box1;
  p.box2 = box2;
  p.incremental = req->penetration_mode == FCLB_PEN_INCREMENTAL_MIN ? 1 : 0;
  for (int k = 0; k < 3; k++) p.dir[k] = req->dir[k];
  p.tol = sp.epa_tol;  // MPR(128, request.distanceTolerance())
  p.out_contacts = contacts;
  FCLB_CUDA(launchScenePairPenetration<S>(p, e.compute));
  e.launches += 1;
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}

// FlatHeightMap<S>::updateHeightsByPointGenerationFunctor (flat_heightmap-inl.h:249-272)
// ---- DefaultGJK_EPA requests on scene geometries: candidate traversal + leaf batch (fclb_scene_gjk_impl.cuh) -----
namespace {
struct DevBuf {  // scope-bound device allocation; stream-ordered (pooled) when a stream is given
  void* p = nullptr;
  cudaStream_t stream = nullptr;
  bool pooled = false;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { reset(); }
  void reset() {
    if (p) {
      if (pooled)
        cudaFreeAsync(p, stream);
      else
        cudaFree(p);
    }
    p = nullptr;
  }
  cudaError_t alloc(size_t bytes) {
    reset();
    pooled = false;
    return cudaMalloc(&p, bytes ? bytes : 16);
  }
  cudaError_t alloc(size_t bytes, cudaStream_t st) {
    reset();
    pooled = true;
    stream = st;
    return cudaMallocAsync(&p, bytes ? bytes : 16, st);
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
};
}  // namespace

// kind1 / scene1: the scene geometry (scene-shape) or side 1 (scene pair); kind2 < 0: scene-shape with table / shape_ids
template <typename S>
static int sceneGjkContactsDev(Engine& e, int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const ShapeTable* t,
                               const uint32_t* shape_ids, const void* poses_a, const void* poses_b, size_t n,
                               const fclb_request* req, uint32_t max_keep, uint32_t* counts, long long* out_b1,
                               long long* out_b2, void* contacts) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  const bool pair = kind2 >= 0;
  if (max_keep == 0 || !out_b1 || !contacts) return fail(FCLB_ERR_BAD_ARG, "scene contacts: max_keep, id and contact arrays are required");
  if (pair && !out_b2) return fail(FCLB_ERR_BAD_ARG, "scene pair contacts: out_b2 is required");
  const void* tris = nullptr;
  int mode = 0;
  if (!pair) {
    mode = kind1 == FCLB_SCENE_BVH ? 0 : 1;
  } else {
    mode = kind2 == FCLB_SCENE_BVH ? 3 : 2;
  }
  fclb_request all = *req;  // pass 1: every leaf that survives the node culls
  all.penetration_mode = FCLB_PEN_DISABLED;
  all.max_contacts = 0x7fffffffu;
  const size_t ps = 12 * sizeof(S);
  size_t cap = std::max<size_t>(size_t(1) << 16, std::min<size_t>(n * 32, size_t(1) << 24));
  const size_t cap_max = size_t(1) << 25;  // candidates per chunk (leaf batch of ~300 B per item)
  DevBuf d_count, d_cq, d_cb1, d_cb2, d_box1, d_box2, d_scratch;
  FCLB_CUDA(d_count.alloc(sizeof(unsigned long long), e.compute));
  FCLB_CUDA(d_scratch.alloc(n * sizeof(uint32_t), e.compute));
  auto allocCand = [&](size_t c) -> int {
    FCLB_CUDA(d_cq.alloc(c * 4, e.compute));
    FCLB_CUDA(d_cb1.alloc(c * 8, e.compute));
    FCLB_CUDA(d_cb2.alloc(c * 8, e.compute));
    FCLB_CUDA(d_box1.alloc(c * 6 * sizeof(S), e.compute));
    FCLB_CUDA(d_box2.alloc(c * 6 * sizeof(S), e.compute));
    return FCLB_OK;
  };
  int rc = allocCand(cap);
  if (rc) return rc;
  // query chunks: a chunk whose candidates exceed the per-chunk budget is halved and run again
  std::vector<std::pair<size_t, size_t>> todo{{0, n}};
  while (!todo.empty()) {
    const size_t c0 = todo.back().first, c1 = todo.back().second;
    todo.pop_back();
    const size_t nc = c1 - c0;
    unsigned long long m64 = 0;
    while (true) {
      FCLB_CUDA(cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), e.compute));
      LeafCandSink cs;
      cs.count = d_count.as<unsigned long long>();
      cs.cap = cap;
      cs.q = d_cq.as<uint32_t>();
      cs.b1 = d_cb1.as<long long>();
      cs.b2 = pair ? d_cb2.as<long long>() : nullptr;
      cs.box1 = mode == 0 ? nullptr : d_box1.p;
      cs.box2 = mode == 2 ? d_box2.p : nullptr;
      const char* pa = static_cast<const char*>(poses_a) + c0 * ps;
      const char* pb = static_cast<const char*>(poses_b) + c0 * ps;
      if (!pair) {
        ContactSink sink;
        sink.cand = cs;
        if (kind1 == FCLB_SCENE_BVH) {
          auto it = bvhTable().find(scene1);
          if (it == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown BVH handle");
          if (it->second->scalar_type != st) return fail(FCLB_ERR_BAD_ARG, "BVH was uploaded for a different scalar type");
          tris = it->second->tris;
          rc = bvhShapeDev<S>(e, it->second, t, shape_ids + c0, pa, pb, nc, &all, d_scratch.as<uint32_t>(), nullptr, sink);
        } else if (kind1 == FCLB_SCENE_HEIGHTMAP) {
          auto it = hmTable().find(scene1);
          if (it == hmTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown heightmap handle");
          rc = heightmapShapeDev<S>(e, it->second, t, shape_ids + c0, pa, pb, nc, &all, d_scratch.as<uint32_t>(), nullptr, sink);
        } else if (kind1 == FCLB_SCENE_OCTREE) {
          auto it = octTable().find(scene1);
          if (it == octTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown octree handle");
          rc = octreeShapeDev<S>(e, it->second, t, shape_ids + c0, pa, pb, nc, &all, d_scratch.as<uint32_t>(), nullptr, sink);
        } else {
          return fail(FCLB_ERR_BAD_ARG, "unknown scene kind");
        }
      } else {
        rc = scenePairDev<S>(e, kind1, scene1, kind2, scene2, pa, pb, nc, &all, 0, d_scratch.as<uint32_t>(), nullptr, nullptr,
                             nullptr, nullptr, &tris, &cs);
      }
      if (rc) return rc;
      FCLB_CUDA(cudaMemcpyAsync(&m64, d_count.p, sizeof(m64), cudaMemcpyDeviceToHost, e.compute));
      FCLB_CUDA(cudaStreamSynchronize(e.compute));
      if (m64 <= cap) break;
      if (m64 > cap_max && nc > 1) break;  // split the chunk instead of growing
      cap = size_t(m64) + size_t(m64) / 8;
      rc = allocCand(cap);
      if (rc) return rc;
    }
    if (m64 > cap) {  // over budget: halve
      const size_t mid = c0 + nc / 2;
      todo.push_back({mid, c1});
      todo.push_back({c0, mid});
      continue;
    }
    const size_t m = size_t(m64);
    DevBuf d_qcount, d_qoff, d_itemq, d_ib1, d_ib2, d_flags, d_lc, d_lcnt;
    FCLB_CUDA(d_qcount.alloc(nc * 4, e.compute));
    FCLB_CUDA(d_qoff.alloc(nc * 4, e.compute));
    FCLB_CUDA(cudaMemsetAsync(d_qcount.p, 0, nc * 4, e.compute));
    FCLB_CUDA(cudaMemsetAsync(d_qoff.p, 0, nc * 4, e.compute));
    if (m > 0) {
      // (query, b1, b2) order: least significant key first, stable radix passes
      DevBuf d_ka, d_kb, d_va, d_vb, d_tmp;
      FCLB_CUDA(d_ka.alloc(m * 8, e.compute));
      FCLB_CUDA(d_kb.alloc(m * 8, e.compute));
      FCLB_CUDA(d_va.alloc(m * 4, e.compute));
      FCLB_CUDA(d_vb.alloc(m * 4, e.compute));
      size_t tmp_bytes = 0;
      FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_ka.as<unsigned long long>(), d_kb.as<unsigned long long>(),
                                                d_va.as<uint32_t>(), d_vb.as<uint32_t>(), int(m), 0, 64, e.compute));
      FCLB_CUDA(d_tmp.alloc(tmp_bytes, e.compute));
      const int g = int(std::min<size_t>((m + 255) / 256, size_t(e.sms) * 8));
      iotaKernel<<<g, 256, 0, e.compute>>>(d_va.as<uint32_t>(), m);
      uint32_t* cur = d_va.as<uint32_t>();
      uint32_t* nxt = d_vb.as<uint32_t>();
      auto pass = [&](int which, int bits) -> int {
        if (which == 2)
          gatherU32Kernel<<<g, 256, 0, e.compute>>>(d_cq.as<uint32_t>(), cur, m, d_ka.as<unsigned long long>());
        else
          gatherI64Kernel<<<g, 256, 0, e.compute>>>(which == 0 ? d_cb2.as<long long>() : d_cb1.as<long long>(), cur, m,
                                                    d_ka.as<unsigned long long>());
        FCLB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_ka.as<unsigned long long>(), d_kb.as<unsigned long long>(),
                                                  cur, nxt, int(m), 0, bits, e.compute));
        std::swap(cur, nxt);
        e.launches += 3;
        return FCLB_OK;
      };
      if (pair && (rc = pass(0, 64))) return rc;
      if ((rc = pass(1, 64))) return rc;
      if ((rc = pass(2, 32))) return rc;
      // the leaf batch
      const uint32_t n_user = pair ? 0u : t->n;
      DevBuf d_table, d_pairs, d_p1, d_p2;
      FCLB_CUDA(d_table.alloc((size_t(n_user) + 2 * m) * sizeof(ShapeD<S>), e.compute));
      FCLB_CUDA(d_pairs.alloc(m * sizeof(fclb_pair), e.compute));
      FCLB_CUDA(d_p1.alloc(m * ps, e.compute));
      FCLB_CUDA(d_p2.alloc(m * ps, e.compute));
      FCLB_CUDA(d_itemq.alloc(m * 4, e.compute));
      FCLB_CUDA(d_ib1.alloc(m * 8, e.compute));
      FCLB_CUDA(d_ib2.alloc(m * 8, e.compute));
      FCLB_CUDA(d_flags.alloc(m, e.compute));
      FCLB_CUDA(d_lc.alloc(m * 4 * 9 * sizeof(S), e.compute));
      FCLB_CUDA(d_lcnt.alloc(m * 4, e.compute));
      if (n_user)
        FCLB_CUDA(cudaMemcpyAsync(d_table.p, t->d_shapes[st], size_t(n_user) * sizeof(ShapeD<S>), cudaMemcpyDeviceToDevice,
                                  e.compute));
      LeafBuildArgs<S> b{};
      b.mode = mode;
      b.octree_pair = pair && kind1 == FCLB_SCENE_OCTREE && kind2 == FCLB_SCENE_OCTREE;
      b.m = m;
      b.q_base = c0;
      b.order = cur;
      b.cq = d_cq.as<uint32_t>();
      b.cb1 = d_cb1.as<long long>();
      b.cb2 = pair ? d_cb2.as<long long>() : nullptr;
      b.box1 = d_box1.as<S>();
      b.box2 = d_box2.as<S>();
      b.shape_ids = shape_ids;
      b.poses_a = static_cast<const S*>(poses_a);
      b.poses_b = static_cast<const S*>(poses_b);
      b.n_user = n_user;
      b.table = d_table.as<ShapeD<S>>();
      b.pairs = d_pairs.as<fclb_pair>();
      b.p1 = d_p1.as<S>();
      b.p2 = d_p2.as<S>();
      b.item_q = d_itemq.as<uint32_t>();
      b.item_b1 = d_ib1.as<long long>();
      b.item_b2 = d_ib2.as<long long>();
      b.item_flags = d_flags.as<uint8_t>();
      b.q_count = d_qcount.as<uint32_t>();
      leafBatchBuildKernel<S><<<g, 256, 0, e.compute>>>(b);
      e.launches += 2;
      FCLB_CUDA(cudaGetLastError());
      rc = collideLeafBatch(e, d_table.p, tris, d_pairs.as<fclb_pair>(), d_p1.p, d_p2.p, m, st, req, d_lc.p, d_lcnt.as<uint32_t>(),
                             uint32_t(n_user + 2 * m));
      if (rc) return rc;
      size_t scan_bytes = 0;
      FCLB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_qcount.as<uint32_t>(), d_qoff.as<uint32_t>(), int(nc), e.compute));
      DevBuf d_scan;
      FCLB_CUDA(d_scan.alloc(scan_bytes, e.compute));
      FCLB_CUDA(cub::DeviceScan::ExclusiveSum(d_scan.p, scan_bytes, d_qcount.as<uint32_t>(), d_qoff.as<uint32_t>(), int(nc), e.compute));
      e.launches += 1;
      FCLB_CUDA(cudaStreamSynchronize(e.compute));  // (d_scan and the sort buffers go out of scope below)
    }
    LeafScatterArgs<S> sc{};
    sc.n_chunk = nc;
    sc.q_base = c0;
    sc.q_off = d_qoff.as<uint32_t>();
    sc.q_count = d_qcount.as<uint32_t>();
    sc.item_b1 = d_ib1.as<long long>();
    sc.item_b2 = d_ib2.as<long long>();
    sc.item_flags = d_flags.as<uint8_t>();
    sc.leaf_contacts = d_lc.as<S>();
    sc.leaf_counts = d_lcnt.as<uint32_t>();
    sc.max_contacts = req->max_contacts;
    sc.max_keep = max_keep;
    sc.counts = counts;
    sc.out_b1 = out_b1;
    sc.out_b2 = out_b2;
    sc.out_contacts = static_cast<S*>(contacts);
    const int gs = int(std::min<size_t>((nc + 127) / 128, size_t(e.sms) * 16));
    leafBatchScatterKernel<S><<<gs, 128, 0, e.compute>>>(sc);
    e.launches += 1;
    FCLB_CUDA(cudaGetLastError());
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
  }
  return FCLB_OK;
}

// request.penetration_mode != Disabled against a scene geometry: MPR modes -> boolean traversal + per-contact MPR
// (sceneContactsDev); DefaultGJK_EPA -> candidate traversal + leaf batch (sceneGjkContactsDev)
static int sceneShapeContactsAny(Engine& e, int kind, fclb_handle scene, const ShapeTable* t, const uint32_t* shape_ids,
                                 const void* poses_scene, const void* poses_shape, size_t n, int scalar_type,
                                 const fclb_request* req, uint32_t max_keep, uint32_t* counts, long long* b1, void* contacts) {
  if (req->max_contacts == 0) {
    FCLB_CUDA(cudaMemsetAsync(counts, 0, n * sizeof(uint32_t), e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    return FCLB_OK;
  }
  if (req->penetration_mode == FCLB_PEN_DEFAULT_GJK_EPA) {
    if (scalar_type == FCLB_F32)
      return sceneGjkContactsDev<float>(e, kind, scene, -1, 0, t, shape_ids, poses_scene, poses_shape, n, req, max_keep, counts, b1,
                                        nullptr, contacts);
    return sceneGjkContactsDev<double>(e, kind, scene, -1, 0, t, shape_ids, poses_scene, poses_shape, n, req, max_keep, counts, b1,
                                       nullptr, contacts);
  }
  if (scalar_type == FCLB_F32)
    return sceneContactsDev<float>(e, kind, scene, t, shape_ids, poses_scene, poses_shape, n, req, max_keep, counts, b1, contacts);
  return sceneContactsDev<double>(e, kind, scene, t, shape_ids, poses_scene, poses_shape, n, req, max_keep, counts, b1, contacts);
}

static int scenePairContactsAny(Engine& e, int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                const void* poses2, size_t n, int scalar_type, const fclb_request* req, uint32_t max_keep,
                                uint32_t* counts, long long* b1, long long* b2, void* contacts) {
  if (req->max_contacts == 0) {
    FCLB_CUDA(cudaMemsetAsync(counts, 0, n * sizeof(uint32_t), e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    return FCLB_OK;
  }
  if (req->penetration_mode == FCLB_PEN_DEFAULT_GJK_EPA) {
    if (kind1 == FCLB_SCENE_OCTREE && kind2 == FCLB_SCENE_HEIGHTMAP)
      return fail(FCLB_ERR_UNSUPPORTED, "scene pair: pass the heightmap first (heightMapOctreeIntersect)");
    if (scalar_type == FCLB_F32)
      return sceneGjkContactsDev<float>(e, kind1, scene1, kind2, scene2, nullptr, nullptr, poses1, poses2, n, req, max_keep, counts,
                                        b1, b2, contacts);
    return sceneGjkContactsDev<double>(e, kind1, scene1, kind2, scene2, nullptr, nullptr, poses1, poses2, n, req, max_keep, counts,
                                       b1, b2, contacts);
  }
  if (scalar_type == FCLB_F32)
    return scenePairContactsDev<float>(e, kind1, scene1, kind2, scene2, poses1, poses2, n, req, max_keep, counts, b1, b2, contacts);
  return scenePairContactsDev<double>(e, kind1, scene1, kind2, scene2, poses1, poses2, n, req, max_keep, counts, b1, b2, contacts);
}

__global__ void narrowFirstIdKernel(const long long* __restrict__ b1, const uint32_t* __restrict__ counts, size_t n, int32_t* out32,
                                    long long* out64) {
  for (size_t q = blockIdx.x * size_t(blockDim.x) + threadIdx.x; q < n; q += size_t(gridDim.x) * blockDim.x) {
    const long long v = counts[q] ? b1[q] : -1;
    if (out32) out32[q] = int32_t(v);
    if (out64) out64[q] = v;
  }
}

// the boolean / counting entry points with a penetration request: the contact path with one kept contact per query
static int sceneShapeCountsViaContacts(Engine& e, int kind, fclb_handle scene, const ShapeTable* t, const uint32_t* shape_ids,
                                       const void* poses_scene, const void* poses_shape, size_t n, int scalar_type,
                                       const fclb_request* req, uint32_t* counts, int32_t* first32, long long* first64) {
  DevBuf tb1, tc;
  FCLB_CUDA(tb1.alloc(n * 8));
  FCLB_CUDA(tc.alloc(n * 7 * 8));
  const int rc = sceneShapeContactsAny(e, kind, scene, t, shape_ids, poses_scene, poses_shape, n, scalar_type, req, 1, counts,
                                       tb1.as<long long>(), tc.p);
  if (rc) return rc;
  if (first32 || first64) {
    const int g = int(std::min<size_t>((n + 255) / 256, size_t(e.sms) * 8));
    narrowFirstIdKernel<<<g, 256, 0, e.compute>>>(tb1.as<long long>(), counts, n, first32, first64);
    e.launches += 1;
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
  }
  return FCLB_OK;
}

// ---- octree2::Octree<S>::rebuildTree on the device (fclb_octree_build_dev.cuh) -------------------------------------
static int octAllocLevel(std::vector<OctLevel>& levels, std::vector<DevBuf>& bufs, int l, uint32_t n, cudaStream_t st) {
  OctLevel& v = levels[size_t(l)];
  v.n = n;
  FCLB_CUDA(bufs[size_t(4 * l)].alloc(size_t(n) * 8, st));
  FCLB_CUDA(bufs[size_t(4 * l + 1)].alloc(size_t(n) * 4, st));
  FCLB_CUDA(bufs[size_t(4 * l + 2)].alloc(size_t(n) * 4, st));
  FCLB_CUDA(bufs[size_t(4 * l + 3)].alloc(size_t(n) * 4, st));
  v.key = bufs[size_t(4 * l)].as<unsigned long long>();
  v.first = bufs[size_t(4 * l + 1)].as<uint32_t>();
  v.parent = bufs[size_t(4 * l + 2)].as<uint32_t>();
  v.index = bufs[size_t(4 * l + 3)].as<uint32_t>();
  return FCLB_OK;
}

template <typename S>
static int octreeBuildDev(Engine& e, const void* d_points, size_t n_points, double resolution, uint32_t bottom_half, OctreeDev* out) {
  int log2h = 0;
  while ((1u << log2h) < bottom_half) log2h++;
  const int L = log2h + 2;
  out->num_layers = L;
  const S res = S(resolution);
  const S inv = S(1.0) / res;
  const S mx = res * S(bottom_half);
  for (int k = 0; k < 3; k++) {
    out->root_box[k] = double(-mx);
    out->root_box[3 + k] = double(mx);
  }
  cudaStream_t st = e.compute;
  auto grid = [&](size_t n) { return int(std::min<size_t>((n + 255) / 256, size_t(e.sms) * 8)); };
  uint32_t n_valid = 0;
  DevBuf k0, i0, k1, i1, tmp, head, scanned;
  size_t tmp_bytes = 0;
  if (n_points) {
    FCLB_CUDA(k0.alloc(n_points * 8, st));
    FCLB_CUDA(i0.alloc(n_points * 4, st));
    FCLB_CUDA(k1.alloc(n_points * 8, st));
    FCLB_CUDA(i1.alloc(n_points * 4, st));
    octKeyKernel<S><<<grid(n_points), 256, 0, st>>>(static_cast<const S*>(d_points), n_points, inv, int(bottom_half), L,
                                                    k0.as<unsigned long long>(), i0.as<uint32_t>());
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0.as<unsigned long long>(), k1.as<unsigned long long>(), i0.as<uint32_t>(),
                                              i1.as<uint32_t>(), int(n_points), 0, 64, st));
    size_t scan_bytes = 0;
    FCLB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, static_cast<uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), int(n_points), st));
    FCLB_CUDA(tmp.alloc(std::max(tmp_bytes, scan_bytes), st));
    tmp_bytes = std::max(tmp_bytes, scan_bytes);
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k0.as<unsigned long long>(), k1.as<unsigned long long>(), i0.as<uint32_t>(),
                                              i1.as<uint32_t>(), int(n_points), 0, 64, st));
    e.launches += 2;
    // points outside the grid carry the key ~0 and sort last
    std::vector<unsigned long long> probe(1);
    size_t lo = 0, hi = n_points;  // first position whose key is ~0 (binary search with single-word reads)
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      FCLB_CUDA(cudaMemcpyAsync(probe.data(), k1.as<unsigned long long>() + mid, 8, cudaMemcpyDeviceToHost, st));
      FCLB_CUDA(cudaStreamSynchronize(st));
      if (probe[0] == ~0ull) hi = mid; else lo = mid + 1;
    }
    n_valid = uint32_t(lo);
    FCLB_CUDA(head.alloc(size_t(n_points) * 4, st));
    FCLB_CUDA(scanned.alloc(size_t(n_points) * 4, st));
  }
  // levels[l]: unique prefixes of l triplets; l = L - 1 voxels, l = L - 2 leaf nodes, 1 .. L - 3 inner nodes
  std::vector<OctLevel> levels(static_cast<size_t>(L));
  OctLevel* lv = levels.data();
  std::vector<DevBuf> bufs(static_cast<size_t>(L) * 4);
  auto allocLevel = [&](int l, uint32_t n) -> int { return octAllocLevel(levels, bufs, l, n, st); };
  auto lastOf = [&](const uint32_t* d_scanned, uint32_t n, uint32_t* out_v) -> int {
    FCLB_CUDA(cudaMemcpyAsync(out_v, d_scanned + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    FCLB_CUDA(cudaStreamSynchronize(st));
    return FCLB_OK;
  };
  int rc = FCLB_OK;
  if (n_valid) {
    uint32_t n_vox = 0;
    octHeadKernel<<<grid(n_valid), 256, 0, st>>>(k1.as<unsigned long long>(), n_valid, 0, head.as<uint32_t>());
    FCLB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, head.as<uint32_t>(), scanned.as<uint32_t>(), int(n_valid), st));
    if ((rc = lastOf(scanned.as<uint32_t>(), n_valid, &n_vox))) return rc;
    if ((rc = allocLevel(L - 1, n_vox))) return rc;
    octUniqueKernel<<<grid(n_valid), 256, 0, st>>>(k1.as<unsigned long long>(), i1.as<uint32_t>(), scanned.as<uint32_t>(), n_valid,
                                                   lv[L - 1].key, lv[L - 1].first);
    e.launches += 3;
    for (int l = L - 1; l >= 2; l--) {  // fold level l into level l - 1
      const uint32_t n = lv[l].n;
      uint32_t n_par = 0;
      octHeadKernel<<<grid(n), 256, 0, st>>>(lv[l].key, n, 3, head.as<uint32_t>());
      FCLB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, head.as<uint32_t>(), scanned.as<uint32_t>(), int(n), st));
      if ((rc = lastOf(scanned.as<uint32_t>(), n, &n_par))) return rc;
      if ((rc = allocLevel(l - 1, n_par))) return rc;
      octFillKernel<<<grid(n_par), 256, 0, st>>>(lv[l - 1].first, n_par, 0xffffffffu);
      octFoldKernel<<<grid(n), 256, 0, st>>>(lv[l].key, lv[l].first, scanned.as<uint32_t>(), n, 3, lv[l - 1].key, lv[l - 1].first, lv[l].parent);
      e.launches += 4;
    }
  }
  // ---- node numbering: creation time = (first point, depth)
  uint32_t n_inner = 1, n_leaf = n_valid ? lv[L - 2].n : 0;
  std::vector<uint32_t> offset(size_t(L), 0);
  uint32_t n_ranked = 0;
  for (int l = 1; l <= L - 3; l++) {
    offset[l] = n_ranked;
    n_ranked += lv[l].n;
  }
  n_inner += n_ranked;
  DevBuf rk0, rk1, rv0, rv1, flat;
  if (n_ranked) {
    FCLB_CUDA(rk0.alloc(size_t(n_ranked) * 8, st));
    FCLB_CUDA(rk1.alloc(size_t(n_ranked) * 8, st));
    FCLB_CUDA(rv0.alloc(size_t(n_ranked) * 4, st));
    FCLB_CUDA(rv1.alloc(size_t(n_ranked) * 4, st));
    FCLB_CUDA(flat.alloc(size_t(n_ranked) * 4, st));
    for (int l = 1; l <= L - 3; l++)
      if (lv[l].n) octRankKeyKernel<<<grid(lv[l].n), 256, 0, st>>>(lv[l].first, lv[l].n, l, offset[l], rk0.as<unsigned long long>(), rv0.as<uint32_t>());
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, rk0.as<unsigned long long>(), rk1.as<unsigned long long>(), rv0.as<uint32_t>(),
                                              rv1.as<uint32_t>(), int(n_ranked), 0, 40, st));
    octAssignIndexKernel<<<grid(n_ranked), 256, 0, st>>>(rv1.as<uint32_t>(), n_ranked, 1u, flat.as<uint32_t>());
    for (int l = 1; l <= L - 3; l++)
      if (lv[l].n) FCLB_CUDA(cudaMemcpyAsync(lv[l].index, flat.as<uint32_t>() + offset[l], size_t(lv[l].n) * 4, cudaMemcpyDeviceToDevice, st));
    e.launches += 3;
  }
  DevBuf lk0, lk1, lv0, lv1;
  if (n_leaf) {
    FCLB_CUDA(lk0.alloc(size_t(n_leaf) * 8, st));
    FCLB_CUDA(lk1.alloc(size_t(n_leaf) * 8, st));
    FCLB_CUDA(lv0.alloc(size_t(n_leaf) * 4, st));
    FCLB_CUDA(lv1.alloc(size_t(n_leaf) * 4, st));
    octRankKeyKernel<<<grid(n_leaf), 256, 0, st>>>(lv[L - 2].first, n_leaf, 0, 0, lk0.as<unsigned long long>(), lv0.as<uint32_t>());
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, lk0.as<unsigned long long>(), lk1.as<unsigned long long>(), lv0.as<uint32_t>(),
                                              lv1.as<uint32_t>(), int(n_leaf), 0, 40, st));
    octAssignIndexKernel<<<grid(n_leaf), 256, 0, st>>>(lv1.as<uint32_t>(), n_leaf, 0u, lv[L - 2].index);
    e.launches += 3;
  }
  // ---- the reference's flat arrays
  out->n_inner = n_inner;
  out->n_leaf = n_leaf;
  FCLB_CUDA(cudaMalloc(&out->children, size_t(32) * n_inner));
  FCLB_CUDA(cudaMalloc(&out->inner_full, n_inner));
  FCLB_CUDA(cudaMalloc(&out->leaf_bits, n_leaf ? n_leaf : 1));
  FCLB_CUDA(cudaMemsetAsync(out->children, 0xff, size_t(32) * n_inner, st));
  FCLB_CUDA(cudaMemsetAsync(out->inner_full, 0, n_inner, st));
  if (n_valid) {
    for (int l = 1; l <= L - 2; l++) {
      if (!lv[l].n) continue;
      if (l == 1) {
        // (level 1 has no parent array: every item hangs off the root)
        octFillKernel<<<grid(lv[l].n), 256, 0, st>>>(lv[l].parent, lv[l].n, 0u);
      }
      octLinkKernel<<<grid(lv[l].n), 256, 0, st>>>(lv[l].key, lv[l].parent, lv[l].index, lv[l].n, l == 1 ? nullptr : lv[l - 1].index,
                                                   out->children);
      e.launches += 1;
    }
    DevBuf bits32;
    FCLB_CUDA(bits32.alloc(size_t(n_leaf) * 4, st));
    FCLB_CUDA(cudaMemsetAsync(bits32.p, 0, size_t(n_leaf) * 4, st));
    octLeafBitsKernel<<<grid(lv[L - 1].n), 256, 0, st>>>(lv[L - 1].key, lv[L - 1].parent, lv[L - 1].n, lv[L - 2].index, bits32.as<uint32_t>());
    octNarrowKernel<<<grid(n_leaf), 256, 0, st>>>(bits32.as<uint32_t>(), n_leaf, out->leaf_bits);
    for (int d = L - 3; d >= 0; d--) {
      const uint32_t n = d == 0 ? 1u : lv[d].n;
      if (n) octFullKernel<<<grid(n), 256, 0, st>>>(d == 0 ? nullptr : lv[d].index, n, out->children, d == L - 3 ? 1 : 0, out->leaf_bits, out->inner_full);
    }
    e.launches += 3;
    FCLB_CUDA(cudaGetLastError());
    FCLB_CUDA(cudaStreamSynchronize(st));  // (bits32 and the level buffers go out of scope)
  }
  FCLB_CUDA(cudaGetLastError());
  FCLB_CUDA(cudaStreamSynchronize(st));
  return FCLB_OK;
}

static int octreeBuildDevEntry(const void* d_points, size_t n_points, double resolution, uint32_t bottom_half, int scalar_type,
                               fclb_handle* octree) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!octree || (n_points && !d_points) || bottom_half < 2 || bottom_half > 16384 || (bottom_half & (bottom_half - 1)) ||
      !(resolution > 0) || n_points > 0x7fffffffu)
    return fail(FCLB_ERR_BAD_ARG, "fclb_octree_build_dev: bad argument (half shape: power of two >= 2)");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  OctreeDev* d = new OctreeDev();
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  rc = scalar_type == FCLB_F32 ? octreeBuildDev<float>(e, d_points, n_points, resolution, bottom_half, d)
                               : octreeBuildDev<double>(e, d_points, n_points, resolution, bottom_half, d);
  if (rc) {
    cudaFree(d->children);
    cudaFree(d->inner_full);
    cudaFree(d->leaf_bits);
    delete d;
    return rc;
  }
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = e.last_call_ms = ms;
  const fclb_handle h = newHandle();
  octTable()[h] = d;
  *octree = h;
  return FCLB_OK;
}

template <typename S>
static void heightsFromPoints(const double* pts, size_t n, S res_x, S res_y, uint32_t half_x, uint32_t half_y,
                              uint16_t* heights) {
  const uint32_t full_x = 2 * half_x, full_y = 2 * half_y;
  for (size_t i = 0; i < n; i++) {
    const S point_x = S(pts[3 * i]), point_y = S(pts[3 * i + 1]), point_z = S(pts[3 * i + 2]);
    if (point_z < 0) continue;
    const int x = floor(point_x / res_x) + uint16_t(half_x);
    const int y = floor(point_y / res_y) + uint16_t(half_y);
    const bool in_range = (x >= 0 && x < int(full_x) && y >= 0 && y < int(full_y));
    if (in_range) {
      const size_t index = size_t(uint16_t(y)) * full_x + uint16_t(x);
      const int z = static_cast<uint16_t>(point_z * 1000);
      if (heights[index] < z) heights[index] = static_cast<uint16_t>(z);
    }
  }
}


// ---- LayeredHeightMap construction on the device ------------------------------------------------
// FlatHeightMap<S>::updateHeightsByPointGenerationFunctor (flat_heightmap-inl.h:249-272) keeps, per pixel, the
// maximum of uint16(z * 1000) over the points that fall into it: a maximum is order-free, so one thread per point
// with an atomicMax gives the sequential loop's result exactly.  Heights are accumulated in a 32-bit grid (no 16-bit
// atomics), packed to the uint16 bottom layer, and the coarser layers follow by 2x2 max pooling
// (LayeredHeightMap::rebuildNextLayer, layered_heightmap-inl.h:77-101), one launch per layer.
template <typename S>
__global__ void hmRasterKernel(const S* __restrict__ pts, size_t n, S res_x, S res_y, int half_x, int half_y,
                               uint32_t* __restrict__ grid) {
  const int full_x = 2 * half_x, full_y = 2 * half_y;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const S px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
    if (pz < 0) continue;
    const int x = int(floor(double(px / res_x)) + double(half_x));
    const int y = int(floor(double(py / res_y)) + double(half_y));
    if (x >= 0 && x < full_x && y >= 0 && y < full_y) {
      const uint32_t z = uint32_t(uint16_t(int(pz * S(1000))));
      atomicMax(&grid[size_t(y) * full_x + x], z);
    }
  }
}
__global__ void hmPackKernel(const uint32_t* __restrict__ grid, uint16_t* __restrict__ bottom, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const uint32_t g = grid[i];
    const uint16_t b = bottom[i];
    bottom[i] = g > b ? uint16_t(g) : b;  // `if (heights[index] < z) heights[index] = z` on top of earlier heights
  }
}
__global__ void hmPoolKernel(const uint16_t* __restrict__ down, uint32_t dfx, uint16_t* __restrict__ up, uint32_t ufx,
                             uint32_t ufy) {
  const size_t total = size_t(ufx) * ufy;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const uint32_t x = uint32_t(i % ufx), y = uint32_t(i / ufx);
    const uint16_t* r0 = down + size_t(2 * y) * dfx + 2 * x;
    const uint16_t* r1 = r0 + dfx;
    const uint16_t a = r0[0] > r0[1] ? r0[0] : r0[1];
    const uint16_t b = r1[0] > r1[1] ? r1[0] : r1[1];
    up[i] = a > b ? a : b;
  }
}

// layer shapes / offsets of a LayeredHeightMap with the given bottom half shape
static void hmLayout(HeightmapDev* d, uint32_t hx, uint32_t hy, size_t* total) {
  d->half_x = hx;
  d->half_y = hy;
  d->fx.clear();
  d->fy.clear();
  d->off.clear();
  uint32_t fx = 2 * hx, fy = 2 * hy, nx = hx, ny = hy;
  size_t t = 0;
  while (true) {
    d->fx.push_back(fx);
    d->fy.push_back(fy);
    d->off.push_back(t);
    t += size_t(fx) * fy;
    if (!(nx > 1 && ny > 1)) break;
    fx /= 2;
    fy /= 2;
    nx /= 2;
    ny /= 2;
  }
  *total = t;
}

template <typename S>
static int heightmapBuildDev(Engine& e, const void* d_points, size_t n_points, double res_x, double res_y, uint32_t hx,
                             uint32_t hy, fclb_handle* hm) {
  HeightmapDev* d = new HeightmapDev();
  d->res_x = res_x;
  d->res_y = res_y;
  size_t total = 0;
  hmLayout(d, hx, hy, &total);
  const size_t n_px = size_t(d->fx[0]) * d->fy[0];
  uint32_t* grid = nullptr;
  if (cudaMalloc(&d->d_layers, total * sizeof(uint16_t)) != cudaSuccess || cudaMalloc(&grid, n_px * sizeof(uint32_t)) != cudaSuccess) {
    cudaFree(d->d_layers);
    delete d;
    return fail(FCLB_ERR_CUDA, "fclb_heightmap_build: cudaMalloc failed");
  }
  cudaMemsetAsync(d->d_layers, 0, n_px * sizeof(uint16_t), e.compute);
  cudaMemsetAsync(grid, 0, n_px * sizeof(uint32_t), e.compute);
  const int cap = e.sms * 16;
  auto blocks = [&](size_t n) { return int(n / 256 + 1 < size_t(cap) ? n / 256 + 1 : size_t(cap)); };
  hmRasterKernel<S><<<blocks(n_points), 256, 0, e.compute>>>(static_cast<const S*>(d_points), n_points, S(res_x), S(res_y),
                                                            int(hx), int(hy), grid);
  hmPackKernel<<<blocks(n_px), 256, 0, e.compute>>>(grid, d->d_layers, n_px);
  for (size_t k = 1; k < d->off.size(); k++)
    hmPoolKernel<<<blocks(size_t(d->fx[k]) * d->fy[k]), 256, 0, e.compute>>>(d->d_layers + d->off[k - 1], d->fx[k - 1],
                                                                           d->d_layers + d->off[k], d->fx[k], d->fy[k]);
  e.launches += 2 + (d->off.size() - 1);
  // height_upper_bound_in_mm = the maximum height = the maximum of the (tiny) top layer
  const size_t top = d->off.size() - 1;
  std::vector<uint16_t> h_top(size_t(d->fx[top]) * d->fy[top]);
  cudaError_t ce = cudaMemcpyAsync(h_top.data(), d->d_layers + d->off[top], h_top.size() * sizeof(uint16_t),
                                   cudaMemcpyDeviceToHost, e.compute);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e.compute);
  cudaFree(grid);
  if (ce != cudaSuccess) {
    cudaFree(d->d_layers);
    delete d;
    return fail(FCLB_ERR_CUDA, cudaGetErrorString(ce));
  }
  uint32_t mx = 0;
  for (uint16_t h : h_top) mx = h > mx ? h : mx;
  d->upper_mm = mx;
  const fclb_handle h = newHandle();
  hmTable()[h] = d;
  *hm = h;
  return FCLB_OK;
}

}  // namespace fclb

using namespace fclb;

// The host pipeline shared by the three scene-vs-shape boolean entry points (mesh, heightmap, octree): every stage's
// shape ids + poses are queued on the copy-in stream up front, the compute stream waits per stage and runs `dev` on it,
// the copy-out stream drains a stage's counts / first ids while later stages upload and traverse.
// ONE stage by default: measured on C4 (700k + 700k queries, profiles/r02_host_head_ab.txt) the traversal kernels lose more
// on part-batches (a query costs 1 - 100x the median, and every launch ends with a tail of the expensive ones) than the
// overlapped upload saves -- 41.6 ms in one stage, 42.4 ms in equal 512k stages, 42.9 ms with a short first stage.
// FCLB_SCENE_HOST_STAGED=1 turns the stages on (first stage e.host_head, doubling up to a quarter of e.host_chunk).
// first_bytes: size of one "first contact id" record (4, or 8 for octree node codes).
template <typename DevCall>
static int sceneShapeHostStaged(Engine& e, const uint32_t* shape_ids, const void* poses_scene, const void* poses_shape, size_t n,
                                int scalar_type, uint32_t* out_counts, void* out_first, size_t first_bytes, DevCall dev) {
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_ids = 0;
  const size_t o_p1 = alignUp(o_ids + n * 4, 256);
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_first = alignUp(o_cnt + n * 4, 256);
  const size_t total = alignUp(o_first + n * first_bytes, 256);
  int rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  std::vector<size_t> c_begin, c_size;
  static const bool staged = getenv("FCLB_SCENE_HOST_STAGED") != nullptr;
  stageSizes(n, staged ? (e.host_chunk / 4 ? e.host_chunk / 4 : 1) : n, staged ? e.host_head : 0, 0, c_begin, c_size);
  const int n_chunks = int(c_size.size());
  rc = ensureChunkEvents(e, n_chunks);
  if (rc) return rc;
  const char* h_p1 = static_cast<const char*>(poses_scene);
  const char* h_p2 = static_cast<const char*>(poses_shape);
  FCLB_CUDA(cudaStreamSynchronize(e.copy_out));  // the staging arena may still be read by an earlier call's copy-out
  for (int c = 0; c < n_chunks; c++) {
    const size_t b0 = c_begin[c], m = c_size[c];
    FCLB_CUDA(cudaMemcpyAsync(base + o_ids + b0 * 4, shape_ids + b0, m * 4, cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaMemcpyAsync(base + o_p1 + b0 * 12 * ss, h_p1 + b0 * 12 * ss, m * 12 * ss, cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaMemcpyAsync(base + o_p2 + b0 * 12 * ss, h_p2 + b0 * 12 * ss, m * 12 * ss, cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaEventRecord(e.ev_in[c], e.copy_in));
  }
  unsigned long long visits[2] = {0, 0};
  for (int c = 0; c < n_chunks; c++) {
    const size_t b0 = c_begin[c], m = c_size[c];
    FCLB_CUDA(cudaStreamWaitEvent(e.compute, e.ev_in[c], 0));
    rc = dev(reinterpret_cast<const uint32_t*>(base + o_ids) + b0, base + o_p1 + b0 * 12 * ss, base + o_p2 + b0 * 12 * ss, m,
             reinterpret_cast<uint32_t*>(base + o_cnt) + b0, out_first ? base + o_first + b0 * first_bytes : nullptr);
    if (rc) {  // drain the queued copies before the caller gets its buffers back
      cudaStreamSynchronize(e.copy_in);
      cudaStreamSynchronize(e.copy_out);
      return rc;
    }
    visits[0] += g_stats[0];
    visits[1] += g_stats[1];
    FCLB_CUDA(cudaEventRecord(e.ev_done[c], e.compute));
    FCLB_CUDA(cudaStreamWaitEvent(e.copy_out, e.ev_done[c], 0));
    FCLB_CUDA(cudaMemcpyAsync(out_counts + b0, base + o_cnt + b0 * 4, m * 4, cudaMemcpyDeviceToHost, e.copy_out));
    if (out_first)
      FCLB_CUDA(cudaMemcpyAsync(static_cast<char*>(out_first) + b0 * first_bytes, base + o_first + b0 * first_bytes, m * first_bytes,
                                cudaMemcpyDeviceToHost, e.copy_out));
  }
  g_stats[0] = visits[0];  // fclb_scene_last_visit_counts: the whole call
  g_stats[1] = visits[1];
  FCLB_CUDA(cudaStreamSynchronize(e.copy_out));
  return FCLB_OK;
}

extern "C" {

int fclb_bvh_shape_collide_batch_dev(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids, const void* poses_mesh,
                                     const void* poses_shape, size_t n, int scalar_type, const fclb_request* req,
                                     uint32_t* out_counts, int32_t* out_first_tri) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = bvhTable().find(bvh);
  if (it == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown BVH handle");
  ShapeTable* t = findTable(e, shapes);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown shape table handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (it->second->scalar_type != scalar_type) return fail(FCLB_ERR_BAD_ARG, "BVH was uploaded for a different scalar type");
  if (!req || !out_counts) return fail(FCLB_ERR_BAD_ARG, "null request / out_counts");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_mesh || !poses_shape) return fail(FCLB_ERR_BAD_ARG, "null input array");
  if (req->penetration_mode != FCLB_PEN_DISABLED)  // contacts are generated (and dropped): numContacts is the contact path's
    return sceneShapeCountsViaContacts(e, FCLB_SCENE_BVH, bvh, t, shape_ids, poses_mesh, poses_shape, n, scalar_type, req,
                                       out_counts, out_first_tri, nullptr);
  if (scalar_type == FCLB_F32)
    return bvhShapeDev<float>(e, it->second, t, shape_ids, poses_mesh, poses_shape, n, req, out_counts, out_first_tri);
  return bvhShapeDev<double>(e, it->second, t, shape_ids, poses_mesh, poses_shape, n, req, out_counts, out_first_tri);
}

static int bvh_shape_collide_batch_host_one(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids, const void* poses_mesh,
                                      const void* poses_shape, size_t n, int scalar_type, const fclb_request* req,
                                      uint32_t* out_counts, int32_t* out_first_tri) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_mesh || !poses_shape || !out_counts) return fail(FCLB_ERR_BAD_ARG, "null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  {
    ShapeTable* t = findTable(e, shapes);
    if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown shape table handle");
    for (size_t q = 0; q < n; q++)
      if (shape_ids[q] >= t->n) return fail(FCLB_ERR_BAD_ARG, "shape id out of range");
  }
  return sceneShapeHostStaged(e, shape_ids, poses_mesh, poses_shape, n, scalar_type, out_counts, out_first_tri, 4,
                              [&](const uint32_t* d_ids, const void* d_p1, const void* d_p2, size_t m, uint32_t* d_cnt, void* d_first) {
                                return fclb_bvh_shape_collide_batch_dev(bvh, shapes, d_ids, d_p1, d_p2, m, scalar_type, req, d_cnt,
                                          static_cast<int32_t*>(d_first));
                              });
}
int fclb_bvh_shape_collide_batch_host(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids, const void* poses_mesh,
                                      const void* poses_shape, size_t n, int scalar_type, const fclb_request* req,
                                      uint32_t* out_counts, int32_t* out_first_tri) {
  if (engineCount() <= 1) return bvh_shape_collide_batch_host_one(bvh, shapes, shape_ids, poses_mesh, poses_shape, n, scalar_type, req, out_counts, out_first_tri);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return bvh_shape_collide_batch_host_one(bvh, shapes, offT(shape_ids, b), offPtr(poses_mesh, b * 12 * ss), offPtr(poses_shape, b * 12 * ss), m_, scalar_type, req, offT(out_counts, b), offT(out_first_tri, b)); });
}

int fclb_heightmap_build_host(const double* points, size_t n_points, double resolution_x, double resolution_y,
                              uint32_t half_shape_x, uint32_t half_shape_y, int scalar_type, uint16_t* heights_mm) {
  if (!points || !heights_mm || half_shape_x == 0 || half_shape_y == 0 || half_shape_x > 32767 || half_shape_y > 32767)
    return fail(FCLB_ERR_BAD_ARG, "fclb_heightmap_build_host: bad argument");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (scalar_type == FCLB_F32)
    heightsFromPoints<float>(points, n_points, float(resolution_x), float(resolution_y), half_shape_x, half_shape_y,
                             heights_mm);
  else
    heightsFromPoints<double>(points, n_points, resolution_x, resolution_y, half_shape_x, half_shape_y, heights_mm);
  return FCLB_OK;
}

static int heightmap_upload_one(const uint16_t* heights_mm, uint32_t full_x, uint32_t full_y, double resolution_x,
                          double resolution_y, uint32_t upper_bound_mm, fclb_handle* hm) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!heights_mm || !hm || full_x < 2 || full_y < 2 || (full_x & 1) || (full_y & 1) || full_x > 65534 || full_y > 65534)
    return fail(FCLB_ERR_BAD_ARG, "fclb_heightmap_upload: bad shape");
  const uint32_t hx = full_x / 2, hy = full_y / 2;
  if ((hx & (hx - 1)) || (hy & (hy - 1)))
    return fail(FCLB_ERR_BAD_ARG, "fclb_heightmap_upload: half shapes must be powers of two (layered_heightmap-inl.h:15-16)");
  if (!(resolution_x > 0) || !(resolution_y > 0)) return fail(FCLB_ERR_BAD_ARG, "fclb_heightmap_upload: bad resolution");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  HeightmapDev* d = new HeightmapDev();
  d->half_x = hx;
  d->half_y = hy;
  d->res_x = resolution_x;
  d->res_y = resolution_y;
  // layers: bottom, then 2x2 max pooling while both half shapes exceed 1 (layered_heightmap-inl.h:17-48,77-101)
  std::vector<std::vector<uint16_t>> layers;
  layers.emplace_back(heights_mm, heights_mm + size_t(full_x) * full_y);
  d->fx.push_back(full_x);
  d->fy.push_back(full_y);
  uint32_t nx = hx, ny = hy;
  while (nx > 1 && ny > 1) {
    const std::vector<uint16_t>& down = layers.back();
    const uint32_t dfx = d->fx.back(), dfy = d->fy.back();
    const uint32_t ufx = dfx / 2, ufy = dfy / 2;
    std::vector<uint16_t> up(size_t(ufx) * ufy, 0);
    for (uint32_t y = 0; y < dfy; y++)
      for (uint32_t x = 0; x < dfx; x++) {
        const uint16_t h = down[size_t(y) * dfx + x];
        uint16_t& u = up[size_t(y / 2) * ufx + x / 2];
        if (h > u) u = h;
      }
    layers.push_back(std::move(up));
    d->fx.push_back(ufx);
    d->fy.push_back(ufy);
    nx /= 2;
    ny /= 2;
  }
  uint32_t mx = 0;
  for (uint16_t h : layers[0]) mx = h > mx ? h : mx;
  d->upper_mm = upper_bound_mm > mx ? upper_bound_mm : mx;
  size_t total = 0;
  for (auto& l : layers) {
    d->off.push_back(total);
    total += l.size();
  }
  if (cudaMalloc(&d->d_layers, total * sizeof(uint16_t)) != cudaSuccess) {
    delete d;
    return fail(FCLB_ERR_CUDA, "fclb_heightmap_upload: cudaMalloc failed");
  }
  for (size_t k = 0; k < layers.size(); k++)
    uploadSync(d->d_layers + d->off[k], layers[k].data(), layers[k].size() * sizeof(uint16_t));
  const fclb_handle h = newHandle();
  hmTable()[h] = d;
  *hm = h;
  return FCLB_OK;
}
int fclb_heightmap_upload(const uint16_t* heights_mm, uint32_t full_x, uint32_t full_y, double resolution_x,
                          double resolution_y, uint32_t upper_bound_mm, fclb_handle* hm) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return heightmap_upload_one(heights_mm, full_x, full_y, resolution_x, resolution_y, upper_bound_mm, hm); });
}

int fclb_heightmap_build_dev(const void* points, size_t n_points, double resolution_x, double resolution_y,
                             uint32_t half_shape_x, uint32_t half_shape_y, int scalar_type, fclb_handle* hm) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!hm || (n_points && !points) || half_shape_x == 0 || half_shape_y == 0 || half_shape_x > 32767 || half_shape_y > 32767 ||
      (half_shape_x & (half_shape_x - 1)) || (half_shape_y & (half_shape_y - 1)))
    return fail(FCLB_ERR_BAD_ARG, "fclb_heightmap_build_dev: bad argument (half shapes must be powers of two)");
  if (!(resolution_x > 0) || !(resolution_y > 0)) return fail(FCLB_ERR_BAD_ARG, "fclb_heightmap_build_dev: bad resolution");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  if (scalar_type == FCLB_F32)
    return heightmapBuildDev<float>(e, points, n_points, double(float(resolution_x)), double(float(resolution_y)), half_shape_x,
                                    half_shape_y, hm);
  return heightmapBuildDev<double>(e, points, n_points, resolution_x, resolution_y, half_shape_x, half_shape_y, hm);
}

static int heightmap_build_points_host_one(const void* points, size_t n_points, double resolution_x, double resolution_y,
                                     uint32_t half_shape_x, uint32_t half_shape_y, int scalar_type, fclb_handle* hm) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n_points && !points) return fail(FCLB_ERR_BAD_ARG, "null points");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t bytes = n_points * 3 * (scalar_type == FCLB_F32 ? 4 : 8);
  rc = ensureStage(e, bytes ? bytes : 16);
  if (rc) return rc;
  if (bytes) FCLB_CUDA(cudaMemcpyAsync(e.d_stage, points, bytes, cudaMemcpyHostToDevice, e.compute));
  return fclb_heightmap_build_dev(e.d_stage, n_points, resolution_x, resolution_y, half_shape_x, half_shape_y, scalar_type, hm);
}
int fclb_heightmap_build_points_host(const void* points, size_t n_points, double resolution_x, double resolution_y,
                                     uint32_t half_shape_x, uint32_t half_shape_y, int scalar_type, fclb_handle* hm) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return heightmap_build_points_host_one(points, n_points, resolution_x, resolution_y, half_shape_x, half_shape_y, scalar_type, hm); });
}

int fclb_heightmap_info(fclb_handle hm, uint32_t* n_layers, uint32_t* full_x, uint32_t* full_y, uint32_t* upper_bound_mm) {
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = hmTable().find(hm);
  if (it == hmTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_heightmap_info: unknown handle");
  if (n_layers) *n_layers = uint32_t(it->second->off.size());
  if (full_x) *full_x = it->second->fx[0];
  if (full_y) *full_y = it->second->fy[0];
  if (upper_bound_mm) *upper_bound_mm = it->second->upper_mm;
  return FCLB_OK;
}

int fclb_heightmap_export(fclb_handle hm, uint32_t layer, uint16_t* heights_mm) {
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = hmTable().find(hm);
  if (it == hmTable().end() || !heights_mm) return fail(FCLB_ERR_BAD_ARG, "fclb_heightmap_export: unknown handle / null buffer");
  const HeightmapDev* d = it->second;
  if (layer >= d->off.size()) return fail(FCLB_ERR_BAD_ARG, "fclb_heightmap_export: no such layer");
  FCLB_CUDA(cudaMemcpy(heights_mm, d->d_layers + d->off[layer], size_t(d->fx[layer]) * d->fy[layer] * sizeof(uint16_t),
                       cudaMemcpyDeviceToHost));
  return FCLB_OK;
}

static int heightmap_release_one(fclb_handle h) {
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = hmTable().find(h);
  if (it == hmTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_heightmap_release: unknown handle");
  cudaFree(it->second->d_layers);
  delete it->second;
  hmTable().erase(it);
  return FCLB_OK;
}
int fclb_heightmap_release(fclb_handle h) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return heightmap_release_one(h); });
}

int fclb_heightmap_shape_collide_batch_dev(fclb_handle hm, fclb_handle shapes, const uint32_t* shape_ids,
                                           const void* poses_hm, const void* poses_shape, size_t n, int scalar_type,
                                           const fclb_request* req, uint32_t* out_counts, int32_t* out_first_pixel) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = hmTable().find(hm);
  if (it == hmTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown heightmap handle");
  ShapeTable* t = findTable(e, shapes);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown shape table handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!req || !out_counts) return fail(FCLB_ERR_BAD_ARG, "null request / out_counts");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_hm || !poses_shape) return fail(FCLB_ERR_BAD_ARG, "null input array");
  if (req->penetration_mode != FCLB_PEN_DISABLED)
    return sceneShapeCountsViaContacts(e, FCLB_SCENE_HEIGHTMAP, hm, t, shape_ids, poses_hm, poses_shape, n, scalar_type, req,
                                       out_counts, out_first_pixel, nullptr);
  if (scalar_type == FCLB_F32)
    return heightmapShapeDev<float>(e, it->second, t, shape_ids, poses_hm, poses_shape, n, req, out_counts,
                                    out_first_pixel);
  return heightmapShapeDev<double>(e, it->second, t, shape_ids, poses_hm, poses_shape, n, req, out_counts,
                                   out_first_pixel);
}

static int heightmap_shape_collide_batch_host_one(fclb_handle hm, fclb_handle shapes, const uint32_t* shape_ids,
                                            const void* poses_hm, const void* poses_shape, size_t n, int scalar_type,
                                            const fclb_request* req, uint32_t* out_counts, int32_t* out_first_pixel) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_hm || !poses_shape || !out_counts) return fail(FCLB_ERR_BAD_ARG, "null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  {
    ShapeTable* t = findTable(e, shapes);
    if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown shape table handle");
    for (size_t q = 0; q < n; q++)
      if (shape_ids[q] >= t->n) return fail(FCLB_ERR_BAD_ARG, "shape id out of range");
  }
  return sceneShapeHostStaged(e, shape_ids, poses_hm, poses_shape, n, scalar_type, out_counts, out_first_pixel, 4,
                              [&](const uint32_t* d_ids, const void* d_p1, const void* d_p2, size_t m, uint32_t* d_cnt, void* d_first) {
                                return fclb_heightmap_shape_collide_batch_dev(hm, shapes, d_ids, d_p1, d_p2, m, scalar_type, req, d_cnt,
                                          static_cast<int32_t*>(d_first));
                              });
}
int fclb_heightmap_shape_collide_batch_host(fclb_handle hm, fclb_handle shapes, const uint32_t* shape_ids,
                                            const void* poses_hm, const void* poses_shape, size_t n, int scalar_type,
                                            const fclb_request* req, uint32_t* out_counts, int32_t* out_first_pixel) {
  if (engineCount() <= 1) return heightmap_shape_collide_batch_host_one(hm, shapes, shape_ids, poses_hm, poses_shape, n, scalar_type, req, out_counts, out_first_pixel);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return heightmap_shape_collide_batch_host_one(hm, shapes, offT(shape_ids, b), offPtr(poses_hm, b * 12 * ss), offPtr(poses_shape, b * 12 * ss), m_, scalar_type, req, offT(out_counts, b), offT(out_first_pixel, b)); });
}

static int octree_upload_one(const uint32_t* inner_children, const uint8_t* inner_full, uint32_t n_inner, const uint8_t* leaf_bits,
                       uint32_t n_leaf, const uint8_t* pruned_or_null, const double* root_aabb, int num_layers,
                       fclb_handle* octree) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!octree || !root_aabb || num_layers < 2 || (n_inner && (!inner_children || !inner_full)) || (n_leaf && !leaf_bits))
    return fail(FCLB_ERR_BAD_ARG, "fclb_octree_upload: bad argument");
  if (n_inner && !fclb::hostbuild::octLinksInRange(inner_children, n_inner, n_leaf, num_layers, nullptr))
    return fail(FCLB_ERR_BAD_ARG, "fclb_octree_upload: child index out of range");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  OctreeDev* d = new OctreeDev();
  d->n_inner = n_inner;
  d->n_leaf = n_leaf;
  d->num_layers = num_layers;
  for (int k = 0; k < 6; k++) d->root_box[k] = root_aabb[k];
  auto up = [](auto** dst, const void* src, size_t bytes) -> bool {
    if (cudaMalloc(reinterpret_cast<void**>(dst), bytes ? bytes : 1) != cudaSuccess) return false;
    return !bytes || uploadSync(*dst, src, bytes) == cudaSuccess;
  };
  bool ok = up(&d->children, inner_children, size_t(32) * n_inner) && up(&d->inner_full, inner_full, n_inner) &&
            up(&d->leaf_bits, leaf_bits, n_leaf);
  if (ok && pruned_or_null) ok = up(&d->pruned, pruned_or_null, n_inner);
  if (!ok) {
    delete d;
    return fail(FCLB_ERR_CUDA, "fclb_octree_upload: device allocation / copy failed");
  }
  const fclb_handle h = newHandle();
  octTable()[h] = d;
  *octree = h;
  return FCLB_OK;
}
int fclb_octree_upload(const uint32_t* inner_children, const uint8_t* inner_full, uint32_t n_inner, const uint8_t* leaf_bits,
                       uint32_t n_leaf, const uint8_t* pruned_or_null, const double* root_aabb, int num_layers,
                       fclb_handle* octree) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return octree_upload_one(inner_children, inner_full, n_inner, leaf_bits, n_leaf, pruned_or_null, root_aabb, num_layers, octree); });
}

namespace {
int octreeBuildHostImpl(const double* points, size_t n_points, double resolution, uint32_t bottom_half_shape, int scalar_type,
                        fclb::hostbuild::OctreeHost& t, const char* who) {
  if ((n_points && !points) || bottom_half_shape < 2 || bottom_half_shape > 16384 ||
      (bottom_half_shape & (bottom_half_shape - 1)) || !(resolution > 0) || n_points > 0x7fffffffu)
    return fail(FCLB_ERR_BAD_ARG, who);
  if (scalar_type == FCLB_F32)
    fclb::hostbuild::octreeFromPoints<float>(points, n_points, float(resolution), bottom_half_shape, t);
  else if (scalar_type == FCLB_F64)
    fclb::hostbuild::octreeFromPoints<double>(points, n_points, resolution, bottom_half_shape, t);
  else
    return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  return FCLB_OK;
}
}  // namespace

int fclb_octree_build_host(const double* points, size_t n_points, double resolution, uint32_t bottom_half_shape,
                           int scalar_type, uint32_t* inner_children, uint8_t* inner_full, uint32_t inner_capacity,
                           uint32_t* n_inner, uint8_t* leaf_bits, uint32_t leaf_capacity, uint32_t* n_leaf,
                           double* root_aabb, int* num_layers) {
  if (!n_inner || !n_leaf) return fail(FCLB_ERR_BAD_ARG, "fclb_octree_build_host: bad argument");
  // The size query keeps its tree for the data call that follows on the same thread with the same arguments AND the
  // same point data (a 64-bit hash of the whole buffer: a cloud rewritten in place, or another buffer at the same
  // address, is a different call), so the usual two-call sequence inserts the points once.
  struct Stash {
    bool valid = false;
    const double* points = nullptr;
    size_t n = 0;
    double res = 0;
    uint64_t hash = 0;
    uint32_t half = 0;
    int st = 0;
    fclb::hostbuild::OctreeHost tree;
  };
  static thread_local Stash stash;
  auto hashPoints = [&]() {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ uint64_t(n_points);
    const size_t words = n_points * 3;
    for (size_t i = 0; i < words; i++) {
      uint64_t w;
      std::memcpy(&w, points + i, 8);
      h = (h ^ w) * 0xff51afd7ed558ccdull;
      h ^= h >> 29;
    }
    return h;
  };
  auto sameCall = [&]() {
    return stash.valid && stash.points == points && stash.n == n_points && stash.res == resolution &&
           stash.half == bottom_half_shape && stash.st == scalar_type && stash.hash == hashPoints();
  };
  const bool size_query = !inner_children || !inner_full || !leaf_bits;
  fclb::hostbuild::OctreeHost local;
  fclb::hostbuild::OctreeHost& t = (size_query || sameCall()) ? stash.tree : local;
  if (!(&t == &stash.tree && !size_query)) {  // not a reuse: build
    stash.valid = false;
    int rc = octreeBuildHostImpl(points, n_points, resolution, bottom_half_shape, scalar_type, t,
                                 "fclb_octree_build_host: bad argument (half shape: power of two >= 2)");
    if (rc) return rc;
    if (size_query) {
      stash.valid = true;
      stash.points = points;
      stash.n = n_points;
      stash.res = resolution;
      stash.half = bottom_half_shape;
      stash.st = scalar_type;
      stash.hash = hashPoints();
    }
  }
  struct Drop {  // a data call consumes the stash whatever its outcome
    Stash& s;
    bool drop;
    ~Drop() {
      if (drop) {
        s.valid = false;
        s.tree = fclb::hostbuild::OctreeHost();
      }
    }
  } drop{stash, !size_query};
  *n_inner = uint32_t(t.n_inner());
  *n_leaf = uint32_t(t.leaf_bits.size());
  if (num_layers) *num_layers = t.num_layers;
  if (root_aabb)
    for (int k = 0; k < 6; k++) root_aabb[k] = t.root_box[k];
  if (!inner_children || !inner_full || !leaf_bits || inner_capacity < *n_inner || leaf_capacity < *n_leaf)
    return fail(FCLB_ERR_CAPACITY, "fclb_octree_build_host: arrays too small (sizes returned)");
  std::copy(t.children.begin(), t.children.end(), inner_children);
  std::copy(t.full.begin(), t.full.end(), inner_full);
  std::copy(t.leaf_bits.begin(), t.leaf_bits.end(), leaf_bits);
  return FCLB_OK;
}

// the tree is built once on the host (mirror of Octree::rebuildTree) and uploaded to every device
int fclb_octree_build(const double* points, size_t n_points, double resolution, uint32_t bottom_half_shape, int scalar_type,
                      fclb_handle* octree) {
  int rc = ensureInit();
  if (rc) return rc;
  fclb::hostbuild::OctreeHost t;
  rc = octreeBuildHostImpl(points, n_points, resolution, bottom_half_shape, scalar_type, t,
                           "fclb_octree_build: bad argument (half shape: power of two >= 2)");
  if (rc) return rc;
  return forEachDevice([&] {
    return octree_upload_one(t.children.data(), t.full.data(), uint32_t(t.n_inner()), t.leaf_bits.data(),
                             uint32_t(t.leaf_bits.size()), nullptr, t.root_box, t.num_layers, octree);
  });
}

int fclb_octree_build_dev(const void* points, size_t n_points, double resolution, uint32_t bottom_half_shape, int scalar_type,
                          fclb_handle* octree) {
  return octreeBuildDevEntry(points, n_points, resolution, bottom_half_shape, scalar_type, octree);
}
static int octree_build_points_host_one(const void* points, size_t n_points, double resolution, uint32_t bottom_half_shape,
                                        int scalar_type, fclb_handle* octree) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n_points && !points) return fail(FCLB_ERR_BAD_ARG, "null points");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t bytes = n_points * 3 * (scalar_type == FCLB_F32 ? 4 : 8);
  DevBuf pts;
  FCLB_CUDA(pts.alloc(bytes));
  if (bytes) FCLB_CUDA(cudaMemcpyAsync(pts.p, points, bytes, cudaMemcpyHostToDevice, e.compute));
  return octreeBuildDevEntry(pts.p, n_points, resolution, bottom_half_shape, scalar_type, octree);
}
int fclb_octree_build_points_host(const void* points, size_t n_points, double resolution, uint32_t bottom_half_shape, int scalar_type,
                                  fclb_handle* octree) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return octree_build_points_host_one(points, n_points, resolution, bottom_half_shape, scalar_type, octree); });
}
int fclb_octree_info(fclb_handle octree, uint32_t* n_inner, uint32_t* n_leaf, int* num_layers, double* root_aabb) {
  auto it = octTable().find(octree);
  if (it == octTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_octree_info: unknown octree handle");
  if (n_inner) *n_inner = it->second->n_inner;
  if (n_leaf) *n_leaf = it->second->n_leaf;
  if (num_layers) *num_layers = it->second->num_layers;
  if (root_aabb)
    for (int k = 0; k < 6; k++) root_aabb[k] = it->second->root_box[k];
  return FCLB_OK;
}
int fclb_octree_export(fclb_handle octree, uint32_t* inner_children, uint8_t* inner_full, uint8_t* leaf_bits) {
  int rc = ensureInit();
  if (rc) return rc;
  auto it = octTable().find(octree);
  if (it == octTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_octree_export: unknown octree handle");
  const OctreeDev* d = it->second;
  if (inner_children) FCLB_CUDA(cudaMemcpy(inner_children, d->children, size_t(32) * d->n_inner, cudaMemcpyDeviceToHost));
  if (inner_full) FCLB_CUDA(cudaMemcpy(inner_full, d->inner_full, d->n_inner, cudaMemcpyDeviceToHost));
  if (leaf_bits && d->n_leaf) FCLB_CUDA(cudaMemcpy(leaf_bits, d->leaf_bits, d->n_leaf, cudaMemcpyDeviceToHost));
  return FCLB_OK;
}

int fclb_octree_prune_host(const uint32_t* inner_children, uint32_t n_inner, uint32_t n_leaf, const double* root_aabb,
                           int num_layers, const double* obb, int scalar_type, uint8_t* pruned, uint8_t* inner_full,
                           uint8_t* leaf_bits) {
  if (!inner_children || !n_inner || !root_aabb || num_layers < 3 || !obb || !pruned || !inner_full || (n_leaf && !leaf_bits))
    return fail(FCLB_ERR_BAD_ARG, "fclb_octree_prune_host: bad argument");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!fclb::hostbuild::octLinksInRange(inner_children, n_inner, n_leaf, num_layers, nullptr))
    return fail(FCLB_ERR_BAD_ARG, "fclb_octree_prune_host: child index out of range");
  const bool ok = scalar_type == FCLB_F32 ? fclb::hostbuild::octreePrune<float>(inner_children, n_inner, n_leaf, num_layers, root_aabb,
                                                                               obb, pruned, inner_full, leaf_bits)
                                          : fclb::hostbuild::octreePrune<double>(inner_children, n_inner, n_leaf, num_layers, root_aabb,
                                                                                obb, pruned, inner_full, leaf_bits);
  return ok ? FCLB_OK : fail(FCLB_ERR_BAD_ARG, "fclb_octree_prune_host: child index out of range");
}

int fclb_octree_consolidate_host(const uint32_t* inner_children, uint32_t n_inner, const uint8_t* pruned,
                                 const uint8_t* leaf_bits, uint32_t n_leaf, int num_layers, uint32_t* out_children,
                                 uint8_t* out_full, uint32_t* out_n_inner, uint8_t* out_leaf_bits, uint32_t* out_n_leaf) {
  if (!inner_children || !n_inner || !pruned || num_layers < 3 || (n_leaf && (!leaf_bits || !out_leaf_bits)) || !out_children ||
      !out_full || !out_n_inner || !out_n_leaf)
    return fail(FCLB_ERR_BAD_ARG, "fclb_octree_consolidate_host: bad argument");
  fclb::hostbuild::OctreeHost t;
  if (!fclb::hostbuild::octreeConsolidate(inner_children, n_inner, n_leaf, pruned, leaf_bits, num_layers, t))
    return fail(FCLB_ERR_BAD_ARG, "fclb_octree_consolidate_host: child index out of range");
  *out_n_inner = uint32_t(t.n_inner());
  *out_n_leaf = uint32_t(t.leaf_bits.size());
  std::copy(t.children.begin(), t.children.end(), out_children);
  std::copy(t.full.begin(), t.full.end(), out_full);
  std::copy(t.leaf_bits.begin(), t.leaf_bits.end(), out_leaf_bits);
  return FCLB_OK;
}

static int octree_release_one(fclb_handle h) {
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = octTable().find(h);
  if (it == octTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_octree_release: unknown handle");
  cudaFree(it->second->children);
  cudaFree(it->second->inner_full);
  cudaFree(it->second->leaf_bits);
  cudaFree(it->second->pruned);
  delete it->second;
  octTable().erase(it);
  return FCLB_OK;
}
int fclb_octree_release(fclb_handle h) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return octree_release_one(h); });
}

int fclb_octree_shape_collide_batch_dev(fclb_handle octree, fclb_handle shapes, const uint32_t* shape_ids,
                                        const void* poses_octree, const void* poses_shape, size_t n, int scalar_type,
                                        const fclb_request* req, uint32_t* out_counts, int64_t* out_first_node) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = octTable().find(octree);
  if (it == octTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown octree handle");
  ShapeTable* t = findTable(e, shapes);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown shape table handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!req || !out_counts) return fail(FCLB_ERR_BAD_ARG, "null request / out_counts");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_octree || !poses_shape) return fail(FCLB_ERR_BAD_ARG, "null input array");
  for (uint32_t i = 0; i < t->n; i++)
    if (t->host[i].type == FCLB_CONVEX) {
      const int nv = e.convex[t->host[i].geom].n_verts;
      if (nv <= 3 || nv == 6)
        return fail(FCLB_ERR_UNSUPPORTED, "octree-shape: Convex with 1, 2, 3 or 6 vertices uses the special OBB fitters "
                                          "(math/bv/utility-inl.h:63-131), which are not on the device");
    }
  if (req->penetration_mode != FCLB_PEN_DISABLED)
    return sceneShapeCountsViaContacts(e, FCLB_SCENE_OCTREE, octree, t, shape_ids, poses_octree, poses_shape, n, scalar_type, req,
                                       out_counts, nullptr, reinterpret_cast<long long*>(out_first_node));
  if (scalar_type == FCLB_F32)
    return octreeShapeDev<float>(e, it->second, t, shape_ids, poses_octree, poses_shape, n, req, out_counts,
                                 reinterpret_cast<long long*>(out_first_node));
  return octreeShapeDev<double>(e, it->second, t, shape_ids, poses_octree, poses_shape, n, req, out_counts,
                                reinterpret_cast<long long*>(out_first_node));
}

static int octree_shape_collide_batch_host_one(fclb_handle octree, fclb_handle shapes, const uint32_t* shape_ids,
                                         const void* poses_octree, const void* poses_shape, size_t n, int scalar_type,
                                         const fclb_request* req, uint32_t* out_counts, int64_t* out_first_node) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_octree || !poses_shape || !out_counts) return fail(FCLB_ERR_BAD_ARG, "null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  {
    ShapeTable* t = findTable(e, shapes);
    if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown shape table handle");
    for (size_t q = 0; q < n; q++)
      if (shape_ids[q] >= t->n) return fail(FCLB_ERR_BAD_ARG, "shape id out of range");
  }
  return sceneShapeHostStaged(e, shape_ids, poses_octree, poses_shape, n, scalar_type, out_counts, out_first_node, 8,
                              [&](const uint32_t* d_ids, const void* d_p1, const void* d_p2, size_t m, uint32_t* d_cnt, void* d_first) {
                                return fclb_octree_shape_collide_batch_dev(octree, shapes, d_ids, d_p1, d_p2, m, scalar_type, req, d_cnt,
                                          static_cast<int64_t*>(d_first));
                              });
}
int fclb_octree_shape_collide_batch_host(fclb_handle octree, fclb_handle shapes, const uint32_t* shape_ids,
                                         const void* poses_octree, const void* poses_shape, size_t n, int scalar_type,
                                         const fclb_request* req, uint32_t* out_counts, int64_t* out_first_node) {
  if (engineCount() <= 1) return octree_shape_collide_batch_host_one(octree, shapes, shape_ids, poses_octree, poses_shape, n, scalar_type, req, out_counts, out_first_node);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return octree_shape_collide_batch_host_one(octree, shapes, offT(shape_ids, b), offPtr(poses_octree, b * 12 * ss), offPtr(poses_shape, b * 12 * ss), m_, scalar_type, req, offT(out_counts, b), offT(out_first_node, b)); });
}

int fclb_scene_shape_contacts_batch_dev(int scene_kind, fclb_handle scene, fclb_handle shapes, const uint32_t* shape_ids,
                                        const void* poses_scene, const void* poses_shape, size_t n, int scalar_type,
                                        const fclb_request* req, uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1,
                                        void* out_contacts) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  ShapeTable* t = findTable(e, shapes);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown shape table handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!req || !out_counts || !out_b1 || !out_contacts || max_keep == 0) return fail(FCLB_ERR_BAD_ARG, "null output / max_keep == 0");
  if (req->penetration_mode > FCLB_PEN_INCREMENTAL_MIN)
    return fail(FCLB_ERR_BAD_ARG, "fclb_scene_shape_contacts_batch: unknown penetration mode");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_scene || !poses_shape) return fail(FCLB_ERR_BAD_ARG, "null input array");
  return sceneShapeContactsAny(e, scene_kind, scene, t, shape_ids, poses_scene, poses_shape, n, scalar_type, req, max_keep,
                               out_counts, reinterpret_cast<long long*>(out_b1), out_contacts);
}

static int scene_shape_contacts_batch_host_one(int scene_kind, fclb_handle scene, fclb_handle shapes, const uint32_t* shape_ids,
                                         const void* poses_scene, const void* poses_shape, size_t n, int scalar_type,
                                         const fclb_request* req, uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1,
                                         void* out_contacts) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_scene || !poses_shape || !out_counts || !out_b1 || !out_contacts || max_keep == 0)
    return fail(FCLB_ERR_BAD_ARG, "null array / max_keep == 0");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  {
    ShapeTable* t = findTable(e, shapes);
    if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown shape table handle");
    for (size_t q = 0; q < n; q++)
      if (shape_ids[q] >= t->n) return fail(FCLB_ERR_BAD_ARG, "shape id out of range");
  }
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_ids = 0;
  const size_t o_p1 = alignUp(o_ids + n * 4, 256);
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_b1 = alignUp(o_cnt + n * 4, 256);
  const size_t o_ct = alignUp(o_b1 + n * size_t(max_keep) * 8, 256);
  const size_t total = alignUp(o_ct + n * size_t(max_keep) * 7 * ss, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_ids, shape_ids, n * 4, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses_scene, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses_shape, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemsetAsync(base + o_b1, 0xff, n * size_t(max_keep) * 8, e.compute));
  rc = fclb_scene_shape_contacts_batch_dev(scene_kind, scene, shapes, reinterpret_cast<const uint32_t*>(base + o_ids),
                                           base + o_p1, base + o_p2, n, scalar_type, req, max_keep,
                                           reinterpret_cast<uint32_t*>(base + o_cnt), reinterpret_cast<int64_t*>(base + o_b1),
                                           base + o_ct);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_counts, base + o_cnt, n * 4, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(out_b1, base + o_b1, n * size_t(max_keep) * 8, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(out_contacts, base + o_ct, n * size_t(max_keep) * 7 * ss, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}
int fclb_scene_shape_contacts_batch_host(int scene_kind, fclb_handle scene, fclb_handle shapes, const uint32_t* shape_ids,
                                         const void* poses_scene, const void* poses_shape, size_t n, int scalar_type,
                                         const fclb_request* req, uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1,
                                         void* out_contacts) {
  if (engineCount() <= 1) return scene_shape_contacts_batch_host_one(scene_kind, scene, shapes, shape_ids, poses_scene, poses_shape, n, scalar_type, req, max_keep, out_counts, out_b1, out_contacts);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return scene_shape_contacts_batch_host_one(scene_kind, scene, shapes, offT(shape_ids, b), offPtr(poses_scene, b * 12 * ss), offPtr(poses_shape, b * 12 * ss), m_, scalar_type, req, max_keep, offT(out_counts, b), offT(out_b1, b * size_t(max_keep)), offPtr(out_contacts, b * size_t(max_keep) * 7 * ss)); });
}

int fclb_scene_pair_collide_batch_dev(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                      const void* poses2, size_t n, int scalar_type, const fclb_request* req,
                                      uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1, int64_t* out_b2) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!req || !out_counts) return fail(FCLB_ERR_BAD_ARG, "null request / out_counts");
  if ((out_b1 == nullptr) != (out_b2 == nullptr) || (out_b1 && max_keep == 0))
    return fail(FCLB_ERR_BAD_ARG, "out_b1 / out_b2 go together and need max_keep > 0");
  if (n == 0) return FCLB_OK;
  if (!poses1 || !poses2) return fail(FCLB_ERR_BAD_ARG, "null pose array");
  if (req->penetration_mode != FCLB_PEN_DISABLED) {  // contacts are generated (and dropped): counts / ids of the contact path
    const uint32_t keep = out_b1 ? max_keep : 1;
    DevBuf tb1, tb2, tc;
    FCLB_CUDA(tb1.alloc(n * size_t(keep) * 8));
    FCLB_CUDA(tb2.alloc(n * size_t(keep) * 8));
    FCLB_CUDA(tc.alloc(n * size_t(keep) * 7 * 8));
    return scenePairContactsAny(e, kind1, scene1, kind2, scene2, poses1, poses2, n, scalar_type, req, keep, out_counts,
                                out_b1 ? reinterpret_cast<long long*>(out_b1) : tb1.as<long long>(),
                                out_b2 ? reinterpret_cast<long long*>(out_b2) : tb2.as<long long>(), tc.p);
  }
  if (scalar_type == FCLB_F32)
    return scenePairDev<float>(e, kind1, scene1, kind2, scene2, poses1, poses2, n, req, max_keep, out_counts,
                               reinterpret_cast<long long*>(out_b1), reinterpret_cast<long long*>(out_b2));
  return scenePairDev<double>(e, kind1, scene1, kind2, scene2, poses1, poses2, n, req, max_keep, out_counts,
                              reinterpret_cast<long long*>(out_b1), reinterpret_cast<long long*>(out_b2));
}

static int scene_pair_collide_batch_host_one(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                       const void* poses2, size_t n, int scalar_type, const fclb_request* req,
                                       uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1, int64_t* out_b2) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!poses1 || !poses2 || !out_counts) return fail(FCLB_ERR_BAD_ARG, "null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t keep = out_b1 ? size_t(max_keep) : 0;
  const size_t o_p1 = 0;
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_b1 = alignUp(o_cnt + n * 4, 256);
  const size_t o_b2 = alignUp(o_b1 + n * keep * 8, 256);
  const size_t total = alignUp(o_b2 + n * keep * 8, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses1, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses2, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  if (keep) FCLB_CUDA(cudaMemsetAsync(base + o_b1, 0xff, o_b2 + n * keep * 8 - o_b1, e.compute));
  rc = fclb_scene_pair_collide_batch_dev(kind1, scene1, kind2, scene2, base + o_p1, base + o_p2, n, scalar_type, req, max_keep,
                                         reinterpret_cast<uint32_t*>(base + o_cnt),
                                         keep ? reinterpret_cast<int64_t*>(base + o_b1) : nullptr,
                                         keep ? reinterpret_cast<int64_t*>(base + o_b2) : nullptr);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_counts, base + o_cnt, n * 4, cudaMemcpyDeviceToHost, e.compute));
  if (keep) {
    FCLB_CUDA(cudaMemcpyAsync(out_b1, base + o_b1, n * keep * 8, cudaMemcpyDeviceToHost, e.compute));
    FCLB_CUDA(cudaMemcpyAsync(out_b2, base + o_b2, n * keep * 8, cudaMemcpyDeviceToHost, e.compute));
  }
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}
int fclb_scene_pair_collide_batch_host(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                       const void* poses2, size_t n, int scalar_type, const fclb_request* req,
                                       uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1, int64_t* out_b2) {
  if (engineCount() <= 1) return scene_pair_collide_batch_host_one(kind1, scene1, kind2, scene2, poses1, poses2, n, scalar_type, req, max_keep, out_counts, out_b1, out_b2);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return scene_pair_collide_batch_host_one(kind1, scene1, kind2, scene2, offPtr(poses1, b * 12 * ss), offPtr(poses2, b * 12 * ss), m_, scalar_type, req, max_keep, offT(out_counts, b), offT(out_b1, b * size_t(max_keep)), offT(out_b2, b * size_t(max_keep))); });
}

int fclb_scene_pair_contacts_batch_dev(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                       const void* poses2, size_t n, int scalar_type, const fclb_request* req,
                                       uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1, int64_t* out_b2,
                                       void* out_contacts) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!req || !out_counts || !out_b1 || !out_b2 || !out_contacts || max_keep == 0)
    return fail(FCLB_ERR_BAD_ARG, "null output / max_keep == 0");
  if (req->penetration_mode == FCLB_PEN_DISABLED || req->penetration_mode > FCLB_PEN_INCREMENTAL_MIN)
    return fail(FCLB_ERR_BAD_ARG, "fclb_scene_pair_contacts_batch: the request must enable a penetration mode");
  if (n == 0) return FCLB_OK;
  if (!poses1 || !poses2) return fail(FCLB_ERR_BAD_ARG, "null pose array");
  return scenePairContactsAny(e, kind1, scene1, kind2, scene2, poses1, poses2, n, scalar_type, req, max_keep, out_counts,
                              reinterpret_cast<long long*>(out_b1), reinterpret_cast<long long*>(out_b2), out_contacts);
}

static int scene_pair_contacts_batch_host_one(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                        const void* poses2, size_t n, int scalar_type, const fclb_request* req,
                                        uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1, int64_t* out_b2,
                                        void* out_contacts) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!poses1 || !poses2 || !out_counts || !out_b1 || !out_b2 || !out_contacts || max_keep == 0)
    return fail(FCLB_ERR_BAD_ARG, "null array / max_keep == 0");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t keep = max_keep;
  const size_t o_p1 = 0;
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_b1 = alignUp(o_cnt + n * 4, 256);
  const size_t o_b2 = alignUp(o_b1 + n * keep * 8, 256);
  const size_t o_ct = alignUp(o_b2 + n * keep * 8, 256);
  const size_t total = alignUp(o_ct + n * keep * 7 * ss, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses1, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses2, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemsetAsync(base + o_b1, 0xff, o_b2 + n * keep * 8 - o_b1, e.compute));
  rc = fclb_scene_pair_contacts_batch_dev(kind1, scene1, kind2, scene2, base + o_p1, base + o_p2, n, scalar_type, req, max_keep,
                                          reinterpret_cast<uint32_t*>(base + o_cnt), reinterpret_cast<int64_t*>(base + o_b1),
                                          reinterpret_cast<int64_t*>(base + o_b2), base + o_ct);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_counts, base + o_cnt, n * 4, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(out_b1, base + o_b1, n * keep * 8, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(out_b2, base + o_b2, n * keep * 8, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(out_contacts, base + o_ct, n * keep * 7 * ss, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}
int fclb_scene_pair_contacts_batch_host(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                        const void* poses2, size_t n, int scalar_type, const fclb_request* req,
                                        uint32_t max_keep, uint32_t* out_counts, int64_t* out_b1, int64_t* out_b2,
                                        void* out_contacts) {
  if (engineCount() <= 1) return scene_pair_contacts_batch_host_one(kind1, scene1, kind2, scene2, poses1, poses2, n, scalar_type, req, max_keep, out_counts, out_b1, out_b2, out_contacts);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return scene_pair_contacts_batch_host_one(kind1, scene1, kind2, scene2, offPtr(poses1, b * 12 * ss), offPtr(poses2, b * 12 * ss), m_, scalar_type, req, max_keep, offT(out_counts, b), offT(out_b1, b * size_t(max_keep)), offT(out_b2, b * size_t(max_keep)), offPtr(out_contacts, b * size_t(max_keep) * 7 * ss)); });
}

/* node / leaf tests executed by the most recent mesh-shape or heightmap-shape batch call */
int fclb_scene_last_visit_counts(uint64_t* n_bv, uint64_t* n_leaf) {
  if (n_bv) *n_bv = g_stats[0];
  if (n_leaf) *n_leaf = g_stats[1];
  return FCLB_OK;
}

}  // extern "C"
