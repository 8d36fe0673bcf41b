// fclb_scene_api.cu -- C ABI entry points for shape-vs-scene queries:
//   fclb_bvh_shape_collide_batch_{dev,host}   mesh (BVHModel<OBBRSS>) vs convex shape
// Kernels: fclb_bvh_shape_impl.cuh (instantiated in fclb_bvh_shape_f32/f64.cu).
#include "fclb_bvh.cuh"
#include "fclb_shapes.cuh"

namespace fclb {

static unsigned long long* g_counters = nullptr;  // [0] work counter, [1..2] stats
static unsigned long long g_stats[2] = {0, 0};

static int tableUniformType(const ShapeTable* t) {
  int type = int(t->host[0].type);
  for (uint32_t i = 1; i < t->n; i++)
    if (int(t->host[i].type) != type) return ST_DYNAMIC;
  return type;
}

template <typename S>
static int bvhShapeDev(Engine& e, const BvhDev* m, const ShapeTable* t, const uint32_t* shape_ids, const void* poses_mesh,
                       const void* poses_shape, size_t n, const fclb_request* req, uint32_t* counts, int32_t* first_tri) {
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
  FCLB_CUDA(cudaMemcpyAsync(g_stats, g_counters + 1, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
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

}  // namespace fclb

using namespace fclb;

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
  if (req->penetration_mode != FCLB_PEN_DISABLED)
    return fail(FCLB_ERR_UNSUPPORTED, "mesh-shape contact generation (penetration modes) is not on the device yet");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_mesh || !poses_shape) return fail(FCLB_ERR_BAD_ARG, "null input array");
  if (scalar_type == FCLB_F32)
    return bvhShapeDev<float>(e, it->second, t, shape_ids, poses_mesh, poses_shape, n, req, out_counts, out_first_tri);
  return bvhShapeDev<double>(e, it->second, t, shape_ids, poses_mesh, poses_shape, n, req, out_counts, out_first_tri);
}

int fclb_bvh_shape_collide_batch_host(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids, const void* poses_mesh,
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
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_ids = 0;
  const size_t o_p1 = alignUp(o_ids + n * 4, 256);
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_ft = alignUp(o_cnt + n * 4, 256);
  const size_t total = alignUp(o_ft + n * 4, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_ids, shape_ids, n * 4, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses_mesh, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses_shape, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  rc = fclb_bvh_shape_collide_batch_dev(bvh, shapes, reinterpret_cast<const uint32_t*>(base + o_ids), base + o_p1,
                                        base + o_p2, n, scalar_type, req, reinterpret_cast<uint32_t*>(base + o_cnt),
                                        out_first_tri ? reinterpret_cast<int32_t*>(base + o_ft) : nullptr);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_counts, base + o_cnt, n * 4, cudaMemcpyDeviceToHost, e.compute));
  if (out_first_tri) FCLB_CUDA(cudaMemcpyAsync(out_first_tri, base + o_ft, n * 4, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}

/* node / leaf tests executed by the most recent mesh-shape or heightmap-shape batch call */
int fclb_scene_last_visit_counts(uint64_t* n_bv, uint64_t* n_leaf) {
  if (n_bv) *n_bv = g_stats[0];
  if (n_leaf) *n_leaf = g_stats[1];
  return FCLB_OK;
}

}  // extern "C"
