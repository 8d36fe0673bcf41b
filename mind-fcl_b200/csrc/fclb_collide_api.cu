// fclb_collide_api.cu -- C ABI entry points for batched fcl::collide
// (shape-shape) and the direct GJK+EPA path.  Kernels: fclb_collide_impl.cuh.
#include "fclb_collide_impl.cuh"
#include "fclb_engine.h"

namespace fclb {

struct CollideWorkspace {
  uint32_t* count = nullptr;
  uint32_t* query = nullptr;
  void* simplex = nullptr;
  int32_t* rank = nullptr;
  uint32_t* defer_count = nullptr;
  uint32_t* defer_item = nullptr;
  size_t cap = 0;      // queries
  size_t scalar = 0;   // bytes per scalar the simplex buffer was sized for
};
static PerDevice<CollideWorkspace> g_ws_pd;
#define g_ws (g_ws_pd.get())

static int ensureWorkspace(size_t n, size_t ss) {
  if (n <= g_ws.cap && ss <= g_ws.scalar) return FCLB_OK;
  if (g_ws.count) cudaFree(g_ws.count);
  if (g_ws.query) cudaFree(g_ws.query);
  if (g_ws.simplex) cudaFree(g_ws.simplex);
  if (g_ws.rank) cudaFree(g_ws.rank);
  if (g_ws.defer_count) cudaFree(g_ws.defer_count);
  if (g_ws.defer_item) cudaFree(g_ws.defer_item);
  g_ws = CollideWorkspace();
  const size_t cap = n > g_ws.cap ? n : g_ws.cap;
  FCLB_CUDA(cudaMalloc(&g_ws.count, sizeof(uint32_t)));
  FCLB_CUDA(cudaMalloc(&g_ws.query, cap * sizeof(uint32_t)));
  FCLB_CUDA(cudaMalloc(&g_ws.simplex, cap * 24 * 8));
  FCLB_CUDA(cudaMalloc(&g_ws.rank, cap * sizeof(int32_t)));
  FCLB_CUDA(cudaMalloc(&g_ws.defer_count, 8 * sizeof(uint32_t)));  // EpaDefer words: count, cursors, done flag
  FCLB_CUDA(cudaMalloc(&g_ws.defer_item, cap * sizeof(uint32_t)));
  g_ws.cap = cap;
  g_ws.scalar = 8;
  return FCLB_OK;
}

template <typename S>
static int collideDev(Engine& e, ShapeTable* t, const fclb_pair* pairs, const void* poses1, const void* poses2, size_t n,
                      const CollideLaunchArgs& base) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  uint32_t counts[kNumKinds], offsets[kNumKinds];
  int uniform = -1;
  FCLB_CUDA(cudaEventRecord(e.ev_call0, e.compute));
  int rc = bucketBatch<S>(e, t, pairs, n, counts, offsets, &uniform);
  if (rc) return rc;
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  int launches = 0;
  e.n_rec = 0;
  FCLB_CUDA(cudaEventRecord(e.rec_ev[0], e.compute));
  for (int k = 0; k < kNumKinds; k++) {
    if (!counts[k]) continue;
    BatchView b{};
    b.shapes = t->d_shapes[st];
    b.convex = e.d_convex_tab[st];
    b.pairs = pairs;
    b.poses1 = poses1;
    b.poses2 = poses2;
    b.tris = base.tris;
    b.perm = (uniform >= 0) ? nullptr : e.d_perm;
    b.begin = (uniform >= 0) ? 0 : offsets[k];
    b.count = counts[k];
    b.type1 = k / kNumTypes;
    b.type2 = k % kNumTypes;
    CollideLaunchArgs a = base;
    if (a.mode & 2) FCLB_CUDA(cudaMemsetAsync(g_ws.count, 0, sizeof(uint32_t), e.compute));
    FCLB_CUDA(launchCollide<S>(b, a, e.compute, &launches));
    if (a.pen_mode) {
      FCLB_CUDA(launchMprPenetration<S>(b, a, e.compute));
      launches += 1;
    }
    e.rec_kind[e.n_rec] = k;
    e.rec_count[e.n_rec] = counts[k];
    e.n_rec++;
    FCLB_CUDA(cudaEventRecord(e.rec_ev[e.n_rec], e.compute));
  }
  e.launches += uint64_t(launches);
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = ms;
  cudaEventElapsedTime(&ms, e.ev_call0, e.ev1);
  e.last_call_ms = ms;
  for (int i = 0; i < e.n_rec; i++) cudaEventElapsedTime(&e.rec_ms[i], e.rec_ev[i], e.rec_ev[i + 1]);
  return FCLB_OK;
}

static int runCollideTable(Engine& e, ShapeTable* t, const void* tris, const fclb_pair* pairs, const void* poses1,
                           const void* poses2, size_t n, int scalar_type, const fclb_request* req, int api_mode,
                           CollideOut out);

static int runCollide(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2, size_t n,
                      int scalar_type, const fclb_request* req, int api_mode, CollideOut out) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  ShapeTable* t = findTable(e, shapes);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown shape table handle");
  return runCollideTable(e, t, nullptr, pairs, poses1, poses2, n, scalar_type, req, api_mode, out);
}

// Leaf batch of the scene contact path (fclb_scene_gjk.cu): the reference's leaf routine ShapeIntersect /
// ShapeSimplexIntersect with contacts == fcl::collide on (leaf geometry, shape) pairs.  d_table: a device ShapeD<S>
// table that holds the caller's shapes followed by one Box / Triangle entry per leaf; contacts: 4 x 9 S per item.
int collideLeafBatch(Engine& e, void* d_table, const void* tris, const fclb_pair* pairs, const void* poses1,
                     const void* poses2, size_t m, int scalar_type, const fclb_request* req, void* contacts,
                     uint32_t* counts, uint32_t n_table) {
  ShapeTable tmp;
  tmp.n = n_table;
  tmp.d_shapes[scalar_type == FCLB_F32 ? 0 : 1] = d_table;
  fclb_request r = *req;
  r.max_contacts = 4;  // every contact of the leaf (boxBox2 makes up to four); the scatter pass clips to the free space
  r.penetration_mode = FCLB_PEN_DEFAULT_GJK_EPA;
  CollideOut out{};
  out.contacts = contacts;
  out.counts = counts;
  out.max_keep = 4;
  return runCollideTable(e, &tmp, tris, pairs, poses1, poses2, m, scalar_type, &r, 0, out);
}

static int runCollideTable(Engine& e, ShapeTable* t, const void* tris, const fclb_pair* pairs, const void* poses1,
                           const void* poses2, size_t n, int scalar_type, const fclb_request* req, int api_mode,
                           CollideOut out) {
  int rc = FCLB_OK;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!req) return fail(FCLB_ERR_BAD_ARG, "null request");
  if (n == 0) return FCLB_OK;
  if (n > 0xffffffffull) return fail(FCLB_ERR_CAPACITY, "batch larger than 2^32-1 queries: split it");
  if (!pairs || !poses1 || !poses2) return fail(FCLB_ERR_BAD_ARG, "null input array");
  const bool mpr_pen = req->penetration_mode == FCLB_PEN_DIRECTED || req->penetration_mode == FCLB_PEN_INCREMENTAL_MIN;
  if (mpr_pen && api_mode == 1) return fail(FCLB_ERR_BAD_ARG, "the MPR penetration modes belong to fclb_collide_batch");
  if (req->penetration_mode > FCLB_PEN_INCREMENTAL_MIN) return fail(FCLB_ERR_BAD_ARG, "unknown penetration mode");
  if (req->epa_max_faces > 1024) return fail(FCLB_ERR_CAPACITY, "epa_max_faces > 1024 does not fit shared memory");
  CollideLaunchArgs a{};
  a.sp = solverParams(scalar_type, req->binary_tol, req->gjk_max_iter, req->distance_tol, req->epa_max_faces,
                      req->epa_max_iter, true);
  a.tris = tris;
  a.out = out;
  a.out.max_contacts = req->max_contacts;
  a.out.penetration = (req->penetration_mode == FCLB_PEN_DEFAULT_GJK_EPA) ? 1 : 0;
  if (api_mode == 1) {
    a.mode = 4 | 2;  // direct GJK then EPA
  } else {
    a.mode = a.out.penetration ? 2 : 1;  // contacts: GJK+EPA; none: MPR with GJK fallback
  }
  if (a.mode & 2) {
    rc = ensureWorkspace(n, scalar_type == FCLB_F32 ? 4 : 8);
    if (rc) return rc;
  }
  a.work.count = g_ws.count;
  a.work.query = g_ws.query;
  a.work.simplex = g_ws.simplex;
  a.work.rank = g_ws.rank;
  a.work.capacity = uint32_t(g_ws.cap);
  a.defer.count = g_ws.defer_count;
  a.aux = e.aux;
  a.ev_aux0 = e.ev_aux0;
  a.ev_aux1 = e.ev_aux1;
  a.item_capacity = g_ws.cap;
  a.defer.item = g_ws.defer_item;
  a.defer.enabled = 0;
  a.defer.consume = 0;
  if (mpr_pen) {  // boolean collide first (collision_penetration-inl.h:246-250), then one MPR contact per hit
    a.pen_mode = int(req->penetration_mode);
    for (int k = 0; k < 3; k++) a.pen_dir[k] = req->dir[k];
  }
  if (scalar_type == FCLB_F32) return collideDev<float>(e, t, pairs, poses1, poses2, n, a);
  return collideDev<double>(e, t, pairs, poses1, poses2, n, a);
}

// host-buffer wrapper: stage inputs, run, copy the requested outputs back
static int runCollideHost(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2, size_t n,
                          int scalar_type, const fclb_request* req, int api_mode, uint32_t max_keep, void* h_contacts,
                          uint32_t* h_counts, int32_t* h_gjk, int32_t* h_epa, void* h_geom) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!pairs || !poses1 || !poses2) return fail(FCLB_ERR_BAD_ARG, "null input array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_pairs = 0;
  const size_t o_p1 = alignUp(o_pairs + n * sizeof(fclb_pair), 256);
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_cont = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t cont_bytes = h_contacts ? n * size_t(max_keep) * 9 * ss : 0;
  const size_t o_cnt = alignUp(o_cont + cont_bytes, 256);
  const size_t o_gjk = alignUp(o_cnt + n * 4, 256);
  const size_t o_epa = alignUp(o_gjk + n * 4, 256);
  const size_t o_geom = alignUp(o_epa + n * 4, 256);
  const size_t total = alignUp(o_geom + n * 7 * ss, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  // Chunked three-stage pipeline, as in fclb_distance_batch_host: every H2D copy is queued up front on the copy-in stream
  // (one event per chunk), the compute stream waits per chunk, and the copy-out stream drains a chunk's contacts while later
  // chunks upload / compute.  A contact record batch is larger than its inputs (4 x 9 S against 2 x 12 S + 8 B per query),
  // so the call is bounded by the copy-out direction instead of the sum of the two.
  // Stages pay only where the copies dominate: boolean requests (one MPR / closed-form test per query) and contact requests
  // on box / sphere tables (closed forms).  GJK + EPA batches are compute-bound and their tiered launches want the whole
  // batch at once -- measured on B200, 1M queries: c1a 2.05e8 -> 2.41e8 q/s with four stages, c1b 9.6e7 -> 6.7e7
  // (profiles/r02_collide_host_pipeline.txt) -- so those keep one stage.
  bool copy_bound = req && req->penetration_mode == FCLB_PEN_DISABLED && api_mode == 0;
  if (!copy_bound && api_mode == 0) {
    if (const ShapeTable* t = findTable(e, shapes)) {
      copy_bound = true;
      for (uint32_t i = 0; i < t->n; i++)
        if (t->host[i].type != FCLB_BOX && t->host[i].type != FCLB_SPHERE) copy_bound = false;
    }
  }
  // (copy-OUT bound: the result copies can only start after the first stage's upload + kernels, so that stage is short)
  const size_t chunk = copy_bound ? std::max<size_t>(e.host_chunk / 8, 1024) : n;
  std::vector<size_t> c_begin, c_size;
  stageSizes(n, chunk, copy_bound ? e.host_head : 0, 0, c_begin, c_size);
  const int n_chunks = int(c_size.size());
  rc = ensureChunkEvents(e, n_chunks);
  if (rc) return rc;
  const char* hp_pairs = reinterpret_cast<const char*>(pairs);
  const char* hp_1 = static_cast<const char*>(poses1);
  const char* hp_2 = static_cast<const char*>(poses2);
  const size_t cb = size_t(max_keep) * 9 * ss;  // contact bytes per query
  FCLB_CUDA(cudaStreamSynchronize(e.copy_out));  // the arena may still be read by a previous call's copy-out
  for (int c = 0; c < n_chunks; c++) {
    const size_t b0 = c_begin[c], m = c_size[c];
    FCLB_CUDA(cudaMemcpyAsync(base + o_pairs + b0 * sizeof(fclb_pair), hp_pairs + b0 * sizeof(fclb_pair), m * sizeof(fclb_pair),
                              cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaMemcpyAsync(base + o_p1 + b0 * 12 * ss, hp_1 + b0 * 12 * ss, m * 12 * ss, cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaMemcpyAsync(base + o_p2 + b0 * 12 * ss, hp_2 + b0 * 12 * ss, m * 12 * ss, cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaEventRecord(e.ev_in[c], e.copy_in));
  }
  for (int c = 0; c < n_chunks; c++) {
    const size_t b0 = c_begin[c], m = c_size[c];
    FCLB_CUDA(cudaStreamWaitEvent(e.compute, e.ev_in[c], 0));
    if (h_contacts) FCLB_CUDA(cudaMemsetAsync(base + o_cont + b0 * cb, 0, m * cb, e.compute));  // slots beyond counts[q] come back as zeros
    CollideOut out{};
    out.contacts = h_contacts ? base + o_cont + b0 * cb : nullptr;
    out.counts = reinterpret_cast<uint32_t*>(base + o_cnt) + b0;
    out.max_keep = h_contacts ? max_keep : 0;
    out.gjk_status = reinterpret_cast<int32_t*>(base + o_gjk) + b0;
    out.epa_status = reinterpret_cast<int32_t*>(base + o_epa) + b0;
    out.geom = base + o_geom + b0 * 7 * ss;
    rc = runCollide(shapes, reinterpret_cast<const fclb_pair*>(base + o_pairs) + b0, base + o_p1 + b0 * 12 * ss,
                    base + o_p2 + b0 * 12 * ss, m, scalar_type, req, api_mode, out);
    if (rc) {  // drain the queued copies before the caller gets its buffers back
      cudaStreamSynchronize(e.copy_in);
      cudaStreamSynchronize(e.compute);
      cudaStreamSynchronize(e.copy_out);
      return rc;
    }
    FCLB_CUDA(cudaEventRecord(e.ev_done[c], e.compute));
    FCLB_CUDA(cudaStreamWaitEvent(e.copy_out, e.ev_done[c], 0));
    if (h_contacts)
      FCLB_CUDA(cudaMemcpyAsync(static_cast<char*>(h_contacts) + b0 * cb, base + o_cont + b0 * cb, m * cb, cudaMemcpyDeviceToHost, e.copy_out));
    if (h_counts) FCLB_CUDA(cudaMemcpyAsync(h_counts + b0, base + o_cnt + b0 * 4, m * 4, cudaMemcpyDeviceToHost, e.copy_out));
    if (h_gjk) FCLB_CUDA(cudaMemcpyAsync(h_gjk + b0, base + o_gjk + b0 * 4, m * 4, cudaMemcpyDeviceToHost, e.copy_out));
    if (h_epa) FCLB_CUDA(cudaMemcpyAsync(h_epa + b0, base + o_epa + b0 * 4, m * 4, cudaMemcpyDeviceToHost, e.copy_out));
    if (h_geom)
      FCLB_CUDA(cudaMemcpyAsync(static_cast<char*>(h_geom) + b0 * 7 * ss, base + o_geom + b0 * 7 * ss, m * 7 * ss, cudaMemcpyDeviceToHost,
                                e.copy_out));
  }
  FCLB_CUDA(cudaStreamSynchronize(e.copy_out));
  return FCLB_OK;
}

}  // namespace fclb

using namespace fclb;

extern "C" {

int fclb_collide_batch_dev(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                           size_t n, int scalar_type, const fclb_request* req, uint32_t max_keep, void* out_contacts,
                           uint32_t* out_counts) {
  if (!out_counts) return fail(FCLB_ERR_BAD_ARG, "fclb_collide_batch: out_counts is required");
  CollideOut out{};
  out.contacts = out_contacts;
  out.counts = out_counts;
  out.max_keep = out_contacts ? max_keep : 0;
  return runCollide(shapes, pairs, poses1, poses2, n, scalar_type, req, 0, out);
}

static int collide_batch_host_one(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                            size_t n, int scalar_type, const fclb_request* req, uint32_t max_keep, void* out_contacts,
                            uint32_t* out_counts) {
  if (!out_counts) return fail(FCLB_ERR_BAD_ARG, "fclb_collide_batch: out_counts is required");
  return runCollideHost(shapes, pairs, poses1, poses2, n, scalar_type, req, 0, max_keep, out_contacts, out_counts,
                        nullptr, nullptr, nullptr);
}
int fclb_collide_batch_host(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                            size_t n, int scalar_type, const fclb_request* req, uint32_t max_keep, void* out_contacts,
                            uint32_t* out_counts) {
  if (engineCount() <= 1) return collide_batch_host_one(shapes, pairs, poses1, poses2, n, scalar_type, req, max_keep, out_contacts, out_counts);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return collide_batch_host_one(shapes, offT(pairs, b), offPtr(poses1, b * 12 * ss), offPtr(poses2, b * 12 * ss), m_, scalar_type, req, max_keep, offPtr(out_contacts, b * size_t(max_keep) * 9 * ss), offT(out_counts, b)); });
}

int fclb_gjk_epa_batch_dev(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                           size_t n, int scalar_type, const fclb_request* req, int32_t* out_gjk, int32_t* out_epa,
                           void* out_geom) {
  if (!out_gjk) return fail(FCLB_ERR_BAD_ARG, "fclb_gjk_epa_batch: out_gjk is required");
  CollideOut out{};
  out.gjk_status = out_gjk;
  out.epa_status = out_epa;
  out.geom = out_geom;
  return runCollide(shapes, pairs, poses1, poses2, n, scalar_type, req, 1, out);
}

static int gjk_epa_batch_host_one(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                            size_t n, int scalar_type, const fclb_request* req, int32_t* out_gjk, int32_t* out_epa,
                            void* out_geom) {
  if (!out_gjk) return fail(FCLB_ERR_BAD_ARG, "fclb_gjk_epa_batch: out_gjk is required");
  return runCollideHost(shapes, pairs, poses1, poses2, n, scalar_type, req, 1, 0, nullptr, nullptr, out_gjk, out_epa,
                        out_geom);
}
int fclb_gjk_epa_batch_host(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                            size_t n, int scalar_type, const fclb_request* req, int32_t* out_gjk, int32_t* out_epa,
                            void* out_geom) {
  if (engineCount() <= 1) return gjk_epa_batch_host_one(shapes, pairs, poses1, poses2, n, scalar_type, req, out_gjk, out_epa, out_geom);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return gjk_epa_batch_host_one(shapes, offT(pairs, b), offPtr(poses1, b * 12 * ss), offPtr(poses2, b * 12 * ss), m_, scalar_type, req, offT(out_gjk, b), offT(out_epa, b), offPtr(out_geom, b * 7 * ss)); });
}

}  // extern "C"
