// fclb_engine.cu -- engine state, geometry upload, query bucketing and the C ABI
// declared in include/fclb200.h.
//
// One process binds one GPU (fclb_init).  A batch call
//   1. (host entry points) stages the batch arrays H2D in chunks on a copy
//      stream while earlier chunks compute and drain D2H;
//   2. buckets the queries by (type1,type2) on the device (histogram + scatter),
//      so every kernel launch sees one pair kind;
//   3. launches the per-kind kernels (fclb_distance_*.cu, fclb_collide_*.cu).
// There is no CPU fallback: with no CUDA device every compute entry point
// returns FCLB_ERR_NO_DEVICE.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "fclb_bound.h"
#include "fclb_engine.h"
#include "fclb_shapes.cuh"

namespace fclb {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

static Engine* g_engines[kMaxDevices] = {};
static std::atomic<int> g_n_engines{0};
static std::mutex g_engines_mu;
static thread_local int t_slot = 0;

Engine& eng() {
  Engine* e = g_engines[t_slot];
  if (!e) {
    std::lock_guard<std::mutex> lk(g_engines_mu);
    if (!g_engines[t_slot]) {
      g_engines[t_slot] = new Engine();
      g_engines[t_slot]->slot = t_slot;
    }
    e = g_engines[t_slot];
  }
  return *e;
}
static std::atomic<fclb_handle> g_next_handle{1};
static thread_local fclb_handle t_forced_handle = 0;  // forEachDevice: the handle the first replica got
static thread_local fclb_handle t_last_handle = 0;
fclb_handle newHandle() {
  t_last_handle = t_forced_handle ? t_forced_handle : g_next_handle.fetch_add(1);
  return t_last_handle;
}
void beginReplicas() { t_forced_handle = 0; t_last_handle = 0; }
void nextReplica() { t_forced_handle = t_last_handle; }
void endReplicas() { t_forced_handle = 0; }
int engineCount() { return g_n_engines.load(); }
int currentSlot() { return t_slot; }
int setSlot(int slot) {
  if (slot < 0 || slot >= kMaxDevices || !g_engines[slot] || !g_engines[slot]->ready)
    return fail(FCLB_ERR_BAD_ARG, "device slot not initialised (fclb_init_devices)");
  t_slot = slot;
  FCLB_CUDA(cudaSetDevice(g_engines[slot]->device));
  return FCLB_OK;
}
const std::string& lastErrorString() { return g_err; }
void setLastErrorString(const std::string& s) { g_err = s; }

int ensureInit() {
  Engine& e = eng();
  if (e.ready) {
    // CUDA's current device is per host thread: a thread that never touched the library starts on device 0
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != e.device) FCLB_CUDA(cudaSetDevice(e.device));
    return FCLB_OK;
  }
  return fclb_init(-1);
}

int shardOverDevices(size_t n, const std::function<int(size_t, size_t)>& fn) {
  const int nd = engineCount();
  if (nd <= 1 || n < size_t(nd)) return fn(0, n);
  const int home = currentSlot();
  std::vector<int> rcs(size_t(nd), FCLB_OK);
  std::vector<std::string> msgs{size_t(nd)};
  std::vector<std::thread> ts;
  const size_t base = n / size_t(nd), rem = n % size_t(nd);
  size_t begin = 0;
  for (int s = 0; s < nd; s++) {
    const size_t cnt = base + (size_t(s) < rem ? 1 : 0);
    const size_t b = begin;
    begin += cnt;
    ts.emplace_back([&, s, b, cnt] {
      int rc = setSlot(s);
      if (rc == FCLB_OK && cnt) rc = fn(b, cnt);
      rcs[size_t(s)] = rc;
      if (rc) msgs[size_t(s)] = g_err;
    });
  }
  for (auto& t : ts) t.join();
  setSlot(home);
  for (int s = 0; s < nd; s++)
    if (rcs[size_t(s)]) return fail(rcs[size_t(s)], "device " + std::to_string(s) + ": " + msgs[size_t(s)]);
  return FCLB_OK;
}

// ---------------------------------------------------------------------------
// bucketing kernels
template <typename S>
__global__ void classifyKernel(const ShapeD<S>* __restrict__ shapes, uint32_t n_shapes, const fclb_pair* __restrict__ pairs,
                               size_t n, uint8_t* __restrict__ kind, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kNumKinds];
  for (int i = threadIdx.x; i < kNumKinds; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (size_t q = blockIdx.x * size_t(blockDim.x) + threadIdx.x; q < n; q += size_t(gridDim.x) * blockDim.x) {
    fclb_pair p = pairs[q];
    if (p.shape1 >= n_shapes || p.shape2 >= n_shapes) {  // reported as FCLB_ERR_BAD_ARG; never read out of bounds
      atomicAdd(&hist[2 * kNumKinds], 1u);
      p.shape1 = p.shape2 = 0;
    }
    const int k = (shapes[p.shape1].type & 7) * kNumTypes + (shapes[p.shape2].type & 7);
    kind[q] = uint8_t(k);
    atomicAdd(&sh[k], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNumKinds; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// hist[0..K) counts -> hist[K..2K) exclusive offsets (also used as cursors)
// The histogram also goes to `host_copy` (pinned, mapped host memory) with plain stores: a cudaMemcpyAsync of these 500
// bytes would queue on the device-to-host copy engine BEHIND the previous pipeline stage's result copy (60 MB, 1.3 ms),
// and the host cannot launch this stage's kernels before it has the counts (profiles/r02_e2e_taper.txt).
__global__ void scanKernel(uint32_t* hist, volatile uint32_t* host_copy) {
  __shared__ uint32_t offs[kNumKinds];
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (int i = 0; i < kNumKinds; i++) {
      hist[kNumKinds + i] = acc;
      offs[i] = acc;
      acc += hist[i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNumKinds; i += blockDim.x) {
    host_copy[i] = hist[i];
    host_copy[kNumKinds + i] = offs[i];
  }
  if (threadIdx.x == 0) host_copy[2 * kNumKinds] = hist[2 * kNumKinds];
  __threadfence_system();
}

// Stable within a block-chunk: each block owns a contiguous chunk of queries,
// reserves space per kind with one atomic per (block, kind), and ranks its
// queries locally in order.
__global__ void scatterKernel(const uint8_t* __restrict__ kind, size_t n, size_t chunk, uint32_t* __restrict__ cursors,
                              uint32_t* __restrict__ perm) {
  __shared__ uint32_t cnt[kNumKinds];
  __shared__ uint32_t base[kNumKinds];
  const size_t b = size_t(blockIdx.x) * chunk;
  const size_t e = (b + chunk < n) ? (b + chunk) : n;
  for (int i = threadIdx.x; i < kNumKinds; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  for (size_t q = b + threadIdx.x; q < e; q += blockDim.x) atomicAdd(&cnt[kind[q]], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < kNumKinds; i += blockDim.x) {
    base[i] = cnt[i] ? atomicAdd(&cursors[i], cnt[i]) : 0;
    cnt[i] = 0;
  }
  __syncthreads();
  // placement: tiles of blockDim queries, one shared-memory atomic per
  // (warp, kind) via match_any aggregation
  const unsigned lane = threadIdx.x & 31;
  for (size_t t = b; t < e; t += blockDim.x) {
    const size_t q = t + threadIdx.x;
    const int k = (q < e) ? int(kind[q]) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, k);
    const unsigned rank_in_warp = __popc(peers & ((1u << lane) - 1));
    uint32_t start = 0;
    if (rank_in_warp == 0 && k >= 0) start = atomicAdd(&cnt[k], uint32_t(__popc(peers)));
    start = __shfl_sync(peers, start, __ffs(peers) - 1);
    if (k >= 0) perm[base[k] + start + rank_in_warp] = uint32_t(q);
  }
}

// ---------------------------------------------------------------------------
static int uploadConvexTables(Engine& e) {
  for (int st = 0; st < 2; st++) {
    if (e.d_convex_tab[st]) {
      cudaFree(e.d_convex_tab[st]);
      e.d_convex_tab[st] = nullptr;
    }
  }
  if (e.convex.empty()) return FCLB_OK;
  {
    std::vector<ConvexD<float>> tf(e.convex.size());
    std::vector<ConvexD<double>> td(e.convex.size());
    for (size_t i = 0; i < e.convex.size(); i++) {
      const ConvexHost& c = e.convex[i];
      tf[i].verts = static_cast<const float*>(c.d_verts[0]);
      td[i].verts = static_cast<const double*>(c.d_verts[1]);
      tf[i].nbr = td[i].nbr = c.d_nbr;
      tf[i].vinfo = td[i].vinfo = static_cast<const int2*>(c.d_vinfo);
      tf[i].n_verts = td[i].n_verts = c.n_verts;
      tf[i].walk = td[i].walk = c.walk;
      for (int k = 0; k < 6; k++) {
        tf[i].seed[k] = c.seed[0][k];
        td[i].seed[k] = c.seed[1][k];
      }
      for (int k = 0; k < 3; k++) {
        tf[i].interior[k] = float(c.interior[0][k]);
        td[i].interior[k] = c.interior[1][k];
      }
    }
    FCLB_CUDA(cudaMalloc(&e.d_convex_tab[0], tf.size() * sizeof(ConvexD<float>)));
    FCLB_CUDA(cudaMalloc(&e.d_convex_tab[1], td.size() * sizeof(ConvexD<double>)));
    FCLB_CUDA(uploadSync(e.d_convex_tab[0], tf.data(), tf.size() * sizeof(ConvexD<float>)));
    FCLB_CUDA(uploadSync(e.d_convex_tab[1], td.data(), td.size() * sizeof(ConvexD<double>)));
  }
  return FCLB_OK;
}

template <typename S>
static void convexDerive(const std::vector<double>& verts_d, int n, int (&seed)[6], double (&interior)[3],
                         std::vector<S>& verts_s) {
  verts_s.resize(size_t(3) * n);
  for (size_t i = 0; i < verts_s.size(); i++) verts_s[i] = S(verts_d[i]);
  // interior point: running sum in S, times (S)(1.0/n)   (convex-inl.h:64-71)
  S sum[3] = {S(0), S(0), S(0)};
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) sum[k] += verts_s[3 * i + k];
  const S inv = S(1.0 / double(n));
  for (int k = 0; k < 3; k++) interior[k] = double(sum[k] * inv);
  // six axis seeds by the naive scan (convex-inl.h:153-200): dot with a unit
  // axis picks one coordinate; first strict maximum wins.
  for (int k = 0; k < 6; k++) {
    const int axis = k / 2;
    const S sgn = (k & 1) ? S(-1) : S(1);
    int best = 0;
    S best_v = sgn * verts_s[axis];
    for (int i = 1; i < n; i++) {
      const S v = sgn * verts_s[3 * i + axis];
      if (v > best_v) {
        best = i;
        best_v = v;
      }
    }
    seed[k] = best;
  }
}

template <typename S>
static int buildShapeTable(Engine& e, ShapeTable* t, int st) {
  std::vector<ShapeD<S>> h(t->n);
  for (uint32_t i = 0; i < t->n; i++) {
    const fclb_shape& s = t->host[i];
    if (s.type > FCLB_CONVEX) return fail(FCLB_ERR_BAD_ARG, "fclb_shapes_upload: unknown shape type");
    if (s.type == FCLB_CONVEX && s.geom >= e.convex.size())
      return fail(FCLB_ERR_BAD_ARG, "fclb_shapes_upload: convex slot out of range");
    h[i].type = int(s.type);
    h[i].geom = int(s.geom);
    for (int k = 0; k < 3; k++) h[i].p[k] = S(s.p[k]);
  }
  FCLB_CUDA(cudaMalloc(&t->d_shapes[st], std::max<size_t>(1, h.size()) * sizeof(ShapeD<S>)));
  FCLB_CUDA(uploadSync(t->d_shapes[st], h.data(), h.size() * sizeof(ShapeD<S>)));
  // bounding polytopes of the primitives (getBoundVertices in the shape's own frame)
  std::vector<BoundD<S>> bd(t->n);
  for (uint32_t i = 0; i < t->n; i++) boundVertices<S>(t->host[i].type, h[i].p, bd[i]);
  FCLB_CUDA(cudaMalloc(&t->d_bound[st], std::max<size_t>(1, bd.size()) * sizeof(BoundD<S>)));
  FCLB_CUDA(uploadSync(t->d_bound[st], bd.data(), bd.size() * sizeof(BoundD<S>)));
  std::vector<LocalAabbD<S>> la(t->n);
  for (uint32_t i = 0; i < t->n; i++) {
    const bool cvx = t->host[i].type == FCLB_CONVEX;
    const ConvexHost* c = cvx ? &e.convex[t->host[i].geom] : nullptr;
    localAabb<S>(t->host[i].type, h[i].p, cvx ? c->h_verts.data() : nullptr, cvx ? c->n_verts : 0, la[i]);
  }
  FCLB_CUDA(cudaMalloc(&t->d_local[st], std::max<size_t>(1, la.size()) * sizeof(LocalAabbD<S>)));
  FCLB_CUDA(uploadSync(t->d_local[st], la.data(), la.size() * sizeof(LocalAabbD<S>)));
  return FCLB_OK;
}

static int ensureScratch(Engine& e, size_t n) {
  if (n <= e.scratch_cap) return FCLB_OK;
  if (e.d_perm) cudaFree(e.d_perm);
  if (e.d_kind) cudaFree(e.d_kind);
  e.d_perm = nullptr;
  e.d_kind = nullptr;
  e.scratch_cap = 0;
  FCLB_CUDA(cudaMalloc(&e.d_perm, n * sizeof(uint32_t)));
  FCLB_CUDA(cudaMalloc(&e.d_kind, n));
  e.scratch_cap = n;
  return FCLB_OK;
}

// Bucket a device-resident batch.  On return counts/offsets hold the per-kind
// histogram; *uniform_kind >= 0 when every query has the same kind (then no
// permutation is needed and perm stays unused).
template <typename S>
int bucketBatch(Engine& e, const ShapeTable* t, const fclb_pair* d_pairs, size_t n, uint32_t* counts,
                       uint32_t* offsets, int* uniform_kind) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  int rc = ensureScratch(e, n);
  if (rc) return rc;
  FCLB_CUDA(cudaMemsetAsync(e.d_hist, 0, (2 * kNumKinds + 1) * sizeof(uint32_t), e.compute));
  const int block = 256;
  const int grid = int(std::min<size_t>((n + block - 1) / block, size_t(e.sms) * 8));
  classifyKernel<S><<<grid, block, 0, e.compute>>>(static_cast<const ShapeD<S>*>(t->d_shapes[st]), t->n, d_pairs, n, e.d_kind,
                                                   e.d_hist);
  scanKernel<<<1, 64, 0, e.compute>>>(e.d_hist, e.h_hist_dev);
  e.launches += 2;
  FCLB_CUDA(cudaGetLastError());
  static const bool hist_copy = getenv("FCLB_HIST_COPY") != nullptr;  // A/B: fetch it with a copy as well (the old path)
  if (hist_copy)
    FCLB_CUDA(cudaMemcpyAsync(e.h_hist, e.d_hist, (2 * kNumKinds + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  if (e.h_hist[2 * kNumKinds]) return fail(FCLB_ERR_BAD_ARG, "a pair names a shape index outside the shape table");
  *uniform_kind = -1;
  for (int k = 0; k < kNumKinds; k++) {
    counts[k] = e.h_hist[k];
    offsets[k] = e.h_hist[kNumKinds + k];
    if (counts[k] == n) *uniform_kind = k;
  }
  if (*uniform_kind < 0) {
    const size_t chunk = 4096;
    const int sgrid = int((n + chunk - 1) / chunk);
    scatterKernel<<<sgrid, 256, 0, e.compute>>>(e.d_kind, n, chunk, e.d_hist + kNumKinds, e.d_perm);
    e.launches += 1;
    FCLB_CUDA(cudaGetLastError());
  }
  return FCLB_OK;
}

SolverParams solverParams(int scalar_type, double gjk_tol, uint32_t gjk_max_iter, double epa_tol, uint32_t epa_max_faces,
                          uint32_t epa_max_iter, bool collide_defaults) {
  SolverParams sp{};
  // constants<S>::eps_78(): pow(eps, 7/8) evaluated in double, rounded to S
  const double eps = scalar_type == FCLB_F32 ? double(1.1920928955078125e-07f) : 2.220446049250313e-16;
  const double e78 = std::pow(eps, 7. / 8.);
  sp.eps78 = scalar_type == FCLB_F32 ? double(float(e78)) : e78;
  // GJKSolver defaults are eps^(7/8) (gjk_solver-inl.h:1121-1130); fcl::collide
  // overrides both tolerances with the request's 1e-6 (collision_interface-inl.h:19-20)
  const double dflt = collide_defaults ? 1e-6 : sp.eps78;
  sp.gjk_tol = gjk_tol > 0 ? gjk_tol : dflt;
  sp.epa_tol = epa_tol > 0 ? epa_tol : dflt;
  sp.gjk_max_iter = gjk_max_iter ? int(gjk_max_iter) : 128;
  sp.epa_max_faces = epa_max_faces ? int(epa_max_faces) : 256;
  sp.epa_max_iter = epa_max_iter ? int(epa_max_iter) : 255;
  return sp;
}
static SolverParams distanceParams(int scalar_type, double gjk_tol, uint32_t gjk_max_iter) {
  return solverParams(scalar_type, gjk_tol, gjk_max_iter, 0.0, 0, 0, false);
}

template int bucketBatch<float>(Engine&, const ShapeTable*, const fclb_pair*, size_t, uint32_t*, uint32_t*, int*);
template int bucketBatch<double>(Engine&, const ShapeTable*, const fclb_pair*, size_t, uint32_t*, uint32_t*, int*);

template <typename S>
static int distanceDev(Engine& e, ShapeTable* t, const fclb_pair* pairs, const void* poses1, const void* poses2, size_t n,
                       const SolverParams& sp, const DistanceOut& out) {
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
    b.perm = (uniform >= 0) ? nullptr : e.d_perm;
    b.begin = (uniform >= 0) ? 0 : offsets[k];
    b.count = counts[k];
    b.type1 = k / kNumTypes;
    b.type2 = k % kNumTypes;
    FCLB_CUDA(launchDistance<S>(b, sp, out, e.compute, &launches));
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

ShapeTable* findTable(Engine& e, fclb_handle h) {
  auto it = e.tables.find(h);
  return it == e.tables.end() ? nullptr : it->second;
}

int ensureChunkEvents(Engine& e, int n) {
  while (int(e.ev_in.size()) < n) {
    cudaEvent_t a = nullptr, b = nullptr;
    FCLB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    FCLB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    e.ev_in.push_back(a);
    e.ev_done.push_back(b);
  }
  return FCLB_OK;
}

unsigned long long* gjkCursor() {
  static unsigned long long* d[kMaxDevices] = {};
  unsigned long long*& p = d[currentSlot()];
  if (!p && cudaMalloc(&p, sizeof(unsigned long long)) != cudaSuccess) p = nullptr;
  return p;
}

int ensureStage(Engine& e, size_t bytes) {
  if (bytes <= e.stage_cap) return FCLB_OK;
  if (e.d_stage) cudaFree(e.d_stage);
  e.d_stage = nullptr;
  e.stage_cap = 0;
  FCLB_CUDA(cudaMalloc(&e.d_stage, bytes));
  e.stage_cap = bytes;
  return FCLB_OK;
}


}  // namespace fclb

using namespace fclb;

// GJKSolver::shapeSignedDistance (gjk_solver-inl.h:810-868) assembled from the two device paths: the generic GJK
// distance answers the separated queries; for the others the GJK + EPA path gives the penetration, reported
// as a negative distance with the EPA witness points mapped by tf1.
template <typename S>
__global__ void signedDistanceCombineKernel(const S* __restrict__ poses1, size_t n, const int32_t* __restrict__ gjk,
                                            const int32_t* __restrict__ epa, const S* __restrict__ geom, S* dist, S* p1,
                                            S* p2, uint8_t* ok) {
  const size_t q = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (q >= n) return;
  if (ok[q] == 1) return;  // Separated with valid witness points: already final
  S d = S(-1);
  V3<S> a = zero3<S>(), b = zero3<S>();
  uint8_t r = 0;
  if (gjk[q] == 0 /* GJK_Status::Intersect */ && epa[q] > 0 /* EPA ran and did not fail */) {
    const Pose<S> tf1 = loadPose(poses1, q);
    const S* g = geom + 7 * q;
    d = -g[0];
    a = apply(tf1, mk<S>(g[1], g[2], g[3]));
    b = apply(tf1, mk<S>(g[4], g[5], g[6]));
    r = 1;
  }
  if (dist) dist[q] = d;
  if (p1) store3(p1, q, a);
  if (p2) store3(p2, q, b);
  ok[q] = r;
}

// FMA-chain microbenchmark: the FP32 / FP64 CUDA-core peak the compute-bound kernels are compared with
// (SURVEY.md 8d asks for measured, not nominal, FP peaks).  16 independent accumulators per thread.
template <typename S>
__global__ void __launch_bounds__(256) fmaPeakKernel(S* out, int iters, S a, S b) {
  S acc[16];
#pragma unroll
  for (int k = 0; k < 16; k++) acc[k] = S(threadIdx.x + k);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = fma(acc[k], a, b);  // explicit: this file is compiled with --fmad=false
  }
  S sum = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) sum += acc[k];
  if (sum == S(-12345)) out[0] = sum;  // keep the chain alive
}

// FCLB_POSE_QT7 -> the 12 S pose the kernels read: Eigen's QuaternionBase::toRotationMatrix arithmetic in S (no FMA
// contraction in this file), so that the expanded pose equals tf.linear() of a Transform3<S> the caller would have built
// from the same quaternion.  28 B instead of 48 B per pose cross PCIe.
template <typename S>
__global__ void expandQt7Kernel(const S* __restrict__ qt, size_t n, S* __restrict__ out) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const S* q = qt + 7 * i;
    const S x = q[0], y = q[1], z = q[2], w = q[3];
    const S tx = S(2) * x, ty = S(2) * y, tz = S(2) * z;
    const S twx = tx * w, twy = ty * w, twz = tz * w;
    const S txx = tx * x, txy = ty * x, txz = tz * x;
    const S tyy = ty * y, tyz = tz * y, tzz = tz * z;
    S* o = out + 12 * i;
    o[0] = S(1) - (tyy + tzz);
    o[1] = txy - twz;
    o[2] = txz + twy;
    o[3] = txy + twz;
    o[4] = S(1) - (txx + tzz);
    o[5] = tyz - twx;
    o[6] = txz - twy;
    o[7] = tyz + twx;
    o[8] = S(1) - (txx + tyy);
    o[9] = q[4];
    o[10] = q[5];
    o[11] = q[6];
  }
}
static int launchExpandQt7(Engine& e, int scalar_type, const void* qt, size_t n, void* out) {
  const int grid = int(std::min<size_t>((n + 255) / 256, size_t(e.sms) * 16));
  if (scalar_type == FCLB_F32)
    expandQt7Kernel<float><<<grid, 256, 0, e.compute>>>(static_cast<const float*>(qt), n, static_cast<float*>(out));
  else
    expandQt7Kernel<double><<<grid, 256, 0, e.compute>>>(static_cast<const double*>(qt), n, static_cast<double*>(out));
  e.launches += 1;
  FCLB_CUDA(cudaGetLastError());
  return FCLB_OK;
}

// L2 read-bandwidth microbenchmark: every CTA streams the same 32 MB buffer (L2-resident after the first pass) with
// 128-bit ld.global.cg loads -- the ceiling for the traversal kernels, whose node / triangle arrays live in L2.
__global__ void __launch_bounds__(256) l2ReadKernel(const uint4* __restrict__ buf, size_t n_vec, int passes, uint4* out) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (int p = 0; p < passes; p++) {
    for (size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x + size_t(p) * 977) % stride; i < n_vec; i += stride) {
      const uint4 v = __ldcg(buf + i);
      acc.x ^= v.x;
      acc.y ^= v.y;
      acc.z ^= v.z;
      acc.w ^= v.w;
    }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) out[0] = acc;  // keep the loads alive
}

extern "C" {

const char* fclb_last_error(void) { return g_err.c_str(); }
const char* fclb_version(void) { return "fclb200 0.1 (sm_100a)"; }

int fclb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// streams, events and scratch of one engine on `device` (the caller holds the engine's mutex)
static int initEngine(Engine& e, int device) {
  FCLB_CUDA(cudaSetDevice(device));
  e.device = device;
  FCLB_CUDA(cudaDeviceGetAttribute(&e.sms, cudaDevAttrMultiProcessorCount, device));
  FCLB_CUDA(cudaStreamCreateWithFlags(&e.compute, cudaStreamNonBlocking));
  FCLB_CUDA(cudaStreamCreateWithFlags(&e.copy_in, cudaStreamNonBlocking));
  FCLB_CUDA(cudaStreamCreateWithFlags(&e.copy_out, cudaStreamNonBlocking));
  FCLB_CUDA(cudaStreamCreateWithFlags(&e.aux, cudaStreamNonBlocking));
  FCLB_CUDA(cudaEventCreateWithFlags(&e.ev_aux0, cudaEventDisableTiming));
  FCLB_CUDA(cudaEventCreateWithFlags(&e.ev_aux1, cudaEventDisableTiming));
  FCLB_CUDA(cudaEventCreate(&e.ev0));
  FCLB_CUDA(cudaEventCreate(&e.ev1));
  FCLB_CUDA(cudaEventCreate(&e.ev_call0));
  for (int i = 0; i <= Engine::kMaxRec; i++) FCLB_CUDA(cudaEventCreate(&e.rec_ev[i]));
  {  // keep the stream-ordered pool's memory across synchronisations (the builders allocate their scratch from it)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  FCLB_CUDA(cudaMalloc(&e.d_hist, (2 * kNumKinds + 1) * sizeof(uint32_t)));
  FCLB_CUDA(cudaHostAlloc(&e.h_hist, (2 * kNumKinds + 1) * sizeof(uint32_t), cudaHostAllocPortable | cudaHostAllocMapped));
  FCLB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&e.h_hist_dev), e.h_hist, 0));
  if (const char* hc = getenv("FCLB_HOST_CHUNK")) {  // queries per pipeline stage of the *_host entry points (tuning)
    const long long v = atoll(hc);
    if (v >= 1024) e.host_chunk = size_t(v);
  }
  if (const char* hh = getenv("FCLB_HOST_HEAD")) {  // first (short) stage of the compute-bound *_host pipelines (0: equal stages)
    const long long v = atoll(hh);
    e.host_head = v <= 0 ? 0 : std::max<size_t>(size_t(v), 1024);
  }
  if (const char* ht = getenv("FCLB_HOST_TAPER")) {  // shortest tapered stage of fclb_distance_batch_*host (0: equal stages)
    const long long v = atoll(ht);
    e.host_taper = v <= 0 ? 0 : std::max<size_t>(size_t(v), 4096);
  }
  e.ready = true;
  return FCLB_OK;
}

int fclb_init(int device) {
  const int home = t_slot;
  t_slot = 0;  // the one-GPU-per-process binding is slot 0
  Engine& e = eng();
  t_slot = home;
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  if (e.ready && (device < 0 || device == e.device)) return FCLB_OK;
  if (e.ready) return fail(FCLB_ERR_BAD_ARG, "fclb_init: engine already bound to another device (one process per GPU)");
  const int n = fclb_device_count();
  if (n <= 0) return fail(FCLB_ERR_NO_DEVICE, "no CUDA device visible: libfclb200 has no CPU fallback");
  if (device < 0) device = 0;
  if (device >= n) return fail(FCLB_ERR_BAD_ARG, "fclb_init: device index out of range");
  const int rc = initEngine(e, device);
  if (rc) return rc;
  if (g_n_engines.load() < 1) g_n_engines.store(1);
  return FCLB_OK;
}

// SURVEY.md 8(b) fclb_init(int n_devices): enumerate the GPUs of the box, one engine (streams, scratch, replicas of every
// geometry uploaded afterwards) per device.  n_devices <= 0: every visible device.  Must precede the first upload.
int fclb_init_devices(int n_devices) {
  const int n = fclb_device_count();
  if (n <= 0) return fail(FCLB_ERR_NO_DEVICE, "no CUDA device visible: libfclb200 has no CPU fallback");
  if (n_devices <= 0 || n_devices > n) n_devices = n;
  if (n_devices > kMaxDevices) n_devices = kMaxDevices;
  const int home = t_slot;
  int rc = FCLB_OK;
  for (int s = 0; s < n_devices && rc == FCLB_OK; s++) {
    t_slot = s;
    Engine& e = eng();
    std::lock_guard<std::recursive_mutex> lk(e.mu);
    if (e.ready) {
      if (e.device != s) rc = fail(FCLB_ERR_BAD_ARG, "fclb_init_devices: slot 0 is already bound to another GPU by fclb_init");
      continue;
    }
    if (s > 0 && (!g_engines[0]->tables.empty() || !g_engines[0]->convex.empty()))
      rc = fail(FCLB_ERR_BAD_ARG, "fclb_init_devices must be called before the first geometry upload");
    else
      rc = initEngine(e, s);
  }
  t_slot = home;
  if (rc) return rc;
  if (g_n_engines.load() < n_devices) g_n_engines.store(n_devices);
  return setSlot(home);
}
int fclb_num_devices(void) { return g_n_engines.load(); }
int fclb_set_device(int slot) { return setSlot(slot); }

int fclb_host_alloc(void** p, size_t bytes) {
  int rc = ensureInit();
  if (rc) return rc;
  FCLB_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));  // usable by every engine's copy streams
  return FCLB_OK;
}
int fclb_host_free(void* p) {
  FCLB_CUDA(cudaFreeHost(p));
  return FCLB_OK;
}
int fclb_dev_alloc(void** p, size_t bytes) {
  int rc = ensureInit();
  if (rc) return rc;
  FCLB_CUDA(cudaMalloc(p, bytes ? bytes : 1));
  return FCLB_OK;
}
int fclb_dev_free(void* p) {
  FCLB_CUDA(cudaFree(p));
  return FCLB_OK;
}
int fclb_memcpy_h2d(void* dst, const void* src, size_t bytes) {
  int rc = ensureInit();
  if (rc) return rc;
  FCLB_CUDA(uploadSync(dst, src, bytes));
  return FCLB_OK;
}
int fclb_memcpy_d2h(void* dst, const void* src, size_t bytes) {
  int rc = ensureInit();
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return FCLB_OK;
}
int fclb_synchronize(void) {
  int rc = ensureInit();
  if (rc) return rc;
  FCLB_CUDA(cudaDeviceSynchronize());
  return FCLB_OK;
}

int fclb_measure_fp_peak(int scalar_type, double* tflops) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!tflops || (scalar_type != FCLB_F32 && scalar_type != FCLB_F64)) return fail(FCLB_ERR_BAD_ARG, "fclb_measure_fp_peak: bad argument");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  void* d_out = nullptr;
  FCLB_CUDA(cudaMalloc(&d_out, 64));
  const int grid = e.sms * 8, block = 256;
  const int iters = scalar_type == FCLB_F32 ? 4096 : 1024;
  float best_ms = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
    if (scalar_type == FCLB_F32)
      fmaPeakKernel<float><<<grid, block, 0, e.compute>>>(static_cast<float*>(d_out), iters, 1.000001f, 0.5f);
    else
      fmaPeakKernel<double><<<grid, block, 0, e.compute>>>(static_cast<double*>(d_out), iters, 1.000001, 0.5);
    FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.ev0, e.ev1);
    if (rep > 0 && ms < best_ms) best_ms = ms;
    e.launches += 1;
  }
  cudaFree(d_out);
  const double flops = 2.0 * 16.0 * double(iters) * double(grid) * double(block);
  *tflops = flops / (double(best_ms) * 1e-3) / 1e12;
  return FCLB_OK;
}

int fclb_measure_l2_bandwidth(double* gbs) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!gbs) return fail(FCLB_ERR_BAD_ARG, "fclb_measure_l2_bandwidth: null argument");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t bytes = size_t(32) << 20, n_vec = bytes / sizeof(uint4);
  uint4* d = nullptr;
  FCLB_CUDA(cudaMalloc(&d, bytes + sizeof(uint4)));
  FCLB_CUDA(cudaMemsetAsync(d, 1, bytes + sizeof(uint4), e.compute));
  const int grid = e.sms * 8, passes = 16;
  float best_ms = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
    l2ReadKernel<<<grid, 256, 0, e.compute>>>(d, n_vec, passes, d + n_vec);
    FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.ev0, e.ev1);
    if (rep > 0 && ms < best_ms) best_ms = ms;
    e.launches += 1;
  }
  cudaFree(d);
  *gbs = double(bytes) * passes / (double(best_ms) * 1e-3) / 1e9;
  return FCLB_OK;
}

static int convex_upload_one(const double* verts, int n_verts, const int* faces, int faces_len, int num_faces,
                       uint32_t* slot) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!verts || !faces || n_verts <= 0 || num_faces <= 0 || !slot)
    return fail(FCLB_ERR_BAD_ARG, "fclb_convex_upload: null or empty input");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  ConvexHost c;
  c.n_verts = n_verts;
  // neighbour CSR (convex-inl.h:379-407): std::set => ascending, de-duplicated
  std::vector<std::set<int>> nb(n_verts);
  std::map<std::pair<int, int>, int> edge_faces;
  int fi = 0;
  for (int f = 0; f < num_faces; f++) {
    if (fi >= faces_len) return fail(FCLB_ERR_BAD_ARG, "fclb_convex_upload: faces array too short");
    const int cnt = faces[fi];
    if (cnt < 1 || fi + cnt > faces_len - 1) return fail(FCLB_ERR_BAD_ARG, "fclb_convex_upload: faces array too short");
    int prev = faces[fi + cnt];
    for (int i = fi + 1; i <= fi + cnt; i++) {
      const int v = faces[i];
      if (v < 0 || v >= n_verts || prev < 0 || prev >= n_verts)
        return fail(FCLB_ERR_BAD_ARG, "fclb_convex_upload: vertex index out of range");
      nb[v].insert(prev);
      nb[prev].insert(v);
      edge_faces[std::make_pair(std::min(v, prev), std::max(v, prev))] += 1;
      prev = v;
    }
    fi += cnt + 1;
  }
  std::vector<int> csr(n_verts);
  bool all_connected = true;
  for (int v = 0; v < n_verts; v++) {
    csr[v] = int(csr.size());
    csr.push_back(int(nb[v].size()));
    csr.insert(csr.end(), nb[v].begin(), nb[v].end());
    if (nb[v].empty()) all_connected = false;
  }
  // ValidateTopology (convex-inl.h:293-368): walk only on a watertight,
  // fully connected mesh with more than 32 vertices (convex.h:259).
  bool watertight = true;
  for (const auto& kv : edge_faces)
    if (kv.second != 2) watertight = false;
  c.walk = (n_verts > 32 && watertight && all_connected) ? 1 : 0;
  std::vector<double> vd(verts, verts + size_t(3) * n_verts);
  c.h_verts = vd;
  std::vector<float> vf;
  std::vector<double> vdd;
  convexDerive<float>(vd, n_verts, c.seed[0], c.interior[0], vf);
  convexDerive<double>(vd, n_verts, c.seed[1], c.interior[1], vdd);
  // device layout: 4 S per vertex (one 128-bit load), and (first neighbour, count) per vertex beside the CSR
  std::vector<float> vf4(size_t(4) * n_verts, 0.f);
  std::vector<double> vd4(size_t(4) * n_verts, 0.0);
  std::vector<int> vinfo(size_t(2) * n_verts);
  for (int v = 0; v < n_verts; v++) {
    for (int k = 0; k < 3; k++) {
      vf4[4 * size_t(v) + k] = vf[3 * size_t(v) + k];
      vd4[4 * size_t(v) + k] = vdd[3 * size_t(v) + k];
    }
    vinfo[2 * size_t(v)] = csr[v] + 1;
    vinfo[2 * size_t(v) + 1] = csr[csr[v]];
  }
  FCLB_CUDA(cudaMalloc(&c.d_verts[0], vf4.size() * sizeof(float)));
  FCLB_CUDA(cudaMalloc(&c.d_verts[1], vd4.size() * sizeof(double)));
  FCLB_CUDA(cudaMalloc(&c.d_nbr, csr.size() * sizeof(int)));
  FCLB_CUDA(cudaMalloc(&c.d_vinfo, vinfo.size() * sizeof(int)));
  FCLB_CUDA(uploadSync(c.d_verts[0], vf4.data(), vf4.size() * sizeof(float)));
  FCLB_CUDA(uploadSync(c.d_verts[1], vd4.data(), vd4.size() * sizeof(double)));
  FCLB_CUDA(uploadSync(c.d_nbr, csr.data(), csr.size() * sizeof(int)));
  FCLB_CUDA(uploadSync(c.d_vinfo, vinfo.data(), vinfo.size() * sizeof(int)));
  e.convex.push_back(c);
  e.convex_epoch++;
  *slot = uint32_t(e.convex.size() - 1);
  return uploadConvexTables(e);
}
int fclb_convex_upload(const double* verts, int n_verts, const int* faces, int faces_len, int num_faces,
                       uint32_t* slot) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return convex_upload_one(verts, n_verts, faces, faces_len, num_faces, slot); });
}

static int shapes_upload_one(const fclb_shape* shapes, uint32_t n_shapes, fclb_handle* table) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!shapes || !table || n_shapes == 0) return fail(FCLB_ERR_BAD_ARG, "fclb_shapes_upload: null or empty input");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  ShapeTable* t = new ShapeTable();
  t->host.assign(shapes, shapes + n_shapes);
  t->n = n_shapes;
  rc = buildShapeTable<float>(e, t, 0);
  if (!rc) rc = buildShapeTable<double>(e, t, 1);
  if (rc) {
    delete t;
    return rc;
  }
  const fclb_handle h = newHandle();
  e.tables[h] = t;
  *table = h;
  return FCLB_OK;
}
int fclb_shapes_upload(const fclb_shape* shapes, uint32_t n_shapes, fclb_handle* table) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return shapes_upload_one(shapes, n_shapes, table); });
}

static int release_one(fclb_handle h) {
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = e.tables.find(h);
  if (it == e.tables.end()) return fail(FCLB_ERR_BAD_ARG, "fclb_release: unknown handle");
  for (int st = 0; st < 2; st++)
  {
    if (it->second->d_shapes[st]) cudaFree(it->second->d_shapes[st]);
    if (it->second->d_bound[st]) cudaFree(it->second->d_bound[st]);
    if (it->second->d_local[st]) cudaFree(it->second->d_local[st]);
  }
  delete it->second;
  e.tables.erase(it);
  return FCLB_OK;
}
int fclb_release(fclb_handle h) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return release_one(h); });
}

int fclb_distance_batch_dev(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                            size_t n, int scalar_type, double gjk_tol, uint32_t gjk_max_iter, void* out_dist,
                            void* out_p1, void* out_p2, uint8_t* out_ok) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  ShapeTable* t = findTable(e, shapes);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "fclb_distance_batch: unknown shape table handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (n > 0xffffffffull) return fail(FCLB_ERR_CAPACITY, "batch larger than 2^32-1 queries: split it");
  if (!pairs || !poses1 || !poses2) return fail(FCLB_ERR_BAD_ARG, "fclb_distance_batch: null input array");
  const SolverParams sp = distanceParams(scalar_type, gjk_tol, gjk_max_iter);
  DistanceOut out{out_dist, out_p1, out_p2, out_ok};
  if (scalar_type == FCLB_F32) return distanceDev<float>(e, t, pairs, poses1, poses2, n, sp, out);
  return distanceDev<double>(e, t, pairs, poses1, poses2, n, sp, out);
}

int fclb_signed_distance_batch_dev(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                   size_t n, int scalar_type, void* out_dist, void* out_p1, void* out_p2, uint8_t* out_ok) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  ShapeTable* t = findTable(e, shapes);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "fclb_signed_distance_batch: unknown shape table handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (n > 0xffffffffull) return fail(FCLB_ERR_CAPACITY, "batch larger than 2^32-1 queries: split it");
  if (!pairs || !poses1 || !poses2 || !out_ok) return fail(FCLB_ERR_BAD_ARG, "fclb_signed_distance_batch: null array (out_ok is required)");
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  // pass 1: generic GJK distance (GJKSolver defaults: eps^(7/8), 128 iterations)
  SolverParams sp = distanceParams(scalar_type, 0.0, 0);
  sp.generic_only = 1;
  DistanceOut out{out_dist, out_p1, out_p2, out_ok};
  rc = scalar_type == FCLB_F32 ? distanceDev<float>(e, t, pairs, poses1, poses2, n, sp, out)
                               : distanceDev<double>(e, t, pairs, poses1, poses2, n, sp, out);
  if (rc) return rc;
  // pass 2: GJK + EPA (256 faces, 255 iterations, eps^(7/8)) for the penetration of the others
  int32_t *d_gjk = nullptr, *d_epa = nullptr;
  void* d_geom = nullptr;
  struct Scratch {  // freed on every return path, the early ones of FCLB_CUDA included
    int32_t*& a;
    int32_t*& b;
    void*& c;
    ~Scratch() {
      cudaFree(a);
      cudaFree(b);
      cudaFree(c);
    }
  } scratch{d_gjk, d_epa, d_geom};
  FCLB_CUDA(cudaMalloc(&d_gjk, n * 4));
  FCLB_CUDA(cudaMalloc(&d_epa, n * 4));
  FCLB_CUDA(cudaMalloc(&d_geom, n * 7 * ss));
  fclb_request req{};
  req.max_contacts = 1;
  req.penetration_mode = FCLB_PEN_DEFAULT_GJK_EPA;
  req.binary_tol = sp.eps78;
  req.distance_tol = sp.eps78;
  rc = fclb_gjk_epa_batch_dev(shapes, pairs, poses1, poses2, n, scalar_type, &req, d_gjk, d_epa, d_geom);
  if (!rc) {
    const int grid = int((n + 255) / 256);
    if (scalar_type == FCLB_F32)
      signedDistanceCombineKernel<float><<<grid, 256, 0, e.compute>>>(static_cast<const float*>(poses1), n, d_gjk, d_epa,
                                                                    static_cast<const float*>(d_geom),
                                                                    static_cast<float*>(out_dist), static_cast<float*>(out_p1),
                                                                    static_cast<float*>(out_p2), out_ok);
    else
      signedDistanceCombineKernel<double><<<grid, 256, 0, e.compute>>>(static_cast<const double*>(poses1), n, d_gjk, d_epa,
                                                                     static_cast<const double*>(d_geom),
                                                                     static_cast<double*>(out_dist),
                                                                     static_cast<double*>(out_p1), static_cast<double*>(out_p2),
                                                                     out_ok);
    e.launches += 1;
    if (cudaStreamSynchronize(e.compute) != cudaSuccess) rc = fail(FCLB_ERR_CUDA, "signed distance combine failed");
  }
  return rc;
}

static int signed_distance_batch_host_one(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                    size_t n, int scalar_type, void* out_dist, void* out_p1, void* out_p2, uint8_t* out_ok) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!pairs || !poses1 || !poses2) return fail(FCLB_ERR_BAD_ARG, "fclb_signed_distance_batch: null input array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  char* base = nullptr;
  const size_t o_pairs = 0;
  const size_t o_p1 = alignUp(o_pairs + n * sizeof(fclb_pair), 256);
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_dist = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_w1 = alignUp(o_dist + n * ss, 256);
  const size_t o_w2 = alignUp(o_w1 + n * 3 * ss, 256);
  const size_t o_ok = alignUp(o_w2 + n * 3 * ss, 256);
  const size_t total = alignUp(o_ok + n, 256);
  FCLB_CUDA(cudaMalloc(reinterpret_cast<void**>(&base), total));  // (the engine's staging arena is used by the inner calls)
  cudaMemcpyAsync(base + o_pairs, pairs, n * sizeof(fclb_pair), cudaMemcpyHostToDevice, e.compute);
  cudaMemcpyAsync(base + o_p1, poses1, n * 12 * ss, cudaMemcpyHostToDevice, e.compute);
  cudaMemcpyAsync(base + o_p2, poses2, n * 12 * ss, cudaMemcpyHostToDevice, e.compute);
  rc = fclb_signed_distance_batch_dev(shapes, reinterpret_cast<const fclb_pair*>(base + o_pairs), base + o_p1, base + o_p2, n,
                                      scalar_type, base + o_dist, base + o_w1, base + o_w2,
                                      reinterpret_cast<uint8_t*>(base + o_ok));
  if (!rc) {
    if (out_dist) cudaMemcpyAsync(out_dist, base + o_dist, n * ss, cudaMemcpyDeviceToHost, e.compute);
    if (out_p1) cudaMemcpyAsync(out_p1, base + o_w1, n * 3 * ss, cudaMemcpyDeviceToHost, e.compute);
    if (out_p2) cudaMemcpyAsync(out_p2, base + o_w2, n * 3 * ss, cudaMemcpyDeviceToHost, e.compute);
    if (out_ok) cudaMemcpyAsync(out_ok, base + o_ok, n, cudaMemcpyDeviceToHost, e.compute);
    if (cudaStreamSynchronize(e.compute) != cudaSuccess) rc = fail(FCLB_ERR_CUDA, "signed distance copy-out failed");
  }
  cudaFree(base);
  return rc;
}
int fclb_signed_distance_batch_host(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                    size_t n, int scalar_type, void* out_dist, void* out_p1, void* out_p2, uint8_t* out_ok) {
  if (engineCount() <= 1) return signed_distance_batch_host_one(shapes, pairs, poses1, poses2, n, scalar_type, out_dist, out_p1, out_p2, out_ok);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return signed_distance_batch_host_one(shapes, offT(pairs, b), offPtr(poses1, b * 12 * ss), offPtr(poses2, b * 12 * ss), m_, scalar_type, offPtr(out_dist, b * ss), offPtr(out_p1, b * 3 * ss), offPtr(out_p2, b * 3 * ss), offT(out_ok, b)); });
}

static int distance_batch_host_fmt(int pose_format, fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                   size_t n, int scalar_type, double gjk_tol, uint32_t gjk_max_iter, void* out_dist,
                                   void* out_p1, void* out_p2, uint8_t* out_ok) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!pairs || !poses1 || !poses2) return fail(FCLB_ERR_BAD_ARG, "fclb_distance_batch: null input array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> batch_lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  // device layout of the staging arena
  const size_t o_pairs = 0;
  const size_t o_p1 = alignUp(o_pairs + n * sizeof(fclb_pair), 256);
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_dist = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_w1 = alignUp(o_dist + n * ss, 256);
  const size_t o_w2 = alignUp(o_w1 + n * 3 * ss, 256);
  const size_t o_ok = alignUp(o_w2 + n * 3 * ss, 256);
  const bool qt = pose_format == FCLB_POSE_QT7;
  const size_t hp = (qt ? 7 : 12) * ss;  // bytes per pose on the host side
  const size_t o_q1 = alignUp(o_ok + n, 256);
  const size_t o_q2 = alignUp(o_q1 + (qt ? n * hp : 0), 256);
  const size_t total = alignUp(o_q2 + (qt ? n * hp : 0), 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  const size_t in1 = qt ? o_q1 : o_p1, in2 = qt ? o_q2 : o_p2;  // where the host poses land
  // Chunked three-stage pipeline: all H2D copies are queued up front on the copy-in
  // stream (one event per chunk); the compute stream waits per chunk, runs the
  // bucketed kernels, and the copy-out stream drains each chunk's results while later
  // chunks are still uploading / computing.  PCIe is full duplex, so with pinned
  // host buffers the call is bounded by the larger of the two copy directions.
  // Stage sizes: e.host_chunk queries while plenty is left, then halving down to e.host_taper -- the call ends one
  // stage's compute + copy-out after the last upload, so the last stages are kept short (profiles/r02_e2e_taper.txt).
  std::vector<size_t> c_begin, c_size;
  stageSizes(n, e.host_chunk, 0, e.host_taper, c_begin, c_size);
  const int n_chunks = int(c_size.size());
  rc = ensureChunkEvents(e, n_chunks);
  if (rc) return rc;
  const bool trace = getenv("FCLB_TRACE_HOST") != nullptr;  // per-stage timeline on stderr (tuning aid)
  std::vector<cudaEvent_t> tr;                              // start, then (in, done, out) per stage
  if (trace) {
    tr.resize(1 + 3 * size_t(n_chunks));
    for (auto& ev : tr) FCLB_CUDA(cudaEventCreate(&ev));
  }
  const char* h_pairs = reinterpret_cast<const char*>(pairs);
  const char* h_p1 = static_cast<const char*>(poses1);
  const char* h_p2 = static_cast<const char*>(poses2);
  // the staging arena may still be read by copy-out work of a previous call
  FCLB_CUDA(cudaStreamSynchronize(e.copy_out));
  if (trace) FCLB_CUDA(cudaEventRecord(tr[0], e.copy_in));
  for (int c = 0; c < n_chunks; c++) {
    const size_t b0 = c_begin[c], m = c_size[c];
    FCLB_CUDA(cudaMemcpyAsync(base + o_pairs + b0 * sizeof(fclb_pair), h_pairs + b0 * sizeof(fclb_pair),
                              m * sizeof(fclb_pair), cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaMemcpyAsync(base + in1 + b0 * hp, h_p1 + b0 * hp, m * hp, cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaMemcpyAsync(base + in2 + b0 * hp, h_p2 + b0 * hp, m * hp, cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaEventRecord(e.ev_in[c], e.copy_in));
    if (trace) FCLB_CUDA(cudaEventRecord(tr[1 + 3 * c], e.copy_in));
  }
  for (int c = 0; c < n_chunks; c++) {
    const size_t b0 = c_begin[c], m = c_size[c];
    FCLB_CUDA(cudaStreamWaitEvent(e.compute, e.ev_in[c], 0));
    if (qt) {
      rc = launchExpandQt7(e, scalar_type, base + o_q1 + b0 * hp, m, base + o_p1 + b0 * 12 * ss);
      if (!rc) rc = launchExpandQt7(e, scalar_type, base + o_q2 + b0 * hp, m, base + o_p2 + b0 * 12 * ss);
      if (rc) return rc;
    }
    rc = fclb_distance_batch_dev(shapes, reinterpret_cast<const fclb_pair*>(base + o_pairs) + b0, base + o_p1 + b0 * 12 * ss,
                                 base + o_p2 + b0 * 12 * ss, m, scalar_type, gjk_tol, gjk_max_iter,
                                 out_dist ? base + o_dist + b0 * ss : nullptr, out_p1 ? base + o_w1 + b0 * 3 * ss : nullptr,
                                 out_p2 ? base + o_w2 + b0 * 3 * ss : nullptr,
                                 out_ok ? reinterpret_cast<uint8_t*>(base + o_ok) + b0 : nullptr);
    if (rc) {  // drain the queued copies before the caller gets its buffers back
      cudaStreamSynchronize(e.copy_in);
      cudaStreamSynchronize(e.copy_out);
      return rc;
    }
    FCLB_CUDA(cudaEventRecord(e.ev_done[c], e.compute));
    if (trace) FCLB_CUDA(cudaEventRecord(tr[2 + 3 * c], e.compute));
    FCLB_CUDA(cudaStreamWaitEvent(e.copy_out, e.ev_done[c], 0));
    if (out_dist)
      FCLB_CUDA(cudaMemcpyAsync(static_cast<char*>(out_dist) + b0 * ss, base + o_dist + b0 * ss, m * ss,
                                cudaMemcpyDeviceToHost, e.copy_out));
    if (out_p1)
      FCLB_CUDA(cudaMemcpyAsync(static_cast<char*>(out_p1) + b0 * 3 * ss, base + o_w1 + b0 * 3 * ss, m * 3 * ss,
                                cudaMemcpyDeviceToHost, e.copy_out));
    if (out_p2)
      FCLB_CUDA(cudaMemcpyAsync(static_cast<char*>(out_p2) + b0 * 3 * ss, base + o_w2 + b0 * 3 * ss, m * 3 * ss,
                                cudaMemcpyDeviceToHost, e.copy_out));
    if (out_ok) FCLB_CUDA(cudaMemcpyAsync(out_ok + b0, base + o_ok + b0, m, cudaMemcpyDeviceToHost, e.copy_out));
    if (trace) FCLB_CUDA(cudaEventRecord(tr[3 + 3 * c], e.copy_out));
  }
  FCLB_CUDA(cudaStreamSynchronize(e.copy_out));
  if (trace) {
    for (int c = 0; c < n_chunks; c++) {
      float t_in = 0.f, t_done = 0.f, t_out = 0.f;
      cudaEventElapsedTime(&t_in, tr[0], tr[1 + 3 * c]);
      cudaEventElapsedTime(&t_done, tr[0], tr[2 + 3 * c]);
      cudaEventElapsedTime(&t_out, tr[0], tr[3 + 3 * c]);
      fprintf(stderr, "fclb trace: stage %d  %zu queries  uploaded %.3f ms  computed %.3f ms  copied out %.3f ms\n", c,
              c_size[c], t_in, t_done, t_out);
    }
    for (auto& ev : tr) cudaEventDestroy(ev);
  }
  return FCLB_OK;
}
static int distance_batch_host_any(int pose_format, fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                   size_t n, int scalar_type, double gjk_tol, uint32_t gjk_max_iter, void* out_dist, void* out_p1,
                                   void* out_p2, uint8_t* out_ok) {
  if (engineCount() <= 1)
    return distance_batch_host_fmt(pose_format, shapes, pairs, poses1, poses2, n, scalar_type, gjk_tol, gjk_max_iter, out_dist, out_p1,
                                   out_p2, out_ok);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t hp = (pose_format == FCLB_POSE_QT7 ? 7 : 12) * ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) {
    return distance_batch_host_fmt(pose_format, shapes, offT(pairs, b), offPtr(poses1, b * hp), offPtr(poses2, b * hp), m_, scalar_type,
                                   gjk_tol, gjk_max_iter, offPtr(out_dist, b * ss), offPtr(out_p1, b * 3 * ss),
                                   offPtr(out_p2, b * 3 * ss), offT(out_ok, b));
  });
}
int fclb_distance_batch_host(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                             size_t n, int scalar_type, double gjk_tol, uint32_t gjk_max_iter, void* out_dist,
                             void* out_p1, void* out_p2, uint8_t* out_ok) {
  return distance_batch_host_any(FCLB_POSE_RT12, shapes, pairs, poses1, poses2, n, scalar_type, gjk_tol, gjk_max_iter, out_dist, out_p1,
                                 out_p2, out_ok);
}
int fclb_distance_batch_qt_host(fclb_handle shapes, const fclb_pair* pairs, const void* qt_poses1, const void* qt_poses2,
                                size_t n, int scalar_type, double gjk_tol, uint32_t gjk_max_iter, void* out_dist,
                                void* out_p1, void* out_p2, uint8_t* out_ok) {
  return distance_batch_host_any(FCLB_POSE_QT7, shapes, pairs, qt_poses1, qt_poses2, n, scalar_type, gjk_tol, gjk_max_iter, out_dist,
                                 out_p1, out_p2, out_ok);
}
int fclb_expand_poses_dev(const void* qt_poses, size_t n, int scalar_type, void* out_poses12) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!qt_poses || !out_poses12) return fail(FCLB_ERR_BAD_ARG, "fclb_expand_poses_dev: null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  rc = launchExpandQt7(e, scalar_type, qt_poses, n, out_poses12);
  if (rc) return rc;
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}

// (collide / gjk_epa entry points: fclb_collide_api.cu)

uint64_t fclb_launch_count(void) { return eng().launches.load(); }
double fclb_last_kernel_ms(void) { return eng().last_ms; }
double fclb_last_call_ms(void) { return eng().last_call_ms; }
int fclb_last_launches(int* kinds, uint64_t* counts, double* ms, int cap) {
  Engine& e = eng();
  const int n = e.n_rec < cap ? e.n_rec : cap;
  for (int i = 0; i < n; i++) {
    if (kinds) kinds[i] = e.rec_kind[i];
    if (counts) counts[i] = e.rec_count[i];
    if (ms) ms[i] = double(e.rec_ms[i]);
  }
  return e.n_rec;
}
void* fclb_stream(void) {
  if (ensureInit()) return nullptr;
  return static_cast<void*>(eng().compute);
}

}  // extern "C"
