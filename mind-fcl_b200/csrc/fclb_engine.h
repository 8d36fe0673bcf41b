// fclb_engine.h -- engine state shared by the C-ABI translation units
// (fclb_engine.cu, fclb_collide_api.cu, fclb_bvh_api.cu).  Internal.
#pragma once
#include <atomic>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "fclb_stages.h"

#include "fclb_internal.h"

namespace fclb {

int fail(int code, const std::string& msg);
#define FCLB_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return ::fclb::fail(FCLB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));  \
  } while (0)

struct ConvexHost {
  // per scalar type device arrays
  void* d_verts[2] = {nullptr, nullptr};
  int* d_nbr = nullptr;
  void* d_vinfo = nullptr;  // int2 per vertex: (first neighbour offset, count)
  int n_verts = 0;
  std::vector<double> h_verts;  // 3 per vertex, as uploaded
  int walk = 0;
  int seed[2][6];
  double interior[2][3];
};

struct ShapeTable {
  void* d_shapes[2] = {nullptr, nullptr};  // ShapeD<float>[], ShapeD<double>[]
  void* d_bound[2] = {nullptr, nullptr};   // BoundD<float>[], BoundD<double>[] (fclb_bound.h)
  void* d_local[2] = {nullptr, nullptr};   // LocalAabbD<float>[], LocalAabbD<double>[] (fclb_bound.h)
  std::vector<fclb_shape> host;
  uint32_t n = 0;
  uint64_t convex_epoch = 0;
};

constexpr int kMaxDevices = 16;

// One engine per GPU.  A process either binds ONE GPU (fclb_init: one process per GPU, the torchrun layout) or
// enumerates several (fclb_init_devices: one process, the batch of a *_host call sharded by query index across the
// engines, geometry replicated on every device).  The calling thread's current engine is thread-local.
struct Engine {
  std::recursive_mutex mu;
  bool ready = false;
  int device = -1;
  int slot = 0;
  int sms = 148;
  cudaStream_t compute = nullptr, copy_in = nullptr, copy_out = nullptr;
  cudaStream_t aux = nullptr;                     // early EPA tier-2 consumers run here, beside tier 1 on `compute`
  cudaEvent_t ev_aux0 = nullptr, ev_aux1 = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<ConvexHost> convex;
  void* d_convex_tab[2] = {nullptr, nullptr};  // ConvexD<S>[]
  uint64_t convex_epoch = 0;
  std::map<fclb_handle, ShapeTable*> tables;
  // scratch for bucketing
  uint32_t* d_perm = nullptr;
  uint8_t* d_kind = nullptr;
  size_t scratch_cap = 0;
  uint32_t* d_hist = nullptr;  // kNumKinds counters + kNumKinds cursors
  uint32_t* h_hist = nullptr;  // pinned, mapped: scanKernel stores the histogram here itself
  uint32_t* h_hist_dev = nullptr;  // the device's pointer to h_hist
  // staging for host entry points
  void* d_stage = nullptr;
  size_t stage_cap = 0;
  size_t host_chunk = size_t(1) << 21;  // queries per pipeline stage of the *_host entry points
  size_t host_head = size_t(1) << 16;   // first stage of the compute-bound *_host pipelines; the stages double up to their chunk (0: equal stages)
  size_t host_taper = size_t(1) << 19;  // shortest stage of the tapered tail of fclb_distance_batch_*host (0: equal stages)
  std::vector<cudaEvent_t> ev_in, ev_done;
  std::atomic<uint64_t> launches{0};
  double last_ms = 0.0;       // kernels of the last call (bucketing excluded)
  double last_call_ms = 0.0;  // whole device side of the last call (bucketing included)
  // per-launch CUDA-event timing of the last call
  static constexpr int kMaxRec = 2 * kNumKinds;
  cudaEvent_t rec_ev[kMaxRec + 1] = {};
  int rec_kind[kMaxRec] = {};
  uint64_t rec_count[kMaxRec] = {};
  float rec_ms[kMaxRec] = {};
  int n_rec = 0;
  cudaEvent_t ev_call0 = nullptr;
};
Engine& eng();            // engine of the calling thread's current device slot
int engineCount();        // engines created by fclb_init / fclb_init_devices (>= 1 once initialised)
int currentSlot();
int setSlot(int slot);    // thread-local: later calls of this thread go to engine `slot` (cudaSetDevice included)
const std::string& lastErrorString();
void setLastErrorString(const std::string& s);

// process-global scratch that must exist once per device
template <typename T>
struct PerDevice {
  T v[kMaxDevices];
  T& get() { return v[currentSlot()]; }
};

// run fn() once per engine (geometry uploads / releases: every device holds a replica under the same handle)
void beginReplicas();
void nextReplica();
void endReplicas();
template <typename Fn>
int forEachDevice(Fn&& fn) {
  const int n = engineCount();
  if (n <= 1) return fn();
  const int home = currentSlot();
  int rc = FCLB_OK;
  beginReplicas();
  for (int s = 0; s < n && rc == FCLB_OK; s++) {
    rc = setSlot(s);
    if (rc == FCLB_OK) rc = fn();
    nextReplica();  // the other replicas are filed under the first one's handle
  }
  endReplicas();
  setSlot(home);
  return rc;
}
inline const void* offPtr(const void* p, size_t bytes) { return p ? static_cast<const char*>(p) + bytes : nullptr; }
inline void* offPtr(void* p, size_t bytes) { return p ? static_cast<char*>(p) + bytes : nullptr; }
template <typename T>
inline T* offT(T* p, size_t elems) { return p ? p + elems : nullptr; }

// shard [0, n) by contiguous query range over the engines, one host thread per device; fn(begin, count) runs the
// single-device implementation on its slice.  Returns the first non-zero code (message copied to the caller's thread).
int shardOverDevices(size_t n, const std::function<int(size_t, size_t)>& fn);

// Handles are process-wide: the replicas of one geometry on every device share one handle value.
fclb_handle newHandle();

int ensureInit();
ShapeTable* findTable(Engine& e, fclb_handle h);
int ensureStage(Engine& e, size_t bytes);
int ensureChunkEvents(Engine& e, int n);
// cudaMemcpy from PAGEABLE host memory returns once the source is staged; the DMA into device memory may still be in
// flight (CUDA runtime, "API synchronization behavior").  The engine's streams are non-blocking, so a kernel launched on
// them right after an upload is not ordered behind that DMA: the first warps of the first query batch could read a
// table that has not landed yet (seen under compute-sanitizer, whose launch timing differs).  Uploads drain the
// legacy stream before they return.
inline cudaError_t uploadSync(void* dst, const void* src, size_t bytes) {
  const cudaError_t e_ = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  return e_ != cudaSuccess ? e_ : cudaStreamSynchronize(cudaStreamLegacy);
}
inline size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }
SolverParams solverParams(int scalar_type, double gjk_tol, uint32_t gjk_max_iter, double epa_tol, uint32_t epa_max_faces,
                          uint32_t epa_max_iter, bool collide_defaults);
// Bucket a device-resident batch by (type1,type2); see fclb_engine.cu.
template <typename S>
int bucketBatch(Engine& e, const ShapeTable* t, const fclb_pair* d_pairs, size_t n, uint32_t* counts,
                uint32_t* offsets, int* uniform_kind);

// leaf batch of the scene contact path: see fclb_collide_api.cu
int collideLeafBatch(Engine& e, void* d_table, const void* tris, const fclb_pair* pairs, const void* poses1,
                     const void* poses2, size_t m, int scalar_type, const fclb_request* req, void* contacts,
                     uint32_t* counts, uint32_t n_table);

}  // namespace fclb
