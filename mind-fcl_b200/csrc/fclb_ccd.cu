// fclb_ccd.cu -- C ABI of the translational continuous collision path (shape-shape): fcl::translational_ccd
// (reference narrowphase/continuous_collision-inl.h:21-36) over a batch.  Kernel: fclb_ccd.cuh.
#include "fclb_ccd.cuh"
#include "fclb_engine.h"

namespace fclb {

static int ccdDev(Engine& e, ShapeTable* t, const fclb_pair* pairs, const void* poses1, const void* poses2, const void* disp,
                  size_t n, int scalar_type, const fclb_ccd_request* req, uint8_t* hit, void* toc) {
  const int st = scalar_type == FCLB_F32 ? 0 : 1;
  CcdArgs a{};
  a.shapes = t->d_shapes[st];
  a.convex = e.d_convex_tab[st];
  a.local = t->d_local[st];
  a.pairs = pairs;
  a.poses1 = poses1;
  a.poses2 = poses2;
  a.disp = disp;
  a.n = n;
  a.request_type = int(req->request_type);
  a.zero_tol = req->zero_movement_tolerance > 0 ? req->zero_movement_tolerance : 1e-4;  // ccd_request.h:28-30
  a.gjk_tol = req->gjk_tolerance > 0 ? req->gjk_tolerance : 1e-6;
  a.max_iter = req->max_gjk_iterations > 0 ? req->max_gjk_iterations : 128;
  a.hit = hit;
  a.toc = toc;
  const int grid = int(std::min<size_t>((n + kBlock - 1) / kBlock, size_t(e.sms) * 8));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  if (st == 0)
    translationalCcdKernel<float><<<grid, kBlock, 0, e.compute>>>(a);
  else
    translationalCcdKernel<double><<<grid, kBlock, 0, e.compute>>>(a);
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  e.launches += 1;
  FCLB_CUDA(cudaGetLastError());
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -6;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

static int ccdCheck(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2, const void* disp,
                    size_t n, int scalar_type, const fclb_ccd_request* req, uint8_t* hit, ShapeTable** t) {
  Engine& e = eng();
  *t = findTable(e, shapes);
  if (!*t) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_batch: unknown shape table handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!req || req->request_type > 2) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_batch: bad request");
  if (n && (!pairs || !poses1 || !poses2 || !disp || !hit)) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_batch: null array");
  if (n > 0xffffffffull) return fail(FCLB_ERR_CAPACITY, "batch larger than 2^32-1 queries: split it");
  return FCLB_OK;
}

}  // namespace fclb

using namespace fclb;

extern "C" {

int fclb_translational_ccd_batch_dev(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                     const void* displacements, size_t n, int scalar_type, const fclb_ccd_request* req,
                                     uint8_t* out_hit, void* out_toc) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  ShapeTable* t = nullptr;
  rc = ccdCheck(shapes, pairs, poses1, poses2, displacements, n, scalar_type, req, out_hit, &t);
  if (rc || n == 0) return rc;
  return ccdDev(e, t, pairs, poses1, poses2, displacements, n, scalar_type, req, out_hit, out_toc);
}

static int translational_ccd_batch_host_one(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                            const void* displacements, size_t n, int scalar_type, const fclb_ccd_request* req,
                                            uint8_t* out_hit, void* out_toc) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  ShapeTable* t = nullptr;
  rc = ccdCheck(shapes, pairs, poses1, poses2, displacements, n, scalar_type, req, out_hit, &t);
  if (rc || n == 0) return rc;
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_pairs = 0;
  const size_t o_p1 = alignUp(o_pairs + n * sizeof(fclb_pair), 256);
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_d = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_hit = alignUp(o_d + n * 4 * ss, 256);
  const size_t o_toc = alignUp(o_hit + n, 256);
  const size_t total = alignUp(o_toc + n * 2 * ss, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_pairs, pairs, n * sizeof(fclb_pair), cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses1, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses2, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_d, displacements, n * 4 * ss, cudaMemcpyHostToDevice, e.compute));
  rc = ccdDev(e, t, reinterpret_cast<const fclb_pair*>(base + o_pairs), base + o_p1, base + o_p2, base + o_d, n, scalar_type, req,
              reinterpret_cast<uint8_t*>(base + o_hit), out_toc ? base + o_toc : nullptr);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_hit, base + o_hit, n, cudaMemcpyDeviceToHost, e.compute));
  if (out_toc) FCLB_CUDA(cudaMemcpyAsync(out_toc, base + o_toc, n * 2 * ss, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}
int fclb_translational_ccd_batch_host(fclb_handle shapes, const fclb_pair* pairs, const void* poses1, const void* poses2,
                                      const void* displacements, size_t n, int scalar_type, const fclb_ccd_request* req,
                                      uint8_t* out_hit, void* out_toc) {
  if (engineCount() <= 1)
    return translational_ccd_batch_host_one(shapes, pairs, poses1, poses2, displacements, n, scalar_type, req, out_hit, out_toc);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  return shardOverDevices(n, [&](size_t b, size_t m_) {
    return translational_ccd_batch_host_one(shapes, offT(pairs, b), offPtr(poses1, b * 12 * ss), offPtr(poses2, b * 12 * ss),
                                            offPtr(displacements, b * 4 * ss), m_, scalar_type, req, offT(out_hit, b),
                                            offPtr(out_toc, b * 2 * ss));
  });
}

}  // extern "C"
