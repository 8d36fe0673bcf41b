// fclb_ccd_mesh.cu -- C ABI of translational continuous collision, shape vs mesh (kernels: fclb_ccd_mesh.cuh).
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <vector>

#include "fclb_ccd_mesh.cuh"
#include "fclb_engine.h"

namespace fclb {

__global__ void ccdIotaKernel(uint32_t* p, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
template <typename S>
__global__ void ccdFillKernel(long long* prim, S* toc, size_t n) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) {
    prim[i] = -1;
    if (toc) {
      toc[2 * i] = S(-1);
      toc[2 * i + 1] = S(-1);
    }
  }
}

// position of every triangle's leaf in the reference's walk: children pushed left then right, the right one popped
// first (bvh_ccd_solver-inl.h:186-192)
static int ensureDfsRank(BvhDev* m) {
  if (m->d_dfs_rank) return FCLB_OK;
  if (int(m->h_child.size()) != m->n_nodes) return fail(FCLB_ERR_BAD_ARG, "mesh CCD: the BVH has no host copy of its child links");
  std::vector<int> rank(size_t(m->n_tris), 0), stack;
  stack.push_back(0);
  int next = 0;
  while (!stack.empty()) {
    const int id = stack.back();
    stack.pop_back();
    const int fc = m->h_child[size_t(id)];
    if (fc < 0) {
      const int tri = -(fc + 1);
      if (tri >= m->n_tris) return fail(FCLB_ERR_BAD_ARG, "mesh CCD: leaf names a triangle beyond the array");
      rank[size_t(tri)] = next++;
    } else {
      stack.push_back(fc);
      stack.push_back(fc + 1);
    }
  }
  FCLB_CUDA(cudaMalloc(&m->d_dfs_rank, rank.size() * sizeof(int)));
  FCLB_CUDA(uploadSync(m->d_dfs_rank, rank.data(), rank.size() * sizeof(int)));
  return FCLB_OK;
}

struct CcdMeshScratch {
  std::vector<void*> ptrs;
  ~CcdMeshScratch() {
    #pragma unroll 1
    for (void* p : ptrs) cudaFree(p);
  }
  template <typename T>
  cudaError_t get(T** p, size_t count) {
    void* v = nullptr;
    const cudaError_t e = cudaMalloc(&v, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(v);
    *p = static_cast<T*>(v);
    return e;
  }
};

template <typename S>
static int ccdMeshDev(Engine& e, BvhDev* m, ShapeTable* t, const uint32_t* shape_ids, const void* poses_shape,
                      const void* poses_mesh, const void* disp, size_t n, const fclb_ccd_request* req, int mesh_moves,
                      uint32_t keep, uint32_t* counts, long long* prim, void* toc) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  int rc = ensureDfsRank(m);
  if (rc) return rc;
  CcdMeshScratch sc;
  uint32_t* d_count = nullptr;
  unsigned long long* d_work = nullptr;
  FCLB_CUDA(sc.get(&d_count, 2));
  FCLB_CUDA(sc.get(&d_work, 1));
  CcdMeshArgs a{};
  a.nodes = m->nodes;
  a.tris = m->tris;
  a.shapes = t->d_shapes[st];
  a.convex = e.d_convex_tab[st];
  a.shape_ids = shape_ids;
  a.poses_shape = poses_shape;
  a.poses_mesh = poses_mesh;
  a.disp = disp;
  a.n = n;
  a.mesh_moves = mesh_moves;
  a.request_type = int(req->request_type);
  a.zero_tol = req->zero_movement_tolerance > 0 ? req->zero_movement_tolerance : 1e-4;
  a.gjk_tol = req->gjk_tolerance > 0 ? req->gjk_tolerance : 1e-6;
  a.max_iter = req->max_gjk_iterations > 0 ? req->max_gjk_iterations : 128;
  a.cand_count = d_count;
  a.work_counter = d_work;
  a.dfs_rank = m->d_dfs_rank;
  const size_t smem = size_t(kCmWarps) * (kCmStackCap * (sizeof(int) + 2 * sizeof(S)) + 3 * kFitMaxPoints * sizeof(S));
  FCLB_CUDA(cudaFuncSetAttribute(ccdMeshTraverseKernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int grid = int(std::min<size_t>((n + kCmWarps - 1) / kCmWarps, size_t(e.sms) * 4));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  // (A) the walk, repeated with a larger list if the candidates did not fit
  size_t cap = std::min<size_t>(std::max<size_t>(n * 16, size_t(1) << 16), size_t(1) << 26);
  uint32_t h_count[2] = {0, 0};
  for (int attempt = 0; attempt < 3; attempt++) {
    FCLB_CUDA(sc.get(&a.cand_q, cap));
    FCLB_CUDA(sc.get(&a.cand_tri, cap));
    S* iv = nullptr;
    FCLB_CUDA(sc.get(&iv, 2 * cap));
    a.cand_iv = iv;
    a.cand_cap = uint32_t(cap);
    FCLB_CUDA(cudaMemsetAsync(d_count, 0, 2 * sizeof(uint32_t), e.compute));
    FCLB_CUDA(cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), e.compute));
    ccdMeshTraverseKernel<S><<<grid, kCmWarps * 32, smem, e.compute>>>(a);
    FCLB_CUDA(cudaGetLastError());
    e.launches += 1;
    FCLB_CUDA(cudaMemcpyAsync(h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    if (h_count[1]) return fail(FCLB_ERR_CAPACITY, "mesh CCD: tree deeper than the per-warp stack allows");
    if (h_count[0] <= cap) break;
    if (attempt == 2 || h_count[0] > (1u << 30)) return fail(FCLB_ERR_CAPACITY, "mesh CCD: too many candidate triangles: split the batch");
    cap = h_count[0];
  }
  const uint32_t n_cand = h_count[0];
  const int fill_grid = int((n * keep + 255) / 256);
  FCLB_CUDA(cudaMemsetAsync(counts, 0, n * sizeof(uint32_t), e.compute));
  if (n * keep) ccdFillKernel<S><<<fill_grid, 256, 0, e.compute>>>(prim, static_cast<S*>(toc), n * keep);
  if (n_cand) {
    // (B) the leaves
    unsigned long long *keys = nullptr, *keys_sorted = nullptr;
    uint32_t *order = nullptr, *order_sorted = nullptr;
    S* cand_toc = nullptr;
    FCLB_CUDA(sc.get(&keys, n_cand));
    FCLB_CUDA(sc.get(&keys_sorted, n_cand));
    FCLB_CUDA(sc.get(&order, n_cand));
    FCLB_CUDA(sc.get(&order_sorted, n_cand));
    FCLB_CUDA(sc.get(&cand_toc, 2 * size_t(n_cand)));
    a.keys = keys;
    a.cand_toc = cand_toc;
    const int lgrid = int(std::min<size_t>((n_cand + kBlock - 1) / kBlock, size_t(e.sms) * 8));
    ccdMeshLeafKernel<S><<<lgrid, kBlock, 0, e.compute>>>(a, n_cand);
    FCLB_CUDA(cudaGetLastError());
    ccdIotaKernel<<<(n_cand + 255) / 256, 256, 0, e.compute>>>(order, n_cand);
    // (C) order the hits by (query, rank in the reference's walk), keep the first max_contacts of every query
    size_t tmp_bytes = 0;
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_sorted, order, order_sorted, int(n_cand), 0, 64,
                                              e.compute));
    unsigned char* tmp = nullptr;
    FCLB_CUDA(sc.get(&tmp, tmp_bytes));
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_sorted, order, order_sorted, int(n_cand), 0, 64,
                                              e.compute));
    ccdMeshSelectKernel<S><<<(n_cand + 255) / 256, 256, 0, e.compute>>>(keys_sorted, order_sorted, n_cand, a.cand_tri, cand_toc,
                                                                        req->max_contacts ? req->max_contacts : 1u, keep, counts,
                                                                        prim, static_cast<S*>(toc));
    FCLB_CUDA(cudaGetLastError());
    e.launches += 5;
  }
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -7;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

template <typename S>
__global__ void ccdFillScalarKernel(S* p, size_t n, S v) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void ccdFillI64Kernel(long long* p, size_t n, long long v) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}

template <typename S>
static int ccdMeshPairDev(Engine& e, BvhDev* m1, BvhDev* m2, const void* poses1, const void* poses2, const void* disp, size_t n,
                          const fclb_ccd_request* req, uint32_t keep, uint32_t* counts, long long* prim, void* toc) {
  CcdMeshScratch sc;
  uint32_t* d_count = nullptr;
  unsigned long long* d_work = nullptr;
  FCLB_CUDA(sc.get(&d_count, 4));
  FCLB_CUDA(sc.get(&d_work, 1));
  CcdMeshPairArgs a{};
  a.nodes1 = m1->nodes;
  a.tris1 = m1->tris;
  a.nodes2 = m2->nodes;
  a.tris2 = m2->tris;
  a.poses1 = poses1;
  a.poses2 = poses2;
  a.disp = disp;
  a.n = n;
  a.request_type = int(req->request_type);
  a.zero_tol = req->zero_movement_tolerance > 0 ? req->zero_movement_tolerance : 1e-4;
  a.gjk_tol = req->gjk_tolerance > 0 ? req->gjk_tolerance : 1e-6;
  a.max_iter = req->max_gjk_iterations > 0 ? req->max_gjk_iterations : 128;
  a.cand_count = d_count;
  a.work_counter = d_work;
  const size_t smem = size_t(kCmWarps) * kCpStackCap * (8 + 8 + 2 * sizeof(S) + 1);
  FCLB_CUDA(cudaFuncSetAttribute(ccdMeshPairTraverseKernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int grid = int(std::min<size_t>((n + kCmWarps - 1) / kCmWarps, size_t(e.sms) * 4));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  size_t cap = std::min<size_t>(std::max<size_t>(n * 32, size_t(1) << 16), size_t(1) << 26);
  uint32_t h_count[4] = {0, 0, 0, 0};
  for (int attempt = 0; attempt < 3; attempt++) {
    FCLB_CUDA(sc.get(&a.cand_q, cap));
    FCLB_CUDA(sc.get(&a.cand_tri, cap));
    FCLB_CUDA(sc.get(&a.cand_path, cap));
    S* iv = nullptr;
    FCLB_CUDA(sc.get(&iv, 2 * cap));
    a.cand_iv = iv;
    a.cand_cap = uint32_t(cap);
    FCLB_CUDA(cudaMemsetAsync(d_count, 0, 4 * sizeof(uint32_t), e.compute));
    FCLB_CUDA(cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), e.compute));
    ccdMeshPairTraverseKernel<S><<<grid, kCmWarps * 32, smem, e.compute>>>(a);
    FCLB_CUDA(cudaGetLastError());
    e.launches += 1;
    FCLB_CUDA(cudaMemcpyAsync(h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    if (h_count[2]) return fail(FCLB_ERR_CAPACITY, "mesh-pair CCD: node pairs deeper than 64 descent steps");
    if (h_count[1]) return fail(FCLB_ERR_CAPACITY, "mesh-pair CCD: trees deeper than the per-warp stack allows");
    if (h_count[0] <= cap) break;
    if (attempt == 2 || h_count[0] > (1u << 30)) return fail(FCLB_ERR_CAPACITY, "mesh-pair CCD: too many candidate triangle pairs: split the batch");
    cap = h_count[0];
  }
  const uint32_t n_cand = h_count[0];
  FCLB_CUDA(cudaMemsetAsync(counts, 0, n * sizeof(uint32_t), e.compute));
  if (n * keep) {
    ccdFillI64Kernel<<<int((n * keep * 2 + 255) / 256), 256, 0, e.compute>>>(prim, n * keep * 2, -1ll);
    if (toc) ccdFillScalarKernel<S><<<int((n * keep * 2 + 255) / 256), 256, 0, e.compute>>>(static_cast<S*>(toc), n * keep * 2, S(-1));
  }
  if (n_cand) {
    unsigned long long* path_sorted = nullptr;
    uint32_t *order = nullptr, *order1 = nullptr, *order2 = nullptr, *qkey = nullptr, *qkey1 = nullptr, *qkey2 = nullptr;
    S* cand_toc = nullptr;
    FCLB_CUDA(sc.get(&path_sorted, n_cand));
    FCLB_CUDA(sc.get(&order, n_cand));
    FCLB_CUDA(sc.get(&order1, n_cand));
    FCLB_CUDA(sc.get(&order2, n_cand));
    FCLB_CUDA(sc.get(&qkey, n_cand));
    FCLB_CUDA(sc.get(&qkey1, n_cand));
    FCLB_CUDA(sc.get(&qkey2, n_cand));
    FCLB_CUDA(sc.get(&cand_toc, 2 * size_t(n_cand)));
    a.qkey = qkey;
    a.cand_toc = cand_toc;
    const int lgrid = int(std::min<size_t>((n_cand + kBlock - 1) / kBlock, size_t(e.sms) * 8));
    ccdMeshPairLeafKernel<S><<<lgrid, kBlock, 0, e.compute>>>(a, n_cand);
    FCLB_CUDA(cudaGetLastError());
    const int g256 = int((n_cand + 255) / 256);
    ccdIotaKernel<<<g256, 256, 0, e.compute>>>(order, n_cand);
    // stable sorts: by path (the reference's visiting order), then by query
    size_t b1 = 0, b2 = 0;
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, a.cand_path, path_sorted, order, order1, int(n_cand), 0, 64, e.compute));
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b2, qkey1, qkey2, order1, order2, int(n_cand), 0, 32, e.compute));
    unsigned char* tmp = nullptr;
    FCLB_CUDA(sc.get(&tmp, std::max(b1, b2)));
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b1, a.cand_path, path_sorted, order, order1, int(n_cand), 0, 64, e.compute));
    ccdGatherKeyKernel<<<g256, 256, 0, e.compute>>>(qkey, order1, n_cand, qkey1);
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b2, qkey1, qkey2, order1, order2, int(n_cand), 0, 32, e.compute));
    ccdMeshPairSelectKernel<S><<<g256, 256, 0, e.compute>>>(qkey2, order2, n_cand, a.cand_tri, cand_toc,
                                                            req->max_contacts ? req->max_contacts : 1u, keep, counts, prim,
                                                            static_cast<S*>(toc));
    FCLB_CUDA(cudaGetLastError());
    e.launches += 6;
  }
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -8;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

}  // namespace fclb

using namespace fclb;

extern "C" {

int fclb_translational_ccd_mesh_batch_dev(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids, const void* poses_shape,
                                          const void* poses_mesh, const void* displacements, size_t n, int scalar_type,
                                          const fclb_ccd_request* req, int mesh_moves, uint32_t max_keep, uint32_t* out_counts,
                                          int64_t* out_prim, void* out_toc) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = bvhTable().find(bvh);
  if (it == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_mesh_batch: unknown BVH handle");
  ShapeTable* t = findTable(e, shapes);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_mesh_batch: unknown shape table handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (it->second->scalar_type != scalar_type) return fail(FCLB_ERR_BAD_ARG, "BVH was uploaded for a different scalar type");
  if (!req || req->request_type > 2) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_mesh_batch: bad request");
  if (n == 0) return FCLB_OK;
  if (n > 0xfffffffeull) return fail(FCLB_ERR_CAPACITY, "batch larger than 2^32-2 queries: split it");
  if (!shape_ids || !poses_shape || !poses_mesh || !displacements || !out_counts || (max_keep && !out_prim))
    return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_mesh_batch: null array");
  // computeBV<OBB, Convex> fits its box with fit1 / fit2 / fit3 / fit6 for hulls of exactly 1, 2, 3 or 6 vertices
  // (math/bv/utility-inl.h:468-490); only the general covariance fit is on the device
  #pragma unroll 1
  for (uint32_t i = 0; i < t->n; i++)
    if (t->host[i].type == FCLB_CONVEX && t->host[i].geom < e.convex.size()) {
      const int nv = e.convex[t->host[i].geom].n_verts;
      if (nv == 1 || nv == 2 || nv == 3 || nv == 6)
        return fail(FCLB_ERR_UNSUPPORTED, "mesh CCD: Convex shapes with 1, 2, 3 or 6 vertices are not supported");
    }
  if (scalar_type == FCLB_F32)
    return ccdMeshDev<float>(e, it->second, t, shape_ids, poses_shape, poses_mesh, displacements, n, req, mesh_moves, max_keep,
                             out_counts, reinterpret_cast<long long*>(out_prim), out_toc);
  return ccdMeshDev<double>(e, it->second, t, shape_ids, poses_shape, poses_mesh, displacements, n, req, mesh_moves, max_keep,
                            out_counts, reinterpret_cast<long long*>(out_prim), out_toc);
}

static int translational_ccd_mesh_batch_host_one(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids,
                                                 const void* poses_shape, const void* poses_mesh, const void* displacements,
                                                 size_t n, int scalar_type, const fclb_ccd_request* req, int mesh_moves,
                                                 uint32_t max_keep, uint32_t* out_counts, int64_t* out_prim, void* out_toc) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_shape || !poses_mesh || !displacements || !out_counts)
    return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_mesh_batch: null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  {
    ShapeTable* t = findTable(e, shapes);
    if (!t) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_mesh_batch: unknown shape table handle");
    #pragma unroll 1
    for (size_t q = 0; q < n; q++)
      if (shape_ids[q] >= t->n) return fail(FCLB_ERR_BAD_ARG, "shape id out of range");
  }
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_ids = 0;
  const size_t o_p1 = alignUp(o_ids + n * 4, 256);
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_d = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_d + n * 4 * ss, 256);
  const size_t o_prim = alignUp(o_cnt + n * 4, 256);
  const size_t o_toc = alignUp(o_prim + n * max_keep * 8, 256);
  const size_t total = alignUp(o_toc + n * max_keep * 2 * ss, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_ids, shape_ids, n * 4, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses_shape, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses_mesh, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_d, displacements, n * 4 * ss, cudaMemcpyHostToDevice, e.compute));
  rc = fclb_translational_ccd_mesh_batch_dev(bvh, shapes, reinterpret_cast<const uint32_t*>(base + o_ids), base + o_p1, base + o_p2,
                                             base + o_d, n, scalar_type, req, mesh_moves, max_keep,
                                             reinterpret_cast<uint32_t*>(base + o_cnt), reinterpret_cast<int64_t*>(base + o_prim),
                                             out_toc ? base + o_toc : nullptr);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_counts, base + o_cnt, n * 4, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_prim) FCLB_CUDA(cudaMemcpyAsync(out_prim, base + o_prim, n * max_keep * 8, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_toc) FCLB_CUDA(cudaMemcpyAsync(out_toc, base + o_toc, n * max_keep * 2 * ss, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}

int fclb_translational_ccd_mesh_batch_host(fclb_handle bvh, fclb_handle shapes, const uint32_t* shape_ids, const void* poses_shape,
                                           const void* poses_mesh, const void* displacements, size_t n, int scalar_type,
                                           const fclb_ccd_request* req, int mesh_moves, uint32_t max_keep, uint32_t* out_counts,
                                           int64_t* out_prim, void* out_toc) {
  if (engineCount() <= 1)
    return translational_ccd_mesh_batch_host_one(bvh, shapes, shape_ids, poses_shape, poses_mesh, displacements, n, scalar_type, req,
                                                 mesh_moves, max_keep, out_counts, out_prim, out_toc);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  return shardOverDevices(n, [&](size_t b, size_t m_) {
    return translational_ccd_mesh_batch_host_one(bvh, shapes, offT(shape_ids, b), offPtr(poses_shape, b * 12 * ss),
                                                 offPtr(poses_mesh, b * 12 * ss), offPtr(displacements, b * 4 * ss), m_, scalar_type,
                                                 req, mesh_moves, max_keep, offT(out_counts, b), offT(out_prim, b * max_keep),
                                                 offPtr(out_toc, b * max_keep * 2 * ss));
  });
}

int fclb_translational_ccd_mesh_pair_batch_dev(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2,
                                               const void* displacements, size_t n, int scalar_type, const fclb_ccd_request* req,
                                               uint32_t max_keep, uint32_t* out_counts, int64_t* out_prim, void* out_toc) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto i1 = bvhTable().find(bvh1), i2 = bvhTable().find(bvh2);
  if (i1 == bvhTable().end() || i2 == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_mesh_pair_batch: unknown BVH handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (i1->second->scalar_type != scalar_type || i2->second->scalar_type != scalar_type)
    return fail(FCLB_ERR_BAD_ARG, "BVH was uploaded for a different scalar type");
  if (!req || req->request_type > 2) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_mesh_pair_batch: bad request");
  if (n == 0) return FCLB_OK;
  if (n > 0xfffffffeull) return fail(FCLB_ERR_CAPACITY, "batch larger than 2^32-2 queries: split it");
  if (!poses1 || !poses2 || !displacements || !out_counts || (max_keep && !out_prim))
    return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_mesh_pair_batch: null array");
  if (scalar_type == FCLB_F32)
    return ccdMeshPairDev<float>(e, i1->second, i2->second, poses1, poses2, displacements, n, req, max_keep, out_counts,
                                 reinterpret_cast<long long*>(out_prim), out_toc);
  return ccdMeshPairDev<double>(e, i1->second, i2->second, poses1, poses2, displacements, n, req, max_keep, out_counts,
                                reinterpret_cast<long long*>(out_prim), out_toc);
}

static int translational_ccd_mesh_pair_batch_host_one(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2,
                                                      const void* displacements, size_t n, int scalar_type,
                                                      const fclb_ccd_request* req, uint32_t max_keep, uint32_t* out_counts,
                                                      int64_t* out_prim, void* out_toc) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!poses1 || !poses2 || !displacements || !out_counts) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_mesh_pair_batch: null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_p1 = 0;
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_d = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_d + n * 4 * ss, 256);
  const size_t o_prim = alignUp(o_cnt + n * 4, 256);
  const size_t o_toc = alignUp(o_prim + n * max_keep * 16, 256);
  const size_t total = alignUp(o_toc + n * max_keep * 2 * ss, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses1, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses2, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_d, displacements, n * 4 * ss, cudaMemcpyHostToDevice, e.compute));
  rc = fclb_translational_ccd_mesh_pair_batch_dev(bvh1, bvh2, base + o_p1, base + o_p2, base + o_d, n, scalar_type, req, max_keep,
                                                  reinterpret_cast<uint32_t*>(base + o_cnt), reinterpret_cast<int64_t*>(base + o_prim),
                                                  out_toc ? base + o_toc : nullptr);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_counts, base + o_cnt, n * 4, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_prim) FCLB_CUDA(cudaMemcpyAsync(out_prim, base + o_prim, n * max_keep * 16, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_toc) FCLB_CUDA(cudaMemcpyAsync(out_toc, base + o_toc, n * max_keep * 2 * ss, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}

int fclb_translational_ccd_mesh_pair_batch_host(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2,
                                                const void* displacements, size_t n, int scalar_type, const fclb_ccd_request* req,
                                                uint32_t max_keep, uint32_t* out_counts, int64_t* out_prim, void* out_toc) {
  if (engineCount() <= 1)
    return translational_ccd_mesh_pair_batch_host_one(bvh1, bvh2, poses1, poses2, displacements, n, scalar_type, req, max_keep,
                                                      out_counts, out_prim, out_toc);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  return shardOverDevices(n, [&](size_t b, size_t m_) {
    return translational_ccd_mesh_pair_batch_host_one(bvh1, bvh2, offPtr(poses1, b * 12 * ss), offPtr(poses2, b * 12 * ss),
                                                      offPtr(displacements, b * 4 * ss), m_, scalar_type, req, max_keep,
                                                      offT(out_counts, b), offT(out_prim, b * max_keep * 2),
                                                      offPtr(out_toc, b * max_keep * 2 * ss));
  });
}

}  // extern "C"
