// fclb_ccd_scene.cu -- translational continuous collision, shape vs heightmap / octree: kernels + C ABI.
//
// Reference: fcl::translational_ccd(shape, tf1, displacement, HeightMapCollisionGeometry | Octree2CollisionGeometry, tf2, ...)
//   TranslationalDisplacementHeightMapSolver::runShapeHeightMap (detail/ccd/heightmap_ccd_solver-inl.h:8-112)
//   TranslationalDisplacementOctreeSolver::runShapeOctree       (detail/ccd/octree2_ccd_solver-inl.h:58-198)
// Both: the shape's OBB (computeBV<OBB, Shape>) becomes an AABB in its own box frame, the scene's node boxes are AABBs in
// the scene frame, and FixedOrientationBoxPairTranslationalCCD (fixed relative rotation) culls a node or narrows the
// time-of-collision interval its children inherit; a terminal box (bottom-layer pixel; fully occupied octree node; voxel
// of a partial leaf -- those without a box test of their own) runs RunShapePair<Shape, Box> with the request's type.
// The scene-first entries (RunHeightMapShape :138-166, RunOctreeShape) move the negated displacement into the shape's
// frame and run the same walk.
//
// Here, as for meshes (fclb_ccd_mesh.cuh): (A) one warp per query walks the hierarchy and appends every terminal box
// that survives, with its PATH -- the digits of the reference's visiting order, most significant first: heightmap roots
// are pushed row by row and popped in reverse, the four / eight children are pushed in index order and popped in
// reverse, the voxels of a partial octree leaf are visited in index order; (B) one thread per candidate runs the shape
// pair; (C) two stable radix sorts (path, query) put the hits in the reference's order and the first max_contacts of
// every query are written.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <vector>

#include "fclb_ccd_mesh.cuh"
#include "fclb_engine.h"
#include "fclb_scene_pair_impl.cuh"

namespace fclb {

struct CcdSceneArgs {
  HmView hm;
  OctView oct;
  const void* shapes;
  const void* convex;
  const void* local;    // LocalAabbD<S>[]
  const uint32_t* shape_ids;
  const void* poses_shape;
  const void* poses_scene;
  const void* disp;
  size_t n;
  int scene_moves;
  int request_type;
  double zero_tol, gjk_tol;
  int max_iter;
  uint32_t* cand_count;  // [0] appended, [1] stack overflows
  uint32_t cand_cap;
  uint32_t* cand_q;
  long long* cand_code;
  void* cand_box;       // 6 S per candidate
  unsigned long long* cand_path;
  unsigned long long* work_counter;
  uint32_t* qkey;
  void* cand_toc;
};

constexpr int kCsWarps = 4;
constexpr int kCsStackCap = 320;

template <typename S>
struct CsElem {
  BoxElem<S> box;
  S lo, hi;
  unsigned long long path;
  int shift;  // free bits left below the path's digits
};

template <typename S>
FCLB_DI V3<S> ccdDisplacementInShapeFrame(const S* disp, size_t q, const Pose<S>& tf_shape, const Pose<S>& tf_scene, int scene_moves) {
  V3<S> unit = mk<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
  if (scene_moves) unit = mulMtV(tf_shape.R, mulMV(tf_scene.R, -unit));
  return unit;
}

template <typename S, int KIND>
__global__ void __launch_bounds__(kCsWarps * 32) ccdSceneTraverseKernel(CcdSceneArgs a) {
  extern __shared__ __align__(16) unsigned char s_cs[];
  using Side = typename SideOf<S, KIND>::type;
  const Side side = SideOf<S, KIND>::make(a.hm, a.oct);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  CsElem<S>* stack = reinterpret_cast<CsElem<S>*>(s_cs) + size_t(warp) * kCsStackCap;
  S* fit_pts = reinterpret_cast<S*>(s_cs + size_t(kCsWarps) * kCsStackCap * sizeof(CsElem<S>)) + size_t(warp) * 3 * kFitMaxPoints;
  const S* __restrict__ disp = static_cast<const S*>(a.disp);
  const unsigned lt_mask = (1u << lane) - 1u;
  const S zero_tol = S(a.zero_tol);
  constexpr int kChildBits = KIND == FCLB_SCENE_HEIGHTMAP ? 2 : 3;
  constexpr int kMaxChildren = KIND == FCLB_SCENE_HEIGHTMAP ? 4 : 8;
  while (true) {
    unsigned long long q64 = 0;
    if (lane == 0) q64 = atomicAdd(a.work_counter, 1ull);
    q64 = __shfl_sync(0xffffffffu, q64, 0);
    if (q64 >= a.n) break;
    const size_t q = size_t(q64);
    const uint32_t sid = a.shape_ids[q];
    const ShapeInst<S> sh = bindShape(static_cast<const ShapeD<S>*>(a.shapes), static_cast<const ConvexD<S>*>(a.convex), sid);
    const Pose<S> tf_s = loadPose(static_cast<const S*>(a.poses_shape), q);
    const Pose<S> tf_g = loadPose(static_cast<const S*>(a.poses_scene), q);
    const V3<S> unit = ccdDisplacementInShapeFrame(disp, q, tf_s, tf_g, a.scene_moves);
    // initializeShapeFixedOrientationBoxTranslationalCCD (ccd_solver_utility-inl.h:8-35)
    Pose<S> tf_box;
    V3<S> b_ext;
    shapeObbForCcd(sh, tf_s, fit_pts, lane, tf_box.R, tf_box.t, b_ext);
    FixedCcd<S> f;
    {
      const Pose<S> rel = compose(inverse(tf_box), tf_g);
      f.R = rel.R;
      f.t = rel.t;
      f.unit = mulMtV(tf_box.R, mulMV(tf_s.R, unit));
      f.scalar = disp[4 * q + 3];
    }
    const V3<S> mn1 = -b_ext, mx1 = b_ext;

    // roots: pushed in index order, popped in reverse
    const int n_roots = side.numRoots();
    int root_bits = 0;
    while ((1 << root_bits) < n_roots) root_bits++;
    int sp = 0;
    bool overflow = n_roots > kCsStackCap - 8 * 32;
    if (!overflow) {
      #pragma unroll 1
      for (int base = 0; base < n_roots; base += 32) {
        const int i = base + lane;
        CsElem<S> e;
        bool ok = false;
        if (i < n_roots) {
          ok = side.root(i, e.box);
          e.lo = S(0.0);
          e.hi = S(1.0);
          e.shift = 64 - root_bits;
          e.path = root_bits ? ((unsigned long long)(n_roots - 1 - i) << e.shift) : 0ull;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) stack[sp + __popc(m & lt_mask)] = e;
        sp += __popc(m);
      }
    }
    __syncwarp();
    while (sp > 0 && !overflow) {
      int take = sp < 32 ? sp : 32;
      if (sp + take * kMaxChildren > kCsStackCap) take = (kCsStackCap - sp) / kMaxChildren > 0 ? 1 : 0;
      if (take == 0) {
        overflow = true;
        break;
      }
      CsElem<S> e;
      if (lane < take) e = stack[sp - 1 - lane];
      sp -= take;
      __syncwarp();
      int n_push = 0, n_cand = 0;
      unsigned child_mask = 0;
      TocInterval<S> iv;
      iv.lo = iv.hi = S(0);
      bool voxels = false;
      if (lane < take) {
        iv.lo = e.lo;
        iv.hi = e.hi;
        const V3<S> mn2 = mk<S>(e.box.mn[0], e.box.mn[1], e.box.mn[2]), mx2 = mk<S>(e.box.mx[0], e.box.mx[1], e.box.mx[2]);
        if (!fixedCcdDisjoint(f, mn1, mx1, mn2, mx2, iv, zero_tol)) {
          if (e.box.meta & 1u) {
            n_cand = 1;  // terminal: the box itself
          } else {
            child_mask = side.childMask(e.box);
            voxels = KIND == FCLB_SCENE_OCTREE && (e.box.meta & 2u);  // partial leaf: its voxels, without a box test of their own
            if (voxels)
              n_cand = __popc(child_mask);
            else if (e.shift < kChildBits)
              overflow = true;
            else
              n_push = __popc(child_mask);
          }
        }
      }
      if (__any_sync(0xffffffffu, overflow)) {
        overflow = true;
        break;
      }
      // warp scans of the push / candidate counts
      int push_off = n_push, cand_off = n_cand;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int p = __shfl_up_sync(0xffffffffu, push_off, o), c = __shfl_up_sync(0xffffffffu, cand_off, o);
        if (lane >= o) {
          push_off += p;
          cand_off += c;
        }
      }
      const int total_push = __shfl_sync(0xffffffffu, push_off, 31), total_cand = __shfl_sync(0xffffffffu, cand_off, 31);
      push_off -= n_push;
      cand_off -= n_cand;
      if (n_push) {
        int k = 0;
        const int shift = e.shift - kChildBits;
        for (int c = 0; c < kMaxChildren; c++) {
          if (!(child_mask & (1u << c))) continue;
          CsElem<S> ch;
          ch.box = side.child(e.box, c);
          ch.lo = iv.lo;
          ch.hi = iv.hi;
          ch.shift = shift;
          ch.path = e.path | ((unsigned long long)(kMaxChildren - 1 - c) << shift);  // the last child pushed is visited first
          stack[sp + push_off + k] = ch;
          k++;
        }
      }
      sp += total_push;
      if (total_cand) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.cand_count, uint32_t(total_cand));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (n_cand) {
          uint32_t slot = base + uint32_t(cand_off);
          auto emit = [&](const BoxElem<S>& bx, unsigned long long path) {
            if (slot < a.cand_cap) {
              a.cand_q[slot] = uint32_t(q);
              a.cand_code[slot] = Side::code(bx);
              S* o = static_cast<S*>(a.cand_box) + 6 * size_t(slot);
#pragma unroll
              for (int j = 0; j < 3; j++) {
                o[j] = bx.mn[j];
                o[3 + j] = bx.mx[j];
              }
              a.cand_path[slot] = path;
            }
            slot++;
          };
          if (!voxels) {
            emit(e.box, e.path);
          } else {  // voxels of a partial leaf, in index order; they inherit the LEAF's place in the walk
            const int shift = e.shift >= 3 ? e.shift - 3 : 0;
            for (int c = 0; c < 8; c++)
              if (child_mask & (1u << c)) emit(side.child(e.box, c), e.path | ((unsigned long long)c << shift));
          }
        }
      }
      __syncwarp();
    }
    if (overflow && lane == 0) atomicAdd(a.cand_count + 1, 1u);
    __syncwarp();
  }
}

// RunShapePair<Shape, Box> per candidate (heightmap_ccd_solver-inl.h:88-112, octree2_ccd_solver-inl.h shapeToBoxProcessLeafPair):
// constructBox(aabb, tf_scene) makes the box; the contact carries no external interval, so the request's type stands
template <typename S>
__global__ void __launch_bounds__(kBlock) ccdSceneLeafKernel(CcdSceneArgs a, uint32_t n_cand) {
  const S* __restrict__ disp = static_cast<const S*>(a.disp);
  const LocalAabbD<S>* __restrict__ local = static_cast<const LocalAabbD<S>*>(a.local);
  const S zero_tol = S(a.zero_tol), tol = S(a.gjk_tol);
  #pragma unroll 1
  for (size_t c = blockIdx.x * size_t(blockDim.x) + threadIdx.x; c < n_cand; c += size_t(gridDim.x) * blockDim.x) {
    const size_t q = a.cand_q[c];
    const uint32_t sid = a.shape_ids[q];
    const Pose<S> tf_s = loadPose(static_cast<const S*>(a.poses_shape), q);
    const Pose<S> tf_g = loadPose(static_cast<const S*>(a.poses_scene), q);
    const V3<S> unit = ccdDisplacementInShapeFrame(disp, q, tf_s, tf_g, a.scene_moves);
    const S* bx = static_cast<const S*>(a.cand_box) + 6 * c;
    const V3<S> mn = mk<S>(bx[0], bx[1], bx[2]), mx = mk<S>(bx[3], bx[4], bx[5]);
    // constructBox (geometry/shape/utility-inl.h:896-900): side = max - min, tf = tf_scene * Translation(center)
    ShapeInst<S> box;
    box.type = ST_BOX;
    box.cvx = nullptr;
    box.p0 = mx.x - mn.x;
    box.p1 = mx.y - mn.y;
    box.p2 = mx.z - mn.z;
    const V3<S> center = (mn + mx) * S(0.5);
    Pose<S> tf_b;
    tf_b.R = tf_g.R;
    tf_b.t = mulMV(tf_g.R, center) + tf_g.t;
    LocalAabbD<S> lb;  // Box::computeLocalAABB(): +- side / 2 around the origin
    lb.mx[0] = S(0.5) * box.p0; lb.mx[1] = S(0.5) * box.p1; lb.mx[2] = S(0.5) * box.p2;
    lb.mn[0] = -lb.mx[0]; lb.mn[1] = -lb.mx[1]; lb.mn[2] = -lb.mx[2];
    lb.center[0] = lb.center[1] = lb.center[2] = S(0);
    lb.radius = S(0);
    TocInterval<S> toc;
    const bool hit = ccdShapePairEval<S>(bindShape(static_cast<const ShapeD<S>*>(a.shapes), static_cast<const ConvexD<S>*>(a.convex), sid),
                                         local[sid], tf_s, box, lb, tf_b, unit, disp[4 * q + 3], a.request_type, zero_tol, tol,
                                         a.max_iter, toc);
    a.qkey[c] = hit ? uint32_t(q) : 0xffffffffu;
    S* o = static_cast<S*>(a.cand_toc) + 2 * c;
    const bool valid = toc.lo >= 0 && toc.hi >= 0;  // writeToContact(contact, toc): otherwise the (invalid) external interval
    o[0] = valid ? toc.lo : S(-1.0);
    o[1] = valid ? toc.hi : S(-1.0);
  }
}

template <typename S>
__global__ void ccdSceneSelectKernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ order, uint32_t n_cand,
                                     const long long* __restrict__ cand_code, const S* __restrict__ cand_toc,
                                     const S* __restrict__ cand_box, uint32_t max_contacts, uint32_t keep,
                                     uint32_t* __restrict__ counts, long long* __restrict__ code, S* __restrict__ toc,
                                     S* __restrict__ box) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n_cand) return;
  const uint32_t q = keys[i];
  if (q == 0xffffffffu) return;
  auto lowerBound = [&](unsigned long long v) {
    size_t lo = 0, hi = n_cand;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if ((unsigned long long)keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const size_t first = lowerBound(q);
  const size_t k = i - first;
  if (k == 0) {
    const size_t cnt = lowerBound((unsigned long long)q + 1) - first;
    counts[q] = uint32_t(cnt < max_contacts ? cnt : max_contacts);
  }
  if (k < max_contacts && k < keep) {
    const uint32_t c = order[i];
    const size_t o = size_t(q) * keep + k;
    code[o] = cand_code[c];
    if (toc) {
      toc[2 * o] = cand_toc[2 * size_t(c)];
      toc[2 * o + 1] = cand_toc[2 * size_t(c) + 1];
    }
    if (box)
      for (int j = 0; j < 6; j++) box[6 * o + j] = cand_box[6 * size_t(c) + j];
  }
}

__global__ void ccdSceneIotaKernel(uint32_t* p, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
__global__ void ccdSceneGatherKernel(const uint32_t* __restrict__ qkey, const uint32_t* __restrict__ order, uint32_t n, uint32_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = qkey[order[i]];
}
template <typename T>
__global__ void ccdSceneFillKernel(T* p, size_t n, T v) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}

struct CcdSceneScratch {
  std::vector<void*> ptrs;
  ~CcdSceneScratch() {
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
static int ccdSceneDev(Engine& e, int kind, fclb_handle scene, ShapeTable* t, const uint32_t* shape_ids, const void* poses_shape,
                       const void* poses_scene, const void* disp, size_t n, const fclb_ccd_request* req, int scene_moves,
                       uint32_t keep, uint32_t* counts, long long* code, void* toc, void* box) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  CcdSceneArgs a{};
  int rc = kind == FCLB_SCENE_HEIGHTMAP ? sceneHmView(scene, st, a.hm) : sceneOctView(scene, a.oct);
  if (rc) return rc;
  CcdSceneScratch sc;
  uint32_t* d_count = nullptr;
  unsigned long long* d_work = nullptr;
  FCLB_CUDA(sc.get(&d_count, 2));
  FCLB_CUDA(sc.get(&d_work, 1));
  a.shapes = t->d_shapes[st];
  a.convex = e.d_convex_tab[st];
  a.local = t->d_local[st];
  a.shape_ids = shape_ids;
  a.poses_shape = poses_shape;
  a.poses_scene = poses_scene;
  a.disp = disp;
  a.n = n;
  a.scene_moves = scene_moves;
  a.request_type = int(req->request_type);
  a.zero_tol = req->zero_movement_tolerance > 0 ? req->zero_movement_tolerance : 1e-4;
  a.gjk_tol = req->gjk_tolerance > 0 ? req->gjk_tolerance : 1e-6;
  a.max_iter = req->max_gjk_iterations > 0 ? req->max_gjk_iterations : 128;
  a.cand_count = d_count;
  a.work_counter = d_work;
  const size_t smem = size_t(kCsWarps) * (kCsStackCap * sizeof(CsElem<S>) + 3 * kFitMaxPoints * sizeof(S));
  auto kern = kind == FCLB_SCENE_HEIGHTMAP ? ccdSceneTraverseKernel<S, FCLB_SCENE_HEIGHTMAP> : ccdSceneTraverseKernel<S, FCLB_SCENE_OCTREE>;
  FCLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int grid = int(std::min<size_t>((n + kCsWarps - 1) / kCsWarps, size_t(e.sms) * 2));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  size_t cap = std::min<size_t>(std::max<size_t>(n * 32, size_t(1) << 16), size_t(1) << 26);
  uint32_t h_count[2] = {0, 0};
  for (int attempt = 0; attempt < 3; attempt++) {
    FCLB_CUDA(sc.get(&a.cand_q, cap));
    FCLB_CUDA(sc.get(&a.cand_code, cap));
    FCLB_CUDA(sc.get(&a.cand_path, cap));
    S* bx = nullptr;
    FCLB_CUDA(sc.get(&bx, 6 * cap));
    a.cand_box = bx;
    a.cand_cap = uint32_t(cap);
    FCLB_CUDA(cudaMemsetAsync(d_count, 0, 2 * sizeof(uint32_t), e.compute));
    FCLB_CUDA(cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), e.compute));
    kern<<<grid, kCsWarps * 32, smem, e.compute>>>(a);
    FCLB_CUDA(cudaGetLastError());
    e.launches += 1;
    FCLB_CUDA(cudaMemcpyAsync(h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    if (h_count[1]) return fail(FCLB_ERR_CAPACITY, "scene CCD: hierarchy wider or deeper than the per-warp stack allows");
    if (h_count[0] <= cap) break;
    if (attempt == 2 || h_count[0] > (1u << 30)) return fail(FCLB_ERR_CAPACITY, "scene CCD: too many candidate boxes: split the batch");
    cap = h_count[0];
  }
  const uint32_t n_cand = h_count[0];
  FCLB_CUDA(cudaMemsetAsync(counts, 0, n * sizeof(uint32_t), e.compute));
  if (n * keep) {
    const int g = int((n * keep * 6 + 255) / 256);
    ccdSceneFillKernel<long long><<<g, 256, 0, e.compute>>>(code, n * keep, -1ll);
    if (toc) ccdSceneFillKernel<S><<<g, 256, 0, e.compute>>>(static_cast<S*>(toc), n * keep * 2, S(-1));
    if (box) ccdSceneFillKernel<S><<<g, 256, 0, e.compute>>>(static_cast<S*>(box), n * keep * 6, S(0));
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
    ccdSceneLeafKernel<S><<<lgrid, kBlock, 0, e.compute>>>(a, n_cand);
    FCLB_CUDA(cudaGetLastError());
    const int g256 = int((n_cand + 255) / 256);
    ccdSceneIotaKernel<<<g256, 256, 0, e.compute>>>(order, n_cand);
    size_t b1 = 0, b2 = 0;
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, a.cand_path, path_sorted, order, order1, int(n_cand), 0, 64, e.compute));
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b2, qkey1, qkey2, order1, order2, int(n_cand), 0, 32, e.compute));
    unsigned char* tmp = nullptr;
    FCLB_CUDA(sc.get(&tmp, std::max(b1, b2)));
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b1, a.cand_path, path_sorted, order, order1, int(n_cand), 0, 64, e.compute));
    ccdSceneGatherKernel<<<g256, 256, 0, e.compute>>>(qkey, order1, n_cand, qkey1);
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b2, qkey1, qkey2, order1, order2, int(n_cand), 0, 32, e.compute));
    ccdSceneSelectKernel<S><<<g256, 256, 0, e.compute>>>(qkey2, order2, n_cand, a.cand_code, cand_toc, static_cast<const S*>(a.cand_box),
                                                         req->max_contacts ? req->max_contacts : 1u, keep, counts, code,
                                                         static_cast<S*>(toc), static_cast<S*>(box));
    FCLB_CUDA(cudaGetLastError());
    e.launches += 6;
  }
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -9;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

// ---- heightmap / octree vs mesh ------------------------------------------------------------------------------------
// runHeightMapObbBVH (heightmap_ccd_solver-inl.h:168-285) and runOctreeObbBVH (octree2_ccd_solver-inl.h:263-438): the walk
// over (box node, mesh node) pairs.  Both boxes are taken to the world as OBBs (convertBV) and tested with
// BoxPairTranslationalCCD::IsDisjoint, every surviving pair narrowing its parent's interval.  Descent: heightmap -- the
// map is descended when the mesh node is a leaf or the map box is the larger one (AABB::size() = |max - min|^2 against
// OBB::size() = |extent|^2, as written); octree -- the mesh is descended when the octree node is a traverse leaf or its box is
// the smaller one.  Leaf pairs run RunShapeSimplex<Box> (box swept, triangle): the contact carries NO external interval,
// so kBoxApproximate pre-checks with the shapes' local AABBs -- and the triangle's is the empty AABB TriangleP is
// constructed with (+max / -max, never computed), which the arithmetic below reproduces as is.
struct CcdSceneMeshArgs {
  HmView hm;
  OctView oct;
  const void* nodes;
  const void* tris;
  const void* poses_scene;
  const void* poses_mesh;
  const void* disp;
  size_t n;
  int mesh_moves;
  int request_type;
  double zero_tol, gjk_tol;
  int max_iter;
  uint32_t* cand_count;
  uint32_t cand_cap;
  uint32_t* cand_q;
  long long* cand_code;
  int* cand_tri;
  void* cand_box;
  unsigned long long* cand_path;
  unsigned long long* work_counter;
  uint32_t* qkey;
  void* cand_toc;
};

constexpr int kCmsStackCap = 448;  // 4 warps x 448 x 96 B (double) = 172 KB

template <typename S>
struct CmsElem {
  BoxElem<S> box;
  int node;
  S lo, hi;
  unsigned long long path;
  int shift;
};

template <typename S>
FCLB_DI V3<S> ccdSceneMeshDisplacement(const S* disp, size_t q, const Pose<S>& tf_scene, const Pose<S>& tf_mesh, int mesh_moves) {
  V3<S> unit = mk<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
  if (mesh_moves) unit = mulMtV(tf_scene.R, mulMV(tf_mesh.R, -unit));  // RunObbBVH_HeightMap / RunObbBVH_Octree
  return unit;
}

template <typename S, int KIND>
__global__ void __launch_bounds__(kCsWarps * 32) ccdSceneMeshTraverseKernel(CcdSceneMeshArgs a) {
  extern __shared__ __align__(16) unsigned char s_cs[];
  using Side = typename SideOf<S, KIND>::type;
  const Side side = SideOf<S, KIND>::make(a.hm, a.oct);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  CmsElem<S>* stack = reinterpret_cast<CmsElem<S>*>(s_cs) + size_t(warp) * kCmsStackCap;
  const S* __restrict__ nodes = static_cast<const S*>(a.nodes);
  const S* __restrict__ tris = static_cast<const S*>(a.tris);
  const S* __restrict__ disp = static_cast<const S*>(a.disp);
  const unsigned lt_mask = (1u << lane) - 1u;
  const S zero_tol = S(a.zero_tol);
  constexpr int kChildBits = KIND == FCLB_SCENE_HEIGHTMAP ? 2 : 3;
  constexpr int kMaxChildren = KIND == FCLB_SCENE_HEIGHTMAP ? 4 : 8;
  while (true) {
    unsigned long long q64 = 0;
    if (lane == 0) q64 = atomicAdd(a.work_counter, 1ull);
    q64 = __shfl_sync(0xffffffffu, q64, 0);
    if (q64 >= a.n) break;
    const size_t q = size_t(q64);
    const Pose<S> tf_g = loadPose(static_cast<const S*>(a.poses_scene), q);
    const Pose<S> tf_m = loadPose(static_cast<const S*>(a.poses_mesh), q);
    const V3<S> unit = ccdSceneMeshDisplacement(disp, q, tf_g, tf_m, a.mesh_moves);
    const S scalar = disp[4 * q + 3];
    const int n_roots = side.numRoots();
    int root_bits = 0;
    while ((1 << root_bits) < n_roots) root_bits++;
    int sp = 0;
    bool overflow = n_roots > kCmsStackCap - 8 * 32;
    if (!overflow) {
      #pragma unroll 1
      for (int base = 0; base < n_roots; base += 32) {
        const int i = base + lane;
        CmsElem<S> e;
        bool ok = false;
        if (i < n_roots) {
          ok = side.root(i, e.box);
          e.node = 0;
          e.lo = S(0.0);
          e.hi = S(1.0);
          e.shift = 64 - root_bits;
          e.path = root_bits ? ((unsigned long long)(n_roots - 1 - i) << e.shift) : 0ull;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) stack[sp + __popc(m & lt_mask)] = e;
        sp += __popc(m);
      }
    }
    __syncwarp();
    while (sp > 0 && !overflow) {
      int take = sp < 32 ? sp : 32;
      if (sp + take * kMaxChildren > kCmsStackCap) take = (kCmsStackCap - sp) / kMaxChildren > 0 ? 1 : 0;
      if (take == 0) {
        overflow = true;
        break;
      }
      CmsElem<S> e;
      if (lane < take) e = stack[sp - 1 - lane];
      sp -= take;
      __syncwarp();
      int n_push = 0, n_cand = 0, tri = -1, mesh_child = -1;
      unsigned child_mask = 0;
      bool on_scene = false, voxels = false;
      TocInterval<S> iv;
      iv.lo = iv.hi = S(0);
      if (lane < take) {
        iv.lo = e.lo;
        iv.hi = e.hi;
        NodeD<S> nd = loadNode(nodes, e.node);
        const bool mesh_leaf = nd.first_child < 0;
        const S mesh_size = sqnorm(nd.extent);  // (of the stored box: read only when the node is internal)
        if (mesh_leaf) {
          V3<S> P[3];
          loadTri(tris, -(nd.first_child + 1), P);
          fitObb3(P, nd.axis, nd.To, nd.extent);
        }
        // convertBV(AABB, tf_scene) (math/bv/utility-inl.h:587-608) and convertBV(OBB, tf_mesh)
        const V3<S> mn = mk<S>(e.box.mn[0], e.box.mn[1], e.box.mn[2]), mx = mk<S>(e.box.mx[0], e.box.mx[1], e.box.mx[2]);
        const V3<S> c = (mn + mx) * S(0.5);
        const V3<S> To1 = mk<S>(((tf_g.R.m[0] * c.x + tf_g.R.m[1] * c.y) + tf_g.R.m[2] * c.z) + tf_g.t.x,
                                ((tf_g.R.m[3] * c.x + tf_g.R.m[4] * c.y) + tf_g.R.m[5] * c.z) + tf_g.t.y,
                                ((tf_g.R.m[6] * c.x + tf_g.R.m[7] * c.y) + tf_g.R.m[8] * c.z) + tf_g.t.z);
        const V3<S> ext1 = (mx - mn) * S(0.5);
        M3<S> a2;
        V3<S> t2;
        obbToWorld(tf_m, nd.axis, nd.To, a2, t2);
        if (!boxPairCcdDisjoint(tf_g.R, To1, ext1, unit, scalar, a2, t2, nd.extent, iv, zero_tol, true)) {
          const bool scene_leaf = KIND == FCLB_SCENE_HEIGHTMAP ? (e.box.meta & 1u) != 0 : (e.box.meta & 3u) != 0;
          if (mesh_leaf && scene_leaf) {
            tri = -(nd.first_child + 1);
            if (KIND == FCLB_SCENE_OCTREE && !(e.box.meta & 1u)) {  // partial leaf: its voxels, each against the triangle
              voxels = true;
              child_mask = side.childMask(e.box);
              n_cand = __popc(child_mask);
            } else {
              n_cand = 1;
            }
          } else {
            const V3<S> full = mx - mn;
            const S scene_size = sqnorm(full);  // AABB::size()
            if (KIND == FCLB_SCENE_HEIGHTMAP)
              on_scene = mesh_leaf || (!scene_leaf && scene_size > mesh_size);
            else
              on_scene = !(scene_leaf || (!mesh_leaf && scene_size < mesh_size));
            if (on_scene) {
              child_mask = side.childMask(e.box);
              if (e.shift < kChildBits) overflow = true;
              n_push = __popc(child_mask);
            } else {
              mesh_child = nd.first_child;
              if (e.shift < 1) overflow = true;
              n_push = 2;
            }
          }
        }
      }
      if (__any_sync(0xffffffffu, overflow)) {
        overflow = true;
        break;
      }
      int push_off = n_push, cand_off = n_cand;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int p = __shfl_up_sync(0xffffffffu, push_off, o), c = __shfl_up_sync(0xffffffffu, cand_off, o);
        if (lane >= o) {
          push_off += p;
          cand_off += c;
        }
      }
      const int total_push = __shfl_sync(0xffffffffu, push_off, 31), total_cand = __shfl_sync(0xffffffffu, cand_off, 31);
      push_off -= n_push;
      cand_off -= n_cand;
      if (n_push) {
        if (on_scene) {
          int k = 0;
          const int shift = e.shift - kChildBits;
          for (int c = 0; c < kMaxChildren; c++) {
            if (!(child_mask & (1u << c))) continue;
            CmsElem<S> ch;
            ch.box = side.child(e.box, c);
            ch.node = e.node;
            ch.lo = iv.lo;
            ch.hi = iv.hi;
            ch.shift = shift;
            ch.path = e.path | ((unsigned long long)(kMaxChildren - 1 - c) << shift);
            stack[sp + push_off + k] = ch;
            k++;
          }
        } else {  // left pushed first, right popped first
          const int shift = e.shift - 1;
          CmsElem<S> ch = e;
          ch.lo = iv.lo;
          ch.hi = iv.hi;
          ch.shift = shift;
          ch.node = mesh_child;
          ch.path = e.path | (1ull << shift);
          stack[sp + push_off] = ch;
          ch.node = mesh_child + 1;
          ch.path = e.path;
          stack[sp + push_off + 1] = ch;
        }
      }
      sp += total_push;
      if (total_cand) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.cand_count, uint32_t(total_cand));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (n_cand) {
          uint32_t slot = base + uint32_t(cand_off);
          auto emit = [&](const BoxElem<S>& bx, unsigned long long path) {
            if (slot < a.cand_cap) {
              a.cand_q[slot] = uint32_t(q);
              a.cand_code[slot] = Side::code(bx);
              a.cand_tri[slot] = tri;
              S* o = static_cast<S*>(a.cand_box) + 6 * size_t(slot);
#pragma unroll
              for (int j = 0; j < 3; j++) {
                o[j] = bx.mn[j];
                o[3 + j] = bx.mx[j];
              }
              a.cand_path[slot] = path;
            }
            slot++;
          };
          if (!voxels) {
            emit(e.box, e.path);
          } else {
            const int shift = e.shift >= 3 ? e.shift - 3 : 0;
            for (int c = 0; c < 8; c++)
              if (child_mask & (1u << c)) emit(side.child(e.box, c), e.path | ((unsigned long long)c << shift));
          }
        }
      }
      __syncwarp();
    }
    if (overflow && lane == 0) atomicAdd(a.cand_count + 1, 1u);
    __syncwarp();
  }
}

// boxToSimplexProcessLeafPair: RunShapeSimplex<Box>(box of the pixel / node swept, triangle)
template <typename S>
__global__ void __launch_bounds__(kBlock) ccdSceneMeshLeafKernel(CcdSceneMeshArgs a, uint32_t n_cand) {
  const S* __restrict__ disp = static_cast<const S*>(a.disp);
  const S zero_tol = S(a.zero_tol), tol = S(a.gjk_tol);
  const S big = sizeof(S) == 4 ? S(3.402823466e+38f) : S(1.7976931348623157e+308);
  #pragma unroll 1
  for (size_t c = blockIdx.x * size_t(blockDim.x) + threadIdx.x; c < n_cand; c += size_t(gridDim.x) * blockDim.x) {
    const size_t q = a.cand_q[c];
    const Pose<S> tf_g = loadPose(static_cast<const S*>(a.poses_scene), q);
    const Pose<S> tf_m = loadPose(static_cast<const S*>(a.poses_mesh), q);
    const V3<S> unit = ccdSceneMeshDisplacement(disp, q, tf_g, tf_m, a.mesh_moves);
    const S* bx = static_cast<const S*>(a.cand_box) + 6 * c;
    const V3<S> mn = mk<S>(bx[0], bx[1], bx[2]), mx = mk<S>(bx[3], bx[4], bx[5]);
    ShapeInst<S> box, tri;
    box.type = ST_BOX;
    box.cvx = nullptr;
    box.p0 = mx.x - mn.x;
    box.p1 = mx.y - mn.y;
    box.p2 = mx.z - mn.z;
    const V3<S> center = (mn + mx) * S(0.5);
    Pose<S> tf_b;
    tf_b.R = tf_g.R;
    tf_b.t = mulMV(tf_g.R, center) + tf_g.t;
    LocalAabbD<S> lb, lt;
    lb.mx[0] = S(0.5) * box.p0; lb.mx[1] = S(0.5) * box.p1; lb.mx[2] = S(0.5) * box.p2;
    lb.mn[0] = -lb.mx[0]; lb.mn[1] = -lb.mx[1]; lb.mn[2] = -lb.mx[2];
    lb.center[0] = lb.center[1] = lb.center[2] = S(0);
    lb.radius = S(0);
    // TriangleP's aabb_local is never computed on this path: the empty AABB of AABB<S>::AABB() (math/bv/AABB-inl.h:47-51)
    for (int k = 0; k < 3; k++) {
      lt.mn[k] = big;
      lt.mx[k] = -big;
      lt.center[k] = (lt.mn[k] + lt.mx[k]) * S(0.5);
    }
    lt.radius = S(0);
    tri.type = ST_TRIANGLE;
    tri.cvx = nullptr;
    tri.p0 = tri.p1 = tri.p2 = S(0);
    loadTri(static_cast<const S*>(a.tris), a.cand_tri[c], tri.tri);
    TocInterval<S> toc;
    const bool hit = ccdShapePairEval<S>(box, lb, tf_b, tri, lt, tf_m, unit, disp[4 * q + 3], a.request_type, zero_tol, tol, a.max_iter, toc);
    a.qkey[c] = hit ? uint32_t(q) : 0xffffffffu;
    S* o = static_cast<S*>(a.cand_toc) + 2 * c;
    const bool valid = toc.lo >= 0 && toc.hi >= 0;
    o[0] = valid ? toc.lo : S(-1.0);
    o[1] = valid ? toc.hi : S(-1.0);
  }
}

template <typename S>
__global__ void ccdSceneMeshSelectKernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ order, uint32_t n_cand,
                                         const long long* __restrict__ cand_code, const int* __restrict__ cand_tri,
                                         const S* __restrict__ cand_toc, const S* __restrict__ cand_box, uint32_t max_contacts,
                                         uint32_t keep, uint32_t* __restrict__ counts, long long* __restrict__ ids, S* __restrict__ toc,
                                         S* __restrict__ box) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n_cand) return;
  const uint32_t q = keys[i];
  if (q == 0xffffffffu) return;
  auto lowerBound = [&](unsigned long long v) {
    size_t lo = 0, hi = n_cand;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if ((unsigned long long)keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const size_t first = lowerBound(q);
  const size_t k = i - first;
  if (k == 0) {
    const size_t cnt = lowerBound((unsigned long long)q + 1) - first;
    counts[q] = uint32_t(cnt < max_contacts ? cnt : max_contacts);
  }
  if (k < max_contacts && k < keep) {
    const uint32_t c = order[i];
    const size_t o = size_t(q) * keep + k;
    ids[2 * o] = cand_code[c];
    ids[2 * o + 1] = cand_tri[c];
    if (toc) {
      toc[2 * o] = cand_toc[2 * size_t(c)];
      toc[2 * o + 1] = cand_toc[2 * size_t(c) + 1];
    }
    if (box)
      for (int j = 0; j < 6; j++) box[6 * o + j] = cand_box[6 * size_t(c) + j];
  }
}

template <typename S>
static int ccdSceneMeshDev(Engine& e, int kind, fclb_handle scene, const BvhDev* m, const void* poses_scene, const void* poses_mesh,
                           const void* disp, size_t n, const fclb_ccd_request* req, int mesh_moves, uint32_t keep, uint32_t* counts,
                           long long* ids, void* toc, void* box) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  CcdSceneMeshArgs a{};
  int rc = kind == FCLB_SCENE_HEIGHTMAP ? sceneHmView(scene, st, a.hm) : sceneOctView(scene, a.oct);
  if (rc) return rc;
  CcdSceneScratch sc;
  uint32_t* d_count = nullptr;
  unsigned long long* d_work = nullptr;
  FCLB_CUDA(sc.get(&d_count, 2));
  FCLB_CUDA(sc.get(&d_work, 1));
  a.nodes = m->nodes;
  a.tris = m->tris;
  a.poses_scene = poses_scene;
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
  const size_t smem = size_t(kCsWarps) * kCmsStackCap * sizeof(CmsElem<S>);
  auto kern = kind == FCLB_SCENE_HEIGHTMAP ? ccdSceneMeshTraverseKernel<S, FCLB_SCENE_HEIGHTMAP> : ccdSceneMeshTraverseKernel<S, FCLB_SCENE_OCTREE>;
  FCLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int grid = int(std::min<size_t>((n + kCsWarps - 1) / kCsWarps, size_t(e.sms) * 2));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  size_t cap = std::min<size_t>(std::max<size_t>(n * 64, size_t(1) << 16), size_t(1) << 26);
  uint32_t h_count[2] = {0, 0};
  for (int attempt = 0; attempt < 3; attempt++) {
    FCLB_CUDA(sc.get(&a.cand_q, cap));
    FCLB_CUDA(sc.get(&a.cand_code, cap));
    FCLB_CUDA(sc.get(&a.cand_tri, cap));
    FCLB_CUDA(sc.get(&a.cand_path, cap));
    S* bx = nullptr;
    FCLB_CUDA(sc.get(&bx, 6 * cap));
    a.cand_box = bx;
    a.cand_cap = uint32_t(cap);
    FCLB_CUDA(cudaMemsetAsync(d_count, 0, 2 * sizeof(uint32_t), e.compute));
    FCLB_CUDA(cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), e.compute));
    kern<<<grid, kCsWarps * 32, smem, e.compute>>>(a);
    FCLB_CUDA(cudaGetLastError());
    e.launches += 1;
    FCLB_CUDA(cudaMemcpyAsync(h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    if (h_count[1]) return fail(FCLB_ERR_CAPACITY, "scene-mesh CCD: hierarchies wider or deeper than the per-warp stack / 64 path bits allow");
    if (h_count[0] <= cap) break;
    if (attempt == 2 || h_count[0] > (1u << 30)) return fail(FCLB_ERR_CAPACITY, "scene-mesh CCD: too many candidate pairs: split the batch");
    cap = h_count[0];
  }
  const uint32_t n_cand = h_count[0];
  FCLB_CUDA(cudaMemsetAsync(counts, 0, n * sizeof(uint32_t), e.compute));
  if (n * keep) {
    const int g = int((n * keep * 6 + 255) / 256);
    ccdSceneFillKernel<long long><<<g, 256, 0, e.compute>>>(ids, n * keep * 2, -1ll);
    if (toc) ccdSceneFillKernel<S><<<g, 256, 0, e.compute>>>(static_cast<S*>(toc), n * keep * 2, S(-1));
    if (box) ccdSceneFillKernel<S><<<g, 256, 0, e.compute>>>(static_cast<S*>(box), n * keep * 6, S(0));
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
    ccdSceneMeshLeafKernel<S><<<lgrid, kBlock, 0, e.compute>>>(a, n_cand);
    FCLB_CUDA(cudaGetLastError());
    const int g256 = int((n_cand + 255) / 256);
    ccdSceneIotaKernel<<<g256, 256, 0, e.compute>>>(order, n_cand);
    size_t b1 = 0, b2 = 0;
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, a.cand_path, path_sorted, order, order1, int(n_cand), 0, 64, e.compute));
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b2, qkey1, qkey2, order1, order2, int(n_cand), 0, 32, e.compute));
    unsigned char* tmp = nullptr;
    FCLB_CUDA(sc.get(&tmp, std::max(b1, b2)));
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b1, a.cand_path, path_sorted, order, order1, int(n_cand), 0, 64, e.compute));
    ccdSceneGatherKernel<<<g256, 256, 0, e.compute>>>(qkey, order1, n_cand, qkey1);
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b2, qkey1, qkey2, order1, order2, int(n_cand), 0, 32, e.compute));
    ccdSceneMeshSelectKernel<S><<<g256, 256, 0, e.compute>>>(qkey2, order2, n_cand, a.cand_code, a.cand_tri, cand_toc,
                                                             static_cast<const S*>(a.cand_box), req->max_contacts ? req->max_contacts : 1u,
                                                             keep, counts, ids, static_cast<S*>(toc), static_cast<S*>(box));
    FCLB_CUDA(cudaGetLastError());
    e.launches += 6;
  }
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -10;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

// ---- heightmap / octree vs heightmap / octree ---------------------------------------------------------------------
// runHeightMapPair (heightmap_ccd_solver-inl.h:636-757), runHeightMapOctree (:369-561, the octree-first entry
// RunOctreeHeightMap :605-632 re-expresses the displacement and runs the same walk with the heightmap first) and
// runOctreePair (octree2_ccd_solver-inl.h:484-922).  Every box is an AABB in its own scene frame and the
// fixed-orientation swept-box test (fixedCcdDisjoint) does all the work: on a node pair it culls or narrows the
// interval the children inherit, on a terminal pair (boxToBoxProcessLeafPair and the octree leaf routines) it runs
// from [0, 1] and a surviving pair IS the contact, its interval the contact's toc.  Descent rules as written there:
// two heightmaps -- map 1 when node 2 is a bottom pixel or node 1 is the shallower one; heightmap / octree -- the map
// when the octree node is a traverse leaf or the map node is the shallower one; two octrees -- tree 2 when node 1 is a
// traverse leaf or its box is the smaller one (AABB::size() = |max - min|^2).  Two partial octree leaves: the pair's own
// test, then (n1 n2 > n1 + n2) voxels pruned against the other leaf's box, then voxel 1 x voxel 2 in index order.
// One launch walks and tests (hits carry their path), two stable radix sorts and the select pass order them.
struct CcdScenePairArgs {
  HmView hm1, hm2;
  OctView oct1, oct2;
  const void* poses_a;
  const void* poses_b;
  const void* disp;
  size_t n;
  int swapped;  // the caller's geometry 1 (the moving one) is side B: (octree, heightmap)
  double zero_tol;
  uint32_t* cand_count;  // [0] hits, [1] overflows
  uint32_t cand_cap;
  uint32_t* cand_q;
  long long* cand_ids;   // 2 per hit, in the caller's argument order
  void* cand_box;        // 12 S per hit
  void* cand_toc;        // 2 S per hit
  unsigned long long* cand_path;
  unsigned long long* work_counter;
};

constexpr int kCspStackCap = 320;

template <typename S>
struct CspElem {
  BoxElem<S> a, b;
  S lo, hi;
  unsigned long long path;
  int shift;
};

template <typename S, int K>
FCLB_DI bool cspTraverseLeaf(const BoxElem<S>& e) {
  return K == FCLB_SCENE_HEIGHTMAP ? (e.meta & 1u) != 0 : (e.meta & 3u) != 0;
}
template <typename S, int K>
FCLB_DI bool cspPartialLeaf(const BoxElem<S>& e) {
  return K == FCLB_SCENE_OCTREE && (e.meta & 2u) && !(e.meta & 1u);
}
template <typename S>
FCLB_DI V3<S> boxMin(const BoxElem<S>& e) { return mk<S>(e.mn[0], e.mn[1], e.mn[2]); }
template <typename S>
FCLB_DI V3<S> boxMax(const BoxElem<S>& e) { return mk<S>(e.mx[0], e.mx[1], e.mx[2]); }

template <typename S, int KA, int KB>
__global__ void __launch_bounds__(kCsWarps * 32) ccdScenePairKernel(CcdScenePairArgs a) {
  extern __shared__ __align__(16) unsigned char s_cs[];
  using SA = SideOf<S, KA>;
  using SB = SideOf<S, KB>;
  const typename SA::type sideA = SA::make(a.hm1, a.oct1);
  const typename SB::type sideB = SB::make(a.hm2, a.oct2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  CspElem<S>* stack = reinterpret_cast<CspElem<S>*>(s_cs) + size_t(warp) * kCspStackCap;
  const S* __restrict__ disp = static_cast<const S*>(a.disp);
  const unsigned lt_mask = (1u << lane) - 1u;
  const S zero_tol = S(a.zero_tol);
  constexpr int kBitsA = KA == FCLB_SCENE_HEIGHTMAP ? 2 : 3, kBitsB = KB == FCLB_SCENE_HEIGHTMAP ? 2 : 3;
  constexpr int kChildA = 1 << kBitsA, kChildB = 1 << kBitsB;
  constexpr bool kRawB = KA == FCLB_SCENE_HEIGHTMAP && KB == FCLB_SCENE_OCTREE;  // b2 = node_vector_index as is (:449-479)
  while (true) {
    unsigned long long q64 = 0;
    if (lane == 0) q64 = atomicAdd(a.work_counter, 1ull);
    q64 = __shfl_sync(0xffffffffu, q64, 0);
    if (q64 >= a.n) break;
    const size_t q = size_t(q64);
    const Pose<S> tf_a = loadPose(static_cast<const S*>(a.poses_a), q);
    const Pose<S> tf_b = loadPose(static_cast<const S*>(a.poses_b), q);
    FixedCcd<S> f;
    {
      const Pose<S> rel = compose(inverse(tf_a), tf_b);
      f.R = rel.R;
      f.t = rel.t;
      f.unit = mk<S>(disp[4 * q], disp[4 * q + 1], disp[4 * q + 2]);
      if (a.swapped) f.unit = mulMtV(tf_a.R, mulMV(tf_b.R, -f.unit));  // RunOctreeHeightMap
      f.scalar = disp[4 * q + 3];
    }
    // one hit: tested from [0, 1], written with its place in the reference's walk
    auto testEmit = [&](const BoxElem<S>& ba, const BoxElem<S>& bb, unsigned long long path) {
      TocInterval<S> t;
      t.lo = S(0.0);
      t.hi = S(1.0);
      if (fixedCcdDisjoint(f, boxMin(ba), boxMax(ba), boxMin(bb), boxMax(bb), t, zero_tol)) return;
      const uint32_t slot = atomicAdd(a.cand_count, 1u);
      if (slot >= a.cand_cap) return;
      a.cand_q[slot] = uint32_t(q);
      const long long ca = SA::type::code(ba), cb = kRawB ? (long long)bb.index : SB::type::code(bb);
      a.cand_ids[2 * size_t(slot)] = a.swapped ? cb : ca;
      a.cand_ids[2 * size_t(slot) + 1] = a.swapped ? ca : cb;
      S* o = static_cast<S*>(a.cand_box) + 12 * size_t(slot);
      const BoxElem<S>& b1 = a.swapped ? bb : ba;
      const BoxElem<S>& b2 = a.swapped ? ba : bb;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        o[j] = b1.mn[j];
        o[3 + j] = b1.mx[j];
        o[6 + j] = b2.mn[j];
        o[9 + j] = b2.mx[j];
      }
      S* ot = static_cast<S*>(a.cand_toc) + 2 * size_t(slot);
      ot[0] = t.lo;
      ot[1] = t.hi;
      a.cand_path[slot] = path;
    };
    // root pairs: map-1 pixels row by row (outer) x map-2 pixels (inner) / the octree root, pushed in that order and
    // popped in reverse.  Taken 32 at a time; the order of the walk itself is free, the path orders the hits.
    const int n_ra = sideA.numRoots(), n_rb = sideB.numRoots();
    const int n_roots = n_ra * n_rb;
    int root_bits = 0;
    while ((1 << root_bits) < n_roots) root_bits++;
    bool overflow = false;
    #pragma unroll 1
    for (int base = 0; base < n_roots && !overflow; base += 32) {
      int sp = 0;
      {
        const int i = base + lane;
        CspElem<S> e;
        bool ok = false;
        if (i < n_roots) {
          ok = sideA.root(i / n_rb, e.a) && sideB.root(i % n_rb, e.b);
          e.lo = S(0.0);
          e.hi = S(1.0);
          e.shift = 64 - root_bits;
          e.path = root_bits ? ((unsigned long long)(n_roots - 1 - i) << e.shift) : 0ull;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) stack[__popc(m & lt_mask)] = e;
        sp = __popc(m);
      }
      __syncwarp();
      while (sp > 0 && !overflow) {
        int take = sp < 32 ? sp : 32;
        if (sp + take * 8 > kCspStackCap) take = (kCspStackCap - sp) / 8 > 0 ? 1 : 0;
        if (take == 0) {
          overflow = true;
          break;
        }
        CspElem<S> e;
        if (lane < take) e = stack[sp - 1 - lane];
        sp -= take;
        __syncwarp();
        int n_push = 0;
        unsigned child_mask = 0;
        bool on_a = false;
        TocInterval<S> iv;
        iv.lo = iv.hi = S(0);
        if (lane < take) {
          const bool leaf_a = cspTraverseLeaf<S, KA>(e.a), leaf_b = cspTraverseLeaf<S, KB>(e.b);
          if (leaf_a && leaf_b) {
            const bool part_a = cspPartialLeaf<S, KA>(e.a), part_b = cspPartialLeaf<S, KB>(e.b);
            if (!part_a && !part_b) {
              testEmit(e.a, e.b, e.path);
            } else if (e.shift < (part_a ? 3 : 0) + (part_b ? 3 : 0)) {
              overflow = true;
            } else if (part_a && part_b) {  // runOctreePairTwoLeafNode, both partial (:846-919)
              TocInterval<S> t;
              t.lo = S(0.0);
              t.hi = S(1.0);
              if (!fixedCcdDisjoint(f, boxMin(e.a), boxMax(e.a), boxMin(e.b), boxMax(e.b), t, zero_tol)) {
                unsigned ma = sideA.childMask(e.a), mb = sideB.childMask(e.b);
                const int na = __popc(ma), nb = __popc(mb);
                if (na * nb > na + nb) {
                  const unsigned ma0 = ma, mb0 = mb;
                  for (int c = 0; c < 8; c++) {
                    if (!(ma0 & (1u << c))) continue;
                    const BoxElem<S> ca = sideA.child(e.a, c);
                    t.lo = S(0.0);
                    t.hi = S(1.0);
                    if (fixedCcdDisjoint(f, boxMin(ca), boxMax(ca), boxMin(e.b), boxMax(e.b), t, zero_tol)) ma &= ~(1u << c);
                  }
                  for (int c = 0; c < 8; c++) {
                    if (!(mb0 & (1u << c))) continue;
                    const BoxElem<S> cb = sideB.child(e.b, c);
                    t.lo = S(0.0);
                    t.hi = S(1.0);
                    if (fixedCcdDisjoint(f, boxMin(e.a), boxMax(e.a), boxMin(cb), boxMax(cb), t, zero_tol)) mb &= ~(1u << c);
                  }
                }
                const int shift = e.shift - 6;
                for (int c1 = 0; c1 < 8; c1++) {
                  if (!(ma & (1u << c1))) continue;
                  const BoxElem<S> ca = sideA.child(e.a, c1);
                  for (int c2 = 0; c2 < 8; c2++)
                    if (mb & (1u << c2)) testEmit(ca, sideB.child(e.b, c2), e.path | ((unsigned long long)(c1 * 8 + c2) << shift));
                }
              }
            } else {  // the voxels of the partial side in index order, without a test of the leaf's own box
              const int shift = e.shift - 3;
              const unsigned m = part_a ? sideA.childMask(e.a) : sideB.childMask(e.b);
              for (int c = 0; c < 8; c++) {
                if (!(m & (1u << c))) continue;
                const unsigned long long path = e.path | ((unsigned long long)c << shift);
                if (part_a)
                  testEmit(sideA.child(e.a, c), e.b, path);
                else
                  testEmit(e.a, sideB.child(e.b, c), path);
              }
            }
          } else {
            iv.lo = e.lo;
            iv.hi = e.hi;
            if (!fixedCcdDisjoint(f, boxMin(e.a), boxMax(e.a), boxMin(e.b), boxMax(e.b), iv, zero_tol)) {
              if (KA == FCLB_SCENE_HEIGHTMAP && KB == FCLB_SCENE_HEIGHTMAP) {
                const int la = a.hm1.n_layers - 1 - int((e.a.meta >> 8) & 0xffu), lb = a.hm2.n_layers - 1 - int((e.b.meta >> 8) & 0xffu);
                on_a = leaf_b || (!leaf_a && la < lb);
              } else if (KA == FCLB_SCENE_HEIGHTMAP) {
                // depth_from_root_1 as the reference carries it: +1 per map descent, and BACK TO 0 on every octree descent (the
                // pushed element leaves it at its default, :545-554); kept in bits 16.. of the map element's meta
                const int da = int(e.a.meta >> 16), db = int(e.b.meta >> 8);
                on_a = leaf_b || (!leaf_a && da < db);
              } else {
                const V3<S> da = boxMax(e.a) - boxMin(e.a), db = boxMax(e.b) - boxMin(e.b);
                const bool on_b = leaf_a || (!leaf_b && sqnorm(da) < sqnorm(db));
                on_a = !on_b;
              }
              if (e.shift < (on_a ? kBitsA : kBitsB)) {
                overflow = true;
              } else {
                child_mask = on_a ? sideA.childMask(e.a) : sideB.childMask(e.b);
                n_push = __popc(child_mask);
              }
            }
          }
        }
        if (__any_sync(0xffffffffu, overflow)) {
          overflow = true;
          break;
        }
        int push_off = n_push;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int p = __shfl_up_sync(0xffffffffu, push_off, o);
          if (lane >= o) push_off += p;
        }
        const int total_push = __shfl_sync(0xffffffffu, push_off, 31);
        push_off -= n_push;
        if (n_push) {
          int k = 0;
          const int bits = on_a ? kBitsA : kBitsB, n_child = on_a ? kChildA : kChildB;
          const int shift = e.shift - bits;
          #pragma unroll 1
          for (int c = 0; c < n_child; c++) {
            if (!(child_mask & (1u << c))) continue;
            CspElem<S> ch;
            if (on_a) {
              ch.a = sideA.child(e.a, c);
              ch.b = e.b;
              if (kRawB) ch.a.meta |= ((e.a.meta >> 16) + 1u) << 16;
            } else {
              ch.a = e.a;
              ch.b = sideB.child(e.b, c);
              if (kRawB) ch.a.meta &= 0xffffu;
            }
            ch.lo = iv.lo;
            ch.hi = iv.hi;
            ch.shift = shift;
            ch.path = e.path | ((unsigned long long)(n_child - 1 - c) << shift);  // the last child pushed is visited first
            stack[sp + push_off + k] = ch;
            k++;
          }
        }
        sp += total_push;
        __syncwarp();
      }
    }
    if (overflow && lane == 0) atomicAdd(a.cand_count + 1, 1u);
    __syncwarp();
  }
}

template <typename S>
__global__ void ccdScenePairSelectKernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ order, uint32_t n_cand,
                                         const long long* __restrict__ cand_ids, const S* __restrict__ cand_toc,
                                         const S* __restrict__ cand_box, uint32_t max_contacts, uint32_t keep,
                                         uint32_t* __restrict__ counts, long long* __restrict__ ids, S* __restrict__ toc,
                                         S* __restrict__ box) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n_cand) return;
  const uint32_t q = keys[i];
  auto lowerBound = [&](unsigned long long v) {
    size_t lo = 0, hi = n_cand;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if ((unsigned long long)keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const size_t first = lowerBound(q);
  const size_t k = i - first;
  if (k == 0) {
    const size_t cnt = lowerBound((unsigned long long)q + 1) - first;
    counts[q] = uint32_t(cnt < max_contacts ? cnt : max_contacts);
  }
  if (k < max_contacts && k < keep) {
    const uint32_t c = order[i];
    const size_t o = size_t(q) * keep + k;
    ids[2 * o] = cand_ids[2 * size_t(c)];
    ids[2 * o + 1] = cand_ids[2 * size_t(c) + 1];
    if (toc) {
      toc[2 * o] = cand_toc[2 * size_t(c)];
      toc[2 * o + 1] = cand_toc[2 * size_t(c) + 1];
    }
    if (box)
      for (int j = 0; j < 12; j++) box[12 * o + j] = cand_box[12 * size_t(c) + j];
  }
}

template <typename S>
static int ccdScenePairDev(Engine& e, int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                           const void* poses2, const void* disp, size_t n, const fclb_ccd_request* req, uint32_t keep,
                           uint32_t* counts, long long* ids, void* toc, void* box) {
  const int st = sizeof(S) == 4 ? 0 : 1;
  CcdScenePairArgs a{};
  // side A / B of the walk: the reference runs (octree, heightmap) as (heightmap, octree) with the displacement re-expressed
  a.swapped = kind1 == FCLB_SCENE_OCTREE && kind2 == FCLB_SCENE_HEIGHTMAP;
  const int ka = a.swapped ? kind2 : kind1, kb = a.swapped ? kind1 : kind2;
  const fclb_handle ha = a.swapped ? scene2 : scene1, hb = a.swapped ? scene1 : scene2;
  int rc = ka == FCLB_SCENE_HEIGHTMAP ? sceneHmView(ha, st, a.hm1) : sceneOctView(ha, a.oct1);
  if (rc) return rc;
  rc = kb == FCLB_SCENE_HEIGHTMAP ? sceneHmView(hb, st, a.hm2) : sceneOctView(hb, a.oct2);
  if (rc) return rc;
  a.poses_a = a.swapped ? poses2 : poses1;
  a.poses_b = a.swapped ? poses1 : poses2;
  a.disp = disp;
  a.n = n;
  a.zero_tol = req->zero_movement_tolerance > 0 ? req->zero_movement_tolerance : 1e-4;
  CcdSceneScratch sc;
  uint32_t* d_count = nullptr;
  unsigned long long* d_work = nullptr;
  FCLB_CUDA(sc.get(&d_count, 2));
  FCLB_CUDA(sc.get(&d_work, 1));
  a.cand_count = d_count;
  a.work_counter = d_work;
  const size_t smem = size_t(kCsWarps) * kCspStackCap * sizeof(CspElem<S>);
  void (*kern)(CcdScenePairArgs) = nullptr;
  if (ka == FCLB_SCENE_HEIGHTMAP && kb == FCLB_SCENE_HEIGHTMAP) kern = ccdScenePairKernel<S, FCLB_SCENE_HEIGHTMAP, FCLB_SCENE_HEIGHTMAP>;
  else if (ka == FCLB_SCENE_HEIGHTMAP) kern = ccdScenePairKernel<S, FCLB_SCENE_HEIGHTMAP, FCLB_SCENE_OCTREE>;
  else kern = ccdScenePairKernel<S, FCLB_SCENE_OCTREE, FCLB_SCENE_OCTREE>;
  FCLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int grid = int(std::min<size_t>((n + kCsWarps - 1) / kCsWarps, size_t(e.sms)));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  size_t cap = std::min<size_t>(std::max<size_t>(n * 64, size_t(1) << 16), size_t(1) << 25);
  uint32_t h_count[2] = {0, 0};
  S* cand_toc = nullptr;
  for (int attempt = 0; attempt < 3; attempt++) {
    FCLB_CUDA(sc.get(&a.cand_q, cap));
    FCLB_CUDA(sc.get(&a.cand_ids, 2 * cap));
    FCLB_CUDA(sc.get(&a.cand_path, cap));
    S* bx = nullptr;
    FCLB_CUDA(sc.get(&bx, 12 * cap));
    FCLB_CUDA(sc.get(&cand_toc, 2 * cap));
    a.cand_box = bx;
    a.cand_toc = cand_toc;
    a.cand_cap = uint32_t(cap);
    FCLB_CUDA(cudaMemsetAsync(d_count, 0, 2 * sizeof(uint32_t), e.compute));
    FCLB_CUDA(cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), e.compute));
    kern<<<grid, kCsWarps * 32, smem, e.compute>>>(a);
    FCLB_CUDA(cudaGetLastError());
    e.launches += 1;
    FCLB_CUDA(cudaMemcpyAsync(h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    if (h_count[1]) return fail(FCLB_ERR_CAPACITY, "scene-pair CCD: hierarchies wider or deeper than the per-warp stack / 64 path bits allow");
    if (h_count[0] <= cap) break;
    if (attempt == 2 || h_count[0] > (1u << 30)) return fail(FCLB_ERR_CAPACITY, "scene-pair CCD: too many contacts: split the batch");
    cap = h_count[0];
  }
  const uint32_t n_cand = h_count[0];
  FCLB_CUDA(cudaMemsetAsync(counts, 0, n * sizeof(uint32_t), e.compute));
  if (n * keep) {
    const int g = int((n * keep * 12 + 255) / 256);
    ccdSceneFillKernel<long long><<<g, 256, 0, e.compute>>>(ids, n * keep * 2, -1ll);
    if (toc) ccdSceneFillKernel<S><<<g, 256, 0, e.compute>>>(static_cast<S*>(toc), n * keep * 2, S(-1));
    if (box) ccdSceneFillKernel<S><<<g, 256, 0, e.compute>>>(static_cast<S*>(box), n * keep * 12, S(0));
  }
  if (n_cand) {
    unsigned long long* path_sorted = nullptr;
    uint32_t *order = nullptr, *order1 = nullptr, *order2 = nullptr, *qkey1 = nullptr, *qkey2 = nullptr;
    FCLB_CUDA(sc.get(&path_sorted, n_cand));
    FCLB_CUDA(sc.get(&order, n_cand));
    FCLB_CUDA(sc.get(&order1, n_cand));
    FCLB_CUDA(sc.get(&order2, n_cand));
    FCLB_CUDA(sc.get(&qkey1, n_cand));
    FCLB_CUDA(sc.get(&qkey2, n_cand));
    const int g256 = int((n_cand + 255) / 256);
    ccdSceneIotaKernel<<<g256, 256, 0, e.compute>>>(order, n_cand);
    size_t b1 = 0, b2 = 0;
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, a.cand_path, path_sorted, order, order1, int(n_cand), 0, 64, e.compute));
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b2, qkey1, qkey2, order1, order2, int(n_cand), 0, 32, e.compute));
    unsigned char* tmp = nullptr;
    FCLB_CUDA(sc.get(&tmp, std::max(b1, b2)));
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b1, a.cand_path, path_sorted, order, order1, int(n_cand), 0, 64, e.compute));
    ccdSceneGatherKernel<<<g256, 256, 0, e.compute>>>(a.cand_q, order1, n_cand, qkey1);
    FCLB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b2, qkey1, qkey2, order1, order2, int(n_cand), 0, 32, e.compute));
    ccdScenePairSelectKernel<S><<<g256, 256, 0, e.compute>>>(qkey2, order2, n_cand, a.cand_ids, cand_toc, static_cast<const S*>(a.cand_box),
                                                             req->max_contacts ? req->max_contacts : 1u, keep, counts, ids,
                                                             static_cast<S*>(toc), static_cast<S*>(box));
    FCLB_CUDA(cudaGetLastError());
    e.launches += 5;
  }
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -11;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

}  // namespace fclb

using namespace fclb;

extern "C" {

int fclb_translational_ccd_scene_batch_dev(int scene_kind, fclb_handle scene, fclb_handle shapes, const uint32_t* shape_ids,
                                           const void* poses_shape, const void* poses_scene, const void* displacements, size_t n,
                                           int scalar_type, const fclb_ccd_request* req, int scene_moves, uint32_t max_keep,
                                           uint32_t* out_counts, int64_t* out_code, void* out_toc, void* out_box) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  if (scene_kind != FCLB_SCENE_HEIGHTMAP && scene_kind != FCLB_SCENE_OCTREE)
    return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_batch: scene_kind must be FCLB_SCENE_HEIGHTMAP or FCLB_SCENE_OCTREE");
  ShapeTable* t = findTable(e, shapes);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_batch: unknown shape table handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!req || req->request_type > 2) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_batch: bad request");
  if (n == 0) return FCLB_OK;
  if (n > 0xfffffffeull) return fail(FCLB_ERR_CAPACITY, "batch larger than 2^32-2 queries: split it");
  if (!shape_ids || !poses_shape || !poses_scene || !displacements || !out_counts || (max_keep && !out_code))
    return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_batch: null array");
  #pragma unroll 1
  for (uint32_t i = 0; i < t->n; i++)
    if (t->host[i].type == FCLB_CONVEX && t->host[i].geom < e.convex.size()) {
      const int nv = e.convex[t->host[i].geom].n_verts;
      if (nv == 1 || nv == 2 || nv == 3 || nv == 6)
        return fail(FCLB_ERR_UNSUPPORTED, "scene CCD: Convex shapes with 1, 2, 3 or 6 vertices are not supported");
    }
  if (scalar_type == FCLB_F32)
    return ccdSceneDev<float>(e, scene_kind, scene, t, shape_ids, poses_shape, poses_scene, displacements, n, req, scene_moves,
                              max_keep, out_counts, reinterpret_cast<long long*>(out_code), out_toc, out_box);
  return ccdSceneDev<double>(e, scene_kind, scene, t, shape_ids, poses_shape, poses_scene, displacements, n, req, scene_moves,
                             max_keep, out_counts, reinterpret_cast<long long*>(out_code), out_toc, out_box);
}

static int translational_ccd_scene_batch_host_one(int scene_kind, fclb_handle scene, fclb_handle shapes, const uint32_t* shape_ids,
                                                  const void* poses_shape, const void* poses_scene, const void* displacements,
                                                  size_t n, int scalar_type, const fclb_ccd_request* req, int scene_moves,
                                                  uint32_t max_keep, uint32_t* out_counts, int64_t* out_code, void* out_toc,
                                                  void* out_box) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses_shape || !poses_scene || !displacements || !out_counts)
    return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_batch: null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  {
    ShapeTable* t = findTable(e, shapes);
    if (!t) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_batch: unknown shape table handle");
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
  const size_t o_code = alignUp(o_cnt + n * 4, 256);
  const size_t o_toc = alignUp(o_code + n * max_keep * 8, 256);
  const size_t o_box = alignUp(o_toc + n * max_keep * 2 * ss, 256);
  const size_t total = alignUp(o_box + n * max_keep * 6 * ss, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_ids, shape_ids, n * 4, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses_shape, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses_scene, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_d, displacements, n * 4 * ss, cudaMemcpyHostToDevice, e.compute));
  rc = fclb_translational_ccd_scene_batch_dev(scene_kind, scene, shapes, reinterpret_cast<const uint32_t*>(base + o_ids), base + o_p1,
                                              base + o_p2, base + o_d, n, scalar_type, req, scene_moves, max_keep,
                                              reinterpret_cast<uint32_t*>(base + o_cnt), reinterpret_cast<int64_t*>(base + o_code),
                                              out_toc ? base + o_toc : nullptr, out_box ? base + o_box : nullptr);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_counts, base + o_cnt, n * 4, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_code) FCLB_CUDA(cudaMemcpyAsync(out_code, base + o_code, n * max_keep * 8, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_toc) FCLB_CUDA(cudaMemcpyAsync(out_toc, base + o_toc, n * max_keep * 2 * ss, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_box) FCLB_CUDA(cudaMemcpyAsync(out_box, base + o_box, n * max_keep * 6 * ss, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}

int fclb_translational_ccd_scene_batch_host(int scene_kind, fclb_handle scene, fclb_handle shapes, const uint32_t* shape_ids,
                                            const void* poses_shape, const void* poses_scene, const void* displacements, size_t n,
                                            int scalar_type, const fclb_ccd_request* req, int scene_moves, uint32_t max_keep,
                                            uint32_t* out_counts, int64_t* out_code, void* out_toc, void* out_box) {
  if (engineCount() <= 1)
    return translational_ccd_scene_batch_host_one(scene_kind, scene, shapes, shape_ids, poses_shape, poses_scene, displacements, n,
                                                  scalar_type, req, scene_moves, max_keep, out_counts, out_code, out_toc, out_box);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  return shardOverDevices(n, [&](size_t b, size_t m_) {
    return translational_ccd_scene_batch_host_one(scene_kind, scene, shapes, offT(shape_ids, b), offPtr(poses_shape, b * 12 * ss),
                                                  offPtr(poses_scene, b * 12 * ss), offPtr(displacements, b * 4 * ss), m_, scalar_type,
                                                  req, scene_moves, max_keep, offT(out_counts, b), offT(out_code, b * max_keep),
                                                  offPtr(out_toc, b * max_keep * 2 * ss), offPtr(out_box, b * max_keep * 6 * ss));
  });
}

int fclb_translational_ccd_scene_mesh_batch_dev(int scene_kind, fclb_handle scene, fclb_handle bvh, const void* poses_scene,
                                                const void* poses_mesh, const void* displacements, size_t n, int scalar_type,
                                                const fclb_ccd_request* req, int mesh_moves, uint32_t max_keep,
                                                uint32_t* out_counts, int64_t* out_ids, void* out_toc, void* out_box) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  if (scene_kind != FCLB_SCENE_HEIGHTMAP && scene_kind != FCLB_SCENE_OCTREE)
    return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_mesh_batch: scene_kind must be FCLB_SCENE_HEIGHTMAP or FCLB_SCENE_OCTREE");
  auto it = bvhTable().find(bvh);
  if (it == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_mesh_batch: unknown BVH handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (it->second->scalar_type != scalar_type) return fail(FCLB_ERR_BAD_ARG, "BVH was uploaded for a different scalar type");
  if (!req || req->request_type > 2) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_mesh_batch: bad request");
  if (n == 0) return FCLB_OK;
  if (n > 0xfffffffeull) return fail(FCLB_ERR_CAPACITY, "batch larger than 2^32-2 queries: split it");
  if (!poses_scene || !poses_mesh || !displacements || !out_counts || (max_keep && !out_ids))
    return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_mesh_batch: null array");
  if (scalar_type == FCLB_F32)
    return ccdSceneMeshDev<float>(e, scene_kind, scene, it->second, poses_scene, poses_mesh, displacements, n, req, mesh_moves, max_keep,
                                  out_counts, reinterpret_cast<long long*>(out_ids), out_toc, out_box);
  return ccdSceneMeshDev<double>(e, scene_kind, scene, it->second, poses_scene, poses_mesh, displacements, n, req, mesh_moves, max_keep,
                                 out_counts, reinterpret_cast<long long*>(out_ids), out_toc, out_box);
}

static int translational_ccd_scene_mesh_batch_host_one(int scene_kind, fclb_handle scene, fclb_handle bvh, const void* poses_scene,
                                                       const void* poses_mesh, const void* displacements, size_t n, int scalar_type,
                                                       const fclb_ccd_request* req, int mesh_moves, uint32_t max_keep,
                                                       uint32_t* out_counts, int64_t* out_ids, void* out_toc, void* out_box) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!poses_scene || !poses_mesh || !displacements || !out_counts) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_mesh_batch: null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_p1 = 0;
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_d = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_d + n * 4 * ss, 256);
  const size_t o_ids = alignUp(o_cnt + n * 4, 256);
  const size_t o_toc = alignUp(o_ids + n * max_keep * 16, 256);
  const size_t o_box = alignUp(o_toc + n * max_keep * 2 * ss, 256);
  const size_t total = alignUp(o_box + n * max_keep * 6 * ss, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses_scene, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses_mesh, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_d, displacements, n * 4 * ss, cudaMemcpyHostToDevice, e.compute));
  rc = fclb_translational_ccd_scene_mesh_batch_dev(scene_kind, scene, bvh, base + o_p1, base + o_p2, base + o_d, n, scalar_type, req,
                                                   mesh_moves, max_keep, reinterpret_cast<uint32_t*>(base + o_cnt),
                                                   reinterpret_cast<int64_t*>(base + o_ids), out_toc ? base + o_toc : nullptr,
                                                   out_box ? base + o_box : nullptr);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_counts, base + o_cnt, n * 4, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_ids) FCLB_CUDA(cudaMemcpyAsync(out_ids, base + o_ids, n * max_keep * 16, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_toc) FCLB_CUDA(cudaMemcpyAsync(out_toc, base + o_toc, n * max_keep * 2 * ss, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_box) FCLB_CUDA(cudaMemcpyAsync(out_box, base + o_box, n * max_keep * 6 * ss, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}

int fclb_translational_ccd_scene_mesh_batch_host(int scene_kind, fclb_handle scene, fclb_handle bvh, const void* poses_scene,
                                                 const void* poses_mesh, const void* displacements, size_t n, int scalar_type,
                                                 const fclb_ccd_request* req, int mesh_moves, uint32_t max_keep,
                                                 uint32_t* out_counts, int64_t* out_ids, void* out_toc, void* out_box) {
  if (engineCount() <= 1)
    return translational_ccd_scene_mesh_batch_host_one(scene_kind, scene, bvh, poses_scene, poses_mesh, displacements, n, scalar_type,
                                                       req, mesh_moves, max_keep, out_counts, out_ids, out_toc, out_box);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  return shardOverDevices(n, [&](size_t b, size_t m_) {
    return translational_ccd_scene_mesh_batch_host_one(scene_kind, scene, bvh, offPtr(poses_scene, b * 12 * ss),
                                                       offPtr(poses_mesh, b * 12 * ss), offPtr(displacements, b * 4 * ss), m_,
                                                       scalar_type, req, mesh_moves, max_keep, offT(out_counts, b),
                                                       offT(out_ids, b * max_keep * 2), offPtr(out_toc, b * max_keep * 2 * ss),
                                                       offPtr(out_box, b * max_keep * 6 * ss));
  });
}

int fclb_translational_ccd_scene_pair_batch_dev(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                                const void* poses2, const void* displacements, size_t n, int scalar_type,
                                                const fclb_ccd_request* req, uint32_t max_keep, uint32_t* out_counts,
                                                int64_t* out_ids, void* out_toc, void* out_box) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  #pragma unroll 1
  for (int k : {kind1, kind2})
    if (k != FCLB_SCENE_HEIGHTMAP && k != FCLB_SCENE_OCTREE)
      return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_pair_batch: kinds must be FCLB_SCENE_HEIGHTMAP or FCLB_SCENE_OCTREE");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!req || req->request_type > 2) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_pair_batch: bad request");
  if (n == 0) return FCLB_OK;
  if (n > 0xfffffffeull) return fail(FCLB_ERR_CAPACITY, "batch larger than 2^32-2 queries: split it");
  if (!poses1 || !poses2 || !displacements || !out_counts || (max_keep && !out_ids))
    return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_pair_batch: null array");
  if (scalar_type == FCLB_F32)
    return ccdScenePairDev<float>(e, kind1, scene1, kind2, scene2, poses1, poses2, displacements, n, req, max_keep, out_counts,
                                  reinterpret_cast<long long*>(out_ids), out_toc, out_box);
  return ccdScenePairDev<double>(e, kind1, scene1, kind2, scene2, poses1, poses2, displacements, n, req, max_keep, out_counts,
                                 reinterpret_cast<long long*>(out_ids), out_toc, out_box);
}

static int translational_ccd_scene_pair_batch_host_one(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                                       const void* poses2, const void* displacements, size_t n, int scalar_type,
                                                       const fclb_ccd_request* req, uint32_t max_keep, uint32_t* out_counts,
                                                       int64_t* out_ids, void* out_toc, void* out_box) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!poses1 || !poses2 || !displacements || !out_counts) return fail(FCLB_ERR_BAD_ARG, "fclb_translational_ccd_scene_pair_batch: null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_p1 = 0;
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_d = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_d + n * 4 * ss, 256);
  const size_t o_ids = alignUp(o_cnt + n * 4, 256);
  const size_t o_toc = alignUp(o_ids + n * max_keep * 16, 256);
  const size_t o_box = alignUp(o_toc + n * max_keep * 2 * ss, 256);
  const size_t total = alignUp(o_box + n * max_keep * 12 * ss, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses1, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses2, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_d, displacements, n * 4 * ss, cudaMemcpyHostToDevice, e.compute));
  rc = fclb_translational_ccd_scene_pair_batch_dev(kind1, scene1, kind2, scene2, base + o_p1, base + o_p2, base + o_d, n, scalar_type,
                                                   req, max_keep, reinterpret_cast<uint32_t*>(base + o_cnt),
                                                   reinterpret_cast<int64_t*>(base + o_ids), out_toc ? base + o_toc : nullptr,
                                                   out_box ? base + o_box : nullptr);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_counts, base + o_cnt, n * 4, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_ids) FCLB_CUDA(cudaMemcpyAsync(out_ids, base + o_ids, n * max_keep * 16, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_toc) FCLB_CUDA(cudaMemcpyAsync(out_toc, base + o_toc, n * max_keep * 2 * ss, cudaMemcpyDeviceToHost, e.compute));
  if (max_keep && out_box) FCLB_CUDA(cudaMemcpyAsync(out_box, base + o_box, n * max_keep * 12 * ss, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}

int fclb_translational_ccd_scene_pair_batch_host(int kind1, fclb_handle scene1, int kind2, fclb_handle scene2, const void* poses1,
                                                 const void* poses2, const void* displacements, size_t n, int scalar_type,
                                                 const fclb_ccd_request* req, uint32_t max_keep, uint32_t* out_counts,
                                                 int64_t* out_ids, void* out_toc, void* out_box) {
  if (engineCount() <= 1)
    return translational_ccd_scene_pair_batch_host_one(kind1, scene1, kind2, scene2, poses1, poses2, displacements, n, scalar_type, req,
                                                       max_keep, out_counts, out_ids, out_toc, out_box);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  return shardOverDevices(n, [&](size_t b, size_t m_) {
    return translational_ccd_scene_pair_batch_host_one(kind1, scene1, kind2, scene2, offPtr(poses1, b * 12 * ss),
                                                       offPtr(poses2, b * 12 * ss), offPtr(displacements, b * 4 * ss), m_, scalar_type,
                                                       req, max_keep, offT(out_counts, b), offT(out_ids, b * max_keep * 2),
                                                       offPtr(out_toc, b * max_keep * 2 * ss), offPtr(out_box, b * max_keep * 12 * ss));
  });
}

}  // extern "C"
