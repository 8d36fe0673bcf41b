// fclb_collide_impl.cuh -- batched fcl::collide for shape-shape pairs and the
// direct GJK+EPA path.
//
// Reference path: fcl::collide -> ShapeShapeCollide (collision_func_matrix-inl.h:340)
// -> ShapePairIntersectSolver::ShapeIntersect (shape_pair_intersect-inl.h:49-116)
// -> GJKSolver::shapeIntersect: closed form for the registered pairs
// (gjk_solver-inl.h:152-243), else MPR (no contact wanted) / GJK (+EPA)
// (gjk_solver-inl.h:71-140).
//
// Three kernel families, one launch per (type1,type2) bucket:
//   collideClosedKernel  thread per query   closed-form pairs (HBM bound)
//   convexBoolKernel     thread per query   MPR and/or GJK boolean; colliding
//                                           queries that need a contact append
//                                           their GJK simplex to a work list
//   epaKernel            WARP per query     EPA on the work list (fclb_epa.cuh)
#pragma once
#include "fclb_boxbox.cuh"
#include "fclb_epa.cuh"
#include "fclb_internal.h"
#include "fclb_mpr.cuh"
#include "fclb_mpr_pen.cuh"
#include "fclb_sphere_triangle.cuh"

namespace fclb {

struct CollideOut {
  void* contacts;    // max_keep x 9 S per query, or nullptr
  uint32_t* counts;  // numContacts per query (collide API) -- may be nullptr in gjk_epa mode
  uint32_t max_keep;
  uint32_t max_contacts;
  int penetration;  // 0 disabled, 1 default GJK/EPA
  // gjk_epa API outputs
  int32_t* gjk_status;
  int32_t* epa_status;
  void* geom;  // 7 S per query {depth, p0, p1}
};

// EPA work list: colliding queries + their GJK simplices
struct EpaWork {
  uint32_t* count;   // device counter
  uint32_t* query;   // [capacity] query index
  void* simplex;     // [capacity] x 24 S (slot s: vertex 3, direction 3)
  int32_t* rank;     // [capacity]
  uint32_t capacity;
};

template <typename S>
FCLB_DI void writeContact(const CollideOut& o, size_t q, uint32_t k, const ContactPt<S>& c) {
  if (!o.contacts || k >= o.max_keep) return;
  S* p = static_cast<S*>(o.contacts) + (q * o.max_keep + k) * 9;
  p[0] = S(-1);
  p[1] = S(-1);
  p[2] = c.normal.x; p[3] = c.normal.y; p[4] = c.normal.z;
  p[5] = c.pos.x; p[6] = c.pos.y; p[7] = c.pos.z;
  p[8] = c.depth;
}
template <typename S>
FCLB_DI void clearContacts(const CollideOut& o, size_t q, uint32_t from) {
  if (!o.contacts) return;
  #pragma unroll 1
  for (uint32_t k = from; k < o.max_keep; k++) {
    S* p = static_cast<S*>(o.contacts) + (q * o.max_keep + k) * 9;
#pragma unroll
    for (int j = 0; j < 9; j++) p[j] = S(0);
  }
}

// std::partial_sort(first, first+k, last, comp) exactly as libstdc++ implements it
// (__heap_select + __sort_heap), on an index array, comp(a,b) = depth[a] > depth[b]
// (shape_pair_intersect-inl.h:96-106 keeps the k deepest contacts; which of two
// equally deep contacts survives is decided by this algorithm).
template <typename S>
struct SmallPartialSort {
  int idx[8];
  const S* depth;
  FCLB_DI bool comp(int a, int b) const { return depth[b] < depth[a]; }
  FCLB_DI void pushHeap(int hole, int top, int value) {
    int parent = (hole - 1) / 2;
    while (hole > top && comp(idx[parent], value)) {
      idx[hole] = idx[parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    idx[hole] = value;
  }
  FCLB_DI void adjustHeap(int hole, int len, int value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (comp(idx[child], idx[child - 1])) child--;
      idx[hole] = idx[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      idx[hole] = idx[child - 1];
      hole = child - 1;
    }
    pushHeap(hole, top, value);
  }
  FCLB_DI void run(int n, int k) {
    // __make_heap(first, middle)
    if (k >= 2) {
      int parent = (k - 2) / 2;
      while (true) {
        const int value = idx[parent];
        adjustHeap(parent, k, value);
        if (parent == 0) break;
        parent--;
      }
    }
    #pragma unroll 1
    for (int i = k; i < n; i++) {
      if (comp(idx[i], idx[0])) {  // __pop_heap(first, middle, i)
        const int value = idx[i];
        idx[i] = idx[0];
        adjustHeap(0, k, value);
      }
    }
    // __sort_heap(first, middle)
    int last = k;
    while (last > 1) {
      --last;
      const int value = idx[last];
      idx[last] = idx[0];
      adjustHeap(0, last, value);
    }
  }
};

// Emit the result of one shape pair the way ShapeIntersect does (:49-116).
template <typename S>
FCLB_DI void emitContacts(const CollideOut& o, size_t q, bool hit, const ContactPt<S>* cps, int n) {
  if (!hit || o.max_contacts == 0) {
    if (o.counts) o.counts[q] = 0;
    clearContacts<S>(o, q, 0);
    return;
  }
  if (!o.penetration) {
    if (o.counts) o.counts[q] = 1;
    clearContacts<S>(o, q, 0);
    if (o.contacts && o.max_keep > 0) {
      S* p = static_cast<S*>(o.contacts) + (q * o.max_keep) * 9;
      p[0] = S(-1);
      p[1] = S(-1);
    }
    return;
  }
  const uint32_t free_space = o.max_contacts;
  uint32_t adding = uint32_t(n);
  if (free_space < uint32_t(n)) {
    S depth[8];
    SmallPartialSort<S> ps;
    #pragma unroll 1
    for (int i = 0; i < n; i++) {
      depth[i] = cps[i].depth;
      ps.idx[i] = i;
    }
    ps.depth = depth;
    ps.run(n, int(free_space));
    adding = free_space;
    #pragma unroll 1
    for (uint32_t k = 0; k < adding; k++) writeContact(o, q, k, cps[ps.idx[k]]);
  } else {
    #pragma unroll 1
    for (uint32_t k = 0; k < adding; k++) writeContact(o, q, k, cps[k]);
  }
  if (o.counts) o.counts[q] = adding;
  clearContacts<S>(o, q, adding);
}

enum ClosedCollide : int {
  CC_NONE = 0,
  CC_SPHERE_SPHERE,
  CC_SPHERE_CAPSULE,
  CC_CAPSULE_SPHERE,
  CC_SPHERE_BOX,
  CC_BOX_SPHERE,
  CC_SPHERE_CYLINDER,
  CC_CYLINDER_SPHERE,
  CC_BOX_BOX,
  CC_SPHERE_TRIANGLE  // leaf batches only: ShapeTransformedTriangleIntersectIndepImpl<S, Sphere<S>> (gjk_solver-inl.h:570-581)
};
inline int closedCollideOf(int t1, int t2) {
  if (t1 == ST_SPHERE && t2 == ST_SPHERE) return CC_SPHERE_SPHERE;
  if (t1 == ST_SPHERE && t2 == ST_CAPSULE) return CC_SPHERE_CAPSULE;
  if (t1 == ST_CAPSULE && t2 == ST_SPHERE) return CC_CAPSULE_SPHERE;
  if (t1 == ST_SPHERE && t2 == ST_BOX) return CC_SPHERE_BOX;
  if (t1 == ST_BOX && t2 == ST_SPHERE) return CC_BOX_SPHERE;
  if (t1 == ST_SPHERE && t2 == ST_CYLINDER) return CC_SPHERE_CYLINDER;
  if (t1 == ST_CYLINDER && t2 == ST_SPHERE) return CC_CYLINDER_SPHERE;
  if (t1 == ST_BOX && t2 == ST_BOX) return CC_BOX_BOX;
  if (t1 == ST_SPHERE && t2 == ST_TRIANGLE) return CC_SPHERE_TRIANGLE;
  return CC_NONE;
}

template <typename S, int CC>
__global__ void __launch_bounds__(kBlock) collideClosedKernel(BatchView b, CollideOut out) {
  const ShapeD<S>* __restrict__ shapes = static_cast<const ShapeD<S>*>(b.shapes);
  const S* __restrict__ poses1 = static_cast<const S*>(b.poses1);
  const S* __restrict__ poses2 = static_cast<const S*>(b.poses2);
  const bool want = out.penetration != 0;
  #pragma unroll 1
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < b.count; i += size_t(gridDim.x) * blockDim.x) {
    const size_t q = b.perm ? size_t(b.perm[b.begin + i]) : (b.begin + i);
    const fclb_pair pr = b.pairs[q];
    const ShapeD<S> a = shapes[pr.shape1];
    const ShapeD<S> c = shapes[pr.shape2];
    const Pose<S> tf1 = loadPose(poses1, q);
    const Pose<S> tf2 = loadPose(poses2, q);
    ContactPt<S> cp[4];
    int n = 1;
    bool hit = false;
    bool flip = false;
    if (CC == CC_SPHERE_SPHERE) {
      hit = sphereSphereIntersect(a.p[0], tf1, c.p[0], tf2, want, cp[0]);
    } else if (CC == CC_SPHERE_CAPSULE) {
      hit = sphereCapsuleIntersect(a.p[0], tf1, c.p[0], c.p[1], tf2, want, cp[0]);
    } else if (CC == CC_CAPSULE_SPHERE) {
      hit = sphereCapsuleIntersect(c.p[0], tf2, a.p[0], a.p[1], tf1, want, cp[0]);
      flip = true;
    } else if (CC == CC_SPHERE_BOX) {
      hit = sphereBoxIntersect(a.p[0], tf1, mk<S>(c.p[0], c.p[1], c.p[2]), tf2, want, cp[0]);
    } else if (CC == CC_BOX_SPHERE) {
      hit = sphereBoxIntersect(c.p[0], tf2, mk<S>(a.p[0], a.p[1], a.p[2]), tf1, want, cp[0]);
      flip = true;
    } else if (CC == CC_SPHERE_CYLINDER) {
      hit = sphereCylinderIntersect(a.p[0], tf1, c.p[0], c.p[1], tf2, want, cp[0]);
    } else if (CC == CC_CYLINDER_SPHERE) {
      hit = sphereCylinderIntersect(c.p[0], tf2, a.p[0], a.p[1], tf1, want, cp[0]);
      flip = true;
    } else if (CC == CC_BOX_BOX) {
      const int code = boxBox2(mk<S>(a.p[0], a.p[1], a.p[2]), tf1, mk<S>(c.p[0], c.p[1], c.p[2]), tf2, cp, &n);
      hit = code != 0;
    } else if (CC == CC_SPHERE_TRIANGLE) {
      const S* t = static_cast<const S*>(b.tris) + size_t(12) * size_t(c.geom);
      hit = sphereTriangleContact(a.p[0], tf1.t, apply(tf2, mk<S>(t[0], t[1], t[2])), apply(tf2, mk<S>(t[4], t[5], t[6])),
                                  apply(tf2, mk<S>(t[8], t[9], t[10])), cp[0]);
    }
    if (flip && want && hit) cp[0].normal = -cp[0].normal;  // flipNormal (gjk_solver-inl.h:186-197)
    emitContacts<S>(out, q, hit, cp, n);
  }
}

// ---- box-box, two phases ---------------------------------------------------------------
// boxBox2 is a 15-axis SAT followed by ~300 lines of contact generation that only the colliding pairs reach
// (ncu on the one-phase kernel: 4.6 of 32 lanes per issued instruction).  Phase 1 runs the SAT for every query
// of a block-sized chunk and queues the colliding ones in shared memory; phase 2 runs the complete routine on
// the compacted queue, a full block at a time.  Without contacts requested the SAT result is the answer
// (boxBoxIntersect returns return_code != 0, box_box-inl.h:824-846) and phase 2 is skipped.
template <typename S>
__global__ void __launch_bounds__(kBlock) boxBoxCollideKernel(BatchView b, CollideOut out) {
  __shared__ uint32_t queue[2 * kBlock];
  __shared__ int qn;
  const ShapeD<S>* __restrict__ shapes = static_cast<const ShapeD<S>*>(b.shapes);
  const S* __restrict__ poses1 = static_cast<const S*>(b.poses1);
  const S* __restrict__ poses2 = static_cast<const S*>(b.poses2);
  const bool want = out.penetration != 0 && out.max_contacts != 0;
  if (threadIdx.x == 0) qn = 0;
  __syncthreads();
  auto phase2 = [&](int first, int count) {
    if (int(threadIdx.x) < count) {
      const size_t q = queue[first + threadIdx.x];
      const fclb_pair pr = b.pairs[q];
      const ShapeD<S> a = shapes[pr.shape1];
      const ShapeD<S> c = shapes[pr.shape2];
      ContactPt<S> cp[4];
      int n = 1;
      const int code = boxBox2<S>(mk<S>(a.p[0], a.p[1], a.p[2]), loadPose(poses1, q), mk<S>(c.p[0], c.p[1], c.p[2]),
                                  loadPose(poses2, q), cp, &n);
      emitContacts<S>(out, q, code != 0, cp, n);
    }
  };
  const size_t chunk = size_t(gridDim.x) * blockDim.x;
  const size_t rounds = (b.count + chunk - 1) / chunk;
  #pragma unroll 1
  for (size_t r = 0; r < rounds; r++) {
    const size_t i = r * chunk + blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    bool hit = false;
    size_t q = 0;
    if (i < b.count) {
      q = b.perm ? size_t(b.perm[b.begin + i]) : (b.begin + i);
      const fclb_pair pr = b.pairs[q];
      const ShapeD<S> a = shapes[pr.shape1];
      const ShapeD<S> c = shapes[pr.shape2];
      ContactPt<S> cp[4];
      int n = 0;
      const int code = boxBox2<S, true>(mk<S>(a.p[0], a.p[1], a.p[2]), loadPose(poses1, q), mk<S>(c.p[0], c.p[1], c.p[2]),
                                        loadPose(poses2, q), cp, &n);
      hit = code != 0;
      if (!hit || !want) emitContacts<S>(out, q, hit, cp, 0);  // final without contacts
    }
    if (want) {
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      int base = 0;
      if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(&qn, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (hit) queue[base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = uint32_t(q);
      __syncthreads();
      if (qn >= kBlock) {
        const int first = qn - kBlock;
        phase2(first, kBlock);
        __syncthreads();
        if (threadIdx.x == 0) qn = first;
        __syncthreads();
      }
    }
  }
  if (want) {
    __syncthreads();
    phase2(0, qn);
  }
}

// ---- generic convex pairs: boolean stage -------------------------------------
// mode bit 0: run MPR first (collide API without penetration, gjk_solver-inl.h:88-98)
// mode bit 1: colliding queries go to the EPA work list
// mode bit 2: gjk_epa API (write GJK status instead of counts)
template <typename S, int T0, int T1>
__global__ void __launch_bounds__(kBlock) convexBoolKernel(BatchView b, S tol, int max_iter, int mode, CollideOut out,
                                                           EpaWork work) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SlotStore<S> st;
  st.base = reinterpret_cast<S*>(smem_raw) + threadIdx.x;
  st.stride = blockDim.x;
  const ShapeD<S>* __restrict__ shapes = static_cast<const ShapeD<S>*>(b.shapes);
  const ConvexD<S>* __restrict__ cvx = static_cast<const ConvexD<S>*>(b.convex);
  const S* __restrict__ poses1 = static_cast<const S*>(b.poses1);
  const S* __restrict__ poses2 = static_cast<const S*>(b.poses2);
  #pragma unroll 1
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < b.count; i += size_t(gridDim.x) * blockDim.x) {
    const size_t q = b.perm ? size_t(b.perm[b.begin + i]) : (b.begin + i);
    const fclb_pair pr = b.pairs[q];
    MinkDiff<S, T0, T1> md;
    md.s0 = bindShape(shapes, cvx, pr.shape1, static_cast<const S*>(b.tris));
    md.s1 = bindShape(shapes, cvx, pr.shape2, static_cast<const S*>(b.tris));
    md.setPoses(loadPose(poses1, q), loadPose(poses2, q));
    int decided = -1;  // -1 undecided, 0 no collision, 1 collision
    if (mode & 1) {
      const int ms = mprIntersect(md, max_iter, tol, nullptr);
      if (ms == MPR_INTERSECT) decided = 1;
      if (ms == MPR_SEPARATED) decided = 0;
    }
    int gs = -1;
    Simp simplex;
    simplex.ord = 0;
    simplex.rank = -1;
    if (decided < 0) {
      gs = gjkEvaluate<S>(md, st, simplex, mk<S>(S(-1), S(0), S(0)), tol, max_iter, nullptr, nullptr);
      decided = (gs == GJK_INTERSECT) ? 1 : 0;
    }
    if (mode & 4) {
      out.gjk_status[q] = gs;
      if (out.epa_status) out.epa_status[q] = -1;
      if (out.geom) {
        S* g = static_cast<S*>(out.geom) + 7 * q;
#pragma unroll
        for (int j = 0; j < 7; j++) g[j] = S(0);
      }
    }
    if (decided == 1 && (mode & 2)) {
      const uint32_t w = atomicAdd(work.count, 1u);
      if (w < work.capacity) {
        work.query[w] = uint32_t(q);
        work.rank[w] = simplex.rank;
        S* sp = static_cast<S*>(work.simplex) + size_t(w) * 24;
        for (int k = 0; k < 4; k++) {
          const int slot = (k < simplex.rank) ? slotOf(simplex, k) : 0;
          const V3<S> v = st.vtx(slot), d = st.dir(slot);
          sp[6 * k + 0] = v.x; sp[6 * k + 1] = v.y; sp[6 * k + 2] = v.z;
          sp[6 * k + 3] = d.x; sp[6 * k + 4] = d.y; sp[6 * k + 5] = d.z;
        }
      }
    } else if (!(mode & 4)) {
      // collide API, final answer without a contact computation
      ContactPt<S> dummy;
      dummy.normal = zero3<S>();
      dummy.pos = zero3<S>();
      dummy.depth = S(0);
      emitContacts<S>(out, q, decided == 1 && !out.penetration, &dummy, 0);
    }
  }
}

// ---- EPA stage: one TILE of T lanes per work item -----------------------------
// Tier 1: T = 8 (four queries per warp) with a small pool; a query whose polytope
// outgrows the pool is appended to `defer` and re-run from scratch by tier 2
// (T = 32, the reference's capacity), so the reported status is always the one the
// reference's pool size produces.
#ifdef FCLB_EPA_MIN_BLOCKS
#define FCLB_EPA_BOUNDS_TAIL , FCLB_EPA_MIN_BLOCKS
#else
#define FCLB_EPA_BOUNDS_TAIL
#endif
#ifndef FCLB_EPA_THREADS
#define FCLB_EPA_THREADS 128
#endif
// FCLB_EPA_CTA_SYNC: every warp of the CTA starts its lockstep iteration together (one __syncthreads per iteration), so
// the instruction lines one warp fetches are the ones the others need next -- the kernel is instruction-fetch bound
// (ncu: 9.8 issue slots lost to stall_no_instruction per issued instruction with free-running warps).
#ifndef FCLB_EPA_CTA_SYNC
#define FCLB_EPA_CTA_SYNC 0
#endif
constexpr int kEpaThreads = FCLB_EPA_THREADS;
template <typename S>
__host__ __device__ inline size_t epaTileBytes(size_t poly_bytes) {
  return (poly_bytes + 24 * sizeof(S) + 16 + 127) / 128 * 128 + 32;
}

struct EpaDefer {
  uint32_t* count;  // device words: [0] deferred items, [1] tier-1 cursor, [2] tier-2 cursor, [3] early consumers' cursor, [4] tier 1 done
  uint32_t* cursor; // device work cursor of this launch (tiles fetch items dynamically: EPA run times vary 100x)
  uint32_t* item;   // work-list indices of the deferred queries; kEpaItemEmpty until written, kEpaItemTaken once a consumer owns it
  int enabled;      // tier 1: defer on pool exhaustion; tier 2: 0
  int consume;      // 1 tier 2: iterate the deferred list instead of the full work list; 2 early tier-2 consumer: runs BESIDE
                    // tier 1 on another stream and takes items as they are published (a query that runs to the iteration
                    // limit costs 255 iterations x 14 us = 3.5 ms of latency -- started inside tier 1's run time instead of after it)
};
constexpr uint32_t kEpaItemEmpty = 0xffffffffu, kEpaItemTaken = 0xfffffffeu;

template <typename S, int T0, int T1, int T>
__global__ void __launch_bounds__(kEpaThreads FCLB_EPA_BOUNDS_TAIL) epaKernel(BatchView b, S tol, int pool_faces, int max_iter, int mode,
                                                         CollideOut out, EpaWork work, EpaDefer defer,
                                                         size_t poly_bytes, int tier1_iters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kTiles = kEpaThreads / T;
  const int warp_lane = threadIdx.x & 31, tile = threadIdx.x / T, lane = threadIdx.x % T;
  // tiles of one warp sit 32 B apart modulo 128 B, so their (tile-uniform) accesses hit different banks
  const size_t per_tile = epaTileBytes<S>(poly_bytes);
  unsigned char* my = smem_raw + size_t(tile) * per_tile;
  SlotStore<S> st;
  st.base = reinterpret_cast<S*>(my);
  st.stride = 1;
  unsigned char* poly_mem = my + ((24 * sizeof(S) + 15) / 16 * 16);
  const ShapeD<S>* __restrict__ shapes = static_cast<const ShapeD<S>*>(b.shapes);
  const ConvexD<S>* __restrict__ cvx = static_cast<const ConvexD<S>*>(b.convex);
  const S* __restrict__ poses1 = static_cast<const S*>(b.poses1);
  const S* __restrict__ poses2 = static_cast<const S*>(b.poses2);
  const uint32_t n_items = defer.consume ? *defer.count : min(*work.count, work.capacity);
  // The tiles of a warp take their EPA iterations in LOCKSTEP: a warp-uniform loop whose body is "tiles
  // without a query fetch one and build its polytope; full-warp barrier; every tile with a query runs one
  // iteration".  (With a plain per-tile `for each query { evaluate }` the tiles drift apart after their first
  // query and never reconverge: ncu showed 13 of 32 lanes per issued instruction.)  Items come from a device-wide
  // cursor: a query that runs to the iteration limit costs 100x the median, a static split would leave its tile
  // the same share of the list as everyone else.
  bool more = true;
  bool active = false;
  size_t q = 0;
  uint32_t w = 0;
  int n_steps = 0;
  MinkDiff<S, T0, T1> md;
  Pose<S> tf1;
  EpaWarp<S, MinkDiff<S, T0, T1>, T> epa(md, poly_mem, pool_faces, defer.enabled != 0, warp_lane, nullptr);
  S depth = S(0);
  V3<S> p0 = zero3<S>(), p1 = zero3<S>();
  auto finish = [&](int es) {
    epa.sync();
    if (lane == 0) {
      if (defer.enabled && es == EPA_MALLOC_FAILED) {
        defer.item[atomicAdd(defer.count, 1u)] = w;
      } else if (mode & 4) {
        if (out.epa_status) out.epa_status[q] = es;
        if (out.geom) {
          S* g = static_cast<S*>(out.geom) + 7 * q;
          g[0] = depth;
          g[1] = p0.x; g[2] = p0.y; g[3] = p0.z;
          g[4] = p1.x; g[5] = p1.y; g[6] = p1.z;
        }
      } else {
        // gjk_solver-inl.h:117-133: contact in the world frame
        bool hit = false;
        ContactPt<S> c;
        c.normal = zero3<S>();
        c.pos = zero3<S>();
        c.depth = S(0);
        if (es != EPA_FAILED) {
          V3<S> n1 = p0 - p1;
          if (sqnorm(n1) <= S(0))
            n1 = mk<S>(S(0), S(0), S(1));
          else
            n1 = normalized(n1);
          const V3<S> pt1 = S(0.5) * (p0 + p1);
          c.pos = apply(tf1, pt1);
          c.normal = mulMV(tf1.R, n1);
          c.depth = depth;
          hit = true;
        }
        emitContacts<S>(out, q, hit, &c, 1);
      }
    }
    epa.sync();
  };
  while (true) {
    int es = epa.kEpaContinue;
    bool finished = false;
    uint32_t it = 0;
    if (!active && more) {
      if (!defer.consume) {
        if (lane == 0) it = atomicAdd(defer.cursor, 1u);
        it = epa.shfl(it, 0);
        more = it < n_items;
      } else {
        // deferred list: an item belongs to whoever swaps kEpaItemTaken into its slot (tier 2 and the early consumers share it)
        uint32_t got = kEpaItemEmpty;
        if (lane == 0) {
          volatile uint32_t* items = defer.item;
          volatile uint32_t* words = defer.count;
          uint32_t beat = words[1], idle = 0;
          while (true) {
            const uint32_t i = atomicAdd(defer.cursor, 1u);
            if (defer.consume == 1 && i >= n_items) break;
            uint32_t wv = items[i];
            while (defer.consume == 2 && wv == kEpaItemEmpty) {  // not published yet
              if (words[4]) {  // tier 1 has ended: whatever it deferred is visible now
                wv = items[i];
                break;
              }
              const uint32_t now = words[1];  // tier 1's cursor moves while it runs; if it stands still (tier 1 not co-scheduled,
              if (now != beat) {              // e.g. launches serialised by a profiler) give up -- tier 2 proper takes the rest
                beat = now;
                idle = 0;
              } else if (++idle > 4000u) {
                break;
              }
              __nanosleep(500);
              wv = items[i];
            }
            if (wv == kEpaItemEmpty) break;
            if (wv != kEpaItemTaken && atomicCAS(defer.item + i, wv, kEpaItemTaken) == wv) {
              got = wv;
              break;
            }
          }
        }
        got = epa.shfl(got, 0);
        more = got != kEpaItemEmpty;
        it = got;
      }
    }
    if (!active && more) {
      w = it;
      q = work.query[w];
      const fclb_pair pr = b.pairs[q];
      md.s0 = bindShape(shapes, cvx, pr.shape1, static_cast<const S*>(b.tris));
      md.s1 = bindShape(shapes, cvx, pr.shape2, static_cast<const S*>(b.tris));
      tf1 = loadPose(poses1, q);
      md.setPoses(tf1, loadPose(poses2, q));
      // GJK simplex -> slots 0..rank-1
      const S* sp = static_cast<const S*>(work.simplex) + size_t(w) * 24;
      for (int k = lane; k < 24; k += T) st.base[k] = sp[k];
      epa.sync();
      Simp sx;
      sx.rank = work.rank[w];
      sx.ord = 0x03020100u;
      depth = S(0);
      p0 = zero3<S>();
      p1 = zero3<S>();
      es = epa.begin(st, sx, tol, depth, p0, p1);
      n_steps = 0;
      if (es == epa.kEpaContinue)
        active = true;
      else
        finished = true;
    }
#if FCLB_EPA_CTA_SYNC
    if (!__syncthreads_or(active || finished || more)) break;
#else
    if (!__any_sync(0xffffffffu, active || finished || more)) break;
    __syncwarp();
#endif
    if (active) {
      es = epa.step(max_iter, tol, depth, p0, p1);
      if (es != epa.kEpaContinue) {
        finished = true;
        active = false;
      } else if (defer.enabled && ++n_steps >= tier1_iters) {
        // a query still running after tier1_iters iterations (0.003 % of box pairs, 0.07 % of the Convex pairs of C1b) most
        // likely runs to the iteration limit: hand it to tier 2 (which restarts it) instead of pinning this tile for 255 iterations
        es = EPA_MALLOC_FAILED;
        finished = true;
        active = false;
      }
    }
    if (finished) finish(es);  // (one call site: the result path is ~1k SASS instructions)
  }
}

// ---- MPR penetration stage (request modes DirectedPenetration / IncrementalMinimumPenetration) --------
// collisionPenetrationMPR (collision_penetration-inl.h:189-252): the boolean collide has already written
// counts[q]; every colliding shape pair gets one contact from computePenetrationMPR with
// MPR(128, request.distanceTolerance()).
template <typename S, int T0, int T1>
__global__ void __launch_bounds__(kBlock) mprPenetrationKernel(BatchView b, S tol, int incremental, S dx, S dy, S dz,
                                                               CollideOut out) {
  const ShapeD<S>* __restrict__ shapes = static_cast<const ShapeD<S>*>(b.shapes);
  const ConvexD<S>* __restrict__ cvx = static_cast<const ConvexD<S>*>(b.convex);
  const S* __restrict__ poses1 = static_cast<const S*>(b.poses1);
  const S* __restrict__ poses2 = static_cast<const S*>(b.poses2);
  const V3<S> dir_world = mk<S>(dx, dy, dz);
  #pragma unroll 1
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < b.count; i += size_t(gridDim.x) * blockDim.x) {
    const size_t q = b.perm ? size_t(b.perm[b.begin + i]) : (b.begin + i);
    if (out.counts[q] == 0) continue;
    const fclb_pair pr = b.pairs[q];
    MinkDiff<S, T0, T1> md;
    md.s0 = bindShape(shapes, cvx, pr.shape1);
    md.s1 = bindShape(shapes, cvx, pr.shape2);
    const Pose<S> tf1 = loadPose(poses1, q);
    md.setPoses(tf1, loadPose(poses2, q));
    ContactPt<S> cp;
    computePenetrationMpr<S>(md, tf1, dir_world, incremental != 0, 128, tol, cp.pos, cp.normal, cp.depth);
    writeContact<S>(out, q, 0, cp);
  }
}

struct CollideLaunchArgs {
  SolverParams sp;
  int mode;
  CollideOut out;
  EpaWork work;
  EpaDefer defer;
  int pen_mode = 0;  // FCLB_PEN_DIRECTED / FCLB_PEN_INCREMENTAL_MIN: run the MPR penetration stage after the boolean
  double pen_dir[3] = {0, 0, 0};
  const void* tris = nullptr;  // leaf batches: triangle array behind the table's ST_TRIANGLE entries
  cudaStream_t aux = nullptr;  // second stream + fork/join events for the early EPA tier-2 consumers (null: not used)
  cudaEvent_t ev_aux0 = nullptr, ev_aux1 = nullptr;
  size_t item_capacity = 0;    // entries behind defer.item
};

// implemented in fclb_collide_f32.cu / fclb_collide_f64.cu (MPR penetration stage of one bucket)
template <typename S>
cudaError_t launchMprPenetration(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st);

// implemented in fclb_epa_f32.cu / fclb_epa_f64.cu (EPA stage of one bucket)
template <typename S>
cudaError_t launchEpa(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st);
// implemented in fclb_collide_f32.cu / fclb_collide_f64.cu
template <typename S>
cudaError_t launchCollide(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st, int* n_launches);

}  // namespace fclb
