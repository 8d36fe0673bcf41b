// fclb_broadphase.cu -- device-side AABB-tree broadphase.
//
// Reference (results contract): detail::BinaryAABB_Tree<S, Alloc>
//   BuildTreeExternal      broadphase/binary_AABB_tree-inl.h:70-193   (median split, nth_element)
//   SelfCollision          :503-573    every pair of leaves whose AABBs overlap, once
//   TreeCollision          :443-501    every (leaf of this, leaf of tree2) with overlapping AABBs
//   SingleObjectCollision  :400-441    every leaf overlapping one query AABB
//   UpdateObjectAABB       :332-398    new leaf box, ancestors re-unioned
//   AABB<S>::overlap       math/bv/AABB-inl.h:82-88
//   CollisionObject<S>::computeAABB   narrowphase/collision_object-inl.h:141-154
// Node boxes are exact unions (min/max only) and the overlap test is a pure
// comparison, so the SET of reported pairs is a property of the leaf boxes alone:
// {(a,b) : AABB_a overlaps AABB_b}.  It does not depend on how the tree is split,
// which is what lets the device use its own tree: a linear BVH over 30-bit Morton
// codes of the box centres (radix sort, Karras-style hierarchy, bottom-up refit)
// and one thread per leaf for the overlap search.  What does depend on the
// reference's tree is the ORDER in which pairs are reported and, for
// SelfCollision, which of the two ids comes first; pairs are therefore returned
// as an unordered list and self pairs carry (lower sorted position, higher).
#include <cub/cub.cuh>

#include <cstring>
#include <unordered_map>
#include <vector>

#include "fclb_bound.h"
#include "fclb_engine.h"
#include "fclb_math.cuh"
#include "fclb_shapes.cuh"

namespace fclb {

template <typename S>
struct Box6 {
  S mn[3], mx[3];
};

// Traversal node: both children's boxes, links and last covered leaf in ONE record (64 B in float, 128 B
// in double), so a traversal step is one aligned fetch and tests two boxes.
template <typename S>
struct alignas(16) FatNode {
  S lb[6], rb[6];
  int lc, rc;  // >= 0 internal node, < 0 leaf ~index
  int ll, rl;  // last leaf (Morton position) under each child
};

struct BpTree {
  int n = 0;
  int cap = 0;                    // objects the device arrays were allocated for
  int scalar_type = 0;
  void* leaf_box = nullptr;       // Box6<S>[n] in Morton order
  uint64_t* leaf_id = nullptr;    // user ids in Morton order
  void* node_box = nullptr;       // Box6<S>[n-1]
  int2* node_child = nullptr;     // >= 0 internal node, < 0 leaf ~index
  int2* node_range = nullptr;     // first / last leaf covered
  int* parent = nullptr;          // [0, n-1): internal nodes, [n-1, 2n-1): leaves
  int* flags = nullptr;           // refit arrival counters
  void* fat = nullptr;            // FatNode<S>[n-1], rebuilt after every refit
  std::unordered_map<uint64_t, int> pos_of;  // user id -> Morton position (host)
};
static std::map<fclb_handle, BpTree*>& bpTable() {
  static std::map<fclb_handle, BpTree*> t[kMaxDevices];
  return t[currentSlot()];
}

__device__ __forceinline__ double atomicMinD(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) <= v) break;
    old = atomicCAS(a, assumed, __double_as_longlong(v));
  } while (assumed != old);
  return __longlong_as_double(old);
}
__device__ __forceinline__ double atomicMaxD(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) >= v) break;
    old = atomicCAS(a, assumed, __double_as_longlong(v));
  } while (assumed != old);
  return __longlong_as_double(old);
}

// bounds[0..2] = min of centres, bounds[3..5] = max of centres
template <typename S>
__global__ void bpBoundsKernel(const Box6<S>* __restrict__ box, int n, double* bounds) {
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  #pragma unroll 1
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double c = 0.5 * (double(box[i].mn[k]) + double(box[i].mx[k]));
      mn[k] = fmin(mn[k], c);
      mx[k] = fmax(mx[k], c);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], off));
      mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], off));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      atomicMinD(&bounds[k], mn[k]);
      atomicMaxD(&bounds[3 + k], mx[k]);
    }
  }
}

__device__ __forceinline__ uint32_t expandBits10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

// key = morton30(centre) << 32 | object index: unique, so the hierarchy needs no tie rule
template <typename S>
__global__ void bpMortonKernel(const Box6<S>* __restrict__ box, int n, const double* __restrict__ bounds,
                               unsigned long long* keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t code = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double c = 0.5 * (double(box[i].mn[k]) + double(box[i].mx[k]));
    const double ext = bounds[3 + k] - bounds[k];
    double u = ext > 0 ? (c - bounds[k]) / ext : 0.0;
    u = fmin(fmax(u * 1024.0, 0.0), 1023.0);
    code |= expandBits10(uint32_t(u)) << (2 - k);
  }
  keys[i] = (static_cast<unsigned long long>(code) << 32) | static_cast<unsigned>(i);
}

template <typename S>
__global__ void bpGatherKernel(const Box6<S>* __restrict__ box, const uint64_t* __restrict__ ids,
                               const unsigned long long* __restrict__ keys, int n, Box6<S>* leaf_box, uint64_t* leaf_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned src = unsigned(keys[i] & 0xffffffffull);
  leaf_box[i] = box[src];
  leaf_id[i] = ids[src];
}

__device__ __forceinline__ int bpDelta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  return __clzll(keys[i] ^ keys[j]);
}

// one thread per internal node (Karras 2012): range, split, children, parents
__global__ void bpHierarchyKernel(const unsigned long long* __restrict__ keys, int n, int2* node_child, int2* node_range,
                                  int* parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (bpDelta(keys, n, i, i + 1) - bpDelta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = bpDelta(keys, n, i, i - d);
  int lmax = 2;
  while (bpDelta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  #pragma unroll 1
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (bpDelta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = bpDelta(keys, n, i, j);
  int s = 0;
  #pragma unroll 1
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (bpDelta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int first = min(i, j), last = max(i, j);
  const int left = (first == gamma) ? ~gamma : gamma;
  const int right = (last == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
  node_child[i] = make_int2(left, right);
  node_range[i] = make_int2(first, last);
  parent[left >= 0 ? left : (n - 1 + ~left)] = i;
  parent[right >= 0 ? right : (n - 1 + ~right)] = i;
  if (i == 0) parent[0] = -1;
}

template <typename S>
FCLB_DI Box6<S> loadBoxCG(const Box6<S>* p) {  // L2 read: the box was written by another thread of this launch
  Box6<S> b;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    b.mn[k] = __ldcg(&p->mn[k]);
    b.mx[k] = __ldcg(&p->mx[k]);
  }
  return b;
}

// one thread per leaf walks up; the second arrival at a node unions its children
template <typename S>
__global__ void bpRefitKernel(const Box6<S>* __restrict__ leaf_box, int n, const int2* __restrict__ node_child,
                              const int* __restrict__ parent, Box6<S>* node_box, int* flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = parent[n - 1 + i];
  while (p >= 0) {
    __threadfence();
    if (atomicAdd(&flags[p], 1) == 0) return;
    const int2 c = node_child[p];
    const Box6<S> a = c.x >= 0 ? loadBoxCG(node_box + c.x) : leaf_box[~c.x];
    const Box6<S> b = c.y >= 0 ? loadBoxCG(node_box + c.y) : leaf_box[~c.y];
    Box6<S> u;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      u.mn[k] = a.mn[k] < b.mn[k] ? a.mn[k] : b.mn[k];
      u.mx[k] = a.mx[k] > b.mx[k] ? a.mx[k] : b.mx[k];
    }
    node_box[p] = u;
    p = parent[p];
  }
}

template <typename S>
__global__ void bpPackKernel(const Box6<S>* __restrict__ leaf_box, const Box6<S>* __restrict__ node_box,
                             const int2* __restrict__ node_child, const int2* __restrict__ node_range, int n_inner,
                             FatNode<S>* fat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inner) return;
  const int2 c = node_child[i];
  const Box6<S> l = c.x >= 0 ? node_box[c.x] : leaf_box[~c.x];
  const Box6<S> r = c.y >= 0 ? node_box[c.y] : leaf_box[~c.y];
  FatNode<S> f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    f.lb[k] = l.mn[k];
    f.lb[3 + k] = l.mx[k];
    f.rb[k] = r.mn[k];
    f.rb[3 + k] = r.mx[k];
  }
  f.lc = c.x;
  f.rc = c.y;
  f.ll = c.x >= 0 ? node_range[c.x].y : ~c.x;
  f.rl = c.y >= 0 ? node_range[c.y].y : ~c.y;
  fat[i] = f;
}

// AABB<S>::overlap (math/bv/AABB-inl.h:82-88)
template <typename S>
FCLB_DI bool boxOverlap(const Box6<S>& a, const Box6<S>& b) {
  if (a.mn[0] > b.mx[0] || a.mn[1] > b.mx[1] || a.mn[2] > b.mx[2]) return false;
  if (a.mx[0] < b.mn[0] || a.mx[1] < b.mn[1] || a.mx[2] < b.mn[2]) return false;
  return true;
}

struct BpQueryArgs {
  const void* leaf_box;
  const uint64_t* leaf_id;
  const void* node_box;
  const int2* node_child;
  const int2* node_range;
  const void* fat;          // FatNode<S>[n-1]
  int n;                    // leaves of the tree
  const void* query_box;    // Box6<S>[n_query]
  const uint64_t* query_id;
  int n_query;
  int self;                 // 1: queries are the tree's own leaves, report j > i only
  int tree_first;           // 1: pair = (tree id, query id); 0: (query id, tree id)
  uint64_t* out_pairs;      // 2 ids per pair
  unsigned long long cap;
  unsigned long long* count;
  unsigned long long* visits;  // node boxes tested (optional)
};

template <typename S>
__global__ void __launch_bounds__(128) bpQueryKernel(BpQueryArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long visits = 0;
  if (i < a.n_query) {
    const Box6<S>* __restrict__ leaf_box = static_cast<const Box6<S>*>(a.leaf_box);
    const Box6<S>* __restrict__ node_box = static_cast<const Box6<S>*>(a.node_box);
    const Box6<S> qb = static_cast<const Box6<S>*>(a.query_box)[i];
    const uint64_t qid = a.query_id[i];
    auto emit = [&](int leaf) {
      const unsigned long long slot = atomicAdd(a.count, 1ull);
      if (slot < a.cap) {
        const uint64_t tid = a.leaf_id[leaf];
        a.out_pairs[2 * slot] = (a.self || !a.tree_first) ? qid : tid;
        a.out_pairs[2 * slot + 1] = (a.self || !a.tree_first) ? tid : qid;
      }
    };
    if (a.n == 1) {
      if (!a.self) {
        visits++;
        if (boxOverlap(qb, leaf_box[0])) emit(0);
      }
    } else {
      const FatNode<S>* __restrict__ fat = static_cast<const FatNode<S>*>(a.fat);
      int stack[64];
      int sp = 0;
      stack[sp++] = 0;
      while (sp > 0) {
        const FatNode<S> nd = fat[stack[--sp]];
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const int ch = side == 0 ? nd.lc : nd.rc;
          const int last = side == 0 ? nd.ll : nd.rl;
          if (a.self && last <= i) continue;  // self pairs once: only leaves behind the query's Morton position
          const S* cb = side == 0 ? nd.lb : nd.rb;
          visits++;
          const bool hit = !(qb.mn[0] > cb[3] || qb.mn[1] > cb[4] || qb.mn[2] > cb[5] || qb.mx[0] < cb[0] ||
                             qb.mx[1] < cb[1] || qb.mx[2] < cb[2]);
          if (!hit) continue;
          if (ch < 0)
            emit(~ch);
          else if (sp < 64)
            stack[sp++] = ch;
        }
      }
    }
  }
  if (a.visits) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) visits += __shfl_xor_sync(0xffffffffu, visits, off);
    if ((threadIdx.x & 31) == 0 && visits) atomicAdd(a.visits, visits);
  }
}

template <typename S>
__global__ void bpUpdateLeavesKernel(Box6<S>* leaf_box, const int* __restrict__ pos, const Box6<S>* __restrict__ nb, int m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  // UpdateObjectAABB grows the leaf: node.bv += new_AABB unless it already contains it (:345-351)
  Box6<S> o = leaf_box[pos[i]];
  const Box6<S> b = nb[i];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    o.mn[k] = b.mn[k] < o.mn[k] ? b.mn[k] : o.mn[k];
    o.mx[k] = b.mx[k] > o.mx[k] ? b.mx[k] : o.mx[k];
  }
  leaf_box[pos[i]] = o;
}

// CollisionObject<S>::computeAABB (collision_object-inl.h:141-154)
template <typename S>
__global__ void computeAabbKernel(const LocalAabbD<S>* __restrict__ local, const uint32_t* __restrict__ shape_ids,
                                  const S* __restrict__ poses, size_t n, Box6<S>* out) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const LocalAabbD<S> l = local[shape_ids[i]];
  const Pose<S> tf = loadPose(poses, i);
  const S prec = sizeof(S) == 4 ? S(1e-5f) : S(1e-12);  // NumTraits<S>::dummy_precision()
  bool ident = true;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const S v = tf.R(r, c);
      if (r == c ? (fabs_(v - S(1)) > prec) : (fabs_(v) > prec)) ident = false;
    }
  Box6<S> b;
  if (ident) {
    b.mn[0] = l.mn[0] + tf.t.x; b.mn[1] = l.mn[1] + tf.t.y; b.mn[2] = l.mn[2] + tf.t.z;
    b.mx[0] = l.mx[0] + tf.t.x; b.mx[1] = l.mx[1] + tf.t.y; b.mx[2] = l.mx[2] + tf.t.z;
  } else {
    const V3<S> c = apply(tf, mk<S>(l.center[0], l.center[1], l.center[2]));
    b.mn[0] = c.x - l.radius; b.mn[1] = c.y - l.radius; b.mn[2] = c.z - l.radius;
    b.mx[0] = c.x + l.radius; b.mx[1] = c.y + l.radius; b.mx[2] = c.z + l.radius;
  }
  out[i] = b;
}

// pairs of object ids -> the per-query arrays fclb_collide_batch_dev consumes
template <typename S>
__global__ void gatherPairsKernel(const uint64_t* __restrict__ id_pairs, size_t n, const uint32_t* __restrict__ shape_ids,
                                  const S* __restrict__ poses, fclb_pair* pairs, S* poses1, S* poses2) {
  const size_t q = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (q >= n) return;
  const uint64_t a = id_pairs[2 * q], b = id_pairs[2 * q + 1];
  fclb_pair p;
  p.shape1 = shape_ids[a];
  p.shape2 = shape_ids[b];
  pairs[q] = p;
#pragma unroll
  for (int k = 0; k < 12; k++) {
    poses1[12 * q + k] = poses[12 * a + k];
    poses2[12 * q + k] = poses[12 * b + k];
  }
}

static void freeTree(BpTree* t) {
  cudaFree(t->leaf_box);
  cudaFree(t->leaf_id);
  cudaFree(t->node_box);
  cudaFree(t->node_child);
  cudaFree(t->node_range);
  cudaFree(t->parent);
  cudaFree(t->flags);
  cudaFree(t->fat);
  delete t;
}

struct BpCounters {
  unsigned long long* counters = nullptr;  // device: [0] pair count, [1] visits
  uint64_t last_visits = 0;
  void* d_in = nullptr;  // staging of fclb_scene_self_collide_host
  size_t d_in_cap = 0;
};
static PerDevice<BpCounters> g_bp_pd;
#define g_bp_counters (g_bp_pd.get().counters)
#define g_bp_last_visits (g_bp_pd.get().last_visits)

template <typename S>
static int refit(Engine& e, BpTree* t) {
  if (t->n < 2) return FCLB_OK;
  FCLB_CUDA(cudaMemsetAsync(t->flags, 0, size_t(t->n - 1) * sizeof(int), e.compute));
  bpRefitKernel<S><<<(t->n + 127) / 128, 128, 0, e.compute>>>(static_cast<const Box6<S>*>(t->leaf_box), t->n, t->node_child,
                                                              t->parent, static_cast<Box6<S>*>(t->node_box), t->flags);
  bpPackKernel<S><<<(t->n - 1 + 127) / 128, 128, 0, e.compute>>>(static_cast<const Box6<S>*>(t->leaf_box),
                                                                 static_cast<const Box6<S>*>(t->node_box), t->node_child,
                                                                 t->node_range, t->n - 1, static_cast<FatNode<S>*>(t->fat));
  e.launches += 2;
  FCLB_CUDA(cudaGetLastError());
  return FCLB_OK;
}

// grow-only device scratch of the tree builder (bounds, sort keys, cub temp storage)
struct BpScratch {
  double* bounds = nullptr;
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  void* tmp = nullptr;
  size_t cap_n = 0, cap_tmp = 0;
};
static PerDevice<BpScratch> g_bp_scratch_pd;
#define g_bp_scratch (g_bp_scratch_pd.get())

// boxes / ids: DEVICE pointers.  A tree whose arrays already hold `cap` >= n objects is rebuilt in place.
template <typename S>
static int buildTreeDev(Engine& e, const void* d_boxes, const uint64_t* d_ids, int n, BpTree* t, bool want_host_map) {
  const Box6<S>* box = static_cast<const Box6<S>*>(d_boxes);
  if (t->cap < n || t->scalar_type != (sizeof(S) == 4 ? FCLB_F32 : FCLB_F64)) {
    cudaFree(t->leaf_box); cudaFree(t->leaf_id); cudaFree(t->node_box); cudaFree(t->node_child);
    cudaFree(t->node_range); cudaFree(t->parent); cudaFree(t->flags); cudaFree(t->fat);
    t->fat = nullptr;
    t->leaf_box = t->node_box = nullptr;
    t->leaf_id = nullptr; t->node_child = t->node_range = nullptr; t->parent = t->flags = nullptr;
    t->cap = 0;
    FCLB_CUDA(cudaMalloc(&t->leaf_box, size_t(n) * sizeof(Box6<S>)));
    FCLB_CUDA(cudaMalloc(&t->leaf_id, size_t(n) * sizeof(uint64_t)));
    const int ni = n > 1 ? n - 1 : 1;
    FCLB_CUDA(cudaMalloc(&t->node_box, size_t(ni) * sizeof(Box6<S>)));
    FCLB_CUDA(cudaMalloc(&t->node_child, size_t(ni) * sizeof(int2)));
    FCLB_CUDA(cudaMalloc(&t->node_range, size_t(ni) * sizeof(int2)));
    FCLB_CUDA(cudaMalloc(&t->parent, size_t(2 * n) * sizeof(int)));
    FCLB_CUDA(cudaMalloc(&t->flags, size_t(ni) * sizeof(int)));
    FCLB_CUDA(cudaMalloc(&t->fat, size_t(ni) * sizeof(FatNode<S>)));
    t->cap = n;
  }
  t->n = n;
  t->scalar_type = sizeof(S) == 4 ? FCLB_F32 : FCLB_F64;
  BpScratch& sc = g_bp_scratch;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, sc.keys, sc.keys2, n, 0, 62, e.compute);
  if (!sc.bounds) FCLB_CUDA(cudaMalloc(&sc.bounds, 6 * sizeof(double)));
  if (sc.cap_n < size_t(n)) {
    cudaFree(sc.keys);
    cudaFree(sc.keys2);
    sc.keys = sc.keys2 = nullptr;
    sc.cap_n = 0;
    FCLB_CUDA(cudaMalloc(&sc.keys, size_t(n) * 8));
    FCLB_CUDA(cudaMalloc(&sc.keys2, size_t(n) * 8));
    sc.cap_n = size_t(n);
  }
  if (sc.cap_tmp < tmp_bytes || !sc.tmp) {
    cudaFree(sc.tmp);
    sc.tmp = nullptr;
    sc.cap_tmp = 0;
    FCLB_CUDA(cudaMalloc(&sc.tmp, tmp_bytes ? tmp_bytes : 8));
    sc.cap_tmp = tmp_bytes ? tmp_bytes : 8;
  }
  double* d_bounds = sc.bounds;
  unsigned long long *d_keys = sc.keys, *d_keys2 = sc.keys2;
  void* d_tmp = sc.tmp;
  const double init[6] = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300};
  FCLB_CUDA(cudaMemcpyAsync(d_bounds, init, sizeof(init), cudaMemcpyHostToDevice, e.compute));
  const int grid = (n + 127) / 128;
  bpBoundsKernel<S><<<grid < 1184 ? grid : 1184, 128, 0, e.compute>>>(box, n, d_bounds);
  bpMortonKernel<S><<<grid, 128, 0, e.compute>>>(box, n, d_bounds, d_keys);
  cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_keys2, n, 0, 62, e.compute);
  bpGatherKernel<S><<<grid, 128, 0, e.compute>>>(box, d_ids, d_keys2, n, static_cast<Box6<S>*>(t->leaf_box), t->leaf_id);
  e.launches += 4;
  if (n > 1) {
    bpHierarchyKernel<<<(n - 1 + 127) / 128, 128, 0, e.compute>>>(d_keys2, n, t->node_child, t->node_range, t->parent);
    e.launches += 1;
    int rc = refit<S>(e, t);
    if (rc) return rc;
  }
  FCLB_CUDA(cudaGetLastError());
  if (want_host_map) {
    std::vector<uint64_t> ids(n);
    FCLB_CUDA(cudaMemcpyAsync(ids.data(), t->leaf_id, size_t(n) * 8, cudaMemcpyDeviceToHost, e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    t->pos_of.reserve(size_t(n) * 2);
    #pragma unroll 1
    for (int i = 0; i < n; i++) t->pos_of[ids[i]] = i;
  }
  return FCLB_OK;  // stream-ordered: later work on e.compute sees the finished tree
}

template <typename S>
static int queryDev(Engine& e, const BpTree* t, const void* q_box, const uint64_t* q_id, int nq, int self, int tree_first,
                    uint64_t* d_out, size_t cap, size_t* n_pairs) {
  if (!g_bp_counters) FCLB_CUDA(cudaMalloc(&g_bp_counters, 2 * sizeof(unsigned long long)));
  FCLB_CUDA(cudaMemsetAsync(g_bp_counters, 0, 2 * sizeof(unsigned long long), e.compute));
  BpQueryArgs a{};
  a.leaf_box = t->leaf_box;
  a.leaf_id = t->leaf_id;
  a.node_box = t->node_box;
  a.node_child = t->node_child;
  a.node_range = t->node_range;
  a.fat = t->fat;
  a.n = t->n;
  a.query_box = q_box;
  a.query_id = q_id;
  a.n_query = nq;
  a.self = self;
  a.tree_first = tree_first;
  a.out_pairs = d_out;
  a.cap = d_out ? cap : 0;
  a.count = g_bp_counters;
  a.visits = g_bp_counters + 1;
  bpQueryKernel<S><<<(nq + 127) / 128, 128, 0, e.compute>>>(a);
  e.launches += 1;
  FCLB_CUDA(cudaGetLastError());
  unsigned long long h[2] = {0, 0};
  FCLB_CUDA(cudaMemcpyAsync(h, g_bp_counters, sizeof(h), cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  g_bp_last_visits = h[1];
  if (n_pairs) *n_pairs = size_t(h[0]);
  if (d_out && h[0] > cap) return fail(FCLB_ERR_CAPACITY, "broadphase: pair buffer too small (n_pairs holds the required size)");
  return FCLB_OK;
}

__global__ void iotaKernel(uint64_t* ids, size_t n) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) ids[i] = i;
}
__global__ void countNonZeroKernel(const uint32_t* __restrict__ v, size_t n, unsigned long long* out) {
  unsigned long long c = 0;
  #pragma unroll 1
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) c += v[i] != 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// workspace of fclb_scene_self_collide_*: grow-only, reused across scenes
struct SceneWs {
  BpTree tree;
  void* boxes = nullptr;
  uint64_t* ids = nullptr;
  size_t cap_obj = 0, obj_scalar = 0;
  uint64_t* id_pairs = nullptr;
  fclb_pair* pairs = nullptr;
  void *poses1 = nullptr, *poses2 = nullptr;
  uint32_t* counts = nullptr;
  size_t cap_pairs = 0, pair_scalar = 0;
  unsigned long long* hits = nullptr;
};
static PerDevice<SceneWs> g_scene_ws_pd;
#define g_scene (g_scene_ws_pd.get())

template <typename S>
static int sceneSelfCollide(Engine& e, fclb_handle shapes, const uint32_t* shape_ids, const void* poses, size_t n,
                            const fclb_request* req, size_t* n_candidates, size_t* n_colliding, uint64_t* out_id_pairs,
                            uint32_t* out_counts, size_t out_cap) {
  SceneWs& w = g_scene;
  const int st = sizeof(S) == 4 ? FCLB_F32 : FCLB_F64;
  if (w.cap_obj < n || w.obj_scalar < sizeof(S)) {
    cudaFree(w.boxes);
    cudaFree(w.ids);
    w.boxes = nullptr;
    w.ids = nullptr;
    w.cap_obj = 0;
    FCLB_CUDA(cudaMalloc(&w.boxes, n * 6 * 8));
    FCLB_CUDA(cudaMalloc(&w.ids, n * 8));
    w.cap_obj = n;
    w.obj_scalar = 8;
  }
  if (!w.hits) FCLB_CUDA(cudaMalloc(&w.hits, sizeof(unsigned long long)));
  int rc = fclb_compute_aabb_batch_dev(shapes, shape_ids, poses, n, st, w.boxes);
  if (rc) return rc;
  iotaKernel<<<int((n + 255) / 256), 256, 0, e.compute>>>(w.ids, n);
  e.launches += 1;
  rc = buildTreeDev<S>(e, w.boxes, w.ids, int(n), &w.tree, false);
  if (rc) return rc;
  size_t found = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    const size_t want = attempt == 0 ? (w.cap_pairs ? w.cap_pairs : 8 * n) : found;
    if (w.cap_pairs < want || w.pair_scalar < sizeof(S)) {
      cudaFree(w.id_pairs); cudaFree(w.pairs); cudaFree(w.poses1); cudaFree(w.poses2); cudaFree(w.counts);
      w.id_pairs = nullptr; w.pairs = nullptr; w.poses1 = w.poses2 = nullptr; w.counts = nullptr;
      w.cap_pairs = 0;
      const size_t cap = want + want / 8 + 1024;
      FCLB_CUDA(cudaMalloc(&w.id_pairs, cap * 16));
      FCLB_CUDA(cudaMalloc(&w.pairs, cap * sizeof(fclb_pair)));
      FCLB_CUDA(cudaMalloc(&w.poses1, cap * 12 * 8));
      FCLB_CUDA(cudaMalloc(&w.poses2, cap * 12 * 8));
      FCLB_CUDA(cudaMalloc(&w.counts, cap * 4));
      w.cap_pairs = cap;
      w.pair_scalar = 8;
    }
    rc = queryDev<S>(e, &w.tree, w.tree.leaf_box, w.tree.leaf_id, w.tree.n, 1, 0, w.id_pairs, w.cap_pairs, &found);
    if (rc == FCLB_OK) break;
    if (rc != FCLB_ERR_CAPACITY || attempt == 1) return rc;
  }
  if (n_candidates) *n_candidates = found;
  unsigned long long hits = 0;
  if (found) {
    rc = fclb_gather_pairs_dev(w.id_pairs, found, shape_ids, poses, st, w.pairs, w.poses1, w.poses2);
    if (rc) return rc;
    rc = fclb_collide_batch_dev(shapes, w.pairs, w.poses1, w.poses2, found, st, req, 0, nullptr, w.counts);
    if (rc) return rc;
    FCLB_CUDA(cudaMemsetAsync(w.hits, 0, sizeof(unsigned long long), e.compute));
    countNonZeroKernel<<<296, 256, 0, e.compute>>>(w.counts, found, w.hits);
    e.launches += 1;
    FCLB_CUDA(cudaMemcpyAsync(&hits, w.hits, sizeof(hits), cudaMemcpyDeviceToHost, e.compute));
    if (out_id_pairs && out_counts) {
      const size_t m = found < out_cap ? found : out_cap;
      FCLB_CUDA(cudaMemcpyAsync(out_id_pairs, w.id_pairs, m * 16, cudaMemcpyDefault, e.compute));
      FCLB_CUDA(cudaMemcpyAsync(out_counts, w.counts, m * 4, cudaMemcpyDefault, e.compute));
    }
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
  }
  if (n_colliding) *n_colliding = size_t(hits);
  return FCLB_OK;
}

static BpTree* findTree(fclb_handle h) {
  auto it = bpTable().find(h);
  return it == bpTable().end() ? nullptr : it->second;
}

}  // namespace fclb

using namespace fclb;

extern "C" {

int fclb_broadphase_build_dev(const void* aabbs, const uint64_t* user_ids, size_t n, int scalar_type, fclb_handle* tree) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!aabbs || !user_ids || !tree || n == 0 || n > 0x7fffffffull) return fail(FCLB_ERR_BAD_ARG, "fclb_broadphase_build: bad argument");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  BpTree* t = new BpTree();
  rc = scalar_type == FCLB_F32 ? buildTreeDev<float>(e, aabbs, user_ids, int(n), t, false)
                               : buildTreeDev<double>(e, aabbs, user_ids, int(n), t, false);
  if (!rc && cudaStreamSynchronize(e.compute) != cudaSuccess) rc = fail(FCLB_ERR_CUDA, "broadphase build failed");
  if (rc) {
    freeTree(t);
    return rc;
  }
  const fclb_handle h = newHandle();
  bpTable()[h] = t;
  *tree = h;
  return FCLB_OK;
}

int fclb_broadphase_build_host(const void* aabbs, const uint64_t* user_ids, size_t n, int scalar_type, fclb_handle* tree) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!aabbs || !user_ids || !tree || n == 0 || n > 0x7fffffffull) return fail(FCLB_ERR_BAD_ARG, "fclb_broadphase_build: bad argument");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  void* d_box = nullptr;
  uint64_t* d_id = nullptr;
  FCLB_CUDA(cudaMalloc(&d_box, n * 6 * ss));
  FCLB_CUDA(cudaMalloc(&d_id, n * 8));
  FCLB_CUDA(cudaMemcpyAsync(d_box, aabbs, n * 6 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(d_id, user_ids, n * 8, cudaMemcpyHostToDevice, e.compute));
  BpTree* t = new BpTree();
  rc = scalar_type == FCLB_F32 ? buildTreeDev<float>(e, d_box, d_id, int(n), t, true)
                               : buildTreeDev<double>(e, d_box, d_id, int(n), t, true);
  cudaFree(d_box);
  cudaFree(d_id);
  if (rc) {
    freeTree(t);
    return rc;
  }
  const fclb_handle h = newHandle();
  bpTable()[h] = t;
  *tree = h;
  return FCLB_OK;
}

int fclb_broadphase_release(fclb_handle tree) {
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = bpTable().find(tree);
  if (it == bpTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_broadphase_release: unknown handle");
  freeTree(it->second);
  bpTable().erase(it);
  return FCLB_OK;
}

/* out_pairs: DEVICE buffer of 2*cap ids (may be NULL to count only) */
int fclb_broadphase_self_pairs_dev(fclb_handle tree, uint64_t* out_pairs, size_t cap, size_t* n_pairs) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  BpTree* t = findTree(tree);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown broadphase tree handle");
  if (t->scalar_type == FCLB_F32)
    return queryDev<float>(e, t, t->leaf_box, t->leaf_id, t->n, 1, 0, out_pairs, cap, n_pairs);
  return queryDev<double>(e, t, t->leaf_box, t->leaf_id, t->n, 1, 0, out_pairs, cap, n_pairs);
}

static int pairsToHost(Engine& e, int rc_query, uint64_t* d_buf, uint64_t* out_pairs, size_t cap, size_t n_found) {
  if (rc_query == FCLB_OK && out_pairs && n_found) {
    const size_t m = n_found < cap ? n_found : cap;
    cudaError_t ce = cudaMemcpy(out_pairs, d_buf, m * 16, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) rc_query = fail(FCLB_ERR_CUDA, cudaGetErrorString(ce));
  }
  if (d_buf) cudaFree(d_buf);
  (void)e;
  return rc_query;
}

int fclb_broadphase_self_pairs_host(fclb_handle tree, uint64_t* out_pairs, size_t cap, size_t* n_pairs) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  uint64_t* d_buf = nullptr;
  if (out_pairs && cap) FCLB_CUDA(cudaMalloc(&d_buf, cap * 16));
  size_t found = 0;
  rc = fclb_broadphase_self_pairs_dev(tree, d_buf, cap, &found);
  if (n_pairs) *n_pairs = found;
  return pairsToHost(e, rc, d_buf, out_pairs, cap, found);
}

int fclb_broadphase_tree_pairs_host(fclb_handle tree_a, fclb_handle tree_b, uint64_t* out_pairs, size_t cap,
                                    size_t* n_pairs) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  BpTree* a = findTree(tree_a);
  BpTree* b = findTree(tree_b);
  if (!a || !b) return fail(FCLB_ERR_BAD_ARG, "unknown broadphase tree handle");
  if (a->scalar_type != b->scalar_type) return fail(FCLB_ERR_BAD_ARG, "trees of different scalar types");
  uint64_t* d_buf = nullptr;
  if (out_pairs && cap) FCLB_CUDA(cudaMalloc(&d_buf, cap * 16));
  size_t found = 0;
  // leaves of A query tree B; pair = (id in A, id in B) as TreeCollision reports (node1 of this, node2 of tree2)
  rc = a->scalar_type == FCLB_F32 ? queryDev<float>(e, b, a->leaf_box, a->leaf_id, a->n, 0, 0, d_buf, cap, &found)
                                  : queryDev<double>(e, b, a->leaf_box, a->leaf_id, a->n, 0, 0, d_buf, cap, &found);
  if (n_pairs) *n_pairs = found;
  return pairsToHost(e, rc, d_buf, out_pairs, cap, found);
}

int fclb_broadphase_query_pairs_host(fclb_handle tree, const void* aabbs, const uint64_t* object_ids, size_t n,
                                     uint64_t* out_pairs, size_t cap, size_t* n_pairs) {
  int rc = ensureInit();
  if (rc) return rc;
  if (n == 0) {
    if (n_pairs) *n_pairs = 0;
    return FCLB_OK;
  }
  if (!aabbs || !object_ids || n > 0x7fffffffull) return fail(FCLB_ERR_BAD_ARG, "fclb_broadphase_query_pairs: bad argument");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  BpTree* t = findTree(tree);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown broadphase tree handle");
  const size_t ss = t->scalar_type == FCLB_F32 ? 4 : 8;
  void* d_box = nullptr;
  uint64_t* d_id = nullptr;
  uint64_t* d_buf = nullptr;
  FCLB_CUDA(cudaMalloc(&d_box, n * 6 * ss));
  FCLB_CUDA(cudaMalloc(&d_id, n * 8));
  if (out_pairs && cap) FCLB_CUDA(cudaMalloc(&d_buf, cap * 16));
  FCLB_CUDA(cudaMemcpyAsync(d_box, aabbs, n * 6 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(d_id, object_ids, n * 8, cudaMemcpyHostToDevice, e.compute));
  size_t found = 0;
  // SingleObjectCollision reports (leaf id, object id)
  rc = t->scalar_type == FCLB_F32 ? queryDev<float>(e, t, d_box, d_id, int(n), 0, 1, d_buf, cap, &found)
                                  : queryDev<double>(e, t, d_box, d_id, int(n), 0, 1, d_buf, cap, &found);
  cudaFree(d_box);
  cudaFree(d_id);
  if (n_pairs) *n_pairs = found;
  return pairsToHost(e, rc, d_buf, out_pairs, cap, found);
}

int fclb_broadphase_update_host(fclb_handle tree, const uint64_t* user_ids, const void* new_aabbs, size_t n) {
  int rc = ensureInit();
  if (rc) return rc;
  if (n == 0) return FCLB_OK;
  if (!user_ids || !new_aabbs) return fail(FCLB_ERR_BAD_ARG, "fclb_broadphase_update: null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  BpTree* t = findTree(tree);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown broadphase tree handle");
  if (t->pos_of.empty()) return fail(FCLB_ERR_UNSUPPORTED, "tree was built from device arrays: no id map on the host");
  const size_t ss = t->scalar_type == FCLB_F32 ? 4 : 8;
  // an id may appear more than once: the growth is a union, so merge its boxes on the host first
  std::vector<int> pos;
  std::vector<unsigned char> merged;
  std::unordered_map<int, size_t> slot_of;
  #pragma unroll 1
  for (size_t i = 0; i < n; i++) {
    auto it = t->pos_of.find(user_ids[i]);
    if (it == t->pos_of.end()) return fail(FCLB_ERR_BAD_ARG, "fclb_broadphase_update: unknown user id");  // UpdateObjectAABB returns false
    const unsigned char* src = static_cast<const unsigned char*>(new_aabbs) + i * 6 * ss;
    auto sl = slot_of.find(it->second);
    if (sl == slot_of.end()) {
      slot_of[it->second] = pos.size();
      pos.push_back(it->second);
      merged.insert(merged.end(), src, src + 6 * ss);
    } else if (ss == 4) {
      float* d = reinterpret_cast<float*>(merged.data() + sl->second * 24);
      const float* b = reinterpret_cast<const float*>(src);
      for (int k = 0; k < 3; k++) {
        d[k] = b[k] < d[k] ? b[k] : d[k];
        d[3 + k] = b[3 + k] > d[3 + k] ? b[3 + k] : d[3 + k];
      }
    } else {
      double* d = reinterpret_cast<double*>(merged.data() + sl->second * 48);
      const double* b = reinterpret_cast<const double*>(src);
      for (int k = 0; k < 3; k++) {
        d[k] = b[k] < d[k] ? b[k] : d[k];
        d[3 + k] = b[3 + k] > d[3 + k] ? b[3 + k] : d[3 + k];
      }
    }
  }
  n = pos.size();
  new_aabbs = merged.data();
  int* d_pos = nullptr;
  void* d_box = nullptr;
  FCLB_CUDA(cudaMalloc(&d_pos, n * 4));
  FCLB_CUDA(cudaMalloc(&d_box, n * 6 * ss));
  FCLB_CUDA(cudaMemcpyAsync(d_pos, pos.data(), n * 4, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(d_box, new_aabbs, n * 6 * ss, cudaMemcpyHostToDevice, e.compute));
  if (t->scalar_type == FCLB_F32) {
    bpUpdateLeavesKernel<float><<<int((n + 127) / 128), 128, 0, e.compute>>>(static_cast<Box6<float>*>(t->leaf_box), d_pos,
                                                                           static_cast<const Box6<float>*>(d_box), int(n));
    rc = refit<float>(e, t);
  } else {
    bpUpdateLeavesKernel<double><<<int((n + 127) / 128), 128, 0, e.compute>>>(static_cast<Box6<double>*>(t->leaf_box), d_pos,
                                                                            static_cast<const Box6<double>*>(d_box), int(n));
    rc = refit<double>(e, t);
  }
  e.launches += 1;
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  cudaFree(d_pos);
  cudaFree(d_box);
  return rc;
}

uint64_t fclb_broadphase_last_visits(void) { return g_bp_last_visits; }

/* One scene end to end on the device: computeAABB for every object, tree build, SelfCollision, and boolean
 * fcl::collide on every candidate pair.  shape_ids / poses: DEVICE arrays (object i = user id i).
 * out_id_pairs / out_counts (optional, host or device, out_cap pairs): the candidate (id, id) pairs and
 * numContacts of each. */
int fclb_scene_self_collide_dev(fclb_handle shapes, const uint32_t* shape_ids, const void* poses, size_t n, int scalar_type,
                                const fclb_request* req, size_t* n_candidates, size_t* n_colliding, uint64_t* out_id_pairs,
                                uint32_t* out_counts, size_t out_cap) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (!req) return fail(FCLB_ERR_BAD_ARG, "null request");
  if (n_candidates) *n_candidates = 0;
  if (n_colliding) *n_colliding = 0;
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses || n > 0x7fffffffull) return fail(FCLB_ERR_BAD_ARG, "fclb_scene_self_collide: bad argument");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  if (scalar_type == FCLB_F32)
    return sceneSelfCollide<float>(e, shapes, shape_ids, poses, n, req, n_candidates, n_colliding, out_id_pairs, out_counts,
                                   out_cap);
  return sceneSelfCollide<double>(e, shapes, shape_ids, poses, n, req, n_candidates, n_colliding, out_id_pairs, out_counts,
                                  out_cap);
}

int fclb_scene_self_collide_host(fclb_handle shapes, const uint32_t* shape_ids, const void* poses, size_t n, int scalar_type,
                                 const fclb_request* req, size_t* n_candidates, size_t* n_colliding, uint64_t* out_id_pairs,
                                 uint32_t* out_counts, size_t out_cap) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return fclb_scene_self_collide_dev(shapes, shape_ids, poses, 0, scalar_type, req, n_candidates, n_colliding,
                                                 out_id_pairs, out_counts, out_cap);
  if (!shape_ids || !poses) return fail(FCLB_ERR_BAD_ARG, "null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  void*& d_in = g_bp_pd.get().d_in;
  size_t& d_in_cap = g_bp_pd.get().d_in_cap;
  const size_t o_p = alignUp(n * 4, 256), total = o_p + n * 12 * ss;
  if (d_in_cap < total) {
    cudaFree(d_in);
    d_in = nullptr;
    d_in_cap = 0;
    FCLB_CUDA(cudaMalloc(&d_in, total));
    d_in_cap = total;
  }
  char* base = static_cast<char*>(d_in);
  FCLB_CUDA(cudaMemcpyAsync(base, shape_ids, n * 4, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p, poses, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  return fclb_scene_self_collide_dev(shapes, reinterpret_cast<const uint32_t*>(base), base + o_p, n, scalar_type, req,
                                     n_candidates, n_colliding, out_id_pairs, out_counts, out_cap);
}

/* CollisionObject<S>::computeAABB for n objects: out = 6 S per object (min xyz, max xyz). DEVICE pointers. */
int fclb_compute_aabb_batch_dev(fclb_handle shapes, const uint32_t* shape_ids, const void* poses, size_t n, int scalar_type,
                                void* out_aabbs) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  ShapeTable* t = findTable(e, shapes);
  if (!t) return fail(FCLB_ERR_BAD_ARG, "unknown shape table handle");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses || !out_aabbs) return fail(FCLB_ERR_BAD_ARG, "null array");
  const int grid = int((n + 127) / 128);
  if (scalar_type == FCLB_F32)
    computeAabbKernel<float><<<grid, 128, 0, e.compute>>>(static_cast<const LocalAabbD<float>*>(t->d_local[0]), shape_ids,
                                                          static_cast<const float*>(poses), n,
                                                          static_cast<Box6<float>*>(out_aabbs));
  else
    computeAabbKernel<double><<<grid, 128, 0, e.compute>>>(static_cast<const LocalAabbD<double>*>(t->d_local[1]), shape_ids,
                                                           static_cast<const double*>(poses), n,
                                                           static_cast<Box6<double>*>(out_aabbs));
  e.launches += 1;
  FCLB_CUDA(cudaGetLastError());
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}

static int compute_aabb_batch_host_one(fclb_handle shapes, const uint32_t* shape_ids, const void* poses, size_t n, int scalar_type,
                                 void* out_aabbs) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!shape_ids || !poses || !out_aabbs) return fail(FCLB_ERR_BAD_ARG, "null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_ids = 0, o_p = alignUp(n * 4, 256), o_out = alignUp(o_p + n * 12 * ss, 256);
  rc = ensureStage(e, alignUp(o_out + n * 6 * ss, 256));
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_ids, shape_ids, n * 4, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p, poses, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  rc = fclb_compute_aabb_batch_dev(shapes, reinterpret_cast<const uint32_t*>(base + o_ids), base + o_p, n, scalar_type,
                                   base + o_out);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpy(out_aabbs, base + o_out, n * 6 * ss, cudaMemcpyDeviceToHost));
  return FCLB_OK;
}
int fclb_compute_aabb_batch_host(fclb_handle shapes, const uint32_t* shape_ids, const void* poses, size_t n, int scalar_type,
                                 void* out_aabbs) {
  if (engineCount() <= 1) return compute_aabb_batch_host_one(shapes, shape_ids, poses, n, scalar_type, out_aabbs);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return compute_aabb_batch_host_one(shapes, offT(shape_ids, b), offPtr(poses, b * 12 * ss), m_, scalar_type, offPtr(out_aabbs, b * 6 * ss)); });
}

/* candidate (object id, object id) pairs -> the per-query arrays of fclb_collide_batch_dev. DEVICE pointers. */
int fclb_gather_pairs_dev(const uint64_t* id_pairs, size_t n_pairs, const uint32_t* shape_ids, const void* poses,
                          int scalar_type, fclb_pair* out_pairs, void* out_poses1, void* out_poses2) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n_pairs == 0) return FCLB_OK;
  if (!id_pairs || !shape_ids || !poses || !out_pairs || !out_poses1 || !out_poses2) return fail(FCLB_ERR_BAD_ARG, "null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const int grid = int((n_pairs + 127) / 128);
  if (scalar_type == FCLB_F32)
    gatherPairsKernel<float><<<grid, 128, 0, e.compute>>>(id_pairs, n_pairs, shape_ids, static_cast<const float*>(poses),
                                                          out_pairs, static_cast<float*>(out_poses1),
                                                          static_cast<float*>(out_poses2));
  else
    gatherPairsKernel<double><<<grid, 128, 0, e.compute>>>(id_pairs, n_pairs, shape_ids, static_cast<const double*>(poses),
                                                           out_pairs, static_cast<double*>(out_poses1),
                                                           static_cast<double*>(out_poses2));
  e.launches += 1;
  FCLB_CUDA(cudaGetLastError());
  return FCLB_OK;
}

}  // extern "C"
