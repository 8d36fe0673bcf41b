// fclb_bvh.cu -- batched mesh-mesh collide over flattened BVHModel<OBBRSS>
// trees: ONE WARP PER QUERY (relative pose).
//
// Reference path (results contract):
//   fcl::collide(BVH, BVH) -> BVHCollide<OBBRSS> -> OrientedNodeBVHSolver::MeshIntersect
//     narrowphase/detail/traversal/collision/bvh_solver-inl.h:75-160
//   overlap(R, T, OBBRSS, OBBRSS) -> OBB only (math/bv/OBBRSS-inl.h:130-135)
//     -> overlap(R0,T0,b1,b2) + obbDisjoint (math/bv/OBB-inl.h:305-436)
//   leaf: SimplexIntersect -> trianglePairIntersect -> Intersect::intersect_Triangle
//     (shape_pair_intersect-inl.h:201-270; traversal/collision/intersect-inl.h:594-612,724-845)
//   relativeTransform (math/geometry-inl.h:409-435)
//
// The reference walks ONE (node,node) pair at a time from a std::stack (DFS).
// The boolean result and the number of intersecting triangle pairs do not
// depend on the visiting order, so here a warp owns the query and
//   * keeps the pair stack in shared memory and pops up to 32 pairs per step,
//     one per lane (adaptive: near capacity it degrades to 1-wide DFS);
//   * every lane fetches its two 64-byte nodes with four 128-bit loads each
//     (128-byte nodes / eight loads in double) and runs the 15-axis OBB test;
//   * surviving pairs are expanded and pushed back with ballot + popc prefix
//     sums; leaf pairs go to a separate shared-memory queue and the 17-axis
//     triangle test runs on full batches of 32 (the "leaf stage");
//   * a hit is reduced with a warp ballot: boolean queries stop at once,
//     counting queries add popc(ballot) until max_contacts is reached.
// Which contact is reported first for max_contacts == 1 is therefore NOT the
// reference's DFS-first pair (SURVEY.md 7 "Traversal-order-dependent outputs");
// out_first_pair holds *a* colliding triangle pair.
#include <cstdio>
#include <cstring>
#include <vector>

#include "fclb_bvh_build.h"
#include "fclb_bvh.cuh"
#include "fclb_mpr_pen.cuh"

namespace fclb {

std::map<fclb_handle, BvhDev*>& bvhTable() {
  static std::map<fclb_handle, BvhDev*> t[kMaxDevices];
  return t[currentSlot()];
}

// project6 (intersect-inl.h:1082-1105); std::min(a,b) = (b<a)?b:a, std::max(a,b) = (a<b)?b:a
template <typename S>
FCLB_DI bool project6(const V3<S>& ax, const V3<S>& p1, const V3<S>& p2, const V3<S>& p3, const V3<S>& q1,
                      const V3<S>& q2, const V3<S>& q3) {
  const S P1 = dot(ax, p1), P2 = dot(ax, p2), P3 = dot(ax, p3);
  const S Q1 = dot(ax, q1), Q2 = dot(ax, q2), Q3 = dot(ax, q3);
  const S mn1 = fmin_(P1, fmin_(P2, P3));
  const S mx2 = fmax_(Q1, fmax_(Q2, Q3));
  if (mn1 > mx2) return false;
  const S mx1 = fmax_(P1, fmax_(P2, P3));
  const S mn2 = fmin_(Q1, fmin_(Q2, Q3));
  if (mn2 > mx1) return false;
  return true;
}

// Intersect::intersect_Triangle, boolean (intersect-inl.h:594-612 -> :724-794)
template <typename S>
FCLB_DI bool triTriIntersect(const V3<S> P[3], const V3<S> Q[3], const M3<S>& R, const V3<S>& T) {
  const V3<S> Q1 = mulMV(R, Q[0]) + T, Q2 = mulMV(R, Q[1]) + T, Q3 = mulMV(R, Q[2]) + T;
  const V3<S> p1 = P[0] - P[0], p2 = P[1] - P[0], p3 = P[2] - P[0];
  const V3<S> q1 = Q1 - P[0], q2 = Q2 - P[0], q3 = Q3 - P[0];
  const V3<S> e1 = p2 - p1, e2 = p3 - p2;
  const V3<S> n1 = cross(e1, e2);
  if (!project6(n1, p1, p2, p3, q1, q2, q3)) return false;
  const V3<S> f1 = q2 - q1, f2 = q3 - q2;
  const V3<S> m1 = cross(f1, f2);
  if (!project6(m1, p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e1, f1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e1, f2), p1, p2, p3, q1, q2, q3)) return false;
  const V3<S> f3 = q1 - q3;
  if (!project6(cross(e1, f3), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e2, f1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e2, f2), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e2, f3), p1, p2, p3, q1, q2, q3)) return false;
  const V3<S> e3 = p1 - p3;
  if (!project6(cross(e3, f1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e3, f2), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e3, f3), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e1, n1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e2, n1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(e3, n1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(f1, m1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(f2, m1), p1, p2, p3, q1, q2, q3)) return false;
  if (!project6(cross(f3, m1), p1, p2, p3, q1, q2, q3)) return false;
  return true;
}

// Contact part of Intersect::intersect_Triangle (intersect-inl.h:795-845) for an intersecting pair:
// buildTrianglePlane (:1036-1050), computeDeepestPoints (:849-888).  P, Qw in mesh 1's frame.
// Returns the number of contact points (0..2); points / depth / normal in mesh 1's frame.
template <typename S>
FCLB_DI void deepestPoints(const V3<S> pts[3], const V3<S>& n, S t, S& depth, V3<S> deepest[3], unsigned& num) {
  const S eps = S(1e-5);  // Intersect<S>::getEpsilon()
  S max_depth = -(sizeof(S) == 4 ? S(3.402823466e+38f) : S(1.7976931348623157e+308));
  unsigned nd = 0, num_neg = 0, num_pos = 0, num_zero = 0;
  for (int i = 0; i < 3; i++) {
    const S dist = -(dot(n, pts[i]) - t);
    if (dist > eps)
      num_pos++;
    else if (dist < -eps)
      num_neg++;
    else
      num_zero++;
    if (dist > max_depth) {
      max_depth = dist;
      nd = 1;
      deepest[0] = pts[i];
    } else if (double(dist) + 1e-6 >= double(max_depth)) {  // "dist + 1e-6": a double literal
      nd++;
      deepest[nd - 1] = pts[i];
    }
  }
  if (max_depth < -eps) nd = 0;
  if (num_zero == 0 && ((num_neg == 0) || (num_pos == 0))) nd = 0;
  depth = max_depth;
  num = nd;
}
template <typename S>
FCLB_DI unsigned triTriContacts(const V3<S> P[3], const V3<S> Q[3], const M3<S>& R, const V3<S>& T, V3<S> pts[2], S& depth,
                                V3<S>& normal) {
  const V3<S> Qw[3] = {mulMV(R, Q[0]) + T, mulMV(R, Q[1]) + T, mulMV(R, Q[2]) + T};
  V3<S> n1 = normalized(cross(P[1] - P[0], P[2] - P[0]));
  const S t1 = dot(n1, P[0]);
  V3<S> n2 = normalized(cross(Qw[1] - Qw[0], Qw[2] - Qw[0]));
  const S t2 = dot(n2, Qw[0]);
  V3<S> deep1[3], deep2[3];
  unsigned num1 = 0, num2 = 0;
  S pd1, pd2;
  deepestPoints(Qw, n1, t1, pd2, deep2, num2);
  deepestPoints(P, n2, t2, pd1, deep1, num1);
  unsigned nc;
  if (pd1 > pd2) {
    nc = num2 < 2u ? num2 : 2u;
    #pragma unroll 1
    for (unsigned i = 0; i < nc; i++) pts[i] = deep2[i];
    normal = n1;
    depth = pd2;
  } else {
    nc = num1 < 2u ? num1 : 2u;
    #pragma unroll 1
    for (unsigned i = 0; i < nc; i++) pts[i] = deep1[i];
    normal = -n2;
    depth = pd1;
  }
  return nc;
}

constexpr int kBvhWarps = 8;        // warps per CTA
constexpr int kStackCap = 1024;     // (node,node) pairs per warp
constexpr int kLeafCap = 64;        // queued leaf pairs per warp

struct BvhArgs {
  const void* nodes1;
  const void* nodes2;
  const void* tris1;
  const void* tris2;
  const void* poses1;
  const void* poses2;
  size_t n;
  uint32_t max_contacts;
  uint32_t* counts;
  int32_t* first_pair;
  unsigned long long* work_counter;
  unsigned long long* stats;  // [0] BV-pair tests, [1] leaf-pair tests (optional)
  // contact generation (request.useDefaultPenetration()): first max_keep contacts of each query
  uint32_t max_keep;
  int32_t* out_ids;      // [n * max_keep * 2] = b1, b2
  void* out_contacts;    // [n * max_keep * 7 S] = normal, pos, depth (world frame)
};

// pop width of a query that wants only a few contacts; measured on C3 / C4 (B200): 32 -> 4.76 / 23.4 ms,
// 16 -> 5.04 / 25.1 ms, 8 -> 5.85 / 28.7 ms (most of the work is proving the non-colliding queries separate)
#ifndef FCLB_EAGER_WIDTH
#define FCLB_EAGER_WIDTH 32
#endif
template <typename S, bool PEN>
__global__ void __launch_bounds__(kBvhWarps * 32) bvhCollideKernel(BvhArgs a) {
  extern __shared__ __align__(16) int2 s_bvh[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int2* stack = s_bvh + size_t(warp) * (kStackCap + kLeafCap);
  int2* leafq = stack + kStackCap;
  const S* __restrict__ nodes1 = static_cast<const S*>(a.nodes1);
  const S* __restrict__ nodes2 = static_cast<const S*>(a.nodes2);
  const S* __restrict__ tris1 = static_cast<const S*>(a.tris1);
  const S* __restrict__ tris2 = static_cast<const S*>(a.tris2);
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned long long st_bv = 0, st_leaf = 0;

  while (true) {
    unsigned long long q64 = 0;
    if (lane == 0) q64 = atomicAdd(a.work_counter, 1ull);
    q64 = __shfl_sync(0xffffffffu, q64, 0);
    if (q64 >= a.n) break;
    const size_t q = size_t(q64);
    // relativeTransform (geometry-inl.h:433-434)
    M3<S> R;
    V3<S> t;
    Pose<S> tf2w;  // SimplexIntersect hands tf2 to trianglePairIntersect as the contact frame (shape_pair_intersect-inl.h:266)
    {
      const Pose<S> tf1 = loadPose(static_cast<const S*>(a.poses1), q);
      const Pose<S> tf2 = loadPose(static_cast<const S*>(a.poses2), q);
      R = mulMtM(tf1.R, tf2.R);
      t = mulMtV(tf1.R, tf2.t - tf1.t);
      tf2w = tf2;
    }
    uint32_t count = 0;
    int first_a = -1, first_b = -1;
    int sp = 1, nleaf = 0;
    if (lane == 0) stack[0] = make_int2(0, 0);
    __syncwarp();
    bool done = (a.max_contacts == 0);

    // A query that needs only a few contacts (boolean collide: one) should not wait for a full batch of leaf pairs:
    // `eager` runs the leaf stage as soon as it has work (C3: 5.25 -> 4.76 ms).
    const bool eager = a.max_contacts <= 8;
    const int width = eager ? FCLB_EAGER_WIDTH : 32;
    while (!done && (sp > 0 || nleaf > 0)) {
      if (sp > 0 && nleaf < 32) {
        // ---- BV stage: pop up to `width` pairs ----
        int take = sp < width ? sp : width;
        if (sp + take > kStackCap - 256) take = 1;  // near capacity: plain DFS (grows by <= 1 per step)
        int2 pr = make_int2(-1, -1);
        if (lane < take) pr = stack[sp - 1 - lane];
        sp -= take;
        __syncwarp();
        bool expand = false, leaf = false;
        int2 c0 = make_int2(0, 0), c1 = make_int2(0, 0);
        if (lane < take) {
          const NodeD<S> n1 = loadNode(nodes1, pr.x);
          const NodeD<S> n2 = loadNode(nodes2, pr.y);
          st_bv++;
          if (obbOverlap(R, t, n1, n2)) {
            const bool l1 = n1.first_child < 0, l2 = n2.first_child < 0;
            if (l1 && l2) {
              leaf = true;
              c0 = make_int2(-(n1.first_child + 1), -(n2.first_child + 1));
            } else {
              expand = true;
              // descend the first tree if the second is a leaf, or if both are
              // inner and bv_1.size() > bv_2.size() (bvh_solver-inl.h:148-160)
              const bool on1 = l2 || (!l1 && sqnorm(n1.extent) > sqnorm(n2.extent));
              if (on1) {
                c0 = make_int2(n1.first_child, pr.y);
                c1 = make_int2(n1.first_child + 1, pr.y);
              } else {
                c0 = make_int2(pr.x, n2.first_child);
                c1 = make_int2(pr.x, n2.first_child + 1);
              }
            }
          }
        }
        const unsigned em = __ballot_sync(0xffffffffu, expand);
        const unsigned lm = __ballot_sync(0xffffffffu, leaf);
        if (sp + 2 * __popc(em) > kStackCap) {  // deeper than the depth-first head room: report, never corrupt
          if (lane == 0) atomicAdd(&a.stats[2], 1ull);
          done = true;
          expand = false;
        }
        if (expand) {
          const int pos = sp + 2 * __popc(em & lt_mask);
          stack[pos] = c0;
          stack[pos + 1] = c1;
        }
        if (leaf) leafq[nleaf + __popc(lm & lt_mask)] = c0;
        if (!done) sp += 2 * __popc(em);
        nleaf += __popc(lm);
        __syncwarp();
      }
      if (nleaf >= 32 || ((sp == 0 || eager) && nleaf > 0)) {
        // ---- leaf stage: one triangle pair per lane ----
        const int batch = nleaf < 32 ? nleaf : 32;
        bool hit = false;
        int2 lp = make_int2(-1, -1);
        int pen_nc = 0;
        S pen_depth = S(0);
        V3<S> pen_normal = zero3<S>(), pen_p0 = zero3<S>(), pen_p1 = zero3<S>();
        if (lane < batch) {
          lp = leafq[nleaf - 1 - lane];
          V3<S> P[3], Q[3];
          loadTri(tris1, lp.x, P);
          loadTri(tris2, lp.y, Q);
          st_leaf++;
          hit = triTriIntersect(P, Q, R, t);
          if (PEN && hit) {
            // trianglePairIntersect with penetration (shape_pair_intersect-inl.h:201-252): <= 2 contacts per pair
            V3<S> pts[2];
            S depth;
            V3<S> normal;
            const unsigned nc = triTriContacts(P, Q, R, t, pts, depth, normal);
            pen_nc = int(nc);
            pen_depth = depth;
            pen_normal = mulMV(tf2w.R, normal);
            pen_p0 = nc > 0 ? apply(tf2w, pts[0]) : zero3<S>();
            pen_p1 = nc > 1 ? apply(tf2w, pts[1]) : zero3<S>();
          }
        }
        nleaf -= batch;
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm) {
          if (first_a < 0) {
            const int src = __ffs(hm) - 1;
            first_a = __shfl_sync(0xffffffffu, lp.x, src);
            first_b = __shfl_sync(0xffffffffu, lp.y, src);
          }
          if (PEN) {
            int off = hit ? pen_nc : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int v = __shfl_up_sync(0xffffffffu, off, o);
              if (lane >= o) off += v;
            }
            const int total = __shfl_sync(0xffffffffu, off, 31);
            off -= hit ? pen_nc : 0;
            if (hit && a.out_ids) {
              #pragma unroll 1
              for (int k = 0; k < pen_nc; k++) {
                const uint32_t slot = count + uint32_t(off + k);
                if (slot < a.max_keep && slot < a.max_contacts) {
                  const size_t r = q * a.max_keep + slot;
                  a.out_ids[2 * r] = lp.x;
                  a.out_ids[2 * r + 1] = lp.y;
                  S* o7 = static_cast<S*>(a.out_contacts) + r * 7;
                  const V3<S> pp = k == 0 ? pen_p0 : pen_p1;
                  o7[0] = pen_normal.x; o7[1] = pen_normal.y; o7[2] = pen_normal.z;
                  o7[3] = pp.x; o7[4] = pp.y; o7[5] = pp.z;
                  o7[6] = pen_depth;
                }
              }
            }
            count += uint32_t(total);
          } else {
            if (hit && a.out_ids) {  // boolean request with more than one contact wanted: the ids of every kept pair
              const uint32_t slot = count + uint32_t(__popc(hm & ((1u << lane) - 1u)));
              if (slot < a.max_keep && slot < a.max_contacts) {
                const size_t r = q * a.max_keep + slot;
                a.out_ids[2 * r] = lp.x;
                a.out_ids[2 * r + 1] = lp.y;
              }
            }
            count += uint32_t(__popc(hm));
          }
          if (count >= a.max_contacts) {
            count = a.max_contacts;
            done = true;
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0) {
      a.counts[q] = count;
      if (a.first_pair) {
        a.first_pair[2 * q] = first_a;
        a.first_pair[2 * q + 1] = first_b;
      }
    }
    __syncwarp();
  }
  if (a.stats) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      st_bv += __shfl_xor_sync(0xffffffffu, st_bv, off);
      st_leaf += __shfl_xor_sync(0xffffffffu, st_leaf, off);
    }
    if (lane == 0) {
      atomicAdd(&a.stats[0], st_bv);
      atomicAdd(&a.stats[1], st_leaf);
    }
  }
}

struct BvhStats {
  unsigned long long* counters = nullptr;  // device: [0] work counter, [1..2] stats
  unsigned long long last[2] = {0, 0};
  unsigned long long host[3] = {0, 0, 0};
};
static PerDevice<BvhStats> g_bvh_pd;
#define g_bvh_counters (g_bvh_pd.get().counters)
#define g_last_stats (g_bvh_pd.get().last)

template <typename S>
static int bvhCollideDev(Engine& e, const BvhDev* m1, const BvhDev* m2, const void* poses1, const void* poses2, size_t n,
                         uint32_t max_contacts, uint32_t* counts, int32_t* first_pair, bool pen = false, uint32_t max_keep = 0,
                         int32_t* out_ids = nullptr, void* out_contacts = nullptr) {
  if (!g_bvh_counters) FCLB_CUDA(cudaMalloc(&g_bvh_counters, 4 * sizeof(unsigned long long)));
  FCLB_CUDA(cudaMemsetAsync(g_bvh_counters, 0, 4 * sizeof(unsigned long long), e.compute));
  BvhArgs a{};
  a.nodes1 = m1->nodes;
  a.nodes2 = m2->nodes;
  a.tris1 = m1->tris;
  a.tris2 = m2->tris;
  a.poses1 = poses1;
  a.poses2 = poses2;
  a.n = n;
  a.max_contacts = max_contacts;
  a.counts = counts;
  a.first_pair = first_pair;
  a.work_counter = g_bvh_counters;
  a.stats = g_bvh_counters + 1;
  a.max_keep = max_keep;
  a.out_ids = out_ids;
  a.out_contacts = out_contacts;
  const size_t need = (n + kBvhWarps - 1) / kBvhWarps;
  const size_t cap = size_t(e.sms) * 3;
  const int grid = int(need < cap ? need : cap);
  FCLB_CUDA(cudaEventRecord(e.ev_call0, e.compute));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  const size_t smem = size_t(kBvhWarps) * (kStackCap + kLeafCap) * sizeof(int2);
  if (pen) {
    FCLB_CUDA(cudaFuncSetAttribute(bvhCollideKernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    bvhCollideKernel<S, true><<<grid, kBvhWarps * 32, smem, e.compute>>>(a);
  } else {
    FCLB_CUDA(cudaFuncSetAttribute(bvhCollideKernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    bvhCollideKernel<S, false><<<grid, kBvhWarps * 32, smem, e.compute>>>(a);
  }
  FCLB_CUDA(cudaGetLastError());
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  e.launches += 1;
  unsigned long long* h_stats = g_bvh_pd.get().host;
  FCLB_CUDA(cudaMemcpyAsync(h_stats, g_bvh_counters + 1, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  g_last_stats[0] = h_stats[0];
  g_last_stats[1] = h_stats[1];
  if (h_stats[2]) return fail(FCLB_ERR_CAPACITY, "mesh-mesh traversal: BVH deeper than the per-warp pair stack allows");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = ms;
  e.last_call_ms = ms;
  e.n_rec = 1;
  e.rec_kind[0] = -1;
  e.rec_count[0] = n;
  e.rec_ms[0] = ms;
  return FCLB_OK;
}

// ---- refit: BVHModel::refitTree(bottomup = false) (BVH_model-inl.h:624-637) ---------------------------------------
// Every node is fitted again from the primitives it covers with the tree's topology unchanged: the OBBRSS fitter of
// build time (detail/BV_fitter-inl.h:324-345: covariance of the triangle corners -> eigen_old -> axisFromEigen ->
// extent and centre along the axes).  One warp per node.  The nine covariance sums are sequential in the reference
// (math/geometry-inl.h:713-780) and their rounding depends on that order, so lane k < 9 accumulates sum k over the
// node's primitives in primitive_indices_ order; the 3x3 Jacobi solve runs on every lane; the extents are minima /
// maxima (order-free) and are reduced across the lanes.  Results are bit-identical to the reference's refit.
// The nine covariance sums of a node (S1[x, y, z], c00, c11, c22, c01, c02, c12: lane k < 9 owns sum k) accumulated in
// primitive order.  The triangles of 32 primitives are fetched by the whole warp into a shared-memory stage (one gather per
// lane instead of one dependent load pair per element of the sequential loop), then every owning lane adds its 32 terms in
// order: same terms, same order, same rounding as the reference's sequential loop.
template <typename S>
FCLB_DI S orderedCovarianceSum(const S* __restrict__ tris, const int* __restrict__ prim, int first, int count, int lane, S* stage) {
  const int a = lane < 3 ? lane : (lane == 3 ? 0 : lane == 4 ? 1 : lane == 5 ? 2 : lane == 6 ? 0 : lane == 7 ? 0 : 1);
  const int b = lane < 3 ? lane : (lane == 3 ? 0 : lane == 4 ? 1 : lane == 5 ? 2 : lane == 6 ? 1 : lane == 7 ? 2 : 2);
  S acc = S(0);
  #pragma unroll 1
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    if (i < count) {
      const S* t = tris + size_t(12) * size_t(prim[first + i]);
#pragma unroll
      for (int v = 0; v < 3; v++)
#pragma unroll
        for (int k = 0; k < 3; k++) stage[lane * 9 + 3 * v + k] = t[4 * v + k];
    }
    __syncwarp();
    const int m = count - base < 32 ? count - base : 32;
    if (lane < 3) {
      #pragma unroll 1
      for (int j = 0; j < m; j++) {
        const S* p = stage + 9 * j;
        acc += (p[a] + p[3 + a]) + p[6 + a];
      }
    } else if (lane < 9) {
      #pragma unroll 1
      for (int j = 0; j < m; j++) {
        const S* p = stage + 9 * j;
        acc += (p[a] * p[b] + p[3 + a] * p[3 + b] + p[6 + a] * p[6 + b]);
      }
    }
    __syncwarp();
  }
  return acc;
}
// lanes 0..2: sum over the primitives, in order, of the centroid component ((p1 + p2) + p3) / 3 (computeRule_mean)
template <typename S>
FCLB_DI S orderedCentroidSum(const S* __restrict__ tris, const int* __restrict__ prim, int first, int count, int lane, S* stage) {
  S acc = S(0);
  #pragma unroll 1
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    if (i < count) {
      const S* t = tris + size_t(12) * size_t(prim[first + i]);
#pragma unroll
      for (int v = 0; v < 3; v++)
#pragma unroll
        for (int k = 0; k < 3; k++) stage[lane * 9 + 3 * v + k] = t[4 * v + k];
    }
    __syncwarp();
    const int m = count - base < 32 ? count - base : 32;
    if (lane < 3)
      #pragma unroll 1
      for (int j = 0; j < m; j++) {
        const S* p = stage + 9 * j;
        acc += ((p[lane] + p[3 + lane]) + p[6 + lane]) / 3;
      }
    __syncwarp();
  }
  return acc;
}

template <typename S>
__global__ void __launch_bounds__(256) bvhRefitKernel(S* __restrict__ nodes, const S* __restrict__ tris, const int2* __restrict__ range,
                                                      const int* __restrict__ prim, int n_nodes) {
  __shared__ S s_stage[8][32 * 9];
  S* stage = s_stage[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  #pragma unroll 1
  for (int node = warp; node < n_nodes; node += n_warps) {
    const int2 r = range[node];
    // which scalars of a triangle this lane's sum needs: S1[k] (lanes 0-2), c00 c11 c22 c01 c02 c12 (lanes 3-8)
    const S acc = orderedCovarianceSum<S>(tris, prim, r.x, r.y, lane, stage);
    S sums[9];
#pragma unroll
    for (int k = 0; k < 9; k++) sums[k] = __shfl_sync(0xffffffffu, acc, k);
    const int n_points = 3 * r.y;
    S M[3][3];
    M[0][0] = sums[3] - sums[0] * sums[0] / n_points;
    M[1][1] = sums[4] - sums[1] * sums[1] / n_points;
    M[2][2] = sums[5] - sums[2] * sums[2] / n_points;
    M[0][1] = sums[6] - sums[0] * sums[1] / n_points;
    M[1][2] = sums[8] - sums[1] * sums[2] / n_points;
    M[0][2] = sums[7] - sums[0] * sums[2] / n_points;
    M[1][0] = M[0][1];
    M[2][0] = M[0][2];
    M[2][1] = M[1][2];
    S d[3] = {0, 0, 0}, vec[3][3];
    if (!hostbuild::jacobi3<S>(M, d, vec)) {
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) vec[i][j] = (i == j) ? S(1) : S(0);
    }
    S ax[9];
    hostbuild::axesFromEigen<S>(vec, d, ax);
    const S big = sizeof(S) == 4 ? S(3.402823466e+38f) : S(1.7976931348623157e+308);
    S mn[3] = {big, big, big}, mx[3] = {-big, -big, -big};
    #pragma unroll 1
    for (int i = lane; i < 3 * r.y; i += 32) {  // one triangle corner per lane and round
      const S* p = tris + size_t(12) * size_t(prim[r.x + i / 3]) + 4 * (i % 3);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const S proj = (ax[0 + k] * p[0] + ax[3 + k] * p[1]) + ax[6 + k] * p[2];
        if (proj > mx[k]) mx[k] = proj;
        if (proj < mn[k]) mn[k] = proj;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const S omx = __shfl_xor_sync(0xffffffffu, mx[k], off), omn = __shfl_xor_sync(0xffffffffu, mn[k], off);
        if (omx > mx[k]) mx[k] = omx;
        if (omn < mn[k]) mn[k] = omn;
      }
    if (lane == 0) {
      S* o = nodes + size_t(16) * size_t(node);
#pragma unroll
      for (int k = 0; k < 9; k++) o[k] = ax[k];
      S c[3];
#pragma unroll
      for (int k = 0; k < 3; k++) c[k] = (mx[k] + mn[k]) / 2;
#pragma unroll
      for (int rr = 0; rr < 3; rr++) o[9 + rr] = (ax[3 * rr] * c[0] + ax[3 * rr + 1] * c[1]) + ax[3 * rr + 2] * c[2];
#pragma unroll
      for (int k = 0; k < 3; k++) o[12 + k] = (mx[k] - mn[k]) / 2;
    }
  }
}

// collisionPenetrationMPR for mesh pairs (collision_penetration-inl.h:189-252): every contact of the boolean collide is a
// triangle pair (b1, b2); computePenetrationMPR on (triangle b1 at tf1, triangle b2 at tf2) fills normal / pos / depth
template <typename S>
__global__ void __launch_bounds__(kBlock) bvhPairPenetrationKernel(const S* __restrict__ tris1, const S* __restrict__ tris2,
                                                                   const S* __restrict__ poses1, const S* __restrict__ poses2, size_t n,
                                                                   uint32_t max_keep, const uint32_t* __restrict__ counts,
                                                                   const int32_t* __restrict__ ids, int incremental, double dx,
                                                                   double dy, double dz, double tol, S* __restrict__ out) {
  const size_t total = n * size_t(max_keep);
  const V3<S> dir_world = mk<S>(S(dx), S(dy), S(dz));
  #pragma unroll 1
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const size_t q = i / max_keep;
    const uint32_t k = uint32_t(i % max_keep);
    S* o = out + i * 7;
    if (k >= counts[q]) {
#pragma unroll
      for (int j = 0; j < 7; j++) o[j] = S(0);
      continue;
    }
    const Pose<S> tf1 = loadPose(poses1, q), tf2 = loadPose(poses2, q);
    MinkDiff<S, ST_TRIANGLE, ST_TRIANGLE> md;
    md.s0.type = md.s1.type = ST_TRIANGLE;
    md.s0.cvx = md.s1.cvx = nullptr;
    md.s0.p0 = md.s0.p1 = md.s0.p2 = md.s1.p0 = md.s1.p1 = md.s1.p2 = S(0);
    loadTri(tris1, ids[2 * i], md.s0.tri);
    loadTri(tris2, ids[2 * i + 1], md.s1.tri);
    md.setPoses(tf1, tf2);
    V3<S> pos, normal;
    S depth;
    computePenetrationMpr<S>(md, tf1, dir_world, incremental != 0, 128, S(tol), pos, normal, depth);
    o[0] = normal.x; o[1] = normal.y; o[2] = normal.z;
    o[3] = pos.x; o[4] = pos.y; o[5] = pos.z;
    o[6] = depth;
  }
}

// ---- bottom-up refit (BVHModel::refitTreeBottomUp, BVH_model-inl.h:580-617) --------------------------------------
// A leaf box is fit(3 points) -> OBB_fit_functions::fit3 (math/bv/utility-inl.h:85-109); an inner box is
// left.bv + right.bv -> OBB<S>::operator+ (math/bv/OBB-inl.h:116-126): merge_largedist when the centres are further
// apart than twice the summed largest half sides (first axis = the centre difference, the other two from the covariance
// of the 16 corners projected off it, :197-244), else merge_smalldist (rotation = the normalised sum of the two
// quaternions, box around the 16 corners, :248-293).  Only the OBB half of OBBRSS is kept (collide reads nothing else).
template <typename S>
struct ObbB {
  S axis[9];  // row-major axis(i, j)
  S To[3], ext[3];
};
template <typename S>
FCLB_DI void obbCorners(const ObbB<S>& b, S v[8][3]) {  // computeVertices (OBB-inl.h:177-193)
  S e0[3], e1[3], e2[3];
  for (int k = 0; k < 3; k++) {
    e0[k] = b.axis[3 * k + 0] * b.ext[0];
    e1[k] = b.axis[3 * k + 1] * b.ext[1];
    e2[k] = b.axis[3 * k + 2] * b.ext[2];
  }
  const int s0[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, s1[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, s2[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
  for (int i = 0; i < 8; i++)
    for (int k = 0; k < 3; k++) {
      S t = s0[i] > 0 ? b.To[k] + e0[k] : b.To[k] - e0[k];
      t = s1[i] > 0 ? t + e1[k] : t - e1[k];
      v[i][k] = s2[i] > 0 ? t + e2[k] : t - e2[k];
    }
}
template <typename S>
FCLB_DI void quatFromAxis(const S* m, S q[4]) {  // Quaternion(Matrix3): Shepperd's method; q = (x, y, z, w)
  S t = m[0] + m[4] + m[8];
  if (t > S(0)) {
    t = fsqrt(t + S(1.0));
    q[3] = S(0.5) * t;
    t = S(0.5) / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = fsqrt(m[4 * i] - m[4 * j] - m[4 * k] + S(1.0));
    q[i] = S(0.5) * t;
    t = S(0.5) / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}
template <typename S>
FCLB_DI void obbMerge(const ObbB<S>& b1, const ObbB<S>& b2, ObbB<S>& b) {
  const S big = sizeof(S) == 4 ? S(3.402823466e+38f) : S(1.7976931348623157e+308);
  S cd[3];
  for (int k = 0; k < 3; k++) cd[k] = b1.To[k] - b2.To[k];
  const S m1 = fmax_(fmax_(b1.ext[0], b1.ext[1]), b1.ext[2]), m2 = fmax_(fmax_(b2.ext[0], b2.ext[1]), b2.ext[2]);
  const S cd2 = (cd[0] * cd[0] + cd[1] * cd[1]) + cd[2] * cd[2];
  S v[16][3];
  obbCorners(b1, v);
  obbCorners(b2, v + 8);
  if (fsqrt(cd2) > 2 * (m1 + m2)) {  // merge_largedist
    S a0[3] = {cd[0], cd[1], cd[2]};
    if (cd2 > S(0)) {
      const S n = fsqrt(cd2);
      for (int k = 0; k < 3; k++) a0[k] = a0[k] / n;
    }
    S S1[3] = {0, 0, 0}, c00 = 0, c11 = 0, c22 = 0, c01 = 0, c02 = 0, c12 = 0;
    for (int i = 0; i < 16; i++) {
      const S d = (v[i][0] * a0[0] + v[i][1] * a0[1]) + v[i][2] * a0[2];
      S p[3];
      for (int k = 0; k < 3; k++) p[k] = v[i][k] - a0[k] * d;
      for (int k = 0; k < 3; k++) S1[k] += p[k];
      c00 += p[0] * p[0];
      c11 += p[1] * p[1];
      c22 += p[2] * p[2];
      c01 += p[0] * p[1];
      c02 += p[0] * p[2];
      c12 += p[1] * p[2];
    }
    const int n_points = 16;
    S M[3][3];
    M[0][0] = c00 - S1[0] * S1[0] / n_points;
    M[1][1] = c11 - S1[1] * S1[1] / n_points;
    M[2][2] = c22 - S1[2] * S1[2] / n_points;
    M[0][1] = c01 - S1[0] * S1[1] / n_points;
    M[1][2] = c12 - S1[1] * S1[2] / n_points;
    M[0][2] = c02 - S1[0] * S1[2] / n_points;
    M[1][0] = M[0][1];
    M[2][0] = M[0][2];
    M[2][1] = M[1][2];
    S d[3] = {0, 0, 0}, vec[3][3];
    if (!hostbuild::jacobi3<S>(M, d, vec)) {
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) vec[i][j] = (i == j) ? S(1) : S(0);
    }
    int mn, md, mx;
    if (d[0] > d[1]) {
      mx = 0;
      mn = 1;
    } else {
      mn = 0;
      mx = 1;
    }
    if (d[2] < d[mn]) {
      md = mn;
      mn = 2;
    } else if (d[2] > d[mx]) {
      md = mx;
      mx = 2;
    } else {
      md = 2;
    }
    (void)mn;
    for (int r = 0; r < 3; r++) {
      b.axis[3 * r + 0] = a0[r];
      b.axis[3 * r + 1] = vec[r][mx];
      b.axis[3 * r + 2] = vec[r][md];
    }
    S lo[3] = {big, big, big}, hi[3] = {-big, -big, -big};
    for (int i = 0; i < 16; i++)
      for (int j = 0; j < 3; j++) {
        const S proj = (b.axis[j] * v[i][0] + b.axis[3 + j] * v[i][1]) + b.axis[6 + j] * v[i][2];
        if (proj > hi[j]) hi[j] = proj;
        if (proj < lo[j]) lo[j] = proj;
      }
    S o[3];
    for (int k = 0; k < 3; k++) o[k] = (hi[k] + lo[k]) / 2;
    for (int r = 0; r < 3; r++) b.To[r] = (b.axis[3 * r] * o[0] + b.axis[3 * r + 1] * o[1]) + b.axis[3 * r + 2] * o[2];
    for (int k = 0; k < 3; k++) b.ext[k] = (hi[k] - lo[k]) * S(0.5);
    return;
  }
  // merge_smalldist
  for (int k = 0; k < 3; k++) b.To[k] = (b1.To[k] + b2.To[k]) * S(0.5);
  S q0[4], q1[4], q[4];
  quatFromAxis(b1.axis, q0);
  quatFromAxis(b2.axis, q1);
  const S dt = ((q0[0] * q1[0] + q0[1] * q1[1]) + q0[2] * q1[2]) + q0[3] * q1[3];
  if (dt < 0)
    for (int k = 0; k < 4; k++) q1[k] = -q1[k];
  for (int k = 0; k < 4; k++) q[k] = q0[k] + q1[k];
  {
    const S z = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
    if (z > S(0)) {
      const S n = fsqrt(z);
      for (int k = 0; k < 4; k++) q[k] = q[k] / n;
    }
  }
  {  // toRotationMatrix
    const S tx = S(2) * q[0], ty = S(2) * q[1], tz = S(2) * q[2];
    const S twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const S txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const S tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    b.axis[0] = S(1) - (tyy + tzz);
    b.axis[1] = txy - twz;
    b.axis[2] = txz + twy;
    b.axis[3] = txy + twz;
    b.axis[4] = S(1) - (txx + tzz);
    b.axis[5] = tyz - twx;
    b.axis[6] = txz - twy;
    b.axis[7] = tyz + twx;
    b.axis[8] = S(1) - (txx + tyy);
  }
  S pmin[3] = {big, big, big}, pmax[3] = {-big, -big, -big};
  for (int i = 0; i < 16; i++) {
    S diff[3];
    for (int k = 0; k < 3; k++) diff[k] = v[i][k] - b.To[k];
    for (int j = 0; j < 3; j++) {
      const S dot = (diff[0] * b.axis[j] + diff[1] * b.axis[3 + j]) + diff[2] * b.axis[6 + j];
      if (dot > pmax[j])
        pmax[j] = dot;
      else if (dot < pmin[j])
        pmin[j] = dot;
    }
  }
  for (int j = 0; j < 3; j++) {
    const S h = S(0.5) * (pmax[j] + pmin[j]);
    for (int k = 0; k < 3; k++) b.To[k] += b.axis[3 * k + j] * h;
    b.ext[j] = S(0.5) * (pmax[j] - pmin[j]);
  }
}
template <typename S>
FCLB_DI void loadObbB(const S* nodes, int i, ObbB<S>& b) {
  const S* p = nodes + size_t(16) * i;
  for (int k = 0; k < 9; k++) b.axis[k] = p[k];
  for (int k = 0; k < 3; k++) {
    b.To[k] = p[9 + k];
    b.ext[k] = p[12 + k];
  }
}
template <typename S>
FCLB_DI void storeObbB(S* nodes, int i, const ObbB<S>& b) {
  S* p = nodes + size_t(16) * i;
  for (int k = 0; k < 9; k++) p[k] = b.axis[k];
  for (int k = 0; k < 3; k++) {
    p[9 + k] = b.To[k];
    p[12 + k] = b.ext[k];
  }
}
// one thread per leaf: fits its box, then climbs; the second thread to arrive at a node merges the two children
template <typename S>
__global__ void __launch_bounds__(128) bvhRefitBottomUpKernel(S* nodes, const S* __restrict__ tris, const int* __restrict__ leaf_node,
                                                              const int* __restrict__ parent, int* __restrict__ arrived, int n_tris) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tris) return;
  int node = leaf_node[t];
  {
    const S* p = tris + size_t(12) * t;
    // OBB_fit_functions::fit3 + getExtentAndCenter_pointcloud
    const S e0[3] = {p[0] - p[4], p[1] - p[5], p[2] - p[6]}, e1[3] = {p[4] - p[8], p[5] - p[9], p[6] - p[10]},
            e2[3] = {p[8] - p[0], p[9] - p[1], p[10] - p[2]};
    const S l0 = (e0[0] * e0[0] + e0[1] * e0[1]) + e0[2] * e0[2], l1 = (e1[0] * e1[0] + e1[1] * e1[1]) + e1[2] * e1[2],
            l2 = (e2[0] * e2[0] + e2[1] * e2[1]) + e2[2] * e2[2];
    int imax = 0;
    if (l1 > l0) imax = 1;
    if (l2 > (imax == 0 ? l0 : l1)) imax = 2;
    S c2[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
    {
      const S z = (c2[0] * c2[0] + c2[1] * c2[1]) + c2[2] * c2[2];
      if (z > S(0)) {
        const S n = fsqrt(z);
        for (int k = 0; k < 3; k++) c2[k] = c2[k] / n;
      }
    }
    S c0[3];
    for (int k = 0; k < 3; k++) c0[k] = imax == 0 ? e0[k] : (imax == 1 ? e1[k] : e2[k]);
    {
      const S z = (c0[0] * c0[0] + c0[1] * c0[1]) + c0[2] * c0[2];
      if (z > S(0)) {
        const S n = fsqrt(z);
        for (int k = 0; k < 3; k++) c0[k] = c0[k] / n;
      }
    }
    const S c1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
    ObbB<S> b;
    for (int r = 0; r < 3; r++) {
      b.axis[3 * r + 0] = c0[r];
      b.axis[3 * r + 1] = c1[r];
      b.axis[3 * r + 2] = c2[r];
    }
    const S big = sizeof(S) == 4 ? S(3.402823466e+38f) : S(1.7976931348623157e+308);
    S lo[3] = {big, big, big}, hi[3] = {-big, -big, -big};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        const S proj = (b.axis[j] * p[4 * i] + b.axis[3 + j] * p[4 * i + 1]) + b.axis[6 + j] * p[4 * i + 2];
        if (proj > hi[j]) hi[j] = proj;
        if (proj < lo[j]) lo[j] = proj;
      }
    S o[3];
    for (int k = 0; k < 3; k++) o[k] = (hi[k] + lo[k]) / 2;
    for (int r = 0; r < 3; r++) b.To[r] = (b.axis[3 * r] * o[0] + b.axis[3 * r + 1] * o[1]) + b.axis[3 * r + 2] * o[2];
    for (int k = 0; k < 3; k++) b.ext[k] = (hi[k] - lo[k]) * S(0.5);
    storeObbB(nodes, node, b);
  }
  while (node != 0) {
    node = parent[node];
    __threadfence();
    if (atomicAdd(&arrived[node], 1) == 0) return;  // the sibling subtree is not done yet: its last thread continues
    __threadfence();
    S* np = nodes + size_t(16) * node;
    int fc;
    if (sizeof(S) == 4)
      fc = __float_as_int(float(np[15]));
    else
      fc = int(__double_as_longlong(double(np[15])));
    ObbB<S> a, c, m;
    loadObbB(nodes, fc, a);
    loadObbB(nodes, fc + 1, c);
    obbMerge(a, c, m);
    storeObbB(nodes, node, m);
  }
}

// 9 S per triangle (upload layout) -> the 12 S device records
template <typename S>
__global__ void triRepackKernel(const S* __restrict__ in9, int n_tris, S* __restrict__ out12) {
  #pragma unroll 1
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < size_t(n_tris) * 3; i += size_t(gridDim.x) * blockDim.x) {
    const S* p = in9 + 3 * i;
    S* o = out12 + 4 * i;
    o[0] = p[0];
    o[1] = p[1];
    o[2] = p[2];
    o[3] = S(0);
  }
}

// first_primitive / num_primitives and primitive_indices_ from the child links: an explicit-stack DFS, left child first
static void deriveRanges(const int32_t* first_child, int n_nodes, int n_tris, std::vector<int2>& range, std::vector<int>& prim) {
  range.assign(size_t(n_nodes), make_int2(0, 0));
  prim.clear();
  prim.reserve(size_t(n_tris));
  std::vector<std::pair<int, int>> stack;  // (node, state: 0 = enter, 1 = leave)
  stack.push_back({0, 0});
  while (!stack.empty()) {
    const auto top = stack.back();
    stack.pop_back();
    const int node = top.first;
    if (top.second == 1) {
      range[size_t(node)].y = int(prim.size()) - range[size_t(node)].x;
      continue;
    }
    range[size_t(node)].x = int(prim.size());
    const int fc = first_child[node];
    if (fc < 0) {
      prim.push_back(-(fc + 1));
      range[size_t(node)].y = 1;
      continue;
    }
    stack.push_back({node, 1});
    stack.push_back({fc + 1, 0});
    stack.push_back({fc, 0});
  }
}

// ---- BVHModel<OBBRSS>::buildTree on the device (BVH_model-inl.h:402-570, detail/BV_fitter-inl.h:324-345,
// detail/BV_splitter-inl.h:361-372), level by level ------------------------------------------------------------
// The reference builds depth-first with an explicit stack: fit the node's box to its primitives, split them at the mean
// of their centroids along the box's first axis by an in-place swap pass over primitive_indices_, allocate the two
// children at the end of the node array and continue with the left one.  Everything observable is reproduced: the
// fit sums run in primitive order (one lane per sum, as in the refit), the swap pass is replayed element by element on
// the node's slice (the flags of 32 elements are evaluated in parallel, the swaps applied in order), and the node ids
// follow from the depth-first order: the children of the k-th inner node in preorder are 2k+1 and 2k+2, and a node's
// preorder rank among inner nodes is its parent's + 1 (left child) or + the left sibling's primitive count (right child).
struct BuildNode {
  int id, first, count, rank;
};

template <typename S>
__global__ void __launch_bounds__(256) bvhBuildFitKernel(const BuildNode* __restrict__ level, int n_level, S* __restrict__ nodes,
                                                         const S* __restrict__ tris, const int* __restrict__ prim,
                                                         int2* __restrict__ range) {
  __shared__ S s_stage[8][32 * 9];
  S* stage = s_stage[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  #pragma unroll 1
  for (int w = warp; w < n_level; w += n_warps) {
    const BuildNode nd = level[w];
    const int2 r = make_int2(nd.first, nd.count);
    const S acc = orderedCovarianceSum<S>(tris, prim, r.x, r.y, lane, stage);
    S sums[9];
#pragma unroll
    for (int k = 0; k < 9; k++) sums[k] = __shfl_sync(0xffffffffu, acc, k);
    const int n_points = 3 * r.y;
    S M[3][3];
    M[0][0] = sums[3] - sums[0] * sums[0] / n_points;
    M[1][1] = sums[4] - sums[1] * sums[1] / n_points;
    M[2][2] = sums[5] - sums[2] * sums[2] / n_points;
    M[0][1] = sums[6] - sums[0] * sums[1] / n_points;
    M[1][2] = sums[8] - sums[1] * sums[2] / n_points;
    M[0][2] = sums[7] - sums[0] * sums[2] / n_points;
    M[1][0] = M[0][1];
    M[2][0] = M[0][2];
    M[2][1] = M[1][2];
    S d[3] = {0, 0, 0}, vec[3][3];
    if (!hostbuild::jacobi3<S>(M, d, vec)) {
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) vec[i][j] = (i == j) ? S(1) : S(0);
    }
    S ax[9];
    hostbuild::axesFromEigen<S>(vec, d, ax);
    const S big = sizeof(S) == 4 ? S(3.402823466e+38f) : S(1.7976931348623157e+308);
    S mn[3] = {big, big, big}, mx[3] = {-big, -big, -big};
    #pragma unroll 1
    for (int i = lane; i < 3 * r.y; i += 32) {
      const S* p = tris + size_t(12) * size_t(prim[r.x + i / 3]) + 4 * (i % 3);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const S proj = (ax[0 + k] * p[0] + ax[3 + k] * p[1]) + ax[6 + k] * p[2];
        if (proj > mx[k]) mx[k] = proj;
        if (proj < mn[k]) mn[k] = proj;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const S omx = __shfl_xor_sync(0xffffffffu, mx[k], off), omn = __shfl_xor_sync(0xffffffffu, mn[k], off);
        if (omx > mx[k]) mx[k] = omx;
        if (omn < mn[k]) mn[k] = omn;
      }
    if (lane == 0) {
      S* o = nodes + size_t(16) * nd.id;
      for (int k = 0; k < 9; k++) o[k] = ax[k];
      S c[3];
      for (int k = 0; k < 3; k++) c[k] = (mx[k] + mn[k]) / 2;
      for (int rr = 0; rr < 3; rr++) o[9 + rr] = (ax[3 * rr] * c[0] + ax[3 * rr + 1] * c[1]) + ax[3 * rr + 2] * c[2];
      for (int k = 0; k < 3; k++) o[12 + k] = (mx[k] - mn[k]) / 2;
      const int fc = nd.count == 1 ? -(prim[nd.first] + 1) : 1 + 2 * nd.rank;
      if (sizeof(S) == 4)
        o[15] = S(__int_as_float(fc));
      else
        o[15] = S(__longlong_as_double((long long)fc));
      range[nd.id] = r;
    }
    __syncwarp();
  }
}

template <typename S>
__global__ void __launch_bounds__(256) bvhBuildSplitKernel(const BuildNode* __restrict__ level, int n_level, const S* __restrict__ nodes,
                                                           const S* __restrict__ tris, int* prim, BuildNode* __restrict__ next,
                                                           int* __restrict__ next_count) {
  __shared__ S s_stage[8][32 * 9];
  S* stage = s_stage[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  #pragma unroll 1
  for (int w = warp; w < n_level; w += n_warps) {
    const BuildNode nd = level[w];
    if (nd.count <= 1) continue;  // (warp-uniform)
    const S* o = nodes + size_t(16) * nd.id;
    const S sv[3] = {o[0], o[3], o[6]};  // the box's first axis
    // mean of the centroids, summed in primitive order (computeRule_mean)
    const S acc = orderedCentroidSum<S>(tris, prim, nd.first, nd.count, lane, stage);
    const S c0 = __shfl_sync(0xffffffffu, acc, 0), c1v = __shfl_sync(0xffffffffu, acc, 1), c2 = __shfl_sync(0xffffffffu, acc, 2);
    const S split_value = (c0 * sv[0] + c1v * sv[1] + c2 * sv[2]) / nd.count;
    // the swap pass, replayed in order
    int c1 = 0;
    #pragma unroll 1
    for (int base = 0; base < nd.count; base += 32) {
      const int i = base + lane;
      bool left = false;
      if (i < nd.count) {
        const S* t = tris + size_t(12) * size_t(prim[nd.first + i]);
        S p[3];
#pragma unroll
        for (int k = 0; k < 3; k++) p[k] = ((t[k] + t[4 + k]) + t[8 + k]) / S(3.0);
        left = !(((sv[0] * p[0] + sv[1] * p[1]) + sv[2] * p[2]) > split_value);
      }
      const unsigned lm = __ballot_sync(0xffffffffu, left);
      if (lane == 0) {
        unsigned m = lm;
        while (m) {
          const int bpos = __ffs(m) - 1;
          m &= m - 1;
          const int ii = nd.first + base + bpos, cc = nd.first + c1;
          const int tmp = prim[ii];
          prim[ii] = prim[cc];
          prim[cc] = tmp;
          c1++;
        }
      }
      c1 = __shfl_sync(0xffffffffu, c1, 0);
      __syncwarp();
    }
    if (c1 == 0 || c1 == nd.count) c1 = nd.count / 2;
    if (lane == 0) {
      const int slot = atomicAdd(next_count, 2);
      next[slot] = BuildNode{1 + 2 * nd.rank, nd.first, c1, nd.rank + 1};
      next[slot + 1] = BuildNode{2 + 2 * nd.rank, nd.first + c1, nd.count - c1, nd.rank + c1};
    }
    __syncwarp();
  }
}

template <typename S>
static int buildObbTreeOnDevice(Engine& e, const double* verts_d, int n_verts, const int32_t* tris, int n_tris,
                                hostbuild::TreeOut<S>& out, float* build_ms) {
  (void)n_verts;
  const size_t n_nodes = size_t(2) * n_tris - 1;
  out.tri.resize(size_t(9) * n_tris);
  std::vector<S> tris12(size_t(12) * n_tris, S(0));
  #pragma unroll 1
  for (int t = 0; t < n_tris; t++)
    for (int v = 0; v < 3; v++)
      for (int k = 0; k < 3; k++) {
        const S x = S(verts_d[size_t(3) * tris[size_t(3) * t + v] + k]);
        out.tri[size_t(9) * t + 3 * v + k] = x;
        tris12[size_t(12) * t + 4 * v + k] = x;
      }
  std::vector<int> ident(static_cast<size_t>(n_tris));
  #pragma unroll 1
  for (int i = 0; i < n_tris; i++) ident[static_cast<size_t>(i)] = i;
  struct Scratch {
    std::vector<void*> p;
    ~Scratch() {
      #pragma unroll 1
      for (void* q : p) cudaFree(q);
    }
  } sc;
  auto alloc = [&](void** ptr, size_t bytes) {
    const cudaError_t err = cudaMalloc(ptr, bytes);
    if (err == cudaSuccess) sc.p.push_back(*ptr);
    return err;
  };
  S *d_tris = nullptr, *d_nodes = nullptr;
  int *d_prim = nullptr, *d_next_count = nullptr;
  int2* d_range = nullptr;
  BuildNode *d_a = nullptr, *d_b = nullptr;
  FCLB_CUDA(alloc(reinterpret_cast<void**>(&d_tris), tris12.size() * sizeof(S)));
  FCLB_CUDA(alloc(reinterpret_cast<void**>(&d_nodes), n_nodes * 16 * sizeof(S)));
  FCLB_CUDA(alloc(reinterpret_cast<void**>(&d_prim), size_t(n_tris) * sizeof(int)));
  FCLB_CUDA(alloc(reinterpret_cast<void**>(&d_range), n_nodes * sizeof(int2)));
  FCLB_CUDA(alloc(reinterpret_cast<void**>(&d_a), size_t(n_tris) * sizeof(BuildNode)));
  FCLB_CUDA(alloc(reinterpret_cast<void**>(&d_b), size_t(n_tris) * sizeof(BuildNode)));
  FCLB_CUDA(alloc(reinterpret_cast<void**>(&d_next_count), sizeof(int)));
  FCLB_CUDA(cudaMemcpyAsync(d_tris, tris12.data(), tris12.size() * sizeof(S), cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(d_prim, ident.data(), ident.size() * sizeof(int), cudaMemcpyHostToDevice, e.compute));
  const BuildNode root{0, 0, n_tris, 0};
  FCLB_CUDA(cudaMemcpyAsync(d_a, &root, sizeof(root), cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  int n_level = 1, levels = 0;
  size_t done = 0;
  while (n_level > 0) {
    if (++levels > 8192) return fail(FCLB_ERR_CAPACITY, "fclb_bvh_build_device: tree deeper than 8192 levels");
    const int grid = int(std::min<size_t>((size_t(n_level) + 7) / 8, size_t(e.sms) * 16));
    bvhBuildFitKernel<S><<<grid, 256, 0, e.compute>>>(d_a, n_level, d_nodes, d_tris, d_prim, d_range);
    FCLB_CUDA(cudaMemsetAsync(d_next_count, 0, sizeof(int), e.compute));
    bvhBuildSplitKernel<S><<<grid, 256, 0, e.compute>>>(d_a, n_level, d_nodes, d_tris, d_prim, d_b, d_next_count);
    FCLB_CUDA(cudaGetLastError());
    e.launches += 2;
    done += size_t(n_level);
    int next = 0;
    FCLB_CUDA(cudaMemcpyAsync(&next, d_next_count, sizeof(int), cudaMemcpyDeviceToHost, e.compute));
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    n_level = next;
    std::swap(d_a, d_b);
  }
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  if (done != n_nodes) return fail(FCLB_ERR_CUDA, "fclb_bvh_build_device: the build visited " + std::to_string(done) + " nodes, expected " + std::to_string(n_nodes));
  if (build_ms) cudaEventElapsedTime(build_ms, e.ev0, e.ev1);
  std::vector<S> nodes(n_nodes * 16);
  FCLB_CUDA(cudaMemcpy(nodes.data(), d_nodes, nodes.size() * sizeof(S), cudaMemcpyDeviceToHost));
  out.obb.resize(n_nodes * 15);
  out.first_child.resize(n_nodes);
  #pragma unroll 1
  for (size_t i = 0; i < n_nodes; i++) {
    for (int k = 0; k < 15; k++) out.obb[15 * i + k] = nodes[16 * i + k];
    if (sizeof(S) == 4) {
      int32_t v;
      memcpy(&v, &nodes[16 * i + 15], 4);
      out.first_child[i] = v;
    } else {
      long long v;
      memcpy(&v, &nodes[16 * i + 15], 8);
      out.first_child[i] = int32_t(v);
    }
  }
  return FCLB_OK;
}

// parent links and the leaf of every triangle, for the bottom-up refit
static int ensureParents(BvhDev* d) {
  if (d->d_parent) return FCLB_OK;
  if (int(d->h_child.size()) != d->n_nodes) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_refit: the BVH has no host copy of its child links");
  std::vector<int> parent(size_t(d->n_nodes), 0), leaf(size_t(d->n_tris), 0);
  #pragma unroll 1
  for (int i = 0; i < d->n_nodes; i++) {
    const int fc = d->h_child[size_t(i)];
    if (fc < 0) {
      leaf[size_t(-(fc + 1))] = i;
    } else {
      parent[size_t(fc)] = i;
      parent[size_t(fc) + 1] = i;
    }
  }
  FCLB_CUDA(cudaMalloc(&d->d_parent, parent.size() * sizeof(int)));
  FCLB_CUDA(cudaMalloc(&d->d_leaf_node, leaf.size() * sizeof(int)));
  FCLB_CUDA(cudaMalloc(&d->d_arrived, parent.size() * sizeof(int)));
  FCLB_CUDA(uploadSync(d->d_parent, parent.data(), parent.size() * sizeof(int)));
  FCLB_CUDA(uploadSync(d->d_leaf_node, leaf.data(), leaf.size() * sizeof(int)));
  return FCLB_OK;
}

template <typename S>
static int refitDev(Engine& e, BvhDev* d, const void* d_tri9, int bottomup) {
  const int g = int(std::min<size_t>((size_t(d->n_tris) * 3 + 255) / 256, size_t(e.sms) * 8));
  triRepackKernel<S><<<g, 256, 0, e.compute>>>(static_cast<const S*>(d_tri9), d->n_tris, static_cast<S*>(d->tris));
  const int warps_per_cta = 8;
  const int grid = int(std::min<size_t>((size_t(d->n_nodes) + warps_per_cta - 1) / warps_per_cta, size_t(e.sms) * 16));
  if (bottomup) {
    const int rc = ensureParents(d);
    if (rc) return rc;
    FCLB_CUDA(cudaMemsetAsync(d->d_arrived, 0, size_t(d->n_nodes) * sizeof(int), e.compute));
  }
  FCLB_CUDA(cudaEventRecord(e.ev0, e.compute));
  if (bottomup)
    bvhRefitBottomUpKernel<S><<<(d->n_tris + 127) / 128, 128, 0, e.compute>>>(static_cast<S*>(d->nodes), static_cast<const S*>(d->tris),
                                                                            d->d_leaf_node, d->d_parent, d->d_arrived, d->n_tris);
  else
    bvhRefitKernel<S><<<grid, warps_per_cta * 32, 0, e.compute>>>(static_cast<S*>(d->nodes), static_cast<const S*>(d->tris), d->d_range,
                                                                 d->d_prim, d->n_nodes);
  FCLB_CUDA(cudaEventRecord(e.ev1, e.compute));
  e.launches += 2;
  FCLB_CUDA(cudaGetLastError());
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.last_ms = e.last_call_ms = ms;
  // keep the host mirror of fclb_bvh_export current
  std::vector<S> nodes(size_t(16) * d->n_nodes);
  FCLB_CUDA(cudaMemcpy(nodes.data(), d->nodes, nodes.size() * sizeof(S), cudaMemcpyDeviceToHost));
  S* ho = reinterpret_cast<S*>(d->h_obb.data());
  #pragma unroll 1
  for (int i = 0; i < d->n_nodes; i++)
    for (int k = 0; k < 15; k++) ho[size_t(15) * i + k] = nodes[size_t(16) * i + k];
  FCLB_CUDA(cudaMemcpy(d->h_tri.data(), d_tri9, d->h_tri.size(), cudaMemcpyDeviceToHost));
  return FCLB_OK;
}

template <typename S>
static int uploadBvh(BvhDev* d, const void* obb, const int32_t* first_child, int n_nodes, const void* tri, int n_tris) {
  const S* o = static_cast<const S*>(obb);
  const S* tv = static_cast<const S*>(tri);
  std::vector<S> nodes(size_t(16) * n_nodes), tris(size_t(12) * n_tris, S(0));
  #pragma unroll 1
  for (int i = 0; i < n_nodes; i++) {
    for (int k = 0; k < 15; k++) nodes[size_t(16) * i + k] = o[size_t(15) * i + k];
    S bits;
    if (sizeof(S) == 4) {
      const int32_t v = first_child[i];
      memcpy(&bits, &v, 4);
    } else {
      const long long v = first_child[i];
      memcpy(&bits, &v, 8);
    }
    nodes[size_t(16) * i + 15] = bits;
  }
  #pragma unroll 1
  for (int t = 0; t < n_tris; t++)
    for (int v = 0; v < 3; v++)
      for (int k = 0; k < 3; k++) tris[size_t(12) * t + 4 * v + k] = tv[size_t(9) * t + 3 * v + k];
  FCLB_CUDA(cudaMalloc(&d->nodes, nodes.size() * sizeof(S)));
  FCLB_CUDA(cudaMalloc(&d->tris, tris.size() * sizeof(S)));
  FCLB_CUDA(uploadSync(d->nodes, nodes.data(), nodes.size() * sizeof(S)));
  FCLB_CUDA(uploadSync(d->tris, tris.data(), tris.size() * sizeof(S)));
  d->h_obb.assign(reinterpret_cast<const unsigned char*>(o), reinterpret_cast<const unsigned char*>(o + size_t(15) * n_nodes));
  d->h_tri.assign(reinterpret_cast<const unsigned char*>(tv), reinterpret_cast<const unsigned char*>(tv + size_t(9) * n_tris));
  d->h_child.assign(first_child, first_child + n_nodes);
  std::vector<int2> range;
  std::vector<int> prim;
  deriveRanges(first_child, n_nodes, n_tris, range, prim);
  if (int(prim.size()) != n_tris) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_upload: the leaves do not cover every triangle exactly once");
  FCLB_CUDA(cudaMalloc(&d->d_range, range.size() * sizeof(int2)));
  FCLB_CUDA(cudaMalloc(&d->d_prim, prim.size() * sizeof(int)));
  FCLB_CUDA(uploadSync(d->d_range, range.data(), range.size() * sizeof(int2)));
  FCLB_CUDA(uploadSync(d->d_prim, prim.data(), prim.size() * sizeof(int)));
  return FCLB_OK;
}

}  // namespace fclb

using namespace fclb;

extern "C" {

static int bvh_upload_one(const void* obb, const int32_t* first_child, int n_nodes, const void* tri_verts, int n_tris,
                    int scalar_type, fclb_handle* h) {
  int rc = ensureInit();
  if (rc) return rc;
  if (!obb || !first_child || !tri_verts || !h || n_nodes <= 0 || n_tris <= 0)
    return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_upload: null or empty input");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  #pragma unroll 1
  for (int i = 0; i < n_nodes; i++) {
    const int fc = first_child[i];
    // children follow their parent in the array (BVHModel::recursiveBuildTree numbers them that way,
    // BVH_model-inl.h:470-560): anything else could be a cycle or a shared subtree
    if (fc >= 0 ? (fc + 1 >= n_nodes || fc <= i) : (-(fc + 1) >= n_tris))
      return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_upload: child / primitive index out of range (children must follow their parent)");
  }
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  BvhDev* d = new BvhDev();
  d->n_nodes = n_nodes;
  d->n_tris = n_tris;
  d->scalar_type = scalar_type;
  rc = scalar_type == FCLB_F32 ? uploadBvh<float>(d, obb, first_child, n_nodes, tri_verts, n_tris)
                               : uploadBvh<double>(d, obb, first_child, n_nodes, tri_verts, n_tris);
  if (rc) {
    delete d;
    return rc;
  }
  const fclb_handle hd = newHandle();
  bvhTable()[hd] = d;
  *h = hd;
  return FCLB_OK;
}
int fclb_bvh_upload(const void* obb, const int32_t* first_child, int n_nodes, const void* tri_verts, int n_tris,
                    int scalar_type, fclb_handle* h) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return bvh_upload_one(obb, first_child, n_nodes, tri_verts, n_tris, scalar_type, h); });
}

// the tree is built once on the host (mirror of BVHModel::endModel) and uploaded to every device
int fclb_bvh_build(const double* verts, int n_verts, const int32_t* tris, int n_tris, int scalar_type, fclb_handle* h) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  if (!verts || !tris || !h || n_verts <= 0 || n_tris <= 0) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_build: null or empty input");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  #pragma unroll 1
  for (size_t i = 0; i < size_t(3) * n_tris; i++)
    if (tris[i] < 0 || tris[i] >= n_verts) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_build: vertex index out of range");
  if (scalar_type == FCLB_F32) {
    hostbuild::TreeOut<float> t;
    hostbuild::buildObbTree<float>(verts, n_verts, tris, n_tris, t);
    return forEachDevice([&] {
      return bvh_upload_one(t.obb.data(), t.first_child.data(), int(t.first_child.size()), t.tri.data(), n_tris, scalar_type, h);
    });
  }
  hostbuild::TreeOut<double> t;
  hostbuild::buildObbTree<double>(verts, n_verts, tris, n_tris, t);
  return forEachDevice([&] {
    return bvh_upload_one(t.obb.data(), t.first_child.data(), int(t.first_child.size()), t.tri.data(), n_tris, scalar_type, h);
  });
}

// the same tree, built on the device (level by level, see bvhBuildFitKernel / bvhBuildSplitKernel)
int fclb_bvh_build_device(const double* verts, int n_verts, const int32_t* tris, int n_tris, int scalar_type, fclb_handle* h) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  if (!verts || !tris || !h || n_verts <= 0 || n_tris <= 0) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_build_device: null or empty input");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  #pragma unroll 1
  for (size_t i = 0; i < size_t(3) * n_tris; i++)
    if (tris[i] < 0 || tris[i] >= n_verts) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_build_device: vertex index out of range");
  float ms = 0.f;
  int rc;
  if (scalar_type == FCLB_F32) {
    hostbuild::TreeOut<float> t;
    {
      Engine& e = eng();
      std::lock_guard<std::recursive_mutex> lk(e.mu);
      rc = buildObbTreeOnDevice<float>(e, verts, n_verts, tris, n_tris, t, &ms);
    }
    if (rc) return rc;
    rc = forEachDevice([&] {
      return bvh_upload_one(t.obb.data(), t.first_child.data(), int(t.first_child.size()), t.tri.data(), n_tris, scalar_type, h);
    });
  } else {
    hostbuild::TreeOut<double> t;
    {
      Engine& e = eng();
      std::lock_guard<std::recursive_mutex> lk(e.mu);
      rc = buildObbTreeOnDevice<double>(e, verts, n_verts, tris, n_tris, t, &ms);
    }
    if (rc) return rc;
    rc = forEachDevice([&] {
      return bvh_upload_one(t.obb.data(), t.first_child.data(), int(t.first_child.size()), t.tri.data(), n_tris, scalar_type, h);
    });
  }
  if (!rc) eng().last_ms = eng().last_call_ms = ms;  // the build launches alone (upload excluded)
  return rc;
}

int fclb_bvh_build_host(const double* verts, int n_verts, const int32_t* tris, int n_tris, int scalar_type, void* obb,
                        int32_t* first_child, void* tri_verts, int* n_nodes) {
  if (!verts || !tris || n_verts <= 0 || n_tris <= 0) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_build_host: null or empty input");
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  #pragma unroll 1
  for (size_t i = 0; i < size_t(3) * n_tris; i++)
    if (tris[i] < 0 || tris[i] >= n_verts) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_build_host: vertex index out of range");
  auto emit = [&](auto& t) {
    if (obb) memcpy(obb, t.obb.data(), t.obb.size() * sizeof(t.obb[0]));
    if (first_child) memcpy(first_child, t.first_child.data(), t.first_child.size() * sizeof(int32_t));
    if (tri_verts) memcpy(tri_verts, t.tri.data(), t.tri.size() * sizeof(t.tri[0]));
    if (n_nodes) *n_nodes = int(t.first_child.size());
  };
  if (scalar_type == FCLB_F32) {
    hostbuild::TreeOut<float> t;
    hostbuild::buildObbTree<float>(verts, n_verts, tris, n_tris, t);
    emit(t);
  } else {
    hostbuild::TreeOut<double> t;
    hostbuild::buildObbTree<double>(verts, n_verts, tris, n_tris, t);
    emit(t);
  }
  return FCLB_OK;
}

int fclb_bvh_info(fclb_handle h, int* n_nodes, int* n_tris, int* scalar_type) {
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = bvhTable().find(h);
  if (it == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_info: unknown handle");
  if (n_nodes) *n_nodes = it->second->n_nodes;
  if (n_tris) *n_tris = it->second->n_tris;
  if (scalar_type) *scalar_type = it->second->scalar_type;
  return FCLB_OK;
}

int fclb_bvh_export(fclb_handle h, void* obb, int32_t* first_child, void* tri_verts) {
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = bvhTable().find(h);
  if (it == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_export: unknown handle");
  const BvhDev* d = it->second;
  if (obb) memcpy(obb, d->h_obb.data(), d->h_obb.size());
  if (first_child) memcpy(first_child, d->h_child.data(), d->h_child.size() * sizeof(int32_t));
  if (tri_verts) memcpy(tri_verts, d->h_tri.data(), d->h_tri.size());
  return FCLB_OK;
}

static int bvh_release_one(fclb_handle h) {
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = bvhTable().find(h);
  if (it == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_release: unknown handle");
  cudaFree(it->second->nodes);
  cudaFree(it->second->tris);
  cudaFree(it->second->d_range);
  cudaFree(it->second->d_prim);
  cudaFree(it->second->d_dfs_rank);
  cudaFree(it->second->d_parent);
  cudaFree(it->second->d_leaf_node);
  cudaFree(it->second->d_arrived);
  delete it->second;
  bvhTable().erase(it);
  return FCLB_OK;
}
int fclb_bvh_release(fclb_handle h) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return bvh_release_one(h); });
}

// BVHModel::beginReplaceModel / replaceSubModel / endReplaceModel(refit = true, bottomup = false): the vertices move, the
// topology stays; every node OBB is fitted again on the device (bvhRefitKernel).
static int bvh_refit_one(fclb_handle bvh, const void* tri_verts, int n_tris, int on_device, int bottomup = 0) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto it = bvhTable().find(bvh);
  if (it == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_refit: unknown BVH handle");
  BvhDev* d = it->second;
  if (!tri_verts || n_tris != d->n_tris) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_refit: the replaced model must have the same triangles");
  const size_t bytes = size_t(9) * n_tris * (d->scalar_type == FCLB_F32 ? 4 : 8);
  const void* d_in = tri_verts;
  if (!on_device) {
    rc = ensureStage(e, bytes);
    if (rc) return rc;
    FCLB_CUDA(cudaMemcpyAsync(e.d_stage, tri_verts, bytes, cudaMemcpyHostToDevice, e.compute));
    d_in = e.d_stage;
  }
  return d->scalar_type == FCLB_F32 ? refitDev<float>(e, d, d_in, bottomup) : refitDev<double>(e, d, d_in, bottomup);
}
int fclb_bvh_refit_host(fclb_handle bvh, const void* tri_verts, int n_tris) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return bvh_refit_one(bvh, tri_verts, n_tris, 0); });
}
int fclb_bvh_refit_dev(fclb_handle bvh, const void* tri_verts, int n_tris) { return bvh_refit_one(bvh, tri_verts, n_tris, 1); }
int fclb_bvh_refit_bottomup_host(fclb_handle bvh, const void* tri_verts, int n_tris) {
  const int rc_init_ = ensureInit();
  if (rc_init_) return rc_init_;
  return forEachDevice([&] { return bvh_refit_one(bvh, tri_verts, n_tris, 0, 1); });
}
int fclb_bvh_refit_bottomup_dev(fclb_handle bvh, const void* tri_verts, int n_tris) { return bvh_refit_one(bvh, tri_verts, n_tris, 1, 1); }

int fclb_bvh_collide_batch_dev(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2, size_t n,
                               int scalar_type, const fclb_request* req, uint32_t* out_counts,
                               int32_t* out_first_pair) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto i1 = bvhTable().find(bvh1), i2 = bvhTable().find(bvh2);
  if (i1 == bvhTable().end() || i2 == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown BVH handle");
  if (i1->second->scalar_type != scalar_type || i2->second->scalar_type != scalar_type)
    return fail(FCLB_ERR_BAD_ARG, "BVH was uploaded for a different scalar type");
  if (!req || !out_counts) return fail(FCLB_ERR_BAD_ARG, "null request / out_counts");
  if (req->penetration_mode != FCLB_PEN_DISABLED)
    return fail(FCLB_ERR_UNSUPPORTED, "this entry point answers boolean requests: contact records (every penetration mode) come from fclb_bvh_collide_contacts_batch_*");
  if (n == 0) return FCLB_OK;
  if (!poses1 || !poses2) return fail(FCLB_ERR_BAD_ARG, "null pose array");
  if (scalar_type == FCLB_F32)
    return bvhCollideDev<float>(e, i1->second, i2->second, poses1, poses2, n, req->max_contacts, out_counts,
                                out_first_pair);
  return bvhCollideDev<double>(e, i1->second, i2->second, poses1, poses2, n, req->max_contacts, out_counts,
                               out_first_pair);
}

static int bvh_collide_batch_host_one(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2, size_t n,
                                int scalar_type, const fclb_request* req, uint32_t* out_counts,
                                int32_t* out_first_pair) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!poses1 || !poses2 || !out_counts) return fail(FCLB_ERR_BAD_ARG, "null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_p1 = 0;
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_fp = alignUp(o_cnt + n * 4, 256);
  const size_t total = alignUp(o_fp + n * 8, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  // Chunked three-stage pipeline, as fclb_distance_batch_host: every chunk's poses are queued on the copy-in stream up
  // front, the compute stream waits per chunk, the copy-out stream drains a chunk's results while later chunks upload
  // and traverse (queries are independent; per-query cost varies 100x, which the kernel's work counter absorbs).
  // (the traversal is compute-bound: the first stage is short so that the kernels start early, then the stages double)
  const size_t chunk = e.host_chunk / 4 ? e.host_chunk / 4 : 1;
  std::vector<size_t> c_begin, c_size;
  stageSizes(n, chunk, e.host_head, 0, c_begin, c_size);
  const int n_chunks = int(c_size.size());
  rc = ensureChunkEvents(e, n_chunks);
  if (rc) return rc;
  const char* h_p1 = static_cast<const char*>(poses1);
  const char* h_p2 = static_cast<const char*>(poses2);
  FCLB_CUDA(cudaStreamSynchronize(e.copy_out));  // the staging arena may still be read by an earlier call's copy-out
  #pragma unroll 1
  for (int c = 0; c < n_chunks; c++) {
    const size_t b0 = c_begin[c], m = c_size[c];
    FCLB_CUDA(cudaMemcpyAsync(base + o_p1 + b0 * 12 * ss, h_p1 + b0 * 12 * ss, m * 12 * ss, cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaMemcpyAsync(base + o_p2 + b0 * 12 * ss, h_p2 + b0 * 12 * ss, m * 12 * ss, cudaMemcpyHostToDevice, e.copy_in));
    FCLB_CUDA(cudaEventRecord(e.ev_in[c], e.copy_in));
  }
  unsigned long long visits[2] = {0, 0};
  #pragma unroll 1
  for (int c = 0; c < n_chunks; c++) {
    const size_t b0 = c_begin[c], m = c_size[c];
    FCLB_CUDA(cudaStreamWaitEvent(e.compute, e.ev_in[c], 0));
    rc = fclb_bvh_collide_batch_dev(bvh1, bvh2, base + o_p1 + b0 * 12 * ss, base + o_p2 + b0 * 12 * ss, m, scalar_type, req,
                                    reinterpret_cast<uint32_t*>(base + o_cnt) + b0,
                                    out_first_pair ? reinterpret_cast<int32_t*>(base + o_fp) + 2 * b0 : nullptr);
    if (rc) {  // drain the queued copies before the caller gets its buffers back
      cudaStreamSynchronize(e.copy_in);
      cudaStreamSynchronize(e.copy_out);
      return rc;
    }
    visits[0] += g_last_stats[0];
    visits[1] += g_last_stats[1];
    FCLB_CUDA(cudaEventRecord(e.ev_done[c], e.compute));
    FCLB_CUDA(cudaStreamWaitEvent(e.copy_out, e.ev_done[c], 0));
    FCLB_CUDA(cudaMemcpyAsync(out_counts + b0, base + o_cnt + b0 * 4, m * 4, cudaMemcpyDeviceToHost, e.copy_out));
    if (out_first_pair)
      FCLB_CUDA(cudaMemcpyAsync(out_first_pair + 2 * b0, base + o_fp + b0 * 8, m * 8, cudaMemcpyDeviceToHost, e.copy_out));
  }
  g_last_stats[0] = visits[0];  // fclb_bvh_last_visit_counts: the whole call
  g_last_stats[1] = visits[1];
  FCLB_CUDA(cudaStreamSynchronize(e.copy_out));
  return FCLB_OK;
}
int fclb_bvh_collide_batch_host(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2, size_t n,
                                int scalar_type, const fclb_request* req, uint32_t* out_counts,
                                int32_t* out_first_pair) {
  if (engineCount() <= 1) return bvh_collide_batch_host_one(bvh1, bvh2, poses1, poses2, n, scalar_type, req, out_counts, out_first_pair);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return bvh_collide_batch_host_one(bvh1, bvh2, offPtr(poses1, b * 12 * ss), offPtr(poses2, b * 12 * ss), m_, scalar_type, req, offT(out_counts, b), offT(out_first_pair, 2 * b)); });
}

int fclb_bvh_collide_contacts_batch_dev(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2, size_t n,
                                        int scalar_type, const fclb_request* req, uint32_t max_keep, uint32_t* out_counts,
                                        int32_t* out_ids, void* out_contacts) {
  int rc = ensureInit();
  if (rc) return rc;
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  auto i1 = bvhTable().find(bvh1), i2 = bvhTable().find(bvh2);
  if (i1 == bvhTable().end() || i2 == bvhTable().end()) return fail(FCLB_ERR_BAD_ARG, "unknown BVH handle");
  if (i1->second->scalar_type != scalar_type || i2->second->scalar_type != scalar_type)
    return fail(FCLB_ERR_BAD_ARG, "BVH was uploaded for a different scalar type");
  if (!req || !out_counts || !out_ids || !out_contacts || max_keep == 0) return fail(FCLB_ERR_BAD_ARG, "null output / max_keep == 0");
  if (req->penetration_mode > FCLB_PEN_INCREMENTAL_MIN) return fail(FCLB_ERR_BAD_ARG, "fclb_bvh_collide_contacts_batch: bad penetration_mode");
  if (n == 0) return FCLB_OK;
  if (!poses1 || !poses2) return fail(FCLB_ERR_BAD_ARG, "null pose array");
  if (req->penetration_mode == FCLB_PEN_DIRECTED || req->penetration_mode == FCLB_PEN_INCREMENTAL_MIN) {
    // boolean collide first (ids of the colliding triangle pairs), then one MPR penetration per kept pair
    rc = scalar_type == FCLB_F32 ? bvhCollideDev<float>(e, i1->second, i2->second, poses1, poses2, n, req->max_contacts, out_counts,
                                                        nullptr, false, max_keep, out_ids, out_contacts)
                                 : bvhCollideDev<double>(e, i1->second, i2->second, poses1, poses2, n, req->max_contacts, out_counts,
                                                         nullptr, false, max_keep, out_ids, out_contacts);
    if (rc) return rc;
    const size_t total = n * size_t(max_keep);
    const int grid = int(std::min<size_t>((total + kBlock - 1) / kBlock, size_t(e.sms) * 16));
    const int inc = req->penetration_mode == FCLB_PEN_INCREMENTAL_MIN ? 1 : 0;
    const double tol = req->distance_tol > 0 ? req->distance_tol : 1e-6;  // MPR(128, request.distanceTolerance())
    if (scalar_type == FCLB_F32)
      bvhPairPenetrationKernel<float><<<grid, kBlock, 0, e.compute>>>(
          static_cast<const float*>(i1->second->tris), static_cast<const float*>(i2->second->tris), static_cast<const float*>(poses1),
          static_cast<const float*>(poses2), n, max_keep, out_counts, out_ids, inc, req->dir[0], req->dir[1], req->dir[2], tol,
          static_cast<float*>(out_contacts));
    else
      bvhPairPenetrationKernel<double><<<grid, kBlock, 0, e.compute>>>(
          static_cast<const double*>(i1->second->tris), static_cast<const double*>(i2->second->tris),
          static_cast<const double*>(poses1), static_cast<const double*>(poses2), n, max_keep, out_counts, out_ids, inc, req->dir[0],
          req->dir[1], req->dir[2], tol, static_cast<double*>(out_contacts));
    FCLB_CUDA(cudaGetLastError());
    e.launches += 1;
    FCLB_CUDA(cudaStreamSynchronize(e.compute));
    return FCLB_OK;
  }
  const bool pen = req->penetration_mode == FCLB_PEN_DEFAULT_GJK_EPA;
  if (scalar_type == FCLB_F32)
    return bvhCollideDev<float>(e, i1->second, i2->second, poses1, poses2, n, req->max_contacts, out_counts, nullptr, pen,
                                max_keep, out_ids, out_contacts);
  return bvhCollideDev<double>(e, i1->second, i2->second, poses1, poses2, n, req->max_contacts, out_counts, nullptr, pen,
                               max_keep, out_ids, out_contacts);
}

static int bvh_collide_contacts_batch_host_one(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2, size_t n,
                                         int scalar_type, const fclb_request* req, uint32_t max_keep, uint32_t* out_counts,
                                         int32_t* out_ids, void* out_contacts) {
  int rc = ensureInit();
  if (rc) return rc;
  if (scalar_type != FCLB_F32 && scalar_type != FCLB_F64) return fail(FCLB_ERR_BAD_ARG, "bad scalar_type");
  if (n == 0) return FCLB_OK;
  if (!poses1 || !poses2 || !out_counts || !out_ids || !out_contacts || max_keep == 0) return fail(FCLB_ERR_BAD_ARG, "null array");
  Engine& e = eng();
  std::lock_guard<std::recursive_mutex> lk(e.mu);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  const size_t o_p1 = 0;
  const size_t o_p2 = alignUp(o_p1 + n * 12 * ss, 256);
  const size_t o_cnt = alignUp(o_p2 + n * 12 * ss, 256);
  const size_t o_ids = alignUp(o_cnt + n * 4, 256);
  const size_t o_ct = alignUp(o_ids + n * size_t(max_keep) * 8, 256);
  const size_t total = alignUp(o_ct + n * size_t(max_keep) * 7 * ss, 256);
  rc = ensureStage(e, total);
  if (rc) return rc;
  char* base = static_cast<char*>(e.d_stage);
  FCLB_CUDA(cudaMemcpyAsync(base + o_p1, poses1, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(base + o_p2, poses2, n * 12 * ss, cudaMemcpyHostToDevice, e.compute));
  FCLB_CUDA(cudaMemsetAsync(base + o_ids, 0xff, n * size_t(max_keep) * 8, e.compute));
  FCLB_CUDA(cudaMemsetAsync(base + o_ct, 0, n * size_t(max_keep) * 7 * ss, e.compute));
  rc = fclb_bvh_collide_contacts_batch_dev(bvh1, bvh2, base + o_p1, base + o_p2, n, scalar_type, req, max_keep,
                                           reinterpret_cast<uint32_t*>(base + o_cnt), reinterpret_cast<int32_t*>(base + o_ids),
                                           base + o_ct);
  if (rc) return rc;
  FCLB_CUDA(cudaMemcpyAsync(out_counts, base + o_cnt, n * 4, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(out_ids, base + o_ids, n * size_t(max_keep) * 8, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaMemcpyAsync(out_contacts, base + o_ct, n * size_t(max_keep) * 7 * ss, cudaMemcpyDeviceToHost, e.compute));
  FCLB_CUDA(cudaStreamSynchronize(e.compute));
  return FCLB_OK;
}
int fclb_bvh_collide_contacts_batch_host(fclb_handle bvh1, fclb_handle bvh2, const void* poses1, const void* poses2, size_t n,
                                         int scalar_type, const fclb_request* req, uint32_t max_keep, uint32_t* out_counts,
                                         int32_t* out_ids, void* out_contacts) {
  if (engineCount() <= 1) return bvh_collide_contacts_batch_host_one(bvh1, bvh2, poses1, poses2, n, scalar_type, req, max_keep, out_counts, out_ids, out_contacts);
  const size_t ss = scalar_type == FCLB_F32 ? 4 : 8;
  (void)ss;
  return shardOverDevices(n, [&](size_t b, size_t m_) { return bvh_collide_contacts_batch_host_one(bvh1, bvh2, offPtr(poses1, b * 12 * ss), offPtr(poses2, b * 12 * ss), m_, scalar_type, req, max_keep, offT(out_counts, b), offT(out_ids, b * size_t(max_keep) * 2), offPtr(out_contacts, b * size_t(max_keep) * 7 * ss)); });
}

int fclb_bvh_last_visit_counts(uint64_t* n_bv, uint64_t* n_leaf) {
  if (n_bv) *n_bv = g_last_stats[0];
  if (n_leaf) *n_leaf = g_last_stats[1];
  return FCLB_OK;
}

}  // extern "C"
