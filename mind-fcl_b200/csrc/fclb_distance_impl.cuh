// fclb_distance_impl.cuh -- batched fcl::distance (== GJKSolver<S>::shapeDistance,
// reference gjk_solver-inl.h:762-808, closed forms :902-988).
//
// One query per thread; one launch per (type1,type2) bucket so a warp never
// mixes pair kinds.  Closed-form pairs are HBM-bound streaming kernels
// (2 poses in, dist + 2 witness points + flag out); GJK pairs are FP-bound and
// use the shared-memory simplex store of fclb_gjk.cuh.
#pragma once
#include <cstdlib>

#include "fclb_gjk.cuh"
#include "fclb_internal.h"
#include "fclb_primitives.cuh"

namespace fclb {

template <typename S>
FCLB_DI void writeDistance(const DistanceOut& o, size_t q, S d, const V3<S>& p1, const V3<S>& p2, uint8_t ok) {
  if (o.dist) static_cast<S*>(o.dist)[q] = d;
  if (o.p1) store3(static_cast<S*>(o.p1), q, p1);
  if (o.p2) store3(static_cast<S*>(o.p2), q, p2);
  if (o.ok) o.ok[q] = ok;
}

// ---- closed-form bucket (block-uniform `kind`) -----------------------------
enum ClosedKind : int {
  CK_NONE = 0,
  CK_SPHERE_BOX,
  CK_BOX_SPHERE,
  CK_SPHERE_CAPSULE,
  CK_CAPSULE_SPHERE,
  CK_SPHERE_CYLINDER,
  CK_CYLINDER_SPHERE,
  CK_SPHERE_SPHERE,
  CK_CAPSULE_CAPSULE
};
inline int closedKindOf(int t1, int t2) {
  if (t1 == ST_SPHERE && t2 == ST_BOX) return CK_SPHERE_BOX;
  if (t1 == ST_BOX && t2 == ST_SPHERE) return CK_BOX_SPHERE;
  if (t1 == ST_SPHERE && t2 == ST_CAPSULE) return CK_SPHERE_CAPSULE;
  if (t1 == ST_CAPSULE && t2 == ST_SPHERE) return CK_CAPSULE_SPHERE;
  if (t1 == ST_SPHERE && t2 == ST_CYLINDER) return CK_SPHERE_CYLINDER;
  if (t1 == ST_CYLINDER && t2 == ST_SPHERE) return CK_CYLINDER_SPHERE;
  if (t1 == ST_SPHERE && t2 == ST_SPHERE) return CK_SPHERE_SPHERE;
  if (t1 == ST_CAPSULE && t2 == ST_CAPSULE) return CK_CAPSULE_CAPSULE;
  return CK_NONE;
}

template <typename S, int CK>
__global__ void __launch_bounds__(kBlock) distanceClosedKernel(BatchView b, S eps78, DistanceOut out) {
  const ShapeD<S>* __restrict__ shapes = static_cast<const ShapeD<S>*>(b.shapes);
  const S* __restrict__ poses1 = static_cast<const S*>(b.poses1);
  const S* __restrict__ poses2 = static_cast<const S*>(b.poses2);
  #pragma unroll 1
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < b.count; i += size_t(gridDim.x) * blockDim.x) {
    const size_t q = b.perm ? size_t(b.perm[b.begin + i]) : (b.begin + i);
    const fclb_pair pr = b.pairs[q];
    const ShapeD<S> a = shapes[pr.shape1];
    const ShapeD<S> c = shapes[pr.shape2];
    const Pose<S> tf1 = loadPose(poses1, q);
    const Pose<S> tf2 = loadPose(poses2, q);
    S d = S(-1);
    V3<S> p1 = zero3<S>(), p2 = zero3<S>();
    bool ok = false;
    if (CK == CK_SPHERE_BOX) {
      ok = sphereBoxDistance(a.p[0], tf1, mk<S>(c.p[0], c.p[1], c.p[2]), tf2, d, p1, p2);
    } else if (CK == CK_BOX_SPHERE) {
      ok = sphereBoxDistance(c.p[0], tf2, mk<S>(a.p[0], a.p[1], a.p[2]), tf1, d, p2, p1);
    } else if (CK == CK_SPHERE_CAPSULE) {
      ok = sphereCapsuleDistance(a.p[0], tf1, c.p[0], c.p[1], tf2, d, p1, p2);
    } else if (CK == CK_CAPSULE_SPHERE) {
      ok = sphereCapsuleDistance(c.p[0], tf2, a.p[0], a.p[1], tf1, d, p2, p1);
    } else if (CK == CK_SPHERE_CYLINDER) {
      ok = sphereCylinderDistance(a.p[0], tf1, c.p[0], c.p[1], tf2, d, p1, p2);
    } else if (CK == CK_CYLINDER_SPHERE) {
      ok = sphereCylinderDistance(c.p[0], tf2, a.p[0], a.p[1], tf1, d, p2, p1);
    } else if (CK == CK_SPHERE_SPHERE) {
      ok = sphereSphereDistance(a.p[0], tf1, c.p[0], tf2, d, p1, p2);
    } else if (CK == CK_CAPSULE_CAPSULE) {
      ok = capsuleCapsuleDistance(a.p[0], a.p[1], tf1, c.p[0], c.p[1], tf2, eps78, d, p1, p2);
    }
    writeDistance(out, q, d, p1, p2, ok ? uint8_t(1) : uint8_t(0));
  }
}

// ---- GJK bucket -------------------------------------------------------------
// ok: 0 = not separated (dist = -1); 1 = separated, witness points valid;
//     3 = separated but the reference's witness extraction reported invalid
//         (it then returns uninitialised points; we return zeros mapped by tf1).
//
// One query per thread, written as a STATE MACHINE around a single support
// site.  The reference runs three loops back to back -- boolean GJK
// (gjk.hpp:90-137), the distance refinement (gjk_distance.hpp:48-105) and the
// witness extraction (:376-470, up to six more supports).  A straight
// translation leaves each lane of a warp in a different loop at a different
// iteration: the first version of this kernel ran at 6.8 of 32 lanes active and
// was instruction-fetch bound (ncu: stall_no_inst 75 %, 7k SASS lines).  Every
// one of those loops is "evaluate support0(d) and support1(-d), then update
// some state", so here all lanes meet at ONE support evaluation per trip and
// only the (short) state update diverges; a lane that finishes a query
// immediately fetches its next one (persistent, strided), so early finishers do
// not idle.  Arithmetic and decision order per query are unchanged.
#ifndef FCLB_GJK_MIN_BLOCKS
#define FCLB_GJK_MIN_BLOCKS 6
#endif
enum GjkPhase : int { PH_FETCH = 0, PH_BOOL_FIRST = 1, PH_BOOL = 2, PH_DIST = 3, PH_EXTRACT = 4 };

template <typename S, int T0, int T1>
__global__ void __launch_bounds__(kBlock, FCLB_GJK_MIN_BLOCKS) distanceGjkKernel(BatchView b, S tol, int max_iter, DistanceOut out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SlotStore<S> st;
  st.base = reinterpret_cast<S*>(smem_raw) + threadIdx.x;
  st.stride = blockDim.x;
  const ShapeD<S>* __restrict__ shapes = static_cast<const ShapeD<S>*>(b.shapes);
  const ConvexD<S>* __restrict__ cvx = static_cast<const ConvexD<S>*>(b.convex);
  const S* __restrict__ poses1 = static_cast<const S*>(b.poses1);
  const S* __restrict__ poses2 = static_cast<const S*>(b.poses2);
  const S tol_sq = tol * tol;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;

  int phase = PH_FETCH;
  size_t q = 0;
  MinkDiff<S, T0, T1> md;
  Simp simplex;
  simplex.ord = 0;
  simplex.rank = -1;
  V3<S> d = mk<S>(S(-1), S(0), S(0));
  V3<S> cur = zero3<S>();
  S book = S(0);
  int it = 0;
  ExtractPlan<S> plan;
  plan.n = 0;
  int ex_k = 0;
  V3<S> p0 = zero3<S>(), p1 = zero3<S>();

  while (true) {
    if (phase == PH_FETCH) {
      if (i >= b.count) break;
      q = b.perm ? size_t(b.perm[b.begin + i]) : (b.begin + i);
      i += stride;
      const fclb_pair pr = b.pairs[q];
      md.s0 = bindShape(shapes, cvx, pr.shape1);
      md.s1 = bindShape(shapes, cvx, pr.shape2);
      md.setPoses(loadPose(poses1, q), loadPose(poses2, q));
      // guess = (1,0,0); Evaluate is called with -guess (gjk_solver-inl.h:768,783)
      d = normalized(mk<S>(S(-1), S(0), S(0)));
      simplex.ord = 0;
      simplex.rank = -1;
      phase = PH_BOOL_FIRST;
    }

    // ---- the single support site ----
    const V3<S> s0 = md.support0(d);
    const V3<S> s1 = md.support1(-d);

    bool done = false;        // query finished this trip
    bool separated = false;   // result flag when done
    bool valid = false;       // witness validity when done && separated
    bool begin_extract = false;

    if (phase == PH_EXTRACT) {
      if (!plan.weighted) {
        p0 = s0;
        p1 = s1;
      } else {
        const S w = (ex_k == 0) ? plan.w0 : ((ex_k == 1) ? plan.w1 : plan.w2);
        if (ex_k == 0) {
          p0 = s0 * w;
          p1 = s1 * w;
        } else {
          p0 = p0 + s0 * w;
          p1 = p1 + s1 * w;
        }
      }
      ex_k += 1;
      if (ex_k >= plan.n) {
        done = true;
        separated = true;
        valid = plan.valid;
      } else {
        d = st.dir((plan.slots >> (8 * ex_k)) & 0xff);
      }
    } else {
      const V3<S> v = s0 - s1;
      bool to_dist = false;
      if (phase == PH_BOOL_FIRST) {  // gjk.hpp:73-88
        addVertex(st, simplex, v, d);
        if (sqnorm(v) <= tol_sq) {
          done = true;  // Intersect
        } else if (dot(v, d) < 0) {
          to_dist = true;
        } else {
          d = d * S(-1);
          it = 0;
          phase = PH_BOOL;
          if (it >= max_iter) done = true;  // IterationLimit
        }
      } else if (phase == PH_BOOL) {  // gjk.hpp:92-141
        it += 1;
        if (dot(v, d) < 0) {
          to_dist = true;
        } else {
          bool dup = false;
          #pragma unroll 1
          for (int j = 0; j < simplex.rank; j++)
            if (sqnorm(st.vtx(slotOf(simplex, j)) - v) < tol_sq) dup = true;
          if (dup || sqnorm(v) <= tol_sq) {
            done = true;  // ConvergeNoProgress / Intersect: not separated either way
          } else {
            addVertex(st, simplex, v, d);
            const int ps = simplexProjection(st, simplex, d, tol);
            if (ps != PROJ_CONTINUE || it >= max_iter) done = true;
          }
        }
      } else {  // PH_DIST: gjk_distance.hpp:48-101, v is the new vertex along d
        it += 1;
        const S delta = dot(d, v - cur);
        bool stop = delta < tol;
        if (!stop) {
          #pragma unroll 1
          for (int j = 0; j < simplex.rank; j++)
            if (sqnorm(st.vtx(slotOf(simplex, j)) - v) < tol_sq) stop = true;
        }
        if (stop) {
          begin_extract = true;
        } else {
          addVertex(st, simplex, v, d);
          const int us = minDistUpdate(st, simplex, cur, tol);
          if (us == 0) {
            begin_extract = true;
          } else if (us == 1) {
            const S nd = norm(cur);
            const S improvement = book - nd;
            if (improvement < tol || nd < tol) {
              begin_extract = true;
            } else {
              book = nd;
              d = (-cur) / book;
              if (it >= max_iter) {  // loop falls through: "return false" (:104-105)
                done = true;
                separated = true;
                valid = false;
                p0 = zero3<S>();
                p1 = zero3<S>();
              }
            }
          } else {
            done = true;
            separated = true;
            valid = false;
            p0 = zero3<S>();
            p1 = zero3<S>();
          }
        }
      }
      if (to_dist) {
        // process_separated_vertex (gjk.hpp:22-52) + the distance loop's prologue
        // (gjk_distance.hpp:14-46): the simplex restarts from the certifying vertex.
        simplex.rank = -1;
        simplex.ord = 0;
        addVertex(st, simplex, v, d);
        cur = v;
        book = norm(cur);
        if (book <= tol) {
          begin_extract = true;
        } else {
          d = (-cur) / book;
          it = 0;
          phase = PH_DIST;
          if (it >= max_iter) {
            done = true;
            separated = true;
            valid = false;
            p0 = zero3<S>();
            p1 = zero3<S>();
          }
        }
      }
      if (begin_extract) {
        plan = planExtraction(st, simplex);
        if (plan.n == 0) {
          done = true;
          separated = true;
          valid = false;
          p0 = zero3<S>();
          p1 = zero3<S>();
        } else {
          ex_k = 0;
          d = st.dir(plan.slots & 0xff);
          phase = PH_EXTRACT;
        }
      }
    }

    if (done) {
      S dist = S(-1);
      V3<S> w1 = zero3<S>(), w2 = zero3<S>();
      uint8_t ok = 0;
      if (separated) {
        // answers are in shape-1's frame; report them in the world frame
        // (gjk_solver-inl.h:785-792)
        const Pose<S> tf1 = loadPose(poses1, q);
        dist = norm(p0 - p1);
        w1 = apply(tf1, p0);
        w2 = apply(tf1, p1);
        ok = valid ? uint8_t(1) : uint8_t(3);
      }
      writeDistance(out, q, dist, w1, w2, ok);
      phase = PH_FETCH;
    }
  }
}

// ---- GJK bucket, phase-binned scheduling -------------------------------------------
// The state machine above keeps one query per lane, so the lanes of a warp sit in five different
// phases (ncu: 5 of 32 lanes active per issued instruction).  Here a warp owns a POOL of query slots in
// shared memory (simplex store, direction, closest point, Minkowski-difference transforms, phase) and in
// every trip it picks the phase most of its slots are in, hands one such slot to each lane, and runs
// the support + update of THAT phase only: the per-trip control flow is warp-uniform, what is left is
// the divergence inside an update (simplex rank).  Queries are fetched in contiguous ranges from an
// atomic cursor, so the pose reads are coalesced.  Per-query arithmetic and decisions are those of
// the kernel above (same helpers), only the order in which queries advance differs.
template <typename S>
struct BinPool {
  // query slots per warp.  Measured on C2 (B200, capsule/cylinder-box kernels, ms per 3.3M queries):
  //   f32: 96 slots 3.6 | 80: 2.95 | 64: 2.70 | 60: 2.54 | 48: 2.70 | 40: 2.93   (60 slots = 53.8 KB per CTA = 4 CTAs per SM)
  //   f64: 48 slots 8.5 | 36: 8.2 | 30: 7.6
  // -- a fuller pool fills the chosen bin better, a smaller one buys occupancy; occupancy wins beyond ~60.
#ifndef FCLB_GJK_SLOTS
#define FCLB_GJK_SLOTS 60
#endif
  static constexpr int kSlots = sizeof(S) == 4 ? FCLB_GJK_SLOTS : FCLB_GJK_SLOTS / 2;
  // simplex 24 | d 3 | {cur 3, book 1} aliased with {p0 3, p1 3, w 3} (distance loop vs witness extraction) 9 |
  // toshape0 12 (toshape1 = transpose(toshape0.R) bit for bit: same products, same summation order)
  static constexpr int kWordsS = 24 + 3 + 9 + 12;
  static constexpr int kWordsU = 8;                        // q phase ord rankit plan pslots pair1 pair2
  static constexpr size_t bytesPerWarp() { return size_t(kSlots) * (kWordsS * sizeof(S) + kWordsU * 4); }
};

// kCta: the four warps of a CTA share ONE pool (4 x kSlots slots) and the bins are (phase, simplex rank): with 240 slots
// the nine bins fill to about a warp each, so a trip runs one projection code (rank 2, 3 or 4) on full warps.  Every
// trip the warps publish the bin masks of their slice, and each warp derives the same list of work units (bin, 32
// consecutive members) from them and takes the unit with its own index -- two CTA barriers per trip, no atomics.
#ifndef FCLB_GJK_CTA_WARPS
#define FCLB_GJK_CTA_WARPS 4
#endif
template <typename S, int T0, int T1, int kCtaWarps>
__global__ void __launch_bounds__(kCtaWarps ? kCtaWarps * 32 : kBlock)
    distanceGjkBinnedKernel(BatchView b, S tol, int max_iter, DistanceOut out, unsigned long long* cursor) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool kCta = kCtaWarps != 0;
  constexpr int kWarps = kCta ? kCtaWarps : kBlock / 32;
  constexpr int kPool = BinPool<S>::kSlots * (kBlock / 32);   // slots per CTA (the same shared memory in both modes)
  constexpr int NW = kCta ? kPool / kWarps : BinPool<S>::kSlots;  // slots per warp slice
  constexpr int NS = kCta ? NW * kWarps : NW;            // slots addressed by one warp
  constexpr int W2 = (NW + 31) / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wbase = kCta ? smem_raw : smem_raw + size_t(warp) * BinPool<S>::bytesPerWarp();
  S* fS = reinterpret_cast<S*>(wbase);                                        // [kWordsS][NS]
  uint32_t* fU = reinterpret_cast<uint32_t*>(wbase + size_t(NS) * BinPool<S>::kWordsS * sizeof(S));  // [kWordsU][NS]
  const ShapeD<S>* __restrict__ shapes = static_cast<const ShapeD<S>*>(b.shapes);
  const ConvexD<S>* __restrict__ cvx = static_cast<const ConvexD<S>*>(b.convex);
  const S* __restrict__ poses1 = static_cast<const S*>(b.poses1);
  const S* __restrict__ poses2 = static_cast<const S*>(b.poses2);
  const S tol_sq = tol * tol;
  enum { U_Q = 0, U_PHASE, U_ORD, U_RANKIT, U_PLAN, U_PSLOTS, U_PAIR1, U_PAIR2 };
  enum { F_D = 24, F_CUR = 27, F_BOOK = 30, F_P0 = 27, F_P1 = 30, F_W = 33, F_TS0R = 36, F_TS0T = 45 };
  auto ldS = [&](int f, int slot) { return fS[f * NS + slot]; };
  auto stS = [&](int f, int slot, S v) { fS[f * NS + slot] = v; };
  auto ld3 = [&](int f, int slot) { return mk<S>(fS[f * NS + slot], fS[(f + 1) * NS + slot], fS[(f + 2) * NS + slot]); };
  auto st3 = [&](int f, int slot, const V3<S>& v) {
    fS[f * NS + slot] = v.x;
    fS[(f + 1) * NS + slot] = v.y;
    fS[(f + 2) * NS + slot] = v.z;
  };
  // scheduling bins: the phase, and for the two iterating phases also the simplex rank the update starts
  // from (the projection / sub-simplex search of rank 2, 3 and 4 simplices are three different codes)
  // (measured: with 64 slots per warp the nine bins fill too thinly -- 3.6 ms against 2.7 ms for phase-only
  // bins on C2 -- so the rank split is compiled out; the census below skips the unused bins)
  constexpr bool kRankBins = kCta;
#ifndef FCLB_GJK_ASYNC
#define FCLB_GJK_ASYNC 0
#endif
  constexpr bool kAsync = FCLB_GJK_ASYNC != 0;
  constexpr int kBins = 9;  // 0 FETCH, 1 BOOL_FIRST, 2..4 BOOL rank 1..3, 5..7 DIST rank 1..3, 8 EXTRACT
  auto binOf = [](int phase, int rank) {
    const int r = !kRankBins ? 1 : (rank < 1 ? 1 : (rank > 3 ? 3 : rank));  // (the phase must survive whatever the rank is)
    return phase == PH_BOOL ? 1 + r : (phase == PH_DIST ? 4 + r : (phase == PH_EXTRACT ? 8 : phase));
  };
  __shared__ unsigned s_mask[kCta ? kWarps : 1][kBins][W2];
  __shared__ int s_more, s_next;
#ifndef FCLB_GJK_MIN_FILL
#define FCLB_GJK_MIN_FILL 1
#endif
#ifndef FCLB_GJK_SHARE_UNITS
#define FCLB_GJK_SHARE_UNITS 0
#endif
  constexpr bool kShareUnits = FCLB_GJK_SHARE_UNITS != 0;
  constexpr int kMinFill = FCLB_GJK_MIN_FILL;
  bool trip_open = false;  // (CTA pool) inside a trip: the unit list below is valid
  int u_nfull = 0, u_rem = 0, u_incl = 0, u_total_full = 0, u_rank = 0, u_count = 0;
  if (kCta) {
    #pragma unroll 1
    for (int k = threadIdx.x; k < NS; k += kWarps * 32) fU[U_PHASE * NS + k] = 0;
    if (threadIdx.x == 0) s_more = 1;
  } else {
    #pragma unroll 1
    for (int k = lane; k < NS; k += 32) fU[U_PHASE * NS + k] = 0;
    __syncwarp();
  }
  bool more = true;  // queries left behind the cursor

  if (kCta && kAsync) __syncthreads();

  while (true) {
    int best_bin = -1, best_n = 0, slot = -1;
    if (kCta && kAsync) {
      // ---- no barriers: a warp reads all slot states, takes the fullest bin and CLAIMS its members with a
      // compare-and-swap on the state word (another warp may have been faster; that lane then idles this trip)
      constexpr int J = (NS + 31) / 32;
      volatile uint32_t* vstate = fU + U_PHASE * NS;
      uint32_t stv[J];
      unsigned long long acc = 0;
      unsigned acc8 = 0;
#pragma unroll
      for (int j = 0; j < J; j++) {
        const int k = j * 32 + lane;
        stv[j] = k < NS ? vstate[k] : 0xffu;
        if (stv[j] < 8u) acc += 1ull << (8 * stv[j]);
        else if (stv[j] == 8u) acc8 += 1;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc8 += __shfl_xor_sync(0xffffffffu, acc8, o);
      }
      more = *reinterpret_cast<volatile int*>(&s_more) != 0;
#pragma unroll
      for (int p = kBins - 1; p >= 0; p--) {
        if (p == 0 && !more) continue;
        const int c = p == 8 ? int(acc8) : int((acc >> (8 * p)) & 0xffu);
        if (c > best_n) {
          best_n = c;
          best_bin = p;
        }
      }
      if (best_bin < 0) break;  // nothing unclaimed and nothing to fetch: the warps holding claimed slots finish them
      const int take = best_n < 32 ? best_n : 32;
      // members in slot order, this warp starts a quarter (1/kWarps) of the way round so that two warps on the same bin collide less
      int want = lane < take ? (lane + (warp * best_n) / kWarps) % best_n : -1;
      unsigned sel = 0;
      int sel_base = 0, sel_want = -1;
#pragma unroll
      for (int j = 0; j < J; j++) {
        const unsigned m = __ballot_sync(0xffffffffu, stv[j] == uint32_t(best_bin));
        const int c = __popc(m);
        if (sel_want < 0 && want >= 0 && want < c) {
          sel = m;
          sel_base = j * 32;
          sel_want = want;
        }
        want -= c;
      }
      if (sel_want >= 0) {
        slot = sel_base + int(__fns(sel, 0, sel_want + 1));
        if (atomicCAS(const_cast<uint32_t*>(&vstate[slot]), uint32_t(best_bin), 0xffu) != uint32_t(best_bin)) slot = -1;
      }
      __threadfence_block();
      best_n = take;
    } else if (kCta) {
      if (!trip_open) {
        __syncthreads();  // the slot states the other warps wrote in the previous trip
        more = *reinterpret_cast<volatile int*>(&s_more) != 0;  // (read between the barriers: fetching warps clear it after the second)
        if (threadIdx.x == 0) s_next = 0;
#pragma unroll
        for (int w = 0; w < W2; w++) {
          const int k = w * 32 + lane;
          const int ph = k < NW ? int(fU[U_PHASE * NS + warp * NW + k]) : -1;
#pragma unroll
          for (int p = 0; p < kBins; p++) {
            const unsigned m = __ballot_sync(0xffffffffu, ph == p);
            if (lane == p) s_mask[warp][p][w] = m;
          }
        }
        __syncthreads();
        // lane l < kBins speaks for bin kBins-1-l.  The trip's work units, the same list in every warp: the full ones
        // (32 members) from the last bin down, then the partial ones, largest first, as long as they fill
        // kMinFill lanes (a thinner bin waits for more members unless nothing else is left)
        const int mybin = kBins - 1 - lane;
        int T = 0;
        if (lane < kBins && !(mybin == 0 && !more)) {
#pragma unroll
          for (int w4 = 0; w4 < kWarps; w4++)
#pragma unroll
            for (int w = 0; w < W2; w++) T += __popc(s_mask[w4][mybin][w]);
        }
        if (__ballot_sync(0xffffffffu, T > 0) == 0) break;  // pool empty and nothing left to fetch (the same in every warp)
        u_nfull = T >> 5;
        u_rem = T & 31;
        int incl = u_nfull;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        u_incl = incl;
        u_total_full = __shfl_sync(0xffffffffu, incl, kBins - 1);
        u_rank = -1;  // (ranked below, only by a warp that draws a partial unit)
        const int n_fill = __popc(__ballot_sync(0xffffffffu, lane < kBins && u_rem >= kMinFill));
        u_count = u_total_full + (n_fill > 0 ? n_fill : (u_total_full == 0 ? 1 : 0));
        trip_open = true;
      }
      // ---- this warp's unit of the trip.  FCLB_GJK_SHARE_UNITS=1 lets the warps draw ALL the trip's units from a counter
      // instead (measured on C2 f32: 3.2 ms per launch against 2.2 ms -- working through the thin bins costs more
      // than the barrier wait it saves), so by default a trip runs its kWarps best units and the rest stay for later
      int u = warp;
      if (kShareUnits) {
        if (lane == 0) u = atomicAdd(&s_next, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
      } else {
        trip_open = false;
      }
      // FCLB_GJK_SHARE_UNITS=2: as 1, but only FULL units are drawn beyond the first kWarps of the list
      const int u_limit = (FCLB_GJK_SHARE_UNITS == 2 && u_count > kWarps) ? (u_total_full > kWarps ? u_total_full : kWarps) : u_count;
      if (u >= u_limit) {
        trip_open = false;
        continue;
      }
      int start = 0;
      if (u < u_total_full) {
        const unsigned owner = __ballot_sync(0xffffffffu, lane < kBins && u >= u_incl - u_nfull && u < u_incl);
        const int l = __ffs(owner) - 1;
        best_bin = kBins - 1 - l;
        start = 32 * (u - __shfl_sync(0xffffffffu, u_incl - u_nfull, l));
        best_n = 32;
      } else {
        if (u_rank < 0) {
          u_rank = 0;
#pragma unroll
          for (int l2 = 0; l2 < kBins; l2++) {
            const int o = __shfl_sync(0xffffffffu, u_rem, l2);
            if (o > u_rem || (o == u_rem && l2 < lane)) u_rank += 1;
          }
        }
        const unsigned owner = __ballot_sync(0xffffffffu, lane < kBins && u_rem > 0 && u_rank == u - u_total_full);
        if (owner) {
          const int l = __ffs(owner) - 1;
          best_bin = kBins - 1 - l;
          start = 32 * __shfl_sync(0xffffffffu, u_nfull, l);
          best_n = __shfl_sync(0xffffffffu, u_rem, l);
        }
      }
      if (best_bin >= 0) {
        int want = lane < best_n ? start + lane : -1;
        unsigned sel = 0;
        int sel_base = 0, sel_want = -1;
#pragma unroll
        for (int w4 = 0; w4 < kWarps; w4++)
#pragma unroll
          for (int w = 0; w < W2; w++) {
            const unsigned m = s_mask[w4][best_bin][w];
            const int c = __popc(m);
            if (sel_want < 0 && want >= 0 && want < c) {
              sel = m;
              sel_base = w4 * NW + w * 32;
              sel_want = want;
            }
            want -= c;
          }
        if (sel_want >= 0) slot = sel_base + int(__fns(sel, 0, sel_want + 1));
      }
    } else {
    // ---- census of the pool: which bin do most slots wait in?
    unsigned masks[kBins][(NS + 31) / 32];
    int cnt[kBins];
#pragma unroll
    for (int p = 0; p < kBins; p++) cnt[p] = 0;
#pragma unroll
    for (int w = 0; w < (NS + 31) / 32; w++) {
      const int k = w * 32 + lane;
      const int ph = k < NS ? int(fU[U_PHASE * NS + k]) : -1;
#pragma unroll
      for (int p = 0; p < kBins; p++) {
        if (!kRankBins && (p == 3 || p == 4 || p == 6 || p == 7)) {
          masks[p][w] = 0;
          continue;
        }
        masks[p][w] = __ballot_sync(0xffffffffu, ph == p);
        cnt[p] += __popc(masks[p][w]);
      }
    }
#pragma unroll
    for (int p = kBins - 1; p >= 0; p--) {  // ties go to the later bin (drains the pool)
      if (p == 0 && !more) continue;
      if (cnt[p] > best_n) {
        best_n = cnt[p];
        best_bin = p;
      }
    }
    if (best_bin < 0) break;  // pool empty and nothing left to fetch
    // ---- hand one slot of that phase to each lane
    {
      int want = lane;
#pragma unroll
      for (int w = 0; w < (NS + 31) / 32; w++) {
        const int c = __popc(masks[best_bin][w]);
        if (slot < 0 && want >= 0 && want < c) slot = w * 32 + int(__fns(masks[best_bin][w], 0, want + 1));
        want -= c;
      }
    }
    }
    const int best = best_bin <= 0 ? PH_FETCH
                                   : (best_bin == 1 ? PH_BOOL_FIRST : (best_bin <= 4 ? PH_BOOL : (best_bin <= 7 ? PH_DIST : PH_EXTRACT)));
    if (kCta && best_bin < 0) continue;  // no unit for this warp in this trip
    int n_act = best_n < 32 ? best_n : 32;

    if (best == PH_FETCH) {
      int idx = lane;  // which of the fetched queries this lane takes
      if (kCta && kAsync) {
        const unsigned got = __ballot_sync(0xffffffffu, slot >= 0);
        n_act = __popc(got);
        idx = slot >= 0 ? __popc(got & ((1u << lane) - 1u)) : 32;
        if (n_act == 0) continue;
      }
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(cursor, (unsigned long long)n_act);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + n_act >= b.count) {
        more = false;
        if (kCta && lane == 0) *reinterpret_cast<volatile int*>(&s_more) = 0;
      }
      if (kCta && kAsync && slot >= 0 && base + idx >= b.count) fU[U_PHASE * NS + slot] = 0;  // claimed for nothing
      if (slot >= 0 && idx < n_act && base + idx < b.count) {
        const size_t i = size_t(base) + idx;
        const size_t q = b.perm ? size_t(b.perm[b.begin + i]) : (b.begin + i);
        const fclb_pair pr = b.pairs[q];
        MinkDiff<S, T0, T1> md;
        md.setPoses(loadPose(poses1, q), loadPose(poses2, q));
#pragma unroll
        for (int k = 0; k < 9; k++) stS(F_TS0R + k, slot, md.toshape0.R.m[k]);
        st3(F_TS0T, slot, md.toshape0.t);
        st3(F_D, slot, normalized(mk<S>(S(-1), S(0), S(0))));  // Evaluate(-guess), guess = (1,0,0)
        fU[U_Q * NS + slot] = uint32_t(q);
        fU[U_PAIR1 * NS + slot] = pr.shape1;
        fU[U_PAIR2 * NS + slot] = pr.shape2;
        fU[U_ORD * NS + slot] = 0;
        fU[U_RANKIT * NS + slot] = 0;  // rank + 1 = 0 (rank -1), it = 0
        if (kCta && kAsync) __threadfence_block();  // the slot's fields before its state
        *reinterpret_cast<volatile uint32_t*>(&fU[U_PHASE * NS + slot]) = 1;
      }
      if (!kCta) __syncwarp();
      continue;
    }

    if (slot >= 0 && lane < n_act) {
      // ---- load the slot
      SlotStore<S> st;
      st.base = fS + slot;
      st.stride = NS;
      MinkDiff<S, T0, T1> md;
      md.s0 = bindShape(shapes, cvx, fU[U_PAIR1 * NS + slot]);
      md.s1 = bindShape(shapes, cvx, fU[U_PAIR2 * NS + slot]);
#pragma unroll
      for (int k = 0; k < 9; k++) md.toshape0.R.m[k] = ldS(F_TS0R + k, slot);
      md.toshape1 = transpose(md.toshape0.R);
      md.toshape0.t = ld3(F_TS0T, slot);
      V3<S> d = ld3(F_D, slot);
      Simp simplex;
      simplex.ord = fU[U_ORD * NS + slot];
      const uint32_t rankit = fU[U_RANKIT * NS + slot];
      simplex.rank = int(rankit & 0xffu) - 1;
      int it = int(rankit >> 8);
      int phase = best;

      // ---- the single support site
      const V3<S> s0 = md.support0(d);
      const V3<S> s1 = md.support1(-d);

      bool done = false, separated = false, valid = false, begin_extract = false;
      V3<S> p0 = zero3<S>(), p1 = zero3<S>();
      V3<S> cur = zero3<S>();
      S book = S(0);

      if (best == PH_EXTRACT) {
        const uint32_t pl = fU[U_PLAN * NS + slot];
        const int plan_n = int(pl & 0xfu);
        const bool weighted = (pl >> 4) & 1u, pvalid = (pl >> 5) & 1u;
        int ex_k = int(pl >> 8);
        const uint32_t pslots = fU[U_PSLOTS * NS + slot];
        if (!weighted) {
          p0 = s0;
          p1 = s1;
        } else {
          const S w = ldS(F_W + ex_k, slot);
          if (ex_k == 0) {
            p0 = s0 * w;
            p1 = s1 * w;
          } else {
            p0 = ld3(F_P0, slot) + s0 * w;
            p1 = ld3(F_P1, slot) + s1 * w;
          }
        }
        ex_k += 1;
        if (ex_k >= plan_n) {
          done = true;
          separated = true;
          valid = pvalid;
        } else {
          st3(F_P0, slot, p0);
          st3(F_P1, slot, p1);
          st3(F_D, slot, st.dir((pslots >> (8 * ex_k)) & 0xff));
          fU[U_PLAN * NS + slot] = (pl & 0xffu) | (uint32_t(ex_k) << 8);
          if (kCta && kAsync) {  // back to the extraction bin (the claim had overwritten the state)
            __threadfence_block();
            *reinterpret_cast<volatile uint32_t*>(&fU[U_PHASE * NS + slot]) = 8u;
          }
        }
      } else {
        const V3<S> v = s0 - s1;
        bool to_dist = false;
        if (best == PH_BOOL_FIRST) {  // gjk.hpp:73-88
          addVertex(st, simplex, v, d);
          if (sqnorm(v) <= tol_sq) {
            done = true;
          } else if (dot(v, d) < 0) {
            to_dist = true;
          } else {
            d = d * S(-1);
            it = 0;
            phase = PH_BOOL;
            if (it >= max_iter) done = true;
          }
        } else if (best == PH_BOOL) {  // gjk.hpp:92-141
          it += 1;
          if (dot(v, d) < 0) {
            to_dist = true;
          } else {
            bool dup = false;
            #pragma unroll 1
            for (int j = 0; j < simplex.rank; j++)
              if (sqnorm(st.vtx(slotOf(simplex, j)) - v) < tol_sq) dup = true;
            if (dup || sqnorm(v) <= tol_sq) {
              done = true;
            } else {
              addVertex(st, simplex, v, d);
              const int ps = simplexProjection(st, simplex, d, tol);
              if (ps != PROJ_CONTINUE || it >= max_iter) done = true;
            }
          }
        } else {  // PH_DIST: gjk_distance.hpp:48-101
          cur = ld3(F_CUR, slot);
          book = ldS(F_BOOK, slot);
          it += 1;
          const S delta = dot(d, v - cur);
          bool stop = delta < tol;
          if (!stop) {
            #pragma unroll 1
            for (int j = 0; j < simplex.rank; j++)
              if (sqnorm(st.vtx(slotOf(simplex, j)) - v) < tol_sq) stop = true;
          }
          if (stop) {
            begin_extract = true;
          } else {
            addVertex(st, simplex, v, d);
            const int us = minDistUpdate(st, simplex, cur, tol);
            if (us == 0) {
              begin_extract = true;
            } else if (us == 1) {
              const S nd = norm(cur);
              const S improvement = book - nd;
              if (improvement < tol || nd < tol) {
                begin_extract = true;
              } else {
                book = nd;
                d = (-cur) / book;
                if (it >= max_iter) {
                  done = true;
                  separated = true;
                  valid = false;
                }
              }
            } else {
              done = true;
              separated = true;
              valid = false;
            }
          }
        }
        if (to_dist) {  // process_separated_vertex (gjk.hpp:22-52) + distance prologue (gjk_distance.hpp:14-46)
          simplex.rank = -1;
          simplex.ord = 0;
          addVertex(st, simplex, v, d);
          cur = v;
          book = norm(cur);
          if (book <= tol) {
            begin_extract = true;
          } else {
            d = (-cur) / book;
            it = 0;
            phase = PH_DIST;
            if (it >= max_iter) {
              done = true;
              separated = true;
              valid = false;
            }
          }
        }
        if (begin_extract) {
          const ExtractPlan<S> plan = planExtraction(st, simplex);
          if (plan.n == 0) {
            done = true;
            separated = true;
            valid = false;
          } else {
            d = st.dir(plan.slots & 0xff);
            phase = PH_EXTRACT;
            fU[U_PLAN * NS + slot] = uint32_t(plan.n) | (plan.weighted ? 16u : 0u) | (plan.valid ? 32u : 0u);
            fU[U_PSLOTS * NS + slot] = plan.slots;
            stS(F_W, slot, plan.w0);
            stS(F_W + 1, slot, plan.w1);
            stS(F_W + 2, slot, plan.w2);
          }
        }
        if (!done) {
          st3(F_D, slot, d);
          if (phase == PH_DIST) {
            st3(F_CUR, slot, cur);
            stS(F_BOOK, slot, book);
          }
          fU[U_ORD * NS + slot] = simplex.ord;
          fU[U_RANKIT * NS + slot] = uint32_t(simplex.rank + 1) | (uint32_t(it) << 8);
          if (kCta && kAsync) __threadfence_block();
          *reinterpret_cast<volatile uint32_t*>(&fU[U_PHASE * NS + slot]) = uint32_t(binOf(phase, simplex.rank));
        }
      }

      if (done) {
        const size_t q = fU[U_Q * NS + slot];
        S dist = S(-1);
        V3<S> w1 = zero3<S>(), w2 = zero3<S>();
        uint8_t ok = 0;
        if (separated) {  // (p0, p1 stay zero on the paths where the reference returns without witness points)
          const Pose<S> tf1 = loadPose(poses1, q);
          dist = norm(p0 - p1);
          w1 = apply(tf1, p0);
          w2 = apply(tf1, p1);
          ok = valid ? uint8_t(1) : uint8_t(3);
        }
        writeDistance(out, q, dist, w1, w2, ok);
        if (kCta && kAsync) __threadfence_block();
        *reinterpret_cast<volatile uint32_t*>(&fU[U_PHASE * NS + slot]) = 0;
      }
    }
    if (!kCta) __syncwarp();
  }
}

inline int gridFor(size_t count, int block, int ctas_per_sm) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t need = (count + block - 1) / block;
  const size_t cap = size_t(sms) * ctas_per_sm;
  return int(need < cap ? (need ? need : 1) : cap);
}

// FCLB_GJK_BINNED=0 selects the one-query-per-lane state machine (kept for comparison)
inline int gjkBinnedMode() {
  static int v = [] {
    const char* e = getenv("FCLB_GJK_BINNED");
    return e ? atoi(e) : 2;
  }();
  return v;  // 0 one query per lane | 1 per-warp pool, phase bins | 2 CTA-wide pool, (phase, rank) bins
}
inline bool gjkBinnedEnabled() { return gjkBinnedMode() != 0; }
unsigned long long* gjkCursor();  // device counter (fclb_engine.cu)

template <typename S, int T0, int T1>
cudaError_t launchGjkDistance(const BatchView& b, const SolverParams& sp, const DistanceOut& out, cudaStream_t st) {
  if (gjkBinnedEnabled()) {
    unsigned long long* cursor = gjkCursor();
    if (!cursor) return cudaErrorMemoryAllocation;
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    const size_t smem = BinPool<S>::bytesPerWarp() * (kBlock / 32);
    const bool cta = gjkBinnedMode() == 2;
    auto kern = cta ? distanceGjkBinnedKernel<S, T0, T1, FCLB_GJK_CTA_WARPS> : distanceGjkBinnedKernel<S, T0, T1, 0>;
    const int block = cta ? FCLB_GJK_CTA_WARPS * 32 : kBlock;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int per_sm = int((227 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    const size_t need = (b.count + BinPool<S>::kSlots * (kBlock / 32) - 1) / (BinPool<S>::kSlots * (kBlock / 32));
    int grid = gridFor(b.count, kBlock, per_sm);
    if (size_t(grid) > need) grid = int(need ? need : 1);
    kern<<<grid, block, smem, st>>>(b, S(sp.gjk_tol), sp.gjk_max_iter, out, cursor);
    return cudaGetLastError();
  }
  const size_t smem = size_t(24) * sizeof(S) * kBlock;
  const int grid = gridFor(b.count, kBlock, 16);
  distanceGjkKernel<S, T0, T1><<<grid, kBlock, smem, st>>>(b, S(sp.gjk_tol), sp.gjk_max_iter, out);
  return cudaGetLastError();
}

template <typename S, int CK>
cudaError_t launchClosedDistance(const BatchView& b, const SolverParams& sp, const DistanceOut& out, cudaStream_t st) {
  const int grid = gridFor(b.count, kBlock, 16);
  distanceClosedKernel<S, CK><<<grid, kBlock, 0, st>>>(b, S(sp.eps78), out);
  return cudaGetLastError();
}

template <typename S>
cudaError_t launchDistance(const BatchView& b, const SolverParams& sp, const DistanceOut& out, cudaStream_t st,
                           int* n_launches) {
  if (b.count == 0) return cudaSuccess;
  if (n_launches) *n_launches += 1;
  switch (sp.generic_only ? CK_NONE : closedKindOf(b.type1, b.type2)) {
    case CK_SPHERE_BOX:
      return launchClosedDistance<S, CK_SPHERE_BOX>(b, sp, out, st);
    case CK_BOX_SPHERE:
      return launchClosedDistance<S, CK_BOX_SPHERE>(b, sp, out, st);
    case CK_SPHERE_CAPSULE:
      return launchClosedDistance<S, CK_SPHERE_CAPSULE>(b, sp, out, st);
    case CK_CAPSULE_SPHERE:
      return launchClosedDistance<S, CK_CAPSULE_SPHERE>(b, sp, out, st);
    case CK_SPHERE_CYLINDER:
      return launchClosedDistance<S, CK_SPHERE_CYLINDER>(b, sp, out, st);
    case CK_CYLINDER_SPHERE:
      return launchClosedDistance<S, CK_CYLINDER_SPHERE>(b, sp, out, st);
    case CK_SPHERE_SPHERE:
      return launchClosedDistance<S, CK_SPHERE_SPHERE>(b, sp, out, st);
    case CK_CAPSULE_CAPSULE:
      return launchClosedDistance<S, CK_CAPSULE_CAPSULE>(b, sp, out, st);
    default:
      break;
  }
  // Shape-specialised GJK for the pair kinds the configs exercise; everything
  // else goes through the run-time switch (same arithmetic, more divergence).
#define FCLB_GJK_CASE(A, B) \
  if (b.type1 == A && b.type2 == B) return launchGjkDistance<S, A, B>(b, sp, out, st);
  FCLB_GJK_CASE(ST_CAPSULE, ST_BOX)
  FCLB_GJK_CASE(ST_CYLINDER, ST_BOX)
  FCLB_GJK_CASE(ST_BOX, ST_BOX)
  FCLB_GJK_CASE(ST_BOX, ST_CAPSULE)
  FCLB_GJK_CASE(ST_BOX, ST_CYLINDER)
  FCLB_GJK_CASE(ST_CONVEX, ST_CONVEX)
  FCLB_GJK_CASE(ST_CONVEX, ST_BOX)
  FCLB_GJK_CASE(ST_BOX, ST_CONVEX)
#undef FCLB_GJK_CASE
  return launchGjkDistance<S, ST_DYNAMIC, ST_DYNAMIC>(b, sp, out, st);
}

}  // namespace fclb
