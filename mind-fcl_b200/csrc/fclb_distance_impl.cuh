// fclb_distance_impl.cuh -- batched fcl::distance (== GJKSolver<S>::shapeDistance,
// reference gjk_solver-inl.h:762-808, closed forms :902-988).
//
// One query per thread; one launch per (type1,type2) bucket so a warp never
// mixes pair kinds.  Closed-form pairs are HBM-bound streaming kernels
// (2 poses in, dist + 2 witness points + flag out); GJK pairs are FP-bound and
// use the shared-memory simplex store of fclb_gjk.cuh.
#pragma once
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

inline int gridFor(size_t count, int block, int ctas_per_sm) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t need = (count + block - 1) / block;
  const size_t cap = size_t(sms) * ctas_per_sm;
  return int(need < cap ? (need ? need : 1) : cap);
}

template <typename S, int T0, int T1>
cudaError_t launchGjkDistance(const BatchView& b, const SolverParams& sp, const DistanceOut& out, cudaStream_t st) {
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
  switch (closedKindOf(b.type1, b.type2)) {
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
