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
template <typename S, int T0, int T1>
__global__ void __launch_bounds__(kBlock) distanceGjkKernel(BatchView b, S tol, int max_iter, DistanceOut out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SlotStore<S> st;
  st.base = reinterpret_cast<S*>(smem_raw) + threadIdx.x;
  st.stride = blockDim.x;
  const ShapeD<S>* __restrict__ shapes = static_cast<const ShapeD<S>*>(b.shapes);
  const ConvexD<S>* __restrict__ cvx = static_cast<const ConvexD<S>*>(b.convex);
  const S* __restrict__ poses1 = static_cast<const S*>(b.poses1);
  const S* __restrict__ poses2 = static_cast<const S*>(b.poses2);
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < b.count; i += size_t(gridDim.x) * blockDim.x) {
    const size_t q = b.perm ? size_t(b.perm[b.begin + i]) : (b.begin + i);
    const fclb_pair pr = b.pairs[q];
    MinkDiff<S, T0, T1> md;
    md.s0 = bindShape(shapes, cvx, pr.shape1);
    md.s1 = bindShape(shapes, cvx, pr.shape2);
    const Pose<S> tf1 = loadPose(poses1, q);
    {
      const Pose<S> tf2 = loadPose(poses2, q);
      md.setPoses(tf1, tf2);
    }
    Simp simplex;
    GjkDistOut<S> dout;
    dout.valid = false;
    dout.p0 = zero3<S>();
    dout.p1 = zero3<S>();
    // guess = (1,0,0); Evaluate is called with -guess (gjk_solver-inl.h:768,783)
    const int status = gjkEvaluate(md, st, simplex, mk<S>(S(-1), S(0), S(0)), tol, max_iter, &dout, nullptr);
    if (status == GJK_SEPARATED) {
      const V3<S> w1 = apply(tf1, dout.p0);
      const V3<S> w2 = apply(tf1, dout.p1);
      const S d = norm(dout.p0 - dout.p1);
      writeDistance(out, q, d, w1, w2, dout.valid ? uint8_t(1) : uint8_t(3));
    } else {
      writeDistance(out, q, S(-1), zero3<S>(), zero3<S>(), uint8_t(0));
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
