// fclb_collide_launch.cuh -- per-bucket launch logic of the collide path
// (included by the per-scalar-type translation units).
#pragma once
#include <cstdlib>

#include "fclb_collide_impl.cuh"
#include "fclb_distance_impl.cuh"  // gridFor

namespace fclb {

template <typename S, int CC>
cudaError_t launchClosedCollide(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st) {
  const int grid = gridFor(b.count, kBlock, 16);
  collideClosedKernel<S, CC><<<grid, kBlock, 0, st>>>(b, a.out);
  return cudaGetLastError();
}

template <typename S, int T0, int T1>
cudaError_t launchConvex(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st, int* n_launches) {
  const size_t smem = size_t(24) * sizeof(S) * kBlock;
  const int grid = gridFor(b.count, kBlock, 8);
  convexBoolKernel<S, T0, T1><<<grid, kBlock, smem, st>>>(b, S(a.sp.gjk_tol), a.sp.gjk_max_iter, a.mode, a.out, a.work);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (a.mode & 2) {
    e = launchEpa<S>(b, a, st);
    if (n_launches) *n_launches += 1;
  }
  return e;
}

template <typename S>
cudaError_t launchMprPenetration(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st) {
  if (b.count == 0) return cudaSuccess;
  const int grid = gridFor(b.count, kBlock, 8);
  const int inc = a.pen_mode == FCLB_PEN_INCREMENTAL_MIN ? 1 : 0;
  const S tol = S(a.sp.epa_tol);  // request.distanceTolerance()
  const S dx = S(a.pen_dir[0]), dy = S(a.pen_dir[1]), dz = S(a.pen_dir[2]);
#define FCLB_PEN_CASE(A, B)                                                                    \
  if (b.type1 == A && b.type2 == B) {                                                          \
    mprPenetrationKernel<S, A, B><<<grid, kBlock, 0, st>>>(b, tol, inc, dx, dy, dz, a.out);    \
    return cudaGetLastError();                                                                 \
  }
  FCLB_PEN_CASE(ST_BOX, ST_BOX)
  FCLB_PEN_CASE(ST_CONVEX, ST_CONVEX)
#undef FCLB_PEN_CASE
  mprPenetrationKernel<S, ST_DYNAMIC, ST_DYNAMIC><<<grid, kBlock, 0, st>>>(b, tol, inc, dx, dy, dz, a.out);
  return cudaGetLastError();
}

template <typename S>
cudaError_t launchCollide(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st, int* n_launches) {
  if (b.count == 0) return cudaSuccess;
  if (n_launches) *n_launches += 1;
  if (!(a.mode & 4)) {  // the gjk_epa API always runs the generic path
    switch (closedCollideOf(b.type1, b.type2)) {
      case CC_SPHERE_SPHERE:
        return launchClosedCollide<S, CC_SPHERE_SPHERE>(b, a, st);
      case CC_SPHERE_CAPSULE:
        return launchClosedCollide<S, CC_SPHERE_CAPSULE>(b, a, st);
      case CC_CAPSULE_SPHERE:
        return launchClosedCollide<S, CC_CAPSULE_SPHERE>(b, a, st);
      case CC_SPHERE_BOX:
        return launchClosedCollide<S, CC_SPHERE_BOX>(b, a, st);
      case CC_BOX_SPHERE:
        return launchClosedCollide<S, CC_BOX_SPHERE>(b, a, st);
      case CC_SPHERE_CYLINDER:
        return launchClosedCollide<S, CC_SPHERE_CYLINDER>(b, a, st);
      case CC_CYLINDER_SPHERE:
        return launchClosedCollide<S, CC_CYLINDER_SPHERE>(b, a, st);
      case CC_SPHERE_TRIANGLE:
        if (b.tris) return launchClosedCollide<S, CC_SPHERE_TRIANGLE>(b, a, st);
        break;
      case CC_BOX_BOX: {
        if (getenv("FCLB_BOXBOX_ONE_PHASE")) return launchClosedCollide<S, CC_BOX_BOX>(b, a, st);
        const int grid = gridFor(b.count, kBlock, 12);
        boxBoxCollideKernel<S><<<grid, kBlock, 0, st>>>(b, a.out);
        return cudaGetLastError();
      }
      default:
        break;
    }
  }
#define FCLB_CVX_CASE(A, B) \
  if (b.type1 == A && b.type2 == B) return launchConvex<S, A, B>(b, a, st, n_launches);
  FCLB_CVX_CASE(ST_BOX, ST_BOX)
  FCLB_CVX_CASE(ST_CONVEX, ST_CONVEX)
#undef FCLB_CVX_CASE
  return launchConvex<S, ST_DYNAMIC, ST_DYNAMIC>(b, a, st, n_launches);
}

}  // namespace fclb
