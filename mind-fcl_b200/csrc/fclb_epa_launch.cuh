// fclb_epa_launch.cuh -- launch of the warp-per-query EPA stage (own translation
// units: the EPA kernels dominate compile time).
#pragma once
#include "fclb_collide_impl.cuh"

namespace fclb {

template <typename S, int T0, int T1>
cudaError_t launchEpaT(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st) {
  const size_t poly = PolyStore<S>::bytes(a.sp.epa_max_faces);
  const size_t per_warp = poly + 24 * sizeof(S) + 16;
  const size_t esmem = per_warp * kEpaWarps;
  auto kern = epaKernel<S, T0, T1>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(esmem));
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = int((227 * 1024) / esmem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  kern<<<sms * per_sm, kEpaWarps * 32, esmem, st>>>(b, S(a.sp.epa_tol), a.sp.epa_max_faces, a.sp.epa_max_iter, a.mode,
                                                    a.out, a.work, poly);
  return cudaGetLastError();
}

template <typename S>
cudaError_t launchEpa(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st) {
  if (b.type1 == ST_BOX && b.type2 == ST_BOX) return launchEpaT<S, ST_BOX, ST_BOX>(b, a, st);
  if (b.type1 == ST_CONVEX && b.type2 == ST_CONVEX) return launchEpaT<S, ST_CONVEX, ST_CONVEX>(b, a, st);
  return launchEpaT<S, ST_DYNAMIC, ST_DYNAMIC>(b, a, st);
}

}  // namespace fclb
