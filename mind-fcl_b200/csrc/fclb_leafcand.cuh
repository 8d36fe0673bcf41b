// fclb_leafcand.cuh -- warp-aggregated append to the leaf-candidate sink (fclb_internal.h LeafCandSink).
#pragma once
#include "fclb_internal.h"

namespace fclb {

// Every lane of the warp must call this (converged); lanes with active == true append one candidate.
// box1 / box2: 6 S each (min xyz, max xyz) or nullptr.
template <typename S>
__device__ __forceinline__ void candAppend(const LeafCandSink& c, bool active, uint32_t q, long long b1, long long b2,
                                           const S* box1, const S* box2) {
  const unsigned m = __ballot_sync(0xffffffffu, active);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(c.count, static_cast<unsigned long long>(__popc(m)));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (!active) return;
  const unsigned long long i = base + __popc(m & ((1u << lane) - 1u));
  if (i >= c.cap) return;
  c.q[i] = q;
  c.b1[i] = b1;
  if (c.b2) c.b2[i] = b2;
  if (c.box1 && box1) {
    S* o = static_cast<S*>(c.box1) + i * 6;
#pragma unroll
    for (int k = 0; k < 6; k++) o[k] = box1[k];
  }
  if (c.box2 && box2) {
    S* o = static_cast<S*>(c.box2) + i * 6;
#pragma unroll
    for (int k = 0; k < 6; k++) o[k] = box2[k];
  }
}

}  // namespace fclb
