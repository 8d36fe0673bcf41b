// fclb_epa_launch.cuh -- launch of the warp-per-query EPA stage (own translation
// units: the EPA kernels dominate compile time).
#pragma once
#include <cstdlib>

#include "fclb_collide_impl.cuh"

namespace fclb {

// tier-1 pool: faces per query (edges = vertices = 1.51 x faces); FCLB_EPA_TIER1_FACES overrides it for tuning
inline int tier1Faces() {
  static int v = [] {
    const char* e = getenv("FCLB_EPA_TIER1_FACES");
    const int x = e ? atoi(e) : 0;
    return x >= 8 ? x : 40;
  }();
  return v;
}
// tier 1 hands a query to tier 2 after this many EPA iterations (FCLB_EPA_TIER1_ITERS; the reference's limit is 255)
inline int tier1Iters() {
  static int v = [] {
    const char* e = getenv("FCLB_EPA_TIER1_ITERS");
    const int x = e ? atoi(e) : 0;
    return x >= 1 ? x : 32;
  }();
  return v;
}
inline int epaBlocksPerSmCap() {
  static int v = [] {
    const char* e = getenv("FCLB_EPA_BLOCKS_PER_SM");
    const int x = e ? atoi(e) : 0;
    return x >= 1 ? x : 4;
  }();
  return v;
}

template <typename S, int T0, int T1, int T>
cudaError_t launchEpaTier(const BatchView& b, const CollideLaunchArgs& a, int pool_faces, const EpaDefer& defer,
                          cudaStream_t st, int grid_override = 0) {
  const size_t poly = PolyStore<S>::bytes(pool_faces, defer.enabled != 0);
  const size_t per_tile = epaTileBytes<S>(poly);
  const size_t esmem = per_tile * (kEpaThreads / T);
  auto kern = epaKernel<S, T0, T1, T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(esmem));
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = int((227 * 1024) / (esmem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > epaBlocksPerSmCap()) per_sm = epaBlocksPerSmCap();
  kern<<<grid_override > 0 ? grid_override : sms * per_sm, kEpaThreads, esmem, st>>>(b, S(a.sp.epa_tol), pool_faces, a.sp.epa_max_iter, a.mode, a.out,
                                                 a.work, defer, poly, tier1Iters());
  return cudaGetLastError();
}

// lanes per query of the first tier; FCLB_EPA_TILE overrides it for tuning (4, 8 or 16; 0 = by pair kind).
// Measured on B200: 8 lanes are best for Convex pairs (round 1, profiles/r01_epa_sweep.txt: 70 ms against 72 / 80 ms
// with 4 / 16 on c1b_convex; after the unroll-1 change 48.1 against 49.8 / 51.7 ms, f64 90 against 100 / 116 ms).
// Box pairs: round 1 picked 16 lanes (7.0 ms against 7.5 / 9.2 ms with 8 / 4 on c1b, when the few queries that run to
// the iteration limit set the run time); since tier 1 hands those to tier 2 after 32 iterations and the kernel is a
// quarter smaller, FOUR lanes win -- 6.34 ms against 6.50 / 7.03 ms with 8 / 16, f64 6.94 against 6.96 / 8.79 ms
// (profiles/r02_epa_knobs_after_unroll.txt): a box polytope rarely has more than a dozen faces.
inline int tier1TileOverride() {
  static int v = [] {
    const char* e = getenv("FCLB_EPA_TILE");
    const int x = e ? atoi(e) : 0;
    return (x == 4 || x == 8 || x == 16) ? x : 0;
  }();
  return v;
}
template <typename S, int T0, int T1>
cudaError_t launchEpaTier1(const BatchView& b, const CollideLaunchArgs& a, int pool_faces, const EpaDefer& defer,
                           cudaStream_t st) {
  const int tile = tier1TileOverride() ? tier1TileOverride() : ((T0 == ST_BOX && T1 == ST_BOX) ? 4 : 8);
  switch (tile) {
    case 4: return launchEpaTier<S, T0, T1, 4>(b, a, pool_faces, defer, st);
    case 16: return launchEpaTier<S, T0, T1, 16>(b, a, pool_faces, defer, st);
    default: return launchEpaTier<S, T0, T1, 8>(b, a, pool_faces, defer, st);
  }
}

// FCLB_EPA_EARLY_TIER2=0 turns the early consumers off (tier 2 then starts after tier 1, as in round 1)
inline int epaEarlyTier2Ctas() {
  static int v = [] {
    const char* e = getenv("FCLB_EPA_EARLY_TIER2");
    // CTAs of 4 warps beside tier 1.  Mid-round, with 16-lane box tiles, they were within noise (c1b 7.79 -> 7.70 ms) and off.
    // Since box pairs run 4-lane tiles, tier 1 is 3.1 ms of c1b's step and tier 2 -- the latency chain of the four queries
    // that run all 255 iterations -- 2.9 ms: started inside tier 1's run time, c1b 6.31 -> 5.99 ms (f64 6.94 -> 6.73 ms),
    // Convex pairs unchanged (profiles/r02_epa_knobs_after_unroll.txt).  On by default with 74 CTAs.
    return e ? atoi(e) : 74;
  }();
  return v;
}

template <typename S, int T0, int T1>
cudaError_t launchEpaT(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st) {
  EpaDefer d = a.defer;
  cudaError_t e = cudaMemsetAsync(d.count, 0, 8 * sizeof(uint32_t), st);  // count, the work cursors, the done flag
  if (e != cudaSuccess) return e;
  if (a.sp.epa_max_faces > tier1Faces()) {
    const bool early = a.aux && a.ev_aux0 && a.ev_aux1 && epaEarlyTier2Ctas() > 0 && b.count >= 4096 && a.item_capacity >= b.count;
    if (early) {
      // the early consumers: the tier-2 kernel on a small grid, launched BEFORE tier 1 on the second stream; they wait
      // for deferred items to appear in the list (kEpaItemEmpty until then) and stop when tier 1 has ended
      e = cudaMemsetAsync(d.item, 0xff, b.count * sizeof(uint32_t), st);
      if (e != cudaSuccess) return e;
      if ((e = cudaEventRecord(a.ev_aux0, st)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(a.aux, a.ev_aux0, 0)) != cudaSuccess) return e;
      EpaDefer c = d;
      c.enabled = 0;
      c.consume = 2;
      c.cursor = d.count + 3;
      e = launchEpaTier<S, T0, T1, 32>(b, a, a.sp.epa_max_faces, c, a.aux, epaEarlyTier2Ctas());
      if (e != cudaSuccess) return e;
    }
    d.enabled = 1;
    d.consume = 0;
    d.cursor = d.count + 1;
    e = launchEpaTier1<S, T0, T1>(b, a, tier1Faces(), d, st);
    if (e != cudaSuccess) return e;
    if (early && (e = cudaMemsetAsync(d.count + 4, 0xff, sizeof(uint32_t), st)) != cudaSuccess) return e;  // "tier 1 has ended"
    d.enabled = 0;
    d.consume = 1;
    d.cursor = d.count + 2;
    e = launchEpaTier<S, T0, T1, 32>(b, a, a.sp.epa_max_faces, d, st);
    if (e != cudaSuccess) return e;
    if (early) {
      if ((e = cudaEventRecord(a.ev_aux1, a.aux)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(st, a.ev_aux1, 0)) != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  d.enabled = 0;
  d.consume = 0;
  d.cursor = d.count + 1;
  return launchEpaTier1<S, T0, T1>(b, a, a.sp.epa_max_faces, d, st);
}

template <typename S>
cudaError_t launchEpa(const BatchView& b, const CollideLaunchArgs& a, cudaStream_t st) {
  if (b.type1 == ST_BOX && b.type2 == ST_BOX) return launchEpaT<S, ST_BOX, ST_BOX>(b, a, st);
  if (b.type1 == ST_CONVEX && b.type2 == ST_CONVEX) return launchEpaT<S, ST_CONVEX, ST_CONVEX>(b, a, st);
  return launchEpaT<S, ST_DYNAMIC, ST_DYNAMIC>(b, a, st);
}

}  // namespace fclb
