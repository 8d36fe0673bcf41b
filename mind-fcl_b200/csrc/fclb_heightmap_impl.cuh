// fclb_heightmap_impl.cuh -- batched heightmap-vs-shape collide, ONE WARP PER QUERY.
//
// Reference path (results contract):
//   fcl::collide(HeightMapCollisionGeometry, tf_hm, Shape, tf_shape)
//     -> HeightMapShapeCollide (collision_func_matrix-inl.h:99-118)
//     -> HeightMapCollisionSolver::heightMapShapeIntersectImpl
//        (traversal/heightmap/heightmap_solver_traverse-inl.h:23-75):
//          shape AABB in the map frame: computeBV<AABB, Shape>(shape, tf_hm^-1 * tf_shape)
//          (geometry/shape/utility-inl.h:73-235), z reject, aabbToPixelROI
//          (geometry/heightmap/flat_heightmap-inl.h:59-92);
//     -> flatHeightMapShapeIntersectImpl (:77-118): every pixel of the ROI with a
//        non-zero height >= the AABB's min z becomes a Box (constructBox,
//        utility-inl.h:896-900) tested by ShapeIntersect<Box, Shape>
//        (heightmap_solver_leaf-inl.h:10-31 -> shape_pair_intersect-inl.h:49-80 ->
//        GJKSolver::shapeIntersect: boxBox2 / sphereBox closed forms, else MPR with
//        GJK on "Failed", gjk_solver-inl.h:71-140).
// The number of colliding pixels does not depend on the visiting order, so the
// warp scans the ROI tile by tile: a coarse layer of the LayeredHeightMap
// (max over 2^k x 2^k pixels, layered_heightmap-inl.h:77-101) rejects whole
// tiles whose maximum height is zero or below the AABB (exactly the per-pixel
// rejections of :89-95, taken for a tile at once), surviving pixels are
// compacted with ballots into a shared-memory queue and the leaf routine runs
// on 32 pixel boxes at a time.
// For an ROI of at least 1/8 of the map the reference switches to a traversal of
// the layer pyramid (:119-185) whose extra culling test (6-axis OBB SAT) is
// conservative; both variants report the same pixels except for boxes within
// rounding of that SAT's boundary (declared).
#pragma once
#include "fclb_boxbox.cuh"
#include "fclb_gjk.cuh"
#include "fclb_internal.h"
#include "fclb_leafcand.cuh"
#include "fclb_mpr.cuh"
#include "fclb_primitives_intersect.cuh"

// (3 CTAs per SM were measured: 80 registers with spills, 21.5 ms against 19.1 ms on C4 -- keep 2)
#ifndef FCLB_HM_MIN_BLOCKS
#define FCLB_HM_MIN_BLOCKS 2
#endif

namespace fclb {

// fabs() in the reference's computeBV resolves to the C double overload for
// S = float (no using-declaration in scope), so the ranges are formed in double
// from S-rounded products and rounded once to S.
template <typename S>
FCLB_DI double fabsd(S v) {
  return fabs(double(v));
}

// computeBV<AABB<S>, Shape>(shape, tf, bv)  (geometry/shape/utility-inl.h:73-235, 770-778)
template <typename S>
FCLB_DI void shapeAabb(const ShapeInst<S>& sh, const Pose<S>& tf, V3<S>& mn, V3<S>& mx) {
  const M3<S>& R = tf.R;
  const V3<S>& T = tf.t;
  if (sh.type == ST_CONVEX) {
    const S big = sizeof(S) == 4 ? S(3.402823466e+38f) : S(1.7976931348623157e+308);
    mn = mk<S>(big, big, big);
    mx = mk<S>(-big, -big, -big);
    const ConvexD<S>& c = *sh.cvx;
    #pragma unroll 1
    for (int i = 0; i < c.n_verts; i++) {
      const V3<S> p = mulMV(R, loadVert(c.verts, i)) + T;
      mn = mk<S>(fmin_(mn.x, p.x), fmin_(mn.y, p.y), fmin_(mn.z, p.z));
      mx = mk<S>(fmax_(mx.x, p.x), fmax_(mx.y, p.y), fmax_(mx.z, p.z));
    }
    return;
  }
  V3<S> d;
  switch (sh.type) {
    case ST_BOX:
      d.x = S(0.5 * (fabsd(R(0, 0) * sh.p0) + fabsd(R(0, 1) * sh.p1) + fabsd(R(0, 2) * sh.p2)));
      d.y = S(0.5 * (fabsd(R(1, 0) * sh.p0) + fabsd(R(1, 1) * sh.p1) + fabsd(R(1, 2) * sh.p2)));
      d.z = S(0.5 * (fabsd(R(2, 0) * sh.p0) + fabsd(R(2, 1) * sh.p1) + fabsd(R(2, 2) * sh.p2)));
      break;
    case ST_SPHERE:
      d = mk<S>(sh.p0, sh.p0, sh.p0);
      break;
    case ST_ELLIPSOID:
      d.x = S(fabsd(R(0, 0) * sh.p0) + fabsd(R(0, 1) * sh.p1) + fabsd(R(0, 2) * sh.p2));
      d.y = S(fabsd(R(1, 0) * sh.p0) + fabsd(R(1, 1) * sh.p1) + fabsd(R(1, 2) * sh.p2));
      d.z = S(fabsd(R(2, 0) * sh.p0) + fabsd(R(2, 1) * sh.p1) + fabsd(R(2, 2) * sh.p2));
      break;
    case ST_CAPSULE:  // p0 = radius, p1 = lz
      d.x = S(0.5 * fabsd(R(0, 2) * sh.p1) + sh.p0);
      d.y = S(0.5 * fabsd(R(1, 2) * sh.p1) + sh.p0);
      d.z = S(0.5 * fabsd(R(2, 2) * sh.p1) + sh.p0);
      break;
    default:  // ST_CONE, ST_CYLINDER
      d.x = S(fabsd(R(0, 0) * sh.p0) + fabsd(R(0, 1) * sh.p0) + 0.5 * fabsd(R(0, 2) * sh.p1));
      d.y = S(fabsd(R(1, 0) * sh.p0) + fabsd(R(1, 1) * sh.p0) + 0.5 * fabsd(R(1, 2) * sh.p1));
      d.z = S(fabsd(R(2, 0) * sh.p0) + fabsd(R(2, 1) * sh.p0) + 0.5 * fabsd(R(2, 2) * sh.p1));
      break;
  }
  mx = T + d;
  mn = T - d;
}

// GJKSolver::shapeIntersect(Box, tf_box, Shape, tf_shape, nullptr)  (gjk_solver-inl.h:71-140,152-243)
template <typename S, int T1>
FCLB_DI bool boxShapeHit(const V3<S>& side, const Pose<S>& tf_box, const ShapeInst<S>& sh, const Pose<S>& tf_shape, S tol,
                         int max_iter, SlotStore<S>& st) {
  const int type = (T1 == ST_DYNAMIC) ? sh.type : T1;
  if (type == ST_SPHERE) {
    ContactPt<S> cp;
    return sphereBoxIntersect(sh.p0, tf_shape, side, tf_box, false, cp);
  } else if (type == ST_BOX) {
    // the boolean answer is boxBox2's return code, final after the 15-axis test: the contact generation that follows it
    // in the reference (box_box-inl.h:437-807) cannot change it and is left out of the traversal kernels
    int n = 0;
    return boxBox2<S, true>(side, tf_box, mk<S>(sh.p0, sh.p1, sh.p2), tf_shape, nullptr, &n) != 0;
  } else {
    MinkDiff<S, ST_BOX, T1> md;
    md.s0.type = ST_BOX;
    md.s0.p0 = side.x;
    md.s0.p1 = side.y;
    md.s0.p2 = side.z;
    md.s0.cvx = nullptr;
    md.s1 = sh;
    md.setPoses(tf_box, tf_shape);
    const int ms = mprIntersect<S>(md, max_iter, tol, nullptr);
    if (ms == MPR_INTERSECT) return true;
    if (ms == MPR_SEPARATED) return false;
    Simp simplex;
    simplex.ord = 0;
    simplex.rank = -1;
    return gjkEvaluate<S>(md, st, simplex, mk<S>(S(-1), S(0), S(0)), tol, max_iter, nullptr, nullptr) == GJK_INTERSECT;
  }
}

constexpr int kHmWarps = kHeightmapWarps;
constexpr int kHmQueue = 64;  // queued candidate pixels per warp

template <typename S, int T1>
__global__ void __launch_bounds__(kHmWarps * 32, FCLB_HM_MIN_BLOCKS) heightmapShapeKernel(HeightmapArgs a) {
  extern __shared__ __align__(16) unsigned char s_hm_raw[];
  // layout: SlotStore (24 S per thread) | per-warp pixel queues
  SlotStore<S> st;
  st.base = reinterpret_cast<S*>(s_hm_raw) + threadIdx.x;
  st.stride = blockDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* queue = reinterpret_cast<uint32_t*>(s_hm_raw + size_t(24) * sizeof(S) * blockDim.x) + warp * kHmQueue;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned long long st_pix = 0, st_leaf = 0;
  const S res_x = S(a.res_x), res_y = S(a.res_y);
  const S half_res_x = S(0.5) * res_x, half_res_y = S(0.5) * res_y;
  const S upper_m = S(a.upper_mm) * S(0.001);
  const int tile = 1 << a.coarse_shift;

  while (true) {
    unsigned long long q64 = 0;
    if (lane == 0) q64 = atomicAdd(a.work_counter, 1ull);
    q64 = __shfl_sync(0xffffffffu, q64, 0);
    if (q64 >= a.n) break;
    const size_t q = size_t(q64);
    const uint32_t sid = a.shape_ids[q];
    const ShapeInst<S> sh = bindShape(static_cast<const ShapeD<S>*>(a.shapes), static_cast<const ConvexD<S>*>(a.convex), sid);
    const Pose<S> tf_hm = loadPose(static_cast<const S*>(a.poses_hm), q);
    const Pose<S> tf_shape = loadPose(static_cast<const S*>(a.poses_shape), q);
    const Pose<S> tf_s2m = compose(inverse(tf_hm), tf_shape);
    V3<S> mn, mx;
    shapeAabb(sh, tf_s2m, mn, mx);
    uint32_t count = 0;
    int first = -1;
    bool live = a.max_contacts != 0;
    if (mx.z < 0) live = false;
    if (mn.z > upper_m) live = false;
    int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
    if (live) {
      // aabbToPixelROI (flat_heightmap-inl.h:59-92)
      const int fx = int(a.full_x), fy = int(a.full_y);
      const int px0 = int(floor(double(mn.x / res_x)) + double(a.half_x));
      const int py0 = int(floor(double(mn.y / res_y)) + double(a.half_y));
      const int px1 = int(floor(double(mx.x / res_x)) + double(a.half_x));
      const int py1 = int(floor(double(mx.y / res_y)) + double(a.half_y));
      auto clampi = [](int v, int ub) { return v <= 0 ? 0 : (v >= ub ? ub : v); };
      x0 = clampi(px0, fx - 1);
      y0 = clampi(py0, fy - 1);
      x1 = clampi(px1, fx - 1);
      y1 = clampi(py1, fy - 1);
      const bool tl_proj = (x0 != px0) || (y0 != py0);
      const bool br_proj = (x1 != px1) || (y1 != py1);
      if (tl_proj && br_proj && ((x0 == x1) || (y0 == y1))) live = false;
    }
    int nq = 0;
    bool done = false;
    // ---- leaf stage: one pixel box per lane, batches of 32 (flush: also a partial batch)
    auto runLeaf = [&](bool flush) {
      while (!done && (nq >= 32 || (flush && nq > 0))) {
        const int batch = nq < 32 ? nq : 32;
        bool hit = false;
        uint32_t pc = 0;
        V3<S> hb_min = zero3<S>(), hb_max = zero3<S>();
        if (lane < batch) {
          pc = queue[nq - 1 - lane];
          const int x = int(pc >> 16), y = int(pc & 0xffffu);
          const uint16_t h = a.bottom[size_t(y) * a.full_x + x];
          const S height_m = S(h) * S(0.001);
          // pixelToPoint2DUnchecked(Center): formed in double, rounded to S (flat_heightmap-inl.h:110-113)
          const S ccx = S((x - int(a.half_x) + 0.5) * res_x);
          const S ccy = S((y - int(a.half_y) + 0.5) * res_y);
          const V3<S> bmin = mk<S>(ccx - half_res_x, ccy - half_res_y, S(0.0));
          const V3<S> bmax = mk<S>(ccx + half_res_x, ccy + half_res_y, height_m);
          const V3<S> side = bmax - bmin;
          const V3<S> center = (bmin + bmax) * S(0.5);
          Pose<S> tf_box;
          tf_box.R = tf_hm.R;
          tf_box.t = mulMV(tf_hm.R, center) + tf_hm.t;
          st_leaf++;
          if (!a.cand.count) hit = boxShapeHit<S, T1>(side, tf_box, sh, tf_shape, S(a.tol), a.max_iter, st);
          hb_min = bmin;
          hb_max = bmax;
        }
        if (a.cand.count) {  // candidate mode: the leaf batch decides (ShapeIntersect<Box, Shape> with contacts)
          const S hb[6] = {hb_min.x, hb_min.y, hb_min.z, hb_max.x, hb_max.y, hb_max.z};
          candAppend<S>(a.cand, lane < batch, uint32_t(q), (long long)pc, -1, hb, nullptr);
        }
        nq -= batch;
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm) {
          if (first < 0) first = int(__shfl_sync(0xffffffffu, pc, __ffs(hm) - 1));
          if (a.out_b1 && hit) {
            const uint32_t slot = count + uint32_t(__popc(hm & lt_mask));
            if (slot < a.max_keep && slot < a.max_contacts) {
              a.out_b1[q * a.max_keep + slot] = (long long)pc;
              S* ob = static_cast<S*>(a.out_box) + (q * a.max_keep + slot) * 6;
              ob[0] = hb_min.x; ob[1] = hb_min.y; ob[2] = hb_min.z;
              ob[3] = hb_max.x; ob[4] = hb_max.y; ob[5] = hb_max.z;
            }
          }
          count += uint32_t(__popc(hm));
          if (count >= a.max_contacts) {
            count = a.max_contacts;
            done = true;
          }
        }
        __syncwarp();
      }
    };
    if (live) {
      // rectifyHeightMapROI (flat_heightmap-inl.h:362-378) is the identity on a clamped ROI
      const int tx0 = x0 >> a.coarse_shift, tx1 = x1 >> a.coarse_shift;
      const int ty0 = y0 >> a.coarse_shift, ty1 = y1 >> a.coarse_shift;
      const int ntx = tx1 - tx0 + 1, nty = ty1 - ty0 + 1;
      const int n_tiles = ntx * nty;
      #pragma unroll 1
      for (int tb = 0; tb < n_tiles && !done; tb += 32) {
        // ---- tile stage: one coarse pixel per lane
        const int ti = tb + lane;
        bool active = false;
        int tx = 0, ty = 0;
        if (ti < n_tiles) {
          tx = tx0 + ti % ntx;
          ty = ty0 + ti / ntx;
          const uint16_t hmax = a.coarse[size_t(ty) * a.coarse_full_x + tx];
          active = (hmax != 0) && !(mn.z > S(hmax) * S(0.001));
        }
        unsigned am = __ballot_sync(0xffffffffu, active);
        while (am && !done) {
          const int src = __ffs(am) - 1;
          am &= am - 1;
          const int cx = __shfl_sync(0xffffffffu, tx, src), cy = __shfl_sync(0xffffffffu, ty, src);
          // pixels of this tile inside the ROI
          const int bx0 = max(cx << a.coarse_shift, x0), bx1 = min((cx << a.coarse_shift) + tile - 1, x1);
          const int by0 = max(cy << a.coarse_shift, y0), by1 = min((cy << a.coarse_shift) + tile - 1, y1);
          const int w = bx1 - bx0 + 1, npx = w * (by1 - by0 + 1);
          #pragma unroll 1
          for (int pb = 0; pb < npx && !done; pb += 32) {
            const int pi = pb + lane;
            bool cand = false;
            uint32_t code = 0;
            if (pi < npx) {
              const int x = bx0 + pi % w, y = by0 + pi / w;
              const uint16_t h = a.bottom[size_t(y) * a.full_x + x];
              st_pix++;
              if (h != 0 && !(mn.z > S(h) * S(0.001))) {
                cand = true;
                code = (uint32_t(x) << 16) | uint32_t(y);
              }
            }
            const unsigned cm = __ballot_sync(0xffffffffu, cand);
            if (cand) queue[nq + __popc(cm & lt_mask)] = code;
            nq += __popc(cm);
            __syncwarp();
            runLeaf(false);
          }
        }
      }
      runLeaf(true);
    }
    if (lane == 0) {
      a.counts[q] = count;
      if (a.first_pixel) a.first_pixel[q] = first;
    }
    __syncwarp();
  }
  if (a.stats) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      st_pix += __shfl_xor_sync(0xffffffffu, st_pix, off);
      st_leaf += __shfl_xor_sync(0xffffffffu, st_leaf, off);
    }
    if (lane == 0) {
      atomicAdd(&a.stats[0], st_pix);
      atomicAdd(&a.stats[1], st_leaf);
    }
  }
}

template <typename S>
cudaError_t launchHeightmapShape(int type1, const HeightmapArgs& a, int grid, cudaStream_t st) {
  const size_t smem = size_t(24) * sizeof(S) * kHmWarps * 32 + size_t(kHmWarps) * kHmQueue * sizeof(uint32_t);
#define FCLB_HM_CASE(T)                                                                                              \
  case T: {                                                                                                          \
    cudaError_t e_ = cudaFuncSetAttribute(heightmapShapeKernel<S, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); \
    if (e_ != cudaSuccess) return e_;                                                                                \
    heightmapShapeKernel<S, T><<<grid, kHmWarps * 32, smem, st>>>(a);                                                \
    break;                                                                                                           \
  }
  switch (type1) {
    FCLB_HM_CASE(ST_BOX)
    FCLB_HM_CASE(ST_SPHERE)
    FCLB_HM_CASE(ST_ELLIPSOID)
    FCLB_HM_CASE(ST_CAPSULE)
    FCLB_HM_CASE(ST_CONE)
    FCLB_HM_CASE(ST_CYLINDER)
    FCLB_HM_CASE(ST_CONVEX)
    default: {
      cudaError_t e_ = cudaFuncSetAttribute(heightmapShapeKernel<S, ST_DYNAMIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e_ != cudaSuccess) return e_;
      heightmapShapeKernel<S, ST_DYNAMIC><<<grid, kHmWarps * 32, smem, st>>>(a);
      break;
    }
  }
#undef FCLB_HM_CASE
  return cudaGetLastError();
}

}  // namespace fclb
