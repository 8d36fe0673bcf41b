// fclb_octree_impl.cuh -- batched octree-vs-shape collide, ONE WARP PER QUERY.
//
// Reference path (results contract):
//   fcl::collide(Octree2CollisionGeometry, tf_octree, Shape, tf_shape)
//     -> OcTree2ShapeCollide (collision_func_matrix-inl.h:253-273)
//     -> CollisionSolverOctree2::octreeShapeIntersectImpl (traversal/octree2/octree2_solver_traverse-inl.h:12-136):
//        shape OBB = computeBV<OBB, Shape>(shape, tf_shape) (geometry/shape/utility-inl.h:93-246, 780-787),
//        FixedRotationBoxDisjoint::initialize(tf_octree, tf_of_that_OBB) + isDisjoint(node box, +-extent, false)
//        (math/fixed_rotation_obb_disjoint-inl.h:10-45 generic, :190-357 float/SSE association),
//        child boxes by computeChildAABB (geometry/octree2/octree_util-inl.h:10-37),
//        leaf layer = 2x2x2 bitmask nodes, fully occupied inner nodes act as one box,
//        every candidate box -> ShapeIntersect<Box, Shape> (octree2_solver_leaf-inl.h:23-44).
//   node arrays: OctreeInnerNode = 8 x u32 children (octree_node.h:35-39), OctreeLeafNode = 1 byte,
//   child layer is the leaf layer when depth + 3 >= num_layers (octree-inl.h:161-163).
// The contact count does not depend on the visiting order, so the warp pops up to 32 stack
// elements per step (one per lane), and pushes children / queues candidate boxes at offsets
// from a warp prefix sum; the leaf routine runs on 32 boxes at a time (same leaf stage as the
// heightmap kernel, fclb_heightmap_impl.cuh boxShapeHit).
#pragma once
#include "fclb_bvh_shape_impl.cuh"   // fitObbPoints
#include "fclb_heightmap_impl.cuh"   // boxShapeHit

namespace fclb {

// computeBV<OBB<S>, Shape>(shape, tf, bv)
template <typename S>
FCLB_DI NodeD<S> shapeObbDirect(const ShapeInst<S>& sh, const Pose<S>& tf) {
  NodeD<S> bv;
  bv.first_child = -1;
  bv.axis = tf.R;
  bv.To = tf.t;
  switch (sh.type) {
    case ST_BOX:
      bv.extent = mk<S>(sh.p0 * S(0.5), sh.p1 * S(0.5), sh.p2 * S(0.5));
      break;
    case ST_SPHERE:
#pragma unroll
      for (int i = 0; i < 9; i++) bv.axis.m[i] = (i % 4 == 0) ? S(1) : S(0);
      bv.extent = mk<S>(sh.p0, sh.p0, sh.p0);
      break;
    case ST_ELLIPSOID:
      bv.extent = mk<S>(sh.p0, sh.p1, sh.p2);
      break;
    case ST_CAPSULE:
      bv.extent = mk<S>(sh.p0, sh.p0, sh.p1 / 2 + sh.p0);
      break;
    case ST_CONE:
    case ST_CYLINDER:
      bv.extent = mk<S>(sh.p0, sh.p0, sh.p1 / 2);
      break;
    default: {  // ST_CONVEX: fit(vertices) in the shape frame, then bv.axis = R * axis, bv.To = tf * To
      const ConvexD<S>& c = *sh.cvx;
      const NodeD<S> local = fitObbPoints<S>(c.n_verts, [&](int i) { return loadVert(c.verts, i); });
      bv.axis = mulMM(tf.R, local.axis);
      bv.To = apply(tf, local.To);
      bv.extent = local.extent;
      break;
    }
  }
  return bv;
}

template <typename S>
struct FixedRot {
  M3<S> R, A;  // rotation_2in1, |rotation_2in1| + 1e-6
  V3<S> t;
};
template <typename S>
FCLB_DI FixedRot<S> makeFixedRot(const Pose<S>& tf1, const Pose<S>& tf2) {
  const Pose<S> rel = compose(inverse(tf1), tf2);
  FixedRot<S> f;
  f.R = rel.R;
  f.t = rel.t;
#pragma unroll
  for (int i = 0; i < 9; i++) f.A.m[i] = fabs_(rel.R.m[i]) + S(1e-6);
  return f;
}
// row i of M times v with the association the reference's scalar type uses:
// double: Eigen left-to-right; float: mat3x4_mul_vec4 without SSE4 = (m0 v0 + m2 v2) + m1 v1
// (math/math_simd_details.h:208-221)
FCLB_DI float rowDotAssoc(const M3<float>& m, int i, const V3<float>& v) {
  return (m(i, 0) * v.x + m(i, 2) * v.z) + m(i, 1) * v.y;
}
FCLB_DI double rowDotAssoc(const M3<double>& m, int i, const V3<double>& v) {
  return (m(i, 0) * v.x + m(i, 1) * v.y) + m(i, 2) * v.z;
}
// isDisjoint(aabb1, aabb2 = [-e2, e2], check_strict_disjoint = false): the 6 face axes
template <typename S>
FCLB_DI bool fixedRotDisjoint6(const FixedRot<S>& f, const V3<S>& mn1, const V3<S>& mx1, const V3<S>& e2) {
  const V3<S> c1 = (mn1 + mx1) * S(0.5);
  const V3<S> a = S(0.5) * (mx1 - mn1);
  const V3<S> T = f.t - c1;  // R * 0 + t - c1
  if (fabs_(T.x) > a.x + rowDotAssoc(f.A, 0, e2)) return true;
  if (fabs_(T.y) > a.y + rowDotAssoc(f.A, 1, e2)) return true;
  if (fabs_(T.z) > a.z + rowDotAssoc(f.A, 2, e2)) return true;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const S s = dot(col(f.R, j), T);
    const S rhs = dot(a, col(f.A, j)) + comp(e2, j);
    if (fabs_(s) > rhs) return true;
  }
  return false;
}

// computeChildAABB (octree_util-inl.h:10-37)
template <typename S>
FCLB_DI void childAabb(const V3<S>& mn, const V3<S>& mx, int child, V3<S>& cmn, V3<S>& cmx) {
  const S hx = (mn.x + mx.x) * S(0.5), hy = (mn.y + mx.y) * S(0.5), hz = (mn.z + mx.z) * S(0.5);
  cmn.x = (child & 1) ? hx : mn.x;
  cmx.x = (child & 1) ? mx.x : hx;
  cmn.y = (child & 2) ? hy : mn.y;
  cmx.y = (child & 2) ? mx.y : hy;
  cmn.z = (child & 4) ? hz : mn.z;
  cmx.z = (child & 4) ? mx.z : hz;
}

template <typename S>
struct OctElem {  // OctreeTraverseStackElement
  S mn[3], mx[3];
  uint32_t index;
  uint32_t meta;  // bit 0: is_leaf_node, bits 8..: depth
};
template <typename S>
struct OctCand {  // candidate voxel box + encodeOctree2Node
  S mn[3], mx[3];
  long long code;
};

constexpr int kOctWarps = kOctreeWarps;
constexpr int kOctStack = 384;   // stack elements per warp
constexpr int kOctDfsReserve = 7 * 16;  // head room of the depth-first fallback
constexpr int kOctQueue = 160;   // queued candidate boxes per warp (< 32 left over + 16 popped leaves x 8 voxels)

template <typename S, int T1>
__global__ void __launch_bounds__(kOctWarps * 32) octreeShapeKernel(OctreeArgs a) {
  extern __shared__ __align__(16) unsigned char s_oct_raw[];
  SlotStore<S> st;
  st.base = reinterpret_cast<S*>(s_oct_raw) + threadIdx.x;
  st.stride = blockDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wbase = s_oct_raw + size_t(24) * sizeof(S) * blockDim.x +
                         size_t(warp) * (kOctStack * sizeof(OctElem<S>) + kOctQueue * sizeof(OctCand<S>));
  OctElem<S>* stack = reinterpret_cast<OctElem<S>*>(wbase);
  OctCand<S>* queue = reinterpret_cast<OctCand<S>*>(wbase + kOctStack * sizeof(OctElem<S>));
  unsigned long long st_node = 0, st_leaf = 0;

  while (true) {
    unsigned long long q64 = 0;
    if (lane == 0) q64 = atomicAdd(a.work_counter, 1ull);
    q64 = __shfl_sync(0xffffffffu, q64, 0);
    if (q64 >= a.n) break;
    const size_t q = size_t(q64);
    const uint32_t sid = a.shape_ids[q];
    const ShapeInst<S> sh = bindShape(static_cast<const ShapeD<S>*>(a.shapes), static_cast<const ConvexD<S>*>(a.convex), sid);
    const Pose<S> tf_oct = loadPose(static_cast<const S*>(a.poses_octree), q);
    const Pose<S> tf_shape = loadPose(static_cast<const S*>(a.poses_shape), q);
    const NodeD<S> obb = shapeObbDirect(sh, tf_shape);
    Pose<S> tf_obb;
    tf_obb.R = obb.axis;
    tf_obb.t = obb.To;
    const FixedRot<S> fr = makeFixedRot(tf_oct, tf_obb);
    const V3<S> e2 = obb.extent;

    uint32_t count = 0;
    long long first = -1;
    int sp = 0, nq = 0;
    bool done = (a.max_contacts == 0) || (a.n_inner == 0);
    if (!done) {
      if (lane == 0) {
        OctElem<S> r;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          r.mn[k] = S(a.root_box[k]);
          r.mx[k] = S(a.root_box[3 + k]);
        }
        r.index = 0;
        r.meta = 0;
        stack[0] = r;
      }
      sp = 1;
    }
    __syncwarp();

    auto runLeaf = [&](bool flush) {
      while (!done && (nq >= 32 || (flush && nq > 0))) {
        const int batch = nq < 32 ? nq : 32;
        bool hit = false;
        long long code = -1;
        OctCand<S> c;
        if (lane < batch) {
          c = queue[nq - 1 - lane];
          code = c.code;
          const V3<S> bmin = mk<S>(c.mn[0], c.mn[1], c.mn[2]), bmax = mk<S>(c.mx[0], c.mx[1], c.mx[2]);
          const V3<S> side = bmax - bmin;
          const V3<S> center = (bmin + bmax) * S(0.5);
          Pose<S> tf_box;
          tf_box.R = tf_oct.R;
          tf_box.t = mulMV(tf_oct.R, center) + tf_oct.t;
          st_leaf++;
          if (!a.cand.count) hit = boxShapeHit<S, T1>(side, tf_box, sh, tf_shape, S(a.tol), a.max_iter, st);
        }
        if (a.cand.count) {  // candidate mode: the leaf batch decides (ShapeIntersect<Box, Shape> with contacts)
          const S hb[6] = {c.mn[0], c.mn[1], c.mn[2], c.mx[0], c.mx[1], c.mx[2]};
          candAppend<S>(a.cand, lane < batch, uint32_t(q), code, -1, hb, nullptr);
        }
        nq -= batch;
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm) {
          if (first < 0) first = __shfl_sync(0xffffffffu, code, __ffs(hm) - 1);
          if (a.out_b1 && hit) {
            const uint32_t slot = count + uint32_t(__popc(hm & ((1u << lane) - 1u)));
            if (slot < a.max_keep && slot < a.max_contacts) {
              a.out_b1[q * a.max_keep + slot] = code;
              S* ob = static_cast<S*>(a.out_box) + (q * a.max_keep + slot) * 6;
#pragma unroll
              for (int k = 0; k < 3; k++) {
                ob[k] = c.mn[k];
                ob[3 + k] = c.mx[k];
              }
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

    while (!done && sp > 0) {
      // a popped element pushes <= 8 children and queues <= 8 boxes: bound both before popping
      // (popping `take` and pushing 8 each leaves sp + 7 take).  Wide pops stop while kOctDfsReserve slots are
      // still free; from there the warp pops one element at a time from the top, i.e. plain depth-first
      // order, whose growth is bounded by 7 per remaining tree level (<= 16 levels: half shapes are uint16).
      int take = sp < 32 ? sp : 32;
      if (sp + 7 * take > kOctStack - kOctDfsReserve) {
        const int fit = (kOctStack - kOctDfsReserve - sp) / 7;
        take = fit < 1 ? 1 : fit;
      }
      if (take * 8 > kOctQueue - nq) take = (kOctQueue - nq) / 8;
      if (take < 1) {  // queue nearly full: drain it first
        runLeaf(true);
        continue;
      }
      OctElem<S> el;
      bool have = false;
      if (lane < take) {
        el = stack[sp - 1 - lane];
        have = true;
      }
      sp -= take;
      __syncwarp();
      int n_push = 0, n_cand = 0;
      unsigned child_mask = 0;   // children to push (inner) or voxels to queue (leaf)
      bool whole = false;        // the element's own box is a candidate
      const bool is_leaf = have && (el.meta & 1u);
      if (have) {
        const bool pruned = a.pruned && !is_leaf && a.pruned[el.index];
        st_node++;
        if (!pruned && !fixedRotDisjoint6(fr, mk<S>(el.mn[0], el.mn[1], el.mn[2]), mk<S>(el.mx[0], el.mx[1], el.mx[2]), e2)) {
          if (is_leaf) {
            const unsigned bits = a.leaf_bits[el.index];
            if (bits == 0xffu) {
              whole = true;
              n_cand = 1;
            } else {
              child_mask = bits;
              n_cand = __popc(bits);
            }
          } else if (a.inner_full[el.index]) {
            whole = true;
            n_cand = 1;
          } else {
#pragma unroll
            for (int c = 0; c < 8; c++)
              if (a.inner_children[size_t(8) * el.index + c] != 0xffffffffu) child_mask |= 1u << c;
            n_push = __popc(child_mask);
          }
        }
      }
      // exclusive prefix sums over the lanes
      int push_off = n_push, cand_off = n_cand;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int pv = __shfl_up_sync(0xffffffffu, push_off, off);
        const int cv = __shfl_up_sync(0xffffffffu, cand_off, off);
        if (lane >= off) {
          push_off += pv;
          cand_off += cv;
        }
      }
      const int tot_push = __shfl_sync(0xffffffffu, push_off, 31), tot_cand = __shfl_sync(0xffffffffu, cand_off, 31);
      if (sp + tot_push > kOctStack) {  // deeper than the depth-first head room: report, never corrupt
        if (lane == 0) atomicAdd(&a.stats[2], 1ull);
        done = true;
        break;
      }
      push_off -= n_push;
      cand_off -= n_cand;
      if (have) {
        const V3<S> mn = mk<S>(el.mn[0], el.mn[1], el.mn[2]), mx = mk<S>(el.mx[0], el.mx[1], el.mx[2]);
        const uint32_t depth = el.meta >> 8;
        if (n_push) {
          const bool child_leaf = int(depth) + 3 >= a.num_layers;  // isChildLayerLeafNode(parent.depth)
          int k = 0;
          for (int c = 0; c < 8; c++) {
            if (!(child_mask & (1u << c))) continue;
            V3<S> cmn, cmx;
            childAabb(mn, mx, c, cmn, cmx);
            OctElem<S> ch;
            ch.mn[0] = cmn.x; ch.mn[1] = cmn.y; ch.mn[2] = cmn.z;
            ch.mx[0] = cmx.x; ch.mx[1] = cmx.y; ch.mx[2] = cmx.z;
            ch.index = a.inner_children[size_t(8) * el.index + c];
            ch.meta = ((depth + 1) << 8) | (child_leaf ? 1u : 0u);
            stack[sp + push_off + k] = ch;
            k++;
          }
        }
        if (n_cand) {
          // encodeOctree2Node(index, is_leaf, child) (octree2_solver_leaf-inl.h:10-20)
          const long long base = (long long)el.index + ((long long)(is_leaf ? 1 : 0) << 48);
          if (whole) {
            OctCand<S> cd;
            cd.mn[0] = mn.x; cd.mn[1] = mn.y; cd.mn[2] = mn.z;
            cd.mx[0] = mx.x; cd.mx[1] = mx.y; cd.mx[2] = mx.z;
            cd.code = base + (8ll << 32);  // inner_child_idx defaults to kInvalidChildIndex = 8 (octree2_solver.h:47-49)
            queue[nq + cand_off] = cd;
          } else {
            int k = 0;
            for (int c = 0; c < 8; c++) {
              if (!(child_mask & (1u << c))) continue;
              V3<S> cmn, cmx;
              childAabb(mn, mx, c, cmn, cmx);
              OctCand<S> cd;
              cd.mn[0] = cmn.x; cd.mn[1] = cmn.y; cd.mn[2] = cmn.z;
              cd.mx[0] = cmx.x; cd.mx[1] = cmx.y; cd.mx[2] = cmx.z;
              cd.code = base + ((long long)c << 32);
              queue[nq + cand_off + k] = cd;
              k++;
            }
          }
        }
      }
      sp += tot_push;
      nq += tot_cand;
      __syncwarp();
      runLeaf(false);
    }
    runLeaf(true);
    if (lane == 0) {
      a.counts[q] = count;
      if (a.first_node) a.first_node[q] = first;
    }
    __syncwarp();
  }
  if (a.stats) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      st_node += __shfl_xor_sync(0xffffffffu, st_node, off);
      st_leaf += __shfl_xor_sync(0xffffffffu, st_leaf, off);
    }
    if (lane == 0) {
      atomicAdd(&a.stats[0], st_node);
      atomicAdd(&a.stats[1], st_leaf);
    }
  }
}

template <typename S>
cudaError_t launchOctreeShape(int type1, const OctreeArgs& a, int grid, cudaStream_t st) {
  const size_t smem = size_t(24) * sizeof(S) * kOctWarps * 32 +
                      size_t(kOctWarps) * (kOctStack * sizeof(OctElem<S>) + kOctQueue * sizeof(OctCand<S>));
#define FCLB_OCT_LAUNCH(T)                                                                                                  \
  {                                                                                                                         \
    cudaError_t e_ = cudaFuncSetAttribute(octreeShapeKernel<S, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); \
    if (e_ != cudaSuccess) return e_;                                                                                       \
    octreeShapeKernel<S, T><<<grid, kOctWarps * 32, smem, st>>>(a);                                                         \
  }
  switch (type1) {
    case ST_BOX: FCLB_OCT_LAUNCH(ST_BOX) break;
    case ST_SPHERE: FCLB_OCT_LAUNCH(ST_SPHERE) break;
    case ST_CONVEX: FCLB_OCT_LAUNCH(ST_CONVEX) break;
    default: FCLB_OCT_LAUNCH(ST_DYNAMIC) break;
  }
#undef FCLB_OCT_LAUNCH
  return cudaGetLastError();
}

}  // namespace fclb
